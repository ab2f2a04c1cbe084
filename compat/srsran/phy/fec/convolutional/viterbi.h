/* srsran/phy/fec/convolutional/viterbi.h (compat): the K=7 r=1/3 tail-biting decoder object.  libltephy_b200 decodes every PDCCH
 * candidate of a subframe on the GPU (k_viterbi.cu); this type only has to exist inside srsran_pdcch_t. */
#ifndef SRSRAN_VITERBI_H
#define SRSRAN_VITERBI_H
#include "srsran/config.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { SRSRAN_VITERBI_27 = 0, SRSRAN_VITERBI_29, SRSRAN_VITERBI_37, SRSRAN_VITERBI_39 } srsran_viterbi_type_t;
typedef struct SRSRAN_API {
  void*    ptr;
  uint32_t R;
  uint32_t K;
  uint32_t framebits;
  bool     tail_biting;
  float    gain_quant;
  int16_t  gain_quant_s;
  uint8_t* tmp;
  uint8_t* symbols_uc;
  uint16_t* symbols_us;
} srsran_viterbi_t;
SRSRAN_API int  srsran_viterbi_decode_f(srsran_viterbi_t* q, float* symbols, uint8_t* data, uint32_t frame_length);
#ifdef __cplusplus
}
#endif
#endif
