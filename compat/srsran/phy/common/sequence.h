/* srsran/phy/common/sequence.h (compat): Gold sequence object (36.211 7.2) */
#ifndef SRSRAN_SEQUENCE_H
#define SRSRAN_SEQUENCE_H
#include "srsran/config.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  uint8_t* c;
  uint8_t* c_bytes;
  float*   c_float;
  short*   c_short;
  int8_t*  c_char;
  uint32_t cur_len;
  uint32_t max_len;
} srsran_sequence_t;
SRSRAN_API int  srsran_sequence_LTE_pr(srsran_sequence_t* q, uint32_t len, uint32_t seed);
#ifdef __cplusplus
}
#endif
#endif
