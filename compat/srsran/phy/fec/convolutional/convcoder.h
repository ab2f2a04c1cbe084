/* srsran/phy/fec/convolutional/convcoder.h (compat) */
#ifndef SRSRAN_CONVCODER_H
#define SRSRAN_CONVCODER_H
#include "srsran/config.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API { uint32_t R; uint32_t K; int poly[3]; bool tail_biting; } srsran_convcoder_t;
SRSRAN_API int srsran_convcoder_encode(srsran_convcoder_t* q, uint8_t* input, uint8_t* output, uint32_t frame_length);
#ifdef __cplusplus
}
#endif
#endif
