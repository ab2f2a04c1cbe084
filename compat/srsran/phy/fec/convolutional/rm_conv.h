/* srsran/phy/fec/convolutional/rm_conv.h (compat): convolutional rate matching (36.212 5.1.4.2) */
#ifndef SRSRAN_RM_CONV_H
#define SRSRAN_RM_CONV_H
#include "srsran/config.h"
#ifdef __cplusplus
extern "C" {
#endif
SRSRAN_API int srsran_rm_conv_tx(uint8_t* input, uint32_t in_len, uint8_t* output, uint32_t out_len);
SRSRAN_API int srsran_rm_conv_rx(float* input, uint32_t in_len, float* output, uint32_t out_len);
SRSRAN_API int srsran_rm_conv_rx_s(int16_t* input, uint32_t in_len, int16_t* output, uint32_t out_len);
#ifdef __cplusplus
}
#endif
#endif
