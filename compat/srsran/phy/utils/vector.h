/* srsran/phy/utils/vector.h (compat): the few vector helpers the reference's PHY-facing sources call */
#ifndef SRSRAN_VECTOR_H
#define SRSRAN_VECTOR_H
#include "srsran/config.h"
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_MAX_VEC(a, b) ((a) > (b) ? (a) : (b))
#define srsran_convert_amplitude_to_dB(V) (20.0f * log10f(V))
#define srsran_convert_power_to_dB(V) (10.0f * log10f(V))
#define srsran_convert_dB_to_amplitude(V) (powf(10.0f, (V) / 20.0f))
#define srsran_convert_dB_to_power(V) (powf(10.0f, (V) / 10.0f))
SRSRAN_API void* srsran_vec_malloc(uint32_t size);
SRSRAN_API cf_t*  srsran_vec_cf_malloc(uint32_t nsamples);
SRSRAN_API float* srsran_vec_f_malloc(uint32_t nsamples);
SRSRAN_API void  srsran_vec_cf_zero(cf_t* ptr, uint32_t nsamples);
SRSRAN_API void  srsran_vec_f_zero(float* ptr, uint32_t nsamples);
SRSRAN_API void  srsran_vec_cf_copy(cf_t* dst, const cf_t* src, uint32_t len);
SRSRAN_API float srsran_vec_avg_power_cf(const cf_t* x, const uint32_t len);
SRSRAN_API float srsran_vec_acc_ff(const float* x, const uint32_t len);
SRSRAN_API void  srsran_vec_fprint_f(FILE* stream, const float* x, const uint32_t len);
SRSRAN_API void  srsran_vec_fprint_b(FILE* stream, const uint8_t* x, const uint32_t len);
SRSRAN_API void  srsran_vec_fprint_hex(FILE* stream, uint8_t* x, const uint32_t len);
SRSRAN_API void  srsran_vec_sprint_hex(char* str, const uint32_t max_str_len, uint8_t* x, const uint32_t len);
#ifdef __cplusplus
}
#endif
#endif
