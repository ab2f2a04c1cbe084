/* srsran/phy/utils/bit.h (compat): bit packing helpers used by falcon_dci.c / falcon_pdcch.c */
#ifndef SRSRAN_BIT_H
#define SRSRAN_BIT_H
#include "srsran/config.h"
#ifdef __cplusplus
extern "C" {
#endif
SRSRAN_API void     srsran_bit_unpack(uint32_t value, uint8_t** bits, int nof_bits);
SRSRAN_API void     srsran_bit_unpack_l(uint64_t value, uint8_t** bits, int nof_bits);
SRSRAN_API uint32_t srsran_bit_pack(uint8_t** bits, int nof_bits);
SRSRAN_API uint64_t srsran_bit_pack_l(uint8_t** bits, int nof_bits);
SRSRAN_API void     srsran_bit_pack_vector(uint8_t* unpacked, uint8_t* packed, int nof_bits);
SRSRAN_API void     srsran_bit_unpack_vector(const uint8_t* packed, uint8_t* unpacked, int nof_bits);
SRSRAN_API void     srsran_bit_fprint(FILE* stream, uint8_t* bits, int nof_bits);
SRSRAN_API uint32_t srsran_bit_diff(const uint8_t* x, const uint8_t* y, int nbits);
SRSRAN_API uint32_t srsran_bit_count(uint32_t n);
#ifdef __cplusplus
}
#endif
#endif
