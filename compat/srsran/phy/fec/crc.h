/* srsran/phy/fec/crc.h (compat) */
#ifndef SRSRAN_CRC_H
#define SRSRAN_CRC_H
#include "srsran/config.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  uint64_t table[256];
  int      polynom;
  int      order;
  uint64_t crcinit;
  uint64_t crcmask;
  uint64_t crchighbit;
  uint32_t srsran_crc_out;
} srsran_crc_t;
SRSRAN_API int      srsran_crc_init(srsran_crc_t* h, uint32_t srsran_crc_poly, int srsran_crc_order);
SRSRAN_API uint32_t srsran_crc_attach(srsran_crc_t* h, uint8_t* data, int len);
SRSRAN_API uint32_t srsran_crc_checksum(srsran_crc_t* h, uint8_t* data, int len);
#ifdef __cplusplus
}
#endif
#endif
