/* srsran/phy/fec/softbuffer.h (compat): HARQ soft buffers.  In libltephy_b200 the soft bits of a HARQ process live in
 * device memory (ltephy_harq_*); this object is the host-side handle the reference allocates per (RNTI, pid, TB)
 * (src/src/HARQ.cc:71-135, src/src/DL_Sniffer_PDSCH.cc:43,62,250) and passes through srsran_pdsch_cfg_t.softbuffers. */
#ifndef SRSRAN_SOFTBUFFER_H
#define SRSRAN_SOFTBUFFER_H
#include "srsran/config.h"
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  uint32_t  max_cb;
  uint32_t  max_cb_size;
  int16_t** buffer_f;
  uint8_t** data;
  bool*     cb_crc;
  bool      tb_crc;
  /* compat: identity of the device-side buffer (0 = none yet) and the transport-block size it was reset for */
  uint64_t  b200_id;
  uint32_t  b200_tbs;
} srsran_softbuffer_rx_t;
typedef struct SRSRAN_API {
  uint32_t  max_cb;
  uint32_t  max_cb_size;
  uint8_t** buffer_b;
} srsran_softbuffer_tx_t;
#define SOFTBUFFER_SIZE 18600
SRSRAN_API int  srsran_softbuffer_rx_init(srsran_softbuffer_rx_t* q, uint32_t nof_prb);
SRSRAN_API int  srsran_softbuffer_rx_init_guru(srsran_softbuffer_rx_t* q, uint32_t max_cb, uint32_t max_cb_size);
SRSRAN_API void srsran_softbuffer_rx_reset(srsran_softbuffer_rx_t* p);
SRSRAN_API void srsran_softbuffer_rx_reset_cb_crc(srsran_softbuffer_rx_t* q, uint32_t nof_cb);
SRSRAN_API void srsran_softbuffer_rx_reset_tbs(srsran_softbuffer_rx_t* q, uint32_t tbs);
SRSRAN_API void srsran_softbuffer_rx_reset_cb(srsran_softbuffer_rx_t* q, uint32_t nof_cb);
SRSRAN_API void srsran_softbuffer_rx_free(srsran_softbuffer_rx_t* p);
SRSRAN_API int  srsran_softbuffer_tx_init(srsran_softbuffer_tx_t* q, uint32_t nof_prb);
SRSRAN_API void srsran_softbuffer_tx_reset(srsran_softbuffer_tx_t* p);
SRSRAN_API void srsran_softbuffer_tx_free(srsran_softbuffer_tx_t* p);
#ifdef __cplusplus
}
#endif
#endif
