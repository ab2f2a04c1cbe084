/* srsran/phy/phch/pusch.h (compat): srsran_pusch_decode (src/src/UL_Sniffer_PUSCH.cc:262) */
#ifndef SRSRAN_PUSCH_H
#define SRSRAN_PUSCH_H
#include "srsran/phy/ch_estimation/chest_ul.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/phch/pusch_cfg.h"
#include "srsran/phy/phch/uci_cfg.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API { srsran_cell_t cell; bool is_ue; uint16_t ue_rnti; uint32_t max_re; bool llr_is_8bit; void* b200; } srsran_pusch_t;
typedef struct SRSRAN_API {
  uint8_t*           data;
  srsran_uci_value_t uci;
  bool               crc;
  float              avg_iterations_block;
  float              evm;
  float              epre_dbfs;
} srsran_pusch_res_t;
SRSRAN_API int srsran_pusch_decode(srsran_pusch_t* q, srsran_ul_sf_cfg_t* sf, srsran_pusch_cfg_t* cfg, srsran_chest_ul_res_t* channel, cf_t* sf_symbols,
                                   srsran_pusch_res_t* data);
#ifdef __cplusplus
}
#endif
#endif
