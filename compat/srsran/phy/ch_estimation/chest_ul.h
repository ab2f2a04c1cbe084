/* srsran/phy/ch_estimation/chest_ul.h (compat): PUSCH DMRS channel estimator (srsran_chest_ul_estimate_pusch,
 * src/src/UL_Sniffer_PUSCH.cc:256); the estimate itself is part of the GPU PUSCH kernel (k_pusch.cu) */
#ifndef SRSRAN_CHEST_UL_H
#define SRSRAN_CHEST_UL_H
#include "srsran/phy/ch_estimation/refsignal_ul.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/phch/pusch_cfg.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  cf_t*    ce;
  uint32_t nof_re;
  float    noise_estimate;
  float    noise_estimate_dbm;
  float    rsrp;
  float    rsrp_dBfs;
  float    epre;
  float    epre_dBfs;
  float    snr;
  float    snr_db;
  float    cfo_hz;
  float    ta_us;
} srsran_chest_ul_res_t;
typedef struct { srsran_cell_t cell; srsran_refsignal_ul_t dmrs_signal; void* b200; } srsran_chest_ul_t;
SRSRAN_API int  srsran_chest_ul_res_init(srsran_chest_ul_res_t* q, uint32_t max_prb);
SRSRAN_API void srsran_chest_ul_res_free(srsran_chest_ul_res_t* q);
SRSRAN_API int  srsran_chest_ul_estimate_pusch(srsran_chest_ul_t* q, srsran_ul_sf_cfg_t* sf, srsran_pusch_cfg_t* cfg, cf_t* input, srsran_chest_ul_res_t* res);
#ifdef __cplusplus
}
#endif
#endif
