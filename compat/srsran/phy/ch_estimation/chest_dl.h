/* srsran/phy/ch_estimation/chest_dl.h (compat): CRS channel-estimator configuration and result (srsran_chest_dl_res_t is what
 * the reference reads after srsran_ue_dl_decode_fft_estimate: src/src/DCISearch.cc:568-569, src/src/SubframeWorker.cc:203) */
#ifndef SRSRAN_CHEST_DL_H
#define SRSRAN_CHEST_DL_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  cf_t*    ce[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS];
  uint32_t nof_re;
  float    noise_estimate;
  float    noise_estimate_dbm;
  float    snr_db;
  float    snr_ant_port_db[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS];
  float    rsrp;
  float    rsrp_dbm;
  float    rsrp_neigh;
  float    rsrp_port_dbm[SRSRAN_MAX_PORTS];
  float    rsrp_ant_port_dbm[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS];
  float    rsrq;
  float    rsrq_db;
  float    rsrq_ant_port_db[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS];
  float    rssi_dbm;
  float    cfo;
  float    sync_error;
} srsran_chest_dl_res_t;
typedef enum SRSRAN_API { SRSRAN_NOISE_ALG_REFS = 0, SRSRAN_NOISE_ALG_PSS, SRSRAN_NOISE_ALG_EMPTY } srsran_chest_dl_noise_alg_t;
typedef enum SRSRAN_API { SRSRAN_CHEST_FILTER_GAUSS = 0, SRSRAN_CHEST_FILTER_TRIANGLE, SRSRAN_CHEST_FILTER_NONE } srsran_chest_filter_t;
typedef enum SRSRAN_API { SRSRAN_ESTIMATOR_ALG_AVERAGE = 0, SRSRAN_ESTIMATOR_ALG_INTERPOLATE, SRSRAN_ESTIMATOR_ALG_WIENER } srsran_chest_dl_estimator_alg_t;
typedef struct SRSRAN_API {
  srsran_chest_dl_estimator_alg_t estimator_alg;
  srsran_chest_dl_noise_alg_t     noise_alg;
  srsran_chest_filter_t           filter_type;
  float                           filter_coef[2];
  uint16_t                        cfo_estimate_sf_mask;
  bool                            cfo_estimate_enable;
  bool                            rsrp_neighbour;
  bool                            sync_error_enable;
} srsran_chest_dl_cfg_t;
typedef struct SRSRAN_API { srsran_cell_t cell; uint32_t nof_rx_antennas; void* b200; } srsran_chest_dl_t;
SRSRAN_API int  srsran_chest_dl_res_init(srsran_chest_dl_res_t* q, uint32_t max_prb);
SRSRAN_API void srsran_chest_dl_res_free(srsran_chest_dl_res_t* q);
SRSRAN_API srsran_chest_dl_estimator_alg_t srsran_chest_dl_str2estimator_alg(const char* str);
#ifdef __cplusplus
}
#endif
#endif
