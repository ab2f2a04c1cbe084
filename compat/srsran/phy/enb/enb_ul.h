/* srsran/phy/enb/enb_ul.h (compat): the uplink receiver object the reference decodes PUSCH with
 * (src/src/SubframeWorker.cc:79,261; src/src/UL_Sniffer_PUSCH.cc:391-392: in_buffer = antenna buffer 1, srsran_enb_ul_fft) */
#ifndef SRSRAN_ENB_UL_H
#define SRSRAN_ENB_UL_H
#include "srsran/phy/ch_estimation/chest_ul.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/dft/ofdm.h"
#include "srsran/phy/phch/pusch.h"
#include "srsran/phy/phch/pusch_cfg.h"
#include "srsran/phy/phch/ra.h"
#include "srsran/phy/utils/debug.h"
#include "srsran/phy/utils/vector.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  srsran_cell_t         cell;
  cf_t*                 sf_symbols;
  cf_t*                 in_buffer;
  srsran_chest_ul_res_t chest_res;
  srsran_ofdm_t         fft;
  srsran_chest_ul_t     chest;
  srsran_pusch_t        pusch;
  void*                 b200; /* compat: CUDA PHY handle behind this object */
} srsran_enb_ul_t;
typedef struct SRSRAN_API { uint32_t dummy; } srsran_pucch_cfg_t;
typedef struct SRSRAN_API {
  srsran_pucch_cfg_t                pucch;
  srsran_pusch_cfg_t                pusch;
  srsran_pusch_hopping_cfg_t        hopping;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_refsignal_srs_cfg_t        srs;
} srsran_ul_cfg_t;
SRSRAN_API int  srsran_enb_ul_init(srsran_enb_ul_t* q, cf_t* in_buffer, uint32_t max_prb);
SRSRAN_API void srsran_enb_ul_free(srsran_enb_ul_t* q);
SRSRAN_API int  srsran_enb_ul_set_cell(srsran_enb_ul_t* q, srsran_cell_t cell, srsran_refsignal_dmrs_pusch_cfg_t* pusch_cfg, srsran_refsignal_srs_cfg_t* srs_cfg);
SRSRAN_API void srsran_enb_ul_fft(srsran_enb_ul_t* q);
#ifdef __cplusplus
}
#endif
#endif
