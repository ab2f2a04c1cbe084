/* srsran/phy/ue/ue_dl.h (compat): the per-worker downlink PHY context the reference hands to every stage of
 * SubframeWorker::work (src/src/SubframeWorker.cc:142-207) -- srsran_ue_dl_t.  One object = one CUDA PHY handle of
 * libltephy_b200 with a batch of one subframe:
 *   srsran_ue_dl_decode_fft_estimate (src/src/DCISearch.cc:562)  -> phase A on the GPU for this subframe: OFDM rx, CRS channel
 *        estimate, PCFICH, PDCCH LLRs AND the whole blind-decode table; q->sf_symbols, q->chest_res, q->pdcch.llr are filled
 *   srsran_pdcch_dci_decode (falcon_pdcch.c:142)                 -> a lookup in that table
 *   srsran_ue_dl_decode_pdsch (src/src/DL_Sniffer_PDSCH.cc:997)  -> phase B for one grant
 * Throughput-oriented callers use the batched tier-1 C-ABI (include/ltephy_b200.h, ltephy_search.h, ltephy_shard.h). */
#ifndef SRSRAN_UE_DL_H
#define SRSRAN_UE_DL_H
#include "srsran/phy/ch_estimation/chest_dl.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/dft/ofdm.h"
#include "srsran/phy/phch/cqi.h"
#include "srsran/phy/phch/dci.h"
#include "srsran/phy/phch/pdcch.h"
#include "srsran/phy/phch/pdsch.h"
#include "srsran/phy/phch/pdsch_cfg.h"
#include "srsran/phy/phch/phich.h"
#include "srsran/phy/phch/ra.h"
#include "srsran/phy/phch/ra_dl.h"
#include "srsran/phy/phch/ra_ul.h"
#include "srsran/phy/phch/regs.h"
#include "srsran/phy/utils/debug.h"
#include "srsran/phy/utils/vector.h"
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_MAX_CANDIDATES_UE 16 // From 36.213 Table 9.1.1-1
#define SRSRAN_MAX_CANDIDATES_COM 6 // From 36.213 Table 9.1.1-1
#define SRSRAN_MAX_CANDIDATES (SRSRAN_MAX_CANDIDATES_UE + SRSRAN_MAX_CANDIDATES_COM)
#define SRSRAN_MAX_FORMATS 4
#define SRSRAN_MI_NOF_REGS ((q->cell.frame_type == SRSRAN_FDD) ? 1 : 6)
#define SRSRAN_MI_MAX_REGS 6
#define SRSRAN_MAX_DCI_MSG SRSRAN_MAX_CARRIERS

typedef struct SRSRAN_API { srsran_cell_t cell; srsran_regs_t* regs; uint32_t nof_rx_antennas; } srsran_pcfich_t;

typedef struct SRSRAN_API {
  srsran_pcfich_t       pcfich;
  srsran_phich_t        phich;
  srsran_pdcch_t        pdcch;
  srsran_pdsch_t        pdsch;
  srsran_regs_t         regs[SRSRAN_MI_MAX_REGS];
  uint32_t              mi_manual_index;
  bool                  mi_auto;
  srsran_chest_dl_t     chest;
  srsran_chest_dl_res_t chest_res;
  srsran_ofdm_t         fft[SRSRAN_MAX_PORTS];
  srsran_dci_msg_t      pending_ul_dci_msg[SRSRAN_MAX_DCI_MSG];
  uint32_t              pending_ul_dci_count;
  cf_t*                 sf_symbols[SRSRAN_MAX_PORTS]; /* per rx antenna, 14 * 12 * nof_prb each */
  srsran_cell_t         cell;
  uint32_t              nof_rx_antennas;
  uint16_t              pregen_rnti;
  void*                 b200; /* compat: the CUDA PHY handle and host mirrors behind this object */
} srsran_ue_dl_t;

typedef struct SRSRAN_API {
  srsran_cqi_report_cfg_t cqi_report;
  srsran_pdsch_cfg_t      pdsch;
  srsran_dci_cfg_t        dci;
  srsran_tm_t             tm;
  bool                    dci_common_ss;
  bool                    pdsch_use_tbs_index_alt;
} srsran_dl_cfg_t;
typedef struct SRSRAN_API {
  srsran_dl_cfg_t       cfg;
  srsran_chest_dl_cfg_t chest_cfg;
  uint32_t              last_ri;
  float                 snr_to_cqi_offset;
} srsran_ue_dl_cfg_t;
typedef struct { uint32_t v_dai_dl; uint32_t n_cce; uint32_t grant_cc_idx; uint32_t tpc_for_pucch; } srsran_pdsch_ack_resource_t;

SRSRAN_API int  srsran_ue_dl_init(srsran_ue_dl_t* q, cf_t* input[SRSRAN_MAX_PORTS], uint32_t max_prb, uint32_t nof_rx_antennas);
SRSRAN_API void srsran_ue_dl_free(srsran_ue_dl_t* q);
SRSRAN_API int  srsran_ue_dl_set_cell(srsran_ue_dl_t* q, srsran_cell_t cell);
SRSRAN_API void srsran_ue_dl_set_rnti(srsran_ue_dl_t* q, uint16_t rnti);
SRSRAN_API int  srsran_ue_dl_decode_fft_estimate(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_ue_dl_cfg_t* cfg);
SRSRAN_API int  srsran_ue_dl_decode_fft_estimate_noguru(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_ue_dl_cfg_t* cfg, cf_t* input[SRSRAN_MAX_PORTS]);
SRSRAN_API int  srsran_ue_dl_decode_pdsch(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_pdsch_cfg_t* pdsch_cfg, srsran_pdsch_res_t data[SRSRAN_MAX_CODEWORDS]);
#ifdef __cplusplus
}
#endif
#endif
