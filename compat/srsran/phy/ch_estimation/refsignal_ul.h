/* srsran/phy/ch_estimation/refsignal_ul.h (compat): uplink DMRS configuration (36.211 5.5) */
#ifndef SRSRAN_REFSIGNAL_UL_H
#define SRSRAN_REFSIGNAL_UL_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_NOF_GROUPS_U 30
#define SRSRAN_NOF_SEQUENCES_U 2
#define SRSRAN_NOF_DELTA_SS 30
#define SRSRAN_NOF_CSHIFT 8
#define SRSRAN_REFSIGNAL_UL_L(ns_idx, cp) ((ns_idx + 1) * SRSRAN_CP_NSYMB(cp) - 4)
typedef struct SRSRAN_API {
  uint32_t cyclic_shift;
  uint32_t delta_ss;
  bool     group_hopping_en;
  bool     sequence_hopping_en;
} srsran_refsignal_dmrs_pusch_cfg_t;
typedef struct SRSRAN_API {
  uint32_t subframe_config;
  uint32_t bw_cfg;
  bool     simul_ack;
  uint32_t B;
  uint32_t b_hop;
  uint32_t n_srs;
  uint32_t I_srs;
  uint32_t k_tc;
  uint32_t n_rrc;
  bool     dedicated_enabled;
  bool     common_enabled;
  bool     configured;
} srsran_refsignal_srs_cfg_t;
typedef struct SRSRAN_API {
  srsran_cell_t cell;
  float*        tmp_arg;
  uint32_t      n_cs_cell[SRSRAN_NSLOTS_X_FRAME][SRSRAN_CP_NORM_NSYMB];
  uint32_t      n_prs_pusch[SRSRAN_NOF_DELTA_SS][SRSRAN_NSLOTS_X_FRAME];
  uint32_t      f_gh[SRSRAN_NSLOTS_X_FRAME];
  uint32_t      u_pucch[SRSRAN_NSLOTS_X_FRAME];
  uint32_t      v_pusch[SRSRAN_NSLOTS_X_FRAME][SRSRAN_NOF_DELTA_SS];
} srsran_refsignal_ul_t;
#ifdef __cplusplus
}
#endif
#endif
