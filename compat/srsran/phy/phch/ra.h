/* srsran/phy/phch/ra.h (compat): resource-allocation types shared by DL and UL (36.213 7.1.6 / 8.1).
 * (The pre-20.04 srsran_ra_dl_dci_t / _dl_grant_t / _ul_dci_t / _ul_grant_t structures of the DCI trace are the reference's own:
 * lib/include/falcon/common/falcon_define.h.) */
#ifndef SRSRAN_RA_H
#define SRSRAN_RA_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef enum SRSRAN_API { SRSRAN_RA_ALLOC_TYPE0 = 0, SRSRAN_RA_ALLOC_TYPE1 = 1, SRSRAN_RA_ALLOC_TYPE2 = 2 } srsran_ra_type_t;
typedef struct SRSRAN_API { uint32_t rbg_bitmask; } srsran_ra_type0_t;
typedef struct SRSRAN_API { uint32_t vrb_bitmask; uint32_t rbg_subset; bool shift; } srsran_ra_type1_t;
typedef struct SRSRAN_API {
  uint32_t riv; // if L_crb==0, DCI message packer will take this value directly
  enum { SRSRAN_RA_TYPE2_NPRB1A_2 = 0, SRSRAN_RA_TYPE2_NPRB1A_3 = 1 } n_prb1a;
  enum { SRSRAN_RA_TYPE2_NG1 = 0, SRSRAN_RA_TYPE2_NG2 = 1 } n_gap;
  enum { SRSRAN_RA_TYPE2_LOC = 0, SRSRAN_RA_TYPE2_DIST = 1 } mode;
} srsran_ra_type2_t;

#define SRSRAN_RA_NOF_TBS_IDX 34

typedef struct SRSRAN_API {
  srsran_mod_t mod;
  int          tbs;
  int          rv;
  uint32_t     nof_bits;
  uint32_t     cw_idx;
  bool         enabled;
  // this is for debugging and metrics purposes
  uint32_t mcs_idx;
} srsran_ra_tb_t;

SRSRAN_API uint32_t srsran_ra_type0_P(uint32_t nof_prb);
SRSRAN_API uint32_t srsran_ra_type2_to_riv(uint32_t L_crb, uint32_t RB_start, uint32_t nof_prb);
SRSRAN_API void     srsran_ra_type2_from_riv(uint32_t riv, uint32_t* L_crb, uint32_t* RB_start, uint32_t nof_prb, uint32_t nof_vrb);
SRSRAN_API int      srsran_ra_tbs_idx_from_mcs(uint32_t mcs, bool use_tbs_index_alt, bool is_ul);
SRSRAN_API srsran_mod_t srsran_ra_dl_mod_from_mcs(uint32_t mcs, bool use_tbs_index_alt);
SRSRAN_API srsran_mod_t srsran_ra_ul_mod_from_mcs(uint32_t mcs);
SRSRAN_API int      srsran_ra_tbs_from_idx(uint32_t tbs_idx, uint32_t n_prb);
SRSRAN_API uint32_t srsran_ra_type1_N_rb(uint32_t nof_prb);
#ifdef __cplusplus
}
#endif
#endif
