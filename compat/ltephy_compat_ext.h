/* ltephy_compat_ext.h -- extensions of libltephy_srsran_compat beyond the srsRAN names.
 *
 * ltephy_compat_inject: hands the NEXT srsran_ue_dl_decode_fft_estimate call on q a ready-made phase-A result (per-subframe record,
 * blind-decode table T[location][size], PDCCH LLRs) instead of running the GPU.  It exists so that the reference's unmodified
 * blind search (src/src/DCISearch.cc, lib/src/phy/falcon_phch/falcon_pdcch.c) can be run on host-supplied tables -- the parity
 * tests use it to compare the reference's own walk with ltephy_search_batch on identical input, with or without a GPU. */
#ifndef LTEPHY_COMPAT_EXT_H
#define LTEPHY_COMPAT_EXT_H
#include "srsran/phy/ue/ue_dl.h"
#include "../include/ltephy_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* info: finalised record of the subframe; table: [LTEPHY_MAX_LOC][LTEPHY_MAX_SIZES]; llr: 72 * info->nof_cce floats (may be NULL: zeros) */
int ltephy_compat_inject(srsran_ue_dl_t* q, const ltephy_sf_info_t* info, const ltephy_cand_t* table, const float* llr);
/* the tier-1 PHY handle behind q (NULL before the first GPU subframe) */
ltephy_t* ltephy_compat_phy(srsran_ue_dl_t* q);
#ifdef __cplusplus
}
#endif
#endif
