/* srsran/phy/phch/ra_dl.h (compat): DL DCI -> grant helpers (36.213 7.1.6, 7.1.7) */
#ifndef SRSRAN_RA_DL_H
#define SRSRAN_RA_DL_H
#include "srsran/phy/phch/dci.h"
#include "srsran/phy/phch/pdsch_cfg.h"
#include "srsran/phy/phch/ra.h"
#ifdef __cplusplus
extern "C" {
#endif
SRSRAN_API int      srsran_ra_dl_grant_to_grant_prb_allocation(const srsran_dci_dl_t* dci, srsran_pdsch_grant_t* grant, uint32_t nof_prb);
SRSRAN_API int      srsran_dl_fill_ra_mcs(srsran_ra_tb_t* tb, int last_tbs, uint32_t nprb, bool pdsch_use_tbs_index_alt);
SRSRAN_API void     srsran_ra_dl_compute_nof_re(const srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_pdsch_grant_t* grant);
SRSRAN_API uint32_t srsran_ra_dl_grant_nof_re(const srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_pdsch_grant_t* grant);
#ifdef __cplusplus
}
#endif
#endif
