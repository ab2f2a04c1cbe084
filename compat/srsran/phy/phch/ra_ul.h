/* srsran/phy/phch/ra_ul.h (compat): UL DCI -> grant helpers (36.213 8.1, 8.4, 8.6) */
#ifndef SRSRAN_RA_UL_H
#define SRSRAN_RA_UL_H
#include "srsran/phy/phch/dci.h"
#include "srsran/phy/phch/pusch_cfg.h"
#ifdef __cplusplus
extern "C" {
#endif
SRSRAN_API int  srsran_ra_ul_nof_re(srsran_pusch_grant_t* grant, srsran_cp_t cp, uint32_t N_srs);
SRSRAN_API int  srsran_ra_ul_dci_to_grant(srsran_cell_t* cell, srsran_ul_sf_cfg_t* sf, srsran_pusch_hopping_cfg_t* hopping_cfg, srsran_dci_ul_t* dci,
                                          srsran_pusch_grant_t* grant);
SRSRAN_API void srsran_ra_ul_compute_nof_re(srsran_pusch_grant_t* grant, srsran_cp_t cp, uint32_t N_srs);
#ifdef __cplusplus
}
#endif
#endif
