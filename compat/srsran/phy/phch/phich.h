/* srsran/phy/phch/phich.h (compat): only the names src/include/HARQ.h needs */
#ifndef SRSRAN_PHICH_H
#define SRSRAN_PHICH_H
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/phch/regs.h"
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_PHICH_NORM_NSEQUENCES 8
#define SRSRAN_PHICH_EXT_NSEQUENCES 4
#define SRSRAN_PHICH_NBITS 3
typedef struct SRSRAN_API { uint32_t ngroup; uint32_t nseq; } srsran_phich_resource_t;
typedef struct SRSRAN_API { uint32_t n_prb_lowest; uint32_t n_dmrs; uint32_t I_phich; } srsran_phich_grant_t;
typedef struct SRSRAN_API { bool ack_value; float distance; } srsran_phich_res_t;
typedef struct SRSRAN_API { srsran_cell_t cell; uint32_t nof_rx_antennas; srsran_regs_t* regs; } srsran_phich_t;
#ifdef __cplusplus
}
#endif
#endif
