/* srsran/phy/phch/regs.h (compat): control-region resource element groups (36.211 6.2.4).  Opaque here: the REG / CCE maps
 * live inside libltephy_b200 (lte_host.cpp: build_ctrl_map). */
#ifndef SRSRAN_REGS_H
#define SRSRAN_REGS_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  srsran_cell_t cell;
  uint32_t      max_ctrl_symbols;
  uint32_t      ngroups_phich;
  uint32_t      ngroups_phich_m1;
  uint32_t      nof_regs;
  void*         b200; /* compat: owner object */
} srsran_regs_t;
#ifdef __cplusplus
}
#endif
#endif
