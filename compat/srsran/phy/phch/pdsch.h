/* srsran/phy/phch/pdsch.h (compat): PDSCH object and srsran_pdsch_decode.  In libltephy_b200 one call decodes every transport block
 * of the grant on the GPU (k_pdsch.cu, k_turbo.cu); the batched form is ltephy_submit_grants / ltephy_get_phase_b. */
#ifndef SRSRAN_PDSCH_H
#define SRSRAN_PDSCH_H
#include "srsran/phy/ch_estimation/chest_dl.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/phch/dci.h"
#include "srsran/phy/phch/pdsch_cfg.h"
#include "srsran/phy/phch/regs.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  srsran_cell_t cell;
  uint32_t      nof_rx_antennas;
  uint32_t      max_re;
  bool          is_ue;
  bool          llr_is_8bit;
  void*         b200; /* compat: owner PHY context */
} srsran_pdsch_t;
typedef struct {
  uint8_t* payload;
  bool     crc;
  float    avg_iterations_block;
  float    evm;
} srsran_pdsch_res_t;
SRSRAN_API int  srsran_pdsch_decode(srsran_pdsch_t* q, srsran_dl_sf_cfg_t* sf, srsran_pdsch_cfg_t* cfg, srsran_chest_dl_res_t* channel,
                                    cf_t* sf_symbols[SRSRAN_MAX_PORTS], srsran_pdsch_res_t data[SRSRAN_MAX_CODEWORDS]);
#ifdef __cplusplus
}
#endif
#endif
