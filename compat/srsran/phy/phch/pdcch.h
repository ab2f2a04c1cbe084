/* srsran/phy/phch/pdcch.h (compat): PDCCH object.  q->llr, q->nof_cce[] and q->nof_regs[] are what
 * lib/src/phy/falcon_phch/falcon_pdcch.c reads (:36-37, :138, :276, :607); srsran_pdcch_dci_decode is the call the reference makes
 * per (location, format) (falcon_pdcch.c:142) -- here it is a lookup in the table the GPU filled for the whole subframe. */
#ifndef SRSRAN_PDCCH_H
#define SRSRAN_PDCCH_H
#include "srsran/phy/ch_estimation/chest_dl.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/fec/convolutional/convcoder.h"
#include "srsran/phy/fec/convolutional/rm_conv.h"
#include "srsran/phy/fec/convolutional/viterbi.h"
#include "srsran/phy/fec/crc.h"
#include "srsran/phy/phch/dci.h"
#include "srsran/phy/phch/regs.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum SRSRAN_API { SEARCH_UE, SEARCH_COMMON } srsran_pdcch_search_mode_t;
typedef struct SRSRAN_API {
  srsran_cell_t cell;
  uint32_t      nof_regs[3];
  uint32_t      nof_cce[3];
  uint32_t      max_bits;
  uint32_t      nof_rx_antennas;
  bool          is_ue;
  srsran_regs_t* regs;
  /* buffers */
  cf_t*    ce[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS];
  cf_t*    symbols[SRSRAN_MAX_PORTS];
  cf_t*    x[SRSRAN_MAX_LAYERS];
  cf_t*    d;
  uint8_t* e;
  float    rm_f[3 * (SRSRAN_DCI_MAX_BITS + 16)];
  float*   llr;
  /* tx & rx objects */
  srsran_viterbi_t   decoder;
  srsran_crc_t       crc;
  srsran_convcoder_t encoder;
  /* compat: the PHY context whose current subframe's candidate table backs srsran_pdcch_dci_decode */
  void* b200;
} srsran_pdcch_t;
SRSRAN_API float    srsran_pdcch_coderate(uint32_t nof_bits, uint32_t l);
SRSRAN_API int      srsran_pdcch_extract_llr(srsran_pdcch_t* q, srsran_dl_sf_cfg_t* sf, srsran_chest_dl_res_t* channel,
                                             cf_t* sf_symbols[SRSRAN_MAX_PORTS]);
SRSRAN_API int      srsran_pdcch_dci_decode(srsran_pdcch_t* q, float* e, uint8_t* data, uint32_t E, uint32_t nof_bits, uint16_t* crc);
SRSRAN_API void     srsran_pdcch_dci_encode_conv(srsran_pdcch_t* q, uint8_t* data, uint32_t nof_bits, uint8_t* coded_data, uint16_t rnti);
SRSRAN_API uint32_t srsran_pdcch_ue_locations_ncce(uint32_t nof_cce, srsran_dci_location_t* c, uint32_t max_candidates, uint32_t sf_idx, uint16_t rnti);
SRSRAN_API uint32_t srsran_pdcch_ue_locations_ncce_L(uint32_t nof_cce, srsran_dci_location_t* c, uint32_t max_candidates, uint32_t sf_idx,
                                                     uint16_t rnti, int L);
SRSRAN_API uint32_t srsran_pdcch_common_locations_ncce(uint32_t nof_cce, srsran_dci_location_t* c, uint32_t max_candidates);
#ifdef __cplusplus
}
#endif
#endif
