/* srsran/phy/dft/ofdm.h (compat): OFDM (de)modulator object; the transform runs on the GPU (k_frontend.cu: ofdm_rx_kernel) */
#ifndef SRSRAN_OFDM_H
#define SRSRAN_OFDM_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SRSRAN_API {
  uint32_t    nof_prb;
  cf_t*       in_buffer;
  cf_t*       out_buffer;
  uint32_t    symbol_sz;
  srsran_cp_t cp;
  float       freq_shift_f;
  float       rx_window_offset;
  bool        normalize;
  bool        keep_dc;
} srsran_ofdm_cfg_t;
typedef struct SRSRAN_API {
  srsran_ofdm_cfg_t cfg;
  uint32_t          nof_symbols;
  uint32_t          nof_guards;
  uint32_t          nof_re;
  uint32_t          slot_sz;
  uint32_t          sf_sz;
} srsran_ofdm_t;
#ifdef __cplusplus
}
#endif
#endif
