/* srsran/phy/phch/cqi.h (compat): CQI report configuration helpers (types in uci_cfg.h) */
#ifndef SRSRAN_CQI_H
#define SRSRAN_CQI_H
#include "srsran/phy/phch/uci_cfg.h"
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_CQI_STR_MAX_CHAR 64
typedef struct SRSRAN_API {
  bool     periodic_configured;
  bool     aperiodic_configured;
  uint32_t pmi_idx;
  uint32_t ri_idx;
  bool     ri_idx_present;
  bool     format_is_subband;
  uint32_t subband_size;
  int      periodic_mode;
  int      aperiodic_mode;
} srsran_cqi_report_cfg_t;
SRSRAN_API int srsran_cqi_size(srsran_cqi_cfg_t* cfg);
SRSRAN_API int srsran_cqi_hl_get_subband_size(int num_prbs);
SRSRAN_API int srsran_cqi_hl_get_no_subbands(int num_prbs);
#ifdef __cplusplus
}
#endif
#endif
