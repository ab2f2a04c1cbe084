/* srsran/phy/phch/prach.h (compat): PRACH configuration (src/src/ULSchedule.cc:147-153) and detector object
 * (src/src/UL_Sniffer_PUSCH.cc:657-713) */
#ifndef SRSRAN_PRACH_H
#define SRSRAN_PRACH_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_PRACH_MAX_LEN (2 * 24576 + 21024) // Maximum Tcp + Tseq
typedef struct SRSRAN_API {
  bool                is_nr;
  uint32_t            config_idx;
  uint32_t            root_seq_idx;
  uint32_t            zero_corr_zone;
  uint32_t            freq_offset;
  uint32_t            num_ra_preambles;
  bool                hs_flag;
  srsran_tdd_config_t tdd_config;
  bool                enable_successive_cancellation;
  bool                enable_freq_domain_offset_calc;
} srsran_prach_cfg_t;
typedef struct SRSRAN_API {
  bool               is_nr;
  uint32_t           f;
  uint32_t           rsi;
  uint32_t           zczc;
  uint32_t           N_ifft_ul;
  uint32_t           N_ifft_prach;
  uint32_t           max_N_ifft_ul;
  uint32_t           N_zc;
  uint32_t           N_cs;
  uint32_t           N_seq;
  uint32_t           N_cp;
  float              T_seq;
  float              T_tot;
  srsran_prach_cfg_t cfg;
  float              detect_factor;
  void*              b200;
} srsran_prach_t;
#ifdef __cplusplus
}
#endif
#endif
