/* srsran/phy/phch/uci_cfg.h (compat): UCI carried on PUSCH (36.212 5.2.2.6) */
#ifndef SRSRAN_UCI_CFG_H
#define SRSRAN_UCI_CFG_H
#include "srsran/phy/common/phy_common.h"
#ifdef __cplusplus
extern "C" {
#endif
#define SRSRAN_UCI_MAX_ACK_BITS 10
#define SRSRAN_UCI_MAX_ACK_SR_BITS (SRSRAN_UCI_MAX_ACK_BITS + 1)
#define SRSRAN_UCI_MAX_M 9
#define SRSRAN_UCI_MAX_CQI_LEN_PUSCH 512
#define SRSRAN_UCI_MAX_CQI_LEN_PUCCH 13
#define SRSRAN_CQI_MAX_BITS 64
typedef struct SRSRAN_API {
  uint8_t  ack_value[SRSRAN_UCI_MAX_ACK_BITS];
  bool     valid;
} srsran_uci_value_ack_t;
typedef struct SRSRAN_API {
  bool     pending_tb[SRSRAN_MAX_CODEWORDS]; //< Indicates whether there was a grant that requires an ACK/NACK
  uint32_t nof_acks;                         //< Number of transport blocks, deduced from transmission mode
  uint32_t ncce[SRSRAN_UCI_MAX_M];
  uint32_t N_bundle;
  uint32_t tdd_ack_M;
  uint32_t tdd_ack_m;
  bool     tdd_is_multiplex;
  uint32_t tpc_for_pucch;
  uint32_t grant_cc_idx;
} srsran_uci_cfg_ack_t;
typedef enum SRSRAN_API { SRSRAN_CQI_TYPE_WIDEBAND = 0, SRSRAN_CQI_TYPE_SUBBAND_UE, SRSRAN_CQI_TYPE_SUBBAND_UE_DIFF, SRSRAN_CQI_TYPE_SUBBAND_HL } srsran_cqi_type_t;
typedef struct SRSRAN_API {
  bool             data_enable;
  bool             pmi_present;
  bool             four_antenna_ports;
  bool             rank_is_not_one;
  bool             subband_label_2_bits;
  uint32_t         scell_index;
  uint32_t         L;
  uint32_t         N;
  srsran_cqi_type_t type;
  uint32_t         ri_len;
} srsran_cqi_cfg_t;
typedef struct SRSRAN_API {
  srsran_uci_cfg_ack_t ack[SRSRAN_MAX_CARRIERS];
  srsran_cqi_cfg_t     cqi;
  bool                 is_scheduling_request_tti;
} srsran_uci_cfg_t;
typedef struct SRSRAN_API {
  uint8_t wideband_cqi;
  uint8_t raw[16];
  bool    data_crc;
} srsran_cqi_value_t;
typedef struct SRSRAN_API {
  bool                   scheduling_request;
  srsran_cqi_value_t     cqi;
  srsran_uci_value_ack_t ack;
  uint8_t                ri;
} srsran_uci_value_t;
typedef struct SRSRAN_API {
  srsran_uci_cfg_t   cfg;
  srsran_uci_value_t value;
} srsran_uci_data_t;
typedef struct SRSRAN_API { uint32_t I_offset_cqi; uint32_t I_offset_ri; uint32_t I_offset_ack; } srsran_uci_offset_cfg_t;
#ifdef __cplusplus
}
#endif
#endif
