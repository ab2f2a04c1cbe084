/* srsran/phy/phch/dci.h (compat): DCI messages and their unpacked forms (36.212 5.3.3).  srsRAN 21.10 names. */
#ifndef SRSRAN_DCI_H
#define SRSRAN_DCI_H
#include "srsran/config.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/phch/ra.h"
#ifdef __cplusplus
extern "C" {
#endif

#define SRSRAN_DCI_MAX_BITS 128
#define SRSRAN_RAR_GRANT_LEN 20
#define SRSRAN_DCI_IS_TB_EN(tb) (!(tb.mcs_idx == 0 && tb.rv == 1))
#define SRSRAN_DCI_TB_DISABLE(tb) do { tb.mcs_idx = 0; tb.rv = 1; } while (0)
#define SRSRAN_DCI_HEXDEBUG 0

typedef enum {
  SRSRAN_DCI_FORMAT0 = 0, SRSRAN_DCI_FORMAT1, SRSRAN_DCI_FORMAT1A, SRSRAN_DCI_FORMAT1B, SRSRAN_DCI_FORMAT1C, SRSRAN_DCI_FORMAT1D,
  SRSRAN_DCI_FORMAT2, SRSRAN_DCI_FORMAT2A, SRSRAN_DCI_FORMAT2B,
  // SRSRAN_DCI_FORMAT3, SRSRAN_DCI_FORMAT3A,
  SRSRAN_DCI_FORMAT_RAR, // Not a real LTE format. Used internally to indicate RAR grant
  SRSRAN_DCI_NOF_FORMATS
} srsran_dci_format_t;

typedef struct {
  bool multiple_csi_request_enabled;
  bool cif_enabled;
  bool cif_present;
  bool srs_request_enabled;
  bool ra_format_enabled;
  bool is_not_ue_ss;
} srsran_dci_cfg_t;

typedef struct SRSRAN_API { uint32_t L; uint32_t ncce; } srsran_dci_location_t;

typedef struct SRSRAN_API {
  uint8_t               payload[SRSRAN_DCI_MAX_BITS];
  uint32_t              nof_bits;
  srsran_dci_location_t location;
  srsran_dci_format_t   format;
  uint16_t              rnti;
} srsran_dci_msg_t;

typedef struct SRSRAN_API { uint32_t mcs_idx; int rv; bool ndi; uint32_t cw_idx; } srsran_dci_tb_t;

typedef struct SRSRAN_API {
  uint16_t              rnti;
  srsran_dci_format_t   format;
  srsran_dci_location_t location;
  uint32_t              ue_cc_idx;
  // Resource Allocation
  srsran_ra_type_t alloc_type;
  union {
    srsran_ra_type0_t type0_alloc;
    srsran_ra_type1_t type1_alloc;
    srsran_ra_type2_t type2_alloc;
  };
  // Codeword information
  srsran_dci_tb_t tb[SRSRAN_MAX_CODEWORDS];
  bool            tb_cw_swap;
  uint32_t        pinfo;
  // Power control
  bool    pconf;
  bool    power_offset;
  uint8_t tpc_pucch;
  // RA order
  bool     is_ra_order;
  uint32_t ra_preamble;
  uint32_t ra_mask_idx;
  // Release 10
  uint32_t cif;
  bool     cif_present;
  bool     srs_request;
  bool     srs_request_present;
  // Other parameters
  uint32_t pid;
  uint32_t dai;
  bool     is_tdd;
  bool     is_dwpts;
  bool     sram_id;
} srsran_dci_dl_t;

/* 36.213 Table 8.4-2: SRSRAN_RA_PUSCH_HOP_HALF is 0 for < 10 Mhz and 10 for > 10 Mhz.
 * SRSRAN_RA_PUSCH_HOP_QUART is 00 for > 10 Mhz and SRSRAN_RA_PUSCH_HOP_QUART_NEG is 01 for > 10 Mhz. */
typedef enum {
  SRSRAN_RA_PUSCH_HOP_DISABLED  = -1,
  SRSRAN_RA_PUSCH_HOP_QUART     = 0,
  SRSRAN_RA_PUSCH_HOP_QUART_NEG = 1,
  SRSRAN_RA_PUSCH_HOP_HALF      = 2,
  SRSRAN_RA_PUSCH_HOP_TYPE2     = 3
} srsran_ra_pusch_hop_t;

typedef struct SRSRAN_API {
  uint16_t              rnti;
  srsran_dci_format_t   format;
  srsran_dci_location_t location;
  uint32_t              ue_cc_idx;
  srsran_ra_type2_t     type2_alloc;
  srsran_ra_pusch_hop_t freq_hop_fl;
  // Codeword information
  srsran_dci_tb_t tb;
  uint32_t        n_dmrs;
  bool            cqi_request;
  // TDD parametres
  uint32_t dai;
  uint32_t ul_idx;
  bool     is_tdd;
  // Power control
  uint8_t tpc_pusch;
  // Release 10
  uint32_t         cif;
  bool             cif_present;
  uint8_t          multiple_csi_request;
  bool             multiple_csi_request_present;
  bool             srs_request;
  bool             srs_request_present;
  srsran_ra_type_t ra_type;
  bool             ra_type_present;
} srsran_dci_ul_t;

typedef struct SRSRAN_API {
  uint32_t rba;
  uint32_t trunc_mcs;
  uint32_t tpc_pusch;
  bool     ul_delay;
  bool     cqi_request;
  bool     hopping_flag;
} srsran_dci_rar_grant_t;

SRSRAN_API int      srsran_dci_msg_unpack_pusch(srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_dci_cfg_t* cfg, srsran_dci_msg_t* msg, srsran_dci_ul_t* dci);
SRSRAN_API int      srsran_dci_msg_unpack_pdsch(srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_dci_cfg_t* cfg, srsran_dci_msg_t* msg, srsran_dci_dl_t* dci);
SRSRAN_API uint32_t srsran_dci_format_sizeof(const srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_dci_cfg_t* cfg, srsran_dci_format_t format);
SRSRAN_API uint32_t srsran_dci_format_max_tb(srsran_dci_format_t format);
SRSRAN_API bool     srsran_dci_location_isvalid(srsran_dci_location_t* c);
SRSRAN_API void     srsran_dci_cfg_set_common_ss(srsran_dci_cfg_t* cfg);
SRSRAN_API uint32_t srsran_dci_dl_info(const srsran_dci_dl_t* dci_dl, char* str, uint32_t str_len);
SRSRAN_API uint32_t srsran_dci_ul_info(srsran_dci_ul_t* dci_ul, char* info_str, uint32_t len);
SRSRAN_API char*    srsran_dci_format_string(srsran_dci_format_t format);
SRSRAN_API char*    srsran_dci_format_string_short(srsran_dci_format_t format);
SRSRAN_API int      srsran_dci_location_set(srsran_dci_location_t* c, uint32_t L, uint32_t nCCE);
#ifdef __cplusplus
}
#endif
#endif
