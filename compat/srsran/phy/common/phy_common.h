/* srsran/phy/common/phy_common.h (compat): cell description, RNTI ranges, numerology macros.  srsRAN 21.10 names. */
#ifndef SRSRAN_PHY_COMMON_H
#define SRSRAN_PHY_COMMON_H
#include "srsran/config.h"
#include <math.h>
#include "srsran/phy/utils/debug.h"
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SRSRAN_SUCCESS 0
#define SRSRAN_ERROR -1
#define SRSRAN_ERROR_INVALID_INPUTS -2
#define SRSRAN_ERROR_TIMEOUT -3
#define SRSRAN_ERROR_INVALID_COMMAND -4
#define SRSRAN_ERROR_OUT_OF_BOUNDS -5
#define SRSRAN_ERROR_CANT_START -6
#define SRSRAN_ERROR_ALREADY_STARTED -7

#define SRSRAN_NOF_SF_X_FRAME 10
#define SRSRAN_NOF_SLOTS_PER_SF 2
#define SRSRAN_NSLOTS_X_FRAME (SRSRAN_NOF_SLOTS_PER_SF * SRSRAN_NOF_SF_X_FRAME)
#define SRSRAN_NSOFT_BITS 250368
#define SRSRAN_PC_MAX 23
#define SRSRAN_NOF_NID_1 168
#define SRSRAN_NOF_NID_2 3
#define SRSRAN_NUM_PCI (SRSRAN_NOF_NID_1 * SRSRAN_NOF_NID_2)
#define SRSRAN_MAX_CARRIERS 5
#define SRSRAN_MAX_PORTS 4
#define SRSRAN_MAX_CHANNELS (SRSRAN_MAX_CARRIERS * SRSRAN_MAX_PORTS)
#define SRSRAN_MAX_LAYERS 4
#define SRSRAN_MAX_CODEWORDS 2
#define SRSRAN_MAX_TB SRSRAN_MAX_CODEWORDS
#define SRSRAN_MAX_QM 8
#define SRSRAN_MAX_CODEBLOCKS 32
#define SRSRAN_MAX_CODEBOOK_IDX_LEN 2
#define SRSRAN_MAX_CODEBOOKS 4
#define SRSRAN_NOF_CFI 3
#define SRSRAN_CFI_ISVALID(x) ((x >= 1 && x <= 3))
#define SRSRAN_CFI_IDX(x) ((x - 1) % SRSRAN_NOF_CFI)
#define SRSRAN_LTE_CRC24A 0x1864CFB
#define SRSRAN_LTE_CRC24B 0X1800063
#define SRSRAN_LTE_CRC24C 0X1B2B117
#define SRSRAN_LTE_CRC16 0x11021
#define SRSRAN_LTE_CRC11 0xE21
#define SRSRAN_LTE_CRC8 0x19B
#define SRSRAN_LTE_CRC6 0x61
#define SRSRAN_MAX_MBSFN_AREA_IDS 256
#define SRSRAN_PMCH_RV 0

typedef enum { SRSRAN_CP_NORM = 0, SRSRAN_CP_EXT } srsran_cp_t;
typedef enum { SRSRAN_SF_NORM = 0, SRSRAN_SF_MBSFN } srsran_sf_t;

#define SRSRAN_INVALID_RNTI 0x0
#define SRSRAN_CRNTI_START 0x000B
#define SRSRAN_CRNTI_END 0xFFF3
#define SRSRAN_RARNTI_START 0x0001
#define SRSRAN_RARNTI_END 0x000A
#define SRSRAN_SIRNTI 0xFFFF
#define SRSRAN_PRNTI 0xFFFE
#define SRSRAN_MRNTI 0xFFFD
#define SRSRAN_RNTI_ISRAR(rnti) (rnti >= SRSRAN_RARNTI_START && rnti <= SRSRAN_RARNTI_END)
#define SRSRAN_RNTI_ISUSER(rnti) (rnti >= SRSRAN_CRNTI_START && rnti <= SRSRAN_CRNTI_END)
#define SRSRAN_RNTI_ISSI(rnti) (rnti == SRSRAN_SIRNTI)
#define SRSRAN_RNTI_ISPA(rnti) (rnti == SRSRAN_PRNTI)
#define SRSRAN_RNTI_ISMBSFN(rnti) (rnti == SRSRAN_MRNTI)
#define SRSRAN_RNTI_ISSIRAPA(rnti) (SRSRAN_RNTI_ISSI(rnti) || SRSRAN_RNTI_ISRAR(rnti) || SRSRAN_RNTI_ISPA(rnti))

#define SRSRAN_CELL_ID_UNKNOWN 1000
#define SRSRAN_MAX_NSYMB 7
#define SRSRAN_MAX_PRB 110
#define SRSRAN_NRE 12
#define SRSRAN_SYMBOL_SZ_MAX 2048
#define SRSRAN_CP_NORM_NSYMB 7
#define SRSRAN_CP_NORM_SF_NSYMB (2 * SRSRAN_CP_NORM_NSYMB)
#define SRSRAN_CP_NORM_0_LEN 160
#define SRSRAN_CP_NORM_LEN 144
#define SRSRAN_CP_EXT_NSYMB 6
#define SRSRAN_CP_EXT_SF_NSYMB (2 * SRSRAN_CP_EXT_NSYMB)
#define SRSRAN_CP_EXT_LEN 512
#define SRSRAN_CP_ISNORM(cp) (cp == SRSRAN_CP_NORM)
#define SRSRAN_CP_ISEXT(cp) (cp == SRSRAN_CP_EXT)
#define SRSRAN_CP_NSYMB(cp) (SRSRAN_CP_ISNORM(cp) ? SRSRAN_CP_NORM_NSYMB : SRSRAN_CP_EXT_NSYMB)
#define SRSRAN_CP_LEN(symbol_sz, c) ((int)ceilf((((float)(c) * (symbol_sz)) / 2048.0f)))
#define SRSRAN_CP_LEN_NORM(symbol, symbol_sz) (((symbol) == 0) ? SRSRAN_CP_LEN((symbol_sz), SRSRAN_CP_NORM_0_LEN) : SRSRAN_CP_LEN((symbol_sz), SRSRAN_CP_NORM_LEN))
#define SRSRAN_CP_LEN_EXT(symbol_sz) (SRSRAN_CP_LEN((symbol_sz), SRSRAN_CP_EXT_LEN))
#define SRSRAN_CP_SZ(symbol_sz, cp) (SRSRAN_CP_LEN(symbol_sz, (SRSRAN_CP_ISNORM(cp) ? SRSRAN_CP_NORM_LEN : SRSRAN_CP_EXT_LEN)))
#define SRSRAN_SYMBOL_SZ(symbol_sz, cp) (symbol_sz + SRSRAN_CP_SZ(symbol_sz, cp))
#define SRSRAN_SLOT_LEN(symbol_sz) (symbol_sz * 15 / 2)
#define SRSRAN_SF_LEN(symbol_sz) (symbol_sz * 15)
#define SRSRAN_SF_LEN_MAX (SRSRAN_SF_LEN(SRSRAN_SYMBOL_SZ_MAX))
#define SRSRAN_SLOT_LEN_PRB(nof_prb) (SRSRAN_SLOT_LEN(srsran_symbol_sz(nof_prb)))
#define SRSRAN_SF_LEN_PRB(nof_prb) ((uint32_t)SRSRAN_SF_LEN(srsran_symbol_sz(nof_prb)))
#define SRSRAN_SLOT_LEN_RE(nof_prb, cp) (nof_prb * SRSRAN_NRE * SRSRAN_CP_NSYMB(cp))
#define SRSRAN_SF_LEN_RE(nof_prb, cp) (2 * SRSRAN_SLOT_LEN_RE(nof_prb, cp))
#define SRSRAN_NOF_RE(cell) (2 * SRSRAN_SLOT_LEN_RE(cell.nof_prb, cell.cp))
#define SRSRAN_TA_OFFSET (10e-6)
#define SRSRAN_LTE_TS (1.0f / (15000.0f * 2048.0f))
#define SRSRAN_RE_IDX(nof_prb, symbol_idx, sample_idx) ((symbol_idx) * (nof_prb) * (SRSRAN_NRE) + sample_idx)
#define SRSRAN_N_TA_OFFSET 0
#define SRSRAN_MAX(a, b) ((a) > (b) ? (a) : (b))
#define SRSRAN_MIN(a, b) ((a) < (b) ? (a) : (b))
#define SRSRAN_CEIL(NUM, DEN) (((NUM) + ((DEN)-1)) / (DEN))
#define SRSRAN_FLOOR(NUM, DEN) ((NUM) / (DEN))
#define SRSRAN_ROUND(NUM, DEN) ((uint32_t)round((double)(NUM) / (double)(DEN)))
#define SRSRAN_MEM_ZERO(Q, TYPE, SIZE) do { memset((Q), 0, sizeof(TYPE) * (size_t)(SIZE)); } while (false)

typedef enum { SRSRAN_PHICH_NORM = 0, SRSRAN_PHICH_EXT } srsran_phich_length_t;
typedef enum { SRSRAN_PHICH_R_1_6 = 0, SRSRAN_PHICH_R_1_2, SRSRAN_PHICH_R_1, SRSRAN_PHICH_R_2 } srsran_phich_r_t;
typedef enum { SRSRAN_FDD = 0, SRSRAN_TDD = 1 } srsran_frame_type_t;

typedef struct SRSRAN_API {
  uint32_t             nof_prb;
  uint32_t             nof_ports;
  uint32_t             id;
  srsran_cp_t          cp;
  srsran_phich_length_t phich_length;
  srsran_phich_r_t     phich_resources;
  srsran_frame_type_t  frame_type;
} srsran_cell_t;

typedef struct SRSRAN_API {
  uint32_t sf_config;
  uint32_t ss_config;
  bool     configured;
} srsran_tdd_config_t;

typedef enum SRSRAN_API {
  SRSRAN_TM1 = 0, SRSRAN_TM2, SRSRAN_TM3, SRSRAN_TM4, SRSRAN_TM5, SRSRAN_TM6, SRSRAN_TM7, SRSRAN_TM8, SRSRAN_TMINV
} srsran_tm_t;
typedef enum SRSRAN_API {
  SRSRAN_TXSCHEME_PORT0, SRSRAN_TXSCHEME_DIVERSITY, SRSRAN_TXSCHEME_SPATIALMUX, SRSRAN_TXSCHEME_CDD
} srsran_tx_scheme_t;
typedef enum SRSRAN_API { SRSRAN_MIMO_DECODER_ZF, SRSRAN_MIMO_DECODER_MMSE } srsran_mimo_decoder_t;
typedef enum SRSRAN_API {
  SRSRAN_MOD_BPSK = 0, SRSRAN_MOD_QPSK, SRSRAN_MOD_16QAM, SRSRAN_MOD_64QAM, SRSRAN_MOD_256QAM, SRSRAN_MOD_NITEMS
} srsran_mod_t;
typedef struct SRSRAN_API {
  srsran_tdd_config_t tdd_config;
  uint32_t            tti;
  uint32_t            cfi;
  srsran_sf_t         sf_type;
  uint32_t            non_mbsfn_region;
} srsran_dl_sf_cfg_t;
typedef struct SRSRAN_API {
  srsran_tdd_config_t tdd_config;
  uint32_t            tti;
  bool                shortened;
} srsran_ul_sf_cfg_t;
typedef struct SRSRAN_API { int id; float fd; } srsran_earfcn_t;
enum band_geographical_area { SRSRAN_BAND_GEO_AREA_ALL, SRSRAN_BAND_GEO_AREA_NAR, SRSRAN_BAND_GEO_AREA_APAC, SRSRAN_BAND_GEO_AREA_EMEA,
                              SRSRAN_BAND_GEO_AREA_JAPAN, SRSRAN_BAND_GEO_AREA_CALA, SRSRAN_BAND_GEO_AREA_NA };
typedef enum SRSRAN_API { SRSRAN_MBSFN_SF_ALLOC_ONE = 0, SRSRAN_MBSFN_SF_ALLOC_FOUR } srsran_sf_alloc_t;
typedef enum { SRSRAN_RAT_LTE = 0, SRSRAN_RAT_NR, SRSRAN_RAT_NULL } srsran_rat_t;

SRSRAN_API bool     srsran_cell_isvalid(srsran_cell_t* cell);
SRSRAN_API bool     srsran_cellid_isvalid(uint32_t cell_id);
SRSRAN_API bool     srsran_nofprb_isvalid(uint32_t nof_prb);
SRSRAN_API bool     srsran_sfidx_isvalid(uint32_t sf_idx);
SRSRAN_API bool     srsran_portid_isvalid(uint32_t port_id);
SRSRAN_API bool     srsran_symbol_sz_isvalid(uint32_t symbol_sz);
SRSRAN_API int      srsran_symbol_sz(uint32_t nof_prb);       /* follows srsran_use_standard_symbol_size() */
SRSRAN_API int      srsran_symbol_sz_power2(uint32_t nof_prb);
SRSRAN_API int      srsran_nof_prb(uint32_t symbol_sz);
SRSRAN_API uint32_t srsran_max_cce(uint32_t nof_prb);
SRSRAN_API int      srsran_sampling_freq_hz(uint32_t nof_prb);
SRSRAN_API void     srsran_use_standard_symbol_size(bool enabled);
SRSRAN_API bool     srsran_symbol_size_is_standard();
SRSRAN_API uint32_t srsran_re_x_prb(uint32_t ns, uint32_t symbol, uint32_t nof_ports, uint32_t nof_symbols);
SRSRAN_API uint32_t srsran_mod_bits_x_symbol(srsran_mod_t mod);
SRSRAN_API char*    srsran_mod_string(srsran_mod_t mod);
SRSRAN_API char*    srsran_cp_string(srsran_cp_t cp);
SRSRAN_API const char* srsran_mimotype2str(srsran_tx_scheme_t mimo_type);
SRSRAN_API uint32_t srsran_tti_interval(uint32_t tti1, uint32_t tti2);
SRSRAN_API void     srsran_cell_fprint(FILE* stream, srsran_cell_t* cell, uint32_t sfn);

#define SRSRAN_DEFAULT_MAX_FRAMES_PBCH 500
#define SRSRAN_DEFAULT_MAX_FRAMES_PSS 10
#define SRSRAN_DEFAULT_NOF_VALID_PSS_FRAMES 10

#ifdef __cplusplus
}
#endif
#endif
