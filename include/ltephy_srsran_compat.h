/*
 * ltephy_srsran_compat.h -- tier-2 of the boundary: the srsRAN / FALCON names the reference calls on
 * the hot path, implemented on top of the batched tier-1 C-ABI with a batch of one subframe.  SOURCE-compatible for the
 * fields the reference touches, not binary-compatible with srsRAN's structs: srsRAN's headers are not
 * in the reference tree, so the structs below carry exactly the members the reference reads or writes, under srsRAN's names.
 * A compatibility path (one synchronous GPU round trip per call), not the performance path -- that is ltephy_decode_subframes.
 *
 *   srsran_ue_dl_init / _set_cell / _free        src/src/SubframeWorker.cc:52,102,94
 *   srsran_ue_dl_decode_fft_estimate             src/src/DCISearch.cc:562   (fills sf_symbols, chest_res, sf->cfi, pdcch.llr)
 *   srsran_pdcch_dci_decode                      lib/src/phy/falcon_phch/falcon_pdcch.c:142 (table lookup: phase A decoded every candidate)
 *   srsran_ue_dl_decode_pdsch                    src/src/DL_Sniffer_PDSCH.cc:257,520,708,997,1110,1207
 */
#ifndef LTEPHY_SRSRAN_COMPAT_H
#define LTEPHY_SRSRAN_COMPAT_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SRSRAN_SUCCESS 0
#define SRSRAN_ERROR -1
#define SRSRAN_ERROR_INVALID_INPUTS -2
#define SRSRAN_MAX_PORTS 4
#define SRSRAN_MAX_CODEWORDS 2
#define SRSRAN_MAX_PRB 110
#define SRSRAN_DCI_MAX_BITS 128

typedef struct {
  float re, im;
} cf_t; /* layout of float _Complex */

typedef enum { SRSRAN_CP_NORM = 0, SRSRAN_CP_EXT } srsran_cp_t;
typedef struct {
  uint32_t    nof_prb;
  uint32_t    nof_ports;
  uint32_t    id;
  srsran_cp_t cp;
  int         phich_length, phich_resources, frame_type;
} srsran_cell_t;

typedef struct {
  uint32_t tti;
  uint32_t cfi; /* written by srsran_ue_dl_decode_fft_estimate (DCISearch.cc:569) */
  int      sf_type;
} srsran_dl_sf_cfg_t;

/* the chest configuration the reference sets (SubframeWorker.cc:379-389) is the one the library implements; the fields are accepted and ignored */
typedef struct {
  struct {
    float filter_coef[2];
    int   filter_type, noise_alg, estimator_alg;
    bool  cfo_estimate_enable, rsrp_neighbour;
    uint32_t cfo_estimate_sf_mask;
  } chest_cfg;
} srsran_ue_dl_cfg_t;

typedef struct {
  cf_t* ce[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS]; /* [port][rx antenna], 14*12*nof_prb each */
  float noise_estimate, noise_estimate_dbm;
  float snr_db;
  float snr_ant_port_db[SRSRAN_MAX_PORTS][SRSRAN_MAX_PORTS];
  float rsrp, rsrp_dbm;
  float cfo;
} srsran_chest_dl_res_t;

typedef struct {
  float*   llr;          /* 72 floats per CCE (falcon_pdcch.c:138,381) */
  uint32_t nof_cce[3];   /* per CFI (falcon_pdcch.c:36-37) */
  uint32_t nof_regs[3];
  uint32_t max_bits;
} srsran_pdcch_t;

typedef struct {
  srsran_cell_t         cell;
  uint32_t              nof_rx_antennas;
  cf_t*                 sf_symbols[SRSRAN_MAX_PORTS]; /* per rx antenna (DCISearch.cc:565) */
  srsran_chest_dl_res_t chest_res;
  srsran_pdcch_t        pdcch;
  void*                 ltephy_priv;
} srsran_ue_dl_t;

typedef enum { SRSRAN_MOD_BPSK = 0, SRSRAN_MOD_QPSK, SRSRAN_MOD_16QAM, SRSRAN_MOD_64QAM, SRSRAN_MOD_256QAM } srsran_mod_t;
typedef enum { SRSRAN_TXSCHEME_PORT0 = 0, SRSRAN_TXSCHEME_DIVERSITY, SRSRAN_TXSCHEME_SPATIALMUX, SRSRAN_TXSCHEME_CDD } srsran_tx_scheme_t;
typedef struct {
  srsran_mod_t mod;
  int          tbs;
  int          rv;
  uint32_t     nof_bits;
  uint32_t     cw_idx;
  bool         enabled;
  uint32_t     mcs_idx;
} srsran_ra_tb_t;
typedef struct {
  srsran_tx_scheme_t tx_scheme;
  uint32_t           pmi;
  bool               prb_idx[2][SRSRAN_MAX_PRB];
  uint32_t           nof_prb;
  uint32_t           nof_re;
  uint32_t           nof_symb_slot[2];
  srsran_ra_tb_t     tb[SRSRAN_MAX_CODEWORDS];
  uint32_t           nof_tb;
  uint32_t           nof_layers;
} srsran_pdsch_grant_t;
typedef struct {
  srsran_pdsch_grant_t grant;
  uint16_t             rnti;
  uint32_t             max_nof_iterations; /* 0 in the reference (zero-initialised cfg): the library then uses its configured maximum */
  int                  decoder_type;
  float                p_a;
  bool                 csi_enable, meas_evm_en;
  struct {
    void* tx[SRSRAN_MAX_CODEWORDS];
    void* rx[SRSRAN_MAX_CODEWORDS]; /* softbuffers: accepted, unused (no HARQ combining yet) */
  } softbuffers;
} srsran_pdsch_cfg_t;
typedef struct {
  uint8_t* payload; /* caller-allocated, tbs/8 bytes are written (DL_Sniffer_PDSCH.cc:47) */
  bool     crc;
  float    avg_iterations_block;
  float    evm;
} srsran_pdsch_res_t;

int  srsran_ue_dl_init(srsran_ue_dl_t* q, cf_t* in_buffer[SRSRAN_MAX_PORTS], uint32_t max_prb, uint32_t nof_rx_antennas);
int  srsran_ue_dl_set_cell(srsran_ue_dl_t* q, srsran_cell_t cell);
void srsran_ue_dl_free(srsran_ue_dl_t* q);
int  srsran_ue_dl_decode_fft_estimate(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_ue_dl_cfg_t* cfg);
/* e must point into q->llr at the first CCE of the location (e = &q->llr[72 * ncce], falcon_pdcch.c:138); E = 72 << L */
int  srsran_pdcch_dci_decode(srsran_pdcch_t* q, float* e, uint8_t* data, uint32_t E, uint32_t nof_bits, uint16_t* crc);
int  srsran_ue_dl_decode_pdsch(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_pdsch_cfg_t* cfg, srsran_pdsch_res_t data[SRSRAN_MAX_CODEWORDS]);

#ifdef __cplusplus
}
#endif
#endif
