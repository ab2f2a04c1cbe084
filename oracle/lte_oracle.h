/*
 * lte_oracle.h -- CPU ORACLE for the per-subframe LTE PHY decode path (TEST INFRASTRUCTURE).
 *
 * PARITY UNPINNED: the reference (SysSec-KAIST/LTESniffer) contains no tests, fixtures or golden
 * vectors for this path (lib/test/CMakeLists.txt is empty) and delegates the arithmetic to srsRAN,
 * which is fetched at configure time and is absent from /root/reference.
 * This file restates the algorithm chain the reference calls --
 *   srsran_ue_dl_decode_fft_estimate (src/src/DCISearch.cc:562): OFDM rx, CRS channel estimate,
 *       PCFICH, PDCCH LLR extraction
 *   srsran_pdcch_dci_decode (lib/src/phy/falcon_phch/falcon_pdcch.c:142): conv rate-dematch,
 *       tail-biting Viterbi (3 concatenated copies, uint8 quantisation with gain 32), CRC16
 *   srsran_ue_dl_decode_pdsch (src/src/DL_Sniffer_PDSCH.cc:997): RE gather, predecoding, soft demod,
 *       descrambling, turbo rate-dematch, max-log-MAP turbo decode, CRC24B/CRC24A
 * -- from 3GPP TS 36.211/212/213 and the in-tree glue.  The signal arithmetic therefore stays PARITY UNPINNED by the reference; what
 * stands in for it: (i) ground truth from the synthetic eNB in sim/ (CRC-self-validating), (ii) 3GPP structural checks of the tables and the
 * transport block sizes of the reference's example captures (tests/test_tables.py), (iii) the primitives this receiver SHARES with that
 * transmitter (sim/lte_common.c: CRCs, Gold sequences, CRS, encoders, rate matching, constellations, DMRS, interleavers, OFDM numerology)
 * checked against third-party code and other-domain formulations (tests/test_independent_primitives.py), so that a shared mistake cannot
 * cancel out.  The DECISION logic around it (search walk, grants, RAR, HARQ bookkeeping, trace lines) is pinned on the reference's own code
 * compiled into oracle/_ref (tests/test_reference_code.py, tests/test_rnti_manager.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Every float expression below has a fixed evaluation order (build with
 * -ffp-contract=off) so that the CUDA path can be compared bit-for-bit.
 */
#ifndef LTE_ORACLE_H
#define LTE_ORACLE_H
#include "../sim/lte_common.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lteo lteo_t;

typedef struct {
  float noise[LTE_MAX_PORTS][LTE_MAX_ANT]; /* per-path noise estimate (linear) */
  float rsrp[LTE_MAX_PORTS][LTE_MAX_ANT];
  float noise_avg, rsrp_avg;
  float cfo_re, cfo_im; /* sum conj(ls[l=0]) * ls[l=7], port 0, antenna 0 */
  float snr_db, cfo;    /* host-side: 10 log10f(rsrp/noise), atan2f(im,re)/(2 pi 7.5) */
} lteo_chest_res_t;

lteo_t* lteo_create(const lte_cell_t* cell);
void    lteo_destroy(lteo_t* q);
uint32_t lteo_nof_cce(lteo_t* q, uint32_t cfi);

/* deterministic float reduction used by every stage (32-lane strided partials + halving tree) */
float lteo_det_sum(const float* x, uint32_t n);

/* K1: one antenna-subframe: iq[sf_len] -> sym[14*12*nof_prb] */
void lteo_ofdm_rx(lteo_t* q, const cf_t* iq, cf_t* sym);
/* geometry of the control region: grid indices of the PDCCH REs in CCE order for this CFI, and of the 16 PCFICH REs (may be NULL); returns nof_cce */
uint32_t lteo_pdcch_re_index(lteo_t* q, uint32_t cfi, uint16_t* idx, uint16_t* pcfich_idx);
/* K2: sym[ant] -> ce[port*nof_rx+ant][14*nsc] */
void lteo_chest(lteo_t* q, uint32_t sf_idx, const cf_t* const* sym, cf_t* const* ce, lteo_chest_res_t* res);
/* K10: per-PRB mean RE power of antenna 0 (linear) -- SubframePower::computePower, src/src/SubframePower.cc:18-58 */
void lteo_rb_power(lteo_t* q, const cf_t* sym0, float* pwr);
/* K3: returns cfi 1..3; corr[3] receives the three correlations */
uint32_t lteo_pcfich_decode(lteo_t* q, uint32_t sf_idx, const cf_t* const* sym, const cf_t* const* ce, float* corr);
/* K4: llr[nof_cce*72]; returns nof_cce */
uint32_t lteo_pdcch_extract_llr(lteo_t* q, uint32_t sf_idx, uint32_t cfi, const cf_t* const* sym, const cf_t* const* ce, float* llr);
/* per-CCE mean |LLR| (double accumulate) -- srsran_pdcch_cce_avg_llr_power, falcon_pdcch.c:595-620 */
void lteo_cce_power(const float* llr, uint32_t nof_cce, float* pwr);
/* K5: returns 0 ok, -1 if all-zero input (no decode) */
int lteo_dci_decode(const float* e, uint32_t E, uint32_t nof_bits, uint8_t* bits, uint16_t* crc_rem);

/* K6: gather + predecode + soft demod + descramble for one grant; llr[cw] receives nof_re*Qm int16 */
int lteo_pdsch_llr(lteo_t* q, uint32_t sf_idx, uint32_t cfi, uint16_t rnti, const lte_dl_grant_t* g, const cf_t* const* sym,
                   const cf_t* const* ce, int16_t* const* llr, cf_t* const* eq_out);
/* K7+K8 for one transport block: e[G] int16 -> payload bytes; returns crc ok (1/0), <0 on error.
 * iters_out (optional) receives per-code-block iteration counts. */
int lteo_dlsch_decode_harq(const int16_t* e, uint32_t G, uint32_t tbs, uint32_t rv, uint32_t Qm, uint32_t NL, uint32_t max_iter, int early_stop,
                           uint8_t* payload, uint32_t* iters_out, int16_t* soft, int combine);
int lteo_dlsch_decode(const int16_t* e, uint32_t G, uint32_t tbs, uint32_t rv, uint32_t Qm, uint32_t NL, uint32_t max_iter,
                      int early_stop, uint8_t* payload, uint32_t* iters_out);
/* K7 alone: rate-dematch one code block into conditioned (sys, p1, p2) streams of K+4 */
void lteo_rm_turbo_rx(const int16_t* e, uint32_t E, uint32_t K, uint32_t F, uint32_t rv, uint32_t Qm, int16_t* d /* 3*(K+4) */);
/* PBCH of subframe 0 (srsran_ue_mib_decode at src/src/LTESniffer_Core.cc:386): equalise the 240 resource elements with the cell's port count, QPSK soft
 * bits, and for each of the four positions q of this frame inside the 40 ms period: descramble, rate-dematch 480 -> 120, tail-biting Viterbi, CRC16;
 * the CRC remainder must be the mask of 1, 2 or 4 antenna ports.  Returns 1 (found: mib[24] bits, *nof_ports, *q = SFN mod 4) or 0. */
int lteo_pbch_decode(lteo_t* q, const cf_t* const* sym, const cf_t* const* ce, uint8_t* mib, uint32_t* nof_ports, uint32_t* frame_q);
/* e^{-j 2 pi f n / (15000 fft)} applied to one subframe of samples, n restarting at 0 (srsran_cfo_correct on the file samples, srsran_ue_sync file mode) */
void lteo_cfo_correct(lteo_t* q, float cfo_hz, const cf_t* in, cf_t* out);
/* HARQ soft combining (reference src/src/HARQ.cc:71-151, DL_Sniffer_PDSCH.cc:955-985): per code block LTEO_HARQ_CB_STRIDE int16 accumulators */
#define LTEO_HARQ_CB_STRIDE 18448
void lteo_rm_turbo_rx_harq(const int16_t* e, uint32_t E, uint32_t K, uint32_t F, uint32_t rv, uint32_t Qm, int16_t* d, int16_t* soft, int combine);
/* K8 alone: decode one code block from conditioned streams; bits[K]; returns iterations run.
 * crc_type: 0 none (run max_iter), 1 CRC24A, 2 CRC24B */
uint32_t lteo_turbo_decode(const int16_t* d, uint32_t K, uint32_t max_iter, int crc_type, uint8_t* bits, int* crc_ok);
/* full PDSCH for one grant */
/* soft[t] (NULL: no HARQ buffer): C x LTEO_HARQ_CB_STRIDE int16 of transport block t; combine[t] 0 = new transmission (buffer reset), 1 = add */
int lteo_pdsch_decode_harq(lteo_t* q, uint32_t sf_idx, uint32_t cfi, uint16_t rnti, const lte_dl_grant_t* g, const cf_t* const* sym,
                           const cf_t* const* ce, uint32_t max_iter, uint8_t* const* payload, int* crc_ok, int16_t* const* soft, const int* combine);
int lteo_pdsch_decode(lteo_t* q, uint32_t sf_idx, uint32_t cfi, uint16_t rnti, const lte_dl_grant_t* g, const cf_t* const* sym,
                      const cf_t* const* ce, uint32_t max_iter, uint8_t* const* payload, int* crc_ok);

/* whole phase A of one subframe (OFDM rx, chest, PCFICH, PDCCH LLR); returns nof_cce */
int lteo_phase_a(lteo_t* q, const cf_t* iq, uint32_t sf_idx, cf_t* sym, cf_t* ce, float* llr, lteo_chest_res_t* res, uint32_t* cfi_out);

#ifdef __cplusplus
}
#endif
#endif

/* ================================================================== uplink (PUSCH) -- ORACLE
 * restates srsran_enb_ul_fft + srsran_chest_ul_estimate_pusch + srsran_pusch_decode as called from
 * PUSCH_Decoder::decode / decode_run (src/src/UL_Sniffer_PUSCH.cc:250-263,389-392) for grants without UCI,
 * without hopping, L_prb >= 3. */
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  float noise, rsrp, snr_db, ta_us;
} lteo_ul_chest_t;
/* K1-UL: iq[sf_len] -> sym[14*nsc] with the 7.5 kHz shift removed */
void lteo_ul_ofdm(lteo_t* q, const cf_t* iq, cf_t* sym);
/* test accessor: the unscaled M-point inverse DFT of the transform de-precoder (M = 12 L_prb, L_prb = 2^a 3^b 5^c); 0 ok */
int lteo_idft(uint32_t M, const cf_t* in, cf_t* out);
/* K9: one PUSCH grant; llr_out (optional) receives the nof_bits descrambled, de-interleaved int16 soft bits */
int lteo_pusch_decode(lteo_t* q, const lte_ul_cfg_t* ucfg, uint32_t sf_idx, const lte_ul_grant_t* g, const cf_t* sym, uint32_t max_iter,
                      uint8_t* payload, int* crc_ok, lteo_ul_chest_t* chest, int16_t* llr_out);
#ifdef __cplusplus
}
#endif
