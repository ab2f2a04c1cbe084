// ref_walk.cc -- ORACLE (test infrastructure): a C wrapper around the reference's OWN blind search and DCI -> grant code,
// compiled UNMODIFIED from /root/reference by oracle/build_ref.sh against the srsRAN-compatible header tree compat/srsran:
//   src/src/{DCISearch,DCICollection,MetaFormats,SubframeInfo,SubframePower,ULSchedule,HARQ,MCSTracking,PhyCommon,...}.cc
//   lib/src/phy/falcon_phch/{falcon_pdcch,falcon_dci,dl_sniffer_pdsch,ul_sniffer_pusch}.c, lib/src/phy/falcon_ue/falcon_ue_dl.c
//   lib/src/util/{RNTIManager,Histogram,Interval}.cc
// The per-candidate srsran_pdcch_dci_decode (falcon_pdcch.c:142) is served by libltephy_srsran_compat from a candidate table
// handed in with ltephy_compat_inject, so the reference's walk and the product's walk (ltephy_search_batch) can be compared on
// IDENTICAL tables -- from the GPU in the -m gpu tests, from the CPU oracle otherwise.
// Set-up mirrors the reference's own: RNTIManager as PhyCommon builds it (src/src/PhyCommon.cc:11), evergreen / forbidden ranges as
// LTESniffer_Core seeds them after the MIB (src/src/LTESniffer_Core.cc:398-417), DCIMetaFormats with split ratio 0.99 refreshed
// every 500 subframes (Settings.h:55-56, LTESniffer_Core.cc:434), one SubframeInfo + DCISearch per subframe as
// SubframeWorker::work does (src/src/SubframeWorker.cc:142-207).
#include "include/DCISearch.h"
#include "include/DCICollection.h"
#include "include/SubframeInfoConsumer.h"
#include "include/MCSTracking.h"
#include "include/HARQ.h"
#include "include/ULSchedule.h"
#include "falcon/util/RNTIManager.h"
#include "ltephy_compat_ext.h"
#include <atomic>
#include <cstring>
#include <unistd.h>

extern "C" {
#include "falcon/phy/falcon_phch/dl_sniffer_pdsch.h"
#include "falcon/phy/falcon_phch/ul_sniffer_pusch.h"

typedef struct {
  uint16_t rnti;
  uint8_t  format, L;
  uint16_t ncce, nof_bits;
  uint32_t histval;
  uint8_t  bits[64];
  // what the reference's dl_sniffer_ra_dl_dci_to_grant + dl_sniffer_config_mimo make of it, for both MCS tables (0: 64QAM, 1: 256QAM)
  int32_t  grant_ret[2]; // 0 ok; < 0: the reference zeroed the RNTI / MIMO configuration refused
  uint32_t nof_prb, nof_re[2], nof_tb[2], tx_scheme[2], pmi[2], nof_layers[2];
  int32_t  tbs[2][2];
  uint8_t  qm[2][2], rv[2][2], tb_en[2][2], cw_idx[2][2];
  uint8_t  prb_mask[2][110];
  // format 0: the reference's srsran_ra_ul_dci_to_grant path (0: Table 8.6.1-1) and ulsniffer_ra_ul_dci_to_grant_256 (1)
  uint32_t ul_L_prb, ul_n_prb[2], ul_n_dmrs;
  int32_t  ul_tbs[2];
  uint8_t  ul_qm[2];
} refwalk_dci_t;
typedef struct {
  uint32_t nof_decoded_locations, nof_cce, nof_missed_cce, nof_subframes, nof_locations;
} refwalk_stats_t;
}

namespace {
uint8_t qm_of(srsran_mod_t m) { return m == SRSRAN_MOD_QPSK ? 2 : m == SRSRAN_MOD_16QAM ? 4 : m == SRSRAN_MOD_64QAM ? 6 : m == SRSRAN_MOD_256QAM ? 8 : 1; }
uint32_t scheme_of(srsran_tx_scheme_t t) { return t == SRSRAN_TXSCHEME_PORT0 ? 0 : t == SRSRAN_TXSCHEME_DIVERSITY ? 1 : t == SRSRAN_TXSCHEME_CDD ? 2 : 3; } // LTEPHY_TX_*
} // namespace

struct refwalk {
  srsran_cell_t        cell{};
  cf_t*                in[SRSRAN_MAX_PORTS]{};
  srsran_ue_dl_t       q{};
  falcon_ue_dl_t*      fq = nullptr;
  RNTIManager*         rm = nullptr;
  DCIMetaFormats*      mf = nullptr;
  std::atomic<float>   cfo{0.0f};
  MCSTracking*         mcs = nullptr;
  HARQ*                harq = nullptr;
  ULSchedule*          ulsche = nullptr;
  srsran_ue_dl_cfg_t   ue_dl_cfg{};
  uint32_t             sf_cnt = 0, update_interval = 500;
  bool                 shortcut = true;
  DCIBlindSearchStats  stats;
  FILE*                trace = nullptr; // DCIToFile target (refwalk_set_trace)
};

extern "C" {

refwalk* refwalk_create_phich(uint32_t nof_prb, uint32_t nof_ports, uint32_t cell_id, uint32_t nof_rx, uint32_t threshold, uint32_t phich_resources, uint32_t phich_length);
refwalk* refwalk_create(uint32_t nof_prb, uint32_t nof_ports, uint32_t cell_id, uint32_t nof_rx, uint32_t threshold)
{
  return refwalk_create_phich(nof_prb, nof_ports, cell_id, nof_rx, threshold, SRSRAN_PHICH_R_1_6, SRSRAN_PHICH_NORM); // file mode, LTESniffer_Core.cc:242-247
}
// the cell as the live mode gets it from the MIB (srsran_pbch_mib_unpack, LTESniffer_Core.cc:389)
refwalk* refwalk_create_phich(uint32_t nof_prb, uint32_t nof_ports, uint32_t cell_id, uint32_t nof_rx, uint32_t threshold, uint32_t phich_resources, uint32_t phich_length)
{
  refwalk* w           = new refwalk();
  w->cell.nof_prb      = nof_prb, w->cell.nof_ports = nof_ports, w->cell.id = cell_id, w->cell.cp = SRSRAN_CP_NORM;
  w->cell.phich_length = (srsran_phich_length_t)phich_length, w->cell.phich_resources = (srsran_phich_r_t)phich_resources;
  static cf_t dummy[2][4];
  w->in[0] = dummy[0], w->in[1] = dummy[1];
  if (srsran_ue_dl_init(&w->q, w->in, nof_prb, nof_rx) || srsran_ue_dl_set_cell(&w->q, w->cell)) {
    delete w;
    return nullptr;
  }
  w->fq = new falcon_ue_dl_t();
  memset((void*)w->fq, 0, sizeof(falcon_ue_dl_t));
  w->fq->q = &w->q;
  w->rm    = new RNTIManager(nof_falcon_ue_all_formats, RNTI_PER_SUBFRAME, threshold); // PhyCommon.cc:11
  int idx  = falcon_dci_index_of_format_in_list(SRSRAN_DCI_FORMAT1A, falcon_ue_all_formats, nof_falcon_ue_all_formats);
  if (idx > -1) w->rm->addEvergreen(SRSRAN_RARNTI_START, SRSRAN_RARNTI_END, (uint32_t)idx), w->rm->addEvergreen(SRSRAN_PRNTI, SRSRAN_SIRNTI, (uint32_t)idx);
  idx = falcon_dci_index_of_format_in_list(SRSRAN_DCI_FORMAT1C, falcon_ue_all_formats, nof_falcon_ue_all_formats);
  if (idx > -1) w->rm->addEvergreen(SRSRAN_RARNTI_START, SRSRAN_RARNTI_END, (uint32_t)idx), w->rm->addEvergreen(SRSRAN_PRNTI, SRSRAN_SIRNTI, (uint32_t)idx);
  for (uint32_t f = 0; f < nof_falcon_ue_all_formats; f++) w->rm->addForbidden(0x0, 0x0, f);
  w->mf = new DCIMetaFormats(nof_falcon_ue_all_formats, 0.99); // DEFAULT_DCI_FORMAT_SPLIT_RATIO
  { // MCSTracking opens ./mcs_statistic.csv in its constructor: keep that out of the caller's directory
    char cwd[4096];
    const bool moved = getcwd(cwd, sizeof(cwd)) && chdir("/tmp") == 0;
    w->mcs           = new MCSTracking(DL_SNIFFER_MCS_MODE_BOTH, 0, false, DL_MODE, 0, w->cfo);
    if (moved && chdir(cwd) != 0) {
    }
  }
  w->harq   = new HARQ();
  w->ulsche = new ULSchedule(0, nullptr, false);
  w->ue_dl_cfg.cfg.dci.multiple_csi_request_enabled = false, w->ue_dl_cfg.cfg.dci.cif_enabled = false; // SubframeWorker.cc:394-399
  return w;
}
void refwalk_destroy(refwalk* w)
{
  if (!w) return;
  if (w->trace) fclose(w->trace);
  delete w->ulsche;
  delete w->harq;
  delete w->mcs;
  delete w->mf;
  delete w->rm;
  srsran_ue_dl_free(&w->q);
  delete w->fq;
  delete w;
}
// every following subframe's accepted DCIs are written to `path` by the reference's DCIToFile (NULL: stop)
int refwalk_set_trace(refwalk* w, const char* path)
{
  if (w->trace) fclose(w->trace);
  w->trace = path ? fopen(path, "w") : nullptr;
  return (path && !w->trace) ? -1 : 0;
}
void refwalk_config(refwalk* w, int shortcut, int skip_secondary, uint32_t update_interval)
{
  w->shortcut = shortcut != 0, w->update_interval = update_interval;
  w->mf->setSkipSecondaryMetaFormats(skip_secondary != 0);
}
void refwalk_activate(refwalk* w, uint16_t rnti, uint32_t format_idx, int reason)
{
  w->rm->activateAndRefresh(rnti, format_idx, (ActivationReason)reason);
}
uint32_t refwalk_frequency(refwalk* w, uint16_t rnti, uint32_t format_idx) { return w->rm->getFrequency(rnti, format_idx); }

// One subframe, exactly SubframeWorker::work's DL part: update_formats on the refresh tick, SubframeInfo, DCISearch::search.
// info / table / llr: the phase-A result for this subframe (ltephy_compat_inject).
int refwalk_subframe(refwalk* w, const ltephy_sf_info_t* info, const ltephy_cand_t* table, const float* llr, refwalk_dci_t* out, uint32_t max_out,
                     uint32_t* n_out)
{
  if (w->update_interval && (w->sf_cnt % w->update_interval) == 0) w->mf->update_formats();
  w->sf_cnt++;
  if (ltephy_compat_inject(&w->q, info, table, llr)) return -1;
  srsran_dl_sf_cfg_t sf{};
  sf.tti = info->tti, sf.sf_type = SRSRAN_SF_NORM;
  SubframeInfo subframeInfo(w->cell, DL_SNIFFER_MCS_MODE_BOTH, w->mcs, 0, w->harq, w->ulsche);
  DCISearch    search(*w->fq, *w->mf, *w->rm, subframeInfo, info->tti % 10, info->tti / 10, &sf, &w->ue_dl_cfg);
  search.setShortcutDiscovery(w->shortcut);
  search.search();
  w->stats += search.getStats();
  if (w->trace) { // the reference's own trace writer, DCIToFile::printDCICollection (SubframeInfoConsumer.cc:66-138)
    DCIToFile tf(w->trace);
    tf.consumeDCICollection(subframeInfo);
    fflush(w->trace);
  }
  // accepted DCIs: the reference keeps DL and UL in separate containers; both in acceptance order
  uint32_t n = 0;
  DCICollection& col = subframeInfo.getDCICollection();
  for (auto& d : col.getDLSnifferDCI_DL()) {
    if (n >= max_out) break;
    refwalk_dci_t& o = out[n++];
    memset(&o, 0, sizeof(o));
    o.rnti = d.rnti, o.format = (uint8_t)d.format, o.L = (uint8_t)d.location.L, o.ncce = (uint16_t)d.location.ncce, o.nof_bits = (uint16_t)d.nof_bits, o.histval = d.histval;
    // payload back from the hex string the reference stores (sprint_hex, falcon_dci.c:37-58)
    for (uint32_t i = 0; i < d.nof_bits && i < 64; i++) {
      const char   c = d.hex[i / 4];
      const uint32_t v = (uint32_t)(c <= '9' ? c - '0' : (c | 0x20) - 'a' + 10);
      o.bits[i]        = (uint8_t)((v >> (3 - (i % 4))) & 1u);
    }
    srsran_pdsch_grant_t* gr[2] = {d.ran_pdsch_grant.get(), d.ran_pdsch_grant_256.get()};
    for (int t = 0; t < 2; t++) {
      // the DCI -> grant conversion ran inside DCICollection::addCandidate (srsran_dci_msg_to_trace_timestamp, mode BOTH -> both
      // tables); a failure there zeroes ran_dci_dl->rnti (falcon_dci.c:286-305).  Re-run it per table to know WHICH table failed.
      srsran_dci_dl_t      dci = *d.ran_dci_dl;
      srsran_pdsch_grant_t g;
      dci.rnti = d.rnti;
      srsran_dl_sf_cfg_t sf2 = sf;
      int                r   = dl_sniffer_ra_dl_dci_to_grant(&w->cell, &sf2, t == 1, &dci, &g);
      if (r == SRSRAN_SUCCESS) r = dl_sniffer_config_mimo(&w->cell, d.format, &dci, &g); // src/src/DL_Sniffer_PDSCH.cc:920
      o.grant_ret[t] = r;
      if (r != SRSRAN_SUCCESS) continue;
      (void)gr;
      o.nof_prb = g.nof_prb, o.nof_re[t] = g.nof_re, o.nof_tb[t] = g.nof_tb, o.tx_scheme[t] = scheme_of(g.tx_scheme), o.pmi[t] = g.pmi, o.nof_layers[t] = g.nof_layers;
      for (int i = 0; i < 2; i++)
        o.tbs[t][i] = g.tb[i].tbs, o.qm[t][i] = qm_of(g.tb[i].mod), o.rv[t][i] = (uint8_t)g.tb[i].rv, o.tb_en[t][i] = g.tb[i].enabled, o.cw_idx[t][i] = (uint8_t)g.tb[i].cw_idx;
      if (t == 0)
        for (int sl = 0; sl < 2; sl++)
          for (uint32_t p = 0; p < w->cell.nof_prb; p++) o.prb_mask[sl][p] = g.prb_idx[sl][p];
    }
  }
  for (auto& d : col.getULSnifferDCI_UL()) {
    if (n >= max_out) break;
    refwalk_dci_t& o = out[n++];
    memset(&o, 0, sizeof(o));
    o.rnti = d.rnti, o.format = (uint8_t)d.format, o.L = (uint8_t)d.location.L, o.ncce = (uint16_t)d.location.ncce, o.nof_bits = (uint16_t)d.nof_bits, o.histval = d.histval;
    for (uint32_t i = 0; i < d.nof_bits && i < 64; i++) {
      const char   c = d.hex[i / 4];
      const uint32_t v = (uint32_t)(c <= '9' ? c - '0' : (c | 0x20) - 'a' + 10);
      o.bits[i]        = (uint8_t)((v >> (3 - (i % 4))) & 1u);
    }
    o.grant_ret[0] = d.ran_ul_dci->rnti ? 0 : -1, o.grant_ret[1] = o.grant_ret[0];
    o.ul_L_prb = d.ran_ul_grant->L_prb, o.ul_n_prb[0] = d.ran_ul_grant->n_prb[0], o.ul_n_prb[1] = d.ran_ul_grant->n_prb[1], o.ul_n_dmrs = d.ran_ul_dci->n_dmrs;
    o.ul_tbs[0] = d.ran_ul_grant->tb.tbs, o.ul_qm[0] = qm_of(d.ran_ul_grant->tb.mod);
    o.ul_tbs[1] = d.ran_ul_grant_256->tb.tbs, o.ul_qm[1] = qm_of(d.ran_ul_grant_256->tb.mod);
  }
  *n_out = n;
  return 0;
}
void refwalk_get_stats(refwalk* w, refwalk_stats_t* st)
{
  st->nof_decoded_locations = w->stats.nof_decoded_locations, st->nof_cce = w->stats.nof_cce, st->nof_missed_cce = w->stats.nof_missed_cce;
  st->nof_subframes = w->stats.nof_subframes, st->nof_locations = w->stats.nof_locations;
}

// ---- single reference functions, for stage-by-stage parity tests -----------------------------------------------------------
uint32_t refwalk_validate_location(uint32_t nof_cce, uint32_t ncce, uint32_t l, uint32_t nsubframe, uint16_t rnti)
{
  return srsran_pdcch_validate_location(nof_cce, ncce, l, nsubframe, rnti); // falcon_pdcch.c:223-250
}
// srsran_pdcch_ue_locations_all_map (falcon_pdcch.c:321-356) + srsran_pdcch_cce_avg_llr_power (:595-620) over the LLRs of a subframe:
// ncce[], L[] of every location, sufficient_power per location, mean |LLR| per CCE; returns the number of locations
uint32_t refwalk_locations(refwalk* w, uint32_t cfi, const float* llr, uint16_t* ncce, uint8_t* L, uint8_t* sufficient_power, float* cce_power)
{
  falcon_dci_location_t            loc[MAX_CANDIDATES_BLIND];
  falcon_cce_to_dci_location_map_t map[MAX_NUM_OF_CCE];
  memset(loc, 0, sizeof(loc)), memset(map, 0, sizeof(map));
  memcpy(w->q.pdcch.llr, llr, sizeof(float) * 72 * w->q.pdcch.nof_cce[cfi - 1]);
  const uint32_t n = srsran_pdcch_ue_locations_all_map(&w->q.pdcch, loc, MAX_CANDIDATES_BLIND, map, MAX_NUM_OF_CCE, 0, cfi);
  srsran_pdcch_cce_avg_llr_power(&w->q.pdcch, cfi, map, MAX_NUM_OF_CCE);
  const uint32_t lim = SRSRAN_MIN(w->q.pdcch.nof_cce[cfi - 1], MAX_NUM_OF_CCE);
  for (uint32_t c = 0; c < lim; c++) cce_power[c] = map[c].power;
  // the sufficient-power rule of recursive_blind_dci_search (DCISearch.cc:473-489)
  for (uint32_t i = 0; i < n; i++) {
    ncce[i] = (uint16_t)loc[i].ncce, L[i] = (uint8_t)loc[i].L;
    uint8_t ok = 1;
    for (uint32_t c = loc[i].ncce; c < loc[i].ncce + (1u << loc[i].L) && c < lim; c++)
      if (map[c].power < PWR_THR) ok = 0;
    sufficient_power[i] = ok;
  }
  return n;
}

// ---- the reference's HARQ bookkeeping (src/src/HARQ.cc, compiled unmodified): what ltesniffer_b200/csrc/harq.cpp restates -------------------------
void* refharq_create()
{
  HARQ* q = new HARQ();
  q->init_HARQ(DL_SNIFFER_HARQ_MODE_ON);
  return q;
}
void refharq_destroy(void* q) { delete static_cast<HARQ*>(q); }
int  refharq_size(void* q) { return static_cast<HARQ*>(q)->harqBufferSize(); }
// HARQ::is_retransmission with the grant PDSCH_Decoder::decode_dl_mode builds at DL_Sniffer_PDSCH.cc:946-953; *buffer = getHARQBuffer (identity of the store)
int refharq_is_retransmission(void* q, uint16_t rnti, int pid, int tid, int ndi, int rv, int tbs, uint32_t tti, const void** buffer)
{
  dl_sniffer_harq_grant_t g = {};
  g.last_decoded = false, g.ndi = ndi != 0, g.ndi_present = true, g.rv = rv, g.tbs = tbs, g.is_first_transmission = false;
  HARQ*     h  = static_cast<HARQ*>(q);
  const int st = h->is_retransmission(rnti, pid, tid, g, tti / 10, tti % 10);
  if (buffer) *buffer = (st == DL_SNIFFER_NEW_TX || st == DL_SNIFFER_RE_TX) ? (const void*)h->getHARQBuffer(rnti, pid, tid) : nullptr;
  return st;
}
// HARQ::updateHARQRNTI with the grant built at DL_Sniffer_PDSCH.cc:1008-1014
void refharq_update(void* q, uint16_t rnti, int pid, int tid, int ndi, int rv, int tbs, uint32_t tti, int decoded)
{
  dl_sniffer_harq_grant_t g = {};
  g.last_decoded = decoded != 0, g.ndi = ndi != 0, g.ndi_present = true, g.rv = rv, g.tbs = tbs, g.is_first_transmission = false;
  static_cast<HARQ*>(q)->updateHARQRNTI(rnti, pid, tid, tti / 10, tti % 10, g);
}

int refharq_last_tbs(void* q, uint16_t rnti, int pid, int tid) { return static_cast<HARQ*>(q)->getlastTbs(rnti, pid, tid); }

// SubframePower::computePower (src/src/SubframePower.cc:18-47, called at DCISearch.cc:565): per-PRB power of antenna 0 in dB
void refwalk_rb_power(uint32_t nof_prb, const cf_t* sf_symbols, float* out_db)
{
  srsran_cell_t cell = {};
  cell.nof_prb       = nof_prb;
  SubframePower p(cell);
  p.computePower(sf_symbols);
  const std::vector<float>& v = p.getRBPowerDL();
  for (uint32_t i = 0; i < nof_prb; i++) out_db[i] = v[i];
}

// One DCI through the reference's grant conversion, outside any walk (randomised parity tests): srsran_dci_msg_unpack_pdsch, then per MCS table
// dl_sniffer_ra_dl_dci_to_grant (lib/src/phy/falcon_phch/dl_sniffer_pdsch.c:95-132) + dl_sniffer_config_mimo (:255-276), as falcon_dci.c:271-310
// and DL_Sniffer_PDSCH.cc:920 do.  bits: one per byte.  Returns -1 if the unpack fails.
int refgrant_dl(refwalk* w, uint32_t format, uint16_t rnti, const uint8_t* bits, uint32_t nof_bits, uint32_t tti, uint32_t cfi, refwalk_dci_t* out)
{
  srsran_dci_msg_t msg;
  memset(&msg, 0, sizeof(msg));
  msg.format = falcon_ue_all_formats[format], msg.rnti = rnti, msg.nof_bits = nof_bits;
  memcpy(msg.payload, bits, nof_bits);
  srsran_dl_sf_cfg_t sf{};
  sf.tti = tti, sf.cfi = cfi, sf.sf_type = SRSRAN_SF_NORM;
  srsran_dci_dl_t dci;
  memset(&dci, 0, sizeof(dci));
  memset(out, 0, sizeof(*out));
  if (srsran_dci_msg_unpack_pdsch(&w->cell, &sf, &w->ue_dl_cfg.cfg.dci, &msg, &dci)) return -1;
  out->rnti = rnti, out->format = (uint8_t)format, out->nof_bits = (uint16_t)nof_bits;
  for (int t = 0; t < 2; t++) {
    srsran_dci_dl_t      d2 = dci;
    srsran_pdsch_grant_t g;
    memset(&g, 0, sizeof(g));
    srsran_dl_sf_cfg_t sf2 = sf;
    int                r   = dl_sniffer_ra_dl_dci_to_grant(&w->cell, &sf2, t == 1, &d2, &g);
    if (r == SRSRAN_SUCCESS) r = dl_sniffer_config_mimo(&w->cell, msg.format, &d2, &g);
    out->grant_ret[t] = r;
    if (r != SRSRAN_SUCCESS) continue;
    out->nof_prb = g.nof_prb, out->nof_re[t] = g.nof_re, out->nof_tb[t] = g.nof_tb, out->tx_scheme[t] = scheme_of(g.tx_scheme), out->pmi[t] = g.pmi, out->nof_layers[t] = g.nof_layers;
    for (int i = 0; i < 2; i++)
      out->tbs[t][i] = g.tb[i].tbs, out->qm[t][i] = qm_of(g.tb[i].mod), out->rv[t][i] = (uint8_t)g.tb[i].rv, out->tb_en[t][i] = g.tb[i].enabled, out->cw_idx[t][i] = (uint8_t)g.tb[i].cw_idx;
    if (t == 0)
      for (int sl = 0; sl < 2; sl++)
        for (uint32_t p = 0; p < w->cell.nof_prb; p++) out->prb_mask[sl][p] = g.prb_idx[sl][p];
  }
  return 0;
}
// One format-0 DCI through the reference's own uplink grant conversion with an arbitrary pusch-HoppingOffset: srsran_dci_msg_unpack_pusch, then
// ul_sniffer_ra_ul_dci_to_grant (Table 8.6.1-1, out table 0) and ulsniffer_ra_ul_dci_to_grant_256 (Table 8.6.1-3, out table 1), both on top of
// ul_sniffer_ra_ul_grant_to_grant_prb_allocation (lib/src/phy/falcon_phch/ul_sniffer_pusch.c:20-87,138-245).  Returns -1 when the unpack refuses.
int refgrant_ul(refwalk* w, uint16_t rnti, const uint8_t* bits, uint32_t nof_bits, uint32_t tti, uint32_t n_rb_ho, refwalk_dci_t* out)
{
  srsran_dci_msg_t msg;
  memset(&msg, 0, sizeof(msg));
  msg.format = SRSRAN_DCI_FORMAT0, msg.rnti = rnti, msg.nof_bits = nof_bits;
  memcpy(msg.payload, bits, nof_bits);
  srsran_dl_sf_cfg_t sf{};
  sf.tti = tti, sf.sf_type = SRSRAN_SF_NORM;
  srsran_dci_ul_t dci;
  memset(out, 0, sizeof(*out));
  if (srsran_dci_msg_unpack_pusch(&w->cell, &sf, &w->ue_dl_cfg.cfg.dci, &msg, &dci)) return -1;
  out->rnti = rnti, out->format = 0, out->nof_bits = (uint16_t)nof_bits, out->ul_n_dmrs = dci.n_dmrs;
  for (int t = 0; t < 2; t++) {
    srsran_dci_ul_t      d2 = dci;
    srsran_pusch_grant_t g;
    memset(&g, 0, sizeof(g));
    srsran_ul_sf_cfg_t ul_sf;
    memset(&ul_sf, 0, sizeof(ul_sf));
    ul_sf.tti = tti;
    srsran_pusch_hopping_cfg_t hop;
    memset(&hop, 0, sizeof(hop));
    hop.n_rb_ho = n_rb_ho;
    out->grant_ret[t] = t == 0 ? ul_sniffer_ra_ul_dci_to_grant(&w->cell, &ul_sf, &hop, &d2, &g) : ulsniffer_ra_ul_dci_to_grant_256(&w->cell, &ul_sf, &hop, &d2, &g);
    if (out->grant_ret[t]) continue;
    out->ul_L_prb = g.L_prb, out->ul_n_prb[0] = g.n_prb[0], out->ul_n_prb[1] = g.n_prb[1];
    out->ul_tbs[t] = g.tb.tbs, out->ul_qm[t] = qm_of(g.tb.mod);
    out->rv[t][0] = (uint8_t)g.tb.rv, out->tx_scheme[t] = (uint8_t)g.freq_hopping;
  }
  return 0;
}
// One 20-bit RAR grant (one bit per byte, as srsran::rar_subh::get_sched_grant hands it over) through the reference's own conversion:
// ul_sniffer_dci_rar_unpack + ul_sniffer_dci_rar_to_ul_dci (lib/src/phy/falcon_phch/falcon_dci.c:648-684) + ul_sniffer_ra_ul_dci_to_grant, as at
// src/src/DL_Sniffer_PDSCH.cc:646-658.  out: grant_ret[0], ul_L_prb, ul_n_prb[2], ul_tbs[0], ul_qm[0], rv[0][0]
int refgrant_rar(refwalk* w, const uint8_t* grant_bits, uint16_t t_crnti, uint32_t tti, uint32_t n_rb_ho, refwalk_dci_t* out)
{
  uint8_t bits[SRSRAN_RAR_GRANT_LEN];
  memcpy(bits, grant_bits, SRSRAN_RAR_GRANT_LEN);
  srsran_dci_rar_grant_t rar;
  memset(&rar, 0, sizeof(rar));
  ul_sniffer_dci_rar_unpack(bits, &rar);
  srsran_dci_ul_t dci;
  ul_sniffer_dci_rar_to_ul_dci(&w->cell, &rar, &dci);
  dci.rnti = t_crnti, dci.format = SRSRAN_DCI_FORMAT_RAR;
  srsran_ul_sf_cfg_t ul_sf;
  memset(&ul_sf, 0, sizeof(ul_sf));
  ul_sf.tti = tti;
  srsran_pusch_hopping_cfg_t hop;
  memset(&hop, 0, sizeof(hop));
  hop.n_rb_ho = n_rb_ho;
  srsran_pusch_grant_t g;
  memset(&g, 0, sizeof(g));
  memset(out, 0, sizeof(*out));
  out->rnti = t_crnti;
  out->grant_ret[0] = ul_sniffer_ra_ul_dci_to_grant(&w->cell, &ul_sf, &hop, &dci, &g);
  if (out->grant_ret[0]) return 0;
  out->ul_L_prb = g.L_prb, out->ul_n_prb[0] = g.n_prb[0], out->ul_n_prb[1] = g.n_prb[1], out->ul_tbs[0] = g.tb.tbs, out->ul_qm[0] = qm_of(g.tb.mod);
  out->rv[0][0] = (uint8_t)g.tb.rv, out->ul_n_dmrs = dci.n_dmrs;
  return 0;
}
int refcqi_no_subbands(int nof_prb) { return ul_sniffer_cqi_hl_get_no_subbands(nof_prb); }
}
