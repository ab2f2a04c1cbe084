// srsran_compat.cpp -- tier-2 shim (include/ltephy_srsran_compat.h): srsRAN / FALCON entry points of the hot path on top of the
// tier-1 C-ABI with a batch of one subframe.  Host code only; built into libltephy_srsran_compat.so, which links libltephy_b200.so.
#include "../../include/ltephy_srsran_compat.h"
#include "../../include/ltephy_b200.h"
#include <cstddef>
#include <cstring>
#include <vector>

namespace {
struct Priv {
  ltephy_t*                  phy = nullptr;
  cf_t*                      in[SRSRAN_MAX_PORTS]{};
  uint32_t                   sf_len = 0, g = 0;
  std::vector<float>         iq;                    // [rx][sf_len] cf32, contiguous staging of the caller's antenna buffers
  std::vector<cf_t>          sym, ce;               // host mirrors handed out through q->sf_symbols / q->chest_res.ce
  std::vector<float>         llr;
  std::vector<ltephy_cand_t> table;                 // T[location][size] of the current subframe
  ltephy_sf_info_t           info{};
  int16_t                    loc_of[4][LTEPHY_MAX_CCE]; // [L][ncce] -> location index for the current CFI
  uint32_t                   loc_cfi = 0;
};
Priv* P(srsran_ue_dl_t* q) { return q ? static_cast<Priv*>(q->ltephy_priv) : nullptr; }
} // namespace

extern "C" int srsran_ue_dl_init(srsran_ue_dl_t* q, cf_t* in_buffer[SRSRAN_MAX_PORTS], uint32_t max_prb, uint32_t nof_rx_antennas)
{
  if (!q || !in_buffer || nof_rx_antennas < 1 || nof_rx_antennas > 2 || max_prb > SRSRAN_MAX_PRB) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(q, 0, sizeof(*q));
  Priv* p = new Priv();
  for (uint32_t a = 0; a < nof_rx_antennas; a++) p->in[a] = in_buffer[a];
  q->nof_rx_antennas = nof_rx_antennas;
  q->ltephy_priv     = p;
  return SRSRAN_SUCCESS;
}
extern "C" int srsran_ue_dl_set_cell(srsran_ue_dl_t* q, srsran_cell_t cell)
{
  Priv* p = P(q);
  if (!p || cell.cp != SRSRAN_CP_NORM) return SRSRAN_ERROR_INVALID_INPUTS;
  if (p->phy) ltephy_destroy(p->phy), p->phy = nullptr;
  ltephy_cfg_t cfg{};
  cfg.nof_prb = cell.nof_prb, cfg.nof_ports = cell.nof_ports, cfg.cell_id = cell.id, cfg.nof_rx = q->nof_rx_antennas;
  cfg.max_subframes = 1, cfg.turbo_max_iter = 8, cfg.flags = 0; // every location is decoded: the caller decides which ones it asks for
  if (ltephy_create(&cfg, &p->phy) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  q->cell   = cell;
  p->sf_len = ltephy_sf_len(p->phy), p->g = 14 * 12 * cell.nof_prb;
  p->iq.assign((size_t)2 * q->nof_rx_antennas * p->sf_len, 0.0f);
  p->sym.assign((size_t)q->nof_rx_antennas * p->g, cf_t{0, 0});
  p->ce.assign((size_t)cell.nof_ports * q->nof_rx_antennas * p->g, cf_t{0, 0});
  p->llr.assign((size_t)72 * LTEPHY_MAX_CCE, 0.0f);
  p->table.assign((size_t)LTEPHY_MAX_LOC * LTEPHY_MAX_SIZES, ltephy_cand_t{});
  for (uint32_t a = 0; a < q->nof_rx_antennas; a++) q->sf_symbols[a] = p->sym.data() + (size_t)a * p->g;
  for (uint32_t pt = 0; pt < cell.nof_ports; pt++)
    for (uint32_t a = 0; a < q->nof_rx_antennas; a++) q->chest_res.ce[pt][a] = p->ce.data() + ((size_t)pt * q->nof_rx_antennas + a) * p->g;
  q->pdcch.llr = p->llr.data();
  for (uint32_t cfi = 1; cfi <= 3; cfi++) q->pdcch.nof_cce[cfi - 1] = ltephy_nof_cce(p->phy, cfi), q->pdcch.nof_regs[cfi - 1] = 9 * q->pdcch.nof_cce[cfi - 1];
  q->pdcch.max_bits = 72 * LTEPHY_MAX_CCE;
  p->loc_cfi        = 0;
  return SRSRAN_SUCCESS;
}
extern "C" void srsran_ue_dl_free(srsran_ue_dl_t* q)
{
  Priv* p = P(q);
  if (!p) return;
  if (p->phy) ltephy_destroy(p->phy);
  delete p;
  memset(q, 0, sizeof(*q));
}
extern "C" int srsran_ue_dl_decode_fft_estimate(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_ue_dl_cfg_t* cfg)
{
  (void)cfg;
  Priv* p = P(q);
  if (!p || !p->phy || !sf) return SRSRAN_ERROR_INVALID_INPUTS;
  for (uint32_t a = 0; a < q->nof_rx_antennas; a++) memcpy(p->iq.data() + (size_t)2 * a * p->sf_len, p->in[a], (size_t)p->sf_len * sizeof(cf_t));
  const uint32_t tti = sf->tti;
  if (ltephy_submit_iq(p->phy, p->iq.data(), &tti, 1) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  if (ltephy_get_phase_a(p->phy, &p->info, p->table.data()) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  if (ltephy_tap(p->phy, LTEPHY_TAP_SYM, p->sym.data(), p->sym.size() * sizeof(cf_t)) || ltephy_tap(p->phy, LTEPHY_TAP_CE, p->ce.data(), p->ce.size() * sizeof(cf_t)) ||
      ltephy_tap(p->phy, LTEPHY_TAP_LLR, p->llr.data(), p->llr.size() * sizeof(float)))
    return SRSRAN_ERROR;
  sf->cfi                      = p->info.cfi;
  q->chest_res.noise_estimate = p->info.noise_avg, q->chest_res.snr_db = p->info.snr_db, q->chest_res.cfo = p->info.cfo, q->chest_res.rsrp = p->info.rsrp_avg;
  if (p->info.cfi >= 1 && p->info.cfi <= 3 && p->loc_cfi != p->info.cfi) { // location index of (L, ncce) in the table of this CFI
    uint16_t ncce[LTEPHY_MAX_LOC];
    uint8_t  L[LTEPHY_MAX_LOC];
    memset(p->loc_of, 0xFF, sizeof(p->loc_of));
    const uint32_t n = ltephy_locations(p->phy, p->info.cfi, ncce, L, LTEPHY_MAX_LOC);
    for (uint32_t i = 0; i < n; i++) p->loc_of[L[i]][ncce[i]] = (int16_t)i;
    p->loc_cfi = p->info.cfi;
  }
  return SRSRAN_SUCCESS;
}
extern "C" int srsran_pdcch_dci_decode(srsran_pdcch_t* pd, float* e, uint8_t* data, uint32_t E, uint32_t nof_bits, uint16_t* crc)
{
  if (!pd || !e || !data || !crc) return SRSRAN_ERROR_INVALID_INPUTS;
  srsran_ue_dl_t* q = reinterpret_cast<srsran_ue_dl_t*>(reinterpret_cast<char*>(pd) - offsetof(srsran_ue_dl_t, pdcch));
  Priv*           p = P(q);
  if (!p || !p->phy || p->loc_cfi == 0) return SRSRAN_ERROR_INVALID_INPUTS;
  const ptrdiff_t d = e - pd->llr;
  uint32_t        L = 0;
  while (L < 4 && (72u << L) != E) L++;
  if (d < 0 || d % 72 || L > 3 || (size_t)d / 72 >= LTEPHY_MAX_CCE) return SRSRAN_ERROR_INVALID_INPUTS;
  const int li = p->loc_of[L][d / 72];
  int       si = -1;
  for (uint32_t f = 0; f < LTEPHY_NOF_FORMATS && si < 0; f++)
    if (ltephy_dci_size(p->phy, f) == nof_bits) si = (int)ltephy_size_index(p->phy, f);
  if (li < 0 || si < 0) return SRSRAN_ERROR_INVALID_INPUTS; // not one of the blind-search locations / payload sizes
  const ltephy_cand_t& c = p->table[(size_t)li * LTEPHY_MAX_SIZES + si];
  for (uint32_t i = 0; i < nof_bits; i++) data[i] = c.valid ? (uint8_t)((c.bits >> (63 - i)) & 1u) : 0;
  *crc = c.valid ? c.rnti : 0;
  return SRSRAN_SUCCESS;
}
extern "C" int srsran_ue_dl_decode_pdsch(srsran_ue_dl_t* q, srsran_dl_sf_cfg_t* sf, srsran_pdsch_cfg_t* cfg, srsran_pdsch_res_t data[SRSRAN_MAX_CODEWORDS])
{
  (void)sf;
  Priv* p = P(q);
  if (!p || !p->phy || !cfg || !data) return SRSRAN_ERROR_INVALID_INPUTS;
  static const uint8_t qm_of[5] = {1, 2, 4, 6, 8};
  const srsran_pdsch_grant_t& s = cfg->grant;
  ltephy_grant_t              g{};
  g.sf = 0, g.rnti = cfg->rnti, g.nof_tb = (uint8_t)s.nof_tb, g.nof_re = s.nof_re, g.pmi = s.pmi;
  g.tx_scheme = s.tx_scheme == SRSRAN_TXSCHEME_PORT0       ? LTEPHY_TX_PORT0
                : s.tx_scheme == SRSRAN_TXSCHEME_DIVERSITY ? LTEPHY_TX_DIVERSITY
                : s.tx_scheme == SRSRAN_TXSCHEME_CDD       ? LTEPHY_TX_CDD
                                                           : LTEPHY_TX_SPATIALMUX;
  for (int sl = 0; sl < 2; sl++)
    for (uint32_t prb = 0; prb < q->cell.nof_prb; prb++)
      if (s.prb_idx[sl][prb]) g.prb_mask[sl][prb >> 5] |= 1u << (prb & 31);
  for (int t = 0; t < SRSRAN_MAX_CODEWORDS; t++) {
    g.tb[t].enabled = s.tb[t].enabled, g.tb[t].tbs = s.tb[t].enabled ? s.tb[t].tbs : 0, g.tb[t].rv = (uint8_t)s.tb[t].rv;
    g.tb[t].qm = (unsigned)s.tb[t].mod < 5 ? qm_of[s.tb[t].mod] : 0;
    data[t].crc = false, data[t].avg_iterations_block = 0.0f;
  }
  if (s.tb[0].enabled && s.tb[1].enabled) // srsran_ra_tb_t.cw_idx (dl_sniffer_pdsch.c:24): DCI 2/2A swap flag
    g.tb[0].cw_idx = (uint8_t)(s.tb[0].cw_idx & 1u), g.tb[1].cw_idx = (uint8_t)(s.tb[1].cw_idx & 1u);
  if (ltephy_submit_grants(p->phy, &g, 1) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  ltephy_tb_result_t   r[2]{};
  std::vector<uint8_t> pl(2 * 16000);
  if (ltephy_get_phase_b(p->phy, r, pl.data(), pl.size()) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  for (int t = 0; t < SRSRAN_MAX_CODEWORDS; t++) {
    if (!r[t].payload_len) continue;
    data[t].crc = r[t].crc != 0, data[t].avg_iterations_block = (float)r[t].avg_iters;
    if (data[t].payload) memcpy(data[t].payload, pl.data() + r[t].payload_off, r[t].payload_len);
  }
  return SRSRAN_SUCCESS;
}
