// compat_check.cpp -- drives the tier-2 shim (the srsRAN-compatible header tree compat/srsran/..., libltephy_srsran_compat.so) the way the reference's SubframeWorker / DCISearch /
// PDSCH_Decoder drive srsRAN, and dumps what it gets so that tests/test_compat_shim.py can compare it with the tier-1 results:
//   per subframe: srsran_ue_dl_decode_fft_estimate; cfi, snr_db, FNV-1a of sf_symbols / ce, the PDCCH LLRs;
//                 srsran_pdcch_dci_decode for every (location, payload size) given on the command line;
//                 srsran_ue_dl_decode_pdsch for the grants of that subframe read from <grants.bin>.
//   compat_check <iq.cf32> <nof_prb> <nof_ports> <cell_id> <nof_rx> <n_sf> <first_tti> <grants.bin> <out.bin> <nof_bits>...
#include "ltephy_b200.h"
#include "srsran/srsran.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static uint64_t fnv(const void* p, size_t n)
{
  const uint8_t* b = static_cast<const uint8_t*>(p);
  uint64_t       h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
template <class T> static void put(FILE* f, const T& v) { fwrite(&v, sizeof(T), 1, f); }

struct GrantRec {
  uint32_t       sf, rnti;
  ltephy_grant_t g;
};

int main(int argc, char** argv)
{
  if (argc < 11) return 2;
  srsran_cell_t cell{};
  cell.nof_prb = (uint32_t)atoi(argv[2]), cell.nof_ports = (uint32_t)atoi(argv[3]), cell.id = (uint32_t)atoi(argv[4]), cell.cp = SRSRAN_CP_NORM;
  const uint32_t nof_rx = (uint32_t)atoi(argv[5]), n_sf = (uint32_t)atoi(argv[6]), tti0 = (uint32_t)atoi(argv[7]);
  std::vector<uint32_t> sizes;
  for (int i = 10; i < argc; i++) sizes.push_back((uint32_t)atoi(argv[i]));
  FILE* in = fopen(argv[1], "rb");
  FILE* gf = fopen(argv[8], "rb");
  FILE* out = fopen(argv[9], "wb");
  if (!in || !gf || !out) return 1;
  std::vector<GrantRec> grants;
  GrantRec              gr;
  while (fread(&gr, sizeof(gr), 1, gf) == 1) grants.push_back(gr);

  srsran_use_standard_symbol_size(true); // the capture handed to this program is sampled at the standard LTE rate (30.72 Msps at 100 PRB)
  const uint32_t sf_len = SRSRAN_SF_LEN_PRB(cell.nof_prb);
  std::vector<cf_t> buf[2];
  cf_t*             in_buffer[SRSRAN_MAX_PORTS] = {};
  for (uint32_t a = 0; a < nof_rx; a++) buf[a].resize(3 * sf_len), in_buffer[a] = buf[a].data(); // SubframeBuffer over-allocates (src/src/SubframeBuffer.cc:25)
  srsran_ue_dl_t q;
  if (srsran_ue_dl_init(&q, in_buffer, cell.nof_prb, nof_rx) || srsran_ue_dl_set_cell(&q, cell)) {
    fprintf(stderr, "init failed: %s\n", ltephy_last_error());
    return 1;
  }
  const uint32_t g = 14 * 12 * cell.nof_prb;
  for (uint32_t i = 0; i < n_sf; i++) {
    for (uint32_t a = 0; a < nof_rx; a++)
      if (fread(buf[a].data(), sizeof(cf_t), sf_len, in) != sf_len) return 1;
    srsran_dl_sf_cfg_t sf{};
    srsran_ue_dl_cfg_t cfg{};
    sf.tti = tti0 + i;
    if (srsran_ue_dl_decode_fft_estimate(&q, &sf, &cfg) != SRSRAN_SUCCESS) return 1;
    put(out, sf.cfi), put(out, q.chest_res.snr_db), put(out, q.chest_res.noise_estimate), put(out, q.chest_res.cfo);
    uint64_t hs = 0, hc = 0;
    for (uint32_t a = 0; a < nof_rx; a++) hs ^= fnv(q.sf_symbols[a], g * sizeof(cf_t)) * (a + 1);
    for (uint32_t p = 0; p < cell.nof_ports; p++)
      for (uint32_t a = 0; a < nof_rx; a++) hc ^= fnv(q.chest_res.ce[p][a], g * sizeof(cf_t)) * (p * 2 + a + 1);
    put(out, hs), put(out, hc);
    const uint32_t ncce = q.pdcch.nof_cce[sf.cfi - 1];
    put(out, ncce);
    fwrite(q.pdcch.llr, sizeof(float), 72 * ncce, out);
    // every location x size, in srsran_pdcch_ue_locations_all_map order (falcon_pdcch.c:321-356)
    const uint32_t lim = ncce < 84 ? ncce : 84;
    for (int L = 3; L >= 0; L--)
      for (uint32_t k = 0; k < lim >> L; k++)
        for (uint32_t nb : sizes) {
          uint8_t  data[SRSRAN_DCI_MAX_BITS];
          uint16_t crc = 0;
          if (srsran_pdcch_dci_decode(&q.pdcch, &q.pdcch.llr[72 * (k << L)], data, 72u << L, nb, &crc) != SRSRAN_SUCCESS) return 1;
          uint64_t bits = 0;
          for (uint32_t b = 0; b < nb; b++) bits |= (uint64_t)(data[b] & 1u) << (63 - b);
          put(out, crc), put(out, bits);
        }
    for (const GrantRec& r : grants) {
      if (r.sf != i) continue;
      srsran_pdsch_cfg_t pc{};
      pc.rnti = (uint16_t)r.rnti;
      srsran_pdsch_grant_t& s = pc.grant;
      s.tx_scheme = r.g.tx_scheme == LTEPHY_TX_PORT0       ? SRSRAN_TXSCHEME_PORT0
                    : r.g.tx_scheme == LTEPHY_TX_DIVERSITY ? SRSRAN_TXSCHEME_DIVERSITY
                    : r.g.tx_scheme == LTEPHY_TX_CDD       ? SRSRAN_TXSCHEME_CDD
                                                           : SRSRAN_TXSCHEME_SPATIALMUX;
      s.pmi = r.g.pmi, s.nof_re = r.g.nof_re, s.nof_tb = r.g.nof_tb;
      for (int sl = 0; sl < 2; sl++)
        for (uint32_t prb = 0; prb < cell.nof_prb; prb++) s.prb_idx[sl][prb] = (r.g.prb_mask[sl][prb >> 5] >> (prb & 31)) & 1u, s.nof_prb += sl == 0 && s.prb_idx[sl][prb];
      for (int t = 0; t < 2; t++) {
        s.tb[t].enabled = r.g.tb[t].enabled, s.tb[t].tbs = r.g.tb[t].tbs, s.tb[t].rv = r.g.tb[t].rv, s.tb[t].cw_idx = r.g.tb[t].cw_idx;
        s.tb[t].mod = r.g.tb[t].qm == 2 ? SRSRAN_MOD_QPSK : r.g.tb[t].qm == 4 ? SRSRAN_MOD_16QAM : r.g.tb[t].qm == 6 ? SRSRAN_MOD_64QAM : SRSRAN_MOD_256QAM;
      }
      std::vector<uint8_t> pl0(16000), pl1(16000); // srsran_vec_u8_malloc(2000 * 8), DL_Sniffer_PDSCH.cc:47
      srsran_pdsch_res_t   res[SRSRAN_MAX_CODEWORDS]{};
      res[0].payload = pl0.data(), res[1].payload = pl1.data();
      if (srsran_ue_dl_decode_pdsch(&q, &sf, &pc, res) != SRSRAN_SUCCESS) return 1;
      for (int t = 0; t < 2; t++) {
        const uint32_t nby = r.g.tb[t].enabled && r.g.tb[t].tbs > 0 ? (uint32_t)r.g.tb[t].tbs / 8 : 0;
        put(out, (uint8_t)res[t].crc), put(out, nby);
        fwrite(t ? pl1.data() : pl0.data(), 1, nby, out);
      }
    }
  }
  srsran_ue_dl_free(&q);
  fclose(out);
  return 0;
}
