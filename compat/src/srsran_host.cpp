// srsran_host.cpp -- host-only part of libltephy_srsran_compat: the srsRAN functions the reference's PHY-facing sources call that
// involve no signal processing (DCI sizes and unpacking, resource allocation, transport-block sizes, search-space candidates, bit
// helpers).  Declared in compat/srsran/...; each function names the srsRAN 21.10 function it stands for and the reference call
// site that needs it.  Built on the same host code as the tier-1 library (csrc/lte_host.cpp, csrc/host_search.cpp).
#include "srsran/srsran.h"
#include "../../include/lte_tables.h"
#include "../../include/ltephy_b200.h"
#include "../../include/ltephy_search.h"
#include "../../ltesniffer_b200/csrc/lte_host.hpp"
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <sys/time.h>

extern "C" {

int srsran_verbose = 0;

// srsran/phy/utils/debug.c: tdata[0] = tdata[2] - tdata[1]
void get_time_interval(struct timeval* tdata)
{
  tdata[0].tv_sec  = tdata[2].tv_sec - tdata[1].tv_sec;
  tdata[0].tv_usec = tdata[2].tv_usec - tdata[1].tv_usec;
  if (tdata[0].tv_usec < 0) {
    tdata[0].tv_sec--;
    tdata[0].tv_usec += 1000000;
  }
}

// ---------------------------------------------------------------------------------------------------- bit helpers (bit.c)
void srsran_bit_unpack(uint32_t value, uint8_t** bits, int nof_bits)
{
  for (int i = 0; i < nof_bits; i++) (*bits)[i] = (uint8_t)((value >> (nof_bits - i - 1)) & 0x1);
  *bits += nof_bits;
}
void srsran_bit_unpack_l(uint64_t value, uint8_t** bits, int nof_bits)
{
  for (int i = 0; i < nof_bits; i++) (*bits)[i] = (uint8_t)((value >> (nof_bits - i - 1)) & 0x1);
  *bits += nof_bits;
}
uint32_t srsran_bit_pack(uint8_t** bits, int nof_bits)
{
  uint32_t value = 0;
  for (int i = 0; i < nof_bits; i++) value |= (uint32_t)(*bits)[i] << (nof_bits - i - 1);
  *bits += nof_bits;
  return value;
}
uint64_t srsran_bit_pack_l(uint8_t** bits, int nof_bits)
{
  uint64_t value = 0;
  for (int i = 0; i < nof_bits; i++) value |= (uint64_t)(*bits)[i] << (nof_bits - i - 1);
  *bits += nof_bits;
  return value;
}
void srsran_bit_pack_vector(uint8_t* unpacked, uint8_t* packed, int nof_bits)
{
  const int nbytes = nof_bits / 8;
  for (int i = 0; i < nbytes; i++) packed[i] = (uint8_t)srsran_bit_pack(&unpacked, 8);
  if (nof_bits % 8) packed[nbytes] = (uint8_t)(srsran_bit_pack(&unpacked, nof_bits % 8) << (8 - (nof_bits % 8)));
}
void srsran_bit_unpack_vector(const uint8_t* packed, uint8_t* unpacked, int nof_bits)
{
  const int nbytes = nof_bits / 8;
  for (int i = 0; i < nbytes; i++) srsran_bit_unpack(packed[i], &unpacked, 8);
  if (nof_bits % 8) srsran_bit_unpack(packed[nbytes] >> (8 - nof_bits % 8), &unpacked, nof_bits % 8);
}
void srsran_bit_fprint(FILE* stream, uint8_t* bits, int nof_bits)
{
  fprintf(stream, "[");
  for (int i = 0; i < nof_bits - 1; i++) fprintf(stream, "%d,", bits[i]);
  fprintf(stream, "%d]\n", nof_bits > 0 ? bits[nof_bits - 1] : 0);
}
uint32_t srsran_bit_diff(const uint8_t* x, const uint8_t* y, int nbits)
{
  uint32_t errors = 0;
  for (int i = 0; i < nbits; i++) errors += x[i] != y[i];
  return errors;
}
uint32_t srsran_bit_count(uint32_t n) { return (uint32_t)__builtin_popcount(n); }

// ---------------------------------------------------------------------------------------------------- vector helpers (vector.c)
void*  srsran_vec_malloc(uint32_t size) { return aligned_alloc(64, ((size_t)size + 63) & ~(size_t)63); }
cf_t*  srsran_vec_cf_malloc(uint32_t n) { return (cf_t*)srsran_vec_malloc((uint32_t)sizeof(cf_t) * n); }
float* srsran_vec_f_malloc(uint32_t n) { return (float*)srsran_vec_malloc((uint32_t)sizeof(float) * n); }
void   srsran_vec_cf_zero(cf_t* ptr, uint32_t n) { memset((void*)ptr, 0, sizeof(cf_t) * n); }
void   srsran_vec_f_zero(float* ptr, uint32_t n) { memset(ptr, 0, sizeof(float) * n); }
void   srsran_vec_cf_copy(cf_t* dst, const cf_t* src, uint32_t len) { memcpy((void*)dst, (const void*)src, sizeof(cf_t) * len); }
float  srsran_vec_avg_power_cf(const cf_t* x, const uint32_t len)
{
  const float* f = reinterpret_cast<const float*>(x);
  float        s = 0.0f;
  for (uint32_t i = 0; i < len; i++) s += f[2 * i] * f[2 * i] + f[2 * i + 1] * f[2 * i + 1];
  return len ? s / (float)len : 0.0f;
}
float srsran_vec_acc_ff(const float* x, const uint32_t len)
{
  float s = 0.0f;
  for (uint32_t i = 0; i < len; i++) s += x[i];
  return s;
}
void srsran_vec_fprint_f(FILE* stream, const float* x, const uint32_t len)
{
  fprintf(stream, "[");
  for (uint32_t i = 0; i < len; i++) fprintf(stream, "%+2.5f, ", x[i]);
  fprintf(stream, "];\n");
}
void srsran_vec_fprint_b(FILE* stream, const uint8_t* x, const uint32_t len)
{
  fprintf(stream, "[");
  for (uint32_t i = 0; i < len; i++) fprintf(stream, "%d, ", x[i]);
  fprintf(stream, "];\n");
}
void srsran_vec_sprint_hex(char* str, const uint32_t max_str_len, uint8_t* x, const uint32_t len)
{
  uint32_t nbytes = len / 8, n = 0;
  if ((nbytes + (len % 8 ? 1 : 0)) * 3 + 2 >= max_str_len) return;
  n += (uint32_t)sprintf(&str[n], "[");
  for (uint32_t i = 0; i < nbytes; i++) {
    uint8_t* p = &x[8 * i];
    n += (uint32_t)sprintf(&str[n], "%02x ", srsran_bit_pack(&p, 8));
  }
  if (len % 8) {
    uint8_t* p = &x[8 * nbytes];
    n += (uint32_t)sprintf(&str[n], "%02x ", srsran_bit_pack(&p, (int)(len % 8)) << (8 - (len % 8)));
  }
  n += (uint32_t)sprintf(&str[n], "]");
  str[max_str_len - 1] = 0;
}
void srsran_vec_fprint_hex(FILE* stream, uint8_t* x, const uint32_t len)
{
  char tmp[1024];
  tmp[0] = 0;
  srsran_vec_sprint_hex(tmp, sizeof(tmp), x, len);
  fprintf(stream, "%s\n", tmp);
}

// ---------------------------------------------------------------------------------------------------- phy_common.c
static bool g_standard_rates =
#ifdef FORCE_STANDARD_RATE
    true;
#else
    false;
#endif
void srsran_use_standard_symbol_size(bool enabled) { g_standard_rates = enabled; }
bool srsran_symbol_size_is_standard() { return g_standard_rates; }
// FFT size of one OFDM symbol: 128 * 2^k with standard rates; the default srsRAN build (no FORCE_STANDARD_RATE, the reference's
// CMakeLists.txt:289-292 only sets it with -DUSE_LTE_RATES) uses 3/4 of that above 15 PRB: 384 / 768 / 1024 / 1536
int srsran_symbol_sz_power2(uint32_t nof_prb) { return nof_prb == 0 ? -1 : nof_prb <= 6 ? 128 : nof_prb <= 15 ? 256 : nof_prb <= 25 ? 512 : nof_prb <= 52 ? 1024 : nof_prb <= 110 ? 2048 : -1; }
int srsran_symbol_sz(uint32_t nof_prb)
{
  if (nof_prb == 0) return SRSRAN_ERROR;
  if (g_standard_rates) return srsran_symbol_sz_power2(nof_prb);
  return nof_prb <= 6 ? 128 : nof_prb <= 15 ? 256 : nof_prb <= 25 ? 384 : nof_prb <= 52 ? 768 : nof_prb <= 79 ? 1024 : nof_prb <= 110 ? 1536 : SRSRAN_ERROR;
}
int  srsran_sampling_freq_hz(uint32_t nof_prb) { const int n = srsran_symbol_sz(nof_prb); return n < 0 ? SRSRAN_ERROR : 15000 * n; }
bool srsran_symbol_sz_isvalid(uint32_t n) { return n == 128 || n == 256 || n == 384 || n == 512 || n == 768 || n == 1024 || n == 1536 || n == 2048; }
int  srsran_nof_prb(uint32_t symbol_sz)
{
  switch (symbol_sz) {
    case 128: return 6;
    case 256: return 15;
    case 384: case 512: return 25;
    case 768: return 50;
    case 1024: return g_standard_rates ? 50 : 75;
    case 1536: return 100;
    case 2048: return 100;
    default: return SRSRAN_ERROR;
  }
}
bool srsran_cellid_isvalid(uint32_t cell_id) { return cell_id < 504; }
bool srsran_nofprb_isvalid(uint32_t nof_prb) { return nof_prb == 1 || (nof_prb >= 6 && nof_prb <= SRSRAN_MAX_PRB); }
bool srsran_cell_isvalid(srsran_cell_t* cell) { return cell && srsran_cellid_isvalid(cell->id) && cell->nof_ports >= 1 && cell->nof_ports <= SRSRAN_MAX_PORTS && srsran_nofprb_isvalid(cell->nof_prb); }
bool srsran_sfidx_isvalid(uint32_t sf_idx) { return sf_idx <= SRSRAN_NOF_SF_X_FRAME; }
bool srsran_portid_isvalid(uint32_t port_id) { return port_id <= SRSRAN_MAX_PORTS; }
uint32_t srsran_mod_bits_x_symbol(srsran_mod_t mod)
{
  switch (mod) {
    case SRSRAN_MOD_BPSK: return 1;
    case SRSRAN_MOD_QPSK: return 2;
    case SRSRAN_MOD_16QAM: return 4;
    case SRSRAN_MOD_64QAM: return 6;
    case SRSRAN_MOD_256QAM: return 8;
    default: return 0;
  }
}
char* srsran_mod_string(srsran_mod_t mod)
{
  switch (mod) {
    case SRSRAN_MOD_BPSK: return (char*)"BPSK";
    case SRSRAN_MOD_QPSK: return (char*)"QPSK";
    case SRSRAN_MOD_16QAM: return (char*)"16QAM";
    case SRSRAN_MOD_64QAM: return (char*)"64QAM";
    case SRSRAN_MOD_256QAM: return (char*)"256QAM";
    default: return (char*)"N/A";
  }
}
char*       srsran_cp_string(srsran_cp_t cp) { return cp == SRSRAN_CP_NORM ? (char*)"Normal  " : (char*)"Extended"; }
const char* srsran_mimotype2str(srsran_tx_scheme_t t)
{
  switch (t) {
    case SRSRAN_TXSCHEME_PORT0: return "p0";
    case SRSRAN_TXSCHEME_DIVERSITY: return "div";
    case SRSRAN_TXSCHEME_SPATIALMUX: return "mux";
    case SRSRAN_TXSCHEME_CDD: return "cdd";
    default: return "N/A";
  }
}
uint32_t srsran_tti_interval(uint32_t tti1, uint32_t tti2) { return tti1 >= tti2 ? tti1 - tti2 : 10240 - tti2 + tti1; }
void     srsran_cell_fprint(FILE* stream, srsran_cell_t* cell, uint32_t sfn)
{
  fprintf(stream, " - Type:            %s\n", cell->frame_type == SRSRAN_FDD ? "FDD" : "TDD");
  fprintf(stream, " - PCI:             %d\n", cell->id);
  fprintf(stream, " - Nof ports:       %d\n", cell->nof_ports);
  fprintf(stream, " - CP:              %s\n", srsran_cp_string(cell->cp));
  fprintf(stream, " - PRB:             %d\n", cell->nof_prb);
  fprintf(stream, " - SFN:             %d\n", sfn);
}
// data REs of a PRB without CRS in one OFDM symbol / with CRS (srsran_re_x_prb)
uint32_t srsran_re_x_prb(uint32_t ns, uint32_t symbol, uint32_t nof_ports, uint32_t nof_symbols)
{
  (void)ns;
  if (symbol == 0 || symbol == nof_symbols - 3) return nof_ports == 1 ? 10 : 8; // CRS of ports 0 / 1
  if (symbol == 1 && nof_ports == 4) return 8;
  return 12;
}
uint32_t srsran_max_cce(uint32_t nof_prb) { return nof_prb <= 6 ? 7 : nof_prb <= 15 ? 20 : nof_prb <= 25 ? 21 : nof_prb <= 50 ? 43 : nof_prb <= 75 ? 65 : 87; }

// ---------------------------------------------------------------------------------------------------- dci.c
static ltehost::Cell host_cell(const srsran_cell_t* c) { return ltehost::Cell{c->nof_prb, c->nof_ports, c->id, 1, (uint32_t)c->phich_resources, (uint32_t)c->phich_length}; }
// srsran_dci_format_sizeof (falcon_pdcch.c:133): payload bits of a format for this cell (FDD, no carrier indicator, no SRS request)
uint32_t srsran_dci_format_sizeof(const srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_dci_cfg_t* cfg, srsran_dci_format_t format)
{
  (void)sf, (void)cfg;
  if (!cell || (unsigned)format >= LTEPHY_NOF_FORMATS) return 0;
  static std::mutex                                             mtx;
  static std::map<std::pair<uint32_t, uint32_t>, ltehost::SizeTable> cache;
  std::lock_guard<std::mutex>                                   lk(mtx);
  auto                                                          key = std::make_pair(cell->nof_prb, cell->nof_ports);
  auto                                                          it  = cache.find(key);
  if (it == cache.end()) it = cache.emplace(key, ltehost::dci_size_table(host_cell(cell))).first;
  return it->second.sizes[it->second.index_of[format]];
}
uint32_t srsran_dci_format_max_tb(srsran_dci_format_t format) { return (format == SRSRAN_DCI_FORMAT2 || format == SRSRAN_DCI_FORMAT2A || format == SRSRAN_DCI_FORMAT2B) ? 2 : 1; }
char*    srsran_dci_format_string(srsran_dci_format_t format)
{
  static const char* n[] = {"Format0 ", "Format1 ", "Format1A", "Format1B", "Format1C", "Format1D", "Format2 ", "Format2A", "Format2B", "FormatRAR"};
  return (unsigned)format <= SRSRAN_DCI_FORMAT_RAR ? (char*)n[format] : (char*)"N/A";
}
char* srsran_dci_format_string_short(srsran_dci_format_t format)
{
  static const char* n[] = {"0", "1", "1A", "1B", "1C", "1D", "2", "2A", "2B", "RAR"};
  return (unsigned)format <= SRSRAN_DCI_FORMAT_RAR ? (char*)n[format] : (char*)"N/A";
}
bool srsran_dci_location_isvalid(srsran_dci_location_t* c) { return c && c->L <= 3 && c->ncce <= 87; }
int  srsran_dci_location_set(srsran_dci_location_t* c, uint32_t L, uint32_t nCCE)
{
  if (L > 3 || nCCE > 87) return SRSRAN_ERROR;
  c->L = L, c->ncce = nCCE;
  return SRSRAN_SUCCESS;
}
void srsran_dci_cfg_set_common_ss(srsran_dci_cfg_t* cfg) { cfg->is_not_ue_ss = true; }

static uint64_t payload_word(const srsran_dci_msg_t* msg)
{
  uint64_t v = 0;
  for (uint32_t i = 0; i < msg->nof_bits && i < 64; i++) v |= (uint64_t)(msg->payload[i] & 1u) << (63 - i);
  return v;
}
// srsran_dci_msg_unpack_pdsch (falcon_dci.c:271): payload -> srsran_dci_dl_t.  Formats 1, 1A, 1C, 2, 2A as LTESniffer decodes them
// (1B / 1D / 2B are refused here; the reference refuses them one step later in dl_sniffer_config_mimo_type, dl_sniffer_pdsch.c:134-178).
int srsran_dci_msg_unpack_pdsch(srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_dci_cfg_t* cfg, srsran_dci_msg_t* msg, srsran_dci_dl_t* dci)
{
  (void)sf, (void)cfg;
  if (!cell || !msg || !dci) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(dci, 0, sizeof(*dci));
  dci->rnti = msg->rnti, dci->location = msg->location, dci->format = msg->format;
  ltehost::DlDciFields f;
  if (ltehost_unpack_dl_dci(host_cell(cell), (uint32_t)msg->format, msg->rnti, payload_word(msg), f) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  dci->alloc_type = (srsran_ra_type_t)f.alloc;
  if (f.alloc == 0)
    dci->type0_alloc.rbg_bitmask = f.rbg_mask;
  else if (f.alloc == 1)
    dci->type1_alloc.vrb_bitmask = f.t1_mask, dci->type1_alloc.rbg_subset = f.t1_subset, dci->type1_alloc.shift = f.t1_shift != 0;
  else {
    dci->type2_alloc.riv     = f.riv;
    dci->type2_alloc.n_prb1a = f.n_prb1a == 2 ? srsran_ra_type2_t::SRSRAN_RA_TYPE2_NPRB1A_2 : srsran_ra_type2_t::SRSRAN_RA_TYPE2_NPRB1A_3;
    dci->type2_alloc.n_gap   = f.ngap2 ? srsran_ra_type2_t::SRSRAN_RA_TYPE2_NG2 : srsran_ra_type2_t::SRSRAN_RA_TYPE2_NG1;
    dci->type2_alloc.mode    = f.dist ? srsran_ra_type2_t::SRSRAN_RA_TYPE2_DIST : srsran_ra_type2_t::SRSRAN_RA_TYPE2_LOC;
  }
  for (int i = 0; i < 2; i++) dci->tb[i].mcs_idx = f.mcs[i], dci->tb[i].rv = f.rv[i], dci->tb[i].ndi = f.ndi[i] != 0, dci->tb[i].cw_idx = 0;
  const bool two_tb_format = msg->format >= SRSRAN_DCI_FORMAT2;
  if (!two_tb_format) SRSRAN_DCI_TB_DISABLE(dci->tb[1]);
  dci->tb_cw_swap = f.tb_cw_swap != 0;
  if (two_tb_format && SRSRAN_DCI_IS_TB_EN(dci->tb[0]) && SRSRAN_DCI_IS_TB_EN(dci->tb[1])) // 36.212 Table 5.3.3.1.5-1
    dci->tb[0].cw_idx = dci->tb_cw_swap ? 1 : 0, dci->tb[1].cw_idx = dci->tb_cw_swap ? 0 : 1;
  dci->pinfo = f.pinfo, dci->pid = f.harq_pid, dci->tpc_pucch = f.tpc;
  return SRSRAN_SUCCESS;
}
// srsran_dci_msg_unpack_pusch (falcon_dci.c:208): DCI format 0
int srsran_dci_msg_unpack_pusch(srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_dci_cfg_t* cfg, srsran_dci_msg_t* msg, srsran_dci_ul_t* dci)
{
  (void)sf, (void)cfg;
  if (!cell || !msg || !dci) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(dci, 0, sizeof(*dci));
  dci->rnti = msg->rnti, dci->location = msg->location, dci->format = msg->format;
  const uint32_t N = cell->nof_prb;
  uint32_t       rivb = 0;
  while ((1u << rivb) < N * (N + 1) / 2) rivb++;
  uint8_t* y = msg->payload;
  if (srsran_bit_pack(&y, 1) != 0) return SRSRAN_ERROR; // format 0 / 1A flag
  const uint32_t hop = srsran_bit_pack(&y, 1);
  uint32_t       n_ul_hop = 0;
  if (hop) { // 36.213 Table 8.4-1 / 8.4-2: 1 hopping bit below 50 PRB, 2 from 50 PRB on
    n_ul_hop          = N < 50 ? 1 : 2;
    const uint32_t hb = srsran_bit_pack(&y, (int)n_ul_hop);
    if (n_ul_hop == 1)
      dci->freq_hop_fl = hb == 0 ? SRSRAN_RA_PUSCH_HOP_HALF : SRSRAN_RA_PUSCH_HOP_TYPE2;
    else
      dci->freq_hop_fl = (srsran_ra_pusch_hop_t)hb; // 00 quarter, 01 -quarter, 10 half, 11 type 2
  } else
    dci->freq_hop_fl = SRSRAN_RA_PUSCH_HOP_DISABLED;
  dci->type2_alloc.riv = srsran_bit_pack(&y, (int)(rivb - n_ul_hop));
  dci->tb.mcs_idx      = srsran_bit_pack(&y, 5);
  dci->tb.ndi          = srsran_bit_pack(&y, 1) != 0;
  dci->tb.rv           = 0;
  dci->tpc_pusch       = (uint8_t)srsran_bit_pack(&y, 2);
  dci->n_dmrs          = srsran_bit_pack(&y, 3);
  dci->cqi_request     = srsran_bit_pack(&y, 1) != 0;
  return SRSRAN_SUCCESS;
}
uint32_t srsran_dci_dl_info(const srsran_dci_dl_t* d, char* str, uint32_t len)
{
  int n = snprintf(str, len, "f=%s, cce=%2d, L=%d, alloc=%d, mcs={%d,%d}, rv={%d,%d}, pid=%d", srsran_dci_format_string_short(d->format), d->location.ncce,
                   d->location.L, (int)d->alloc_type, d->tb[0].mcs_idx, d->tb[1].mcs_idx, d->tb[0].rv, d->tb[1].rv, d->pid);
  return n < 0 ? 0 : (uint32_t)n;
}
uint32_t srsran_dci_ul_info(srsran_dci_ul_t* d, char* str, uint32_t len)
{
  int n = snprintf(str, len, "f=0, cce=%2d, L=%d, riv=%d, mcs=%d, rv=%d, ndi=%d, n_dmrs=%d", d->location.ncce, d->location.L, d->type2_alloc.riv, d->tb.mcs_idx,
                   d->tb.rv, d->tb.ndi, d->n_dmrs);
  return n < 0 ? 0 : (uint32_t)n;
}

// ---------------------------------------------------------------------------------------------------- ra.c / ra_dl.c
uint32_t srsran_ra_type0_P(uint32_t nof_prb) { return nof_prb <= 10 ? 1 : nof_prb <= 26 ? 2 : nof_prb <= 63 ? 3 : 4; }
uint32_t srsran_ra_type2_to_riv(uint32_t L_crb, uint32_t RB_start, uint32_t nof_prb)
{
  return (L_crb - 1) <= nof_prb / 2 ? nof_prb * (L_crb - 1) + RB_start : nof_prb * (nof_prb - L_crb + 1) + nof_prb - 1 - RB_start;
}
// srsran_ra_type2_from_riv (ul_sniffer_pusch.c:28)
void srsran_ra_type2_from_riv(uint32_t riv, uint32_t* L_crb, uint32_t* RB_start, uint32_t nof_prb, uint32_t nof_vrb)
{
  *L_crb    = riv / nof_prb + 1;
  *RB_start = riv % nof_prb;
  if (*L_crb > nof_vrb - *RB_start) {
    *L_crb    = nof_prb - riv / nof_prb + 1;
    *RB_start = nof_prb - 1 - riv % nof_prb;
  }
}
// srsran_ra_tbs_from_idx (dl_sniffer_pdsch.c:49, ul_sniffer_pusch.c:186-193): 36.213 Table 7.1.7.2.1-1, rows 0..33
int srsran_ra_tbs_from_idx(uint32_t tbs_idx, uint32_t n_prb)
{
  if (tbs_idx < LTE_TBS_NOF_ITBS && n_prb > 0 && n_prb <= SRSRAN_MAX_PRB) return lte_tbs_table[tbs_idx][n_prb - 1];
  return SRSRAN_ERROR;
}
int srsran_ra_tbs_idx_from_mcs(uint32_t mcs, bool use_tbs_index_alt, bool is_ul)
{
  if (mcs >= 29) return SRSRAN_ERROR;
  if (is_ul) return mcs <= 10 ? (int)mcs : mcs <= 20 ? (int)mcs - 1 : (int)mcs - 2;
  return use_tbs_index_alt ? (mcs < 28 ? lte_dl_mcs_itbs_alt[mcs] : SRSRAN_ERROR) : lte_dl_mcs_itbs[mcs];
}
static srsran_mod_t mod_of_qm(int qm) { return qm == 2 ? SRSRAN_MOD_QPSK : qm == 4 ? SRSRAN_MOD_16QAM : qm == 6 ? SRSRAN_MOD_64QAM : SRSRAN_MOD_256QAM; }
srsran_mod_t srsran_ra_dl_mod_from_mcs(uint32_t mcs, bool alt) { return mod_of_qm(mcs < 32 ? (alt ? lte_dl_mcs_qm_alt[mcs] : lte_dl_mcs_qm[mcs]) : 2); }
srsran_mod_t srsran_ra_ul_mod_from_mcs(uint32_t mcs) { return mcs <= 10 ? SRSRAN_MOD_QPSK : mcs <= 20 ? SRSRAN_MOD_16QAM : SRSRAN_MOD_64QAM; }
// srsran_dl_fill_ra_mcs (dl_sniffer_pdsch.c:78): modulation and TBS of one transport block; returns the TBS, 0 for a retransmission
// MCS (29..31: "TBS from the latest PDCCH", which LTESniffer's zeroed grant does not have), negative on error
int srsran_dl_fill_ra_mcs(srsran_ra_tb_t* tb, int last_tbs, uint32_t nprb, bool pdsch_use_tbs_index_alt)
{
  if (!tb || tb->mcs_idx > 31) return SRSRAN_ERROR;
  const int itbs = pdsch_use_tbs_index_alt ? lte_dl_mcs_itbs_alt[tb->mcs_idx] : lte_dl_mcs_itbs[tb->mcs_idx];
  tb->mod        = srsran_ra_dl_mod_from_mcs(tb->mcs_idx, pdsch_use_tbs_index_alt);
  int tbs        = 0;
  if (itbs >= 0) {
    tbs     = srsran_ra_tbs_from_idx((uint32_t)itbs, nprb);
    tb->tbs = tbs;
  } else
    tb->tbs = last_tbs;
  return tbs;
}
static void to_fields(const srsran_dci_dl_t* d, ltehost::DlDciFields& f)
{
  f        = ltehost::DlDciFields{};
  f.format = (uint8_t)d->format, f.rnti = d->rnti, f.alloc = (uint32_t)d->alloc_type;
  f.rbg_mask = d->type0_alloc.rbg_bitmask;
  if (d->alloc_type == SRSRAN_RA_ALLOC_TYPE1) f.t1_mask = d->type1_alloc.vrb_bitmask, f.t1_subset = d->type1_alloc.rbg_subset, f.t1_shift = d->type1_alloc.shift ? 1 : 0;
  if (d->alloc_type == SRSRAN_RA_ALLOC_TYPE2) {
    f.riv = d->type2_alloc.riv, f.dist = d->type2_alloc.mode == srsran_ra_type2_t::SRSRAN_RA_TYPE2_DIST, f.ngap2 = d->type2_alloc.n_gap == srsran_ra_type2_t::SRSRAN_RA_TYPE2_NG2;
    f.n_prb1a = d->type2_alloc.n_prb1a == srsran_ra_type2_t::SRSRAN_RA_TYPE2_NPRB1A_2 ? 2 : 3;
  }
}
// srsran_ra_dl_grant_to_grant_prb_allocation (dl_sniffer_pdsch.c:103): 36.213 7.1.6 resource allocation types 0, 1, 2 (localised and
// distributed virtual resource blocks) -> PRBs of both slots
int srsran_ra_dl_grant_to_grant_prb_allocation(const srsran_dci_dl_t* dci, srsran_pdsch_grant_t* grant, uint32_t nof_prb)
{
  if (!dci || !grant || nof_prb == 0 || nof_prb > SRSRAN_MAX_PRB) return SRSRAN_ERROR;
  ltehost::DlDciFields f;
  to_fields(dci, f);
  uint32_t mask[2][4], n = 0;
  if (ltehost_dl_prb_allocation(nof_prb, f, mask, &n) != LTEPHY_SUCCESS) return SRSRAN_ERROR;
  for (uint32_t sl = 0; sl < 2; sl++)
    for (uint32_t p = 0; p < SRSRAN_MAX_PRB; p++) grant->prb_idx[sl][p] = p < nof_prb && ((mask[sl][p >> 5] >> (p & 31u)) & 1u);
  grant->nof_prb = n;
  return SRSRAN_SUCCESS;
}
// srsran_ra_dl_compute_nof_re (dl_sniffer_pdsch.c:110): data REs of the grant (CRS, PDCCH symbols, PSS / SSS / PBCH excluded) and
// the coded bits of every enabled transport block
uint32_t srsran_ra_dl_grant_nof_re(const srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_pdsch_grant_t* grant)
{
  const ltehost::Cell c = host_cell(cell);
  uint16_t            kk[12];
  uint32_t            n = 0;
  const uint32_t      cfi = sf->cfi >= 1 && sf->cfi <= 3 ? sf->cfi : 1;
  for (uint32_t l = 0; l < 14; l++)
    for (uint32_t p = 0; p < cell->nof_prb; p++)
      if (grant->prb_idx[l / 7][p]) n += ltehost::pdsch_re_in_prb(c, sf->tti % 10, cfi, l, p, kk);
  return n;
}
void srsran_ra_dl_compute_nof_re(const srsran_cell_t* cell, srsran_dl_sf_cfg_t* sf, srsran_pdsch_grant_t* grant)
{
  grant->nof_re = srsran_ra_dl_grant_nof_re(cell, sf, grant);
  const uint32_t cfi = sf->cfi >= 1 && sf->cfi <= 3 ? sf->cfi : 1;
  grant->nof_symb_slot[0] = 7 - cfi, grant->nof_symb_slot[1] = 7;
  for (int i = 0; i < SRSRAN_MAX_TB; i++)
    if (grant->tb[i].enabled) grant->tb[i].nof_bits = grant->nof_re * srsran_mod_bits_x_symbol(grant->tb[i].mod);
}

// ---------------------------------------------------------------------------------------------------- ra_ul.c
// srsran_ra_ul_compute_nof_re (ul_sniffer_pusch.c:155): 12 data symbols per subframe with normal CP (one more lost to SRS)
void srsran_ra_ul_compute_nof_re(srsran_pusch_grant_t* grant, srsran_cp_t cp, uint32_t N_srs)
{
  grant->nof_symb = 2 * (SRSRAN_CP_NSYMB(cp) - 1) - N_srs;
  grant->nof_re   = grant->nof_symb * grant->L_prb * SRSRAN_NRE;
  if (grant->tb.mod < SRSRAN_MOD_NITEMS) grant->tb.nof_bits = grant->nof_re * srsran_mod_bits_x_symbol(grant->tb.mod);
}
int srsran_ra_ul_nof_re(srsran_pusch_grant_t* grant, srsran_cp_t cp, uint32_t N_srs) { return (int)((2 * (SRSRAN_CP_NSYMB(cp) - 1) - N_srs) * grant->L_prb * SRSRAN_NRE); }
// srsran_ra_ul_dci_to_grant (falcon_dci.c:222): DCI format 0 -> PUSCH grant, 36.213 8.1 (allocation), 8.4 (type-1 hopping offsets),
// 8.6.1 Table 8.6.1-1 (MCS) -- the same steps as the reference's own copy ul_sniffer_ra_ul_dci_to_grant (ul_sniffer_pusch.c:209-239)
int srsran_ra_ul_dci_to_grant(srsran_cell_t* cell, srsran_ul_sf_cfg_t* sf, srsran_pusch_hopping_cfg_t* hopping_cfg, srsran_dci_ul_t* dci, srsran_pusch_grant_t* grant)
{
  (void)sf;
  if (!cell || !hopping_cfg || !dci || !grant) return SRSRAN_ERROR_INVALID_INPUTS;
  const uint32_t N = cell->nof_prb;
  uint32_t       n_prb_1 = 0, n_rb_ho = hopping_cfg->n_rb_ho;
  srsran_ra_type2_from_riv(dci->type2_alloc.riv, &grant->L_prb, &n_prb_1, N, N);
  if (n_rb_ho % 2) n_rb_ho++;
  if (dci->freq_hop_fl == SRSRAN_RA_PUSCH_HOP_DISABLED || dci->freq_hop_fl == SRSRAN_RA_PUSCH_HOP_TYPE2) {
    grant->n_prb[0] = grant->n_prb[1] = n_prb_1;
    grant->freq_hopping = dci->freq_hop_fl == SRSRAN_RA_PUSCH_HOP_DISABLED ? 0 : 2;
  } else { // type 1: fixed offset between the slots
    const uint32_t n_rb_pusch = N - n_rb_ho - (N % 2);
    grant->n_prb[0]           = n_prb_1;
    if (n_prb_1 < n_rb_ho / 2) return SRSRAN_ERROR;
    switch (dci->freq_hop_fl) {
      case SRSRAN_RA_PUSCH_HOP_QUART: grant->n_prb[1] = (n_rb_pusch / 4 + n_prb_1) % n_rb_pusch; break;
      case SRSRAN_RA_PUSCH_HOP_QUART_NEG: grant->n_prb[1] = n_prb_1 < n_rb_pusch / 4 ? n_rb_pusch + n_prb_1 - n_rb_pusch / 4 : n_prb_1 - n_rb_pusch / 4; break;
      case SRSRAN_RA_PUSCH_HOP_HALF: grant->n_prb[1] = (n_rb_pusch / 2 + n_prb_1) % n_rb_pusch; break;
      default: break;
    }
    grant->freq_hopping = 1;
  }
  if (!(grant->n_prb[0] + grant->L_prb <= N && grant->n_prb[1] + grant->L_prb <= N)) return SRSRAN_ERROR;
  grant->tb.mcs_idx = dci->tb.mcs_idx, grant->tb.rv = dci->tb.rv;
  grant->n_dmrs     = dci->n_dmrs; // cyclic shift FIELD of the DCI; 36.211 Table 5.5.2.1.1-1 is applied by the DMRS generator
  srsran_ra_tb_t* tb = &grant->tb;
  if (tb->mcs_idx <= 28) {
    tb->mod = srsran_ra_ul_mod_from_mcs(tb->mcs_idx);
    tb->tbs = srsran_ra_tbs_from_idx((uint32_t)srsran_ra_tbs_idx_from_mcs(tb->mcs_idx, false, true), grant->L_prb);
  } else if (tb->mcs_idx == 29 && dci->cqi_request && grant->L_prb <= 4) {
    tb->mod = SRSRAN_MOD_QPSK, tb->tbs = 0, tb->rv = 1;
  } else {
    tb->tbs = grant->last_tb.tbs, tb->mod = grant->last_tb.mod, tb->rv = (int)tb->mcs_idx - 28;
  }
  srsran_ra_ul_compute_nof_re(grant, cell->cp, 0);
  for (uint32_t i = 0; i < 2; i++) grant->n_prb_tilde[i] = grant->n_prb[i];
  if (grant->nof_symb == 0 || grant->nof_re == 0) return SRSRAN_ERROR;
  return SRSRAN_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------- pdcch.c (search spaces)
// srsran_pdcch_ue_locations_ncce / srsran_pdcch_common_locations_ncce (falcon_pdcch.c:183-196): 36.213 9.1.1 candidates
uint32_t srsran_pdcch_ue_locations_ncce_L(uint32_t nof_cce, srsran_dci_location_t* c, uint32_t max_candidates, uint32_t sf_idx, uint16_t rnti, int Ls)
{
  static const int cand[4] = {6, 6, 2, 2};
  uint32_t         k = 0, Yk = rnti;
  if (!c) return 0;
  for (uint32_t m = 0; m < sf_idx + 1; m++) Yk = (39827u * Yk) % 65537u;
  for (int l = 3; l >= 0; l--) { // All aggregation levels from 8 to 1
    if (Ls >= 0 && Ls != l) continue;
    const uint32_t L = 1u << l;
    if (nof_cce < L) continue;
    for (int i = 0; i < cand[l]; i++) { // for each candidate as given in table 9.1.1-1
      const uint32_t ncce = L * ((Yk + (uint32_t)i) % (nof_cce / L));
      if (k < max_candidates && ncce + L <= nof_cce) {
        c[k].L = (uint32_t)l, c[k].ncce = ncce;
        k++;
      }
    }
  }
  return k;
}
uint32_t srsran_pdcch_ue_locations_ncce(uint32_t nof_cce, srsran_dci_location_t* c, uint32_t max_candidates, uint32_t sf_idx, uint16_t rnti)
{
  return srsran_pdcch_ue_locations_ncce_L(nof_cce, c, max_candidates, sf_idx, rnti, -1);
}
uint32_t srsran_pdcch_common_locations_ncce(uint32_t nof_cce, srsran_dci_location_t* c, uint32_t max_candidates)
{
  uint32_t k = 0;
  for (int l = 3; l > 1; l--) {
    const uint32_t L = 1u << l;
    if (nof_cce < L) continue;
    for (uint32_t i = 0; i < SRSRAN_MIN(nof_cce, 16) / L; i++) {
      const uint32_t ncce = L * (i % (nof_cce / L));
      if (k < max_candidates && ncce + L <= nof_cce) {
        c[k].L = (uint32_t)l, c[k].ncce = ncce;
        k++;
      }
    }
  }
  return k;
}
float srsran_pdcch_coderate(uint32_t nof_bits, uint32_t l) { return (float)(nof_bits + 16) / (4 * (72u << l) / 8.0f); }

// ---------------------------------------------------------------------------------------------------- CPU coders that are NOT part of
// the accelerated path: falcon_pdcch.c:376-445 (dci_decode_and_check_list, re-encoding check) is dead code in LTESniffer (only
// srsran_pdcch_decode_msg_limit_avg_llr_power is called, src/src/DCISearch.cc:133); the symbols exist so that the file links.
int srsran_rm_conv_tx(uint8_t*, uint32_t, uint8_t*, uint32_t) { return SRSRAN_ERROR; }
int srsran_rm_conv_rx(float*, uint32_t, float*, uint32_t) { return SRSRAN_ERROR; }
int srsran_rm_conv_rx_s(int16_t*, uint32_t, int16_t*, uint32_t) { return SRSRAN_ERROR; }
int srsran_viterbi_decode_f(srsran_viterbi_t*, float*, uint8_t*, uint32_t) { return SRSRAN_ERROR; }
int srsran_convcoder_encode(srsran_convcoder_t*, uint8_t*, uint8_t*, uint32_t) { return SRSRAN_ERROR; }
void srsran_pdcch_dci_encode_conv(srsran_pdcch_t*, uint8_t*, uint32_t, uint8_t*, uint16_t) {}
uint32_t srsran_crc_checksum(srsran_crc_t* h, uint8_t* data, int len)
{ // bitwise CRC over unpacked bits (used only by the dead re-encoding check above)
  const uint32_t poly = (uint32_t)h->polynom;
  const int      order = h->order;
  uint32_t       reg = 0;
  for (int i = 0; i < len + order; i++) {
    reg = (reg << 1) | (i < len ? (data[i] & 1u) : 0u);
    if (reg & (1u << order)) reg ^= poly;
  }
  return reg & ((1u << order) - 1u);
}
uint32_t srsran_crc_attach(srsran_crc_t* h, uint8_t* data, int len)
{
  const uint32_t c = srsran_crc_checksum(h, data, len);
  uint8_t*       p = &data[len];
  srsran_bit_unpack(c, &p, h->order);
  return c;
}
int srsran_crc_init(srsran_crc_t* h, uint32_t poly, int order)
{
  memset(h, 0, sizeof(*h));
  h->polynom = (int)poly, h->order = order;
  return SRSRAN_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------- softbuffer.c
// The soft bits live on the GPU (HARQ store of libltephy_b200); the host object only carries their identity.
static uint64_t g_softbuffer_ids = 0;
int srsran_softbuffer_rx_init(srsran_softbuffer_rx_t* q, uint32_t nof_prb)
{
  if (!q) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(q, 0, sizeof(*q));
  const int tbs = srsran_ra_tbs_from_idx(33, nof_prb > 0 && nof_prb <= SRSRAN_MAX_PRB ? nof_prb : SRSRAN_MAX_PRB);
  q->max_cb     = (uint32_t)(tbs > 0 ? tbs : 97896) / 6120 + 1; // rounded up to the number of code blocks of the largest TBS
  q->max_cb_size = SOFTBUFFER_SIZE;
  q->b200_id     = __atomic_add_fetch(&g_softbuffer_ids, 1, __ATOMIC_RELAXED);
  return SRSRAN_SUCCESS;
}
int  srsran_softbuffer_rx_init_guru(srsran_softbuffer_rx_t* q, uint32_t max_cb, uint32_t max_cb_size)
{
  if (!q) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(q, 0, sizeof(*q));
  q->max_cb = max_cb, q->max_cb_size = max_cb_size, q->b200_id = __atomic_add_fetch(&g_softbuffer_ids, 1, __ATOMIC_RELAXED);
  return SRSRAN_SUCCESS;
}
void srsran_softbuffer_rx_free(srsran_softbuffer_rx_t* q)
{
  if (q) memset(q, 0, sizeof(*q));
}
void srsran_softbuffer_rx_reset(srsran_softbuffer_rx_t* q)
{
  if (q) q->b200_tbs = 0, q->tb_crc = false, q->b200_id = __atomic_add_fetch(&g_softbuffer_ids, 1, __ATOMIC_RELAXED); // new identity = empty buffer
}
void srsran_softbuffer_rx_reset_tbs(srsran_softbuffer_rx_t* q, uint32_t tbs)
{
  if (!q) return;
  srsran_softbuffer_rx_reset(q);
  q->b200_tbs = tbs;
}
void srsran_softbuffer_rx_reset_cb(srsran_softbuffer_rx_t* q, uint32_t) { srsran_softbuffer_rx_reset(q); }
void srsran_softbuffer_rx_reset_cb_crc(srsran_softbuffer_rx_t* q, uint32_t) { if (q) q->tb_crc = false; }
int  srsran_softbuffer_tx_init(srsran_softbuffer_tx_t* q, uint32_t) { if (q) memset(q, 0, sizeof(*q)); return SRSRAN_SUCCESS; }
void srsran_softbuffer_tx_reset(srsran_softbuffer_tx_t*) {}
void srsran_softbuffer_tx_free(srsran_softbuffer_tx_t*) {}

int  srsran_chest_dl_res_init(srsran_chest_dl_res_t* q, uint32_t) { if (q) memset(q, 0, sizeof(*q)); return SRSRAN_SUCCESS; }
void srsran_chest_dl_res_free(srsran_chest_dl_res_t*) {}
srsran_chest_dl_estimator_alg_t srsran_chest_dl_str2estimator_alg(const char* str)
{
  if (str && !strcmp(str, "average")) return SRSRAN_ESTIMATOR_ALG_AVERAGE;
  if (str && !strcmp(str, "wiener")) return SRSRAN_ESTIMATOR_ALG_WIENER;
  return SRSRAN_ESTIMATOR_ALG_INTERPOLATE;
}
int srsran_cqi_hl_get_subband_size(int nof_prb) { return nof_prb < 7 ? 0 : nof_prb <= 26 ? 4 : nof_prb <= 63 ? 6 : nof_prb <= 110 ? 8 : -1; }
int srsran_cqi_hl_get_no_subbands(int nof_prb)
{
  const int hl = srsran_cqi_hl_get_subband_size(nof_prb);
  return hl > 0 ? (int)ceil((float)nof_prb / hl) : 0;
}

} // extern "C"
