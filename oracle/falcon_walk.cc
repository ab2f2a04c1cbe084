// falcon_walk.cc -- ORACLE (test infrastructure): restatement of FALCON's recursive blind DCI search
// exactly as LTESniffer runs it -- DCISearch::search / recursive_blind_dci_search /
// inspect_dci_location_recursively (reference src/src/DCISearch.cc:102-578), DCIMetaFormats::update_formats
// (src/src/MetaFormats.cc:41-89) -- driven by the REFERENCE'S OWN RNTIManager, which build_ref.sh compiles
// from /root/reference/lib/src/util/{RNTIManager,Histogram,Interval}.cc into the same shared object.
// Like the reference it decodes candidates lazily, one srsran_pdcch_dci_decode equivalent
// (lteo_dci_decode) per (location, format) visited.
#include "falcon/util/RNTIManager.h"
#include "lte_oracle.h"
#include <cmath>
#include <cstring>
#include <vector>

extern "C" {
typedef struct {
  uint16_t rnti;
  uint8_t  format, L;
  uint16_t ncce, nof_bits;
  uint8_t  bits[LTE_DCI_MAX_BITS];
  uint32_t histval;
} lteo_dci_out_t;
typedef struct {
  uint32_t nof_decoded_locations, nof_cce, nof_missed_cce, nof_subframes, nof_locations;
} lteo_walk_stats_t;
}

namespace {
const int NOF_FORMATS = 9;
struct Location {
  uint32_t L, ncce;
  bool     used, occupied, checked, sufficient_power;
};
struct CceMap {
  Location* location[4];
  float     power;
};
struct Cand {
  uint16_t rnti;
  uint8_t  bits[LTE_DCI_MAX_BITS];
  uint32_t nof_bits;
  int      format; // decoded format
  uint32_t ssm;
};
struct Meta {
  int      format;
  uint32_t global_index, hits;
};
struct TempDci0 {
  uint16_t rnti;
  uint32_t L, ncce;
  int      format;
  Cand     cand;
};

struct Walk {
  lte_cell_t            cell;
  RNTIManager           rm;
  Meta                  all[NOF_FORMATS];
  std::vector<Meta*>    primary, secondary;
  double                split_ratio = 0.99;
  bool                  skip_secondary = false, shortcut = true;
  uint32_t              update_interval = 500, sf_cnt = 0;
  lteo_walk_stats_t     stats{};
  // per subframe
  const float*          llr = nullptr;
  uint32_t              sf_idx = 0, cfi = 0, nof_cce = 0;
  std::vector<TempDci0> temp_dci0;
  std::vector<lteo_dci_out_t>* out = nullptr;

  Walk(const lte_cell_t& c, uint32_t thr) : cell(c), rm(NOF_FORMATS, RNTI_PER_SUBFRAME, thr)
  {
    for (int i = 0; i < NOF_FORMATS; i++) all[i] = {i, (uint32_t)i, 0};
    update_formats();
  }
  void update_formats() // MetaFormats.cc:41-89
  {
    Meta*  sorted[NOF_FORMATS];
    double total = 0;
    for (int i = 0; i < NOF_FORMATS; i++) {
      sorted[i] = &all[i];
      total += all[i].hits;
    }
    for (int i = 0; i < NOF_FORMATS - 1; i++) {
      int mx = i;
      for (int j = mx; j < NOF_FORMATS; j++)
        if (sorted[j]->hits > sorted[mx]->hits) mx = j;
      std::swap(sorted[i], sorted[mx]);
    }
    double thr = total * split_ratio, cum = 0;
    primary.clear(), secondary.clear();
    for (int i = 0; i < NOF_FORMATS; i++) {
      (cum <= thr ? primary : secondary).push_back(sorted[i]);
      cum += sorted[i]->hits;
      sorted[i]->hits = 0;
    }
  }
  void add_dci(const Cand& c, uint32_t L, uint32_t ncce, uint32_t histval)
  {
    lteo_dci_out_t o{};
    o.rnti = c.rnti, o.format = (uint8_t)c.format, o.L = (uint8_t)L, o.ncce = (uint16_t)ncce, o.nof_bits = (uint16_t)c.nof_bits, o.histval = histval;
    memcpy(o.bits, c.bits, c.nof_bits);
    out->push_back(o);
  }
  // srsran_pdcch_decode_msg_limit_avg_llr_power with bound 0 (falcon_pdcch.c:110-170)
  void decode(const Location* loc, int format, Cand& c)
  {
    c           = Cand{};
    c.nof_bits  = lte_dci_sizeof(&cell, (lte_dci_format_t)format);
    uint32_t  E = 72u << loc->L;
    const float* e = &llr[loc->ncce * 72];
    double    mean = 0;
    for (uint32_t i = 0; i < E; i++) mean += fabsf(e[i]);
    mean /= E;
    c.format = 0; // cand objects are calloc'ed in the reference (falcon_alloc_candidates)
    if (mean > 0) {
      uint16_t crc = 0;
      if (lteo_dci_decode(e, E, c.nof_bits, c.bits, &crc) == 0) {
        c.rnti = crc;
        if (format == LTE_DCI_FORMAT0 || format == LTE_DCI_FORMAT1A)
          c.format = c.bits[0] == 0 ? LTE_DCI_FORMAT0 : LTE_DCI_FORMAT1A;
        else
          c.format = format;
      }
    }
  }
  int inspect(CceMap* cce_map, uint32_t ncce, uint32_t L, uint32_t max_depth, std::vector<Meta*>& mf, uint32_t discovery, const Cand* parent)
  {
    const uint32_t nf = (uint32_t)mf.size();
    int            hist_max_idx = -1;
    unsigned       hist_max_val = 0, n_above = 0;
    std::vector<Cand> cand(nf);
    Location*      loc = cce_map[ncce].location[L];
    if (loc && !loc->occupied && !loc->checked && loc->sufficient_power) {
      for (uint32_t f = 0; f < nf; f++) {
        decode(loc, mf[f]->format, cand[f]);
        stats.nof_decoded_locations++;
        if (rm.getActivationReason(cand[f].rnti) == RM_ACT_RAR && cand[f].format == 0) {
          bool add = true;
          for (auto& m : temp_dci0)
            if (m.format == cand[f].format && m.rnti == cand[f].rnti && m.ncce == ncce) add = false;
          if (add) temp_dci0.push_back({cand[f].rnti, L, ncce, cand[f].format, cand[f]});
        }
        if (mf[f]->format != cand[f].format) {
          cand[f].rnti = 0;
          continue;
        }
        if (mf[f]->format == LTE_DCI_FORMAT1C && cand[f].rnti > LTE_RARNTI_END && cand[f].rnti < LTE_PRNTI) {
          cand[f].rnti = 0;
          continue;
        }
        if (cand[f].rnti > LTE_RARNTI_START && cand[f].rnti < LTE_RARNTI_END) {
          if (mf[f]->format != LTE_DCI_FORMAT1A && mf[f]->format != LTE_DCI_FORMAT1C) {
            cand[f].rnti = 0;
            continue;
          }
        }
        if (shortcut && discovery && parent && parent[f].rnti == cand[f].rnti && !rm.isForbidden(cand[f].rnti, mf[f]->global_index))
          return -((int)f + 1);
        cand[f].ssm = lte_pdcch_validate_location(nof_cce, ncce, L, sf_idx, cand[f].rnti);
        if (cand[f].ssm == 0) {
          cand[f].rnti = 0;
          continue;
        }
        if (rm.validateAndRefresh(cand[f].rnti, mf[f]->global_index)) {
          n_above++;
          hist_max_idx = (int)f;
          hist_max_val = rm.getFrequency(cand[f].rnti, mf[f]->global_index);
        }
      }
      if (n_above > 1) {
        hist_max_idx = -1;
        uint32_t hmax = 0;
        for (uint32_t f = 0; f < nf; f++)
          if (cand[f].rnti != 0) {
            uint32_t h = rm.getFrequency(cand[f].rnti, mf[f]->global_index);
            if (h > hmax) {
              hmax = h, hist_max_idx = (int)f;
              hist_max_val = h;
            }
          }
        if (hist_max_idx == -1) n_above = 0;
      }
      loc->checked = true;
      int disamb = 0;
      if (n_above > 0 && cand[hist_max_idx].ssm == 1) {
        if (L > 0 && max_depth > 0) disamb = inspect(cce_map, ncce + (1u << (L - 1)), L - 1, max_depth - 1, mf, 0, nullptr);
      } else if (n_above == 0) {
        int rr = 0;
        if (L > 0 && max_depth > 0) {
          rr += inspect(cce_map, ncce, L - 1, max_depth - 1, mf, discovery, cand.data());
          if (rr < 0) {
            hist_max_idx = -rr - 1;
            hist_max_val = rm.getFrequency(cand[hist_max_idx].rnti, mf[hist_max_idx]->global_index);
            n_above      = 1;
            if (cand[hist_max_idx].ssm == 1)
              disamb = inspect(cce_map, ncce + (1u << (L - 1)), L - 1, (max_depth < 99 ? max_depth : 99) - 1, mf, 0, nullptr);
            rm.activateAndRefresh(cand[hist_max_idx].rnti, mf[hist_max_idx]->global_index, RM_ACT_SHORTCUT);
          } else
            rr += inspect(cce_map, ncce + (1u << (L - 1)), L - 1, max_depth - 1, mf, discovery, nullptr);
        }
        if (rr == 0) {
          if (discovery)
            for (uint32_t f = 0; f < nf; f++)
              if (cand[f].rnti != 0) rm.addCandidate(cand[f].rnti, mf[f]->global_index);
          return 0;
        } else if (rr > 0)
          return rr;
      }
      if (n_above > 0) {
        loc->used = true;
        for (uint32_t c = ncce; c < ncce + (1u << L); c++)
          for (int a = 0; a < 4; a++)
            if (cce_map[c].location[a]) cce_map[c].location[a]->occupied = cce_map[c].location[a]->checked = true;
        rm.addCandidate(cand[hist_max_idx].rnti, mf[hist_max_idx]->global_index);
        mf[hist_max_idx]->hits++;
        uint32_t L_dis = disamb > 0 ? L - 1 : L;
        Cand&    b     = cand[hist_max_idx];
        if (b.rnti != 0) {
          bool add = true;
          if (b.format == 0)
            for (auto& m : temp_dci0)
              if (m.format == b.format && m.rnti == b.rnti && m.ncce == ncce) add = false;
          if (add) add_dci(b, L_dis, ncce, hist_max_val);
          for (auto& m : temp_dci0) add_dci(m.cand, m.L, m.ncce, rm.getFrequency(m.rnti, (uint32_t)m.format));
          temp_dci0.clear();
        }
        return 1 + disamb;
      }
    }
    return 0;
  }
  int search(uint32_t sf_idx_, uint32_t cfi_, uint32_t nof_cce_, const float* llr_, float snr_db, bool update_fmt, std::vector<lteo_dci_out_t>& o)
  {
    sf_idx = sf_idx_, cfi = cfi_, nof_cce = nof_cce_, llr = llr_, out = &o;
    if (update_fmt) update_formats(); // SubframeWorker::work, src/src/SubframeWorker.cc:148-151
    int ret = -1;
    if (snr_db > 6.0f) {
      temp_dci0.clear();
      Location locations[160];
      CceMap   cce_map[84];
      memset(cce_map, 0, sizeof(cce_map));
      stats.nof_cce += nof_cce;
      uint32_t k = 0, lim = nof_cce < 84 ? nof_cce : 84;
      for (int l = 3; l >= 0; l--) {
        uint32_t Lc = 1u << l;
        for (uint32_t i = 0; i < lim / Lc; i++)
          if (k < 160) {
            locations[k] = {(uint32_t)l, Lc * (i % (nof_cce / Lc)), false, false, false, true};
            for (uint32_t m = locations[k].ncce; m < locations[k].ncce + Lc; m++) cce_map[m].location[l] = &locations[k];
            k++;
          }
      }
      stats.nof_locations += k;
      for (uint32_t c = 0; c < nof_cce && c < 84; c++) {
        double mean = 0;
        for (int i = 0; i < 72; i++) mean += fabsf(llr[c * 72 + i]);
        cce_map[c].power = (float)(mean / 72);
        if (cce_map[c].power < 0.7f)
          for (int a = 0; a < 4; a++)
            if (cce_map[c].location[a]) cce_map[c].location[a]->sufficient_power = false;
      }
      ret = 0;
      for (uint32_t i = 0; i < k; i++) ret += inspect(cce_map, locations[i].ncce, locations[i].L, 99, primary, 1, nullptr);
      if (!skip_secondary) {
        for (uint32_t i = 0; i < k; i++) locations[i].checked = false;
        for (uint32_t i = 0; i < k; i++) ret += inspect(cce_map, locations[i].ncce, locations[i].L, 99, secondary, 1, nullptr);
      }
      uint32_t missed = 0;
      for (uint32_t c = 0; c < nof_cce && c < 84; c++) {
        if (cce_map[c].power < 0.7f) continue;
        bool m = true;
        for (int a = 0; a < 4; a++)
          if (cce_map[c].location[a] && cce_map[c].location[a]->used) {
            m = false;
            break;
          }
        missed += m;
      }
      stats.nof_missed_cce += missed;
      rm.stepTime();
    }
    stats.nof_subframes++;
    return ret;
  }
};
} // namespace

extern "C" {
void* lteo_walk_create(const lte_cell_t* cell, uint32_t threshold)
{
  Walk* w = new Walk(*cell, threshold);
  // LTESniffer_Core.cc:398-417
  for (int f : {LTE_DCI_FORMAT1A, LTE_DCI_FORMAT1C}) {
    w->rm.addEvergreen(LTE_RARNTI_START, LTE_RARNTI_END, (uint32_t)f);
    w->rm.addEvergreen(LTE_PRNTI, LTE_SIRNTI, (uint32_t)f);
  }
  for (uint32_t f = 0; f < (uint32_t)NOF_FORMATS; f++) w->rm.addForbidden(0, 0, f);
  return w;
}
void lteo_walk_destroy(void* w) { delete static_cast<Walk*>(w); }
void lteo_walk_config(void* w, int shortcut, int skip_secondary, uint32_t update_interval)
{
  Walk* k = static_cast<Walk*>(w);
  k->shortcut = shortcut, k->skip_secondary = skip_secondary, k->update_interval = update_interval;
}
int lteo_walk_subframe(void* w, uint32_t sf_idx, uint32_t cfi, uint32_t nof_cce, const float* llr, float snr_db, lteo_dci_out_t* out, uint32_t max,
                       uint32_t* n)
{
  Walk*                       k = static_cast<Walk*>(w);
  std::vector<lteo_dci_out_t> v;
  bool upd = k->update_interval && (k->sf_cnt % k->update_interval) == 0;
  k->sf_cnt++;
  int ret = k->search(sf_idx, cfi, nof_cce, llr, snr_db, upd, v);
  *n      = (uint32_t)(v.size() < max ? v.size() : max);
  memcpy(out, v.data(), *n * sizeof(lteo_dci_out_t));
  return ret;
}
void lteo_walk_stats(void* w, lteo_walk_stats_t* s) { *s = static_cast<Walk*>(w)->stats; }
void lteo_walk_activate(void* w, uint16_t rnti, uint32_t format_idx, int reason)
{
  static_cast<Walk*>(w)->rm.activateAndRefresh(rnti, format_idx, (ActivationReason)reason);
}
}
