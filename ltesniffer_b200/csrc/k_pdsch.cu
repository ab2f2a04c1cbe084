// k_pdsch.cu -- K6: PDSCH RE gather + equalise + layer demap + max-log soft demodulation + Gold
// descrambling -> int16 LLRs; K7: turbo rate de-matching into the decoder's stream layout.
// Restates the front half of srsran_pdsch_decode reached through srsran_ue_dl_decode_pdsch
// (reference src/src/DL_Sniffer_PDSCH.cc:997): srsran_predecoding_type, srsran_demod_soft_demodulate_s
// (int16, scales 100*sqrt2/400/700/1000), srsran_scrambling, srsran_rm_turbo_rx_lut.
#include "dev_common.cuh"
#include "dev_eq.cuh"

// ---- scrambling sequences: c = x1 ^ XOR_{b in c_init} basis[b]  (36.211 7.2 without the 1600-step warm-up)
__global__ void __launch_bounds__(256) scr_seq_kernel(const DevGrant* __restrict__ grants, const uint32_t* __restrict__ x1,
                                                      const uint32_t* __restrict__ basis, uint32_t basis_words, uint32_t cell_id,
                                                      uint32_t* __restrict__ seq_pool)
{
  const DevGrant& g  = grants[blockIdx.y >> 1];
  const uint32_t  cw = blockIdx.y & 1u;
  if (cw >= g.ncw) return;
  const uint32_t nwords = (g.nof_re * g.qm[cw] + 31) / 32;
  const uint32_t wi     = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= nwords || wi >= basis_words) return;
  const uint32_t c_init = (g.rnti << 14) + (cw << 13) + (g.sf_idx << 9) + cell_id;
  uint32_t       v      = x1[wi];
#pragma unroll
  for (uint32_t b = 0; b < 31; b++)
    if ((c_init >> b) & 1u) v ^= basis[(size_t)b * basis_words + wi];
  seq_pool[g.scr_off[cw] + wi] = v;
}

// ---- data REs of one PRB in one symbol (same rule as srsran_ra_dl_compute_nof_re / pdsch RE mapping)
__device__ __forceinline__ uint32_t re_in_prb(const DevCell& c, uint32_t sf_idx, uint32_t cfi, uint32_t l, uint32_t prb, uint16_t* kk)
{
  if (l < (c.nof_prb <= 10 ? cfi + 1 : cfi)) return 0;
  const uint32_t lo = 6 * c.nof_prb - 36, hi = lo + 72;
  const bool     crs = (l % 7 == 0) || (l % 7 == 4);
  uint32_t       n = 0;
  for (uint32_t k = 12 * prb; k < 12 * prb + 12; k++) {
    if (crs && (c.nof_ports == 1 ? (k % 6 == c.crs_off[0][(l % 7) ? 1 : 0]) : (k % 3 == c.cell_id % 3))) continue;
    if (k >= lo && k < hi) {
      if ((sf_idx == 0 || sf_idx == 5) && (l == 5 || l == 6)) continue;
      if (sf_idx == 0 && l >= 7 && l <= 10) continue;
    }
    if (kk) kk[n] = (uint16_t)k;
    n++;
  }
  return n;
}

__device__ __forceinline__ short f2s(float v)
{
  v = fminf(fmaxf(v, -32767.0f), 32767.0f);
  return (short)(int)v; // truncation toward zero
}
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ void demod_s(float2 x, uint32_t qm, short* z)
{
  if (qm == 2) {
    z[0] = (short)-f2s(x.x * 141.421356f);
    z[1] = (short)-f2s(x.y * 141.421356f);
  } else if (qm == 4) {
    const int yr = f2s(x.x * 400.0f), yi = f2s(x.y * 400.0f);
    z[0] = (short)-yr, z[1] = (short)-yi;
    z[2] = (short)(iabs(yr) - 252), z[3] = (short)(iabs(yi) - 252);
  } else if (qm == 6) {
    const int yr = f2s(x.x * 700.0f), yi = f2s(x.y * 700.0f);
    z[0] = (short)-yr, z[1] = (short)-yi;
    z[2] = (short)(iabs(yr) - 432), z[3] = (short)(iabs(yi) - 432);
    z[4] = (short)(iabs(z[2]) - 216), z[5] = (short)(iabs(z[3]) - 216);
  } else {
    const int yr = f2s(x.x * 1000.0f), yi = f2s(x.y * 1000.0f);
    z[0] = (short)-yr, z[1] = (short)-yi;
    z[2] = (short)(iabs(yr) - 613), z[3] = (short)(iabs(yi) - 613);
    z[4] = (short)(iabs(z[2]) - 306), z[5] = (short)(iabs(z[3]) - 306);
    z[6] = (short)(iabs(z[4]) - 153), z[7] = (short)(iabs(z[5]) - 153);
  }
}
__device__ __forceinline__ void emit(const DevGrant& g, uint32_t cw, uint32_t gi, float2 x, const uint32_t* __restrict__ seq_pool,
                                     short* __restrict__ llr_pool)
{
  const uint32_t qm = g.qm[cw];
  short          z[8];
  demod_s(x, qm, z);
  const uint32_t* seq = seq_pool + g.scr_off[cw];
  short*          out = llr_pool + g.llr_off[cw] + (size_t)gi * qm;
  const uint32_t  b0  = gi * qm;
  for (uint32_t i = 0; i < qm; i++) {
    const uint32_t b = b0 + i, s = (seq[b >> 5] >> (b & 31)) & 1u;
    out[i]           = s ? (short)-z[i] : z[i];
  }
}

// one CTA per (OFDM symbol, grant)
__global__ void __launch_bounds__(256) pdsch_demod_kernel(const __grid_constant__ DevCell c, const DevGrant* __restrict__ grants,
                                                          const float2* __restrict__ sym, const float2* __restrict__ ce,
                                                          const uint32_t* __restrict__ seq_pool, short* __restrict__ llr_pool)
{
  __shared__ uint16_t klist[12 * LTEPHY_MAX_PRB];
  __shared__ uint16_t pref[LTEPHY_MAX_PRB + 1];
  const DevGrant&    g = grants[blockIdx.y];
  const uint32_t     l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const uint32_t     nre = g.re_off[l + 1] - g.re_off[l];
  if (nre == 0) return;
  const uint32_t slot = l / 7;
  if (tid == 0) {
    uint32_t acc = 0;
    for (uint32_t prb = 0; prb < c.nof_prb; prb++) {
      pref[prb] = (uint16_t)acc;
      if ((g.prb_mask[slot][prb >> 5] >> (prb & 31)) & 1u) acc += re_in_prb(c, g.sf_idx, g.cfi, l, prb, nullptr);
    }
    pref[c.nof_prb] = (uint16_t)acc;
  }
  __syncthreads();
  for (uint32_t prb = tid; prb < c.nof_prb; prb += nt)
    if (pref[prb + 1] > pref[prb]) re_in_prb(c, g.sf_idx, g.cfi, l, prb, &klist[pref[prb]]);
  __syncthreads();
  const SfView   v    = make_view(c, sym, ce, g.sf);
  const uint32_t base = g.re_off[l];
  if (g.tx_scheme == LTEPHY_TX_PORT0) {
    for (uint32_t i = tid; i < nre; i += nt) emit(g, 0, base + i, eq_port0(c, v, l * c.nsc + klist[i]), seq_pool, llr_pool);
  } else if (g.tx_scheme == LTEPHY_TX_DIVERSITY) {
    for (uint32_t i = 2 * tid; i + 1 < nre; i += 2 * nt) {
      float2 x0, x1;
      eq_sfbc(c, v, l * c.nsc + klist[i], l * c.nsc + klist[i + 1], x0, x1);
      emit(g, 0, base + i, x0, seq_pool, llr_pool);
      emit(g, 0, base + i + 1, x1, seq_pool, llr_pool);
    }
  } else if (g.tx_scheme == LTEPHY_TX_CDD) {
    for (uint32_t i = tid; i < nre; i += nt) {
      float2 x0, x1;
      eq_cdd(v, l * c.nsc + klist[i], ((base + i) & 1u) != 0, x0, x1);
      emit(g, 0, base + i, x0, seq_pool, llr_pool);
      emit(g, 1, base + i, x1, seq_pool, llr_pool);
    }
  }
}

// ---- K7: one CTA per code block.  For every stream position: saturating sum of the soft bits that
// land on it (first, first+nn, ...), conditioning shift/clamp, store into the turbo stream buffers
// (window-transposed, two code blocks interleaved as int16x2).
__device__ __forceinline__ int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

__global__ void __launch_bounds__(256) rm_turbo_rx_kernel(const DevCb* __restrict__ cbs, const DevPair* __restrict__ pairs,
                                                          const short* __restrict__ llr_pool, const uint32_t* __restrict__ rm_pool,
                                                          uint32_t* __restrict__ turbo_pool)
{
  const DevCb&    cb = cbs[blockIdx.x];
  const DevPair&  pr = pairs[cb.pair];
  const uint32_t  K = cb.K, D = K + 4, NW = pr.NW;
  const short*    e   = llr_pool + cb.llr_off;
  const uint32_t* tab = rm_pool + cb.rm_tab;
  short*          buf = reinterpret_cast<short*>(turbo_pool + pr.buf_off);
  for (uint32_t t = threadIdx.x; t < 3 * D; t += blockDim.x) {
    const uint32_t s = t / D, i = t % D, first = tab[t];
    int            acc = 0;
    if (first != 0xFFFFFFFFu)
      for (uint32_t k = first; k < cb.E; k += cb.rm_nn) acc = sat16(acc + (int)e[k]);
    int v = acc >> cb.shift;
    v     = v > 255 ? 255 : (v < -255 ? -255 : v);
    if (s < 2 && i < cb.F) v = -255;
    uint32_t word; // word index inside the pair buffer
    if (i < K)
      word = s * 32 * NW + (i & 31u) * NW + (i >> 5);
    else
      word = 5 * 32 * NW + s * 4 + (i - K);
    buf[2 * word + cb.half] = (short)v;
  }
}

extern "C" void launch_pdsch_front(const DevCell& c, const DevGrant* grants, uint32_t ngrants, uint32_t max_words, const float2* sym,
                                   const float2* ce, const uint32_t* gold_x1, const uint32_t* gold_basis, uint32_t basis_words,
                                   uint32_t* seq_pool, short* llr_pool, cudaStream_t st, uint64_t* launches)
{
  if (!ngrants) return;
  scr_seq_kernel<<<dim3((max_words + 255) / 256, 2 * ngrants), 256, 0, st>>>(grants, gold_x1, gold_basis, basis_words, c.cell_id, seq_pool);
  pdsch_demod_kernel<<<dim3(14, ngrants), 256, 0, st>>>(c, grants, sym, ce, seq_pool, llr_pool);
  *launches += 2;
}
extern "C" void launch_rm_turbo_rx(const DevCb* cbs, uint32_t ncb, const DevPair* pairs, const short* llr_pool, const uint32_t* rm_pool,
                                   uint32_t* turbo_pool, cudaStream_t st, uint64_t* launches)
{
  if (!ncb) return;
  rm_turbo_rx_kernel<<<ncb, 256, 0, st>>>(cbs, pairs, llr_pool, rm_pool, turbo_pool);
  *launches += 1;
}
