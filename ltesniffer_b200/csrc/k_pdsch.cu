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
// soft-demodulate one equalised symbol, descramble with Qm bits taken from a 64-bit window of the sequence, store
__device__ __forceinline__ void emit(const DevGrant& g, uint32_t cw, uint32_t gi, float2 x, const uint32_t* __restrict__ seq_pool,
                                     short* __restrict__ llr_pool)
{
  const uint32_t qm = g.qm[cw];
  short          z[8];
  demod_s(x, qm, z);
  const uint32_t*    seq = seq_pool + g.scr_off[cw];
  const uint32_t     b0  = gi * qm, w0 = b0 >> 5;
  const unsigned long long win = ((unsigned long long)seq[w0] | ((unsigned long long)seq[w0 + 1] << 32)) >> (b0 & 31u);
  uint32_t*          out = reinterpret_cast<uint32_t*>(llr_pool + g.llr_off[cw] + (size_t)gi * qm); // Qm even, offsets 8-aligned
#pragma unroll
  for (uint32_t i = 0; i < 8; i += 2) {
    if (i >= qm) break;
    const int a = ((win >> i) & 1ull) ? -(int)z[i] : (int)z[i], b = ((win >> (i + 1)) & 1ull) ? -(int)z[i + 1] : (int)z[i + 1];
    out[i >> 1] = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
  }
}

// One CTA per grant.  Work items are (OFDM symbol, allocated PRB, subcarrier) slots; the 12-bit data-RE
// mask of the PRB gives the rank of the RE inside the grant (36.211 6.3.5 mapping order: k first, then l).
__global__ void __launch_bounds__(256) pdsch_demod_kernel(const __grid_constant__ DevCell c, const DevGrant* __restrict__ grants,
                                                          const float2* __restrict__ sym, const float2* __restrict__ ce,
                                                          const uint32_t* __restrict__ seq_pool, short* __restrict__ llr_pool)
{
  __shared__ uint16_t pref[14][LTEPHY_MAX_PRB + 2];
  __shared__ DevGrant gs;
  const uint32_t tid = threadIdx.x, nt = blockDim.x, N = c.nof_prb;
  for (uint32_t i = tid; i < sizeof(DevGrant) / 4; i += nt) reinterpret_cast<uint32_t*>(&gs)[i] = reinterpret_cast<const uint32_t*>(&grants[blockIdx.x])[i];
  __syncthreads();
  const DevGrant& g    = gs;
  const uint16_t* mask = c.re_mask + ((size_t)(g.cls * 3 + g.cfi - 1) * 14) * N;
  if (tid < 14) {
    const uint32_t l = tid, sl = l / 7, np = g.np[sl];
    uint32_t       acc = 0;
    for (uint32_t j = 0; j < np; j++) {
      pref[l][j] = (uint16_t)acc;
      acc += __popc(mask[l * N + g.plist[sl][j]]);
    }
    pref[l][np] = (uint16_t)acc;
  }
  __syncthreads();
  const SfView   v  = make_view(c, sym, ce, g.sf);
  const uint32_t n0 = 7 * 12 * g.np[0], n1 = 7 * 12 * g.np[1];
  for (uint32_t it = tid; it < n0 + n1; it += nt) {
    uint32_t sl, rem;
    if (it < n0)
      sl = 0, rem = it;
    else
      sl = 1, rem = it - n0;
    const uint32_t per = 12 * g.np[sl], l = 7 * sl + rem / per, q = rem % per, j = q / 12, kk = q % 12;
    const uint32_t prb = g.plist[sl][j], m = mask[l * N + prb];
    if (!((m >> kk) & 1u)) continue;
    const uint32_t r = pref[l][j] + __popc(m & ((1u << kk) - 1u)), gi = g.re_off[l] + r, idx = l * c.nsc + 12 * prb + kk;
    if (g.tx_scheme == LTEPHY_TX_PORT0) {
      emit(g, 0, gi, eq_port0(c, v, idx), seq_pool, llr_pool);
    } else if (g.tx_scheme == LTEPHY_TX_DIVERSITY) {
      if (gi & 1u) continue; // handled together with its even partner
      const uint32_t rest = m >> (kk + 1);
      uint32_t       idx2;
      if (rest)
        idx2 = idx + 1 + (uint32_t)(__ffs(rest) - 1);
      else { // partner is the first data RE of the next allocated PRB (never happens for even per-PRB counts)
        const uint32_t prb2 = g.plist[sl][j + 1], m2 = mask[l * N + prb2];
        idx2 = l * c.nsc + 12 * prb2 + (uint32_t)(__ffs(m2) - 1);
      }
      float2 x0, x1;
      eq_sfbc(c, v, idx, idx2, x0, x1);
      emit(g, 0, gi, x0, seq_pool, llr_pool);
      emit(g, 0, gi + 1, x1, seq_pool, llr_pool);
    } else if (g.tx_scheme == LTEPHY_TX_CDD) {
      float2 x0, x1;
      eq_cdd(c, v, idx, (gi & 1u) != 0, x0, x1);
      emit(g, 0, gi, x0, seq_pool, llr_pool);
      emit(g, 1, gi, x1, seq_pool, llr_pool);
    } else if (g.tx_scheme == LTEPHY_TX_SPATIALMUX) {
      if (g.ncw == 1) {
        emit(g, 0, gi, eq_spmux1(c, v, idx, spmux_w(1, g.pmi)), seq_pool, llr_pool);
      } else {
        float2 x0, x1;
        eq_spmux2(c, v, idx, spmux_w(2, g.pmi), x0, x1);
        emit(g, 0, gi, x0, seq_pool, llr_pool);
        emit(g, 1, gi, x1, seq_pool, llr_pool);
      }
    }
  }
}

// ---- K7: one CTA per code block, iterating in RECEIVED order: soft bit k (and its repeats k+nn, ...) lands on
// pair-buffer word order[k]; reads of e are coalesced, writes follow the runs of the sub-block interleaver
// (consecutive k -> consecutive words of the window-transposed stream).  Unreceived positions are zeroed here,
// so the stream buffers need no memset.
__device__ __forceinline__ int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

__global__ void __launch_bounds__(256) rm_turbo_rx_kernel(const DevCb* __restrict__ cbs, const DevPair* __restrict__ pairs,
                                                          const short* __restrict__ llr_pool, const uint32_t* __restrict__ rm_pool,
                                                          uint32_t* __restrict__ turbo_pool, short* __restrict__ harq_pool, uint32_t gen)
{
  const DevCb     cb = cbs[blockIdx.x];
  if (cb.harq_gen != gen) return;
  const DevPair&  pr = pairs[cb.pair];
  const uint32_t  NW = pr.NW;
  const bool      alone = pr.ncb == 1;
  const short*    e   = llr_pool + cb.llr_off;
  const uint32_t* ord = rm_pool + cb.rm_tab;
  uint32_t*       bw  = turbo_pool + pr.buf_off;
  short*          bh  = reinterpret_cast<short*>(bw);
  // HARQ store (reference src/src/HARQ.cc, DL_Sniffer_PDSCH.cc:955-985): the accumulators of this code block live at the pair-buffer word index,
  // which does not depend on the redundancy version; a retransmission starts from what earlier transmissions left there
  short* hq = cb.harq_op ? harq_pool + cb.harq_off : nullptr;
  for (uint32_t k = threadIdx.x; k < cb.rm_nn; k += blockDim.x) {
    const uint32_t w = ord[k];
    int            acc = cb.harq_op == LTEPHY_HARQ_RETX ? (int)hq[w] : 0;
    for (uint32_t kk = k; kk < cb.E; kk += cb.rm_nn) acc = sat16(acc + (int)e[kk]);
    if (hq) hq[w] = (short)acc;
    int v = acc >> cb.shift;
    v     = v > 255 ? 255 : (v < -255 ? -255 : v);
    if (alone)
      bw[w] = (uint32_t)v & 0xFFFFu;
    else
      bh[2 * w + cb.half] = (short)v;
  }
  // filler bits are known zeros: strong "0" in the systematic and first parity streams
  for (uint32_t i = threadIdx.x; i < 2 * cb.F; i += blockDim.x) {
    const uint32_t st = i / cb.F, ii = i % cb.F, w = st * 32 * NW + (ii & 31u) * NW + (ii >> 5);
    if (alone)
      bw[w] = (uint32_t)(-255) & 0xFFFFu;
    else
      bh[2 * w + cb.half] = (short)-255;
  }
}

extern "C" void launch_pdsch_front(const DevCell& c, const DevGrant* grants, uint32_t ngrants, uint32_t max_words, const float2* sym,
                                   const float2* ce, const uint32_t* gold_x1, const uint32_t* gold_basis, uint32_t basis_words,
                                   uint32_t* seq_pool, short* llr_pool, cudaStream_t st, uint64_t* launches)
{
  if (!ngrants) return;
  scr_seq_kernel<<<dim3((max_words + 255) / 256, 2 * ngrants), 256, 0, st>>>(grants, gold_x1, gold_basis, basis_words, c.cell_id, seq_pool);
  pdsch_demod_kernel<<<ngrants, 256, 0, st>>>(c, grants, sym, ce, seq_pool, llr_pool);
  *launches += 2;
}
extern "C" void launch_rm_turbo_rx(const DevCb* cbs, uint32_t ncb, const DevPair* pairs, const short* llr_pool, const uint32_t* rm_pool,
                                   uint32_t* turbo_pool, short* harq_pool, uint32_t gen, cudaStream_t st, uint64_t* launches)
{
  if (!ncb) return;
  rm_turbo_rx_kernel<<<ncb, 256, 0, st>>>(cbs, pairs, llr_pool, rm_pool, turbo_pool, harq_pool, gen);
  *launches += 1;
}
