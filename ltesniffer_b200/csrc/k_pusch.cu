// k_pusch.cu -- K9: PUSCH receive chain for one UL grant per CTA column: DMRS least-squares estimate on the two
// reference symbols, 3-tap smoothing, noise / RSRP, time interpolation, zero-forcing equalisation, M_sc-point
// IDFT (transform de-precoding), int16 soft demodulation, descrambling, channel de-interleaving.
// Restates srsran_chest_ul_estimate_pusch + the front half of srsran_pusch_decode as called from
// PUSCH_Decoder::decode_run (reference src/src/UL_Sniffer_PUSCH.cc:250-263); the UL OFDM demodulation
// (srsran_enb_ul_fft, :392) is ofdm_rx_kernel in UL mode.  Grants without UCI, without hopping, L_prb >= 3.
// The IDFT is evaluated as the plain sum z[k] = sum_i x[i] W[(i k) mod M] in index order so that it is
// bit-identical to the oracle for every 2^a 3^b 5^c size; a mixed-radix version is a later optimisation.
#include "dev_common.cuh"
#include "dev_ul.cuh"

__device__ __forceinline__ short ul_f2s(float v)
{
  v = fminf(fmaxf(v, -32767.0f), 32767.0f);
  return (short)(int)v;
}
__device__ __forceinline__ int ul_iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ void ul_demod_s(float2 x, uint32_t qm, short* z)
{
  if (qm == 2) {
    z[0] = (short)-ul_f2s(x.x * 141.421356f);
    z[1] = (short)-ul_f2s(x.y * 141.421356f);
  } else if (qm == 4) {
    const int yr = ul_f2s(x.x * 400.0f), yi = ul_f2s(x.y * 400.0f);
    z[0] = (short)-yr, z[1] = (short)-yi;
    z[2] = (short)(ul_iabs(yr) - 252), z[3] = (short)(ul_iabs(yi) - 252);
  } else if (qm == 6) {
    const int yr = ul_f2s(x.x * 700.0f), yi = ul_f2s(x.y * 700.0f);
    z[0] = (short)-yr, z[1] = (short)-yi;
    z[2] = (short)(ul_iabs(yr) - 432), z[3] = (short)(ul_iabs(yi) - 432);
    z[4] = (short)(ul_iabs(z[2]) - 216), z[5] = (short)(ul_iabs(z[3]) - 216);
  } else { // 256QAM (third attempt of PUSCH_Decoder::decode, src/src/UL_Sniffer_PUSCH.cc:508-520)
    const int yr = ul_f2s(x.x * 1000.0f), yi = ul_f2s(x.y * 1000.0f);
    z[0] = (short)-yr, z[1] = (short)-yi;
    z[2] = (short)(ul_iabs(yr) - 613), z[3] = (short)(ul_iabs(yi) - 613);
    z[4] = (short)(ul_iabs(z[2]) - 306), z[5] = (short)(ul_iabs(z[3]) - 306);
    z[6] = (short)(ul_iabs(z[4]) - 153), z[7] = (short)(ul_iabs(z[5]) - 153);
  }
}

__global__ void __launch_bounds__(256) scr_seq_ul_kernel(const DevUlGrant* __restrict__ grants, const uint32_t* __restrict__ x1,
                                                         const uint32_t* __restrict__ basis, uint32_t basis_words, uint32_t cell_id,
                                                         uint32_t* __restrict__ seq_pool)
{
  const DevUlGrant& g = grants[blockIdx.y];
  const uint32_t nwords = (12 * g.M * g.qm + 31) / 32, wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= nwords || wi >= basis_words) return;
  const uint32_t c_init = (g.rnti << 14) + (g.sf_idx << 9) + cell_id;
  uint32_t       v      = x1[wi];
#pragma unroll
  for (uint32_t b = 0; b < 31; b++)
    if ((c_init >> b) & 1u) v ^= basis[(size_t)b * basis_words + wi];
  seq_pool[g.scr_off + wi] = v;
}

// grid (12 data symbols, grants)
__global__ void __launch_bounds__(256) pusch_kernel(const __grid_constant__ DevCell c, const DevUlGrant* __restrict__ grants,
                                                    const float2* __restrict__ ulsym, const float2* __restrict__ dmrs_pool,
                                                    const float2* __restrict__ idft_pool, const uint32_t* __restrict__ seq_pool,
                                                    short* __restrict__ llr_pool, ltephy_ul_chest_t* __restrict__ chest)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DevUlGrant g = grants[blockIdx.y];
  const uint32_t   M = g.M, tid = threadIdx.x, nt = blockDim.x, nsc = c.nsc, cidx = blockIdx.x;
  float2*          ls = reinterpret_cast<float2*>(smem_raw); // [2][M]
  float2*          sm = ls + 2 * M;                          // [2][M]
  float2*          x  = sm + 2 * M;                          // [M]
  float2*          W  = x + M;                               // [M]
  __shared__ float red[2][2];
  const float2*    y  = ulsym + (size_t)g.sf * 14 * nsc;
  const uint32_t   DATA_SYM[12] = {0, 1, 2, 4, 5, 6, 7, 8, 9, 11, 12, 13};

  for (uint32_t i = tid; i < 2 * M; i += nt) {
    const uint32_t sl = i / M, n = i % M;
    const float2   v = y[(7 * sl + 3) * nsc + g.k0 + n], r = dmrs_pool[g.dmrs_off[sl] + n];
    ls[i]            = make_float2(v.x * r.x + v.y * r.y, v.y * r.x - v.x * r.y);
  }
  for (uint32_t i = tid; i < M; i += nt) W[i] = idft_pool[g.idft_off + i];
  __syncthreads();
  const float f0 = 0.3333f, f1 = 1.0f - 2.0f * 0.3333f, f2 = 0.3333f;
  for (uint32_t i = tid; i < 2 * M; i += nt) {
    const uint32_t sl = i / M, n = i % M;
    float          ar = 0.0f, ai = 0.0f, ws = 0.0f;
    if (n > 0) ar = ar + f0 * ls[sl * M + n - 1].x, ai = ai + f0 * ls[sl * M + n - 1].y, ws = ws + f0;
    ar = ar + f1 * ls[i].x, ai = ai + f1 * ls[i].y, ws = ws + f1;
    if (n + 1 < M) ar = ar + f2 * ls[sl * M + n + 1].x, ai = ai + f2 * ls[sl * M + n + 1].y, ws = ws + f2;
    sm[i] = make_float2(ar / ws, ai / ws);
  }
  __syncthreads();
  if (cidx == 0) {
    const uint32_t warp = tid >> 5, lane = tid & 31;
    if (warp < 2) {
      float pn = 0.0f, pp = 0.0f;
      for (uint32_t n = lane; n < M; n += 32) {
        const float2 s = sm[warp * M + n], r = ls[warp * M + n];
        const float  dr = r.x - s.x, di = r.y - s.y;
        pn = pn + (dr * dr + di * di);
        pp = pp + (s.x * s.x + s.y * s.y);
      }
      pn = warp_tree_sum(pn);
      pp = warp_tree_sum(pp);
      if (lane == 0) red[warp][0] = pn, red[warp][1] = pp;
    }
    __syncthreads();
    if (tid == 0) {
      float nsum = 0.0f, psum = 0.0f;
      nsum = nsum + red[0][0], nsum = nsum + red[1][0];
      psum = psum + red[0][1], psum = psum + red[1][1];
      const float ncorr = (1.0f - 2.0f * f1) + (f0 * f0 + f1 * f1 + f2 * f2);
      chest[blockIdx.y].noise = (nsum / (float)(2 * M)) / ncorr;
      chest[blockIdx.y].rsrp  = psum / (float)(2 * M);
    }
  }
  // ---- this CTA's data symbol: equalise, IDFT, demap, descramble, de-interleave --------------------
  const uint32_t l = DATA_SYM[cidx];
  const float    t = (float)((int)l - 3) / 7.0f;
  for (uint32_t n = tid; n < M; n += nt) {
    const float2 A = sm[n], B = sm[M + n];
    const float2 h = make_float2(A.x + (B.x - A.x) * t, A.y + (B.y - A.y) * t);
    const float2 v = y[l * nsc + g.k0 + n];
    const float  den = h.x * h.x + h.y * h.y;
    x[n]             = make_float2((v.x * h.x + v.y * h.y) / den, (v.y * h.x - v.x * h.y) / den);
  }
  __syncthreads();
  const float     scl = 1.0f / sqrtf((float)M);
  const uint32_t* seq = seq_pool + g.scr_off;
  short*          out = llr_pool + g.llr_off;
  const uint32_t  qm  = g.qm;
  for (uint32_t k = tid; k < M; k += nt) {
    float    ar = 0.0f, ai = 0.0f;
    uint32_t idx = 0; // (i * k) mod M, advanced incrementally
    for (uint32_t i = 0; i < M; i++) {
      const float2 w = W[idx], xi = x[i];
      ar  = ar + (xi.x * w.x - xi.y * w.y);
      ai  = ai + (xi.x * w.y + xi.y * w.x);
      idx += k;
      if (idx >= M) idx -= M;
    }
    short z[8];
    ul_demod_s(make_float2(ar * scl, ai * scl), qm, z);
    const uint32_t hb0 = (cidx * M + k) * qm;
    for (uint32_t b = 0; b < qm; b++) {
      const uint32_t hb = hb0 + b, sbit = (seq[hb >> 5] >> (hb & 31)) & 1u;
      out[(k * 12 + cidx) * qm + b] = sbit ? (short)-z[b] : z[b];
    }
  }
}

extern "C" void launch_pusch(const DevCell& c, const DevUlGrant* grants, uint32_t ngrants, uint32_t max_M, uint32_t max_words, const float2* ulsym,
                             const float2* dmrs_pool, const float2* idft_pool, const uint32_t* x1, const uint32_t* basis, uint32_t basis_words,
                             uint32_t* seq_pool, short* llr_pool, ltephy_ul_chest_t* chest, cudaStream_t st, uint64_t* launches)
{
  if (!ngrants) return;
  scr_seq_ul_kernel<<<dim3((max_words + 255) / 256, ngrants), 256, 0, st>>>(grants, x1, basis, basis_words, c.cell_id, seq_pool);
  const size_t smem = (size_t)6 * max_M * sizeof(float2);
  cudaFuncSetAttribute(pusch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 1200 * (int)sizeof(float2)); // per device / context
  pusch_kernel<<<dim3(12, ngrants), 256, smem, st>>>(c, grants, ulsym, dmrs_pool, idft_pool, seq_pool, llr_pool, chest);
  *launches += 2;
}
