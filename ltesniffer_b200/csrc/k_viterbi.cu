// k_viterbi.cu -- K5: exhaustive blind-DCI decode.  One warp decodes TWO PDCCH candidates of the same
// payload size at once (their path metrics share registers as packed int16x2, VIADD.16x2 / VIMNMX.S16x2);
// the 64 trellis states are spread two per lane.  Restates srsran_pdcch_dci_decode as called from
// srsran_pdcch_decode_msg_limit_avg_llr_power (reference lib/src/phy/falcon_phch/falcon_pdcch.c:110-170):
// conv rate-dematch with accumulation, uint8 quantisation (gain 32 / max|x|), K=7 r=1/3 tail-biting
// Viterbi run over three concatenated copies of which the middle one is kept, CRC16, RNTI = parity ^ crc.
// The table T[location][size] this kernel fills is what DCISearch::inspect_dci_location_recursively
// (src/src/DCISearch.cc:102-447) consults one entry at a time.
#include "dev_common.cuh"

#define VIT_KMAX 80          // nof_bits <= 64 -> K = nof_bits + 16 <= 80
#define VIT_WARPS 4
#define VIT_RENORM 16

struct __align__(16) VitWarpSmem {
  float    rm[2][3 * VIT_KMAX]; // dematched soft bits, stream-major
  uint32_t R[3][VIT_KMAX];      // quantised symbols 2q-255, packed (cand0 lo, cand1 hi)
  uint32_t S[VIT_KMAX][8];      // branch metrics for the 8 output sign patterns (o0 o1 o2), packed; S[k][p ^ 7] = -S[k][p]
  uint4    dec[2 * VIT_KMAX];   // survivor decisions of steps K..3K-1: {c0 even, c0 odd, c1 even, c1 odd}
};

__device__ __forceinline__ uint32_t pk(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int      lo16(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int      hi16(uint32_t v) { return (int)(short)(v >> 16); }

__global__ void __launch_bounds__(VIT_WARPS * 32) dci_viterbi_kernel(const __grid_constant__ DevCell c, const float* __restrict__ llr_all,
                                                                      const DevSfInfo* __restrict__ info, ltephy_cand_t* __restrict__ cands)
{
  __shared__ VitWarpSmem sm_all[VIT_WARPS];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pair = blockIdx.x * VIT_WARPS + warp, si = blockIdx.y, sf = blockIdx.z;
  VitWarpSmem&   sm   = sm_all[warp];

  const uint32_t cfi = info[sf].cfi;
  if (cfi < 1 || cfi > 3) return;
  const uint32_t nloc = c.nloc[cfi - 1];
  if (2 * pair >= nloc) return;
  const uint32_t nb = c.sizes[si], K = nb + 16, n3 = 3 * K;
  const float*   llr = llr_all + (size_t)sf * LLR_STRIDE;
  const uint16_t* tab = c.conv_tab[si];

  // ---- prologue: rate-dematch (accumulating), quantise -------------------------------------------
  bool     valid[2];
  uint32_t loc_i[2];
  for (int cd = 0; cd < 2; cd++) {
    loc_i[cd] = 2 * pair + cd;
    valid[cd] = loc_i[cd] < nloc;
    uint32_t ncce = 0, L = 0;
    if (valid[cd]) {
      const uint32_t e = c.loc_tab[cfi - 1][loc_i[cd]];
      ncce = e & 0xFFu, L = e >> 8;
      if (c.flags & LTEPHY_FLAG_SKIP_LOW_POWER)
        for (uint32_t i = ncce; i < ncce + (1u << L); i++)
          if (info[sf].cce_power[i] < 0.7f) valid[cd] = false;
    }
    const uint32_t E  = 72u << L;
    const float*   e  = llr + 72 * ncce;
    float          mx = 0.0f;
    for (uint32_t j = lane; j < n3; j += 32) {
      float acc = 0.0f;
      if (valid[cd])
        for (uint32_t k = j; k < E; k += n3) acc = acc + e[k];
      sm.rm[cd][tab[j]] = acc;
      mx                = fmaxf(mx, fabsf(acc));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (!(mx > 0.0f)) valid[cd] = false;
    const float gain = valid[cd] ? 32.0f / mx : 0.0f;
    __syncwarp();
    for (uint32_t j = lane; j < n3; j += 32) {
      float v = sm.rm[cd][j] * gain + 127.5f;
      v       = fminf(fmaxf(v, 0.0f), 255.0f);
      const int r = 2 * (int)v - 255;
      uint16_t* dst = reinterpret_cast<uint16_t*>(&sm.R[j / K][j % K]);
      dst[cd]       = (uint16_t)(short)r;
    }
    __syncwarp();
  }
  if (!valid[0] && !valid[1]) {
    if (lane < 2 && loc_i[lane] < LTEPHY_MAX_LOC) {
      ltephy_cand_t o{};
      cands[((size_t)sf * LTEPHY_MAX_LOC + loc_i[lane]) * LTEPHY_MAX_SIZES + si] = o;
    }
    return;
  }
  // branch-metric table for all 8 output patterns p = o0*4 + o1*2 + o2: bm = sum_i (o_i ? +r_i : -r_i)
  for (uint32_t k = lane; k < K; k += 32) {
    const uint32_t r0 = sm.R[0][k], r1 = sm.R[1][k], r2 = sm.R[2][k];
    const uint32_t n0 = __vneg2(r0), n1 = __vneg2(r1), n2 = __vneg2(r2);
#pragma unroll
    for (uint32_t pat = 0; pat < 8; pat++)
      sm.S[k][pat] = __vadd2(__vadd2((pat & 4u) ? r0 : n0, (pat & 2u) ? r1 : n1), (pat & 1u) ? r2 : n2);
  }
  __syncwarp();

  // ---- per-lane constants: lane j owns old states j and j+32, produces new states 2j and 2j+1 ------
  // state bit i = c_{k-1-i}; outputs for (state j < 32, input 0): parity(j & mask); polys 133,171,165
  const uint32_t o0 = __popc(lane & 0x36u) & 1u, o1 = __popc(lane & 0x27u) & 1u, o2 = __popc(lane & 0x2Bu) & 1u;
  const uint32_t pat = o0 * 4 + o1 * 2 + o2; // m = bm(j, input 0) = S[k][pat], -m = S[k][pat ^ 7]
  const uint32_t odd = lane & 1u, lo_half = lane < 16;
  const uint32_t src1 = odd ? 16 + (lane >> 1) : (lane >> 1);
  const uint32_t src2 = odd ? (lane >> 1) : 16 + (lane >> 1);
  const uint32_t* Sp = &sm.S[0][pat];
  const uint32_t* Sn = &sm.S[0][pat ^ 7u];

  uint32_t X0 = 0, X1 = 0; // packed path metrics of states j and j+32
  // one trellis step; STORE: keep the survivor decisions of this step
  auto step = [&](uint32_t k, uint4* dst, bool store) {
    const uint32_t m = Sp[k * 8], mn = Sn[k * 8];
    // new 2j   (input 0): max(X0 + m, X1 - m) ; new 2j+1 (input 1): max(X0 - m, X1 + m); ties keep the j branch
    bool           p0h, p0l, p1h, p1l;
    const uint32_t N0 = __vibmax_s16x2(__vadd2(X0, m), __vadd2(X1, mn), &p0h, &p0l);
    const uint32_t N1 = __vibmax_s16x2(__vadd2(X0, mn), __vadd2(X1, m), &p1h, &p1l);
    if (store) {
      uint4 d;
      d.x = __ballot_sync(0xffffffffu, !p0l); // cand0, new state 2j   -> bit j
      d.y = __ballot_sync(0xffffffffu, !p1l); // cand0, new state 2j+1
      d.z = __ballot_sync(0xffffffffu, !p0h); // cand1
      d.w = __ballot_sync(0xffffffffu, !p1h);
      if (lane == 0) *dst = d;
    }
    // re-distribute: lane j needs new[j], new[j+32]
    const uint32_t v1 = lo_half ? N0 : N1, v2 = lo_half ? N1 : N0;
    const uint32_t r1 = __shfl_sync(0xffffffffu, v1, src1), r2 = __shfl_sync(0xffffffffu, v2, src2);
    X0 = odd ? r2 : r1;
    X1 = odd ? r1 : r2;
  };
  auto renorm = [&]() {
    const uint32_t ref = __vneg2(__shfl_sync(0xffffffffu, X0, 0));
    X0 = __vadd2(X0, ref);
    X1 = __vadd2(X1, ref);
  };
  // three concatenated copies of the K-step frame; decisions are kept for copies 2 and 3
  for (uint32_t pass = 0; pass < 3; pass++) {
    uint4* dst = &sm.dec[(pass ? pass - 1 : 0) * K];
    uint32_t k = 0;
    for (; k + VIT_RENORM <= K; k += VIT_RENORM) {
      if (pass == 0) {
#pragma unroll
        for (uint32_t u = 0; u < VIT_RENORM; u++) step(k + u, nullptr, false);
      } else {
#pragma unroll
        for (uint32_t u = 0; u < VIT_RENORM; u++) step(k + u, dst + k + u, true);
      }
      renorm();
    }
    for (; k < K; k++) step(k, dst + k, pass != 0);
    renorm();
  }
  __syncwarp();
  // ---- best final state (lowest index on ties), per candidate ------------------------------------
  int best_s[2];
  for (int cd = 0; cd < 2; cd++) {
    int v0 = cd ? hi16(X0) : lo16(X0), v1 = cd ? hi16(X1) : lo16(X1);
    int bv = v0, bs = (int)lane;
    if (v1 > bv) bv = v1, bs = (int)lane + 32;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const int ov = __shfl_xor_sync(0xffffffffu, bv, off), os = __shfl_xor_sync(0xffffffffu, bs, off);
      if (ov > bv || (ov == bv && os < bs)) bv = ov, bs = os;
    }
    best_s[cd] = bs;
  }
  // ---- traceback (lanes 0,1: one candidate each), steps 3K-1 .. K, keep K..2K-1 --------------------
  if (lane < 2) {
    const int       cd = (int)lane;
    uint32_t        st = (uint32_t)best_s[cd];
    const uint32_t* dw = reinterpret_cast<const uint32_t*>(sm.dec) + cd * 2; // word (t-K)*4 + cd*2 + (st&1)
    auto back = [&](int idx) { // idx = t - K
      const uint32_t word = dw[idx * 4 + (int)(st & 1u)];
      st                  = (st >> 1) | (((word >> (st >> 1)) & 1u) << 5);
    };
    for (int idx = 2 * (int)K - 1; idx >= (int)K; idx--) back(idx); // third copy: only gives traceback depth
    // second copy: data bit i = idx (K-1 .. 0), bit i at position 31 - (i & 31) of word i >> 5
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    int      idx = (int)K - 1;
    for (; idx >= 64; idx--) {
      w2 |= (st & 1u) << (31 - (idx & 31));
      back(idx);
    }
    for (; idx >= 32; idx--) {
      w1 |= (st & 1u) << (31 - (idx & 31));
      back(idx);
    }
    for (; idx >= 0; idx--) {
      w0 |= (st & 1u) << (31 - (idx & 31));
      back(idx);
    }
    // CRC16 (poly 0x11021, zero init) over the first nb bits; RNTI = received parity ^ computed
    const unsigned long long lo64 = ((unsigned long long)w0 << 32) | (unsigned long long)w1; // bits 0..63, bit i at 63 - i
    uint32_t                 reg  = 0;
    for (uint32_t i = 0; i < nb + 16; i++) {
      const uint32_t bit = i < nb ? (uint32_t)((lo64 >> (63 - i)) & 1ull) : 0u;
      reg                = (reg << 1) | bit;
      if (reg & 0x10000u) reg ^= 0x11021u;
    }
    // the 16 parity bits nb .. nb+15 (bit i of the frame: word i>>5)
    uint32_t par = 0;
    for (uint32_t i = nb; i < nb + 16; i++) {
      const uint32_t wsel = (i >> 5) == 0 ? w0 : ((i >> 5) == 1 ? w1 : w2);
      par                 = (par << 1) | ((wsel >> (31 - (i & 31))) & 1u);
    }
    ltephy_cand_t o{};
    unsigned long long bits = lo64;
    if (nb < 64) bits &= ~((~0ull) >> nb);
    o.bits  = bits;
    o.rnti  = (uint16_t)((par ^ reg) & 0xFFFFu);
    o.valid = valid[cd] ? 1 : 0;
    if (!valid[cd]) o.bits = 0, o.rnti = 0;
    if (loc_i[cd] < LTEPHY_MAX_LOC) cands[((size_t)sf * LTEPHY_MAX_LOC + loc_i[cd]) * LTEPHY_MAX_SIZES + si] = o;
  }
}

extern "C" void launch_viterbi(const DevCell& c, const float* llr, const DevSfInfo* info, ltephy_cand_t* cands, uint32_t n, cudaStream_t st,
                               uint64_t* launches)
{
  const uint32_t max_pairs = (LTEPHY_MAX_LOC / 2 + VIT_WARPS - 1) / VIT_WARPS;
  dci_viterbi_kernel<<<dim3(max_pairs, c.nsizes, n), VIT_WARPS * 32, 0, st>>>(c, llr, info, cands);
  *launches += 1;
}

// ---------------------------------------------------------------------------------------------------
// Survivor selection (ltephy_compact_t, include/ltephy_b200.h): the RNTI-history-independent part of
// DCISearch::inspect_dci_location_recursively (src/src/DCISearch.cc:133-190) done for every entry of the table at once:
// sufficient-power rule (:473-489), srsran_pdcch_validate_location (falcon_pdcch.c:223-250), zero-RNTI and
// first-child-equals-parent (shortcut, :163-178) tests.  One CTA per subframe, one thread per location; the block-wide
// exclusive scan makes the list order (location, then size column) deterministic.
__device__ __forceinline__ bool dev_in_ue_space(uint32_t nof_cce, uint32_t ncce, uint32_t l, uint32_t Yk)
{
  const uint32_t L = 1u << l, M = l < 2 ? 6u : 2u;
  if (nof_cce < L || (ncce & (L - 1))) return false;
  const uint32_t n = nof_cce >> l, q = ncce >> l;
  if (q >= n) return false;
  return (q + n - Yk % n) % n < M;
}
__device__ __forceinline__ bool dev_in_common_space(uint32_t nof_cce, uint32_t ncce, uint32_t l)
{
  if (l < 2) return false;
  const uint32_t L = 1u << l;
  if (nof_cce < L || (ncce & (L - 1))) return false;
  const uint32_t lim = min(nof_cce, 16u) >> l, q = ncce >> l;
  return q < lim && q < (nof_cce >> l);
}
__device__ __forceinline__ uint32_t dev_validate_location(uint32_t nof_cce, uint32_t ncce, uint32_t l, uint32_t sf_idx, uint32_t rnti)
{
  bool ue = false, common = false;
  if (rnti >= 0x0001u && rnti <= 0x000Au)
    common = true;
  else if (rnti >= 0x000Bu && rnti <= 0xFFF3u)
    ue = common = true;
  else if (rnti >= 0xFFFDu)
    common = true;
  else
    return 0;
  uint32_t Yk = rnti;
  if (ue)
    for (uint32_t m = 0; m < sf_idx + 1; m++) Yk = (39827u * Yk) % 65537u;
  const bool valid = (ue && dev_in_ue_space(nof_cce, ncce, l, Yk)) || (common && dev_in_common_space(nof_cce, ncce, l));
  if (!valid) return 0;
  const bool amb = l > 0 && ((ue && dev_in_ue_space(nof_cce, ncce, l - 1, Yk)) || (common && dev_in_common_space(nof_cce, ncce, l - 1)));
  return amb ? 1u : 2u;
}

__global__ void __launch_bounds__(LTEPHY_MAX_LOC) cand_compact_kernel(const __grid_constant__ DevCell c, const DevSfInfo* __restrict__ info,
                                                                       const ltephy_cand_t* __restrict__ cands, ltephy_compact_t* __restrict__ out)
{
  static_assert(LTEPHY_MAX_LOC % 32 == 0 && LTEPHY_MAX_SIZES == 8, "layout");
  __shared__ uint32_t wtot[LTEPHY_MAX_LOC / 32];
  __shared__ uint8_t  smask[LTEPHY_MAX_LOC];
  const uint32_t sf = blockIdx.x, li = threadIdx.x, lane = li & 31u, warp = li >> 5;
  const uint32_t cfi = info[sf].cfi;
  const bool     ok  = cfi >= 1 && cfi <= 3;
  const uint32_t nloc = ok ? c.nloc[cfi - 1] : 0;
  uint4          e[LTEPHY_MAX_SIZES];
  uint32_t       mask = 0, my_L = 0, my_q = 0, lim = 0;
  if (li < nloc) {
    const uint32_t ent = c.loc_tab[cfi - 1][li], ncce = ent & 0xFFu, L = ent >> 8;
    const uint32_t ncce_sf = c.nof_cce[cfi - 1], sf_idx = info[sf].tti % 10;
    lim = min(ncce_sf, (uint32_t)LTEPHY_SEARCH_MAX_CCE), my_L = L, my_q = ncce >> L;
    bool           suff = true;
    for (uint32_t i = ncce; i < ncce + (1u << L); i++)
      if (i < lim && info[sf].cce_power[i] < 0.7f) suff = false;
    if (suff) {
      int par = -1; // location index of (L + 1, ncce): levels are laid out 3,2,1,0 with lim >> l entries each
      if (L < 3 && (ncce & ((2u << L) - 1u)) == 0) {
        const uint32_t lp = L + 1, q = ncce >> lp;
        uint32_t       base = 0;
        for (uint32_t l2 = 3; l2 > lp; l2--) base += lim >> l2;
        if (q < (lim >> lp) && base + q < nloc) par = (int)(base + q);
      }
      const uint4* row  = reinterpret_cast<const uint4*>(cands + ((size_t)sf * LTEPHY_MAX_LOC + li) * LTEPHY_MAX_SIZES);
      const uint4* prow = reinterpret_cast<const uint4*>(cands + ((size_t)sf * LTEPHY_MAX_LOC + (par >= 0 ? par : 0)) * LTEPHY_MAX_SIZES);
#pragma unroll
      for (uint32_t si = 0; si < LTEPHY_MAX_SIZES; si++) {
        if (si >= c.nsizes) continue;
        uint4          v  = row[si];
        const uint32_t vl = (v.z >> 16) & 0xFFu, r = vl ? (v.z & 0xFFFFu) : 0u;
        const uint32_t sm = dev_validate_location(ncce_sf, ncce, L, sf_idx, r);
        bool           eq = false;
        if (par >= 0) {
          const uint4 pv = prow[si];
          eq             = (((pv.z >> 16) & 0xFFu) ? (pv.z & 0xFFFFu) : 0u) == r;
        }
        if (sm == 0 && r != 0 && !eq) continue;
        mask |= 1u << si;
        if (!vl) v.x = 0, v.y = 0;
        v.z   = r | (vl << 16) | ((sm | ((r == 0) << 2) | ((uint32_t)eq << 3)) << 24);
        v.w   = li | (si << 8);
        e[si] = v;
      }
    }
  }
  const uint32_t cnt = __popc(mask);
  uint32_t       inc = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
    if ((int)lane >= off) inc += t;
  }
  if (lane == 31) wtot[warp] = inc;
  smask[li] = (uint8_t)mask;
  __syncthreads();
  uint32_t sub = mask; // union of the masks over the location's subtree: levels are laid out 3,2,1,0 with lim >> l entries each
  if (li < nloc)
    for (uint32_t l2 = 0; l2 < my_L; l2++) {
      uint32_t b = 0;
      for (uint32_t l3 = 3; l3 > l2; l3--) b += lim >> l3;
      const uint32_t span = 1u << (my_L - l2);
      for (uint32_t j = 0; j < span; j++) {
        const uint32_t q2 = my_q * span + j;
        if (q2 < (lim >> l2) && b + q2 < nloc) sub |= smask[b + q2];
      }
    }
  uint32_t base = 0, total = 0;
#pragma unroll
  for (uint32_t w = 0; w < LTEPHY_MAX_LOC / 32; w++) {
    if (w < warp) base += wtot[w];
    total += wtot[w];
  }
  ltephy_compact_t& o   = out[sf];
  uint32_t          pos = base + inc - cnt;
  ltephy_cloc_t     cl;
  cl.off = li < nloc ? (uint16_t)min(pos, 0xFFFFu) : (uint16_t)0, cl.mask = (uint8_t)mask, cl.pad = (uint8_t)sub;
  o.loc[li] = cl;
  if (li == 0) o.count = total, o.reserved = 0;
  uint4* lst = reinterpret_cast<uint4*>(o.list);
#pragma unroll
  for (uint32_t si = 0; si < LTEPHY_MAX_SIZES; si++)
    if ((mask >> si) & 1u) {
      if (pos < LTEPHY_COMPACT_CAP) lst[pos] = e[si];
      pos++;
    }
}

extern "C" void launch_compact(const DevCell& c, const DevSfInfo* info, const ltephy_cand_t* cands, ltephy_compact_t* out, uint32_t n, cudaStream_t st,
                               uint64_t* launches)
{
  cand_compact_kernel<<<n, LTEPHY_MAX_LOC, 0, st>>>(c, info, cands, out);
  *launches += 1;
}
