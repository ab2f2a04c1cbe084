// ltephy_internal.cuh -- the PHY handle and its helpers, shared by ltephy_capi.cu (single-GPU entry points) and shard.cu
// (sharded operation).  Private to libltephy_b200: nothing here is part of the C-ABI.
#pragma once
#include "../../include/ltephy_b200.h"
#include "../../include/ltephy_search.h"
#include "dev_common.cuh"
#include "dev_ul.cuh"
#include "lte_host.hpp"
#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <tuple>
#include <vector>

extern thread_local std::string ltephy_g_err;
static inline int fail(int code, const char* fmt, ...)
{
  char    buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  ltephy_g_err = buf;
  return code;
}
#define CU(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) return fail(LTEPHY_ERROR, "%s: %s", #x, cudaGetErrorString(e_));        \
  } while (0)

template <typename T>
struct DevBuf { // growable device buffer
  T*     p   = nullptr;
  size_t cap = 0;
  int    reserve(size_t n)
  {
    if (n <= cap) return 0;
    size_t want = n + n / 4 + 1024;
    if (p) {
      cudaDeviceSynchronize();
      cudaFree(p);
      p = nullptr;
    }
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) {
      cap = 0;
      return -1;
    }
    cap = want;
    return 0;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr, cap = 0;
  }
};
template <typename T>
struct PinBuf { // growable pinned host buffer
  T*     p   = nullptr;
  size_t cap = 0;
  int    reserve(size_t n)
  {
    if (n <= cap) return 0;
    size_t want = n + n / 4 + 1024;
    if (p) cudaFreeHost(p);
    p = nullptr;
    if (cudaMallocHost(&p, want * sizeof(T)) != cudaSuccess) {
      cap = 0;
      return -1;
    }
    cap = want;
    return 0;
  }
  void release()
  {
    if (p) cudaFreeHost(p);
    p = nullptr, cap = 0;
  }
};

struct ltephy {
  ltephy_cfg_t       cfg{};
  ltehost::Cell      cell;
  ltehost::CtrlMap   cm;
  ltehost::SizeTable st;
  DevCell            dc{};
  cudaStream_t       stream = nullptr;
  cudaEvent_t        ev[6]{}, mark[2]{}, ev_h2d = nullptr;
  DevBuf<float2>     d_cfo;                  // per-sample rotation of ltephy_set_cfo
  DevBuf<uint32_t>   d_vwork;                // work list of the Viterbi kernel: count + (subframe, pair, size) items
  DevBuf<uint32_t>   d_mib;                  // 4 words per subframe (pbch_kernel)
  DevBuf<short>      d_harq;                 // HARQ store: harq_slots x LTEPHY_HARQ_SLOT_BYTES
  uint32_t           harq_slots = 0, harq_max_gen = 0;
  std::map<uint32_t, uint32_t> harq_uses;    // slot -> uses in the batch being built
  std::vector<void*> tables; // device tables freed at destroy
  uint64_t           launches = 0;

  // phase A
  DevBuf<float2>        d_iq, d_sym, d_pil, d_ce; // d_pil: smoothed CRS estimates [n][port][rx][4][2 nof_prb]; d_ce: interpolated grid, filled on demand (tap)
  DevBuf<float>         d_llr;
  DevBuf<DevSfInfo>     d_info;
  DevBuf<ltephy_cand_t> d_cands;
  PinBuf<DevSfInfo>     h_info;
  DevBuf<ltephy_compact_t> d_compact;
  PinBuf<ltephy_compact_t> h_compact;
  uint32_t              n_cur = 0;
  std::vector<uint8_t>  re_cnt; // [3 sf class][3 cfi][14][nof_prb]

  // phase B
  std::vector<DevGrant> grants;
  std::vector<DevCb>    cbs;
  std::vector<DevPair>  pairs;
  std::vector<DevTb>    tbs;
  std::vector<uint32_t> pair_pi_off;
  std::vector<uint32_t> tb_slot; // result slot [grant*2 + tb] -> tb index or ~0
  DevBuf<DevGrant>      d_grants;
  DevBuf<DevCb>         d_cbs;
  DevBuf<DevPair>       d_pairs;
  DevBuf<DevTb>         d_tbs;
  DevBuf<uint32_t>      d_pair_pi_off;
  DevBuf<uint32_t>      d_seq, d_rm, d_turbo;
  DevBuf<uint32_t>      d_tscratch, d_tqueue; // persistent turbo kernel: per-CTA state slots, pair counters of the launches
  DevBuf<short>         d_pllr;
  DevBuf<uint16_t>      d_pi;
  DevBuf<uint8_t>       d_payload, d_cb_iters, d_cb_crc;
  DevBuf<ltephy_tb_result_t> d_res;
  PinBuf<ltephy_tb_result_t> h_res;
  PinBuf<uint8_t>            h_payload;
  PinBuf<uint8_t>            h_stage;      // pinned arena for the job descriptors of one phase B (see pull())
  size_t                     stage_used = 0;
  bool                       stage_busy = false; // descriptors staged and possibly still being pulled
  size_t                     payload_bytes = 0, pllr_elems = 0;
  uint32_t *                 d_gold_x1 = nullptr, *d_gold_basis = nullptr, gold_words = 0;
  uint32_t *                 d_xpowA = nullptr, *d_xpowB = nullptr;
  std::map<std::tuple<uint32_t, uint32_t, uint32_t>, std::pair<uint32_t, uint32_t>> rm_cache; // (K,F,rv) -> (offset, nn)
  int64_t  rm_fast[188][4];   // F == 0 fast path: offset or -1
  uint32_t rm_fast_nn[188][4];
  int64_t  pi_fast[188];
  std::vector<ltehost::Segm> segm_fast; // index tbs/8, C == 0 means "not computed"
  // uplink
  ltephy_ul_cfg_t            ulcfg{};
  bool                       ulcfg_set = false;
  uint32_t                   n_prs[20]{}, ul_u[20]{}, ul_v[20]{}; // per slot: n_PRS, base-sequence group u and sequence v
  DevBuf<float2>             d_uliq, d_ulsym, d_ulpool; // d_ulpool: DMRS sequences and IDFT twiddles
  DevBuf<DevUlGrant>         d_ulgrants;
  DevBuf<DevUlChest>         d_ulchest;
  PinBuf<DevUlChest>         h_ulchest;
  std::vector<DevUlGrant>    ulgrants;
  std::map<uint64_t, uint32_t> ul_tab_cache; // (kind, M, ncs) -> offset in d_ulpool
  size_t                     ulpool_used = 0;
  uint32_t                   n_ul = 0;
  size_t                                                                            rm_used = 0;
  std::map<uint32_t, uint32_t>                                                      pi_cache; // K -> offset
  size_t                                                                            pi_used = 0;
  std::map<uint32_t, ltehost::Segm>                                                 segm_cache;
  float                                                                             t_ms[4]{};
};


// kernel-driven copy of a few bytes from pinned host memory (see ltephy_capi.cu: small transfers avoid the copy engine)
void ltephy_pull(ltephy* h, void* dst_dev, const void* src_pinned, size_t bytes, cudaStream_t st);
