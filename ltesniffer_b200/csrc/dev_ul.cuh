// dev_ul.cuh -- device descriptors of the PUSCH path
#pragma once
#include "dev_common.cuh"

struct DevUlGrant {
  uint32_t sf, sf_idx, rnti;
  uint32_t M, k0, qm;     // M_sc = 12 L_prb, first subcarrier
  uint32_t dmrs_off[2];   // offsets (float2) of the two slots' DMRS sequences in the DMRS pool
  uint32_t idft_off;      // offset of exp(+j 2 pi m / M), m < M
  uint32_t llr_off;       // int16 offset of the codeword in the LLR pool
  uint32_t scr_off;       // word offset of the scrambling sequence
};
