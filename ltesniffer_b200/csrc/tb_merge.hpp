// tb_merge.hpp -- placement of the transport-block results of one grant into the per-DCI result table (private helper of
// ltephy_decode_subframes and ltephy_decode_subframes_sharded).
#pragma once
#include "../../include/ltephy_b200.h"
#include "../../include/ltephy_search.h"

// One result pair per DCI: the reading of the 64QAM MCS table unless only the 256QAM-table reading (the speculative second
// grant of the same DCI, flagged LTEPHY_GRANT_ALT_TABLE) passes a CRC, in which case its transport blocks are reported with
// crc = 2 (the batched form of src/src/DL_Sniffer_PDSCH.cc:1089-1210).  Grants of one DCI are adjacent, primary first.
static inline void ltephy_place_grant_result(ltephy_tb_result_t* tbs, uint32_t grant_dci, const ltephy_tb_result_t& r0, const ltephy_tb_result_t& r1)
{
  const uint32_t di  = grant_dci & ~LTEPHY_GRANT_ALT_TABLE;
  const bool     alt = (grant_dci & LTEPHY_GRANT_ALT_TABLE) != 0;
  if (!alt) {
    tbs[2 * di] = r0, tbs[2 * di + 1] = r1;
    return;
  }
  const bool have_primary = tbs[2 * di].payload_len || tbs[2 * di + 1].payload_len;
  if (have_primary && (tbs[2 * di].crc || tbs[2 * di + 1].crc)) return;
  if (!have_primary || r0.crc || r1.crc) {
    tbs[2 * di] = r0, tbs[2 * di + 1] = r1;
    for (int t = 0; t < 2; t++)
      if (tbs[2 * di + t].crc) tbs[2 * di + t].crc = 2;
  }
}
