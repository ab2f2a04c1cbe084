// harq.cpp -- host bookkeeping of the HARQ soft-combining mode: which transport block of which C-RNTI is a new transmission, a
// retransmission to combine, or already decoded.  Restates HARQ::is_retransmission / updateProcess / updateHARQRNTI (reference
// src/src/HARQ.cc:60-188) as they are driven from PDSCH_Decoder::decode_dl_mode (src/src/DL_Sniffer_PDSCH.cc:942-1018), without the
// per-process mutexes: batches are prepared by one thread in subframe order, so DL_SNIFFER_HARQ_BUSY cannot occur.
// The soft values themselves live on the GPU (ltephy_harq_reserve, rm_turbo_rx_kernel); this object only hands out slot numbers.
#include "../../include/ltephy_b200.h"
#include "../../include/ltephy_search.h"
#include <cstring>
#include <vector>

namespace {
struct TbState { // dl_sniffer_harq_tb_t + dl_sniffer_harq_grant_t (src/include/HARQ.h:48-75)
  uint32_t tti = 0;
  uint8_t  ndi = 0, rv = 0;
  int32_t  tbs = 0;
  bool     is_first_transmission = true, last_decoded = false;
};
struct Entity {
  uint16_t rnti = 0;
  TbState  tb[8][2];
};
} // namespace
struct ltephy_harq {
  std::vector<Entity> db;
};

extern "C" ltephy_harq_t* ltephy_harq_create(uint32_t max_rnti)
{
  if (max_rnti == 0 || max_rnti > 65536) return nullptr;
  ltephy_harq* q = new ltephy_harq();
  q->db.resize(max_rnti);
  return q;
}
extern "C" void ltephy_harq_destroy(ltephy_harq_t* q) { delete q; }

static Entity* find(ltephy_harq* q, uint16_t rnti, Entity** avail)
{
  Entity* hit = nullptr;
  *avail      = nullptr;
  for (Entity& e : q->db) { // HARQ.cc:83-90: the last match wins, the last free entity is the one handed out
    if (e.rnti == rnti)
      hit = &e;
    else if (e.rnti == 0)
      *avail = &e;
  }
  return hit;
}

extern "C" int ltephy_harq_classify(ltephy_harq_t* q, uint16_t rnti, uint32_t pid, uint32_t tb, uint32_t ndi, int32_t tbs, uint32_t tti, uint32_t* slot)
{
  if (!q || pid >= 8 || tb >= 2 || rnti == 0) return LTEPHY_ERROR_INVALID_INPUTS;
  Entity *avail, *e = find(q, rnti, &avail);
  if (!e) {
    if (!avail) return LTEPHY_HARQ_FULL_BUFFER;
    avail->rnti = rnti; // HARQ.cc:91-95
    if (slot) *slot = (uint32_t)(((avail - q->db.data()) * 8 + pid) * 2 + tb);
    return LTEPHY_HARQ_NEW_TX;
  }
  if (slot) *slot = (uint32_t)(((e - q->db.data()) * 8 + pid) * 2 + tb);
  const TbState& s = e->tb[pid][tb];
  const uint32_t d = (tti + 10240u - s.tti) % 10240u; // comparetti (HARQ.cc:60-68): a retransmission comes exactly 8 ms later
  if (d != 8) return LTEPHY_HARQ_NEW_TX;
  if (ndi != s.ndi || s.is_first_transmission || s.tbs != tbs) return LTEPHY_HARQ_NEW_TX; // ndi_present is always set (DL_Sniffer_PDSCH.cc:950)
  return s.last_decoded ? LTEPHY_HARQ_DECODED : LTEPHY_HARQ_RE_TX;
}

extern "C" void ltephy_harq_update(ltephy_harq_t* q, uint16_t rnti, uint32_t pid, uint32_t tb, uint32_t ndi, uint32_t rv, int32_t tbs, uint32_t tti, int decoded)
{
  if (!q || pid >= 8 || tb >= 2) return;
  Entity *avail, *e = find(q, rnti, &avail);
  if (!e) return; // updateHARQRNTI: unknown RNTIs are not added (HARQ.cc:184-186)
  TbState& s = e->tb[pid][tb];
  s.tti = tti % 10240u, s.ndi = (uint8_t)ndi, s.rv = (uint8_t)rv, s.tbs = tbs, s.is_first_transmission = false, s.last_decoded = decoded != 0;
}

// HARQ::getlastTbs (HARQ.cc:262-274): size of the last transmission recorded for (rnti, pid, tb); 0 when the RNTI has no entity yet
extern "C" int32_t ltephy_harq_last_tbs(ltephy_harq_t* q, uint16_t rnti, uint32_t pid, uint32_t tb)
{
  if (!q || pid >= 8 || tb >= 2) return 0;
  Entity *avail, *e = find(q, rnti, &avail);
  return e ? e->tb[pid][tb].tbs : 0;
}

extern "C" int ltephy_harq_prepare_grant(ltephy_harq_t* q, const ltephy_dci_fields_t* f, uint32_t tti, ltephy_grant_t* g, int status[2])
{
  if (!q || !f || !g || !status) return LTEPHY_ERROR_INVALID_INPUTS;
  for (int t = 0; t < 2; t++) {
    status[t] = -1;
    if (!g->tb[t].enabled) continue;
    // a reserved MCS (29-31, or 28-31 of the 256QAM table) carries no size: ltephy_dci_to_grant leaves tbs = 0 and the reference takes the size of
    // the process' last transmission (DCICollection::addCandidate, src/src/DCICollection.cc:236-252).  Still 0: nothing known, the block is not decoded
    if (g->tb[t].tbs == 0) g->tb[t].tbs = ltephy_harq_last_tbs(q, g->rnti, f->harq_pid, (uint32_t)t);
    if (g->tb[t].tbs <= 0) continue;
    uint32_t slot = 0;
    status[t]     = ltephy_harq_classify(q, g->rnti, f->harq_pid, (uint32_t)t, f->ndi[t], g->tb[t].tbs, tti, &slot);
    g->tb[t].harq_op = LTEPHY_HARQ_NONE, g->tb[t].harq_slot = 0;
    if (status[t] == LTEPHY_HARQ_NEW_TX)
      g->tb[t].harq_op = LTEPHY_HARQ_NEW, g->tb[t].harq_slot = slot;
    else if (status[t] == LTEPHY_HARQ_RE_TX)
      g->tb[t].harq_op = LTEPHY_HARQ_RETX, g->tb[t].harq_slot = slot;
  }
  // already decoded: pdsch_cfg->grant.tb[i].enabled = false (DL_Sniffer_PDSCH.cc:970-972).  Only when the block travels alone: switching one of two
  // codewords off would change the layer de-mapping of the other one, so a decoded block of a two-codeword grant is simply decoded again
  const int n_en = (g->tb[0].enabled ? 1 : 0) + (g->tb[1].enabled ? 1 : 0);
  for (int t = 0; t < 2; t++)
    if (status[t] == LTEPHY_HARQ_DECODED && n_en == 1) g->tb[t].enabled = 0;
  return LTEPHY_SUCCESS;
}
