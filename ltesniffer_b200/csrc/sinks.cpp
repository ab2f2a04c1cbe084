// sinks.cpp -- result back-end wire formats (include/ltephy_sinks.h), host only.
// MAC-LTE pcap framing as LTESniffer_pcap_writer::pack_and_write hands it to srsRAN's LTE_PCAP_MAC_WritePDU (reference
// src/src/PcapWriter.cc:93-118).  srsRAN is absent from the reference tree, so the byte layout is restated from the reference's
// own example captures (pcap_file_example/*.pcap), which tests/test_sinks.py reproduces bit for bit:
//   pcap file header: magic a1b2c3d4, version 2.4, thiszone 0, sigfigs 0, snaplen 65535, network 147 (DLT_USER0 = MAC_LTE_DLT)
//   record header:    ts_sec, ts_usec, incl_len, orig_len (all uint32, host order = little endian in the fixtures)
//   context:          radioType (1 = FDD), direction (0 UL / 1 DL), rntiType,
//                     02 rnti(be16), 03 ueid(be16), 04 (sfn << 4 | sf)(be16), 07 crcStatus, 0a carrierId, 0f nbIotMode, 01 = payload follows
#include "../../include/ltephy_sinks.h"
#include <cstdio>
#include <cstring>

struct ltephy_pcap {
  FILE* f;
};

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16), p[3] = (uint8_t)(v >> 24); }

extern "C" ltephy_pcap_t* ltephy_pcap_open(const char* path)
{
  if (!path) return nullptr;
  FILE* f = fopen(path, "wb");
  if (!f) return nullptr;
  uint8_t h[24];
  put32(h, 0xa1b2c3d4u);
  h[4] = 2, h[5] = 0, h[6] = 4, h[7] = 0; // version 2.4
  put32(h + 8, 0), put32(h + 12, 0), put32(h + 16, 65535), put32(h + 20, 147);
  if (fwrite(h, 1, sizeof(h), f) != sizeof(h)) {
    fclose(f);
    return nullptr;
  }
  ltephy_pcap* p = new ltephy_pcap{f};
  return p;
}
extern "C" void ltephy_pcap_close(ltephy_pcap_t* p)
{
  if (!p) return;
  if (p->f) fclose(p->f);
  delete p;
}
extern "C" int ltephy_pcap_write(ltephy_pcap_t* p, const uint8_t* pdu, uint32_t len, uint16_t rnti, uint8_t rnti_type, uint8_t direction, uint32_t tti,
                                 int crc_ok, uint16_t ue_id, uint32_t ts_sec, uint32_t ts_usec)
{
  if (!p || !p->f || (!pdu && len)) return LTEPHY_ERROR_INVALID_INPUTS;
  uint8_t        c[16 + 19];
  const uint32_t sfn_sf = (((tti / 10) & 0xFFFu) << 4) | (tti % 10);
  put32(c, ts_sec), put32(c + 4, ts_usec), put32(c + 8, 19 + len), put32(c + 12, 19 + len);
  uint8_t* x = c + 16;
  x[0] = 1, x[1] = direction, x[2] = rnti_type;
  x[3] = 0x02, x[4] = (uint8_t)(rnti >> 8), x[5] = (uint8_t)rnti;
  x[6] = 0x03, x[7] = (uint8_t)(ue_id >> 8), x[8] = (uint8_t)ue_id;
  x[9] = 0x04, x[10] = (uint8_t)(sfn_sf >> 8), x[11] = (uint8_t)sfn_sf;
  x[12] = 0x07, x[13] = crc_ok ? 1 : 0;
  x[14] = 0x0a, x[15] = 0;
  x[16] = 0x0f, x[17] = 0;
  x[18] = 0x01;
  if (fwrite(c, 1, sizeof(c), p->f) != sizeof(c)) return LTEPHY_ERROR;
  if (len && fwrite(pdu, 1, len, p->f) != len) return LTEPHY_ERROR;
  return LTEPHY_SUCCESS;
}
extern "C" uint8_t ltephy_rnti_type(uint16_t rnti)
{
  if (rnti == 0xFFFF) return LTEPHY_RNTI_SI;
  if (rnti == 0xFFFE) return LTEPHY_RNTI_P;
  if (rnti > 0x0001 && rnti < 0x000A) return LTEPHY_RNTI_RA; // PDSCH_Decoder::rnti_name, src/src/DL_Sniffer_PDSCH.cc:1398-1418 (bounds exclusive)
  return LTEPHY_RNTI_C;
}
extern "C" int ltephy_pcap_write_dl_batch(ltephy_pcap_t* p, const uint32_t* tti, const ltephy_dci_t* dcis, uint32_t n_dcis, const ltephy_tb_result_t* tbs,
                                          const uint8_t* payload, uint16_t ue_id, uint32_t ts_sec, uint32_t ts_usec)
{
  if (!p || !tti || (!dcis && n_dcis) || !tbs || !payload) return LTEPHY_ERROR_INVALID_INPUTS;
  int n = 0;
  for (uint32_t i = 0; i < n_dcis; i++)
    for (uint32_t t = 0; t < 2; t++) {
      const ltephy_tb_result_t& r = tbs[2 * i + t];
      if (!r.crc || !r.payload_len) continue;
      const int rc = ltephy_pcap_write(p, payload + r.payload_off, r.payload_len, dcis[i].rnti, ltephy_rnti_type(dcis[i].rnti), LTEPHY_DIR_DL,
                                       tti[dcis[i].sf], 1, ue_id, ts_sec, ts_usec);
      if (rc) return rc;
      n++;
    }
  return n;
}
extern "C" int ltephy_pcap_write_ul_batch(ltephy_pcap_t* p, const uint32_t* tti, const ltephy_ul_grant_t* grants, uint32_t n_grants,
                                          const ltephy_tb_result_t* res, const uint8_t* payload, uint16_t ue_id, uint32_t ts_sec, uint32_t ts_usec)
{
  if (!p || !tti || (!grants && n_grants) || !res || !payload) return LTEPHY_ERROR_INVALID_INPUTS;
  int n = 0;
  for (uint32_t i = 0; i < n_grants; i++) {
    if (!res[i].crc || !res[i].payload_len) continue;
    const int rc = ltephy_pcap_write(p, payload + res[i].payload_off, res[i].payload_len, grants[i].rnti, LTEPHY_RNTI_C, LTEPHY_DIR_UL, tti[grants[i].sf],
                                     1, ue_id, ts_sec, ts_usec);
    if (rc) return rc;
    n++;
  }
  return n;
}
