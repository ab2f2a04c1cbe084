/*
 * ltephy_sinks.h -- result back-end wire formats, host only: the decoded transport blocks and the
 * accepted DCIs of a batch written in the byte formats LTESniffer's own sinks produce, straight from the result buffers of
 * ltephy_decode_subframes / ltephy_get_ul.
 *
 *  - MAC-LTE pcap (DLT 147): LTESniffer_pcap_writer::pack_and_write -> srsRAN LTE_PCAP_MAC_WritePDU
 *    (reference src/src/PcapWriter.cc:93-118).  The record layout is pinned by the reference's own example captures
 *    pcap_file_example/{ltesniffer_dl_mode,ltesniffer_ul_mode,api_collector}.pcap: tests/test_sinks.py re-writes every record
 *    of those files through this API and requires identical bytes.
 *  - DCI trace line: DCIToFile::printDCICollection (reference src/src/SubframeInfoConsumer.cc:66-138).
 */
#ifndef LTEPHY_SINKS_H
#define LTEPHY_SINKS_H
#include "ltephy_search.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltephy_pcap ltephy_pcap_t;

/* rntiType of the MAC-LTE context (wireshark packet-mac-lte.h) */
enum { LTEPHY_RNTI_NONE = 0, LTEPHY_RNTI_P = 1, LTEPHY_RNTI_RA = 2, LTEPHY_RNTI_C = 3, LTEPHY_RNTI_SI = 4 };
enum { LTEPHY_DIR_UL = 0, LTEPHY_DIR_DL = 1 };

/* creates / truncates the file and writes the pcap global header (magic a1b2c3d4, v2.4, snaplen 65535, network 147) */
ltephy_pcap_t* ltephy_pcap_open(const char* path);
void           ltephy_pcap_close(ltephy_pcap_t* p);
/* one MAC PDU: context {FDD, direction, rntiType, rnti, ueid, sfn<<4|sf, crc status, carrier 0, nb-iot 0} + payload */
int ltephy_pcap_write(ltephy_pcap_t* p, const uint8_t* pdu, uint32_t len, uint16_t rnti, uint8_t rnti_type, uint8_t direction, uint32_t tti,
                      int crc_ok, uint16_t ue_id, uint32_t ts_sec, uint32_t ts_usec);
/* SI-RNTI / P-RNTI / RA-RNTI (1..10 at FDD) / C-RNTI, as PDSCH_Decoder::decode_dl_mode sorts them (src/src/DL_Sniffer_PDSCH.cc:912-940) */
uint8_t ltephy_rnti_type(uint16_t rnti);
/* every CRC-passing transport block of a ltephy_decode_subframes result (write_pcap is only reached with crc && len > 0,
 * DL_Sniffer_PDSCH.cc:1127-1131); tti[sf] = 10*sfn + sf_idx of the batch's subframes.  Returns the number of records written. */
int ltephy_pcap_write_dl_batch(ltephy_pcap_t* p, const uint32_t* tti, const ltephy_dci_t* dcis, uint32_t n_dcis, const ltephy_tb_result_t* tbs,
                               const uint8_t* payload, uint16_t ue_id, uint32_t ts_sec, uint32_t ts_usec);
/* same for the results of ltephy_get_ul (write_ul_crnti) */
int ltephy_pcap_write_ul_batch(ltephy_pcap_t* p, const uint32_t* tti, const ltephy_ul_grant_t* grants, uint32_t n_grants, const ltephy_tb_result_t* res,
                               const uint8_t* payload, uint16_t ue_id, uint32_t ts_sec, uint32_t ts_usec);

/* one line of the DCI trace file for an accepted DCI (DL formats and format 0); returns the line length or a negative code.
 * use_256qam_table selects the MCS table for the TBS columns of a DL DCI. */
int ltephy_dci_trace_line(const ltephy_search_t* s, const ltephy_dci_t* dci, uint32_t tti, uint32_t cfi, int use_256qam_table, uint32_t ts_sec,
                          uint32_t ts_usec, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
