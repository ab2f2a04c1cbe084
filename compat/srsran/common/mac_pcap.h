// srsran/common/mac_pcap.h (compat): placeholder -- the MAC pcap writer is outside the hot path; libltephy_b200 writes the
// reference's pcap / DCI-trace formats from its result buffers (include/ltephy_sinks.h)
#ifndef SRSRAN_MAC_PCAP_H
#define SRSRAN_MAC_PCAP_H
#endif
