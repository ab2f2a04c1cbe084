// srsran/asn1/rrc/bcch_msg.h (compat): the fields of SystemInformationBlockType2 the reference's PHY-facing code reads
// (src/src/ULSchedule.cc:140-158).  The ASN.1 decoder itself is outside the hot path (out of scope).
#ifndef SRSASN1_RRC_BCCH_MSG_H
#define SRSASN1_RRC_BCCH_MSG_H
#include <cstdint>
namespace asn1 {
namespace rrc {
struct ul_ref_sigs_pusch_s {
  bool    group_hop_enabled = false;
  uint8_t group_assign_pusch = 0;
  bool    seq_hop_enabled    = false;
  uint8_t cyclic_shift       = 0;
};
struct pusch_cfg_common_s {
  struct pusch_cfg_basic_s_ {
    uint8_t n_sb = 1;
    int     hop_mode = 0;
    uint8_t pusch_hop_offset = 0;
    bool    enable64_qam = false;
  } pusch_cfg_basic;
  ul_ref_sigs_pusch_s ul_ref_sigs_pusch;
};
struct prach_cfg_info_s {
  uint8_t prach_cfg_idx = 0;
  bool    high_speed_flag = false;
  uint8_t zero_correlation_zone_cfg = 0;
  uint8_t prach_freq_offset = 0;
};
struct prach_cfg_sib_s {
  uint16_t         root_seq_idx = 0;
  prach_cfg_info_s prach_cfg_info;
};
struct rr_cfg_common_sib_s {
  prach_cfg_sib_s    prach_cfg;
  pusch_cfg_common_s pusch_cfg_common;
};
struct sib_type2_s {
  rr_cfg_common_sib_s rr_cfg_common;
};
} // namespace rrc
} // namespace asn1
#endif
