// srsran/asn1/rrc/dl_ccch_msg.h (compat): RRCConnectionSetup-r8-IEs as an opaque value (src/include/ULSchedule.h:125 stores it)
#ifndef SRSASN1_RRC_DL_CCCH_MSG_H
#define SRSASN1_RRC_DL_CCCH_MSG_H
#include <cstdint>
namespace asn1 {
namespace rrc {
struct rrc_conn_setup_r8_ies_s {
  struct {
    bool phys_cfg_ded_present = false;
    struct {
      bool cqi_report_cfg_present = false, pusch_cfg_ded_present = false;
      struct { uint8_t beta_offset_ack_idx = 0, beta_offset_ri_idx = 0, beta_offset_cqi_idx = 0; } pusch_cfg_ded;
    } phys_cfg_ded;
  } rr_cfg_ded;
};
} // namespace rrc
} // namespace asn1
#endif
