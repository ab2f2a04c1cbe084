/* srsran/srsran.h (compat): umbrella of the hot-path headers libltephy_srsran_compat provides */
#ifndef SRSRAN_H
#define SRSRAN_H
#include "srsran/config.h"
#include "srsran/phy/ch_estimation/chest_dl.h"
#include "srsran/phy/ch_estimation/chest_ul.h"
#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/enb/enb_ul.h"
#include "srsran/phy/fec/softbuffer.h"
#include "srsran/phy/phch/dci.h"
#include "srsran/phy/phch/pdcch.h"
#include "srsran/phy/phch/pdsch.h"
#include "srsran/phy/phch/pusch.h"
#include "srsran/phy/phch/ra.h"
#include "srsran/phy/phch/ra_dl.h"
#include "srsran/phy/phch/ra_ul.h"
#include "srsran/phy/ue/ue_dl.h"
#include "srsran/phy/utils/bit.h"
#include "srsran/phy/utils/debug.h"
#include "srsran/phy/utils/vector.h"
#endif
