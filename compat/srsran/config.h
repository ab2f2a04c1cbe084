/* compat/srsran/config.h -- part of the srsRAN-compatible header tree of libltephy_srsran_compat.
 * These headers declare, at srsRAN's own include paths, the types and functions the reference's PHY-facing sources use
 * (SURVEY.md section 8b / Appendix A), so that those sources compile UNMODIFIED against the B200 library:
 *   lib/src/phy/falcon_phch/{falcon_pdcch,falcon_dci,dl_sniffer_pdsch,ul_sniffer_pusch}.c, lib/src/phy/falcon_ue/falcon_ue_dl.c,
 *   src/src/{DCISearch,MetaFormats,SubframePower}.cc ...
 * Layout and field names follow srsRAN 21.10 (the release the reference pins, external/cmake/srsRAN.CMakeLists.txt.in:8-9);
 * the binary layout is this library's own -- the reference is recompiled against these headers, not linked to a stock libsrsran. */
#ifndef SRSRAN_CONFIG_H
#define SRSRAN_CONFIG_H
#define SRSRAN_API __attribute__((visibility("default")))
#define SRSRAN_LOCAL __attribute__((visibility("hidden")))
#include <complex.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#ifdef __cplusplus
#include <complex>
typedef std::complex<float> cf_t;
#undef I
#else
typedef _Complex float cf_t;
#endif
#endif
