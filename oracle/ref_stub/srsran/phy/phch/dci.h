/* Stub so that the reference's lib/include/falcon/util/rnti_manager_c.h (which includes this srsRAN
 * header without using anything from it) compiles without the absent srsRAN tree.  Oracle build only. */
#pragma once
