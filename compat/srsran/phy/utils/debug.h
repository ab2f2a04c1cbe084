/* srsran/phy/utils/debug.h (compat): verbosity macros */
#ifndef SRSRAN_DEBUG_H
#define SRSRAN_DEBUG_H
#include "srsran/config.h"
#define SRSRAN_VERBOSE_DEBUG 3
#define SRSRAN_VERBOSE_INFO 2
#define SRSRAN_VERBOSE_WARN 1
#define SRSRAN_VERBOSE_NONE 0
#ifdef __cplusplus
extern "C" {
#endif
SRSRAN_API extern int srsran_verbose;
SRSRAN_API void  get_time_interval(struct timeval* tdata);
#ifdef __cplusplus
}
#endif
#define SRSRAN_VERBOSE_ISINFO() (srsran_verbose >= SRSRAN_VERBOSE_INFO)
#define SRSRAN_VERBOSE_ISDEBUG() (srsran_verbose >= SRSRAN_VERBOSE_DEBUG)
#define SRSRAN_VERBOSE_ISNONE() (srsran_verbose == SRSRAN_VERBOSE_NONE)
#define PRINT_DEBUG srsran_verbose = SRSRAN_VERBOSE_DEBUG
#define PRINT_INFO srsran_verbose = SRSRAN_VERBOSE_INFO
#define PRINT_WARN srsran_verbose = SRSRAN_VERBOSE_WARN
#define PRINT_NONE srsran_verbose = SRSRAN_VERBOSE_NONE
#define DEBUG(_fmt, ...) do { if (SRSRAN_VERBOSE_ISDEBUG()) fprintf(stdout, "[DEBUG]: " _fmt "\n", ##__VA_ARGS__); } while (0)
#define INFO(_fmt, ...) do { if (SRSRAN_VERBOSE_ISINFO()) fprintf(stdout, "[INFO]: " _fmt "\n", ##__VA_ARGS__); } while (0)
#define ERROR(_fmt, ...) do { fprintf(stderr, "\e[31m%s:%d: " _fmt "\e[0m\n", __FILE__, __LINE__, ##__VA_ARGS__); } while (0)
#endif
