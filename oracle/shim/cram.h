// oracle/shim/cram.h -- TEST INFRASTRUCTURE. CRAM is not supported by the oracle shim;
// the reference only touches it when samFile::is_cram is set (read_chimeric_alignments.cpp:567-568),
// which the shim never sets.
#ifndef ARB_ORACLE_SHIM_CRAM_H
#define ARB_ORACLE_SHIM_CRAM_H
#include "sam.h"
#ifdef __cplusplus
extern "C" {
#endif
enum hts_fmt_option { CRAM_OPT_REFERENCE = 6 };
int cram_set_option(struct cram_fd* fd, enum hts_fmt_option opt, ...);
#ifdef __cplusplus
}
#endif
#endif
