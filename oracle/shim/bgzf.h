// oracle/shim/bgzf.h -- TEST INFRASTRUCTURE. gzip/BGZF stream reader over system zlib,
// standing in for htslib's <bgzf.h> (reference call sites: read_compressed_file.cpp:22-37).
#ifndef ARB_ORACLE_SHIM_BGZF_H
#define ARB_ORACLE_SHIM_BGZF_H
#include <stddef.h>
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct BGZF BGZF;
BGZF* bgzf_open(const char* path, const char* mode);
ssize_t bgzf_read(BGZF* fp, void* data, size_t length);
int bgzf_close(BGZF* fp);
#ifdef __cplusplus
}
#endif
#endif
