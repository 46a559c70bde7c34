/* See lz4.h in this directory. */
#ifndef ORACLE_STUB_LZ4HC_H
#define ORACLE_STUB_LZ4HC_H
#ifdef __cplusplus
extern "C" {
#endif
int LZ4_compress_HC(const char* src, char* dst, int srcSize, int dstCapacity, int compressionLevel);
#ifdef __cplusplus
}
#endif
#endif
