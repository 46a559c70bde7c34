/* Build shim for oracle/_ref only: the image ships liblz4.so.1 without its dev
 * headers.  The four LZ4 entry points the reference's file I/O names are
 * defined as failing stubs in ref_harness.cpp (compressed I/O is not on the
 * integration path and is never exercised by the oracle). */
#ifndef ORACLE_STUB_LZ4_H
#define ORACLE_STUB_LZ4_H
#ifdef __cplusplus
extern "C" {
#endif
int LZ4_compressBound(int inputSize);
int LZ4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);
int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
#ifdef __cplusplus
}
#endif
#endif
