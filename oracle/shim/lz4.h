/* Declaration-only stand-in for <lz4.h> (dev headers are absent in this image;
 * the runtime liblz4.so.1 is present). Only the three entry points the reference
 * calls at cloudini_lib/src/codec_common.cpp:232-278 and cloudini.cpp:281 are declared. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
int LZ4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity);
int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
int LZ4_compressBound(int inputSize);
#ifdef __cplusplus
}
#endif
