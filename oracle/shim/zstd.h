/* Declaration-only stand-in for <zstd.h> (dev headers are absent in this image;
 * the runtime libzstd.so.1 is present). Only the entry points the reference calls at
 * cloudini_lib/src/codec_common.cpp:242-292 and cloudini.cpp:284 are declared. */
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char* ZSTD_getErrorName(size_t code);
#ifdef __cplusplus
}
#endif
