// V5 adaptive integer sections, encode side (placeholder until the section kernels land).
#include "cldn_device.cuh"
#include "cldn_kernels.h"
namespace cldn {
int launch_encode_sections(const Plan&, const SecLaunch&, cudaStream_t) { return -1; }
int launch_place_sections(const Plan&, const SecLaunch&, const uint64_t*, uint32_t, uint32_t, cudaStream_t) { return -1; }
}  // namespace cldn
