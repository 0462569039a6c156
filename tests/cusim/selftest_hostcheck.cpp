// Self-test of CUSIM_HOSTCHECK=1: device memory is reachable from kernels and cudaMemcpy, a host dereference is fatal.
#include "cuda_runtime.h"
#include <stdio.h>
#include <string.h>
__global__ void fill(int* p) { p[threadIdx.x] = static_cast<int>(threadIdx.x); }
int main(int argc, char** argv) {
  int* d = nullptr;
  if (cudaMalloc(&d, 64 * sizeof(int)) != cudaSuccess) return 2;
  ::cusim::launch(dim3(1), dim3(64), 0, [&]() { fill(d); }, reinterpret_cast<const void*>(+fill));
  int h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (h[63] != 63) return 3;
  if (argc > 1 && !strcmp(argv[1], "--touch")) {
    volatile int v = d[5];  // what a host-side `*device_ptr` does on the GPU box
    printf("host read of device memory went through: %d\n", v);
  }
  cudaFree(d);
  printf("ok\n");
  return 0;
}
