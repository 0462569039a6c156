// Self-test of CUSIM_HOSTCHECK=1 (device memory is reachable from kernels and cudaMemcpy, a host dereference is fatal)
// and of CUSIM_ASYNC=1 (--async: a cudaMemcpyAsync result is not there before the host synchronises with the stream).
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
  if (argc > 1 && !strcmp(argv[1], "--async")) {
    cudaStream_t st;
    cudaEvent_t ev;
    cudaStreamCreate(&st);
    cudaEventCreate(&ev);
    int a[64], b[64];
    cudaMemcpyAsync(a, d, sizeof(a), cudaMemcpyDeviceToHost, st);
    cudaEventRecord(ev, st);
    cudaMemcpyAsync(b, d, sizeof(b), cudaMemcpyDeviceToHost, st);
    const bool early = a[63] == 63 || b[63] == 63;          // nothing may have arrived yet
    cudaEventSynchronize(ev);
    const bool first_only = a[63] == 63 && b[63] != 63;     // the event covers the first copy only
    cudaStreamSynchronize(st);
    const bool both = a[63] == 63 && b[63] == 63;
    printf("early %d first_only %d both %d\n", early, first_only, both);
    return (!early && first_only && both) ? 0 : 4;
  }
  if (argc > 1 && !strcmp(argv[1], "--touch")) {
    volatile int v = d[5];  // what a host-side `*device_ptr` does on the GPU box
    printf("host read of device memory went through: %d\n", v);
  }
  cudaFree(d);
  printf("ok\n");
  return 0;
}
