// Self-test of the emulation's dynamic shared memory launch rule (48 KB unless cudaFuncSetAttribute raised it, 227 KB at most).
#include "cuda_runtime.h"
#include <stdio.h>
static int ran = 0;
__global__ void k(int) { ran = 1; }
int main() {
  ::cusim::launch(dim3(1), dim3(32), 60000, [&]() { k(0); }, reinterpret_cast<const void*>(+k));
  int e1 = cudaGetLastError(); int r1 = ran;
  int a = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  ::cusim::launch(dim3(1), dim3(32), 60000, [&]() { k(0); }, reinterpret_cast<const void*>(+k));
  int e2 = cudaGetLastError();
  int b = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 300000);
  printf("first: err %d ran %d | attr %d second: err %d ran %d | too large attr %d\n", e1, r1, a, e2, ran, b);
  return !(e1 == 1 && r1 == 0 && a == 0 && e2 == 0 && ran == 1 && b == 1);
}
