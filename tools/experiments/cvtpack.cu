// cvt.pack.sat.s8.s32.b32: byte order probe + issue rate (tools/experiments; nvcc -arch=sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned* out) {
  unsigned d;
  int a = 1, b = 2;
  unsigned c = 0x0403;
  asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  out[0] = d;
  a = 300; b = -300;
  asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  out[1] = d;
}
__global__ void rate(unsigned* out, int n) {
  int a = threadIdx.x, b = threadIdx.x * 3;
  unsigned c = 0, d0 = 1, d1 = 2, d2 = 3, d3 = 4;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %0;" : "+r"(d0) : "r"(a), "r"(b));
    asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %0;" : "+r"(d1) : "r"(a), "r"(b));
    asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %0;" : "+r"(d2) : "r"(a), "r"(b));
    asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %0;" : "+r"(d3) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2] = (unsigned)(t1 - t0); }
  out[3 + threadIdx.x % 4] = d0 ^ d1 ^ d2 ^ d3 ^ c;
}
int main() {
  unsigned* d; cudaMalloc(&d, 64);
  probe<<<1, 1>>>(d);
  rate<<<1, 256>>>(d, 10000);
  unsigned h[8]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("pack(1,2,c=0x0403) = 0x%08x   pack(300,-300) = 0x%08x\n", h[0], h[1]);
  printf("256 threads (2 warps/SMSP): 40000 I2IP per thread in %u clk -> %.3f warp-instr/clk/SMSP\n", h[2], 40000.0 * 2 / h[2]);
  return 0;
}
