// Dependent-chain latency of the instructions the Hadamard kernels chain (sm_100a), one warp per SM.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#define ITERS 4000
template <int KIND>
__global__ void k(float* out, long long* cyc, float seed) {
  float c[4] = {seed, seed * 2, seed * 3, seed * 4};
  float f = seed; uint32_t u = __float_as_uint(seed) + threadIdx.x; uint64_t d = ((uint64_t)u << 32) | u;
  uint32_t a0 = u, a1 = u * 3, a2 = u * 5, a3 = u * 7, b0 = u * 11, b1 = u * 13;
  long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < ITERS; ++it) {
    if (KIND == 0) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    if (KIND == 1) {  // HMMA whose A operand depends on the previous result (C -> F2FP -> A)
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a0) : "f"(c[1]), "f"(c[0]));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                   : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
    }
    if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(seed), "f"(seed));
    if (KIND == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d) : "l"(d ^ 1));
    if (KIND == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(__uint_as_float(u)), "f"(f));
    if (KIND == 5) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; sub.rn.f32.f16 %0, l, %0;}" : "+f"(f) : "r"(u));
    if (KIND == 6) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u) : "r"(a1));
    if (KIND == 7) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u) : "r"(a1));
    if (KIND == 8) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; cvt.f32.f16 %0, l;}" : "=f"(f) : "r"(__float_as_uint(f)));
    if (KIND == 9) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"((u & 0xfc) ) : "memory");
  }
  long long t1 = clock64();
  out[blockIdx.x * 32 + threadIdx.x] = c[0] + c[1] + c[2] + c[3] + f + __uint_as_float(u) + __uint_as_float((uint32_t)d);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int KIND> void run(const char* name, int per) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
  k<KIND><<<148, 32>>>(out, cyc, 1.25f); cudaDeviceSynchronize();
  k<KIND><<<148, 32>>>(out, cyc, 1.25f); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-44s %.1f cycles per step\n", name, c / ITERS / per);
}
int main() {
  run<0>("HMMA.16816 accumulate chain", 1); run<1>("HMMA -> F2FP -> HMMA (A operand) chain", 1); run<2>("FFMA chain", 1);
  run<3>("FFMA2 chain", 1); run<4>("F2FP chain", 1); run<5>("FHADD chain", 1); run<6>("PRMT chain", 1); run<7>("HMNMX2 chain", 1);
  run<8>("HADD2.F32 (cvt) chain", 1); run<9>("LDS chain", 1);
  return 0;
}
