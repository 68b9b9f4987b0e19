// Issue-rate microbenchmark (sm_100a): warp-instructions per cycle per SM sub-partition for the
// instruction classes the quantiser / Hadamard kernels are made of, alone and mixed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

#define ITERS 2000
#define UN 8

template <int KIND>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float seed) {
  float f[UN]; uint32_t u[UN]; uint64_t d[UN];
  float c4[UN][4];
  for (int i = 0; i < UN; ++i) { f[i] = seed + i + threadIdx.x; u[i] = __float_as_uint(f[i]) * 2654435761u; d[i] = ((uint64_t)u[i] << 32) | (uint32_t)(u[i] * 7u);
    for (int j = 0; j < 4; ++j) c4[i][j] = f[i] * j; }
  const float a = seed * 1.0001f, b = seed * 0.5f;
  const uint64_t a2 = ((uint64_t)__float_as_uint(a) << 32) | __float_as_uint(a), b2 = ((uint64_t)__float_as_uint(b) << 32) | __float_as_uint(b);
  const uint32_t hlim = 0x56005600u;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UN; ++i) {
      if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
      if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(a2), "l"(b2));
      if (KIND == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(b2));
      if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % UN]), "r"(hlim));
      if (KIND == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u[i]) : "r"(hlim));
      if (KIND == 5) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(hlim));
      if (KIND == 6) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; cvt.f32.f16 %0, l;}" : "=f"(f[i]) : "r"(u[i] + __float_as_uint(f[i])));  // cvt + add
      if (KIND == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(__uint_as_float(u[i])));
      if (KIND == 8) asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; sub.rn.f32.f16 %0, l, %0;}" : "+f"(f[i]) : "r"(u[i]));
      if (KIND == 9) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, %1; selp.u32 %0, %0, %2, p;}" : "+r"(u[i]) : "r"(hlim), "r"(u[(i + 1) % UN]));
      if (KIND == 10) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(hlim), "r"(u[(i+1)%UN]));
      if (KIND == 11) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                   : "+f"(c4[i][0]), "+f"(c4[i][1]), "+f"(c4[i][2]), "+f"(c4[i][3]) : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
      if (KIND == 12) {  // mix: FFMA2 + LOP3
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(a2), "l"(b2));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % UN]), "r"(hlim));
      }
      if (KIND == 13) {  // mix: FFMA + LOP3
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % UN]), "r"(hlim));
      }
      if (KIND == 14) {  // mix: FFMA + FFMA2
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(a2), "l"(b2));
      }
      if (KIND == 15) {  // mix: min.f16x2 + cvt(HADD2.F32) + FFMA2 + LOP3 (the quantiser)
        asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(hlim));
        asm volatile("{.reg .b16 l,h; mov.b32 {l,h}, %1; cvt.f32.f16 %0, l;}" : "=f"(f[i]) : "r"(u[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(a2), "l"(b2));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"((uint32_t)d[i]), "r"(__float_as_uint(f[i])));
      }
      if (KIND == 16) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(b));
      if (KIND == 17) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(a2));
      if (KIND == 18) {  // HMMA + FFMA mix
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c4[i][0]), "+f"(c4[i][1]), "+f"(c4[i][2]), "+f"(c4[i][3]) : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
        for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + j) % UN]) : "f"(a), "f"(b));
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UN; ++i) { s += f[i] + __uint_as_float(u[i]) + __uint_as_float((uint32_t)d[i]) + __uint_as_float((uint32_t)(d[i] >> 32)); for (int j = 0; j < 4; ++j) s += c4[i][j]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_slot, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<KIND><<<148, threads>>>(out, cyc, 1.5f);
  cudaDeviceSynchronize();
  k<KIND><<<148, threads>>>(out, cyc, 1.5f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double winst = (double)ITERS * UN * ops_per_slot * (threads / 32) / 4.0;  // per SMSP
  printf("%-34s threads %4d: %.3f warp-instr/clk/SMSP (%.0f cyc)\n", name, threads, winst / c, c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {256, 1024}) {
    run<0>("FFMA", 1, th); run<16>("FADD", 1, th); run<1>("FFMA2", 1, th); run<2>("FADD2", 1, th); run<17>("FMUL2", 1, th);
    run<3>("LOP3", 1, th); run<4>("PRMT", 1, th); run<5>("HMNMX2", 1, th); run<6>("cvt.f32.f16 (+IADD)", 2, th);
    run<7>("F2FP f16x2", 1, th); run<8>("FHADD", 1, th); run<9>("ISETP+SEL", 2, th); run<10>("IMAD", 1, th);
    run<11>("HMMA.16816", 1, th); run<12>("FFMA2+LOP3", 2, th); run<13>("FFMA+LOP3", 2, th); run<14>("FFMA+FFMA2", 2, th);
    run<15>("HMNMX2+cvt+FFMA2+LOP3", 4, th); run<18>("HMMA+8 FFMA", 9, th);
  }
  return 0;
}
