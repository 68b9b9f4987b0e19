// tcgen05.mma issue/throughput microbenchmark (sm_100a): cycles per kind::f16 MMA (M = 128, K = 16)
// as a function of N, of the B operand's major-ness and of whether consecutive MMAs reuse A.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../mquant_b200/csrc -o mma_rates mma_rates.cu
#include <cstdint>
#include <cstdio>
#include <cuda.h>

#include "mq_tc.cuh"
using namespace mq::tc;

#define NMMA 256
// (issued from warp-uniform code through elect_one(): otherwise every UTCHMMA sits in an ELECT/R2UR
// loop and the issuing thread, ~70-120 clk per MMA, is what gets measured)

// mode: 0 = A walks 4 k-steps of one 128-row tile (like a K loop), 1 = A fixed
// bg: what warps 4..11 do meanwhile: 0 nothing, 1 FFMA chains, 2 tcgen05.ld loop, 3 256-bit global stores, 4 LDS loop
__global__ void __launch_bounds__(384) k(long long* cyc, int N, int b_mn_major, int mode, int a_rows_stride, int bg, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  for (int i = threadIdx.x; i < 48 * 1024; i += 384) smem[i] = 0;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sA = base, sB = base + 48 * 1024;
    const long long t0 = clock64();
    if (mode == 2) {
      // lean issue: descriptors = base + compile-time constants, 16 MMAs per loop trip (the K loop of
      // the Hadamard kernel: A walks 4 k-steps x 3 k-blocks, B 16 rows per step)
      const uint64_t ad0 = make_sw128_desc(sA, 0, 1024);
      const uint64_t bd0 = b_mn_major ? make_sw128_desc(sB, 16384, 1024) : make_sw128_desc(sB, 0, 1024);
      for (int i = 0; i < NMMA; i += 16) {
        const uint32_t tmd = tm + (uint32_t)((i >> 4) & 1) * 256u;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          umma_f16(tmd, ad0 + (uint64_t)((((j >> 2) % 3) * 16384 + (j & 3) * 32) >> 4), bd0 + (uint64_t)((j & 3) * (b_mn_major ? 128 : 2)), idesc, j == 0 ? 0u : 1u);
      }
    } else
    for (int i = 0; i < NMMA; ++i) {
      const uint32_t ks = (uint32_t)(i & 3);
      const uint32_t aoff = mode == 0 ? ks * 32u + (uint32_t)((i >> 2) % 3) * (uint32_t)a_rows_stride : 0u;
      const uint64_t ad = make_sw128_desc(sA + aoff, 0, 1024);
      const uint64_t bd = b_mn_major ? make_sw128_desc(sB + ks * 2048u, 16384, 1024) : make_sw128_desc(sB + ks * 32u, 0, 1024);
      umma_f16(tm + (uint32_t)((i >> 4) & 1) * 256u, ad, bd, idesc, (uint32_t)((i & 15) != 0));
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    cyc[2 * blockIdx.x] = t1 - t0;
    cyc[2 * blockIdx.x + 1] = t2 - t0;
    stop = 1;
  } else if (warp >= 4 && bg != 0) {
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    uint32_t acc = 0;
    uint32_t* dst = sink + ((size_t)blockIdx.x * 384 + threadIdx.x) * 8;
    while (!stop) {
      if (bg == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { f0 = f0 * 1.0001f + 0.5f; f1 = f1 * 1.0001f + 0.5f; f2 = f2 * 1.0001f + 0.5f; f3 = f3 * 1.0001f + 0.5f; }
      } else if (bg == 2) {
        uint32_t r[32];
        tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + 128u * (uint32_t)((warp >> 2) & 1), r);
        acc ^= r[0] ^ r[31];
      } else if (bg == 3) {
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst), "r"(acc) : "memory");
        dst += 384 * 8 * 148;
        if (dst > sink + (size_t)60 * 1024 * 1024) dst = sink + ((size_t)blockIdx.x * 384 + threadIdx.x) * 8;
        acc++;
      } else if (bg == 4) {
        uint32_t vx, vy, vz, vw;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vx), "=r"(vy), "=r"(vz), "=r"(vw) : "r"(base + 100 * 1024 + (threadIdx.x * 16) % 8192));
        acc ^= vx ^ vw;
      }
    }
    if (acc == 0x12345678u || f0 + f1 + f2 + f3 == 1.2345f) sink[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  uint32_t* sink;
  cudaMalloc(&sink, (size_t)256 * 1024 * 1024);
  for (int bg = 0; bg < 5; ++bg)
  for (int mode = 0; mode < 3; ++mode)
    for (int mn = 0; mn < 2; ++mn)
      for (int N : {64, 128, 256}) {
        if (bg > 0 && (mode == 1 || mn != 1 || N != 128)) continue;
        for (int rep = 0; rep < 2; ++rep) k<<<148, 384, 200 * 1024>>>(cyc, N, mn, mode, 16384, bg, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[296];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double a = 0, b = 0;
        for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
        printf("bg %d M128 N%-3d K16 B %s-major, A %s: issue %.1f clk/MMA, complete %.1f clk/MMA (math floor %d) %s\n", bg, N,
               mn ? "MN" : "K ", mode == 2 ? "walking, lean issue" : mode ? "fixed  " : "walking", a / 148 / NMMA, b / 148 / NMMA, N / 2, cudaGetErrorString(e));
      }
  return 0;
}
