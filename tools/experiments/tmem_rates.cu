// TMEM read-port microbenchmark (sm_100a): bytes per clock per SM that tcgen05.ld delivers, as a
// function of the number of reading warps, the load width and the loads kept in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rates tmem_rates.cu && ./tmem_rates
#include <cstdint>
#include <cstdio>

#define ITERS 2000

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: x32 loads, one in flight; 1: x32, two in flight; 2: x16, four in flight; 3: x32 x 4 in flight
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const uint32_t col = (uint32_t)((it * 128 + (warp >> 2) * 32) & 511);
    if (MODE == 0) {
      uint32_t r[32];
      ld32(base + (col & 480), r);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 8) acc ^= r[i];
    } else if (MODE == 1) {
      uint32_t r0[32], r1[32];
      ld32(base + (col & 480), r0);
      ld32(base + ((col + 32) & 480), r1);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 8) acc ^= r0[i] ^ r1[i];
    } else if (MODE == 2) {
      uint32_t r0[16], r1[16], r2[16], r3[16];
      ld16(base + (col & 496), r0);
      ld16(base + ((col + 16) & 496), r1);
      ld16(base + ((col + 32) & 496), r2);
      ld16(base + ((col + 48) & 496), r3);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 16; i += 8) acc ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
    } else {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      ld32(base + (col & 480), r0);
      ld32(base + ((col + 32) & 480), r1);
      ld32(base + ((col + 64) & 480), r2);
      ld32(base + ((col + 96) & 480), r3);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 8) acc ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int bytes_per_iter_per_warp, int threads) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, threads>>>(out, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double bytes = (double)ITERS * bytes_per_iter_per_warp * (threads / 32);
  printf("%-28s warps %2d: %7.1f B/clk/SM (%.0f cyc) %s\n", name, threads / 32, bytes / c, c, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0>("x32, 1 in flight", 4096, th);
    run<1>("x32, 2 in flight", 8192, th);
    run<2>("x16, 4 in flight", 8192, th);
    run<3>("x32, 4 in flight", 16384, th);
  }
  return 0;
}
