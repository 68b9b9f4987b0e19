// Pieces shared by the 1-CTA and 2-CTA W4A8 GEMM kernels.
#pragma once
#include <cuda.h>

#include "mq_common.cuh"
#include "mq_tc.cuh"
#include "mq_w4_layout.cuh"

namespace mq {
namespace gemm {
using namespace mq::tc;

// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row x 128-byte atoms, SBO = 1024 B
// between atoms along M/N, one atom along K (LBO unused).  cute::UMMA::SmemDescriptor bit layout.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)0 << 16;                      // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

struct Params {
  const uint8_t* w4;
  const float* s_a;
  int s_a_per_row;
  const float* s_w;
  const __half* bias;
  const __half* x0;
  const __half* w0;
  __half* out;
  int64_t ldo;
  int M, N, K;
  int m_tiles, n_tiles, k_tiles;
  unsigned dbg;  // bottleneck probes (MQ_PROBES builds only; results are garbage when non-zero)
};

// Bottleneck probes (tools/gemm_probe.py): compiled OUT of the shipped library.  A probe build
// (python -m mquant_b200.build --probes) defines MQ_PROBES and reads the mask from MQ_GEMM_DEBUG.
#ifdef MQ_PROBES
#define MQ_PROBE(p, mask) (((p).dbg & (mask)) != 0u)
unsigned probe_mask();
#else
#define MQ_PROBE(p, mask) false
static inline unsigned probe_mask() { return 0u; }
#endif

// F.silu in fp32 with torch's CUDA expression (ATen ActivationSiluKernel.cu: x / (1 + exp(-x)), expf
// and an IEEE division), rounded to fp16 like the fp16 tensor F.silu returns
__device__ __forceinline__ float silu_f32(float x) { return __fdiv_rn(x, __fadd_rn(1.0f, expf(-x))); }
// 8 x (a1 * silu(a2)) on fp16 values: silu -> fp16 (rounding point), product -> fp16 (rounding point)
__device__ __forceinline__ uint4 silu_mul8(const uint4& a1, const uint4& a2) {
  uint4 r;
  const __half2* h1 = reinterpret_cast<const __half2*>(&a1);
  const __half2* h2 = reinterpret_cast<const __half2*>(&a2);
  __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 g = __half22float2(h2[i]);
    o[i] = __hmul2(h1[i], __floats2half2_rn(silu_f32(g.x), silu_f32(g.y)));
  }
  return r;
}

__device__ __forceinline__ void unpack_store(uint32_t brow, int rsw, int j, const uint4& p) {
  // low nibbles -> k = 32j .. 32j+15 (16-byte chunk 2j), high nibbles -> chunk 2j+1; values x16
  const uint32_t lo0 = (p.x << 4) & 0xF0F0F0F0u, lo1 = (p.y << 4) & 0xF0F0F0F0u;
  const uint32_t lo2 = (p.z << 4) & 0xF0F0F0F0u, lo3 = (p.w << 4) & 0xF0F0F0F0u;
  const uint32_t hi0 = p.x & 0xF0F0F0F0u, hi1 = p.y & 0xF0F0F0F0u;
  const uint32_t hi2 = p.z & 0xF0F0F0F0u, hi3 = p.w & 0xF0F0F0F0u;
  const uint32_t a0 = brow + ((uint32_t)((2 * j) ^ rsw) << 4);
  const uint32_t a1 = brow + ((uint32_t)((2 * j + 1) ^ rsw) << 4);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(lo0), "r"(lo1), "r"(lo2), "r"(lo3) : "memory");
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(hi0), "r"(hi1), "r"(hi2), "r"(hi3) : "memory");
}


// mq_gemm_w4a8_v2.cu: 128 x 256 tile per CTA, double-buffered TMEM accumulator
int launch_w4a8_gemm_v2(const CUtensorMap& tmap_a, const Params& p, cudaStream_t st);
// mq_gemm_w4a8_v4.cu: 256 x 256 tile per CTA, one converted weight k-block shared by two M halves
// (tmap_out: fp16 output, 32 x 32 boxes, SWIZZLE_64B -- the epilogue stores through TMA)
// gated: interleaved w1 | w2 rows, output [M, N / 2] = a1 * silu(a2) (tmap_out describes that tensor)
int launch_w4a8_gemm_v4(const CUtensorMap& tmap_a, const CUtensorMap& tmap_out, const Params& p, cudaStream_t st,
                        bool gated = false);
// mq_gemm_w4a8_v5.cu: 128 channels x 256 tokens per CTA, converted weights as the TMEM-resident A operand
int launch_w4a8_gemm_v5(const CUtensorMap& tmap_a, const Params& p, cudaStream_t st);
// mq_gemm_w4a8_sk.cu: split-K flavour of the 128 x 256 kernel for mid-sized M (int32 partial sums in
// a caller-provided self-cleaning workspace); splitk_factor() == 1 means "not worth it"
int splitk_factor(int64_t M, int64_t N, int64_t K);
size_t splitk_workspace_bytes(int64_t M, int64_t N, int64_t K);
int launch_w4a8_gemm_sk(const CUtensorMap& tmap_a, const Params& p, void* workspace, cudaStream_t st);

}  // namespace gemm

// mq_gemv_w4a8.cu: weight-streaming kernel for M <= 16 (split-K, int32 atomics in a workspace)
size_t gemv_workspace_bytes(int64_t N);
int launch_w4a8_gemv(const int8_t* qa, int64_t lda, const gemm::Params& g, void* workspace, cudaStream_t st);

}  // namespace mq
