// W4A8 GEMM for sm_100a: INT8 activations x INT4 weights on tcgen05 (kind::i8), accumulators in
// TMEM, activations by TMA, weights unpacked INT4 -> INT8 into the UMMA shared-memory layout by
// converter warps, scale / rank-1 / bias epilogue straight out of TMEM.
//
//   out[m,n] = fp16( fp32(sum_k qa[m,k] * w[n,k]) * (s_a[m]*s_w[n]) (+ x0[m]*w0[n]) (+ bias[n]) )
//
// Replaces ActQuantWrapper's `self.module(x)` (fake_quant/quant_utils.py:384) and the fp32
// L1/L2 split (:374-376).  The integer sum is exact; the epilogue is specified operation by
// operation (oracle/mq_oracle.py:qlinear_int) and reproduced bit for bit.
//
// Persistent kernel, one CTA per SM, 512 threads:
//   warp 0      TMA producer: A tile [128 x 128 B] per k-block, SWIZZLE_128B, zero-filled past M
//   warp 1      MMA issuer: 4 x tcgen05.mma (M128 N256 K32) per k-block, commits free the stage
//   warp 2      TMEM allocator (512 columns = 2 accumulator buffers of 256)
//   warps 4-7   epilogue: tcgen05.ld 32x32b.x32, scale/bias, 64-byte row segments to HBM
//   warps 8-15  weight converters: packed INT4 (mq_w4_layout.cuh) global -> registers (software
//               prefetched two k-blocks ahead) -> two logic ops per 8 weights -> INT8 (x16) in the
//               K-major SWIZZLE_128B layout -> fence.proxy.async -> mbarrier arrive
// Pipelines: smem ring of 4 stages (full: TMA tx + 256 converter arrivals; empty: tcgen05.commit),
// TMEM double buffer (tmem_full: commit after the last k-block; tmem_empty: 128 epilogue arrivals).
//
// Roofline: tensor pipe.  ops = 2*M*N*K; one k-block (128x256x128 MACs) is 512 cycles at the
// 8192 MAC/clk/SM INT8 rate.  Bytes: N*K/2 (weights, L2 resident across m-tiles) + M*K + 2*M*N.
#include <cuda.h>

#include <cstdlib>

#include "mq_common.cuh"
#include "mq_w4_layout.cuh"

namespace mq {
namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 128;  // bytes == int8 elements
constexpr int UMMA_K = 32;
constexpr int STAGES = 4;
constexpr int A_STAGE = BM * BK;  // 16 KB
constexpr int B_STAGE = BN * BK;  // 32 KB
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4;
constexpr int CONV_WARP0 = 8;
constexpr int NUM_CONV = 256;
constexpr int TMEM_COLS = 512;

// dynamic shared memory carve-up (base aligned to 1024 for SWIZZLE_128B)
constexpr int SM_A = 0;
constexpr int SM_B = SM_A + STAGES * A_STAGE;
constexpr int SM_PARAM = SM_B + STAGES * B_STAGE;       // 2 x {s_w16[256], bias[256], w0[256]} fp32
constexpr int SM_BAR = SM_PARAM + 2 * 3 * BN * 4;       // full[4] empty[4] tfull[2] tempty[2]
constexpr int SM_TMEMPTR = SM_BAR + (2 * STAGES + 4) * 8;
constexpr int SM_TOTAL = SM_TMEMPTR + 16;
constexpr int SM_ALLOC = SM_TOTAL + 1024;  // slack for manual 1024-byte alignment

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

// cute::UMMA::InstrDescriptor for kind::i8, S32 accumulate, both operands K-major
constexpr uint32_t kInstrDesc = (2u << 4)      // c_format = S32
                                | (1u << 7)    // a_format = INT8
                                | (1u << 10)   // b_format = INT8
                                | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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
  unsigned dbg;  // MQ_GEMM_DEBUG bottleneck probes (results are garbage when non-zero)
};

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

__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + SM_A, sB = base + SM_B;
  const uint32_t bar0 = base + SM_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + NUM_CONV);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + SM_TMEMPTR),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (A) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mb = t / p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (p.dbg & 2u) {
            mbar_arrive(full_bar(stage));
          } else {
            mbar_arrive_expect_tx(full_bar(stage), A_STAGE);
            tma_load_2d(sA + stage * A_STAGE, &tmap_a, kb * BK, mb * BM, full_bar(stage));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < p.k_tiles; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_kmajor_sw128_desc(sA + stage * A_STAGE);
          const uint64_t bdesc = make_kmajor_sw128_desc(sB + stage * B_STAGE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 bytes along K inside the 128-byte swizzle atom = +2 in the 16-byte address field
            if (!(p.dbg & 8u))
              umma_i8(tmem_d, adesc + 2u * k, bdesc + 2u * k, kInstrDesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(stage));
          if (kb == p.k_tiles - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== weight converters =====================
    const int r = threadIdx.x - CONV_WARP0 * 32;  // row of the 256-row B tile
    const int rsw = r & 7;
    const uint32_t brow_off = (uint32_t)r * 128u;
    // Packed weights come straight from L2 into registers, software-prefetched PF k-blocks ahead
    // of the conversion cursor (PF == STAGES == 4: ~2000 cycles of cover for an L2 round trip,
    // and the stage index of the unrolled slot is a compile-time constant).
    constexpr int PF = STAGES;
    int pf_t = blockIdx.x, pf_kb = 0;
    uint4 buf[PF][4];
    auto load4 = [&](uint4 (&b)[4]) {
      if (pf_t < total_tiles && !(p.dbg & 1u)) {
        const int nb = pf_t % p.n_tiles;
        const uint8_t* src = p.w4 + w4_blob_offset(nb, pf_kb, p.k_tiles) + (size_t)r * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = ldg_stream(src + (size_t)j * kW4TileN * 16);
      }
      if (++pf_kb == p.k_tiles) { pf_kb = 0; pf_t += gridDim.x; }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u) load4(buf[u]);
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_it = my_tiles * p.k_tiles;
    const bool do_sts = !(p.dbg & (1u | 16u));
    uint32_t phase = 0;
    for (int it0 = 0; it0 < total_it; it0 += PF, phase ^= 1u) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        if (it0 + u < total_it) {
          mbar_wait(empty_bar(u), phase ^ 1u);
          const uint32_t brow = sB + u * B_STAGE + brow_off;
          if (do_sts) {
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack_store(brow, rsw, j, buf[u][j]);
          }
          load4(buf[u]);
          fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
          mbar_arrive(full_bar(u));
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    float* params = reinterpret_cast<float*>(smem + SM_PARAM);
    int local = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++local) {
      const int mb = t / p.n_tiles, nb = t % p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
      float* ps = params + acc * 3 * BN;
      for (int c = et; c < BN; c += 128) {
        const int n = nb * BN + c;
        const bool ok = n < p.N;
        ps[c] = ok ? __fmul_rn(p.s_w[n], 0.0625f) : 0.f;  // the INT8 operand holds 16 x code
        ps[BN + c] = (ok && p.bias != nullptr) ? __half2float(p.bias[n]) : 0.f;
        ps[2 * BN + c] = (ok && p.w0 != nullptr) ? __half2float(p.w0[n]) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int m = mb * BM + quad * 32 + lane;
      const bool row_ok = m < p.M;
      float sa = 0.f, x0v = 0.f;
      if (row_ok) {
        sa = p.s_a_per_row ? p.s_a[m] : p.s_a[0];
        if (p.x0 != nullptr) x0v = __half2float(p.x0[m]);
      }
      const bool has_r1 = p.x0 != nullptr, has_bias = p.bias != nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      __half* orow = p.out + (int64_t)m * p.ldo + (int64_t)nb * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            if (nb * BN + col < p.N) {
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float sc = __fmul_rn(sa, ps[col + e]);
                float val = __fmul_rn((float)(int)v[g * 8 + e], sc);
                if (has_r1) val = __fadd_rn(val, __fmul_rn(x0v, ps[2 * BN + col + e]));
                if (has_bias) val = __fadd_rn(val, ps[BN + col + e]);
                y[e] = val;
              }
              if (!(p.dbg & 4u)) stg_stream(orow + col, pack8_f16(y));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace gemm
}  // namespace mq

extern "C" int mq_w4a8_gemm(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                            const float* s_w, const void* bias, const void* x0, const void* w0, void* out,
                            int64_t ldo, int64_t M, int64_t N, int64_t K, void* stream) {
  using namespace mq;
  using namespace mq::gemm;
  MQ_REQUIRE(qa && w4 && s_a && s_w && out, "mq_w4a8_gemm: null pointer");
  MQ_REQUIRE(M >= 0 && N > 0 && K > 0, "mq_w4a8_gemm: bad shape (M=%lld N=%lld K=%lld)", (long long)M,
             (long long)N, (long long)K);
  MQ_REQUIRE(K % BK == 0, "mq_w4a8_gemm: K=%lld must be a multiple of %d (zero-pad the codes)", (long long)K, BK);
  MQ_REQUIRE(N % 8 == 0 && ldo % 8 == 0 && ldo >= N, "mq_w4a8_gemm: N and ldo must be multiples of 8 (N=%lld ldo=%lld)",
             (long long)N, (long long)ldo);
  MQ_REQUIRE(lda % 16 == 0 && lda >= K, "mq_w4a8_gemm: lda=%lld must be a multiple of 16 and >= K", (long long)lda);
  MQ_REQUIRE(aligned16(qa) && aligned16(w4) && aligned16(out), "mq_w4a8_gemm: pointers must be 16-byte aligned");
  MQ_REQUIRE((x0 == nullptr) == (w0 == nullptr), "mq_w4a8_gemm: x0 and w0 go together");
  MQ_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "mq_w4a8_gemm: dimension exceeds 2^31");
  if (M == 0) return MQ_OK;

  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_w4a8_gemm: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)lda};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)qa, gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("mq_w4a8_gemm: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  Params p;
  p.w4 = w4; p.s_a = s_a; p.s_a_per_row = s_a_per_row; p.s_w = s_w;
  p.bias = (const __half*)bias; p.x0 = (const __half*)x0; p.w0 = (const __half*)w0;
  p.out = (__half*)out; p.ldo = ldo;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.m_tiles = (int)((M + BM - 1) / BM);
  p.n_tiles = (int)((N + BN - 1) / BN);
  p.k_tiles = (int)(K / BK);
  const char* dbg_env = getenv("MQ_GEMM_DEBUG");
  p.dbg = dbg_env ? (unsigned)atoi(dbg_env) : 0u;
  MQ_CUDA(cudaFuncSetAttribute(w4a8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC));
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  w4a8_gemm_kernel<<<grid, NUM_THREADS, SM_ALLOC, (cudaStream_t)stream>>>(tmap, p);
  MQ_LAUNCH_CHECK("w4a8_gemm_kernel");
  return MQ_OK;
}
