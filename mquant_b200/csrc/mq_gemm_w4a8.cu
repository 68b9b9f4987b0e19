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
//   warp 3      packed-weight producer: one 16 KB cp.async.bulk per k-block into a staging ring
//               (mq_w4_layout.cuh makes each (n-block, k-block) blob contiguous)
//   warps 8-15  weight converters: staging -> registers -> two logic ops per 8 weights -> INT8
//               (x16) in the K-major SWIZZLE_128B layout -> fence.proxy.async -> mbarrier arrive.
//               (A first version prefetched the packed weights with LDG into registers; the proxy
//               fence is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS and stalled on the outstanding
//               LDGs every k-block: 1.35 -> see profiles/ for the staged version.)
// Pipelines: A/B ring of 3 stages (full: TMA tx + 256 converter arrivals; empty: tcgen05.commit),
// packed-weight ring of 4 (wfull: bulk-copy tx; wempty: 256 converter arrivals), TMEM double
// buffer (tmem_full: commit after the last k-block; tmem_empty: 128 epilogue arrivals).
//
// Roofline: tensor pipe.  ops = 2*M*N*K; one k-block (128x256x128 MACs) is 512 cycles at the
// 8192 MAC/clk/SM INT8 rate.  Bytes: N*K/2 (weights, L2 resident across m-tiles) + M*K + 2*M*N.
#include <cuda.h>

#include <cstdlib>

#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
using namespace mq::tc;

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 128;  // bytes == int8 elements
constexpr int UMMA_K = 32;
constexpr int STAGES = 3;   // A (TMA) + B (converted INT8) ring
constexpr int WSTAGES = 4;  // packed-INT4 staging ring (bulk async copies)
constexpr int A_STAGE = BM * BK;  // 16 KB
constexpr int B_STAGE = BN * BK;  // 32 KB
constexpr int W_STAGE = kW4BlobBytes;  // 16 KB
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4;
constexpr int CONV_WARP0 = 8;
constexpr int NUM_CONV = 256;
constexpr int TMEM_COLS = 512;

// dynamic shared memory carve-up (base aligned to 1024 for SWIZZLE_128B)
constexpr int SM_A = 0;
constexpr int SM_B = SM_A + STAGES * A_STAGE;
constexpr int SM_W = SM_B + STAGES * B_STAGE;
constexpr int SM_PARAM = SM_W + WSTAGES * W_STAGE;      // 2 x {s_w16[256], bias[256], w0[256]} fp32
constexpr int SM_BAR = SM_PARAM + 2 * 3 * BN * 4;       // full[S] empty[S] wfull[W] wempty[W] tfull[2] tempty[2]
constexpr int SM_TMEMPTR = SM_BAR + (2 * STAGES + 2 * WSTAGES + 4) * 8;
constexpr int SM_TOTAL = SM_TMEMPTR + 16;
constexpr int SM_ALLOC = SM_TOTAL + 1024;  // slack for manual 1024-byte alignment

// cute::UMMA::InstrDescriptor for kind::i8, S32 accumulate, both operands K-major
constexpr uint32_t kInstrDesc = (2u << 4)      // c_format = S32
                                | (1u << 7)    // a_format = INT8
                                | (1u << 10)   // b_format = INT8
                                | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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
  auto wfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto wempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + WSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 * WSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 * WSTAGES + 2 + a); };
  const uint32_t sW = base + SM_W;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler: role code stays on the uniform datapath
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + NUM_CONV / 32);  // TMA thread + one arrive per converter warp
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), NUM_CONV / 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
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
  } else if (warp == 3) {
    // ===================== packed-weight producer =====================
    // one 16 KB bulk async copy per k-block: the tile-major layout makes every blob contiguous
    if (lane == 0 && !(p.dbg & 1u)) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nb = t % p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(wempty_bar(ws), wphase ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(ws), W_STAGE);
          bulk_load(sW + ws * W_STAGE, p.w4 + w4_blob_offset(nb, kb, p.k_tiles), W_STAGE, wfull_bar(ws));
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
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
    // staging (packed INT4, landed by the bulk copy) -> registers -> INT8 x16 -> B stage.
    // Nothing but shared-memory traffic is outstanding when the proxy fence (MEMBAR.ALL.CTA +
    // FENCE.VIEW.ASYNC) executes, so it costs tens of cycles, not an L2 round trip.
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_it = my_tiles * p.k_tiles;
    const bool do_w = !(p.dbg & 1u);
    const bool do_sts = !(p.dbg & (1u | 16u));
    int stage = 0, ws = 0;
    uint32_t phase = 0, wphase = 0;
    for (int it = 0; it < total_it; ++it) {
      uint4 pk[4];
      if (do_w) {
        mbar_wait(wfull_bar(ws), wphase);
        const uint32_t src = sW + ws * W_STAGE + (uint32_t)r * 16u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(pk[j].x), "=r"(pk[j].y), "=r"(pk[j].z), "=r"(pk[j].w)
                       : "r"(src + (uint32_t)j * kW4TileN * 16u));
        }
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t brow = sB + stage * B_STAGE + brow_off;
      if (do_sts) {
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack_store(brow, rsw, j, pk[j]);
      }
      fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();        // ... and ordered before lane 0's release-arrive (one arrival per warp)
      if (lane == 0) {
        mbar_arrive(full_bar(stage));
        if (do_w) mbar_arrive(wempty_bar(ws));  // the loads above have been consumed by the stores
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
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
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace gemm
}  // namespace mq

extern "C" size_t mq_w4a8_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  // one call uses at most one of the two (both leave the workspace zeroed), so they may alias
  const size_t a = M <= 16 ? mq::gemv_workspace_bytes(N) : 0;
  const size_t b = K % mq::gemm::BK == 0 ? mq::gemm::splitk_workspace_bytes(M, N, K) : 0;
  return a > b ? a : b;
}

extern "C" int mq_w4a8_gemm(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                            const float* s_w, const void* bias, const void* x0, const void* w0, void* out,
                            int64_t ldo, int64_t M, int64_t N, int64_t K, void* stream) {
  return mq_w4a8_gemm_ws(qa, lda, w4, s_a, s_a_per_row, s_w, bias, x0, w0, out, ldo, M, N, K, nullptr, 0, stream);
}

extern "C" int mq_w4a8_gemm_ws(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                               const float* s_w, const void* bias, const void* x0, const void* w0, void* out,
                               int64_t ldo, int64_t M, int64_t N, int64_t K, void* workspace,
                               size_t workspace_bytes, void* stream) {
  using namespace mq;
  using namespace mq::gemm;
  if (M == 0) return MQ_OK;  // empty batch: nothing to do (its tensors have null data pointers)
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

  tc::EncodeTiledFn encode = tc::get_encode_fn();
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
  const char* kenv = getenv("MQ_GEMM_KERNEL");
  // decode-sized batches: weight streaming with split-K when the caller provides the workspace
  // (any MQ_GEMM_KERNEL override selects a tensor-core kernel instead).  Above 8 rows the dp4a rate
  // starts to bind and the tcgen05 kernel wins once N gives it >= 32 CTAs (profiles/README.md).
  const bool gemv_shape = M <= 8 || (M <= 16 && p.n_tiles < 32);
  if (gemv_shape && workspace != nullptr && workspace_bytes >= gemv_workspace_bytes(N) && p.dbg == 0 && kenv == nullptr) {
    const int rc = launch_w4a8_gemv(qa, lda, p, workspace, (cudaStream_t)stream);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemv_kernel");
    return MQ_OK;
  }
  // mid-sized M: too few tiles for the persistent kernels -> split K across CTAs
  if (kenv == nullptr && p.dbg == 0 && workspace != nullptr && splitk_factor(M, N, K) >= 2 &&
      workspace_bytes >= splitk_workspace_bytes(M, N, K)) {
    const int rc = launch_w4a8_gemm_sk(tmap, p, workspace, (cudaStream_t)stream);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_sk_kernel");
    return MQ_OK;
  }
  // Kernel selection (MQ_GEMM_KERNEL = v1 | v2 | v3 | v4 | 2cta overrides): v4 = 256 x 256 tile, one
  // weight conversion shared by two M halves (default for large M); v2 = 128 x 256 tile, 8-warp
  // pipelined epilogue with staged stores (default below); 2cta = its cta_group::2 flavour (correct, but its TMA path currently
  // under-performs, profiles/README.md); v1 = this file's first-generation kernel.
  // default: the cheaper of the 256 x 256 kernel (v4) and the 128 x 256 kernel (v2) under a
  // wave model fitted to tools/latency.py (us per wave of one tile per SM: v2 0.506 K + 6.5,
  // v4 0.60 K + 7.5, K in k-blocks; profiles/README.md): v2 only wins while it needs a single wave
  const int64_t sms = num_sms();
  const int64_t waves2 = (((M + 127) / 128) * (int64_t)p.n_tiles + sms - 1) / sms;
  const int64_t waves4 = (((M + 255) / 256) * (int64_t)p.n_tiles + sms - 1) / sms;
  const double cost2 = (double)waves2 * (0.506 * p.k_tiles + 6.5);
  const double cost4 = (double)waves4 * (0.60 * p.k_tiles + 7.5);
  int kernel = cost4 < cost2 ? 5 : 2;
  if (kenv) {
    if (kenv[0] == '2') kernel = 3;                         // "2cta"
    else if (kenv[0] == 'v' && kenv[1] == '1') kernel = 1;  // "v1"
    else if (kenv[0] == 'v' && kenv[1] == '3') kernel = 4;  // "v3"
    else if (kenv[0] == 'v' && kenv[1] == '4') kernel = 5;  // "v4"
    else kernel = 2;                                        // "v2"
  }
  if (kernel == 5) {
    // the 256 x 256 kernel writes its output with TMA: 32 x 32 fp16 boxes out of swizzled staging
    CUtensorMap tmap_out;
    const cuuint64_t odim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t ostride[1] = {(cuuint64_t)ldo * 2};
    const cuuint32_t obox[2] = {32, 32};
    const CUresult cro = encode(&tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, odim, ostride, obox, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cro != CUDA_SUCCESS) {
      set_error("mq_w4a8_gemm: cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)cro);
      return MQ_ERR_CUDA;
    }
    const int rc = launch_w4a8_gemm_v4(tmap, tmap_out, p, (cudaStream_t)stream);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_v4_kernel");
    return MQ_OK;
  }
  if (kernel == 4) {
    const int rc = launch_w4a8_gemm_v3(tmap, p, (cudaStream_t)stream);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_v3_kernel");
    return MQ_OK;
  }
  if (kernel >= 2) {
    const int rc = launch_w4a8_gemm_v2(tmap, p, (cudaStream_t)stream, kernel == 3 && M > BM);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_v2_kernel");
    return MQ_OK;
  }
  MQ_CUDA(cudaFuncSetAttribute(w4a8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC));
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  w4a8_gemm_kernel<<<grid, NUM_THREADS, SM_ALLOC, (cudaStream_t)stream>>>(tmap, p);
  MQ_LAUNCH_CHECK("w4a8_gemm_kernel");
  return MQ_OK;
}
