// SUPERSEDED (round 1, first-generation 128 x 256 kernel; not compiled into libmquant_b200.so)
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
