// W4A8 GEMM, second-generation kernel, in two flavours of one template:
//
//   w4a8_gemm_v2_kernel<false>  one CTA per 128 x 256 tile (cta_group::1)
//   w4a8_gemm_v2_kernel<true>   a CTA pair (cluster of 2) per 256 x 256 tile (cta_group::2): each
//                               CTA TMA-loads its own 128 activation rows and converts only ITS
//                               half (128 channels) of the weight tile; one thread of the leader
//                               issues tcgen05.mma.cta_group::2 (M = 256), which reads A and B from
//                               both CTAs' shared memory and writes each CTA's 128 x 256
//                               accumulator into that CTA's TMEM
//
// Differences to the first kernel (mq_gemm_w4a8.cu), all driven by its ncu profile
// (profiles/r1_gemm_full.csv: LSU shared-memory wavefronts 69 % of peak; probes: epilogue-only
// skeleton = 60 % of the full run time, 16-byte-per-row output stores = 13 %):
//   * 8 epilogue warps (4 lane quadrants x 2 column halves) instead of 4, TMEM loads software
//     pipelined against the arithmetic, output staged per warp in swizzled shared memory and
//     written as whole 64-byte row segments
//   * 4 converter warps; one mbarrier arrival per warp (not per thread)
//   * 2-CTA: `full[s]` completes per CTA (own TMA bytes + own converter warps); the peer's idle
//     MMA warp relays it to the leader with ONE remote arrive per k-block; `empty[s]` and
//     `tmem_full[a]` are signalled in both CTAs by a multicast tcgen05.commit; `tmem_empty[a]`
//     lives in the leader.  Waits are plain CTA-scope try_waits (an acquire.cluster wait costs a
//     CCTL.IVALL per k-block).
//
// Same contract and bit-exact results as mq_gemm_w4a8.cu (identical epilogue arithmetic).
#include <cstdio>
#include <cstdlib>

#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
namespace v2 {

constexpr int BM = 128;  // rows per CTA
constexpr int BN = 256;  // output channels per tile
constexpr int BK = 128;
constexpr int UMMA_K = 32;
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4;  // 8 epilogue warps: lane quadrant = warp & 3, column half = (warp - 4) >> 2
constexpr int NUM_EPI_WARPS = 8;
constexpr int CONV_WARP0 = 12;  // 4 converter warps
constexpr int NUM_CONV = 128;
constexpr int TMEM_COLS = 512;
constexpr int A_STAGE = BM * BK;  // 16 KB

template <bool TWO>
struct Cfg {
  static constexpr int CTAS = TWO ? 2 : 1;
  static constexpr int BNH = BN / CTAS;             // weight rows converted per CTA
  static constexpr int ROWS = BNH / NUM_CONV;       // rows per converter thread
  static constexpr int STAGES = TWO ? 5 : 3;
  static constexpr int WSTAGES = TWO ? 4 : 3;
  static constexpr int B_STAGE = BNH * BK;
  static constexpr int W_STAGE = BNH * BK / 2;
  static constexpr int SM_A = 0;
  static constexpr int SM_B = SM_A + STAGES * A_STAGE;
  static constexpr int SM_W = SM_B + STAGES * B_STAGE;
  static constexpr int SM_EPI = SM_W + WSTAGES * W_STAGE;      // 8 warps x 32 rows x 64 B
  static constexpr int SM_PARAM = SM_EPI + NUM_EPI_WARPS * 2048;  // 2 x {s_w16, bias, w0}[256] fp32
  static constexpr int SM_BAR = SM_PARAM + 2 * 3 * BN * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 2 * WSTAGES + 4;
  static constexpr int SM_TMEMPTR = SM_BAR + NUM_BARS * 8;
  static constexpr int SM_TOTAL = SM_TMEMPTR + 16;
  static constexpr int SM_ALLOC = SM_TOTAL + 1024;
  static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");
  static constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)((CTAS * BM) >> 4) << 24);
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <bool TWO>
__device__ __forceinline__ void umma_i8_v2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  if (TWO) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_i8(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
template <bool TWO>
__device__ __forceinline__ void umma_commit_v2(uint32_t bar) {
  if (TWO) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
  } else {
    umma_commit(bar);
  }
}
template <bool TWO>
__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_v2_kernel(const __grid_constant__ CUtensorMap tmap_a, const Params p) {
  using C = Cfg<TWO>;
  constexpr int STAGES = C::STAGES, WSTAGES = C::WSTAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + C::SM_A, sB = base + C::SM_B, sW = base + C::SM_W;
  const uint32_t bar0 = base + C::SM_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto wfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto wempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + WSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 * WSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 * WSTAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::SM_TMEMPTR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler: role code stays on the uniform datapath
  const int lane = threadIdx.x & 31;
  const uint32_t cta = TWO ? cluster_ctarank() : 0u;  // 0 = leader
  const int unit = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // CTA or CTA pair
  const int num_units = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = p.m_tiles * p.n_tiles;  // m_tiles counts (CTAS*128)-row tiles

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      // own TMA thread + own converter warps (+ the peer's relay, in the leader of a pair)
      mbar_init(full_bar(s), 1 + NUM_CONV / 32 + ((TWO && cta == 0) ? 1 : 0));
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), NUM_CONV / 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), C::CTAS * NUM_EPI_WARPS);  // one arrive per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + C::SM_TMEMPTR),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(base + C::SM_TMEMPTR, TMEM_COLS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (A): this CTA's 128 rows, on its own barrier
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        const int mb = t / p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (MQ_PROBE(p, 2u)) {
            mbar_arrive(full_bar(stage));
          } else {
            mbar_arrive_expect_tx(full_bar(stage), A_STAGE);
            tma_load_2d(sA + stage * A_STAGE, &tmap_a, kb * BK, mb * (C::CTAS * BM) + (int)cta * BM,
                        full_bar(stage));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== packed-weight producer: bulk copies of this CTA's rows of the blob
    if (lane == 0 && !MQ_PROBE(p, 1u)) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        const int nb = t % p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(wempty_bar(ws), wphase ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(ws), C::W_STAGE);
          const uint8_t* blob = p.w4 + w4_blob_offset(nb, kb, p.k_tiles);
          if (TWO) {
            // rows [128*cta, 128*cta + 128) of the 256-row blob = 4 runs (one per k group) of 2 KB
#pragma unroll
            for (int j = 0; j < 4; ++j)
              bulk_load(sW + ws * C::W_STAGE + j * (C::BNH * 16),
                        blob + (size_t)cta * (C::BNH * 16) + (size_t)j * (kW4TileN * 16), C::BNH * 16,
                        wfull_bar(ws));
          } else {
            bulk_load(sW + ws * C::W_STAGE, blob, C::W_STAGE, wfull_bar(ws));
          }
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (TWO && cta != 0) {
      // ===================== readiness relay (peer CTA): one remote arrive per k-block
      const uint32_t full_leader = mapa(bar0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (lane == 0) mbar_arrive_cluster(full_leader + 8u * stage);
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
      // ===================== MMA issuer =====================
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = unit; t < total_tiles; t += num_units, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_kmajor_sw128_desc(sA + stage * A_STAGE);
            const uint64_t bdesc = make_kmajor_sw128_desc(sB + stage * C::B_STAGE);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (!MQ_PROBE(p, 8u))
                umma_i8_v2<TWO>(tmem_d, adesc + 2u * k, bdesc + 2u * k, C::IDESC, (uint32_t)((kb | k) != 0));
            }
            umma_commit_v2<TWO>(empty_bar(stage));
            if (kb == p.k_tiles - 1) umma_commit_v2<TWO>(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== weight converters: 128 threads, ROWS full 128-byte rows each
    const int r0 = threadIdx.x - CONV_WARP0 * 32;
    const int my_tiles = (total_tiles - unit + num_units - 1) / num_units;
    const int total_it = my_tiles * p.k_tiles;
    const bool do_w = !MQ_PROBE(p, 1u);
    const bool do_sts = !MQ_PROBE(p, 1u | 16u);
    int stage = 0, ws = 0;
    uint32_t phase = 0, wphase = 0;
    // Software pipelined: the staging reads of k-block it+1 are issued BEFORE the proxy fence of
    // k-block it (the fence is a MEMBAR: it would otherwise serialise LDS latency + STS drain of
    // consecutive k-blocks, and the converters' per-k-block latency exceeded the 512-cycle MMA time).
    uint4 pk[2][C::ROWS][4];
    auto load_pk = [&](uint4 (&dst)[C::ROWS][4], int slot) {
#pragma unroll
      for (int rr = 0; rr < C::ROWS; ++rr) {
        const uint32_t src = sW + slot * C::W_STAGE + (uint32_t)(r0 + rr * NUM_CONV) * 16u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(dst[rr][j].x), "=r"(dst[rr][j].y), "=r"(dst[rr][j].z), "=r"(dst[rr][j].w)
                       : "r"(src + (uint32_t)j * (C::BNH * 16)));
        }
      }
    };
    auto store_pk = [&](const uint4 (&src)[C::ROWS][4], int st) {
#pragma unroll
      for (int rr = 0; rr < C::ROWS; ++rr) {
        const int r = r0 + rr * NUM_CONV;
        const uint32_t brow = sB + st * C::B_STAGE + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack_store(brow, r & 7, j, src[rr][j]);
      }
    };
    if (do_w && total_it > 0) {
      mbar_wait(wfull_bar(0), 0);
      load_pk(pk[0], 0);
    }
    for (int it = 0; it < total_it; it += 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (it + half < total_it) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (do_sts) store_pk(pk[half], stage);
          // prefetch the next k-block's packed weights into the other register set
          int ws_next = ws + 1;
          uint32_t wphase_next = wphase;
          if (ws_next == WSTAGES) { ws_next = 0; wphase_next ^= 1u; }
          if (do_w && it + half + 1 < total_it) {
            mbar_wait(wfull_bar(ws_next), wphase_next);
            load_pk(pk[half ^ 1], ws_next);
          }
          fence_async_smem();
          // one arrival per warp; __syncwarp orders the lanes' (already proxy-fenced) stores before
          // lane 0's release-arrive
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(full_bar(stage));
            if (do_w) mbar_arrive(wempty_bar(ws));  // slot `ws` was consumed by the stores above
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          ws = ws_next;
          wphase = wphase_next;
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: 8 warps = 4 lane quadrants x 2 column halves ===============
    const int quad = warp & 3;
    const int chalf = (warp - EPI_WARP0) >> 2;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
    float* params = reinterpret_cast<float*>(smem + C::SM_PARAM);
    const uint32_t stg = base + C::SM_EPI + (uint32_t)(warp - EPI_WARP0) * 2048u;
    const uint32_t tempty_leader = TWO ? mapa(tempty_bar(0), 0) : 0u;
    int local = 0;
    for (int t = unit; t < total_tiles; t += num_units, ++local) {
      const int mb = t / p.n_tiles, nb = t % p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
      float* ps = params + acc * 3 * BN;
      {
        const int n = nb * BN + et;
        const bool ok = n < p.N;
        ps[et] = ok ? __fmul_rn(p.s_w[n], 0.0625f) : 0.f;  // the INT8 operand holds 16 x code
        ps[BN + et] = (ok && p.bias != nullptr) ? __half2float(p.bias[n]) : 0.f;
        ps[2 * BN + et] = (ok && p.w0 != nullptr) ? __half2float(p.w0[n]) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int m_warp0 = mb * (C::CTAS * BM) + (int)cta * BM + quad * 32;
      const int m = m_warp0 + lane;
      float sa = 0.f, x0v = 0.f;
      if (m < p.M) {
        sa = p.s_a_per_row ? p.s_a[m] : p.s_a[0];
        if (p.x0 != nullptr) x0v = __half2float(p.x0[m]);
      }
      const bool has_r1 = p.x0 != nullptr, has_bias = p.bias != nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (MQ_PROBE(p, 32u)) {
        // probe: no TMEM drain at all -- does the drain running beside the MMAs slow the main loop?
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (!TWO || cta == 0) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(tempty_leader + 8u * acc);
        }
        continue;
      }
      const int col0 = chalf * (BN / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col0);
      uint32_t va[32], vb[32];
      tmem_ld32_nowait(taddr, va);
      auto process = [&](const uint32_t (&v)[32], int cc) {
        const int colb = col0 + cc * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = colb + g * 8;
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float sc = __fmul_rn(sa, ps[col + e]);
            float val = __fmul_rn((float)(int)v[g * 8 + e], sc);
            if (has_r1) val = __fadd_rn(val, __fmul_rn(x0v, ps[2 * BN + col + e]));
            if (has_bias) val = __fadd_rn(val, ps[BN + col + e]);
            y[e] = val;
          }
          const uint4 hv = pack8_f16(y);
          // per-warp staging: 32 rows x 64 B; 16-byte chunk g of row `lane` sits at g ^ ((lane>>1)&3)
          const uint32_t dst = stg + lane * 64 + ((uint32_t)(g ^ ((lane >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w) : "memory");
        }
        __syncwarp();
        // 4 lanes cover one row's 64 bytes: 8 rows per instruction, whole 32-byte sectors
        const int ch = lane & 3;
        const int ncol = nb * BN + colb + ch * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 8 + (lane >> 2);
          uint4 hv;
          const uint32_t src = stg + row * 64 + ((uint32_t)(ch ^ ((row >> 1) & 3)) << 4);
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "r"(src));
          const int mr = m_warp0 + row;
          if (mr < p.M && ncol < p.N && !MQ_PROBE(p, 4u)) stg_stream(p.out + (int64_t)mr * p.ldo + ncol, hv);
        }
        __syncwarp();
      };
      // TMEM loads are software pipelined: the load of chunk c+1 is in flight while c is processed
#pragma unroll 1
      for (int cc = 0; cc < (BN / 2) / 32; cc += 2) {
        tmem_wait_ld();
        tmem_ld32_nowait(taddr + (cc + 1) * 32, vb);
        process(va, cc);
        tmem_wait_ld();
        if (cc + 2 < (BN / 2) / 32) tmem_ld32_nowait(taddr + (cc + 2) * 32, va);
        process(vb, cc + 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!TWO || cta == 0) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cluster(tempty_leader + 8u * acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still touch this CTA
  if (warp == 2) {
    if (TWO)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace v2

// Only the single-CTA flavour is instantiated: the cta_group::2 branches of the template
// (TWO_CTA = true, a CTA pair sharing one converted weight tile) are correct but their TMA path
// under-performs (profiles/README.md items 5-6); they stay in the source as the starting point of
// that experiment and are not part of the library.
int launch_w4a8_gemm_v2(const CUtensorMap& tmap_a, const Params& p_in, cudaStream_t st) {
  using namespace v2;
  Params p = p_in;
  p.m_tiles = (p.M + BM - 1) / BM;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  MQ_DYN_SMEM(w4a8_gemm_v2_kernel<false>, Cfg<false>::SM_ALLOC);
  const unsigned grid = (unsigned)(tiles < num_sms() ? tiles : num_sms());
  w4a8_gemm_v2_kernel<false><<<grid, NUM_THREADS, Cfg<false>::SM_ALLOC, st>>>(tmap_a, p);
  return MQ_OK;
}

}  // namespace gemm
}  // namespace mq
