// W4A8 GEMM, cta_group::2 version: a CTA pair (thread-block cluster of 2, same TPC) owns one
// 256 x 256 output tile.
//
//   * each CTA TMA-loads its own 128 activation rows and converts only ITS half (128 output
//     channels) of the INT4 weight tile to INT8; `tcgen05.mma.cta_group::2` (M = 256, issued by
//     one thread of the leader CTA) reads A and B from both CTAs' shared memory and writes each
//     CTA's 128 x 256 accumulator into that CTA's TMEM
//   * per-CTA shared-memory work per k-block drops from 32 KB converted + 16 KB staged to
//     16 KB + 8 KB: the 1-CTA kernel is bound by LSU shared-memory wavefronts (ncu: 69 % of peak,
//     profiles/r1_gemm_full.csv), this one halves them; A and W L2 traffic per MAC also halve
//   * barriers: `full[s]` lives in the leader (expect_tx for both CTAs' TMA bytes, 2 x 256
//     converter arrivals, the peer's arrive remotely); `empty[s]` and `tmem_full[a]` are signalled
//     in both CTAs by a multicast tcgen05.commit; `tmem_empty[a]` lives in the leader (2 x 128
//     epilogue arrivals)
//   * epilogue: TMEM -> registers -> fp16 -> per-warp swizzled staging in shared memory -> row
//     segments of 128 contiguous bytes to HBM (the 1-CTA kernel's 16-byte-per-row stores cost 13 %)
//
// Same contract and bit-exact results as mq_gemm_w4a8.cu (the epilogue arithmetic is identical).
#include <cstdlib>

#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
namespace two {

constexpr int BM = 128;    // rows per CTA (256 per pair)
constexpr int BN = 256;    // output channels per pair
constexpr int BNH = 128;   // ... converted per CTA
constexpr int BK = 128;
constexpr int UMMA_K = 32;
constexpr int STAGES = 5;
constexpr int WSTAGES = 4;
constexpr int A_STAGE = BM * BK;        // 16 KB
constexpr int B_STAGE = BNH * BK;       // 16 KB
constexpr int W_STAGE = BNH * BK / 2;   // 8 KB = 4 runs of 2 KB out of the 256-row blob
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4;
constexpr int CONV_WARP0 = 8;
constexpr int NUM_CONV = 256;
constexpr int TMEM_COLS = 512;

constexpr int SM_A = 0;
constexpr int SM_B = SM_A + STAGES * A_STAGE;
constexpr int SM_W = SM_B + STAGES * B_STAGE;
constexpr int SM_EPI = SM_W + WSTAGES * W_STAGE;      // 4 warps x 32 rows x 128 B
constexpr int SM_PARAM = SM_EPI + 4 * 4096;           // 2 x {s_w16[256], bias[256], w0[256]} fp32
constexpr int SM_BAR = SM_PARAM + 2 * 3 * BN * 4;
constexpr int SM_TMEMPTR = SM_BAR + (2 * STAGES + 2 * WSTAGES + 4) * 8;
constexpr int SM_TOTAL = SM_TMEMPTR + 16;
constexpr int SM_ALLOC = SM_TOTAL + 1024;
static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");

constexpr uint32_t kInstrDesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((2 * BM) >> 4) << 24);

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
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                                 uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + SM_A, sB = base + SM_B, sW = base + SM_W;
  const uint32_t bar0 = base + SM_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto wfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto wempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + WSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 * WSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 * WSTAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_tiles = p.m_tiles * p.n_tiles;  // m_tiles counts 256-row tiles here

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + 2 * NUM_CONV);  // used in the leader only
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), NUM_CONV);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * 128);  // used in the leader only
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + SM_TMEMPTR),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (A): 128 rows per CTA, tx bytes land on the leader's barrier
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < total_tiles; t += num_pairs) {
        const int mb = t / p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (p.dbg & 2u) {
            if (cta == 0) mbar_arrive(full_bar(stage));
          } else {
            if (cta == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * A_STAGE);
            tma_load_2d_2cta(sA + stage * A_STAGE, &tmap_a, kb * BK, mb * (2 * BM) + (int)cta * BM,
                             mapa(full_bar(stage), 0));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== packed-weight producer: this CTA's 128 rows = 4 runs of 2 KB per blob
    if (lane == 0 && !(p.dbg & 1u)) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int t = pair; t < total_tiles; t += num_pairs) {
        const int nb = t % p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(wempty_bar(ws), wphase ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(ws), W_STAGE);
          const uint8_t* blob = p.w4 + w4_blob_offset(nb, kb, p.k_tiles) + (size_t)cta * (BNH * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            bulk_load(sW + ws * W_STAGE + j * (BNH * 16), blob + (size_t)j * (kW4TileN * 16), BNH * 16,
                      wfull_bar(ws));
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (cta == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = pair; t < total_tiles; t += num_pairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait_cluster(full_bar(stage), phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t adesc = make_kmajor_sw128_desc(sA + stage * A_STAGE);
            const uint64_t bdesc = make_kmajor_sw128_desc(sB + stage * B_STAGE);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (!(p.dbg & 8u))
                umma_i8_2cta(tmem_d, adesc + 2u * k, bdesc + 2u * k, kInstrDesc2, (uint32_t)((kb | k) != 0));
            }
            umma_commit_2cta(empty_bar(stage));
            if (kb == p.k_tiles - 1) umma_commit_2cta(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== weight converters: 256 threads, (row, half of the k-block) each
    const int ct = threadIdx.x - CONV_WARP0 * 32;
    const int r = ct & (BNH - 1);  // row of this CTA's 128-row B half
    const int h = ct >> 7;         // 0: k 0..63, 1: k 64..127 (warp uniform)
    const int rsw = r & 7;
    const uint32_t brow_off = (uint32_t)r * 128u;
    const uint32_t full_leader = mapa(bar0, 0);  // leader's barrier block (same offsets)
    const int my_tiles = (total_tiles - pair + num_pairs - 1) / num_pairs;
    const int total_it = my_tiles * p.k_tiles;
    const bool do_w = !(p.dbg & 1u);
    const bool do_sts = !(p.dbg & (1u | 16u));
    int stage = 0, ws = 0;
    uint32_t phase = 0, wphase = 0;
    for (int it = 0; it < total_it; ++it) {
      uint4 pk[2];
      if (do_w) {
        mbar_wait(wfull_bar(ws), wphase);
        const uint32_t src = sW + ws * W_STAGE + (uint32_t)(2 * h) * (BNH * 16) + (uint32_t)r * 16u;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(pk[jj].x), "=r"(pk[jj].y), "=r"(pk[jj].z), "=r"(pk[jj].w)
                       : "r"(src + (uint32_t)jj * (BNH * 16)));
        }
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t brow = sB + stage * B_STAGE + brow_off;
      if (do_sts) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) unpack_store(brow, rsw, 2 * h + jj, pk[jj]);
      }
      fence_async_smem();
      if (cta == 0) mbar_arrive(full_bar(stage));
      else mbar_arrive_cluster(full_leader + 8u * stage);
      if (do_w) mbar_arrive(wempty_bar(ws));
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int et = threadIdx.x - EPI_WARP0 * 32;
    float* params = reinterpret_cast<float*>(smem + SM_PARAM);
    const uint32_t stg = base + SM_EPI + quad * 4096;
    const uint32_t tempty_leader = mapa(tempty_bar(0), 0);
    int local = 0;
    for (int t = pair; t < total_tiles; t += num_pairs, ++local) {
      const int mb = t / p.n_tiles, nb = t % p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
      float* ps = params + acc * 3 * BN;
      for (int c = et; c < BN; c += 128) {
        const int n = nb * BN + c;
        const bool ok = n < p.N;
        ps[c] = ok ? __fmul_rn(p.s_w[n], 0.0625f) : 0.f;
        ps[BN + c] = (ok && p.bias != nullptr) ? __half2float(p.bias[n]) : 0.f;
        ps[2 * BN + c] = (ok && p.w0 != nullptr) ? __half2float(p.w0[n]) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int m_warp0 = mb * (2 * BM) + (int)cta * BM + quad * 32;
      const int m = m_warp0 + lane;
      float sa = 0.f, x0v = 0.f;
      if (m < p.M) {
        sa = p.s_a_per_row ? p.s_a[m] : p.s_a[0];
        if (p.x0 != nullptr) x0v = __half2float(p.x0[m]);
      }
      const bool has_r1 = p.x0 != nullptr, has_bias = p.bias != nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c2 = 0; c2 < BN / 64; ++c2) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v[32];
          tmem_ld32(taddr + c2 * 64 + hf * 32, v);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c2 * 64 + hf * 32 + g * 8;
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
            const uint32_t dst = stg + lane * 128 + ((uint32_t)((hf * 4 + g) ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w) : "memory");
          }
        }
        __syncwarp();
        // 8 lanes cover one row's 128 bytes: 4 rows per instruction, full 128-byte lines
        const int ch = lane & 7;
        const int ncol = nb * BN + c2 * 64 + ch * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3);
          uint4 hv;
          const uint32_t src = stg + row * 128 + ((uint32_t)(ch ^ (row & 7)) << 4);
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "r"(src));
          const int mr = m_warp0 + row;
          if (mr < p.M && ncol < p.N && !(p.dbg & 4u))
            stg_stream(p.out + (int64_t)mr * p.ldo + ncol, hv);
        }
        __syncwarp();
      }
      tc_fence_before();
      if (cta == 0) mbar_arrive(tempty_bar(acc));
      else mbar_arrive_cluster(tempty_leader + 8u * acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still touch this CTA
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace two

int launch_w4a8_gemm_2cta(const CUtensorMap& tmap_a, const Params& p_in, cudaStream_t st) {
  using namespace two;
  Params p = p_in;
  p.m_tiles = (p.M + 2 * BM - 1) / (2 * BM);  // 256-row tiles
  MQ_CUDA(cudaFuncSetAttribute(w4a8_gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC));
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  const int max_pairs = num_sms() / 2;
  const int pairs = (int)(tiles < max_pairs ? tiles : max_pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SM_ALLOC;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MQ_CUDA(cudaLaunchKernelEx(&cfg, w4a8_gemm_2cta_kernel, tmap_a, p));
  return MQ_OK;
}

}  // namespace gemm
}  // namespace mq
