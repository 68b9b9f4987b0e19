// W4A8 GEMM, split-K flavour of the 128 x 256 kernel for mid-sized M (17 .. ~1000 rows).
//
// With M/128 * N/256 tiles well below the SM count the persistent kernels leave most of the
// machine idle and a single CTA walks all K/128 k-blocks of its tile (K = 11008: 86 x ~1000
// cycles = 45 us).  Here a work item is (tile, K slice): every CTA accumulates its slice in TMEM
// exactly like mq_gemm_w4a8_v2.cu, adds the int32 partial sums to a caller-provided, self-cleaning
// workspace (integer adds: exact and order independent; layout [tile][column][row] so that the
// 32 lanes of a warp hit 128 consecutive bytes), and the CTA that arrives last at the tile's
// counter reads the sums back, clears them and applies the usual epilogue -> bit-identical
// results.  Roles, rings and the weight path are those of the v2 kernel.
#include <cstdio>
#include <cstdlib>

#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
namespace sk {

constexpr int BM = 128, BN = 256, BK = 128, UMMA_K = 32;
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4, NUM_EPI_WARPS = 8;
constexpr int CONV_WARP0 = 12, NUM_CONV = 128;
constexpr int ROWS = BN / NUM_CONV;
constexpr int TMEM_COLS = 512;
constexpr int STAGES = 3, WSTAGES = 3;
constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK, W_STAGE = BN * BK / 2;
constexpr int SM_A = 0;
constexpr int SM_B = SM_A + STAGES * A_STAGE;
constexpr int SM_W = SM_B + STAGES * B_STAGE;
constexpr int SM_EPI = SM_W + WSTAGES * W_STAGE;              // 8 warps x 32 rows x 64 B
constexpr int SM_PARAM = SM_EPI + NUM_EPI_WARPS * 2048;       // {s_w16, bias, w0}[256] fp32
constexpr int SM_BAR = SM_PARAM + 3 * BN * 4;
constexpr int NUM_BARS = 2 * STAGES + 2 * WSTAGES + 4;
constexpr int SM_FLAG = SM_BAR + NUM_BARS * 8;
constexpr int SM_TMEMPTR = SM_FLAG + 8;
constexpr int SM_TOTAL = SM_TMEMPTR + 16;
constexpr int SM_ALLOC = SM_TOTAL + 1024;
static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Split {
  int ksplit, kbs;    // K slices per tile, k-blocks per slice
  int* ws_acc;        // [tiles][256 columns][128 rows] int32, zero on entry, zero on exit
  unsigned* ws_cnt;   // [tiles] arrival counters, zero on entry, zero on exit
  int no_atomics;     // MQ_SK_PROBE=1: timing probe, wrong results
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_a, const Params p, const Split sp) {
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
  volatile uint32_t* last_flag = reinterpret_cast<volatile uint32_t*>(smem + SM_FLAG);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int total_work = p.m_tiles * p.n_tiles * sp.ksplit;
  // work item w -> tile t = w / ksplit (mb = t / n_tiles, nb = t % n_tiles), k-blocks [kb0, kb1)
  auto k_range = [&](int w, int& t, int& kb0, int& kb1) {
    t = w / sp.ksplit;
    const int ks = w - t * sp.ksplit;
    kb0 = ks * sp.kbs;
    kb1 = min(kb0 + sp.kbs, p.k_tiles);
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + NUM_CONV / 32);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), NUM_CONV / 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(base + SM_TMEMPTR, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (A) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int t, kb0, kb1;
        k_range(w, t, kb0, kb1);
        const int mb = t / p.n_tiles;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), A_STAGE);
          tma_load_2d(sA + stage * A_STAGE, &tmap_a, kb * BK, mb * BM, full_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== packed-weight producer =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int t, kb0, kb1;
        k_range(w, t, kb0, kb1);
        const int nb = t % p.n_tiles;
        for (int kb = kb0; kb < kb1; ++kb) {
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
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++local) {
      int t, kb0, kb1;
      k_range(w, t, kb0, kb1);
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_kmajor_sw128_desc(sA + stage * A_STAGE);
          const uint64_t bdesc = make_kmajor_sw128_desc(sB + stage * B_STAGE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_i8(tmem_d, adesc + 2u * k, bdesc + 2u * k, IDESC, (uint32_t)(kb != kb0 || k != 0));
          umma_commit(empty_bar(stage));
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== weight converters (as in v2) =====================
    const int r0 = threadIdx.x - CONV_WARP0 * 32;
    int total_it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int t, kb0, kb1;
      k_range(w, t, kb0, kb1);
      total_it += kb1 - kb0;
    }
    int stage = 0, ws = 0;
    uint32_t phase = 0, wphase = 0;
    uint4 pk[2][ROWS][4];
    auto load_pk = [&](uint4 (&dst)[ROWS][4], int slot) {
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const uint32_t src = sW + slot * W_STAGE + (uint32_t)(r0 + rr * NUM_CONV) * 16u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(dst[rr][j].x), "=r"(dst[rr][j].y), "=r"(dst[rr][j].z), "=r"(dst[rr][j].w)
                       : "r"(src + (uint32_t)j * (BN * 16)));
        }
      }
    };
    auto store_pk = [&](const uint4 (&src)[ROWS][4], int st) {
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const int r = r0 + rr * NUM_CONV;
        const uint32_t brow = sB + st * B_STAGE + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack_store(brow, r & 7, j, src[rr][j]);
      }
    };
    if (total_it > 0) {
      mbar_wait(wfull_bar(0), 0);
      load_pk(pk[0], 0);
    }
    for (int it = 0; it < total_it; it += 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (it + half < total_it) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          store_pk(pk[half], stage);
          int ws_next = ws + 1;
          uint32_t wphase_next = wphase;
          if (ws_next == WSTAGES) { ws_next = 0; wphase_next ^= 1u; }
          if (it + half + 1 < total_it) {
            mbar_wait(wfull_bar(ws_next), wphase_next);
            load_pk(pk[half ^ 1], ws_next);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(full_bar(stage));
            mbar_arrive(wempty_bar(ws));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          ws = ws_next;
          wphase = wphase_next;
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: partial sums -> workspace; last CTA of a tile finishes it ====
    const int quad = warp & 3;
    const int chalf = (warp - EPI_WARP0) >> 2;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
    float* ps = reinterpret_cast<float*>(smem + SM_PARAM);
    const uint32_t stg = base + SM_EPI + (uint32_t)(warp - EPI_WARP0) * 2048u;
    const bool has_r1 = p.x0 != nullptr, has_bias = p.bias != nullptr;
    const int col0 = chalf * (BN / 2);
    const int row = quad * 32 + lane;
    int local = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++local) {
      int t, kb0, kb1;
      k_range(w, t, kb0, kb1);
      const int mb = t / p.n_tiles, nb = t % p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
      int* wsbase = sp.ws_acc + ((size_t)t * BN + col0) * BM + row;  // + column * 128
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col0);
#pragma unroll 1
      for (int cc = 0; cc < (BN / 2) / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(taddr + cc * 32, v);
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (v[e] != 0u && !sp.no_atomics) atomicAdd(wsbase + (cc * 32 + e) * BM, (int)v[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) *last_flag = (atomicAdd(sp.ws_cnt + t, 1u) == (unsigned)(sp.ksplit - 1)) ? 1u : 0u;
      {
        // column parameters of this tile (only the finishing CTA uses them)
        const int n = nb * BN + et;
        const bool ok = n < p.N;
        ps[et] = ok ? __fmul_rn(p.s_w[n], 0.0625f) : 0.f;  // the INT8 operand holds 16 x code
        ps[BN + et] = (ok && has_bias) ? __half2float(p.bias[n]) : 0.f;
        ps[2 * BN + et] = (ok && has_r1) ? __half2float(p.w0[n]) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool last = *last_flag != 0u;
      if (last) {
        __threadfence();
        if (et == 0) sp.ws_cnt[t] = 0;  // leave the workspace clean for the next call
        const int m_warp0 = mb * BM + quad * 32;
        const int m = m_warp0 + lane;
        float sa = 0.f, x0v = 0.f;
        if (m < p.M) {
          sa = p.s_a_per_row ? p.s_a[m] : p.s_a[0];
          if (has_r1) x0v = __half2float(p.x0[m]);
        }
#pragma unroll 1
        for (int cc = 0; cc < (BN / 2) / 32; ++cc) {
          const int colb = col0 + cc * 32;
          int v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __ldcg(wsbase + (cc * 32 + e) * BM);
#pragma unroll
          for (int e = 0; e < 32; ++e) wsbase[(cc * 32 + e) * BM] = 0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = colb + g * 8;
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float sc = __fmul_rn(sa, ps[col + e]);
              float val = __fmul_rn((float)v[g * 8 + e], sc);
              if (has_r1) val = __fadd_rn(val, __fmul_rn(x0v, ps[2 * BN + col + e]));
              if (has_bias) val = __fadd_rn(val, ps[BN + col + e]);
              y[e] = val;
            }
            const uint4 hv = pack8_f16(y);
            const uint32_t dst = stg + lane * 64 + ((uint32_t)(g ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w) : "memory");
          }
          __syncwarp();
          const int ch = lane & 3;
          const int ncol = nb * BN + colb + ch * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rw = i * 8 + (lane >> 2);
            uint4 hv;
            const uint32_t src = stg + rw * 64 + ((uint32_t)(ch ^ ((rw >> 1) & 3)) << 4);
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "r"(src));
            const int mr = m_warp0 + rw;
            if (mr < p.M && ncol < p.N) stg_stream(p.out + (int64_t)mr * p.ldo + ncol, hv);
          }
          __syncwarp();
        }
      }
      // the parameters / flag are rewritten for the next work item only after everyone is done
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace sk

// Split factor for a shape, 1 = not worth it (enough tiles, or K too short)
int splitk_factor(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ((M + sk::BM - 1) / sk::BM) * ((N + sk::BN - 1) / sk::BN);
  const int k_tiles = (int)(K / sk::BK);
  if (3 * tiles > 2 * num_sms() || k_tiles < 8) return 1;
  int ksplit = (int)(num_sms() / tiles);
  if (ksplit > k_tiles / 4) ksplit = k_tiles / 4;
  // measured (tools/latency.py, hot code / cold weights): the fixed cost of the reduction (atomics,
  // fence, counter, final pass by one CTA) is ~8 us -- a 2- or 3-way split only pays for long K
  if (ksplit < 2 || (ksplit < 4 && k_tiles < 64)) return 1;
  return ksplit;
}

size_t splitk_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (splitk_factor(M, N, K) < 2) return 0;
  const int64_t tiles = ((M + sk::BM - 1) / sk::BM) * ((N + sk::BN - 1) / sk::BN);
  return (size_t)(tiles * sk::BM * sk::BN * sizeof(int) + tiles * sizeof(unsigned));
}

int launch_w4a8_gemm_sk(const CUtensorMap& tmap_a, const Params& p_in, void* workspace, cudaStream_t st) {
  using namespace sk;
  Params p = p_in;
  p.m_tiles = (p.M + BM - 1) / BM;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  Split sp;
  const int want = splitk_factor(p.M, p.N, p.K);
  sp.kbs = (p.k_tiles + want - 1) / want;
  sp.ksplit = (p.k_tiles + sp.kbs - 1) / sp.kbs;
  sp.ws_acc = reinterpret_cast<int*>(workspace);
  sp.ws_cnt = reinterpret_cast<unsigned*>(sp.ws_acc + (size_t)tiles * BM * BN);
  sp.no_atomics = 0;
  const int64_t work = tiles * sp.ksplit;
  MQ_DYN_SMEM(w4a8_gemm_sk_kernel, SM_ALLOC);
  const int grid = (int)(work < num_sms() ? work : num_sms());
  w4a8_gemm_sk_kernel<<<grid, NUM_THREADS, SM_ALLOC, st>>>(tmap_a, p, sp);
  return MQ_OK;
}

}  // namespace gemm
}  // namespace mq
