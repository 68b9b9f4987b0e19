// W4A8 GEMM, third kernel: no shared-memory staging of the packed weights.
//
// What the profiles of v1/v2 showed (profiles/README.md): at 128 x 256 x 128 per k-block the
// kernel is bound by the SM's single 128 B/clk shared-memory port, which every client shares:
//     TMA write of A 16 KB + bulk-copy write of packed W 16 KB + LDS of packed W 16 KB
//     + STS of converted B 32 KB + tensor-core reads of A and B 48 KB  = 128 KB = 1024 cycles
// against 512 cycles of MMA time (measured: ~1000 cycles per k-block).  The staging ring costs
// 32 KB of that.  v1's register prefetch avoided the staging but stalled on
// `fence.proxy.async` (= MEMBAR.ALL.CTA in SASS, which waits for the outstanding LDGs).
//
// v3 keeps the register path and removes the stall by ORDER and by ALTERNATION:
//   * two converter groups (4 warps each) take alternate k-blocks; a group's sequence is
//       wait(empty) -> unpack + STS -> fence.proxy.async -> arrive(full) -> LDG (its NEXT k-block)
//     so no LDG is in flight when the MEMBAR executes (the previous ones were consumed by the
//     STS just before it), and the new LDGs have two k-block times (~1600 cycles) to land
//   * shared memory per k-block drops to 96 KB (768 cycles), and the freed 48 KB buys a 4-deep
//     A/B ring
// Roles (512 threads): warp 0 TMA (A), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// epilogue (TMEM loads double buffered, output staged per warp and written as 64-byte row
// segments), warps 8-11 / 12-15 converter groups 0 / 1 (two 128-byte rows of B per thread).
// Same contract and bit-exact results as the other kernels (identical epilogue arithmetic).
#include <cstdio>
#include <cstdlib>

#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
namespace v3 {

constexpr int BM = 128, BN = 256, BK = 128, UMMA_K = 32;
constexpr int STAGES = 4;  // even: group g owns stages g and g + 2
constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4, NUM_EPI_WARPS = 4;
constexpr int CONV_WARP0 = 8, CONV_WARPS_PER_GROUP = 4;
constexpr int TMEM_COLS = 512;

constexpr int SM_A = 0;
constexpr int SM_B = SM_A + STAGES * A_STAGE;
constexpr int SM_EPI = SM_B + STAGES * B_STAGE;           // 4 warps x 32 rows x 64 B
constexpr int SM_PARAM = SM_EPI + NUM_EPI_WARPS * 2048;   // 2 x {s_w16, bias, w0}[256] fp32
constexpr int SM_BAR = SM_PARAM + 2 * 3 * BN * 4;
constexpr int SM_TMEMPTR = SM_BAR + (2 * STAGES + 4) * 8;
constexpr int SM_TOTAL = SM_TMEMPTR + 16;
constexpr int SM_ALLOC = SM_TOTAL + 1024;
static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");

constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);

__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_v3_kernel(const __grid_constant__ CUtensorMap tmap_a, const Params p) {
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

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + CONV_WARPS_PER_GROUP);  // TMA thread + the owning group's warps
      mbar_init(empty_bar(s), 1);
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
            if (!(p.dbg & 8u)) umma_i8(tmem_d, adesc + 2u * k, bdesc + 2u * k, IDESC, (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(stage));
          if (kb == p.k_tiles - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== weight converters: two groups, alternate k-blocks =====================
    const int grp = (warp - CONV_WARP0) / CONV_WARPS_PER_GROUP;
    const int r0 = (threadIdx.x - CONV_WARP0 * 32) & 127;  // rows r0 and r0 + 128 of the 256-row tile
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_it = my_tiles * p.k_tiles;
    const bool do_w = !(p.dbg & 1u);
    const bool do_sts = !(p.dbg & (1u | 16u));
    uint4 pk[2][4];
    auto load_pk = [&](int it) {  // packed weights of flattened k-block index `it`, straight from L2
      if (it < total_it && do_w) {
        const int lt = it / p.k_tiles, kb = it - lt * p.k_tiles;
        const int nb = ((int)blockIdx.x + lt * (int)gridDim.x) % p.n_tiles;
        const uint8_t* blob = p.w4 + w4_blob_offset(nb, kb, p.k_tiles);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) pk[rr][j] = ldg_stream(blob + (size_t)w4_chunk_offset(j, r0 + rr * 128));
        }
      }
    };
    load_pk(grp);
    for (int it = grp; it < total_it; it += 2) {
      const int stage = it & (STAGES - 1);
      const uint32_t phase = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(stage), phase ^ 1u);
      if (do_sts) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = r0 + rr * 128;
          const uint32_t brow = sB + stage * B_STAGE + (uint32_t)r * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) unpack_store(brow, r & 7, j, pk[rr][j]);
        }
      }
      fence_async_smem();  // MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: nothing but the STS above is outstanding
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      load_pk(it + 2);  // lands while the other group converts the next k-block
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: 4 warps, one TMEM lane quadrant each =====================
    const int quad = warp & 3;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    float* params = reinterpret_cast<float*>(smem + SM_PARAM);
    const uint32_t stg = base + SM_EPI + (uint32_t)quad * 2048u;
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
      const int m_warp0 = mb * BM + quad * 32;
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
      uint32_t va[32], vb[32];
      tmem_ld32_nowait(taddr, va);
      auto process = [&](const uint32_t (&v)[32], int cc) {
        const int colb = cc * 32;
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
          const uint32_t dst = stg + lane * 64 + ((uint32_t)(g ^ ((lane >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w) : "memory");
        }
        __syncwarp();
        const int ch = lane & 3;
        const int ncol = nb * BN + colb + ch * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 8 + (lane >> 2);
          uint4 hv;
          const uint32_t src = stg + row * 64 + ((uint32_t)(ch ^ ((row >> 1) & 3)) << 4);
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "r"(src));
          const int mr = m_warp0 + row;
          if (mr < p.M && ncol < p.N && !(p.dbg & 4u)) stg_stream(p.out + (int64_t)mr * p.ldo + ncol, hv);
        }
        __syncwarp();
      };
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; cc += 2) {
        tmem_wait_ld();
        tmem_ld32_nowait(taddr + (cc + 1) * 32, vb);
        process(va, cc);
        tmem_wait_ld();
        if (cc + 2 < BN / 32) tmem_ld32_nowait(taddr + (cc + 2) * 32, va);
        process(vb, cc + 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace v3

int launch_w4a8_gemm_v3(const CUtensorMap& tmap_a, const Params& p_in, cudaStream_t st) {
  using namespace v3;
  Params p = p_in;
  p.m_tiles = (p.M + BM - 1) / BM;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  MQ_CUDA(cudaFuncSetAttribute(w4a8_gemm_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC));
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  w4a8_gemm_v3_kernel<<<grid, NUM_THREADS, SM_ALLOC, st>>>(tmap_a, p);
  return MQ_OK;
}

}  // namespace gemm
}  // namespace mq
