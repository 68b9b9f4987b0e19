// W4A8 GEMM with the WEIGHTS as the TMEM-resident A operand ("v5"): 128 channels x 256 tokens per CTA.
//
// Why: the 128 x 256 kernel (mq_gemm_w4a8_v2.cu) is bound by the SM's 128 B/clk shared-memory port --
// per k-block 128 KB cross it, 96 KB of them on the weight path (16 KB packed in, 16 KB read by the
// converters, 32 KB INT8 written, 32 KB read by the tensor core) -- and runs its main loop at ~1000
// cycles per k-block against 512 cycles of MMA time.  The 256 x 256 kernel halves the weight traffic per
// MAC by sharing one conversion between two M halves, but needs M >= 2 x 256 tokens per SM-wave to
// pay off; a single 1024-token request (one wave of tiles) stays on the slow kernel.
//
// tcgen05.mma takes its A operand from TENSOR MEMORY as well.  Swap the roles: D^T = W . X^T, the weight
// tile is A (M = 128 channels = TMEM lanes, K along the columns, four INT8 per 32-bit cell), the
// activations are B (N = 256 tokens, K-major SWIZZLE_128B straight from TMA).  A converter thread owns
// ONE weight row = one TMEM lane: four 128-bit loads of packed INT4 from the staging ring, two logic ops
// per eight weights, one tcgen05.st of 32 columns.  The converted weights never touch shared memory:
//
//   per k-block   packed W: 8 KB in + 8 KB read          activations: 32 KB in + 32 KB read by the MMA
//                 = 80 KB for 128 x 256 x 128 MACs (the old kernel: 128 KB), 625 cycles of port time
//
// TMEM: columns 0..255 accumulator (128 channels x 256 tokens, int32), 256..383 a 4-deep ring of
// converted weight k-blocks (32 columns each).  One accumulator: the epilogue of a tile is not
// overlapped with the next tile's main loop -- this kernel serves the one-wave regime, where a CTA
// has one tile (large batches run the 256 x 256 kernel).
//
// The accumulator is transposed (lane = channel, column = token): an epilogue thread keeps its
// channel's s_w / bias / w0 in registers and reads the per-token factors from shared memory; a warp
// writes 32 consecutive channels of one token (64 B) per store.  Same contract and bit-exact results
// as the other W4A8 kernels (identical epilogue arithmetic).
#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
namespace v5 {

constexpr int BCH = 128;  // channels per tile  (MMA M, TMEM lanes)
constexpr int BT = 256;   // tokens per tile    (MMA N)
constexpr int BK = 128;
constexpr int UMMA_K = 32;
constexpr int NUM_THREADS = 512;
constexpr int CONV_WARP0 = 4;  // 4 converter warps: warp & 3 = TMEM lane quadrant
constexpr int EPI_WARP0 = 8;   // 8 epilogue warps: lane quadrant = warp & 3, token half = (warp - 8) >> 2
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = BT;
constexpr int TSTAGES = 4, T_COLS = BK / 4;  // TMEM ring of converted weights: 32 columns per k-block
constexpr int XSTAGES = 4, X_STAGE = BT * BK;      // activations: 32 KB per stage (two 128-row TMA boxes)
constexpr int WSTAGES = 4, W_STAGE = BCH * BK / 2;  // packed weights: 8 KB per stage
constexpr int SM_X = 0;
constexpr int SM_W = SM_X + XSTAGES * X_STAGE;
constexpr int SM_TOK = SM_W + WSTAGES * W_STAGE;  // per-token factors of the tile: s_a fp32[256], x0 fp32[256]
constexpr int SM_BAR = SM_TOK + 2 * BT * 4;
constexpr int NUM_BARS = 2 * XSTAGES + 2 * WSTAGES + 2 * TSTAGES + 2;
constexpr int SM_TMEMPTR = SM_BAR + NUM_BARS * 8;
constexpr int SM_TOTAL = SM_TMEMPTR + 16;
constexpr int SM_ALLOC = SM_TOTAL + 1024;
static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(ACC_COLS + TSTAGES * T_COLS <= TMEM_COLS, "tensor memory budget");
// s8 x s8 -> s32, A (TMEM) and B K-major, M = 128, N = 256
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT >> 3) << 17) | ((uint32_t)(BCH >> 4) << 24);

__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool HAS_R1, bool HAS_BIAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_v5_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sX = base + SM_X, sW = base + SM_W;
  const uint32_t bar0 = base + SM_BAR;
  auto xfull_bar = [&](int s) { return bar0 + 8u * s; };
  auto xempty_bar = [&](int s) { return bar0 + 8u * (XSTAGES + s); };
  auto wfull_bar = [&](int s) { return bar0 + 8u * (2 * XSTAGES + s); };
  auto wempty_bar = [&](int s) { return bar0 + 8u * (2 * XSTAGES + WSTAGES + s); };
  auto afull_bar = [&](int s) { return bar0 + 8u * (2 * XSTAGES + 2 * WSTAGES + s); };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (2 * XSTAGES + 2 * WSTAGES + TSTAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * XSTAGES + 2 * WSTAGES + 2 * TSTAGES);
  const uint32_t tempty_bar = tfull_bar + 8u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int unit = (int)blockIdx.x, num_units = (int)gridDim.x;
  // p.m_tiles counts 256-token tiles, p.n_tiles 128-channel tiles (set by the launcher)
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < XSTAGES; ++s) {
      mbar_init(xfull_bar(s), 1);
      mbar_init(xempty_bar(s), 1);
    }
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), 4);  // the four converter warps
    }
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(afull_bar(s), 4);
      mbar_init(aempty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(base + SM_TMEMPTR, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a0 = tmem_base + (uint32_t)ACC_COLS;

  if (warp == 0) {
    // ===================== TMA producer: 256 tokens x 128 k of INT8 activations per k-block
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        const int mb = t / p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(xempty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(xfull_bar(stage), X_STAGE);
          tma_load_2d(sX + stage * X_STAGE, &tmap_x, kb * BK, mb * BT, xfull_bar(stage));
          tma_load_2d(sX + stage * X_STAGE + 128 * BK, &tmap_x, kb * BK, mb * BT + 128, xfull_bar(stage));
          if (++stage == XSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== packed-weight producer: this tile's 128 rows of the 256-row blob =
    // four runs (one per 32-wide k group) of 2 KB
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        const int nb = t % p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          mbar_wait(wempty_bar(ws), wphase ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(ws), W_STAGE);
          const uint8_t* blob = p.w4 + w4_blob_offset(nb >> 1, kb, p.k_tiles) + (size_t)(nb & 1) * (BCH * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            bulk_load(sW + ws * W_STAGE + j * (BCH * 16), blob + (size_t)j * (kW4TileN * 16), BCH * 16, wfull_bar(ws));
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: D[channel, token] += W[channel, k] . X[token, k]
    int xs = 0, ts = 0;
    uint32_t xphase = 0, tphase = 0;
    int local = 0;
    for (int t = unit; t < total_tiles; t += num_units, ++local) {
      mbar_wait(tempty_bar, (uint32_t)(local & 1) ^ 1u);  // the previous tile has left the accumulator
      tc_fence_after();
      for (int kb = 0; kb < p.k_tiles; ++kb) {
        mbar_wait(xfull_bar(xs), xphase);
        mbar_wait(afull_bar(ts), tphase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bdesc = make_kmajor_sw128_desc(sX + xs * X_STAGE);
          const uint32_t ta = tmem_a0 + (uint32_t)(ts * T_COLS);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_i8_ts(tmem_base, ta + (uint32_t)(k * (UMMA_K / 4)), bdesc + 2u * k, IDESC, (uint32_t)((kb | k) != 0));
          umma_commit(xempty_bar(xs));
          umma_commit(aempty_bar(ts));
          if (kb == p.k_tiles - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        if (++xs == XSTAGES) { xs = 0; xphase ^= 1u; }
        if (++ts == TSTAGES) { ts = 0; tphase ^= 1u; }
      }
    }
  } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
    // ===================== weight converters: thread = one weight row = one TMEM lane
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int my_tiles = unit < total_tiles ? (total_tiles - unit + num_units - 1) / num_units : 0;
    const int total_it = my_tiles * p.k_tiles;
    int ws = 0, ts = 0;
    uint32_t wphase = 0, tphase = 0;
    for (int it = 0; it < total_it; ++it) {
      mbar_wait(wfull_bar(ws), wphase);
      uint4 pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(pk[j].x), "=r"(pk[j].y), "=r"(pk[j].z), "=r"(pk[j].w)
                     : "r"(sW + ws * W_STAGE + (uint32_t)(j * BCH + r) * 16u));
      }
      // k order of a row: group j -> low nibbles = k 32j .. 32j+15, high nibbles = k 32j+16 .. 32j+31; values x 16
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[8 * j + 0] = (pk[j].x << 4) & 0xF0F0F0F0u;
        v[8 * j + 1] = (pk[j].y << 4) & 0xF0F0F0F0u;
        v[8 * j + 2] = (pk[j].z << 4) & 0xF0F0F0F0u;
        v[8 * j + 3] = (pk[j].w << 4) & 0xF0F0F0F0u;
        v[8 * j + 4] = pk[j].x & 0xF0F0F0F0u;
        v[8 * j + 5] = pk[j].y & 0xF0F0F0F0u;
        v[8 * j + 6] = pk[j].z & 0xF0F0F0F0u;
        v[8 * j + 7] = pk[j].w & 0xF0F0F0F0u;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(wempty_bar(ws));  // the staging slot is in registers now
      mbar_wait(aempty_bar(ts), tphase ^ 1u);      // the MMAs that read this TMEM slot have completed
      tc_fence_after();
      tmem_st32(tmem_a0 + lane_addr + (uint32_t)(ts * T_COLS), v);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull_bar(ts));
      if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
      if (++ts == TSTAGES) { ts = 0; tphase ^= 1u; }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: thread = one channel (TMEM lane), 128 tokens (columns)
    const int quad = warp & 3;
    const int thalf = (warp - EPI_WARP0) >> 2;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
    float* tok_sa = reinterpret_cast<float*>(smem + SM_TOK);
    float* tok_x0 = tok_sa + BT;
    int local = 0;
    for (int t = unit; t < total_tiles; t += num_units, ++local) {
      const int mb = t / p.n_tiles, nb = t % p.n_tiles;
      if (local > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // the last tile's factors are no longer read
      {
        const int m = mb * BT + et;
        const bool ok = m < p.M;
        tok_sa[et] = ok ? (p.s_a_per_row ? p.s_a[m] : p.s_a[0]) : 0.f;
        if (HAS_R1) tok_x0[et] = ok ? __half2float(p.x0[m]) : 0.f;
      }
      const int n = nb * BCH + quad * 32 + lane;
      const bool n_ok = n < p.N;
      const float sw16 = n_ok ? __fmul_rn(p.s_w[n], 0.0625f) : 0.f;  // the INT8 operand holds 16 x code
      float bias_n = 0.f, w0_n = 0.f;
      if (HAS_BIAS && n_ok) bias_n = __half2float(p.bias[n]);
      if (HAS_R1 && n_ok) w0_n = __half2float(p.w0[n]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull_bar, (uint32_t)(local & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(thalf * (BT / 2));
      const int m0 = mb * BT + thalf * (BT / 2);
      __half* orow = p.out + n;
      uint32_t va[32], vb[32];
      tmem_ld32_nowait(taddr, va);
      auto process = [&](const uint32_t (&v)[32], int cc) {
        const int mc = m0 + cc * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 sa4 = *reinterpret_cast<const float4*>(tok_sa + thalf * (BT / 2) + cc * 32 + e);
          const float sa[4] = {sa4.x, sa4.y, sa4.z, sa4.w};
          float x0[4] = {0.f, 0.f, 0.f, 0.f};
          if (HAS_R1) {
            const float4 x4 = *reinterpret_cast<const float4*>(tok_x0 + thalf * (BT / 2) + cc * 32 + e);
            x0[0] = x4.x; x0[1] = x4.y; x0[2] = x4.z; x0[3] = x4.w;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float val = __fmul_rn((float)(int)v[e + u], __fmul_rn(sa[u], sw16));
            if (HAS_R1) val = __fadd_rn(val, __fmul_rn(x0[u], w0_n));
            if (HAS_BIAS) val = __fadd_rn(val, bias_n);
            const int m = mc + e + u;
            if (n_ok && m < p.M) orow[(int64_t)m * p.ldo] = __float2half_rn(val);
          }
        }
      };
#pragma unroll 1
      for (int cc = 0; cc < (BT / 2) / 32; cc += 2) {
        tmem_wait_ld();
        tmem_ld32_nowait(taddr + (cc + 1) * 32, vb);
        process(va, cc);
        tmem_wait_ld();
        if (cc + 2 < (BT / 2) / 32) tmem_ld32_nowait(taddr + (cc + 2) * 32, va);
        process(vb, cc + 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace v5

template <bool HAS_R1, bool HAS_BIAS>
static int launch_v5_variant(const CUtensorMap& tmap_x, const Params& p, int grid, cudaStream_t st) {
  using namespace v5;
  auto* kern = w4a8_gemm_v5_kernel<HAS_R1, HAS_BIAS>;
  MQ_DYN_SMEM(kern, SM_ALLOC);
  kern<<<grid, NUM_THREADS, SM_ALLOC, st>>>(tmap_x, p);
  return MQ_OK;
}

int launch_w4a8_gemm_v5(const CUtensorMap& tmap_x, const Params& p_in, cudaStream_t st) {
  using namespace v5;
  Params p = p_in;
  p.m_tiles = (p.M + BT - 1) / BT;
  p.n_tiles = (p.N + BCH - 1) / BCH;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  const bool r1 = p.x0 != nullptr, bias = p.bias != nullptr;
  if (r1) return bias ? launch_v5_variant<true, true>(tmap_x, p, grid, st) : launch_v5_variant<true, false>(tmap_x, p, grid, st);
  return bias ? launch_v5_variant<false, true>(tmap_x, p, grid, st) : launch_v5_variant<false, false>(tmap_x, p, grid, st);
}

}  // namespace gemm
}  // namespace mq
