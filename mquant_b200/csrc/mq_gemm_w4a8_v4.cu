// W4A8 GEMM, 256 x 256 tile per CTA ("v4"): the converted weight tile is shared by two M halves.
//
// The v2 kernel is bound by the SM's 128 B/clk shared-memory port: per 128 x 256 x 128 k-block
// 128 KB cross it (TMA writes 16 KB of A + 16 KB of packed W, converters read 16 KB and write
// 32 KB, the tensor core reads 16 KB of A + 32 KB of B) for 512 cycles of MMA.  Half of that is the
// weight path.  This kernel converts one 256-channel weight k-block ONCE and issues the MMAs of
// two 128-row activation halves against it:
//   per k-block 2 x (A 16 KB in + 16 KB read) + W 16 + 16 + 32 + B read 2 x 32 = 192 KB for
//   2 x 512 MMA cycles  ->  96 KB per 128 x 256 x 128 unit instead of 128 KB.
// Price: the two 128 x 256 int32 accumulators fill all 512 TMEM columns, so the epilogue of a tile
// cannot overlap the main loop of the next one (only per-half: the first half drains while the
// last MMAs of the second half run; the first MMAs of the next tile start when the first half is
// drained).  Worth it from K >= 2048 on; the dispatcher uses it for M >= 512.
//
// Rings: A 3 x 32 KB, B 2 x 32 KB (A comes over TMA with L2 latency and needs the deeper ring; B is
// produced from the staging ring a few hundred cycles away), packed-W staging 2 x 16 KB (+ one
// k-block in converter registers).
//
// Same contract and bit-exact results as the other W4A8 kernels (identical epilogue arithmetic).
#include <cstdio>
#include <cstdlib>

#include "mq_gemm_common.cuh"

// Waits that normally last long (the epilogue waiting for a whole tile, the producers waiting for
// a ring slot): with MQ_GEMM_SLEEPWAIT a failed probe backs off instead of re-issuing at once
#ifdef MQ_GEMM_SLEEPWAIT
#define MQ_WAIT_LONG(bar, par) mbar_wait_sleep(bar, par)
#else
#define MQ_WAIT_LONG(bar, par) mbar_wait(bar, par)
#endif

namespace mq {
namespace gemm {
namespace v4 {

constexpr int BMH = 128;      // rows per accumulator (one tcgen05.mma M)
constexpr int BM = 2 * BMH;   // rows per tile
constexpr int BN = 256;
constexpr int BK = 128;
constexpr int UMMA_K = 32;
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4;  // 8 epilogue warps: lane quadrant = warp & 3, column half = (warp - 4) >> 2
constexpr int NUM_EPI_WARPS = 8;
constexpr int CONV_WARP0 = 12;  // 4 converter warps
constexpr int NUM_CONV = 128;
constexpr int ROWS = BN / NUM_CONV;
constexpr int TMEM_COLS = 512;
constexpr int ASTAGES = 3, BSTAGES = 2, WSTAGES = 2;
constexpr int AH_STAGE = BMH * BK;  // 16 KB per half
constexpr int A_STAGE = BM * BK;    // 32 KB
constexpr int B_STAGE = BN * BK;    // 32 KB
constexpr int W_STAGE = BN * BK / 2;
constexpr int SM_A = 0;
constexpr int SM_B = SM_A + ASTAGES * A_STAGE;
constexpr int SM_W = SM_B + BSTAGES * B_STAGE;
// Output staging per epilogue warp: 32 rows x 64 B, two buffers (warp sampling: with one buffer
// 14 % of the active epilogue time waits for the previous TMA store to have read it).  The
// bias / w0 parameter arrays exist only in the variants that use them; with both there is room
// for one staging buffer only.
constexpr int STG_BYTES = 2048;
constexpr int SM_EPI = SM_W + WSTAGES * W_STAGE;
constexpr int NUM_BARS = 2 * ASTAGES + 2 * BSTAGES + 2 * WSTAGES + 4;
template <bool HAS_R1, bool HAS_BIAS>
struct Lay {
  static constexpr int STG_BUFS = (HAS_R1 && HAS_BIAS) ? 1 : 2;
  static constexpr int SM_PARAM = SM_EPI + NUM_EPI_WARPS * STG_BUFS * STG_BYTES;  // s_w16 fp32[256]
  static constexpr int SM_BIAS = SM_PARAM + BN * 4;                               // bias fp16[256]
  static constexpr int SM_W0 = SM_BIAS + (HAS_BIAS ? BN * 2 : 0);                 // w0 fp16[256]
  static constexpr int SM_BAR = SM_W0 + (HAS_R1 ? BN * 2 : 0);
  static constexpr int SM_TMEMPTR = SM_BAR + NUM_BARS * 8;
  static constexpr int SM_TOTAL = SM_TMEMPTR + 16;
  static constexpr int SM_ALLOC = SM_TOTAL + 1024;  // slack to align the window to 1024 bytes (SWIZZLE_128B)
  static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");
};
// s8 x s8 -> s32, K-major A and B, N = 256, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BMH >> 4) << 24);

// HAS_R1 / HAS_BIAS are compile-time: one straight-line epilogue per variant (with run-time flags
// the compiler emitted four code copies per 8-column group and the epilogue warps, two per
// scheduler, stalled on instruction fetch for ~20 % of their active time -- profiles/README.md).
//
// GATED (model/modeling_qwen_opt.py:382-386): the weight rows are interleaved in
// blocks of 128 -- columns 0..127 of a tile are channels f0..f0+127 of w1, columns 128..255 the same
// channels of w2 -- and the epilogue writes  fp16(a1) * fp16(silu(fp16(a2)))  to a [M, N/2] output:
// the reference's three fp16 rounding points (both Linear outputs, F.silu's result, the product),
// silu in fp32 with torch's expression x / (1 + exp(-x)).  A thread needs a1 and a2 of the same
// channel, so a warp owns 64 gated channels (two 32-column blocks of BOTH column halves).
template <bool HAS_R1, bool HAS_BIAS, bool GATED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_v4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_out,
                    const Params p) {
  using L = Lay<HAS_R1, HAS_BIAS>;
  constexpr int SM_BAR = L::SM_BAR, SM_TMEMPTR = L::SM_TMEMPTR, SM_PARAM = L::SM_PARAM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base + SM_A, sB = base + SM_B, sW = base + SM_W;
  const uint32_t bar0 = base + SM_BAR;
  auto afull_bar = [&](int s) { return bar0 + 8u * s; };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (ASTAGES + s); };
  auto bfull_bar = [&](int s) { return bar0 + 8u * (2 * ASTAGES + s); };
  auto bempty_bar = [&](int s) { return bar0 + 8u * (2 * ASTAGES + BSTAGES + s); };
  auto wfull_bar = [&](int s) { return bar0 + 8u * (2 * ASTAGES + 2 * BSTAGES + s); };
  auto wempty_bar = [&](int s) { return bar0 + 8u * (2 * ASTAGES + 2 * BSTAGES + WSTAGES + s); };
  auto tfull_bar = [&](int h) { return bar0 + 8u * (2 * ASTAGES + 2 * BSTAGES + 2 * WSTAGES + h); };
  auto tempty_bar = [&](int h) { return bar0 + 8u * (2 * ASTAGES + 2 * BSTAGES + 2 * WSTAGES + 2 + h); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler: role code stays on the uniform datapath
  const int lane = threadIdx.x & 31;
  const int unit = (int)blockIdx.x;
  const int num_units = (int)gridDim.x;
  const int total_tiles = p.m_tiles * p.n_tiles;  // m_tiles counts 256-row tiles

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    for (int s = 0; s < ASTAGES; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < BSTAGES; ++s) {
      mbar_init(bfull_bar(s), NUM_CONV / 32);
      mbar_init(bempty_bar(s), 1);
    }
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), NUM_CONV / 32);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(tfull_bar(h), 1);
      mbar_init(tempty_bar(h), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(base + SM_TMEMPTR, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: 256 activation rows per k-block as two 128-row boxes
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        const int mb = t / p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          MQ_WAIT_LONG(aempty_bar(stage), phase ^ 1u);
          if (MQ_PROBE(p, 2u)) {
            mbar_arrive(afull_bar(stage));
          } else {
            mbar_arrive_expect_tx(afull_bar(stage), A_STAGE);
            tma_load_2d(sA + stage * A_STAGE, &tmap_a, kb * BK, mb * BM, afull_bar(stage));
            tma_load_2d(sA + stage * A_STAGE + AH_STAGE, &tmap_a, kb * BK, mb * BM + BMH, afull_bar(stage));
          }
          if (++stage == ASTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== packed-weight producer: one 16 KB bulk copy per k-block
    if (lane == 0 && !MQ_PROBE(p, 1u)) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int t = unit; t < total_tiles; t += num_units) {
        const int nb = t % p.n_tiles;
        for (int kb = 0; kb < p.k_tiles; ++kb) {
          MQ_WAIT_LONG(wempty_bar(ws), wphase ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(ws), W_STAGE);
          bulk_load(sW + ws * W_STAGE, p.w4 + w4_blob_offset(nb, kb, p.k_tiles), W_STAGE, wfull_bar(ws));
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 2 x 4 MMAs per k-block against the same B stage
    // The first and the last two k-blocks of a tile are issued half-major (h0:k, h0:k+1, h1:k,
    // h1:k+1): the first half finishes two k-blocks' worth of MMAs before the second, so its
    // epilogue starts earlier, and at the start of the next tile the first half already runs while
    // the epilogue still drains the second one.
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t tile_phase = 0;
    auto issue = [&](int h, int sa_, int sb_, int kb) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_kmajor_sw128_desc(sA + sa_ * A_STAGE + h * AH_STAGE);
        const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb_ * B_STAGE);
        const uint32_t tmem_d = tmem_base + (uint32_t)(h * BN);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          if (!MQ_PROBE(p, 8u)) umma_i8(tmem_d, adesc + 2u * k, bdesc + 2u * k, IDESC, (uint32_t)((kb | k) != 0));
        }
      }
      __syncwarp();
    };
    auto commit = [&](uint32_t bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    for (int t = unit; t < total_tiles; t += num_units, tile_phase ^= 1u) {
      const int K = p.k_tiles;
      const bool edges = K >= 4 && !MQ_PROBE(p, 128u);
      int kb = 0;
      while (kb < K) {
        if (edges && (kb == 0 || kb == K - 2)) {
          int sa1 = sa + 1, sb1 = sb + 1;
          uint32_t pa1 = pa, pb1 = pb;
          if (sa1 == ASTAGES) { sa1 = 0; pa1 ^= 1u; }
          if (sb1 == BSTAGES) { sb1 = 0; pb1 ^= 1u; }
          mbar_wait(afull_bar(sa), pa);
          mbar_wait(bfull_bar(sb), pb);
          if (kb == 0) mbar_wait(tempty_bar(0), tile_phase ^ 1u);
          issue(0, sa, sb, kb);
          mbar_wait(afull_bar(sa1), pa1);
          mbar_wait(bfull_bar(sb1), pb1);
          issue(0, sa1, sb1, kb + 1);
          if (kb != 0) commit(tfull_bar(0));
          if (kb == 0) mbar_wait(tempty_bar(1), tile_phase ^ 1u);
          issue(1, sa, sb, kb);
          commit(aempty_bar(sa));
          commit(bempty_bar(sb));
          issue(1, sa1, sb1, kb + 1);
          commit(aempty_bar(sa1));
          commit(bempty_bar(sb1));
          if (kb != 0) commit(tfull_bar(1));
          sa = sa1; pa = pa1; sb = sb1; pb = pb1;
          kb += 2;
        } else {
          mbar_wait(afull_bar(sa), pa);
          mbar_wait(bfull_bar(sb), pb);
          const bool last = kb == K - 1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (kb == 0) mbar_wait(tempty_bar(h), tile_phase ^ 1u);  // half h drained by the epilogue
            issue(h, sa, sb, kb);
            if (last) commit(tfull_bar(h));
          }
          commit(aempty_bar(sa));
          commit(bempty_bar(sb));
          kb += 1;
        }
        if (++sa == ASTAGES) { sa = 0; pa ^= 1u; }
        if (++sb == BSTAGES) { sb = 0; pb ^= 1u; }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== weight converters: 128 threads, 2 full 128-byte rows each
    const int r0 = threadIdx.x - CONV_WARP0 * 32;
    const int my_tiles = (total_tiles - unit + num_units - 1) / num_units;
    const int total_it = my_tiles * p.k_tiles;
    const bool do_w = !MQ_PROBE(p, 1u);
    const bool do_sts = !MQ_PROBE(p, 1u | 16u);
    int stage = 0, ws = 0;
    uint32_t phase = 0, wphase = 0;
    // software pipelined as in v2: staging reads of k-block it+1 are issued before the proxy fence
    // (a MEMBAR) of k-block it
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
    if (do_w && total_it > 0) {
      mbar_wait(wfull_bar(0), 0);
      load_pk(pk[0], 0);
    }
    for (int it = 0; it < total_it; it += 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (it + half < total_it) {
          // the registers hold slot `ws`: it can be refilled while we wait for a free B stage
          __syncwarp();
          if (lane == 0 && do_w) mbar_arrive(wempty_bar(ws));
          mbar_wait(bempty_bar(stage), phase ^ 1u);
          if (do_sts) store_pk(pk[half], stage);
          int ws_next = ws + 1;
          uint32_t wphase_next = wphase;
          if (ws_next == WSTAGES) { ws_next = 0; wphase_next ^= 1u; }
          if (do_w && it + half + 1 < total_it) {
            mbar_wait(wfull_bar(ws_next), wphase_next);
            load_pk(pk[half ^ 1], ws_next);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bfull_bar(stage));
          if (++stage == BSTAGES) { stage = 0; phase ^= 1u; }
          ws = ws_next;
          wphase = wphase_next;
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: 8 warps = 4 lane quadrants x 2 column halves, two row halves
    const int quad = warp & 3;
    const int chalf = (warp - EPI_WARP0) >> 2;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
    float* ps = reinterpret_cast<float*>(smem + SM_PARAM);
    __half* ps_bias = reinterpret_cast<__half*>(smem + L::SM_BIAS);
    __half* ps_w0 = reinterpret_cast<__half*>(smem + L::SM_W0);
    const uint32_t stg0 = base + SM_EPI + (uint32_t)(warp - EPI_WARP0) * (L::STG_BUFS * STG_BYTES);
    uint32_t sbuf = 0;
    constexpr bool has_r1 = HAS_R1, has_bias = HAS_BIAS;
    int local = 0;
    uint32_t tile_phase = 0;
    for (int t = unit; t < total_tiles; t += num_units, ++local, tile_phase ^= 1u) {
      const int mb = t / p.n_tiles, nb = t % p.n_tiles;
      if (local > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // the last tile's parameters are no longer read
      {
        const int n = nb * BN + et;
        const bool ok = n < p.N;
        ps[et] = ok ? __fmul_rn(p.s_w[n], 0.0625f) : 0.f;  // the INT8 operand holds 16 x code
        if (HAS_BIAS) ps_bias[et] = ok ? p.bias[n] : __float2half(0.f);
        if (HAS_R1) ps_w0[et] = ok ? p.w0[n] : __float2half(0.f);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int col0 = chalf * (BN / 2);
      // per-row factors of both halves, fetched while the main loop still runs
      float sa2[2] = {0.f, 0.f}, x0v2[2] = {0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = mb * BM + h * BMH + quad * 32 + lane;
        if (m < p.M) {
          sa2[h] = p.s_a_per_row ? p.s_a[m] : p.s_a[0];
          if (HAS_R1) x0v2[h] = __half2float(p.x0[m]);
        }
      }
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int m_warp0 = mb * BM + h * BMH + quad * 32;
        const float sa = h ? sa2[1] : sa2[0], x0v = h ? x0v2[1] : x0v2[0];
        MQ_WAIT_LONG(tfull_bar(h), tile_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * BN + col0);
        uint32_t va[32], vb[32];
        if constexpr (!GATED) tmem_ld32_nowait(taddr, va);
        const uint64_t sa2p = f2_pack(sa, sa), x0v2p = f2_pack(x0v, x0v);
        auto process = [&](const uint32_t (&v)[32], int cc) {
          const int colb = col0 + cc * 32;
          if (MQ_PROBE(p, 32u)) return;  // probe: TMEM drain only
          // the warp's staging tile (32 rows x 64 B, CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk g of
          // row r sits at g ^ ((r >> 1) & 3)) is free once the previous TMA store has read it
          const uint32_t stg = stg0 + sbuf;
          if (L::STG_BUFS == 2) {
            sbuf ^= STG_BYTES;
            if (lane == 0) bulk_wait_read_le1();  // the store issued two blocks ago has read this buffer
          } else {
            if (lane == 0) bulk_wait_read_all();  // the previous store has read the staging tile
          }
          __syncwarp();
          // s_w of the next 8 columns is fetched while the current 8 are computed (the sampled
          // profile showed the multiplies waiting on these shared-memory loads)
          float4 swa = *reinterpret_cast<const float4*>(ps + colb), swb = *reinterpret_cast<const float4*>(ps + colb + 4);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = colb + g * 8;
            const float sw[8] = {swa.x, swa.y, swa.z, swa.w, swb.x, swb.y, swb.z, swb.w};
            if (g < 3) {
              swa = *reinterpret_cast<const float4*>(ps + col + 8);
              swb = *reinterpret_cast<const float4*>(ps + col + 12);
            }
            // the two multiplications handle two columns per instruction (packed fp32: each half
            // rounds like the scalar op); the additions stay scalar -- ptxas contracts
            // mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the .rn qualifiers, which changes 1 in
            // 10^4 outputs by an fp16 ulp (caught by the GPU parity tests)
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const uint64_t sc = f2_mul(sa2p, f2_pack(sw[e], sw[e + 1]));
              f2_unpack(f2_mul(f2_pack((float)(int)v[g * 8 + e], (float)(int)v[g * 8 + e + 1]), sc), y[e], y[e + 1]);
              if (has_r1) {
                const float2 w0f = __half22float2(*reinterpret_cast<const __half2*>(ps_w0 + col + e));
                float t0, t1;
                f2_unpack(f2_mul(x0v2p, f2_pack(w0f.x, w0f.y)), t0, t1);
                y[e] = __fadd_rn(y[e], t0);
                y[e + 1] = __fadd_rn(y[e + 1], t1);
              }
              if (has_bias) {
                const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(ps_bias + col + e));
                y[e] = __fadd_rn(y[e], bf.x);
                y[e + 1] = __fadd_rn(y[e + 1], bf.y);
              }
            }
            const uint4 hv = pack8_f16(y);
            const uint32_t dst = stg + lane * 64 + ((uint32_t)(g ^ ((lane >> 1) & 3)) << 4);
            // (no "memory" clobber: the parameter loads of the next group may move above this store;
            // volatile keeps it ordered with the fence and the TMA store below)
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w));
          }
          if (!MQ_PROBE(p, 64u)) fence_async_smem();
          __syncwarp();
          // one TMA store per 32 x 32 block: whole 64-byte row segments, rows >= M and columns >= N
          // clipped by the tensor map
          if (lane == 0 && !MQ_PROBE(p, 4u)) {
            tma_store_2d(&tmap_out, stg, nb * BN + colb, m_warp0);
            bulk_commit_group();
          }
        };
        if constexpr (GATED) {
          // 16 gated channels per step: a1 from columns c .. c+15, a2 from 128 + c ..; the loads of the
          // next step are in flight while this one is computed; two steps fill one 32 x 32 output box
          (void)va;
          const uint32_t tg = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * BN + chalf * 64);
          uint32_t g1[2][16], g2[2][16];
          tmem_ld16_nowait(tg, g1[0]);
          tmem_ld16_nowait(tg + 128, g2[0]);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            tmem_wait_ld();
            if (s4 < 3) {
              tmem_ld16_nowait(tg + (s4 + 1) * 16, g1[(s4 + 1) & 1]);
              tmem_ld16_nowait(tg + 128 + (s4 + 1) * 16, g2[(s4 + 1) & 1]);
            }
            const uint32_t stg = stg0 + sbuf;
            if ((s4 & 1) == 0) {
              if (L::STG_BUFS == 2) {
                if (lane == 0) bulk_wait_read_le1();
              } else {
                if (lane == 0) bulk_wait_read_all();
              }
              __syncwarp();
            }
            const int c1 = chalf * 64 + s4 * 16;  // column of a1 inside the tile; a2 sits 128 further
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float y1[8], y2[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int col = c1 + g * 8 + e;
                y1[e] = __fmul_rn((float)(int)g1[s4 & 1][g * 8 + e], __fmul_rn(sa, ps[col]));
                y2[e] = __fmul_rn((float)(int)g2[s4 & 1][g * 8 + e], __fmul_rn(sa, ps[128 + col]));
                if (has_bias) {
                  y1[e] = __fadd_rn(y1[e], __half2float(ps_bias[col]));
                  y2[e] = __fadd_rn(y2[e], __half2float(ps_bias[128 + col]));
                }
              }
              const uint4 hv = silu_mul8(pack8_f16(y1), pack8_f16(y2));
              const uint32_t dst = stg + lane * 64 + ((uint32_t)(((s4 & 1) * 2 + g) ^ ((lane >> 1) & 3)) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w));
            }
            if (s4 & 1) {
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmap_out, stg, nb * (BN / 2) + chalf * 64 + (s4 >> 1) * 32, m_warp0);
                bulk_commit_group();
              }
              if (L::STG_BUFS == 2) sbuf ^= STG_BYTES;
            }
          }
        } else
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
        if (lane == 0) mbar_arrive(tempty_bar(h));
      }
    }
  }

  if (warp >= EPI_WARP0 && warp < CONV_WARP0 && lane == 0) bulk_wait_read_all();  // staging outlives the stores
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace v4

template <bool HAS_R1, bool HAS_BIAS, bool GATED>
static int launch_variant(const CUtensorMap& tmap_a, const CUtensorMap& tmap_out, const Params& p, int grid,
                          cudaStream_t st) {
  using namespace v4;
  auto* kern = w4a8_gemm_v4_kernel<HAS_R1, HAS_BIAS, GATED>;
  constexpr int SM_ALLOC = Lay<HAS_R1, HAS_BIAS>::SM_ALLOC;
  MQ_DYN_SMEM(kern, SM_ALLOC);
  kern<<<grid, NUM_THREADS, SM_ALLOC, st>>>(tmap_a, tmap_out, p);
  return MQ_OK;
}

int launch_w4a8_gemm_v4(const CUtensorMap& tmap_a, const CUtensorMap& tmap_out, const Params& p_in, cudaStream_t st,
                        bool gated) {
  using namespace v4;
  Params p = p_in;
  p.m_tiles = (p.M + BM - 1) / BM;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  const bool r1 = p.x0 != nullptr, bias = p.bias != nullptr;
  if (gated) {
    if (r1) return MQ_ERR_UNSUPPORTED;
    return bias ? launch_variant<false, true, true>(tmap_a, tmap_out, p, grid, st)
                : launch_variant<false, false, true>(tmap_a, tmap_out, p, grid, st);
  }
  if (r1) return bias ? launch_variant<true, true, false>(tmap_a, tmap_out, p, grid, st)
                      : launch_variant<true, false, false>(tmap_a, tmap_out, p, grid, st);
  return bias ? launch_variant<false, true, false>(tmap_a, tmap_out, p, grid, st)
              : launch_variant<false, false, false>(tmap_a, tmap_out, p, grid, st);
}

}  // namespace gemm
}  // namespace mq
