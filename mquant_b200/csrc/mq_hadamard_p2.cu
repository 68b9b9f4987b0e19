// Online Walsh-Hadamard rotation for power-of-two sizes n = 64 A (A = 16, 32, 64, 128: n = 1024 ... 8192),
// fused with the static INT8 quantiser: BOTH Kronecker factors on tcgen05, fp32-grade accuracy.
//
//   y = x . H_n * norm   (matmul_hadU_cuda with K == 1, fake_quant/hadamard_utils.py:117-120: the
//   fast_hadamard_transform butterflies in fp32, ONE rounding to fp16), then UniformQuantizer.quant.
//
// Why (profiles/README.md, round 2): the mma.sync kernel (mq_hadamard_mma.cu) spends ~12 thread
// instructions per element and is issue bound at 0.49-0.58 of the HBM roofline.  H_n = H_A (x) H_64
// over the index split i = 64 k + j, so a token is the A x 64 matrix X[k, j] and y = H_A . X . H_64:
//
//   TMA    tile = 128 rows of 64 elements: the A rows of S = 128 / A ADJACENT tokens (one 3-D box;
//          rows past n_in / 64 and tokens past M are out of bounds = zero filled)
//   MMA1   D1[(s,k), j'] = X[(s,k), j] . H_64[j, j']     M = 128, N = 64, K = 64: exact products, fp32
//          accumulation
//   epi-1  thread = one row (TMEM lane): v = D1 / 64 (exact, |v| <= 65504) is split into hi = fp16(v) and
//          lo' = fp16(2^11 (v - hi)) -- |lo'| <= |v|, no underflow before |v| < 2^-35 -- and written as
//          rows (s,k) and 128 + (s,k) of T, the MN-major B operand of MMA2
//   MMA2   D2[(s,k'), (p,j')] = (I_S (x) H_A)[(s,k'), (s,k)] . T[(s,k), (p,j')]    M = 128, N = 128 (two
//          tiles p = 0, 1: a job is 2 S tokens), K = 128, twice: first the lo' rows, then the hi rows
//          with  D = A.B + D * 2^-11  (the scale-input-d operand of tcgen05.mma) on their first k-step:
//          D2 = H . (hi + lo) to ~2^-22 relative, one A operand for both passes.  For A < 128 the A
//          operand is block diagonal: S tokens share the 128 lanes, so every warp of the second
//          epilogue has a full quadrant of rows.
//   epi-2  thread = one output row k' of one token: 64 consecutive j' -> * (64 norm) -> fp16 (THE
//          rounding point) -> fp16 store and / or the exact INT8 quantiser with the token's MSQ scale ->
//          256-bit stores (a warp writes 32 rows x 64 outputs of a token: contiguous)
//
// Persistent, one CTA per SM, 512 threads: warps 0-3 first epilogue, 4-11 second epilogue (set = tile
// p), 12 TMA, 13 MMA1 issuer (+ TMEM allocation), 14 MMA2 issuer.  TMEM: 2 D1 tiles (64 columns
// each) + 3 D2 jobs (128 columns each).  Shared memory: I_S (x) H_A 32 KB (its top-left block doubles as
// H_64 for A >= 64; A < 64 keeps a separate 8 KB H_64), X ring 4 x 16 KB (3 for A < 64), T ring 2 x 64 KB.
// HBM traffic per token: 2*n_in read + n (codes) [+ 2*n fp16].
#include <cuda.h>

#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"
#include "mq_hadamard_epi.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace hadp2 {
using namespace mq::tc;
using namespace mq::hadepi;

constexpr int NUM_THREADS = 512;
// issuers at the highest warp ids, then the second epilogue, the first epilogue lowest (see mq_hadamard_tc3.cu)
constexpr int EPI1_WARP0 = 0, EPI2_WARP0 = 4, W_TMA = 12, W_MMA1 = 13, W_MMA2 = 14;
constexpr int TMEM_COLS = 512;
constexpr int D1N = 2, D2N = 3;
constexpr int D2_COL0 = D1N * 64;
constexpr int TJ = 2;
constexpr int X_SLOT = 128 * 128;      // 128 rows x 64 fp16
constexpr int T_BLK = 256 * 128;       // one 64-column block of T: 128 hi rows + 128 lo' rows, 128 B each
constexpr int T_JOB = 2 * T_BLK;
constexpr int A2_KB = 128 * 128;       // one k-block (64 k) of the A operand: 128 rows x 128 B
// For A >= 64 the top-left 64 x 64 block of I_S (x) H_A IS H_64 (rows 0..63 of k-block 0, the same
// K-major layout): MMA1 reads its B operand there and the 8 KB buy a fourth X slot (loads in flight).
template <int A>
struct Lay {
  static constexpr bool ALIAS_H = A >= 64;
  static constexpr int XS = ALIAS_H ? 4 : 3;
  static constexpr int SM_A2 = 0;
  static constexpr int SM_H = ALIAS_H ? SM_A2 : SM_A2 + 2 * A2_KB;
  static constexpr int SM_X = SM_A2 + 2 * A2_KB + (ALIAS_H ? 0 : 64 * 128);
  static constexpr int SM_T = SM_X + XS * X_SLOT;
  static constexpr int SM_BAR = SM_T + TJ * T_JOB;
  static constexpr int NUM_BARS = 2 * XS + 2 * D1N + 2 * TJ + 2 * D2N;
  static constexpr int SM_TMEMPTR = SM_BAR + NUM_BARS * 8;
  static constexpr int SM_TOTAL = SM_TMEMPTR + 16;
  static constexpr int SM_ALLOC = SM_TOTAL + 1024;  // slack to align the window to 1024 bytes (SWIZZLE_128B)
  static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");
};

// D = A.B + D * 2^-11
__device__ __forceinline__ void umma_f16_scale_d11(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}

// (v0, v1) fp32 -> hi = fp16 pair, lo' = fp16 pair of 2^11 (v - hi)   (v - float(hi) is exact in fp32)
__device__ __forceinline__ void split_pair(uint32_t b0, uint32_t b1, uint32_t& hi, uint32_t& lo) {
  float v0, v1;  // v = D1 / 64 (exact): |v| <= 65504
  f2_unpack(f2_mul(f2_pack(__uint_as_float(b0), __uint_as_float(b1)), f2_pack(0.015625f, 0.015625f)), v0, v1);
  hi = f16x2_rn(v0, v1);
  float r0, r1;  // hi - v (mixed-precision subtract: no separate conversion)
  asm("{.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tsub.rn.f32.f16 %0, l, %3;\n\tsub.rn.f32.f16 %1, h, %4;}"
      : "=f"(r0), "=f"(r1)
      : "r"(hi), "f"(v0), "f"(v1));
  float s0, s1;
  f2_unpack(f2_mul(f2_pack(r0, r1), f2_pack(-2048.0f, -2048.0f)), s0, s1);
  lo = f16x2_rn(s0, s1);
}

template <int A>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hadp2_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const HadArgs a, const float norm2, const int num_jobs) {
  constexpr int S = 128 / A;  // tokens per tile
  constexpr int JT = 2 * S;   // tokens per job
  using L = Lay<A>;
  constexpr int XS = L::XS, SM_A2 = L::SM_A2, SM_H = L::SM_H, SM_X = L::SM_X, SM_T = L::SM_T, SM_BAR = L::SM_BAR,
                SM_TMEMPTR = L::SM_TMEMPTR;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA2 = base + SM_A2, sH = base + SM_H, sX = base + SM_X, sT = base + SM_T;
  const uint32_t bar0 = base + SM_BAR;
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (XS + s); };
  auto d1_full = [&](int s) { return bar0 + 8u * (2 * XS + s); };
  auto d1_empty = [&](int s) { return bar0 + 8u * (2 * XS + D1N + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * XS + 2 * D1N + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * XS + 2 * D1N + TJ + s); };
  auto d2_full = [&](int s) { return bar0 + 8u * (2 * XS + 2 * D1N + 2 * TJ + s); };
  auto d2_empty = [&](int s) { return bar0 + 8u * (2 * XS + 2 * D1N + 2 * TJ + D2N + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int my_jobs = (int)blockIdx.x < num_jobs ? (num_jobs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- one-time setup: barriers, TMEM, the two +-1 matrices --------------------------------------
  if (warp == W_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < XS; ++s) {
      mbar_init(x_full(s), 1);
      mbar_init(x_empty(s), 1);
    }
    for (int s = 0; s < D1N; ++s) {
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), 4);  // the four first-epilogue warps
    }
    for (int s = 0; s < TJ; ++s) {
      mbar_init(t_full(s), 8);  // four first-epilogue warps x two tiles
      mbar_init(t_empty(s), 1);
    }
    for (int s = 0; s < D2N; ++s) {
      mbar_init(d2_full(s), 1);
      mbar_init(d2_empty(s), 8);  // the eight second-epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first XS tiles are requested before the matrices are staged: a cold load (tensor map, HBM
    // page) takes longer than the rest of the setup
    for (int t = 0; t < XS && t < 2 * my_jobs; ++t) {
      mbar_arrive_expect_tx(x_full(t), (uint32_t)X_SLOT);
      tma_load_3d(sX + t * X_SLOT, &tmap_x, 0, 0, ((int)blockIdx.x + (t >> 1) * (int)gridDim.x) * JT + (t & 1) * S, x_full(t));
    }
  }
  if (warp == W_MMA1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  {
    if (!L::ALIAS_H) {
      // H_64 (symmetric): row j', 64 k = 128 B, K-major SWIZZLE_128B; 512 threads = 64 rows x 8 chunks
      const int row = threadIdx.x >> 3, ch = threadIdx.x & 7;
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k0 = ch * 8 + 2 * e;
        const uint32_t lo = (__popc(row & k0) & 1) ? 0xBC00u : 0x3C00u;
        const uint32_t hi = (__popc(row & (k0 + 1)) & 1) ? 0xBC00u : 0x3C00u;
        w[e] = lo | (hi << 16);
      }
      st_shared_v4(sH + row * 128 + ((uint32_t)(ch ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
    }
    // I_S (x) H_A: [k-block kb][row (s,k')][64 k] K-major SWIZZLE_128B; 128 rows x 16 chunks of 8 k
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int id = threadIdx.x + it * NUM_THREADS;
      const int r = id >> 4, c = id & 15;
      uint32_t v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k0 = c * 8 + 2 * e;  // (k0, k0 + 1 are in the same diagonal block: A >= 16)
        const bool on = (r / A) == (k0 / A);
        const uint32_t lo = (__popc((r & (A - 1)) & (k0 & (A - 1))) & 1) ? 0xBC00u : 0x3C00u;
        const uint32_t hi = (__popc((r & (A - 1)) & ((k0 + 1) & (A - 1))) & 1) ? 0xBC00u : 0x3C00u;
        v[e] = on ? (lo | (hi << 16)) : 0u;
      }
      st_shared_v4(sA2 + (uint32_t)(c >> 3) * A2_KB + (uint32_t)r * 128u + ((uint32_t)((c & 7) ^ (r & 7)) << 4), v[0], v[1], v[2], v[3]);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ===================== TMA producer: one tile (S adjacent tokens x A rows x 128 B) per step ======
    if (lane == 0) {  // (the lane that requested the first XS tiles during the setup)
      int xs = 0;
      uint32_t xph = 1;
      for (int t = XS; t < 2 * my_jobs; ++t) {
        const int tok0 = ((int)blockIdx.x + (t >> 1) * (int)gridDim.x) * JT;
        mbar_wait_sleep(x_empty(xs), xph ^ 1u);
        mbar_arrive_expect_tx(x_full(xs), (uint32_t)X_SLOT);
        tma_load_3d(sX + xs * X_SLOT, &tmap_x, 0, 0, tok0 + (t & 1) * S, x_full(xs));
        if (++xs == XS) { xs = 0; xph ^= 1u; }
      }
    }
  } else if (warp == W_MMA1) {
    // ===================== MMA1 issuer: D1 = X . H64 ================================================
    // F16 x F16 -> F32, A K-major, B K-major, M = 128, N = 64
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int xs = 0, d1 = 0;
    uint32_t xph = 0, d1ph = 0;
    for (int g = 0; g < 2 * my_jobs; ++g) {
      mbar_wait_sleep(x_full(xs), xph);
      mbar_wait_sleep(d1_empty(d1), d1ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xa = sX + xs * X_SLOT;
        const uint32_t tm = tmem_base + (uint32_t)(d1 * 64);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tm, make_sw128_desc(xa + ks * 32, 0, 1024), make_sw128_desc(sH + ks * 32, 0, 1024), idesc, (uint32_t)(ks != 0));
        umma_commit(x_empty(xs));
        umma_commit(d1_full(d1));
      }
      __syncwarp();
      if (++xs == XS) { xs = 0; xph ^= 1u; }
      if (++d1 == D1N) { d1 = 0; d1ph ^= 1u; }
    }
  } else if (warp == W_MMA2) {
    // ===================== MMA2 issuer: D2 = (I (x) H_A) . (lo' rows), then . (hi rows) with D * 2^-11
    // A K-major, B MN-major (T rows = k, 64 j' per 128 B, column blocks T_BLK apart), M = 128, N = 128
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_sw128_desc(sA2, 0, 1024);
    int tg = 0, d2 = 0;
    uint32_t tph = 0, d2ph = 0;
    for (int g = 0; g < my_jobs; ++g) {
      mbar_wait_sleep(t_full(tg), tph);
      mbar_wait_sleep(d2_empty(d2), d2ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tm = tmem_base + (uint32_t)(D2_COL0 + d2 * 128);
        const uint64_t bd = make_sw128_desc(sT + tg * T_JOB, T_BLK, 1024);
        // descriptor address fields are in 16-byte units: k-step ks reads the A operand at k-block
        // ks / 4, byte 32 (ks % 4) of the row, and rows 16 ks .. of T (2048 B per k-step)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16(tm, ad + (uint64_t)(((ks >> 2) * A2_KB + (ks & 3) * 32) >> 4), bd + (uint64_t)((128 * 128 + ks * 2048) >> 4), idesc, ks == 0 ? 0u : 1u);
        umma_f16_scale_d11(tm, ad, bd, idesc);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks)
          umma_f16(tm, ad + (uint64_t)(((ks >> 2) * A2_KB + (ks & 3) * 32) >> 4), bd + (uint64_t)((ks * 2048) >> 4), idesc, 1u);
        umma_commit(d2_full(d2));
        umma_commit(t_empty(tg));
      }
      __syncwarp();
      if (++d2 == D2N) { d2 = 0; d2ph ^= 1u; }
      if (++tg == TJ) { tg = 0; tph ^= 1u; }
    }
  } else if (warp >= EPI2_WARP0 && warp < EPI2_WARP0 + 8) {
    // ===================== second epilogue: D2 -> * 64 norm -> fp16 -> fp16 store / INT8 codes -> HBM
    const int set = (warp - EPI2_WARP0) >> 2, quad = warp & 3;
    const int l = quad * 32 + lane;
    const int s = l / A, kp = l % A;
    const bool want_q = a.q != nullptr, want_y = a.y != nullptr;
    QScale q_text = {}, q_vis = {};
    if (want_q) {
      q_text = make_qscale(__ldg(a.scales));
      q_vis = make_qscale(__ldg(a.scales + 1));
    }
    const bool split = (a.flags & MQ_FLAG_SPLIT) != 0;
    const bool q32 = (reinterpret_cast<uintptr_t>(a.q) & 31u) == 0, y32 = (reinterpret_cast<uintptr_t>(a.y) & 31u) == 0;
    const bool rescale = norm2 != 1.0f;
    const uint64_t n2 = f2_pack(norm2, norm2);
    int d2 = 0;
    uint32_t d2ph = 0;
    for (int g = 0; g < my_jobs; ++g) {
      const int64_t tok = (int64_t)((int)blockIdx.x + g * (int)gridDim.x) * JT + set * S + s;
      const bool have = tok < a.M;
      const bool vis = have && a.modality != nullptr && __ldg(a.modality + tok) != 0;
      QScale qs;  // field-wise select keeps it in registers
      qs.s = vis ? q_vis.s : q_text.s;
      qs.inv_lo = vis ? q_vis.inv_lo : q_text.inv_lo;
      qs.inv_hi = vis ? q_vis.inv_hi : q_text.inv_hi;
      qs.lim_lo = vis ? q_vis.lim_lo : q_text.lim_lo;
      qs.lim_hi = vis ? q_vis.lim_hi : q_text.lim_hi;
      qs.narrow = vis ? q_vis.narrow : q_text.narrow;
      if (have && want_q && kp == 0 && a.srow_out != nullptr) a.srow_out[tok] = qs.s;
      mbar_wait_sleep(d2_full(d2), d2ph);
      tc_fence_after();
      const bool any = __any_sync(0xffffffffu, have);
      uint32_t va[32], vb[32];
      if (any) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D2_COL0 + d2 * 128 + set * 64);
        tmem_ld32_nowait(taddr, va);
        tmem_ld32_nowait(taddr + 32, vb);
        tmem_wait_ld();
      }
      // (the slot is handed back after the quantiser and the stores, see mq_hadamard_tc3.cu)
      if (any) {
        int8_t* qrow = want_q ? a.q + tok * (int64_t)a.n + (int64_t)kp * 64 : nullptr;
        __half* yrow = want_y ? a.y + tok * (int64_t)a.n + (int64_t)kp * 64 : nullptr;
        if (rescale) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float lo, hi;
            f2_unpack(f2_mul(f2_pack(__uint_as_float(va[2 * e]), __uint_as_float(va[2 * e + 1])), n2), lo, hi);
            va[2 * e] = __float_as_uint(lo);
            va[2 * e + 1] = __float_as_uint(hi);
            f2_unpack(f2_mul(f2_pack(__uint_as_float(vb[2 * e]), __uint_as_float(vb[2 * e + 1])), n2), lo, hi);
            vb[2 * e] = __float_as_uint(lo);
            vb[2 * e + 1] = __float_as_uint(hi);
          }
        }
        const bool first = split && kp == 0;
        // (the test for the division-free quantiser is hoisted out of the element loops, see mq_hadamard_tc3.cu)
        if (qs.narrow) {
          epi2_emit<true>(va, qrow, yrow, 0, qs, have, first, q32, y32, a.x0_out, tok);
          epi2_emit<true>(vb, qrow, yrow, 32, qs, have, false, q32, y32, a.x0_out, tok);
        } else {
          epi2_emit<false>(va, qrow, yrow, 0, qs, have, first, q32, y32, a.x0_out, tok);
          epi2_emit<false>(vb, qrow, yrow, 32, qs, have, false, q32, y32, a.x0_out, tok);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(d2));
      if (++d2 == D2N) { d2 = 0; d2ph ^= 1u; }
    }
  } else if (warp >= EPI1_WARP0 && warp < EPI1_WARP0 + 4) {
    // ===================== first epilogue: D1 -> hi / lo' rows of T ===================================
    const int quad = warp & 3;
    const int l = quad * 32 + lane;
    const uint32_t sw = (uint32_t)(l & 7);
    int d1 = 0, tg = 0;
    uint32_t d1ph = 0, tph = 0;
    for (int g = 0; g < my_jobs; ++g) {
      mbar_wait_sleep(t_empty(tg), tph ^ 1u);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        mbar_wait_sleep(d1_full(d1), d1ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(d1 * 64);
        const uint32_t hrow = sT + (uint32_t)(tg * T_JOB + p * T_BLK) + (uint32_t)l * 128u;
        const uint32_t lrow = hrow + 128u * 128u;
        uint32_t va[32], vb[32];
        tmem_ld32_nowait(taddr, va);
        tmem_ld32_nowait(taddr + 32, vb);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty(d1));  // the tile is in registers: MMA1 may refill the slot
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t(&v)[32] = half ? vb : va;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t h[4], w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pair(v[cc * 8 + 2 * e], v[cc * 8 + 2 * e + 1], h[e], w[e]);
            const uint32_t off = (((uint32_t)(half * 4 + cc) ^ sw) << 4);
            st_shared_v4(hrow + off, h[0], h[1], h[2], h[3]);
            st_shared_v4(lrow + off, w[0], w[1], w[2], w[3]);
          }
        }
        fence_async_smem();  // T is read by the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_full(tg));
        if (++d1 == D1N) { d1 = 0; d1ph ^= 1u; }
      }
      if (++tg == TJ) { tg = 0; tph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int A>
static int launch(const HadArgs& a, cudaStream_t st) {
  constexpr int S = 128 / A;
  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  // [M][n_in / 64][64]: rows k >= n_in / 64 (the zero padding) and tokens >= M are out of bounds -> zero filled
  CUtensorMap tmap;
  const cuuint32_t estride[3] = {1, 1, 1};
  const cuuint64_t gdim[3] = {64, (cuuint64_t)(a.n_in / 64), (cuuint64_t)a.M};
  const cuuint64_t gstride[2] = {128, (cuuint64_t)a.ldx * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)A, (cuuint32_t)S};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.x, gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  const int num_jobs = (int)((a.M + 2 * S - 1) / (2 * S));
  const int grid = num_jobs < num_sms() ? num_jobs : num_sms();
  MQ_DYN_SMEM(hadp2_tc_kernel<A>, Lay<A>::SM_ALLOC);
  hadp2_tc_kernel<A><<<grid, NUM_THREADS, Lay<A>::SM_ALLOC, st>>>(tmap, a, a.norm * 64.0f, num_jobs);
  return MQ_OK;
}

}  // namespace hadp2

// had_k == 1, n = 1024 ... 8192, whole 64-element rows of input.  Default for n >= 2048 with enough
// tokens to fill the pipeline of every SM (measured at M = 16384, profiles/README.md: 8192 0.49 -> 0.68
// of HBM, 4096 0.58 -> 0.60, 2048 0.44 -> 0.47, 1024 0.42 -> 0.38: seven jobs per SM do not amortise
// the pipeline fill, the warp-per-token mma.sync kernel keeps that size and the small batches).
// mq_debug_force_hadamard_kernel(3) keeps the mma.sync kernel, (4) takes this one for every M and n.
bool hadp2_tc_supported(const HadArgs& a) {
  if (a.had_k != 1) return false;
  if (a.n != 1024 && a.n != 2048 && a.n != 4096 && a.n != 8192) return false;
  if (a.q == nullptr && a.y == nullptr) return false;
  if (a.n_in % 64 != 0 || a.ldx % 8 != 0 || a.M >= (1ll << 30)) return false;
  const int f = forced_hadamard_kernel();
  if (f == 4) return true;
  return f == 0 && a.M >= 1024 && a.n >= 2048;
}

int launch_hadp2_tc(const HadArgs& a, cudaStream_t st) {
  switch (a.n) {
    case 1024: return hadp2::launch<16>(a, st);
    case 2048: return hadp2::launch<32>(a, st);
    case 4096: return hadp2::launch<64>(a, st);
    default: return hadp2::launch<128>(a, st);
  }
}

}  // namespace mq
