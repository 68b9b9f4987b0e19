// Online Hadamard (had_k > 1, n / had_k in {64, 128}) + static INT8 quantisation, third generation:
// BOTH Kronecker factors on tcgen05, hadK as the A operand.
//
//   y[tok, k'*P + j'] = sum_k hadK[k',k] * fp16( norm * sum_j x[tok, k*P + j] * H_P[j,j'] )
//   (matmul_hadU_cuda, fake_quant/hadamard_utils.py:115-128: FHT over the last dim of the
//    [-1, K, n/K] view, fp16 rounding, hadK @ ., fp16 rounding), then UniformQuantizer.quant.
//
// Why (profiles/README.md, round 2): the second generation (mq_hadamard_tc2.cu) spends ~22
// instructions per element -- a 64-point butterfly in registers, and an epilogue whose thread owns
// ONE j' and 16 k', i.e. bytes of 16 different output rows that have to be transposed through
// shuffles -- and is issue bound at 45 % of the HBM roofline.  The premise it was built on (TMEM
// drains at 64 B/clk per SM) is wrong: tools/experiments/tmem_rates.cu measures 64-90 B/clk PER
// WARP, 480-800 B/clk per SM with 8-16 warps.  So both factors go to the tensor pipe and the
// CUDA cores only round, quantise and store (~7 instructions per element):
//
//   TMA    tile = XROWS rows (k) of one token, 128 B (64 j) each per sub-block, SWIZZLE_128B; rows
//          past n_in / P are out of bounds = zero filled (the reference's zero padding)
//   MMA1   D1[k, j'] = X[k, j] . H64[j, j']          M = 128 rows, N = 64, K = 64 (exact products,
//          fp32 accumulation); P = 128: once per 64-wide sub-block, the H2 stage is one add per
//          element in the first epilogue
//   epi-1  thread = one row k (TMEM lane): D1 -> * norm -> fp16 (first rounding point) -> one 128 B
//          row of T, the MN-major B operand of MMA2
//   MMA2   D2[k', j'] = hadK[k', k] . T[k, j']        M = 128 (rows 0..127) and, for had_k > 128,
//          a second tile of rows Kpad-128 .. Kpad-1 (its first lanes recompute rows the first
//          tile already has and are ignored), N = P, K = Kpad
//   epi-2  thread = one output row k': P consecutive j' -> fp16 (second rounding point) -> fp16
//          store and/or the exact INT8 quantiser with the token's MSQ scale -> 256-bit stores
//          (a warp writes 32 rows x P bytes of one token: contiguous)
//
// Persistent, one CTA per SM, 512 threads: warp 0 TMA, warp 1 MMA1 issuer (+ TMEM allocation),
// warp 2 MMA2 issuer, warps 4-7 first epilogue (one per TMEM lane quadrant), warps 8-15 second
// epilogue (two sets that alternate jobs).  TMEM: columns 0..255 ring of D1 tiles, 256..511 ring
// of D2 jobs (job = token x M-tile).  HBM traffic per token: 2*n_in read + n (codes) [+ 2*n fp16].
#include <cuda.h>

#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace hadtc3 {
using namespace mq::tc;

constexpr int NUM_THREADS = 512;
constexpr int EPI1_WARP0 = 4, EPI2_WARP0 = 8;
constexpr int TMEM_COLS = 512;
constexpr int D2_COL0 = 256;
constexpr int MAX_XS = 8, MAX_TB = 4;

struct Cfg {
  int had_k, Kpad, kb2;  // Kpad = round_up(had_k, 16); kb2 = ceil(Kpad / 64) k-blocks of hadK
  int xrows;             // rows per X slot = per D1 tile
  int tpt;               // tiles per token
  int mt;                // M-tiles of MMA2
  int xs, tb;            // ring depths: X slots, T buffers (tokens)
  int x_bytes, t_bytes;
  int a2_off, h_off, x_off, t_off, bar_off, total;
};

struct Params {
  HadArgs a;
  Cfg c;
};

__device__ __forceinline__ void tmem_ld32_at(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32_nowait(taddr, r); }

__device__ __forceinline__ void st_shared_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32 bytes per lane: one 256-bit store when the destination allows it (al32), else two 128-bit ones
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&w)[8], bool al32) {
  if (al32) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  } else {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
__device__ __forceinline__ uint32_t f16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// NS = P / 64 sub-blocks per row
template <int NS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hadk_tc3_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params prm) {
  constexpr int P = 64 * NS;
  constexpr int D1N = 4 / NS;  // D1 tiles in TMEM (P columns each)
  constexpr int D2N = 4 / NS;  // D2 jobs in TMEM
  const HadArgs& a = prm.a;
  const Cfg& c = prm.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA2 = base + c.a2_off, sH = base + c.h_off, sX = base + c.x_off, sT = base + c.t_off;
  const uint32_t bar0 = base + c.bar_off;
  // barrier map: x_full[8] x_empty[8] d1_full[4] d1_empty[4] t_full[4] t_empty[4] d2_full[4] d2_empty[4]
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (MAX_XS + s); };
  auto d1_full = [&](int s) { return bar0 + 8u * (2 * MAX_XS + s); };
  auto d1_empty = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 4 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 8 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 8 + MAX_TB + s); };
  auto d2_full = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 8 + 2 * MAX_TB + s); };
  auto d2_empty = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 12 + 2 * MAX_TB + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + c.bar_off + 8 * (2 * MAX_XS + 16 + 2 * MAX_TB));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int my_tokens = blockIdx.x < a.M ? (int)((a.M - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  // ---- one-time setup: barriers, TMEM, hadK (A of MMA2), H64 (B of MMA1) ------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < MAX_XS; ++s) {
      mbar_init(x_full(s), 1);
      mbar_init(x_empty(s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), 4);  // the four first-epilogue warps
      mbar_init(d2_full(s), 1);
      mbar_init(d2_empty(s), 4);  // the four warps of one second-epilogue set
    }
    for (int s = 0; s < MAX_TB; ++s) {
      mbar_init(t_full(s), 4 * c.tpt);  // every first-epilogue warp, once per tile of the token
      mbar_init(t_empty(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  // hadK as [k' rows][k] K-major SWIZZLE_128B, zero padded to Kpad x (kb2*64): k-block kb at kb * Kpad * 128
  for (int id = threadIdx.x; id < c.Kpad * c.kb2 * 8; id += NUM_THREADS) {
    const int row = id / (c.kb2 * 8), rem = id % (c.kb2 * 8), kb = rem / 8, ch = rem % 8;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = kb * 64 + ch * 8 + 2 * e;
      uint32_t lo = 0, hi = 0;
      if (row < c.had_k) {
        if (k0 < c.had_k) lo = __half_as_ushort(a.hadk[row * c.had_k + k0]);
        if (k0 + 1 < c.had_k) hi = __half_as_ushort(a.hadk[row * c.had_k + k0 + 1]);
      }
      w[e] = lo | (hi << 16);
    }
    st_shared_v4(sA2 + kb * (c.Kpad * 128) + row * 128 + ((uint32_t)(ch ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
  }
  // H64 (symmetric): row j', 64 k = 128 B, same swizzle
  {
    const int row = threadIdx.x >> 3, ch = threadIdx.x & 7;  // 512 threads = 64 rows x 8 chunks
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
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: one tile (NS boxes of xrows x 128 B) per step =============
    if (elect_one()) {
      int xs = 0;
      uint32_t xph = 0;
      for (int lt = 0; lt < my_tokens; ++lt) {
        const int tok = (int)(blockIdx.x + (int64_t)lt * gridDim.x);
        for (int i = 0; i < c.tpt; ++i) {
          mbar_wait(x_empty(xs), xph ^ 1u);
          mbar_arrive_expect_tx(x_full(xs), (uint32_t)c.x_bytes);
#pragma unroll
          for (int sub = 0; sub < NS; ++sub)
            tma_load_3d(sX + xs * c.x_bytes + sub * (c.xrows * 128), &tmap_x, 64 * sub, i * c.xrows, tok, x_full(xs));
          if (++xs == c.xs) { xs = 0; xph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA1 issuer: D1 = X . H64 per sub-block ================================
    // F16 x F16 -> F32, A K-major, B K-major, M = 128, N = 64
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int xs = 0, d1 = 0;
    uint32_t xph = 0, d1ph = 0;
    const int total_tiles = my_tokens * c.tpt;
    for (int g = 0; g < total_tiles; ++g) {
      mbar_wait(x_full(xs), xph);
      mbar_wait(d1_empty(d1), d1ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) {
          const uint32_t xa = sX + xs * c.x_bytes + sub * (c.xrows * 128);
          const uint32_t tm = tmem_base + (uint32_t)(d1 * P + sub * 64);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tm, make_sw128_desc(xa + ks * 32, 0, 1024), make_sw128_desc(sH + ks * 32, 0, 1024), idesc, (uint32_t)(ks != 0));
        }
        umma_commit(x_empty(xs));
        umma_commit(d1_full(d1));
      }
      __syncwarp();
      if (++xs == c.xs) { xs = 0; xph ^= 1u; }
      if (++d1 == D1N) { d1 = 0; d1ph ^= 1u; }
    }
  } else if (warp == 2) {
    // ===================== MMA2 issuer: D2 = hadK . T =============================================
    // A K-major (hadK rows), B MN-major (T rows = k, 64 j' per 128 B), M = 128, N = P
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo = (uint32_t)c.Kpad * 128u;  // between the 64-wide sub-blocks of T
    int tb = 0, d2 = 0;
    uint32_t tph = 0, d2ph = 0;
    for (int lt = 0; lt < my_tokens; ++lt) {
      mbar_wait(t_full(tb), tph);
      for (int mt = 0; mt < c.mt; ++mt) {
        mbar_wait(d2_empty(d2), d2ph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t arow = sA2 + (uint32_t)(mt == 0 ? 0 : (c.Kpad - 128)) * 128u;
          const uint32_t tbase = sT + tb * c.t_bytes;
          const uint32_t tm = tmem_base + (uint32_t)(D2_COL0 + d2 * P);
          for (int ks = 0; ks < c.Kpad / 16; ++ks) {
            const uint32_t aoff = (uint32_t)(ks >> 2) * (uint32_t)(c.Kpad * 128) + (uint32_t)(ks & 3) * 32u;
            umma_f16(tm, make_sw128_desc(arow + aoff, 0, 1024), make_sw128_desc(tbase + (uint32_t)ks * 2048u, lbo, 1024), idesc,
                     (uint32_t)(ks != 0));
          }
          umma_commit(d2_full(d2));
          if (mt == c.mt - 1) umma_commit(t_empty(tb));
        }
        __syncwarp();
        if (++d2 == D2N) { d2 = 0; d2ph ^= 1u; }
      }
      if (++tb == c.tb) { tb = 0; tph ^= 1u; }
    }
  } else if (warp >= EPI2_WARP0) {
    // ===================== second epilogue: D2 -> fp16 -> fp16 store / INT8 codes -> HBM ==========
    const int set = (warp - EPI2_WARP0) >> 2, quad = warp & 3;
    const int l = quad * 32 + lane;
    const bool want_q = a.q != nullptr, want_y = a.y != nullptr;
    QScale q_text = {}, q_vis = {};
    if (want_q) {
      q_text = make_qscale(__ldg(a.scales));
      q_vis = make_qscale(__ldg(a.scales + 1));
    }
    const bool split = (a.flags & MQ_FLAG_SPLIT) != 0;
    const bool q32 = (reinterpret_cast<uintptr_t>(a.q) & 31u) == 0, y32 = (reinterpret_cast<uintptr_t>(a.y) & 31u) == 0;
    const int total_jobs = my_tokens * c.mt;
    const int lane0_mt1 = 256 - c.Kpad;  // first lane of the second M-tile that holds a new row
    for (int j = set; j < total_jobs; j += 2) {
      const int lt = c.mt == 2 ? (j >> 1) : j, mt = c.mt == 2 ? (j & 1) : 0;
      const int d2 = j % D2N;
      const uint32_t d2ph = (uint32_t)(j / D2N) & 1u;
      const int64_t tok = blockIdx.x + (int64_t)lt * gridDim.x;
      const int kp = mt == 0 ? l : (c.Kpad - 128 + l);
      const bool lane_ok = (mt == 0 ? true : (l >= lane0_mt1)) && kp < c.had_k;
      const bool warp_ok = mt == 0 ? (quad * 32 < c.had_k) : (quad * 32 + 31 >= lane0_mt1 && c.Kpad - 128 + quad * 32 < c.had_k);
      const bool vis = a.modality != nullptr && __ldg(a.modality + tok) != 0;
      QScale qs;  // field-wise select keeps it in registers
      qs.s = vis ? q_vis.s : q_text.s;
      qs.inv_lo = vis ? q_vis.inv_lo : q_text.inv_lo;
      qs.inv_hi = vis ? q_vis.inv_hi : q_text.inv_hi;
      qs.lim_lo = vis ? q_vis.lim_lo : q_text.lim_lo;
      qs.lim_hi = vis ? q_vis.lim_hi : q_text.lim_hi;
      qs.narrow = vis ? q_vis.narrow : q_text.narrow;
      if (want_q && mt == 0 && l == 0 && a.srow_out != nullptr) a.srow_out[tok] = qs.s;
      mbar_wait(d2_full(d2), d2ph);
      tc_fence_after();
      if (warp_ok) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D2_COL0 + d2 * P);
        int8_t* qrow = want_q ? a.q + tok * (int64_t)a.n + (int64_t)kp * P : nullptr;
        __half* yrow = want_y ? a.y + tok * (int64_t)a.n + (int64_t)kp * P : nullptr;
        uint32_t va[32], vb[32];
        tmem_ld32_at(taddr, va);
#pragma unroll
        for (int cc = 0; cc < P / 32; ++cc) {
          tmem_wait_ld();
          uint32_t(&cur)[32] = (cc & 1) ? vb : va;
          uint32_t(&nxt)[32] = (cc & 1) ? va : vb;
          if (cc + 1 < P / 32) tmem_ld32_at(taddr + (cc + 1) * 32, nxt);
          uint32_t h[16];  // second rounding point (hadK @ . -> fp16)
#pragma unroll
          for (int e = 0; e < 16; ++e) h[e] = f16x2_rn(__uint_as_float(cur[2 * e]), __uint_as_float(cur[2 * e + 1]));
          if (lane_ok) {
            if (want_y) {
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const uint32_t w[8] = {h[8 * v], h[8 * v + 1], h[8 * v + 2], h[8 * v + 3], h[8 * v + 4], h[8 * v + 5], h[8 * v + 6], h[8 * v + 7]};
                st_global_v8(yrow + cc * 32 + v * 16, w, y32);
              }
            }
            if (want_q) {
              uint32_t w[8];
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                const uint2 r = quant8(make_uint4(h[4 * v], h[4 * v + 1], h[4 * v + 2], h[4 * v + 3]), qs);
                w[2 * v] = r.x;
                w[2 * v + 1] = r.y;
              }
              if (split && cc == 0 && kp == 0) {
                a.x0_out[tok] = __ushort_as_half((unsigned short)(h[0] & 0xffffu));
                w[0] &= 0xffffff00u;
              }
              st_global_v8(qrow + cc * 32, w, q32);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(d2));
    }
  } else if (warp >= EPI1_WARP0) {
    // ===================== first epilogue: D1 -> * norm -> fp16 -> T ===============================
    const int quad = warp & 3;
    const int l = quad * 32 + lane;
    const uint64_t norm2 = f2_pack(a.norm, a.norm);
    int d1 = 0, tb = 0;
    uint32_t d1ph = 0, tph = 0;
    for (int lt = 0; lt < my_tokens; ++lt) {
      for (int i = 0; i < c.tpt; ++i) {
        const int rows_here = min(c.xrows, c.Kpad - i * c.xrows);
        if (i == 0) mbar_wait(t_empty(tb), tph ^ 1u);
        mbar_wait(d1_full(d1), d1ph);
        tc_fence_after();
        if (quad * 32 < rows_here) {
          const int k = i * c.xrows + l;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(d1 * P);
          const uint32_t trow = sT + tb * c.t_bytes + (uint32_t)k * 128u;
          const uint32_t sw = (uint32_t)(k & 7);
          const bool row_ok = l < rows_here;
          if (NS == 1) {
            uint32_t va[32], vb[32];
            tmem_ld32_at(taddr, va);
            tmem_ld32_at(taddr + 32, vb);
            tmem_wait_ld();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t(&v)[32] = half ? vb : va;
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                uint32_t h[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float lo, hi;
                  f2_unpack(f2_mul(f2_pack(__uint_as_float(v[cc * 8 + 2 * e]), __uint_as_float(v[cc * 8 + 2 * e + 1])), norm2), lo, hi);
                  h[e] = f16x2_rn(lo, hi);
                }
                if (row_ok) st_shared_v4(trow + (((uint32_t)(half * 4 + cc) ^ sw) << 4), h[0], h[1], h[2], h[3]);
              }
            }
          } else {
            // P = 128: columns j' (sub-block 0) and 64 + j' (sub-block 1) of a row combine into the
            // H2 stage: (u, v) -> (u + v, u - v), then both halves are rounded and stored
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
              uint32_t va[32], vb[32];
              tmem_ld32_at(taddr + q4 * 32, va);
              tmem_ld32_at(taddr + 64 + q4 * 32, vb);
              tmem_wait_ld();
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                uint32_t hs[4], hd[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint64_t u = f2_pack(__uint_as_float(va[cc * 8 + 2 * e]), __uint_as_float(va[cc * 8 + 2 * e + 1]));
                  const uint64_t w = f2_pack(__uint_as_float(vb[cc * 8 + 2 * e]), __uint_as_float(vb[cc * 8 + 2 * e + 1]));
                  uint64_t df;
                  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(df) : "l"(u), "l"(w));
                  float lo, hi;
                  f2_unpack(f2_mul(f2_add(u, w), norm2), lo, hi);
                  hs[e] = f16x2_rn(lo, hi);
                  f2_unpack(f2_mul(df, norm2), lo, hi);
                  hd[e] = f16x2_rn(lo, hi);
                }
                if (row_ok) {
                  const uint32_t off = (((uint32_t)(q4 * 4 + cc) ^ sw) << 4);
                  st_shared_v4(trow + off, hs[0], hs[1], hs[2], hs[3]);
                  st_shared_v4(trow + (uint32_t)c.Kpad * 128u + off, hd[0], hd[1], hd[2], hd[3]);
                }
              }
            }
          }
          fence_async_smem();  // T is read by the tensor core (async proxy)
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t_full(tb));
          mbar_arrive(d1_empty(d1));
        }
        if (++d1 == D1N) { d1 = 0; d1ph ^= 1u; }
      }
      if (++tb == c.tb) { tb = 0; tph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int NS>
static int launch(const HadArgs& a, cudaStream_t st) {
  constexpr int P = 64 * NS;
  Params prm;
  prm.a = a;
  Cfg& c = prm.c;
  c.had_k = a.had_k;
  c.Kpad = round_up(a.had_k, 16);
  c.kb2 = (c.Kpad + 63) / 64;
  c.mt = c.Kpad > 128 ? 2 : 1;
  c.t_bytes = NS * c.Kpad * 128;
  const int a2 = round_up(c.kb2 * c.Kpad * 128, 1024);
  const int hb = 64 * 128;
  const int bars = 8 * (2 * MAX_XS + 16 + 2 * MAX_TB) + 16;
  const int budget = 227 * 1024 - 1024 - round_up(bars, 128);
  // T: at least two tokens (one being filled while the other is consumed); X: what is left, in
  // 128-row tiles if four of them fit, else 64-row tiles
  c.tb = 3;
  c.xrows = 128;
  auto fits = [&](int tb, int xrows, int xs) { return a2 + hb + tb * c.t_bytes + xs * NS * xrows * 128 <= budget; };
  if (!fits(c.tb, c.xrows, 4)) c.tb = 2;
  if (!fits(c.tb, c.xrows, 4)) c.xrows = 64;
  if (!fits(c.tb, c.xrows, 3)) return MQ_ERR_UNSUPPORTED;
  c.x_bytes = NS * c.xrows * 128;
  c.xs = (budget - a2 - hb - c.tb * c.t_bytes) / c.x_bytes;
  if (c.xs > MAX_XS) c.xs = MAX_XS;
  c.tpt = (c.Kpad + c.xrows - 1) / c.xrows;
  c.a2_off = 0;
  c.h_off = a2;
  c.x_off = c.h_off + hb;
  c.t_off = c.x_off + c.xs * c.x_bytes;
  c.bar_off = c.t_off + c.tb * c.t_bytes;
  c.total = c.bar_off + bars;

  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  // [M][n_in / P][P]: rows k >= n_in / P (the zero padding) are out of bounds -> zero filled by TMA
  CUtensorMap tmap;
  const cuuint32_t estride[3] = {1, 1, 1};
  const cuuint64_t gdim[3] = {(cuuint64_t)P, (cuuint64_t)(a.n_in / P), (cuuint64_t)a.M};
  const cuuint64_t gstride[2] = {(cuuint64_t)P * 2, (cuuint64_t)a.ldx * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)c.xrows, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.x, gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  const int smem = c.total + 1024;
  MQ_DYN_SMEM(hadk_tc3_kernel<NS>, smem);
  const int grid = (int)(a.M < num_sms() ? a.M : num_sms());
  hadk_tc3_kernel<NS><<<grid, NUM_THREADS, smem, st>>>(tmap, prm);
  return MQ_OK;
}

}  // namespace hadtc3

// had_k > 1, n / had_k in {64, 128}, fp16 rounding points of matmul_hadU_cuda: codes and/or fp16 output
bool hadk_tc3_supported(const HadArgs& a) {
  const int P = a.n / a.had_k;
  if (a.had_k <= 1 || a.had_k > 256 || (P != 64 && P != 128)) return false;
  if (a.q == nullptr && a.y == nullptr) return false;
  if (a.flags & MQ_FLAG_FP32_HAD) return false;
  if (a.n_in % P != 0 || a.M >= (1ll << 31)) return false;
  const int f = forced_hadamard_kernel();
  return f != 1 && f != 2;
}

int launch_hadk_tc3(const HadArgs& a, cudaStream_t st) {
  return a.n / a.had_k == 64 ? hadtc3::launch<1>(a, st) : hadtc3::launch<2>(a, st);
}

}  // namespace mq
