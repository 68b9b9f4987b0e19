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
// CUDA cores only round, quantise and store (~8 instructions per element):
//
//   TMA    tile = XROWS rows (k) of one token, 128 B (64 j) each per sub-block, SWIZZLE_128B; rows
//          past n_in / P are out of bounds = zero filled (the reference's zero padding)
//   MMA1   D1[k, j'] = X[k, j] . H64[j, j']          M = 128 rows, N = 64, K = 64 (exact products,
//          fp32 accumulation); P = 128: once per 64-wide sub-block, the H2 stage is one add per
//          element in the first epilogue
//   epi-1  thread = one row k (TMEM lane): D1 -> * norm -> fp16 (first rounding point) -> one 128 B
//          row of T, the MN-major B operand of MMA2
//   MMA2   D2[k', (tok, j')] = hadK[k', k] . T[k, (tok, j')]   M = 128 (rows 0..127) and, for
//          had_k > 128, a second tile of rows Kpad-128 .. Kpad-1 (its first lanes recompute rows
//          the first tile already has and are ignored), K = Kpad, N = 128: G = 128 / P tokens per
//          job.  (tools/experiments/mma_rates.cu: an MMA whose A operand changes costs >= 80 clk
//          whatever N <= 128 is, so N = 64 would halve the rate.)
//   epi-2  thread = one output row k': P consecutive j' of a token -> fp16 (second rounding point)
//          -> fp16 store and/or the exact INT8 quantiser with the token's MSQ scale -> 256-bit
//          stores (a warp writes 32 rows x P bytes of one token: contiguous)
//
// Persistent, one CTA per SM, 512 threads: warps 0-3 first epilogue (one per TMEM lane quadrant), warps
// 4-11 second epilogue (two sets; all twelve first stage hadK while the pipeline fills), warp 12
// TMA, warp 13 MMA1 issuer (+ TMEM allocation), warp 14 MMA2 issuer.  TMEM: columns 0..255 ring
// of D1 tiles, 256..511 two D2 jobs (job = token group x M-tile).
// HBM traffic per token: 2*n_in read + n (codes) [+ 2*n fp16].
#include <cuda.h>

#include <cstdlib>

#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"
#include "mq_hadamard_epi.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace hadtc3 {
using namespace mq::tc;
using namespace mq::hadepi;

constexpr int NUM_THREADS = 512;
// The warp scheduler favours HIGH warp ids (B300_MICROARCH.md, arbiter priority): the issuers of
// TMA / MMA get the highest ids, then the second epilogue (the stage every output passes through; 0.5 %
// over the opposite order in a same-box A/B), the light first epilogue the lowest.
// (With the issuers at ids 0-2 the MMAs of a job were issued at half speed whenever the second
// epilogue of the previous job was running on the same scheduler.)
constexpr int EPI1_WARP0 = 0, EPI2_WARP0 = 4, W_TMA = 12, W_MMA1 = 13, W_MMA2 = 14;
constexpr int TMEM_COLS = 512;
constexpr int MAX_XS = 8, MAX_TG = 4;

struct Cfg {
  int had_k, Kpad, kb2;  // Kpad = round_up(had_k, 16); kb2 = ceil(Kpad / 64) k-blocks of hadK
  int src_k, expand;     // had_k = src_k * expand: the A operand is hadK (x) H_expand (n / src_k in {256, 512})
  int xrows;             // rows per X slot = per D1 tile
  int tpt;               // tiles per token
  int mt;                // M-tiles of MMA2
  int xs, tg;            // ring depths: X slots, T group slots (G tokens each)
  int x_bytes, t_bytes;  // bytes per X slot, per token of T
  int a2_off, h_off, x_off, t_off, bar_off, total;
};

struct Params {
  HadArgs a;
  Cfg c;
};

#ifdef MQ_STAMPS
// clock64 timeline of CTA 0 (tools/had_probe.py --timeline3, build.py --stamps; the stamps themselves
// slow the kernel by a third: read the order of events, not the durations): 16 slots per token, first 96 tokens
__device__ long long g_stamps[96 * 16];
__device__ __forceinline__ void stamp(int lt, int slot) {
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && lt < 96) g_stamps[lt * 16 + slot] = clock64();
}
#else
__device__ __forceinline__ void stamp(int, int) {}
#endif

// The K loop of one MMA2 job.  The issuing thread shares its scheduler with three epilogue warps:
// every instruction between two MMAs costs it an issue slot it has to win (measured: 117 clk per
// MMA with the descriptors rebuilt per k-step, 240 clk while the second epilogue is busy), so the
// descriptors are the base plus compile-time constants (address fields are in 16-byte units):
// k-step ks reads hadK at k-block ks / 4, byte 32 * (ks % 4) of the row, and rows 16 ks .. of T.
template <int NKS>
__device__ __forceinline__ void issue_mma2(uint32_t tm, uint64_t ad, uint64_t bd, uint32_t idesc, int nks_rt = 0) {
  constexpr int KPAD = 16 * NKS;
  if (NKS > 0) {
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks)
      umma_f16(tm, ad + (uint64_t)(((ks >> 2) * KPAD * 128 + (ks & 3) * 32) >> 4), bd + (uint64_t)(ks * 128), idesc, ks == 0 ? 0u : 1u);
  } else {
    const uint64_t kb_inc = (uint64_t)((nks_rt * 16 * 128 - 96) >> 4);
    umma_f16(tm, ad, bd, idesc, 0u);
    for (int ks = 1; ks < nks_rt; ++ks) {
      ad += (ks & 3) == 0 ? kb_inc : 2u;
      bd += 128u;
      umma_f16(tm, ad, bd, idesc, 1u);
    }
  }
}

// NS = P / 64 sub-blocks per row; G = 2 / NS tokens per MMA2 job (N = 128 columns); D1N = D1 tiles in TMEM
template <int NS, int D1N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hadk_tc3_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params prm) {
  constexpr int P = 64 * NS;
  constexpr int G = 2 / NS;
  // TMEM: D1 tiles (P columns each) first, then the D2 jobs (128 columns each).  The first epilogue
  // is never behind (two D1 tiles suffice); the MMA2 <-> second-epilogue hand-over is the
  // latency-critical loop, so it gets the third slot where the columns allow (P = 64).
  constexpr int D2_COL0 = D1N * P;
  constexpr int D2N = (TMEM_COLS - D2_COL0) / 128;
  const HadArgs& a = prm.a;
  const Cfg& c = prm.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA2 = base + c.a2_off, sH = base + c.h_off, sX = base + c.x_off, sT = base + c.t_off;
  const uint32_t bar0 = base + c.bar_off;
  // barrier map: x_full[8] x_empty[8] d1_full[4] d1_empty[4] t_full[4] t_empty[4] d2_full[4] d2_empty[4] a2_ready
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (MAX_XS + s); };
  auto d1_full = [&](int s) { return bar0 + 8u * (2 * MAX_XS + s); };
  auto d1_empty = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 4 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 8 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 8 + MAX_TG + s); };
  auto d2_full = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 8 + 2 * MAX_TG + s); };
  auto d2_empty = [&](int s) { return bar0 + 8u * (2 * MAX_XS + 12 + 2 * MAX_TG + s); };
  const uint32_t a2_ready = bar0 + 8u * (2 * MAX_XS + 16 + 2 * MAX_TG);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + c.bar_off + 8 * (2 * MAX_XS + 17 + 2 * MAX_TG));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int my_tokens = blockIdx.x < a.M ? (int)((a.M - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int my_groups = (my_tokens + G - 1) / G;

  // ---- one-time setup: barriers, TMEM, H64 (B of MMA1); hadK is staged by the epi-2 warps --------
  if (warp == W_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < MAX_XS; ++s) {
      mbar_init(x_full(s), 1);
      mbar_init(x_empty(s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), 4);  // the four first-epilogue warps
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(d2_full(s), 1);
      mbar_init(d2_empty(s), 8);  // the eight second-epilogue warps
    }
    for (int s = 0; s < MAX_TG; ++s) {
      mbar_init(t_full(s), 4 * c.tpt * G);  // every first-epilogue warp, once per tile of the group's tokens
      mbar_init(t_empty(s), 1);
    }
    mbar_init(a2_ready, 12);  // the twelve epilogue warps stage hadK
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first ring of X tiles is requested before the matrices are staged (cold HBM latency)
    const int pre = min(c.xs, my_tokens * c.tpt);
    for (int t = 0; t < pre; ++t) {
      const int lt = t / c.tpt, i = t - lt * c.tpt;
      const int tok = (int)(blockIdx.x + (int64_t)lt * gridDim.x);
      mbar_arrive_expect_tx(x_full(t), (uint32_t)c.x_bytes);
#pragma unroll
      for (int sub = 0; sub < NS; ++sub)
        tma_load_3d(sX + t * c.x_bytes + sub * (c.xrows * 128), &tmap_x, 64 * sub, i * c.xrows, tok, x_full(t));
    }
  }
  if (warp == W_MMA1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  // H64 (symmetric): row j', 64 k = 128 B, K-major SWIZZLE_128B
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
  // hadK (the A operand of MMA2) as [k' rows][k] K-major SWIZZLE_128B, zero padded to Kpad x (kb2*64),
  // k-block kb at kb * Kpad * 128, staged by the twelve epilogue warps.  Before the barrier: the region
  // is zeroed (the padding) and the first STAGE_WORDS 16-byte words per thread of the flat matrix are
  // requested -- a cold miss of every CTA of the grid at once takes longer than the rest of the setup
  // (timeline of the previous version: MMA2 of the first group waited ~9000 cycles for this matrix).
  // After it: the words are scattered to their swizzled places; a word is two 4-element halves, each
  // inside one row whenever had_k % 4 == 0.
  constexpr int STAGERS = 12 * 32, STAGE_WORDS = 10;
  const bool stager = warp < 12;  // the twelve epilogue warps
  const bool flat_ok = c.expand == 1 && (c.had_k & 3) == 0 && (reinterpret_cast<uintptr_t>(a.hadk) & 15u) == 0;
  const int nhalf = c.had_k * c.had_k / 4, nword = (nhalf + 1) / 2;
  uint4 pre[STAGE_WORDS];
  auto load_word = [&](int w) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const uint4* src = reinterpret_cast<const uint4*>(a.hadk);
    if (2 * w + 1 < nhalf) v = __ldg(src + w);
    else if (2 * w < nhalf) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(src + w)); v.x = t.x; v.y = t.y; }
    return v;
  };
  if (stager) {
    if (flat_ok) {
#pragma unroll
      for (int u = 0; u < STAGE_WORDS; ++u) pre[u] = load_word((int)threadIdx.x + u * STAGERS);
    }
    for (int i = threadIdx.x; i < c.h_off / 16; i += STAGERS) st_shared_v4(sA2 + (uint32_t)i * 16u, 0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (stager) {
    const int tid = threadIdx.x;
    auto sdst = [&](int row, int k) {
      return sA2 + (uint32_t)(k >> 6) * (uint32_t)(c.Kpad * 128) + (uint32_t)row * 128u +
             ((uint32_t)(((k & 63) >> 3) ^ (row & 7)) << 4) + (uint32_t)(k & 7) * 2u;
    };
    if (flat_ok) {
      // row = e / had_k by multiplication: e < 2^16, had_k <= 256 (exact, see mq_common.cuh FastDiv)
      const uint32_t inv = (uint32_t)((0x100000000ull + (uint32_t)c.had_k - 1) / (uint32_t)c.had_k);
      auto put_word = [&](int w, const uint4& v) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int e = (2 * w + hf) * 4;  // flat element index of this half
          if (e < c.had_k * c.had_k) {
            const int row = (int)__umulhi((uint32_t)e, inv), k = e - row * c.had_k;
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sdst(row, k)), "r"(hf ? v.z : v.x), "r"(hf ? v.w : v.y) : "memory");
          }
        }
      };
#pragma unroll
      for (int u = 0; u < STAGE_WORDS; ++u) put_word(tid + u * STAGERS, pre[u]);
      for (int w0 = tid + STAGE_WORDS * STAGERS; w0 < nword; w0 += STAGE_WORDS * STAGERS) {
#pragma unroll
        for (int u = 0; u < STAGE_WORDS; ++u) pre[u] = load_word(w0 + u * STAGERS);
#pragma unroll
        for (int u = 0; u < STAGE_WORDS; ++u) put_word(w0 + u * STAGERS, pre[u]);
      }
    } else if (c.expand > 1) {
      // composite factor hadK (x) H_E: element (row, k) = hadK[row / E][k / E] * (-1)^popc((row % E) & (k % E))
      // (the source matrix is a few hundred elements: L1 hits)
      const int le = c.expand == 2 ? 1 : c.expand == 4 ? 2 : 3, em = c.expand - 1;
      for (int e = tid; e < c.had_k * c.had_k; e += STAGERS) {
        const int row = e / c.had_k, k = e - row * c.had_k;
        unsigned short v = __half_as_ushort(__ldg(a.hadk + (row >> le) * c.src_k + (k >> le)));
        if (__popc((row & em) & (k & em)) & 1) v ^= (unsigned short)0x8000u;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(sdst(row, k)), "h"(v) : "memory");
      }
    } else {
      for (int e = tid; e < c.had_k * c.had_k; e += STAGERS) {
        const int row = e / c.had_k, k = e - row * c.had_k;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(sdst(row, k)), "h"(__half_as_ushort(__ldg(a.hadk + e))) : "memory");
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(a2_ready);
  }

  if (warp == W_TMA) {
    // ===================== TMA producer: one tile (NS boxes of xrows x 128 B) per step =============
    if (elect_one()) {
      // (tiles 0 .. c.xs - 1 were requested during the setup: the ring has wrapped once)
      int xs = 0;
      uint32_t xph = 1;
      const int total_tiles = my_tokens * c.tpt;
      int lt = c.xs / c.tpt, i = c.xs - lt * c.tpt;
      for (int t = c.xs; t < total_tiles; ++t) {
        const int tok = (int)(blockIdx.x + (int64_t)lt * gridDim.x);
        mbar_wait_sleep(x_empty(xs), xph ^ 1u);
        mbar_arrive_expect_tx(x_full(xs), (uint32_t)c.x_bytes);
#pragma unroll
        for (int sub = 0; sub < NS; ++sub)
          tma_load_3d(sX + xs * c.x_bytes + sub * (c.xrows * 128), &tmap_x, 64 * sub, i * c.xrows, tok, x_full(xs));
        if (i == 0) stamp(lt, 0);
        if (++xs == c.xs) { xs = 0; xph ^= 1u; }
        if (++i == c.tpt) { i = 0; ++lt; }
      }
    }
  } else if (warp == W_MMA1) {
    // ===================== MMA1 issuer: D1 = X . H64 per sub-block ================================
    // F16 x F16 -> F32, A K-major, B K-major, M = 128, N = 64
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int xs = 0, d1 = 0;
    uint32_t xph = 0, d1ph = 0;
    const int total_tiles = my_tokens * c.tpt;
    int ti = 0, lt1 = 0;  // (tile of the token, token: for the probes only)
    for (int g = 0; g < total_tiles; ++g) {
      mbar_wait_sleep(x_full(xs), xph);
      if (ti == 0) stamp(lt1, 1);
      mbar_wait_sleep(d1_empty(d1), d1ph ^ 1u);
      if (ti == 0) stamp(lt1, 2);
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
      if (++ti == c.tpt) { stamp(lt1, 3); ti = 0; ++lt1; }
      if (++xs == c.xs) { xs = 0; xph ^= 1u; }
      if (++d1 == D1N) { d1 = 0; d1ph ^= 1u; }
    }
  } else if (warp == W_MMA2) {
    // ===================== MMA2 issuer: D2 = hadK . T, one job per (token group, M-tile) ==========
    // A K-major (hadK rows), B MN-major (T rows = k, 64 j' per 128 B), M = 128, N = P * tokens
    const uint32_t idesc0 = (1u << 4) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
    // 64-wide column blocks of B: sub-blocks of a token (P = 128) or the tokens of a group (P = 64)
    const uint32_t lbo = NS == 2 ? (uint32_t)c.Kpad * 128u : (uint32_t)c.t_bytes;
    const int nks = c.Kpad / 16;
    int tg = 0, d2 = 0;
    uint32_t tph = 0, d2ph = 0;
    mbar_wait_sleep(a2_ready, 0);
    for (int g = 0; g < my_groups; ++g) {
      const int tig = min(G, my_tokens - g * G);  // tokens in this group
      const uint32_t idesc = idesc0 | ((uint32_t)((P * tig) >> 3) << 17);
      mbar_wait_sleep(t_full(tg), tph);
      stamp(g * G, 4);
      for (int mt = 0; mt < c.mt; ++mt) {
        mbar_wait_sleep(d2_empty(d2), d2ph ^ 1u);
        if (mt == 0) stamp(g * G, 5);
        tc_fence_after();
        if (elect_one()) {
          // second M-tile: rows Kpad-128 .. Kpad-1 (new rows in the LAST lanes) for even groups, rows
          // 128 .. 255 (new rows in the FIRST lanes; what it reads past hadK is finite garbage or
          // somebody's data, and only feeds rows nobody reads) for odd groups: the second epilogue of
          // this job then alternates between the TMEM lane quadrants 2-3 and 0-1, whose warps would
          // otherwise carry two jobs per group against one
          const uint32_t arow = sA2 + (uint32_t)(mt == 0 ? 0 : ((g & 1) ? 128 : c.Kpad - 128)) * 128u;
          const uint32_t tbase = sT + tg * (G * c.t_bytes);
          const uint32_t tm = tmem_base + (uint32_t)(D2_COL0 + d2 * 128);
          const uint64_t ad = make_sw128_desc(arow, 0, 1024), bd = make_sw128_desc(tbase, lbo, 1024);
          if (G == 2) stamp(g * G + 1, 4 + 2 * mt);  // (probes, odd rows: issue of job A / B starts, ends)
          switch (nks) {  // (the common hadK sizes 108 / 140 / 156 / 172 fully unrolled)
            case 7: issue_mma2<7>(tm, ad, bd, idesc); break;
            case 9: issue_mma2<9>(tm, ad, bd, idesc); break;
            case 10: issue_mma2<10>(tm, ad, bd, idesc); break;
            case 11: issue_mma2<11>(tm, ad, bd, idesc); break;
            default: issue_mma2<0>(tm, ad, bd, idesc, nks); break;
          }
          if (G == 2) stamp(g * G + 1, 5 + 2 * mt);
          umma_commit(d2_full(d2));
          if (mt == c.mt - 1) umma_commit(t_empty(tg));
        }
        __syncwarp();
        if (++d2 == D2N) { d2 = 0; d2ph ^= 1u; }
      }
      stamp(g * G, 6);
      if (++tg == c.tg) { tg = 0; tph ^= 1u; }
    }
  } else if (warp >= EPI2_WARP0 && warp < EPI2_WARP0 + 8) {
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
    const int lane0_mt1 = 256 - c.Kpad;  // first lane of the second M-tile that holds a new row
    // Work split inside a job (4 chunks of 32 columns): P = 64 -> set s owns token s of the group;
    // P = 128 -> set s owns chunks s and s + 2 of the token.
    int d2 = 0;
    uint32_t d2ph = 0;
    for (int g = 0; g < my_groups; ++g) {
      const int tig = min(G, my_tokens - g * G);
      const int lt = g * G + (G == 2 ? set : 0);
      const bool have = G == 1 || set < tig;
      const int64_t tok = blockIdx.x + (int64_t)lt * gridDim.x;
      const bool vis = have && a.modality != nullptr && __ldg(a.modality + tok) != 0;
      QScale qs;  // field-wise select keeps it in registers
      qs.s = vis ? q_vis.s : q_text.s;
      qs.inv_lo = vis ? q_vis.inv_lo : q_text.inv_lo;
      qs.inv_hi = vis ? q_vis.inv_hi : q_text.inv_hi;
      qs.lim_lo = vis ? q_vis.lim_lo : q_text.lim_lo;
      qs.lim_hi = vis ? q_vis.lim_hi : q_text.lim_hi;
      qs.narrow = vis ? q_vis.narrow : q_text.narrow;
      if (have && want_q && l == 0 && (G == 2 || set == 0) && a.srow_out != nullptr) a.srow_out[tok] = qs.s;
      for (int mt = 0; mt < c.mt; ++mt) {
        const bool alt = (g & 1) != 0;  // (see the MMA2 issuer)
        const int kp = mt == 0 ? l : alt ? 128 + l : (c.Kpad - 128 + l);
        const bool lane_ok = have && (mt == 0 || alt || l >= lane0_mt1) && kp < c.had_k;
        const bool warp_ok = have && (mt == 0   ? (quad * 32 < c.had_k)
                                      : alt     ? (128 + quad * 32 < c.had_k)
                                                : (quad * 32 + 31 >= lane0_mt1 && c.Kpad - 128 + quad * 32 < c.had_k));
        mbar_wait_sleep(d2_full(d2), d2ph);
        if (warp == EPI2_WARP0 && mt == 0) stamp(lt, 7);
        if (warp == EPI2_WARP0 + 3 && mt == c.mt - 1) stamp(lt, 9);
        tc_fence_after();
        uint32_t va[32], vb[32];
        const int cc0 = G == 2 ? 2 * set : set, cstep = G == 2 ? 1 : 2;  // this set's chunks: cc0, cc0 + cstep
        // (the slot is handed back AFTER the quantiser and the stores: releasing it as soon as the columns
        // are in registers was 7 % SLOWER in a same-box A/B, 130.0 vs 121.8 us at 16384 x 11008)
        if (warp_ok) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D2_COL0 + d2 * 128);
          int8_t* qrow = want_q ? a.q + tok * (int64_t)a.n + (int64_t)kp * P : nullptr;
          __half* yrow = want_y ? a.y + tok * (int64_t)a.n + (int64_t)kp * P : nullptr;
          const int col0 = (G == 2 ? 0 : cc0) * 32, col1 = (G == 2 ? 1 : cc0 + cstep) * 32;  // columns of the token
          const bool first = split && col0 == 0 && kp == 0;
          tmem_ld32_nowait(taddr + cc0 * 32, va);
          tmem_ld32_nowait(taddr + (cc0 + cstep) * 32, vb);
          tmem_wait_ld();
          // (the test for the division-free quantiser is hoisted: inside quant8 it would cut the 64
          // independent elements of a thread into basic blocks of 8 -- measured IPC 0.19 per warp)
          if (qs.narrow) {
            epi2_emit<true>(va, qrow, yrow, col0, qs, lane_ok, first, q32, y32, a.x0_out, tok);
            epi2_emit<true>(vb, qrow, yrow, col1, qs, lane_ok, false, q32, y32, a.x0_out, tok);
          } else {
            epi2_emit<false>(va, qrow, yrow, col0, qs, lane_ok, first, q32, y32, a.x0_out, tok);
            epi2_emit<false>(vb, qrow, yrow, col1, qs, lane_ok, false, q32, y32, a.x0_out, tok);
          }
        }
        if (warp == EPI2_WARP0 && mt == 0) stamp(lt, 8);
        if (warp == EPI2_WARP0 + 3 && mt == c.mt - 1) stamp(lt, 10);
        if (G == 2 && mt == 1 && quad >= 2) stamp(g * G + 1, warp == EPI2_WARP0 + 2 ? 8 : warp == EPI2_WARP0 + 3 ? 9 : warp == EPI2_WARP0 + 6 ? 10 : 15);  // (probes: job B done)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty(d2));
        if (++d2 == D2N) { d2 = 0; d2ph ^= 1u; }
      }
    }
  } else if (warp >= EPI1_WARP0 && warp < EPI1_WARP0 + 4) {
    // ===================== first epilogue: D1 -> * norm -> fp16 -> T ===============================
    const int quad = warp & 3;
    const int l = quad * 32 + lane;
    const uint64_t norm2 = f2_pack(a.norm, a.norm);
    int d1 = 0, tg = 0, tin = 0;  // D1 slot, T group slot, token inside the group
    uint32_t d1ph = 0, tph = 0;
    for (int lt = 0; lt < my_tokens; ++lt) {
      const uint32_t tbuf = sT + (uint32_t)(tg * G + tin) * (uint32_t)c.t_bytes;
      for (int i = 0; i < c.tpt; ++i) {
        const int rows_here = min(c.xrows, c.Kpad - i * c.xrows);
        if (i == 0 && tin == 0) mbar_wait_sleep(t_empty(tg), tph ^ 1u);
        if (warp == EPI1_WARP0 && i == 0) stamp(lt, 11);
        mbar_wait_sleep(d1_full(d1), d1ph);
        if (warp == EPI1_WARP0 && i == 0) stamp(lt, 12);
        tc_fence_after();
        if (quad * 32 < rows_here) {
          const int k = i * c.xrows + l;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(d1 * P);
          const uint32_t trow = tbuf + (uint32_t)k * 128u;
          const uint32_t sw = (uint32_t)(k & 7);
          const bool row_ok = l < rows_here;
          if (NS == 1) {
            uint32_t va[32], vb[32];
            tmem_ld32_nowait(taddr, va);
            tmem_ld32_nowait(taddr + 32, vb);
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
              tmem_ld32_nowait(taddr + q4 * 32, va);
              tmem_ld32_nowait(taddr + 64 + q4 * 32, vb);
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
        if (warp == EPI1_WARP0) stamp(lt, i == 0 ? 13 : 14);
        if (lane == 0) {
          mbar_arrive(t_full(tg));
          mbar_arrive(d1_empty(d1));
        }
        if (++d1 == D1N) { d1 = 0; d1ph ^= 1u; }
      }
      if (++tin == G || lt == my_tokens - 1) {
        // (an incomplete last group: the arrivals of the missing tokens)
        if (lane == 0)
          for (int m = 0; m < (G - tin) * c.tpt; ++m) mbar_arrive(t_full(tg));
        tin = 0;
        if (++tg == c.tg) { tg = 0; tph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int NS>
static int launch(const HadArgs& a, cudaStream_t st, int expand = 1) {
  constexpr int P = 64 * NS;
  constexpr int G = 2 / NS;
  Params prm;
  prm.a = a;
  Cfg& c = prm.c;
  c.src_k = a.had_k;
  c.expand = expand;
  c.had_k = a.had_k * expand;
  c.Kpad = round_up(c.had_k, 16);
  c.kb2 = (c.Kpad + 63) / 64;
  c.mt = c.Kpad > 128 ? 2 : 1;
  c.t_bytes = NS * c.Kpad * 128;
  const int a2 = round_up(c.kb2 * c.Kpad * 128, 1024);
  const int hb = 64 * 128;
  const int bars = 8 * (2 * MAX_XS + 17 + 2 * MAX_TG) + 16;
  const int budget = 227 * 1024 - 1024 - round_up(bars, 128);
  // X tiles: the token's Kpad rows in equal parts of at most 128 rows (a multiple of 16: whole
  // k-steps of MMA2 and whole swizzle atoms).  MMA1 always reads 128 rows; what lies past a short
  // tile is the next slot (or T), whose D1 rows nobody reads.
  c.tpt = (c.Kpad + 127) / 128;
  c.xrows = round_up((c.Kpad + c.tpt - 1) / c.tpt, 16);
  // T: two groups at least (one being filled while the other is consumed), three if X keeps four slots
  auto fits = [&](int tg, int xs) { return a2 + hb + tg * G * c.t_bytes + xs * NS * c.xrows * 128 <= budget; };
  c.tg = fits(3, 4) ? 3 : 2;
  if (!fits(c.tg, 2)) return MQ_ERR_UNSUPPORTED;
  c.x_bytes = NS * c.xrows * 128;
  c.xs = (budget - a2 - hb - c.tg * G * c.t_bytes) / c.x_bytes;
  if (c.xs > MAX_XS) c.xs = MAX_XS;
  c.a2_off = 0;
  c.h_off = a2;
  c.x_off = c.h_off + hb;
  c.t_off = c.x_off + c.xs * c.x_bytes;
  c.bar_off = c.t_off + c.tg * G * c.t_bytes;
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
  const int grid = (int)(a.M < num_sms() ? a.M : num_sms());
#ifdef MQ_PROBES
  if (NS == 1 && getenv("MQ_HAD_D1N4") != nullptr) {  // (experiment: four D1 tiles + two D2 jobs)
    MQ_DYN_SMEM((hadk_tc3_kernel<NS, 4 / NS>), smem);
    hadk_tc3_kernel<NS, 4 / NS><<<grid, NUM_THREADS, smem, st>>>(tmap, prm);
    return MQ_OK;
  }
#endif
  MQ_DYN_SMEM((hadk_tc3_kernel<NS, 2>), smem);
  hadk_tc3_kernel<NS, 2><<<grid, NUM_THREADS, smem, st>>>(tmap, prm);
  return MQ_OK;
}

}  // namespace hadtc3

#ifdef MQ_STAMPS
extern "C" int mq_debug_hadk_tc3_stamps(long long* dst_host) {
  return cudaMemcpyFromSymbol(dst_host, hadtc3::g_stamps, sizeof(long long) * 96 * 16) == cudaSuccess ? 0 : -3;
}
#endif

// had_k > 1, fp16 rounding points of matmul_hadU_cuda, codes and/or fp16 output.  n / had_k in {64, 128}
// directly; n / had_k in {256, 512} (InternVL2-8B: 14336 = 28 x 512) as the composite factor
// hadK (x) H_E with E = n / had_k / 128 against 128-point blocks: the element index k * P + jh * 128 + jl
// is (k * E + jh) * 128 + jl, and H_P = H_E (x) H_128.  The first fp16 rounding point then sits after
// the 128-point part of the FHT instead of after all of it: same two roundings, different place, so the
// result differs from the reference's by the size of its own intermediate-rounding noise (measured: up
// to a few fp16 ulp on entries far below the row maximum).  That is more than the summation-order
// tolerance of the default path, hence OPT-IN (MQ_FLAG_FAST_HAD); without the flag these shapes run on
// mq_hadamard_tc2.cu, which keeps the reference's rounding points.
static int tc3_expand(const mq::HadArgs& a) {
  const int P = a.n / a.had_k;
  if (P == 64 || P == 128) return 1;
  if ((P == 256 || P == 512) && a.had_k * (P / 128) <= 256 && (a.flags & MQ_FLAG_FAST_HAD)) return P / 128;
  return 0;
}

bool hadk_tc3_supported(const HadArgs& a) {
  if (a.had_k <= 1 || a.had_k > 256 || a.n % a.had_k != 0 || tc3_expand(a) == 0) return false;
  const int P = a.n / a.had_k / tc3_expand(a);
  if (a.q == nullptr && a.y == nullptr) return false;
  if (a.flags & MQ_FLAG_FP32_HAD) return false;
  if (a.n_in % P != 0 || a.M >= (1ll << 31)) return false;
  const int f = forced_hadamard_kernel();
  return f != 1 && f != 2;
}

int launch_hadk_tc3(const HadArgs& a, cudaStream_t st) {
  const int e = tc3_expand(a);
  if (e > 1) return hadtc3::launch<2>(a, st, e);
  return a.n / a.had_k == 64 ? hadtc3::launch<1>(a, st) : hadtc3::launch<2>(a, st);
}

}  // namespace mq
