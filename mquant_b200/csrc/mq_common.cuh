// Shared helpers for the mquant_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/mquant_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mquant_b200 is written for sm_100a (B200) only"
#endif

namespace mq {

// ---- error plumbing (thread-local message, never throws across the ABI) -------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int num_sms();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of per launch
int ensure_dyn_smem(const void* kernel, int bytes);
// test / tooling overrides (mq_debug_force_*): 0 = automatic dispatch
int forced_hadamard_kernel();

#define MQ_REQUIRE(cond, ...)       \
  do {                              \
    if (!(cond)) {                  \
      mq::set_error(__VA_ARGS__);   \
      return MQ_ERR_INVALID;        \
    }                               \
  } while (0)

#define MQ_CUDA(call)                                   \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return mq::cuda_fail(_e, #call); \
  } while (0)

#define MQ_DYN_SMEM(kernel, bytes)                                             \
  do {                                                                         \
    const int _rc = mq::ensure_dyn_smem((const void*)(kernel), (int)(bytes));  \
    if (_rc != MQ_OK) return _rc;                                              \
  } while (0)

#define MQ_LAUNCH_CHECK(name)                                      \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return mq::cuda_fail(_e, name " launch"); \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Division of a 31-bit index by a run-time constant with one multiply-high (Granlund-Montgomery;
// the flattened (row, chunk) loops of the HBM-bound kernels would otherwise pay a 64-bit divide
// per 16-byte chunk).
struct FastDiv {
  uint32_t mul, shr, d;
  FastDiv() : mul(0), shr(0), d(1) {}
  explicit FastDiv(uint32_t denom) : mul(0), shr(0), d(denom) {
    if (denom > 1) {
      uint32_t lg = 0;
      while ((1u << lg) < denom) ++lg;
      const uint32_t p = 31 + lg;
      mul = (uint32_t)(((1ull << p) + denom - 1) / denom);
      shr = p - 32;
    }
  }
#ifdef __CUDACC__
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
#endif
};

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__

// 128-bit streaming global accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream8(void* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ---- the reference quantiser -----------------------------------------------------------------
// code = clamp(rne(fl32(x / s)), lo, hi) with a TRUE fp32 division (uniform.py:28-31; zero_point
// is 0 for the signed symmetric types that are reachable).
//
// A correctly rounded division per element would make these HBM-bound kernels issue-bound
// (IEEE div ~8 instructions, F2I/FRND run at 1/8 rate).  Instead: r = x * fl(1/s) is within
// 3*2^-24 relative of fl(x/s); rne(r) can differ from rne(fl(x/s)) only if r lies within that
// distance of a half-integer.  Those (rare) elements are recomputed with __fdiv_rn, everything
// else takes FMUL + clamp + magic-number rounding (1.5*2^23: FADD in round-to-nearest-even IS
// rne, and the integer sits in the low mantissa bits).  The result is bit-identical to the
// division for every input.
#define MQ_RND_MAGIC 12582912.0f

__device__ __forceinline__ float quant_code_exact(float x, float s, float lo, float hi) {
  // "+ zero_point" (an exact +0.0 for the symmetric types) turns a -0.0 quotient into +0.0
  // (uniform.py:28); rintf keeps the sign of a tiny negative quotient (-0.0)
  float q = rintf(__fadd_rn(__fdiv_rn(x, s), 0.f));
  return fminf(fmaxf(q, lo), hi);
}

// generic-width code as a float (fake-quant path, |hi| <= 2^19); the sign of a zero result
// follows torch.round, see quant_code_exact
__device__ __forceinline__ float quant_code(float x, float s, float inv, float lo, float hi) {
  const float r = __fadd_rn(__fmul_rn(x, inv), 0.f);
  const float rc = fminf(fmaxf(r, lo), hi);
  const float rr = __fadd_rn(__fadd_rn(rc, MQ_RND_MAGIC), -MQ_RND_MAGIC);
  const float d = fabsf(fabsf(rc - rr) - 0.5f);
  if (d <= fmaxf(fabsf(rc), 1.f) * 1e-6f) return quant_code_exact(x, s, lo, hi);
  return copysignf(rr, rc);
}

__device__ __forceinline__ uint32_t q8_exact(float x, float s) {
  return (uint32_t)__float2int_rn(quant_code_exact(x, s, -128.f, 127.f));
}

// ---- INT8 fast path ("bracketing") -----------------------------------------------------------
// Per row-scale constants.  fl(x/s) lies strictly between x*inv_lo and x*inv_hi (inv_* = 1/s
// moved 2^-22 apart; the error budget is 2^-24 each for 1/s, for the quotient rounding and for
// forming inv_*).  Rounding to integer is monotone, so when both ends round to the same integer
// -- one FFMA each against the 1.5*2^23 magic constant -- that integer IS rne(fl(x/s)).  The
// (rare) chunks where some element disagrees are redone with the true division.  The clamp is
// applied to x in the fp16 domain (two packed ops per pair): any fp16 value in
// [126.6 s, 127.4 s] works as the upper limit because everything above it has code 127, and
// fp16's 2^-11 spacing always leaves such a value (same for -128).
struct QScale {
  float s, inv_lo, inv_hi;
  __half2 lim_lo, lim_hi;
  bool narrow;  // s has at most 11 significant bits (an fp16- or bf16-valued scale): see quant8
};
__device__ __forceinline__ QScale make_qscale(float s) {
  QScale q;
  q.s = s;
  const float inv = __frcp_rn(s);
  q.inv_lo = __fmul_rn(inv, 1.0f - 2.384185791015625e-07f);
  q.inv_hi = __fmul_rn(inv, 1.0f + 2.384185791015625e-07f);
  q.lim_hi = __float2half2_rn(fminf(__fmul_rn(127.0f, s), 65504.f));
  q.lim_lo = __float2half2_rn(fmaxf(__fmul_rn(-128.0f, s), -65504.f));
  q.narrow = (__float_as_uint(s) & 0x1fffu) == 0u;
  return q;
}

__device__ __forceinline__ uint32_t pack4_lowbytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// 8 halfs (one uint4) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8_f16(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// 8 halfs (one uint4) -> 8 int8 codes (one uint2); bit-exact w.r.t. quant_code_exact for ANY
// fp32 scale (dynamic per-token scales, percentile scales)
__device__ __forceinline__ uint2 quant8_bracket(const uint4& v, const QScale& q) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  uint32_t t[8];
  uint32_t diff = 0;
  // both ends of the bracket for two elements per instruction (sm_100 packed fp32 FMA: each half
  // is an ordinary round-to-nearest fma)
  uint64_t inv_lo2, inv_hi2, magic2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(inv_lo2) : "f"(q.inv_lo));
  asm("mov.b64 %0, {%1, %1};" : "=l"(inv_hi2) : "f"(q.inv_hi));
  asm("mov.b64 %0, {%1, %1};" : "=l"(magic2) : "f"(MQ_RND_MAGIC));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(__hmin2(__hmax2(h[i], q.lim_lo), q.lim_hi));
    uint64_t x, lo, hi;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(f.x), "f"(f.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(lo) : "l"(x), "l"(inv_lo2), "l"(magic2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(hi) : "l"(x), "l"(inv_hi2), "l"(magic2));
    uint32_t h0, h1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t[2 * i]), "=r"(t[2 * i + 1]) : "l"(lo));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(h0), "=r"(h1) : "l"(hi));
    diff |= t[2 * i] ^ h0;  // (a ^ b) | diff is one LOP3
    diff |= t[2 * i + 1] ^ h1;
  }
  // Undecided elements (the bracket straddles a half-integer: mostly quotients that round to an
  // exact tie, ~1e-4 .. 1e-2 of elements depending on how "round" the scale is) take the true
  // division, element by element so that a warp pays only for the lanes that need it.
  if (diff != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = __half2float(reinterpret_cast<const __half*>(&v)[i]);
      const float xc = fminf(fmaxf(x, __low2float(q.lim_lo)), __low2float(q.lim_hi));
      if (t[i] != __float_as_uint(__fmaf_rn(xc, q.inv_hi, MQ_RND_MAGIC))) t[i] = q8_exact(x, q.s);
    }
  }
  uint2 r;
  r.x = pack4_lowbytes(t[0], t[1], t[2], t[3]);
  r.y = pack4_lowbytes(t[4], t[5], t[6], t[7]);
  return r;
}
// ---- narrow scales: no comparison, no slow path ------------------------------------------------
// MinmaxObserver scales are fp16 (bf16) VALUES: s = S * 2^b with an integer S < 2^11, like every
// activation x = X * 2^a.  Then x / s - (k + 1/2) = (X 2^(a-b+1) - S (2k+1)) / (2 S) is either 0
// (an exact tie) or at least 1/(2S) > 2^-12 away from zero, i.e. >= 2^-20 relative for |x/s| <=
// 128 -- wider than the bracket (x*inv_lo, x*inv_hi), 2^-21.4 relative.  So the two ends of the
// bracket round to different integers ONLY at exact ties, where they are k and k+1, and
//   fl(t_lo + t_hi) = 2M + 2k + 1 -> rounded to a multiple of 2 (ulp of [2^24, 2^25)), ties to an
//   even mantissa = 2 (M + even(k, k+1))
// is round-half-to-even: bits = 0x4BC00000 + code whether the ends agree or not.  One packed add
// replaces the per-element comparison, the branch and the division fallback.
// packed constants of a narrow scale
struct QNarrow {
  uint64_t inv_lo2, inv_hi2;
  __half2 lim_lo, lim_hi;
};
__device__ __forceinline__ QNarrow make_qnarrow(const QScale& q) {
  QNarrow n;
  asm("mov.b64 %0, {%1, %1};" : "=l"(n.inv_lo2) : "f"(q.inv_lo));
  asm("mov.b64 %0, {%1, %1};" : "=l"(n.inv_hi2) : "f"(q.inv_hi));
  n.lim_lo = q.lim_lo;
  n.lim_hi = q.lim_hi;
  return n;
}
// two fp16 values (one 32-bit word) -> two fp32 words whose LOW BYTES are the codes
__device__ __forceinline__ void quant2_narrow(uint32_t w, const QNarrow& q, uint32_t& c0, uint32_t& c1) {
  uint64_t magic2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(magic2) : "f"(MQ_RND_MAGIC));
  const float2 f = __half22float2(__hmin2(__hmax2(*reinterpret_cast<const __half2*>(&w), q.lim_lo), q.lim_hi));
  uint64_t x, lo, hi, sum;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(f.x), "f"(f.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(lo) : "l"(x), "l"(q.inv_lo2), "l"(magic2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(hi) : "l"(x), "l"(q.inv_hi2), "l"(magic2));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum) : "l"(lo), "l"(hi));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(c0), "=r"(c1) : "l"(sum));
}
// four fp16 values in two words (a = elements 0,1; b = elements 2,3) -> 4 codes, byte i = element i
__device__ __forceinline__ uint32_t quant4_narrow(uint32_t a, uint32_t b, const QNarrow& q) {
  uint32_t t0, t1, t2, t3;
  quant2_narrow(a, q, t0, t1);
  quant2_narrow(b, q, t2, t3);
  return pack4_lowbytes(t0, t1, t2, t3);
}
__device__ __forceinline__ uint2 quant8_narrow(const uint4& v, const QNarrow& n) {
  uint2 r;
  r.x = quant4_narrow(v.x, v.y, n);
  r.y = quant4_narrow(v.z, v.w, n);
  return r;
}
// 8 halfs -> 8 codes: narrow scales take the branch-free path, any other scale the bracket with
// its division fallback (kernels hoist the test out of their loops where it matters)
__device__ __forceinline__ uint2 quant8(const uint4& v, const QScale& q) {
  if (q.narrow) return quant8_narrow(v, make_qnarrow(q));
  return quant8_bracket(v, q);
}
// same from 8 floats that already hold fp16 values (Hadamard epilogues)
__device__ __forceinline__ uint2 quant8(const float (&f)[8], const QScale& q) {
  return quant8(pack8_f16(f), q);
}

// fp16 round trip: the reference's rounding point after each Hadamard factor
__device__ __forceinline__ float round_f16(float x) { return __half2float(__float2half_rn(x)); }

#endif  // __CUDACC__

}  // namespace mq
