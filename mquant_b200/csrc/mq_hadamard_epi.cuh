// Epilogue helpers shared by the tcgen05 Hadamard kernels (mq_hadamard_tc3.cu, mq_hadamard_p2.cu):
// a thread owns 32 consecutive outputs of one token row -> fp16 rounding -> fp16 store and / or the
// exact INT8 quantiser -> 256-bit stores.
#pragma once
#include "mq_common.cuh"

namespace mq {
namespace hadepi {

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

// 32 fp32 accumulators of one output row -> fp16 (the second rounding point: hadK @ . -> fp16) ->
// fp16 store and/or INT8 codes, 32 columns starting at `col` of the token's row
template <bool NARROW>
__device__ __forceinline__ void epi2_emit(const uint32_t (&v)[32], int8_t* qrow, __half* yrow, int col, const QScale& qs,
                                          bool lane_ok, bool first, bool q32, bool y32, __half* x0_out, int64_t tok) {
  uint32_t h[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) h[e] = f16x2_rn(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
  if (yrow != nullptr && lane_ok) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t w[8] = {h[8 * u], h[8 * u + 1], h[8 * u + 2], h[8 * u + 3], h[8 * u + 4], h[8 * u + 5], h[8 * u + 6], h[8 * u + 7]};
      st_global_v8(yrow + col + u * 16, w, y32);
    }
  }
  if (qrow != nullptr) {
    uint32_t w[8];
    const QNarrow qn = make_qnarrow(qs);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 hv = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
      const uint2 r = NARROW ? quant8_narrow(hv, qn) : quant8_bracket(hv, qs);
      w[2 * u] = r.x;
      w[2 * u + 1] = r.y;
    }
    if (first) {  // RMS split: channel 0 travels in fp16, its code is zero
      if (lane_ok) x0_out[tok] = __ushort_as_half((unsigned short)(h[0] & 0xffffu));
      w[0] &= 0xffffff00u;
    }
    if (lane_ok) st_global_v8(qrow + col, w, q32);
  }
}

}  // namespace hadepi
}  // namespace mq
