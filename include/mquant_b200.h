/*
 * mquant_b200 -- C ABI of the B200-native quantized-linear hot path.
 *
 * This is the drop-in boundary: a plain C interface (device pointers + sizes, no torch
 * types) that replaces the native seam the reference sits on today and the eager torch
 * chains around it.  Reference = StiphyJay/MQuant, paths relative to its root:
 *
 *   - fast_hadamard_transform.hadamard_transform(x, scale)    (third-party CUDA ext,
 *       imported at fake_quant/hadamard_utils.py:2, called at :118,:124)
 *   - hadK.to(dtype) @ input                                  fake_quant/hadamard_utils.py:127
 *   - UniformQuantizer.quant / dequantize                     fake_quant/quantizer/uniform.py:20-43
 *   - MinmaxObserver.update / PercentileObserver.update       fake_quant/observer/minmax.py:13-28,
 *                                                             fake_quant/observer/percentile.py:26-55
 *   - ActQuantizer.find_params (dynamic per-token)            fake_quant/quant_utils.py:205-268
 *   - WeightQuantizer.find_params / quantize, pack_i4         fake_quant/quant_utils.py:446-518, :61-68
 *   - ActQuantWrapper.forward: self.module(x) / L1,L2 split   fake_quant/quant_utils.py:367-384
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (torch tensor .data_ptr()) unless
 *     the parameter is documented as host.  The library allocates nothing persistent.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises.
 *   - Return value: 0 on success, negative mq_status on failure; mq_last_error() returns a
 *     thread-local message.  Nothing throws across the ABI.
 *   - Row-major matrices; `ld*` are row strides in ELEMENTS.  Activations are IEEE fp16
 *     (the reference runs model.half(), exam/quant_qwenvl.py:82).
 *   - "codes" are the reference's integers: int8 in [-128,127] for activations
 *     (bit_type.py:44), int4 in [-8,7] for weights (quant_utils.py:441-442).
 *   - Static activation scales live in device memory as float[2] = {s_text, s_visual}
 *     (each the fp16-valued scale of MinmaxObserver widened to fp32).  `modality` is a
 *     nullable uint8[M]; nonzero selects s_visual for that token (MSQ).  NULL => s_text.
 */
#ifndef MQUANT_B200_H
#define MQUANT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQ_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum mq_status {
  MQ_OK = 0,
  MQ_ERR_INVALID = -1,     /* bad shape / alignment / null pointer */
  MQ_ERR_UNSUPPORTED = -2, /* valid request this build has no kernel for */
  MQ_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed */
  MQ_ERR_WORKSPACE = -4    /* caller-provided workspace too small */
} mq_status;

/* flags shared by the activation kernels */
#define MQ_FLAG_SPLIT 1u    /* RMS split (quant_utils.py:367-376): channel 0 is carried in fp16
                               (x0_out[m] = x[m,0]) and its code is forced to 0 */
#define MQ_FLAG_FP32_HAD 2u /* --fp32_had (quant_utils.py:336-339): one rounding after the whole
                               rotation instead of fp16 rounding after each factor */
#define MQ_FLAG_FAST_HAD 4u /* opt-in, n / had_k in {256, 512} only: the first fp16 rounding point
                              sits after the 128-point part of the FHT factor instead of after all of it
                              (the rest of the factor joins hadK on the tensor core as hadK (x) H_E).
                              Same two roundings, different place: outputs differ from the reference's
                              by its own intermediate-rounding noise (a few fp16 ulp on entries far
                              below the row maximum, codes +-1 on < 1 %); 3.4x faster at 28 x 512. */

int mq_version(void);
const char* mq_last_error(void);

/* Number of SMs / device id the library sees on the current device (diagnostics). */
int mq_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- (a) activation quantisation ------------------------------------------------------- */

/* Static INT8 quantisation.  Replaces UniformQuantizer.quant (uniform.py:20-32):
 *   q[m,c] = clamp(rne(float(x[m,c]) / s_m), -128, 127),  s_m = scales[modality[m] != 0]
 * Columns [n_in, n_out) of q are zero (utils.revise_down_input, fake_quant/utils.py:465-471).
 * srow_out (nullable float[M]) receives s_m for the GEMM epilogue.
 * Requires n_in % 8 == 0, n_out % 16 == 0, ldx % 8 == 0, 16-byte aligned pointers. */
int mq_act_quant_i8(const void* x, int64_t ldx, int8_t* q, float* srow_out, void* x0_out,
                    const uint8_t* modality, const float* scales, int64_t M, int n_in,
                    int n_out, unsigned flags, void* stream);

/* Fake quantisation, drop-in for BaseQuantizer.forward (quantizer/base.py:44-50):
 *   y = fp16( clamp(rne(float(x)/s + zp), lo, hi) - zp ) * s ),  lo/hi from `bits` (signed).
 * scale_mode 0: scales = float[2] + modality (as above); 1: scales = float[C] per channel
 * (calibration_mode="channel_wise").  y may alias x. */
int mq_act_fake_quant_f16(const void* x, int64_t ldx, void* y, int64_t ldy,
                          const uint8_t* modality, const float* scales, int scale_mode,
                          int bits, int64_t M, int C, void* stream);

/* Dynamic per-token symmetric quantisation (ActQuantizer.find_params, quant_utils.py:236-250,
 * then sym_quant :42-45): srow_out[m] = max(|min(x_m,0)|, max(x_m,0)) * clip_ratio / 127
 * (1 for an all-zero row), q = clamp(rne(x / s_m), -128, 127). */
int mq_act_quant_i8_dynamic(const void* x, int64_t ldx, int8_t* q, float* srow_out, int64_t M,
                            int n_in, int n_out, float clip_ratio, void* stream);

/* RMS normalisation fused with the static quantiser (RMSN.forward,
 * fake_quant/module_util.py:42-61, followed by the ActQuantizer of the Linears that read it):
 *   r_m = 1 / sqrt( sum_c float(x[m,c])^2 / mean_dim + eps )      (fp32)
 *   y[m,c] = fp16( float(x[m,c]) * r_m )                          (nullable: plain RMSN when q is NULL)
 *   q[m,c] = clamp(rne(y[m,c] / s_m), -128, 127), zero for C <= c < n_out; s_m by modality as
 *   in mq_act_quant_i8 (srow_out nullable).  ldy = C. */
int mq_rmsn_quant_i8(const void* x, int64_t ldx, int8_t* q, void* y_f16, float* srow_out,
                     const uint8_t* modality, const float* scales, int64_t M, int C, int n_out,
                     float mean_dim, float eps, void* stream);

/* ---- (a) online Hadamard, fused with the quantiser ------------------------------------- */

/* y = x . (hadK (x) H_{n/had_k})^T / sqrt(n)  (matmul_hadU_cuda, hadamard_utils.py:115-128),
 * x zero-padded from n_in to n first, fp16 rounding after the FHT factor and after the hadK
 * factor as the reference does (unless MQ_FLAG_FP32_HAD), then the static INT8 quantiser of
 * mq_act_quant_i8.  had_k == 1 -> n is a power of two.  had_k > 1 -> `hadk` is the
 * [had_k, had_k] +-1 matrix of get_hadK as fp16, row-major, in device memory.
 * Exactly one of q / y_f16 may be NULL: q = INT8 codes, y_f16 = rotated fp16 activations
 * (what calibration observes and what matmul_hadU_cuda returns). */
int mq_hadamard_quant_i8(const void* x, int64_t ldx, int8_t* q, void* y_f16, float* srow_out,
                         void* x0_out, const uint8_t* modality, const float* scales,
                         const void* hadk, int had_k, int64_t M, int n_in, int n,
                         unsigned flags, void* stream);

/* The native seam of the reference, as it stands: fast_hadamard_transform.hadamard_transform(x,
 * scale) (third-party CUDA extension; imported at fake_quant/hadamard_utils.py:2 and called at
 * :118,:124,:168,:178,:188, quant_utils.py:351, rotation_utils.py:11):
 *   y = fp16( fp32(x . H_n) * scale ),  H_n the Sylvester matrix, n a power of two in [64, 8192],
 * fp32 butterflies, one rounding.  x: fp16 [M, n] row stride ldx; y: fp16 [M, n] contiguous
 * (ldy == n).  compat/fast_hadamard_transform binds it under the original module name. */
int mq_hadamard_transform_f16(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, int n,
                              float scale, void* stream);

/* ---- (c) observers ----------------------------------------------------------------------- */

/* MinmaxObserver.update (minmax.py:13-28).  minmax_io is float[2] = {min, max} (layer_wise)
 * or float[2*C] = {min[C], max[C]} (per_channel != 0), holding the RUNNING statistics
 * (initialise to 0: the reference folds 0 in on the first update).  Merged atomically, so
 * several calls / streams / ranks may feed one buffer.  Values are exact (max/min of fp16). */
int mq_observer_minmax(const void* x, int64_t ldx, int64_t M, int C, int per_channel,
                       float* minmax_io, void* stream);

/* The same for Modality-Specific Static scales (layer_wise only): ONE pass merges the rows with
 * modality[m] == 0 into minmax_text_io and the others into minmax_visual_io (each float[2]
 * {min, max}) and adds the number of rows of either kind to seen_io (unsigned[2] {text, visual}),
 * so that the host needs neither x[mask] gathers nor a device round trip to learn whether a
 * modality occurred.  Replaces two MinmaxObserver.update calls (minmax.py:13-28) on the two
 * halves of the batch. */
int mq_observer_minmax_msq(const void* x, int64_t ldx, int64_t M, int C, const uint8_t* modality,
                           float* minmax_text_io, float* minmax_visual_io, unsigned* seen_io,
                           void* stream);

/* PercentileObserver.update's order statistics (percentile.py:31-32): exact
 * torch.quantile(v.flatten().float(), q) for q in {alpha, 1-alpha}, linear interpolation,
 * via a two-pass counting select on the fp16 bit patterns.  out = float[2] {q_lo, q_hi}
 * (written, not merged: the EMA of percentile.py:44-55 is order dependent and stays on the
 * host).  workspace: >= mq_observer_percentile_workspace() bytes, zeroed by the call. */
size_t mq_observer_percentile_workspace(void);
int mq_observer_percentile(const void* x, int64_t ldx, int64_t M, int C, double alpha,
                           float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- (b) weights and the W4A8 GEMM ------------------------------------------------------ */

/* WeightQuantizer per-output-channel symmetric RTN (quant_utils.py:446-466, :42-45):
 *   s_w[n] = max(|min(W_n,0)|, max(W_n,0), 1e-5) / (2^(bits-1)-1)   (fp32)
 *   codes[n,k] = clamp(rne(float(W[n,k]) / s_w[n]), -2^(bits-1), 2^(bits-1)-1)  (int8)
 * If scales_in != NULL those scales are used instead (GPTQ / clip-search output). */
int mq_weight_quant_rtn(const void* w_f16, int64_t ldw, int8_t* codes, float* s_w,
                        const float* scales_in, int64_t N, int K, int bits, void* stream);

/* Bytes of the GEMM's packed-weight buffer for an [N, K] matrix. */
size_t mq_w4_packed_bytes(int64_t N, int64_t K);

/* int8 codes [N, K] (values in [-8, 7]) -> the GEMM's tile-major, nibble-pre-shuffled INT4
 * layout (documented in DESIGN.md).  Column 0 can be zeroed (split). */
int mq_pack_w4(const int8_t* codes, int64_t ldc, uint8_t* w4, int64_t N, int64_t K,
               unsigned flags, void* stream);
/* Same, from the reference's interchange format pack_i4 (quant_utils.py:61-68): [N, K/2]
 * uint8, even k in the low nibble. */
int mq_repack_w4_from_ref(const uint8_t* ref_packed, uint8_t* w4, int64_t N, int64_t K,
                          void* stream);
/* Inverse of mq_pack_w4 (for checkpoint export / tests). */
int mq_unpack_w4(const uint8_t* w4, int8_t* codes, int64_t ldc, int64_t N, int64_t K,
                 void* stream);

/* out[m,n] = fp16( fp32(sum_k qa[m,k]*w[n,k]) * (s_a[m] * s_w[n]) (+ x0[m]*w0[n]) (+ bias[n]) )
 * -- exact int32 accumulation on tcgen05 (kind::i8, TMEM accumulators), INT4 weights unpacked
 * to INT8 in shared memory, each epilogue operation one fp32 rounding in the order written.
 * Replaces self.module(x) (quant_utils.py:384) and the fp32 L1/L2 split (:374-376).
 *   qa   int8 [M, K] row stride lda (K % 128 == 0: zero-padded codes)
 *   w4   from mq_pack_w4 for [N, K]
 *   s_a  float[M] if s_a_per_row else float[1]
 *   s_w  float[N];  bias fp16[N] nullable;  x0 fp16[M] + w0 fp16[N] nullable (rank-1 split term)
 *   out  fp16 [M, N] row stride ldo */
int mq_w4a8_gemm(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a,
                 int s_a_per_row, const float* s_w, const void* bias, const void* x0,
                 const void* w0, void* out, int64_t ldo, int64_t M, int64_t N, int64_t K,
                 void* stream);

/* Same contract with an optional workspace of at least mq_w4a8_gemm_workspace_bytes(M, N, K)
 * bytes, which enables the kernels that split K across CTAs and merge int32 partial sums
 * (integer adds: exact and order independent, results stay bit-identical):
 *   M <= 8 (<= 16 for N < 8192)  decode: weight-streaming dp4a kernel, bound by the N*K/2-byte
 *                                weight stream rather than by the tensor pipe;
 *   up to ~1000 rows             the tcgen05 kernel with K slices, while M/128 * N/256 tiles are
 *                                too few to fill the SMs.
 * The workspace must be ZERO before its first use; every call leaves it zero again, so one buffer
 * can be reused by all later calls on the same stream without clearing (calls on different
 * streams need their own).  workspace == NULL: identical to mq_w4a8_gemm.
 * mq_w4a8_gemm_workspace_bytes returns 0 when no workspace is useful for the shape. */
size_t mq_w4a8_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int mq_w4a8_gemm_ws(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a,
                    int s_a_per_row, const float* s_w, const void* bias, const void* x0,
                    const void* w0, void* out, int64_t ldo, int64_t M, int64_t N, int64_t K,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- (f-3) gated MLP: a1 * silu(a2) as the epilogue of the w1 | w2 GEMM ---------------------- */

/* QWenMLP.forward (model/modeling_qwen_opt.py:382-386): a1 = w1(x), a2 = w2(x), a1 * F.silu(a2), with
 * the two Linears as ONE W4A8 GEMM whose packed weight holds the rows interleaved in blocks of 128:
 * packed rows 256 b .. 256 b + 127 = channels 128 b .. of w1, rows 256 b + 128 .. = the same channels
 * of w2 (s_w and bias in the same order; N2 = 2 F, F % 128 == 0).
 *   out[m, f] = fp16( fp16(a1[m, f]) * fp16( silu( fp16(a2[m, f]) ) ) ),  silu(x) = x / (1 + exp(-x)) in fp32
 * (the reference's three fp16 rounding points; a1 / a2 = the mq_w4a8_gemm epilogue).  out: [M, F]
 * fp16, row stride ldo.  Shapes the 256 x 256 kernel serves (mq_w4a8_gemm_plan == 4) never touch
 * tmp; the others run the plain GEMM into tmp ([M, N2] fp16, required then) followed by
 * mq_silu_mul_interleaved_f16 -- same bits either way. */
int mq_w4a8_gemm_gated_ws(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a,
                          int s_a_per_row, const float* s_w, const void* bias, void* out, int64_t ldo,
                          int64_t M, int64_t N2, int64_t K, void* tmp, void* workspace,
                          size_t workspace_bytes, void* stream);

/* out[m, f] = in[m, 256 (f / 128) + f % 128] * silu(in[m, 256 (f / 128) + 128 + f % 128]), fp16,
 * the rounding points above; in: [M, 2 F] (row stride ldi), out: [M, F]. */
int mq_silu_mul_interleaved_f16(const void* in, int64_t ldi, void* out, int64_t ldo, int64_t M,
                                int64_t F, void* stream);

/* ---- (f-2) GPTQ acceleration --------------------------------------------------------------- */

/* GPTQ.add_batch (fake_quant/gptq/gptq_utils.py:182-195): H = alpha * H + beta * X^T X with
 * alpha = n / (n + t), beta = 2 / (n + t).  x: [T, C] fp16 (row stride ldx), H: [C, C] fp32,
 * contiguous.  fp16 x fp16 products are exact in fp32; the sums run on the tensor core in fp32. */
int mq_gptq_hessian_update(const void* x, int64_t ldx, int64_t T, int C, float* H, float alpha,
                           float beta, void* stream);

/* Inner loop of GPTQ.fasterquant (gptq_utils.py:254-283) for one block of `count` <= 128 columns,
 * symmetric per-row scales (WeightQuantizer.quantize -> sym_quant_dequant, quant_utils.py:42-57):
 *   for i: q = scale * clamp(round(w_i / scale), -(maxq+1), maxq);  Q[:, i] = q
 *          Losses[:, i] = (w_i - q)^2 / d^2;  err = (w_i - q) / d,  d = Hinv1[i, i];  Err[:, i] = err
 *          w[i:] -= err * Hinv1[i, i:]
 * W points at the block's first column (row stride ldw) and is updated in place; Hinv1 at the
 * block's diagonal element of the upper Cholesky factor (row stride ldh).  Every operation is an
 * fp32 operation in the reference's order. */
int mq_gptq_block_f32(float* W, int64_t ldw, const float* Hinv1, int64_t ldh, const float* scale,
                      float maxq, float* Q, float* Err, float* Losses, int rows, int count,
                      void* stream);

/* Which kernel mq_w4a8_gemm_ws runs for a shape and workspace size: 1 = weight-streaming dp4a
 * (decode), 2 = split-K tcgen05, 4 = 256 x 256 tcgen05, 5 = 128 channels x 256 tokens with the
 * converted weights resident in tensor memory (3 = the older 128 x 256 kernel, never planned);
 * 0 = invalid shape.  Pure function of its arguments and the SM count (diagnostics / tests). */
int mq_w4a8_gemm_plan(int64_t M, int64_t N, int64_t K, size_t workspace_bytes);

/* Test / tooling overrides of the dispatchers (process-wide; return the previous value):
 *   mq_debug_force_gemm_kernel(3 | 4 | 5)  always run the 128 x 256 / 256 x 256 / weights-in-TMEM tcgen05 kernel, 0 = auto
 *   mq_debug_force_hadamard_kernel(1)  CUDA-core Hadamard kernels only (cross-checks), 2 = the
 *                                      previous tensor-core generation for had_k > 1, 3 = power-of-two
 *                                      sizes on the mma.sync kernel for every M, 4 = ... on the tcgen05
 *                                      kernel for every M, 0 = auto */
int mq_debug_force_gemm_kernel(int kernel);
int mq_debug_force_hadamard_kernel(int kind);

/* CUDA-core reference of the same contract (slow; cross-check only, never on the product
 * path). */
int mq_w4a8_gemm_simt(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a,
                      int s_a_per_row, const float* s_w, const void* bias, const void* x0,
                      const void* w0, void* out, int64_t ldo, int64_t M, int64_t N, int64_t K,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MQUANT_B200_H */
