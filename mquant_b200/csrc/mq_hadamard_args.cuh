// Argument block shared by the Hadamard kernels.
#pragma once
#include "mq_common.cuh"

namespace mq {

struct HadArgs {
  const __half* x;
  int64_t ldx;
  int8_t* q;        // nullable
  __half* y;        // nullable
  float* srow_out;  // nullable
  __half* x0_out;   // nullable (split)
  const uint8_t* modality;
  const float* scales;  // nullable when q == nullptr
  const __half* hadk;   // [had_k, had_k] +-1
  int had_k;
  int64_t M;
  int n_in;
  int n;
  unsigned flags;
  float norm;  // float32(1/sqrt(n))
};


// mq_hadamard_tc2.cu: power-of-two factor in registers, hadK factor on tcgen05 (n / had_k in {256, 512}; mq_debug_force_hadamard_kernel(2))
bool hadk_tc2_supported(const HadArgs& a);
int launch_hadk_tc2(const HadArgs& a, cudaStream_t st);

// mq_hadamard_tc3.cu: n / had_k in {64, 128}: both factors on tcgen05, hadK as the A operand (default)
bool hadk_tc3_supported(const HadArgs& a);
int launch_hadk_tc3(const HadArgs& a, cudaStream_t st);

// mq_hadamard_mma.cu: power-of-two n in {1024, 2048, 4096}: 16-point factors on mma.sync, warp per token
bool fwht_mma_supported(const HadArgs& a);
int launch_fwht_mma(const HadArgs& a, cudaStream_t st);

// mq_hadamard_p2.cu: power-of-two n in {1024 ... 8192}, M >= 1024: both Kronecker factors on tcgen05 with a
// hi / lo split of the intermediate (fp32 grade); mq_debug_force_hadamard_kernel(3) = off, (4) = for every M
bool hadp2_tc_supported(const HadArgs& a);
int launch_hadp2_tc(const HadArgs& a, cudaStream_t st);

}  // namespace mq
