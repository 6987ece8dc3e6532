// Shared device/host helpers for the sm_100a kernels of libmmlrec_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmlrec_b200.h"

namespace mmlrec {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MMLREC_CHECK_ARG(cond, msg)                       \
  do {                                                    \
    if (!(cond)) {                                        \
      mmlrec::set_error("%s: %s", __func__, msg);         \
      return -1;                                          \
    }                                                     \
  } while (0)

// call after a kernel launch
#define MMLREC_RETURN_LAUNCH(n)                                               \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      mmlrec::set_error("%s: %s", __func__, cudaGetErrorString(e__));         \
      return (int)e__;                                                        \
    }                                                                         \
    mmlrec::count_launch(n);                                                  \
    return 0;                                                                 \
  } while (0)

#define MMLREC_CHECK_LAUNCH(n)                                                \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      mmlrec::set_error("%s: %s", __func__, cudaGetErrorString(e__));         \
      return (int)e__;                                                        \
    }                                                                         \
    mmlrec::count_launch(n);                                                  \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ---------------------------------------------------------------------------
// A training step is ~25 short kernels in one stream / CUDA graph; a plain launch starts only after the previous
// grid has drained completely.  Launched with the programmatic-stream-serialization attribute, the next grid is
// scheduled as soon as every CTA of the previous one has started (each kernel begins with pdl_launch_dependents())
// and its CTAs wait in pdl_wait() -- which returns when the previous grid has completed and its writes are visible
// -- before touching any data a predecessor produced.  EVERY thread of such a kernel executes pdl_wait() before it
// can exit, so completion stays transitive along the stream.  MMLREC_PDL selects who uses it (capi.cu): by default only the tensor-core GEMM, whose barrier / TMEM / table set-up then overlaps the previous kernel's tail (-4.5 % step time); kernels that can only wait at their very top lose by being scheduled early (profiles/pdl_r02.txt).
bool pdl_enabled();
bool pdl_enabled_gemm();
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;   // one F2FP: d = {hi, lo}, round to nearest even (same result as two scalar conversions)
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == MMLREC_ACT_RELU) return v > 0.f ? v : 0.f;  // matches at::relu for finite inputs
  if (act == MMLREC_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == MMLREC_ACT_SIGMOID2) return 2.f / (1.f + expf(-v));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cp.async 16-byte global -> shared (LDGSTS); src must be 16B aligned
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// One optimizer row/element update, shared by the fused embedding update and the dense sweep.
// Op order follows torch's _single_tensor_adam / _single_tensor_adagrad / rmsprop (fp32).
__device__ __forceinline__ void optimizer_update(float& p, float g, float& s1, float& s2, const MmlrecHyper& h) {
  if (h.optimizer == MMLREC_OPT_ADAM) {
    // exp_avg.lerp_(grad, 1-beta1)  (weight < 0.5 branch of ATen lerp: a + w*(b-a))
    s1 = s1 + h.one_minus_beta1 * (g - s1);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    s2 = s2 * h.beta2 + h.one_minus_beta2 * g * g;
    float denom = sqrtf(s2) / h.bc2_sqrt + h.eps;
    p = p - h.step_size * (s1 / denom);
  } else if (h.optimizer == MMLREC_OPT_ADAGRAD) {
    s1 = s1 + g * g;
    p = p - h.lr * (g / (sqrtf(s1) + h.eps));
  } else if (h.optimizer == MMLREC_OPT_RMSPROP) {
    s1 = s1 * h.alpha + h.one_minus_alpha * g * g;
    p = p - h.lr * (g / (sqrtf(s1) + h.eps));
  } else {  // SGD, no momentum
    p = p - h.lr * g;
  }
}

}  // namespace mmlrec
