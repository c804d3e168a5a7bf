// Shared helpers for libvqw.so (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/vqw.h"

// arithmetic modes (include/vqw.h): split hi/lo operands with 3 MMAs per product / IEEE-fp16 planes
static inline bool vqw_mode_x3(int m) { return m == VQW_MODE_BF16X3 || m == VQW_MODE_FP16X3; }
static inline bool vqw_mode_f16(int m) { return m == VQW_MODE_FP16 || m == VQW_MODE_FP16X3; }
static inline bool vqw_mode_tc(int m) {
  return m == VQW_MODE_BF16X3 || m == VQW_MODE_BF16 || m == VQW_MODE_FP16 || m == VQW_MODE_FP16X3;
}

namespace vqw {

// thread-local error text returned by vqw_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define VQW_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) return ::vqw::set_error(-1, __VA_ARGS__);  \
  } while (0)

// every kernel launch of the library passes through here (vqw_launch_count())
void count_launch();
// VQW_DEBUG_SYNC=1 in the environment: synchronise after every launch so a device fault is
// attributed to the kernel that raised it
bool debug_sync();

#define VQW_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    ::vqw::count_launch();                                                            \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ == cudaSuccess && ::vqw::debug_sync()) e__ = cudaDeviceSynchronize();     \
    if (e__ != cudaSuccess)                                                           \
      return ::vqw::set_error((int)e__, "%s: %s", name, cudaGetErrorString(e__));     \
  } while (0)

#define VQW_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return ::vqw::set_error((int)e__, "%s: %s", #expr, cudaGetErrorString(e__));    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }
// tanh through one exp: accurate to a few ulp of fp32 over the whole range
__device__ __forceinline__ float tanhf_(float v) {
  float a = fabsf(v);
  float e = __expf(-2.0f * a);
  float r = (1.0f - e) / (1.0f + e);
  return copysignf(r, v);
}

// conv.cu
int launch_conv(const vqw_conv_desc& d, float* out, cudaStream_t stream);
int launch_wgrad(const vqw_wgrad_desc& d, float* gw, float* gb, cudaStream_t stream);
// resblock_tc.cu (tcgen05 path)
bool resnet_tc_supported(const vqw_resnet_desc& d);
int64_t resnet_tc_workspace(const vqw_resnet_desc& d);
int resnet_forward_tc(const vqw_resnet_desc& d, const float* x, const float* cond,
                      const vqw_resblock_weights* weights, float* const* residuals, float* skip,
                      float* const* gate_tanh, float* const* gate_sig, void* workspace,
                      void* saved, cudaStream_t stream);
int64_t resnet_tc_saved_bytes(const vqw_resnet_desc& d);

// tc_gemm.cu (tcgen05 backward)
int64_t resnet_backward_tc_workspace(const vqw_resnet_desc& d);
int resnet_backward_tc(const vqw_resnet_desc& d, const float* g_skip, const float* g_last_res,
                       float* const* gate_tanh, float* const* gate_sig,
                       const vqw_resblock_weights* weights, float* gx0, float* gcond,
                       const vqw_resblock_wgrads* wgrads, void* workspace, const void* saved,
                       cudaStream_t stream);

bool head_tc_supported(const vqw_head_desc& d);
int64_t head_tc_workspace(const vqw_head_desc& d);
int64_t head_tc_saved_bytes(const vqw_head_desc& d);
int head_forward_tc(const vqw_head_desc& d, const float* skip, const float* W1, const float* b1,
                    const float* W2, const float* b2, float* y, void* workspace, void* saved,
                    cudaStream_t stream);
int head_loss_forward_tc(const vqw_head_desc& d, const float* skip, const float* W1, const float* b1,
                         const float* W2, const float* b2, const int32_t* tgt_i, const float* tgt_f,
                         int quantize, float log_scale_min, double* loss, float* y, void* workspace,
                         void* saved, cudaStream_t stream);
int head_backward_tc(const vqw_head_desc& d, const float* gy, const float* g_loss, const float* W1,
                     const float* W2, float* gskip, float* gW1, float* gb1, float* gW2, float* gb2,
                     void* workspace, const void* saved, cudaStream_t stream);

bool embed_bwd_tc_supported(int B, int T, int Cr, int Q);
int64_t embed_bwd_tc_workspace(int B, int T, int Cr, int Q);
int embed_backward_tc(const int32_t* q, const float* g, float* gW, float* gb, int B, int T, int Cr,
                      int Q, int mode, void* workspace, cudaStream_t stream);

}  // namespace vqw
