// Optimiser-side elementwise kernels over flat fp32 ranges (HBM bound, one pass each).
//
// vqw_adam_step   chainer.optimizers.Adam's update rule (train.py:101) [dep]:
//                   m += (1-b1)(g-m);  v += (1-b2)(g*g-v);  p -= lr * m / (sqrt(v) + eps)
//                 with lr = alpha*sqrt(1-b2^t)/(1-b1^t) computed by the caller (eps is NOT bias
//                 corrected).  Replaces ~250 per-parameter update launches.
// vqw_ema_update  ExponentialMovingAverage.__call__, utils.py:153-154:
//                   ema = decay*target + (1-decay)*ema   (decay multiplies the TARGET)
// vqw_softmax_ce  chainer.functions.softmax_cross_entropy (train.py:95): loss and d loss / d y
//                 in one pass over the logits.
#include "common.cuh"

namespace vqw {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t n, float lr_value, const float* __restrict__ lr_ptr,
            float omb1, float omb2, float eps) {
  // omb1 / omb2 = float(1 - beta) with the subtraction done in DOUBLE on the host: NumPy (the
  // reference's arithmetic) rounds the Python-float scalar 1 - beta2 = 0.001 to float32 once;
  // 1.0f - 0.999f would be 1.3e-5 off (caught by tests/test_gpu_optim.py)
  // lr from device memory when the step is replayed from a CUDA graph (it changes every step)
  const float lr = lr_ptr ? __ldg(lr_ptr) : lr_value;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      M[k] += omb1 * (G[k] - M[k]);
      V[k] += omb2 * (G[k] * G[k] - V[k]);
      P[k] -= lr * M[k] / (sqrtf(V[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    m[i] += omb1 * (g[i] - m[i]);
    v[i] += omb2 * (g[i] * g[i] - v[i]);
    p[i] -= lr * m[i] / (sqrtf(v[i]) + eps);
  }
}

__global__ void __launch_bounds__(256)
ema_kernel(float* __restrict__ ema, const float* __restrict__ target, int64_t n, float decay,
           float omd) {   // omd = float(1 - decay), subtraction in double (see adam_kernel)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    ema[i] = decay * target[i] + omd * ema[i];
}

// labels outside [0, Q) are ignored (Chainer's ignore_label = -1): they add nothing to the loss,
// get a zero gradient row and do not count in the normalisation (normalize=True)
__global__ void __launch_bounds__(256)
count_valid_kernel(const int32_t* __restrict__ tgt, int64_t n, int Q, double* __restrict__ count) {
  int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k = tgt[i];
    c += (k >= 0 && k < Q) ? 1 : 0;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (double)c);
}

// thread = one time step (coalesced along T); loops over the Q classes three times
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ y, const int32_t* __restrict__ tgt,
                  float* __restrict__ gy, double* __restrict__ loss, int B, int Q, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const double nvalid = loss[1];
  const float inv_n = nvalid > 0.0 ? (float)(1.0 / nvalid) : 0.0f;
  double local = 0.0;
  if (t < T) {
    const float* yc = y + (int64_t)b * Q * T + t;
    const int k = tgt[(int64_t)b * T + t];
    const bool valid = k >= 0 && k < Q;
    if (valid) {
      float mx = -INFINITY;
      for (int q = 0; q < Q; ++q) mx = fmaxf(mx, __ldg(yc + (int64_t)q * T));
      float s = 0.0f;
      for (int q = 0; q < Q; ++q) s += expf(__ldg(yc + (int64_t)q * T) - mx);
      const float lse = mx + logf(s);
      local = (double)(lse - __ldg(yc + (int64_t)k * T));
      if (gy) {
        float* gc = gy + (int64_t)b * Q * T + t;
        for (int q = 0; q < Q; ++q) {
          const float pq = expf(__ldg(yc + (int64_t)q * T) - lse);
          gc[(int64_t)q * T] = (pq - (q == k ? 1.0f : 0.0f)) * inv_n;
        }
      }
    } else if (gy) {
      float* gc = gy + (int64_t)b * Q * T + t;
      for (int q = 0; q < Q; ++q) gc[(int64_t)q * T] = 0.0f;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(loss, local * (double)inv_n);
}

}  // namespace vqw

extern "C" int vqw_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr,
                             double beta1, double beta2, double eps, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(n >= 0, "vqw_adam_step: negative size");
  if (n == 0) return 0;
  VQW_REQUIRE(p && g && m && v, "vqw_adam_step: null pointer");
  VQW_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                  ((uintptr_t)v % 16 == 0), "vqw_adam_step: buffers must be 16-byte aligned");
  adam_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr, nullptr,
                                                         (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                         (float)eps);
  VQW_CHECK_LAUNCH("adam_kernel");
  return 0;
}

extern "C" int vqw_adam_step_dev(float* p, const float* g, float* m, float* v, long long n,
                                 const float* lr_dev, double beta1, double beta2, double eps,
                                 vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(n >= 0, "vqw_adam_step_dev: negative size");
  if (n == 0) return 0;
  VQW_REQUIRE(p && g && m && v && lr_dev, "vqw_adam_step_dev: null pointer");
  VQW_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                  ((uintptr_t)v % 16 == 0), "vqw_adam_step_dev: buffers must be 16-byte aligned");
  adam_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, 0.0f, lr_dev,
                                                         (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                         (float)eps);
  VQW_CHECK_LAUNCH("adam_kernel");
  return 0;
}

extern "C" int vqw_ema_update(float* ema, const float* target, long long n, double decay,
                              vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(n >= 0, "vqw_ema_update: negative size");
  if (n == 0) return 0;
  VQW_REQUIRE(ema && target, "vqw_ema_update: null pointer");
  ema_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(ema, target, n, (float)decay,
                                                        (float)(1.0 - decay));
  VQW_CHECK_LAUNCH("ema_kernel");
  return 0;
}

extern "C" int vqw_softmax_ce(const float* y, const int32_t* t, float* gy, double* loss, int B, int Q,
                              int T, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(B >= 0 && Q > 0 && T >= 0, "vqw_softmax_ce: bad sizes");
  if (B == 0 || T == 0) return 0;
  VQW_REQUIRE(y && t && loss, "vqw_softmax_ce: null pointer");
  VQW_REQUIRE(B <= 65535, "vqw_softmax_ce: B > 65535");
  count_valid_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(t, (int64_t)B * T, Q, loss + 1);
  VQW_CHECK_LAUNCH("count_valid_kernel");
  dim3 grid(ceil_div(T, 256), B);
  softmax_ce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, t, gy, loss, B, Q, T);
  VQW_CHECK_LAUNCH("softmax_ce_kernel");
  return 0;
}
