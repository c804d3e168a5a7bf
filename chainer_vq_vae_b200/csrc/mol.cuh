// Discretised mixture-of-logistics loss of ONE time step (WaveNet.calculate_logistic_loss,
// modules.py:169-230) and its gradient, shared by the stand-alone loss kernel (mol.cu) and the
// loss epilogue of the head GEMM (tc_gemm.cu).
#pragma once
#include "common.cuh"

namespace vqw {

__device__ __forceinline__ float sigmoid_chainer(float v) { return tanhf(v * 0.5f) * 0.5f + 0.5f; }
__device__ __forceinline__ float softplus_chainer(float v) {
  return fmaxf(v, 0.0f) + log1pf(expf(-fabsf(v)));
}

// log-probability of component k before the mixture weight (modules.py:181-226) and, optionally,
// its derivatives with respect to the mean and the (floored) log scale
__device__ __forceinline__ float mol_component(float x, float mean, float ls, float half, float lo,
                                               float hi, float* d_mean, float* d_ls) {
  const float c = x - mean;
  const float inv = expf(-ls);
  const float pin = inv * (c + half), min_ = inv * (c - half);
  const float cp = sigmoid_chainer(pin), cm = sigmoid_chainer(min_);
  float f;
  if (x < lo) {
    f = pin - softplus_chainer(pin);                      // log cdf_plus
    if (d_mean) { *d_mean = -inv * (1.0f - cp); *d_ls = -pin * (1.0f - cp); }
  } else if (x > hi) {
    f = -softplus_chainer(min_);                          // log (1 - cdf_min)
    if (d_mean) { *d_mean = inv * cm; *d_ls = min_ * cm; }
  } else {
    const float delta = cp - cm;
    f = logf(fmaxf(delta, 1e-12f));
    if (d_mean) {
      if (delta >= 1e-12f) {   // F.maximum routes the gradient to its first argument on >=
        const float sp = cp * (1.0f - cp), sm = cm * (1.0f - cm);
        *d_mean = -inv * (sp - sm) / delta;
        *d_ls = -(pin * sp - min_ * sm) / delta;
      } else {
        *d_mean = 0.0f;
        *d_ls = 0.0f;
      }
    }
  }
  return f;
}

// One position: y[k*stride] for k < 3 nr = {logit_probs, means, log_scales}, target value t in
// [-1, 1].  Returns -logsumexp_k(log_probs_k) (the position's term of modules.py:229, before the
// mean) and, if gy != nullptr, writes d(-lse)/dy * inv_n to gy[k*stride].
__device__ __forceinline__ float mol_position(const float* y, int64_t stride, float t, int nr,
                                              float half, float log_scale_min, float inv_n, float* gy,
                                              int64_t gstride) {
  const float x = 127.5f * t;
  const float lo = 127.5f * -0.999f, hi = 127.5f * 0.999f;
  // log_softmax(logit_probs)
  float mx = -INFINITY;
  for (int k = 0; k < nr; ++k) mx = fmaxf(mx, y[(int64_t)k * stride]);
  float s = 0.0f;
  for (int k = 0; k < nr; ++k) s += expf(y[(int64_t)k * stride] - mx);
  const float lse_l = mx + logf(s);
  // logsumexp_k(log_probs_k)
  float m2 = -INFINITY;
  for (int k = 0; k < nr; ++k) {
    const float ls = fmaxf(y[(int64_t)(2 * nr + k) * stride], log_scale_min);
    const float lp = mol_component(x, y[(int64_t)(nr + k) * stride], ls, half, lo, hi, nullptr, nullptr) +
                     (y[(int64_t)k * stride] - lse_l);
    m2 = fmaxf(m2, lp);
  }
  float s2 = 0.0f;
  for (int k = 0; k < nr; ++k) {
    const float ls = fmaxf(y[(int64_t)(2 * nr + k) * stride], log_scale_min);
    const float lp = mol_component(x, y[(int64_t)(nr + k) * stride], ls, half, lo, hi, nullptr, nullptr) +
                     (y[(int64_t)k * stride] - lse_l);
    s2 += expf(lp - m2);
  }
  const float lse = m2 + logf(s2);
  if (gy) {
    for (int k = 0; k < nr; ++k) {
      const float raw = y[(int64_t)(2 * nr + k) * stride];
      const float ls = fmaxf(raw, log_scale_min);
      const float lk = y[(int64_t)k * stride];
      float dm, dl;
      const float lp = mol_component(x, y[(int64_t)(nr + k) * stride], ls, half, lo, hi, &dm, &dl) +
                       (lk - lse_l);
      const float w = expf(lp - lse);            // posterior responsibility of component k
      const float pi = expf(lk - lse_l);         // prior mixture weight
      gy[(int64_t)k * gstride] = -(w - pi) * inv_n;
      gy[(int64_t)(nr + k) * gstride] = -w * dm * inv_n;
      gy[(int64_t)(2 * nr + k) * gstride] = (raw >= log_scale_min) ? -w * dl * inv_n : 0.0f;
    }
  }
  return -lse;
}

}  // namespace vqw
