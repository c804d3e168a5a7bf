// Discretised mixture-of-logistics loss and the ConditionEmbed upsampling, one pass each.
//
// vqw_mol_loss       WaveNet.calculate_logistic_loss, modules.py:169-230, term by term in fp32
//                    (log_scales floor :178-179, +-half-bin sigmoids :185-188, the three branches
//                    at +-0.999*127.5 :198-226 with the log(max(cdf_delta, 1e-12)) floor :214-215,
//                    + log_softmax(logit_probs) :228, -mean(logsumexp) :229) and its gradient
//                    with respect to y (what Chainer's autograd derives from those ~20 functions).
// vqw_upsample_concat  ConditionEmbed.__call__'s tail, net.py:58-63: F.resize_images of the local
//                    embedding to 64x its length (1-D align-corners linear interpolation, exact
//                    integer coordinates), the speaker embedding broadcast over time (a resize
//                    from length 1), and the channel concat (local first); and its backward.
#include "mol.cuh"

namespace vqw {

// thread = one (b, t) position (coalesced along T); sweeps over the nr mixture components
__global__ void __launch_bounds__(256)
mol_loss_kernel(const float* __restrict__ y, const float* __restrict__ tgt, float* __restrict__ gy,
                double* __restrict__ loss, int B, int nr, int T, float half, float log_scale_min,
                float inv_n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  double local = 0.0;
  if (t < T) {
    const int64_t off = (int64_t)b * 3 * nr * T + t;
    local = (double)mol_position(y + off, T, tgt[(int64_t)b * T + t], nr, half, log_scale_min, inv_n,
                                 gy ? gy + off : nullptr, T);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(loss, local * (double)inv_n);
}

// out (B, Cl + Cg, T_out): channels [0, Cl) = align-corners linear resize of local (B, Cl, H),
// channels [Cl, Cl + Cg) = glob[b, c] for every t.  Coordinates in exact integer arithmetic:
// v = i (H-1) / (T_out-1), v0 = min(floor v, H-2), frac = v - v0  (Chainer computes the same v in
// float64; fp32 coordinates would drift by 9e-5 at T_out = 24000).
__global__ void __launch_bounds__(256)
upsample_concat_fwd_kernel(const float* __restrict__ local, const float* __restrict__ glob,
                           float* __restrict__ out, int Cl, int Cg, int H, int T_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (t >= T_out) return;
  float v;
  if (c < Cl) {
    const float* row = local + ((int64_t)b * Cl + c) * H;
    if (H == 1) {
      v = row[0];
    } else {
      const long long num = (long long)t * (H - 1), den = T_out > 1 ? T_out - 1 : 1;
      long long v0 = num / den;
      if (v0 > H - 2) v0 = H - 2;
      const float frac = (float)((double)(num - v0 * den) / (double)den);
      v = (1.0f - frac) * __ldg(row + v0) + frac * __ldg(row + v0 + 1);
    }
  } else {
    v = __ldg(glob + (int64_t)b * Cg + (c - Cl));
  }
  out[((int64_t)b * (Cl + Cg) + c) * T_out + t] = v;
}

// backward: g_local[b, c, h] = sum_t w(t, h) g[b, c, t];  g_glob[b, c] = sum_t g[b, Cl + c, t].
// One warp per output element; the output steps that touch frame h lie in a window of about
// 2 (T_out-1)/(H-1) steps around it.
__global__ void __launch_bounds__(256)
upsample_concat_bwd_kernel(const float* __restrict__ g, float* __restrict__ g_local,
                           float* __restrict__ g_glob, int B, int Cl, int Cg, int H, int T_out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_local = (int64_t)B * Cl * H, n_all = n_local + (int64_t)B * Cg;
  if (wid >= n_all) return;
  float acc = 0.0f;
  if (wid < n_local) {
    const int h = (int)(wid % H);
    const int64_t bc = wid / H;
    const int b = (int)(bc / Cl), c = (int)(bc % Cl);
    const float* row = g + ((int64_t)b * (Cl + Cg) + c) * T_out;
    if (H == 1) {
      for (int t = lane; t < T_out; t += 32) acc += __ldg(row + t);
    } else {
      const long long den = T_out > 1 ? T_out - 1 : 1;
      // steps with v in (h-1, h+1): t in ((h-1) den / (H-1), (h+1) den / (H-1))
      long long t_lo = ((long long)(h - 1) * den) / (H - 1) - 1, t_hi = ((long long)(h + 1) * den) / (H - 1) + 1;
      if (t_lo < 0) t_lo = 0;
      if (t_hi > T_out - 1) t_hi = T_out - 1;
      for (long long t = t_lo + lane; t <= t_hi; t += 32) {
        const long long num = t * (H - 1);
        long long v0 = num / den;
        if (v0 > H - 2) v0 = H - 2;
        const float frac = (float)((double)(num - v0 * den) / (double)den);
        float w = 0.0f;
        if (v0 == h) w = 1.0f - frac;
        else if (v0 + 1 == h) w = frac;
        if (w != 0.0f) acc = fmaf(w, __ldg(row + t), acc);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) g_local[wid] = acc;
  } else {
    const int64_t e = wid - n_local;
    const int b = (int)(e / Cg), c = (int)(e % Cg);
    const float* row = g + ((int64_t)b * (Cl + Cg) + Cl + c) * T_out;
    for (int t = lane; t < T_out; t += 32) acc += __ldg(row + t);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) g_glob[e] = acc;
  }
}

}  // namespace vqw

extern "C" int vqw_mol_loss(const float* y, const float* t, float* gy, double* loss, int B, int n_mix,
                            int T, int quantize, float log_scale_min, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(B >= 0 && n_mix > 0 && T >= 0 && quantize > 1, "vqw_mol_loss: bad sizes");
  if (B == 0 || T == 0) return 0;
  VQW_REQUIRE(y && t && loss, "vqw_mol_loss: null pointer");
  VQW_REQUIRE(B <= 65535, "vqw_mol_loss: B > 65535");
  dim3 grid(ceil_div(T, 256), B);
  mol_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      y, t, gy, loss, B, n_mix, T, (float)(127.5 / (quantize - 1)), log_scale_min,
      1.0f / ((float)B * (float)T));
  VQW_CHECK_LAUNCH("mol_loss_kernel");
  return 0;
}

extern "C" int vqw_upsample_concat_forward(const float* local, const float* glob, float* out, int B,
                                           int Cl, int Cg, int H, int T_out, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(B >= 0 && Cl >= 0 && Cg >= 0 && H >= 1 && T_out >= 0, "vqw_upsample_concat_forward: bad sizes");
  if (B == 0 || T_out == 0 || Cl + Cg == 0) return 0;
  VQW_REQUIRE(out && (Cl == 0 || local) && (Cg == 0 || glob), "vqw_upsample_concat_forward: null pointer");
  VQW_REQUIRE(B <= 65535 && Cl + Cg <= 65535, "vqw_upsample_concat_forward: B or channels > 65535");
  dim3 grid(ceil_div(T_out, 256), Cl + Cg, B);
  upsample_concat_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(local, glob, out, Cl, Cg, H, T_out);
  VQW_CHECK_LAUNCH("upsample_concat_fwd_kernel");
  return 0;
}

extern "C" int vqw_upsample_concat_backward(const float* g, float* g_local, float* g_glob, int B, int Cl,
                                            int Cg, int H, int T_out, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(B >= 0 && Cl >= 0 && Cg >= 0 && H >= 1 && T_out >= 0, "vqw_upsample_concat_backward: bad sizes");
  if (B == 0 || Cl + Cg == 0) return 0;
  VQW_REQUIRE(g && (Cl == 0 || g_local) && (Cg == 0 || g_glob), "vqw_upsample_concat_backward: null pointer");
  const int64_t warps = (int64_t)B * Cl * H + (int64_t)B * Cg;
  const int64_t blocks = (warps * 32 + 255) / 256;
  VQW_REQUIRE(blocks <= 0x7fffffffLL, "vqw_upsample_concat_backward: too many elements");
  upsample_concat_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, g_local, g_glob, B, Cl,
                                                                               Cg, H, T_out);
  VQW_CHECK_LAUNCH("upsample_concat_bwd_kernel");
  return 0;
}
