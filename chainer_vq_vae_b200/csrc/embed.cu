// Causal embedding of mu-law indices (sm_100a, HBM-bound gather).
//
// Replaces the `embed` Convolution2D((2,1), pad=(1,0)) + [:, :, :length] of WaveNet.__call__
// (modules.py:127-128,151-152) when its input is a one-hot tensor: a (2,1) convolution over a
// one-hot signal is a 2-column gather, out[b,c,t] = bias[c] + W[c,q[t-1],0] + W[c,q[t],1].
// The 126 MB one-hot tensor of the reference (B,256,T,1 f32) is never materialised.
// Backward: gW[c,q,j] = sum over (b,t) with q[b,t-1+j] == q of g[b,c,t]  (a per-channel
// histogram), accumulated in shared memory per (channel tile, batch item) and flushed once.
#include "common.cuh"

namespace vqw {

constexpr int EG_CT = 8;   // channels per CTA in the backward histogram

__global__ void __launch_bounds__(256)
embed_gather_fwd_kernel(const int32_t* __restrict__ q, const float* __restrict__ W,
                        const float* __restrict__ bias, float* __restrict__ out, int B, int T,
                        int Cr, int Q) {
  // grid: (ceil(T/256), Cr/ctile, B); thread = time step, loops over a channel tile so the two
  // index loads are amortised; stores are coalesced along T.
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 16;
  if (t >= T) return;
  // an index outside [0, Q) is an all-zero input column (generate.py:51 starts from one)
  const int q1 = q[(int64_t)b * T + t];
  const int q0 = (t > 0) ? q[(int64_t)b * T + t - 1] : -1;
  const bool ok0 = q0 >= 0 && q0 < Q, ok1 = q1 >= 0 && q1 < Q;
#pragma unroll 4
  for (int c = c0; c < min(c0 + 16, Cr); ++c) {
    float v = bias ? __ldg(bias + c) : 0.0f;
    const float* wr = W + (int64_t)c * Q * 2;
    if (ok0) v += __ldg(wr + 2 * q0);
    if (ok1) v += __ldg(wr + 2 * q1 + 1);
    out[((int64_t)b * Cr + c) * T + t] = v;
  }
}

__global__ void __launch_bounds__(256)
embed_gather_bwd_kernel(const int32_t* __restrict__ q, const float* __restrict__ g,
                        float* __restrict__ gW, float* __restrict__ gb, int B, int T, int Cr,
                        int Q) {
  extern __shared__ float hist[];   // [EG_CT][Q][2] + [EG_CT] bias sums
  float* bsum = hist + EG_CT * Q * 2;
  const int c0 = blockIdx.x * EG_CT, b = blockIdx.y;
  for (int i = threadIdx.x; i < EG_CT * Q * 2 + EG_CT; i += blockDim.x) hist[i] = 0.0f;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int q1 = q[(int64_t)b * T + t];
    const int q0 = (t > 0) ? q[(int64_t)b * T + t - 1] : -1;
    const bool ok0 = q0 >= 0 && q0 < Q, ok1 = q1 >= 0 && q1 < Q;
#pragma unroll
    for (int cc = 0; cc < EG_CT; ++cc) {
      int c = c0 + cc;
      if (c >= Cr) break;
      float gv = __ldg(g + ((int64_t)b * Cr + c) * T + t);
      if (ok0) atomicAdd(hist + (cc * Q + q0) * 2, gv);
      if (ok1) atomicAdd(hist + (cc * Q + q1) * 2 + 1, gv);
      atomicAdd(bsum + cc, gv);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < EG_CT * Q * 2; i += blockDim.x) {
    int cc = i / (Q * 2);
    if (c0 + cc < Cr && hist[i] != 0.0f) atomicAdd(gW + (int64_t)c0 * Q * 2 + i, hist[i]);
  }
  if (gb && threadIdx.x < EG_CT && c0 + threadIdx.x < Cr)
    atomicAdd(gb + c0 + threadIdx.x, bsum[threadIdx.x]);
}

}  // namespace vqw

extern "C" int vqw_embed_gather_forward(const int32_t* q, const float* W, const float* bias,
                                        float* out, int B, int T, int Cr, int Q,
                                        vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(q && W && out, "vqw_embed_gather_forward: null pointer");
  VQW_REQUIRE(B >= 0 && T >= 0 && Cr > 0 && Q > 0, "vqw_embed_gather_forward: bad sizes");
  if (B == 0 || T == 0) return 0;
  VQW_REQUIRE(B <= 65535, "vqw_embed_gather_forward: B > 65535");
  dim3 grid(ceil_div(T, 256), ceil_div(Cr, 16), B);
  embed_gather_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q, W, bias, out, B, T, Cr, Q);
  VQW_CHECK_LAUNCH("embed_gather_fwd_kernel");
  return 0;
}

extern "C" int vqw_embed_gather_backward(const int32_t* q, const float* gout, float* gW, float* gb,
                                         int B, int T, int Cr, int Q, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(q && gout && gW, "vqw_embed_gather_backward: null pointer");
  VQW_REQUIRE(B >= 0 && T >= 0 && Cr > 0 && Q > 0, "vqw_embed_gather_backward: bad sizes");
  if (B == 0 || T == 0) return 0;
  VQW_REQUIRE(B <= 65535, "vqw_embed_gather_backward: B > 65535");
  size_t smem = sizeof(float) * ((size_t)EG_CT * Q * 2 + EG_CT);
  VQW_REQUIRE(smem <= 200 * 1024, "vqw_embed_gather_backward: quantize=%d too large", Q);
  VQW_CHECK_CUDA(cudaFuncSetAttribute(embed_gather_bwd_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(Cr, EG_CT), B);
  embed_gather_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(q, gout, gW, gb, B, T, Cr, Q);
  VQW_CHECK_LAUNCH("embed_gather_bwd_kernel");
  return 0;
}

// tensor-core variant (see tc_gemm.cu): mode = VQW_MODE_BF16X3 / VQW_MODE_BF16
extern "C" int64_t vqw_embed_gather_backward_tc_workspace(int B, int T, int Cr, int Q) {
  return vqw::embed_bwd_tc_supported(B, T, Cr, Q) ? vqw::embed_bwd_tc_workspace(B, T, Cr, Q) : -1;
}
extern "C" int vqw_embed_gather_backward_tc(const int32_t* q, const float* gout, float* gW, float* gb,
                                            int B, int T, int Cr, int Q, int mode, void* workspace,
                                            vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(vqw_mode_tc(mode),
              "vqw_embed_gather_backward_tc: tensor-core modes only");
  return embed_backward_tc(q, gout, gW, gb, B, T, Cr, Q, mode, workspace, (cudaStream_t)stream);
}
