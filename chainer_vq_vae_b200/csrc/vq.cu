// VQ nearest-codebook lookup and its codebook gradient (sm_100a, CUDA cores: the op is
// latency/HBM-scale, 0.19 GFLOP and ~1.1 MB at the B200 config -- not tensor-core work).
//
// Replaces StraightThrough.forward / .backward of the reference, utils.py:176-231.
// Layout: z (B,d,T) f32 with T contiguous, W (k,d) f32.
//
// Forward kernel: one warp per (b,t) column, the 32 lanes split the codebook (lane l scans
// codes l, l+32, ...).  Each lane accumulates sum_i (z_i - W_ki)^2 SEQUENTIALLY over i in
// fp32 with explicit round-to-nearest sub/mul/add (no FMA contraction) -- bit-identical to
// NumPy's `sum((xs - W) ** 2, axis=2)` -- keeps its first minimum, then a warp-shuffle
// lexicographic (distance, index) min reduction reproduces argmin's first-occurrence rule.
// The codebook is staged in shared memory in tiles of KT codes with a (d+1) row pitch so the
// 32 lanes hit 32 different banks.
#include "common.cuh"

namespace vqw {

constexpr int VQ_WARPS = 8;

__global__ void __launch_bounds__(VQ_WARPS * 32)
vq_forward_kernel(const float* __restrict__ z, const float* __restrict__ W,
                  int32_t* __restrict__ idx, float* __restrict__ e, float* __restrict__ count,
                  float* __restrict__ zsum, double* __restrict__ sqerr, int B, int d, int T, int k,
                  int KT) {
  extern __shared__ float smem[];
  float* Ws = smem;                              // [KT][d+1]
  float* zs = smem + (size_t)KT * (d + 1);       // [VQ_WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t N = (int64_t)B * T;
  const int64_t n = (int64_t)blockIdx.x * VQ_WARPS + warp;
  const bool active = n < N;
  const int b = active ? (int)(n / T) : 0;
  const int t = active ? (int)(n % T) : 0;
  const float* zcol = z + ((int64_t)b * d) * T + t;
  if (active)
    for (int i = lane; i < d; i += 32) zs[warp * d + i] = zcol[(int64_t)i * T];

  float best = INFINITY;
  int best_k = 0x7fffffff;
  const float* myz = zs + warp * d;
  for (int k0 = 0; k0 < k; k0 += KT) {
    const int kt = min(KT, k - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < kt * d; j += blockDim.x) {
      int kk = j / d, i = j - kk * d;
      Ws[kk * (d + 1) + i] = W[(int64_t)(k0 + kk) * d + i];
    }
    __syncthreads();
    if (active) {
      for (int kk = lane; kk < kt; kk += 32) {
        const float* wr = Ws + kk * (d + 1);
        float acc = 0.0f;
        for (int i = 0; i < d; ++i) {
          float diff = __fsub_rn(myz[i], wr[i]);
          acc = __fadd_rn(acc, __fmul_rn(diff, diff));
        }
        // strict '<' keeps the first minimum; a NaN distance never wins (NumPy would
        // return the first NaN -- documented difference, inputs are finite)
        if (acc < best) { best = acc; best_k = k0 + kk; }
      }
    }
  }
  // warp lexicographic min (distance, then lowest index)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, off);
    int ok = __shfl_xor_sync(0xffffffffu, best_k, off);
    if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
  }
  if (!active) return;
  if (best_k == 0x7fffffff) best_k = 0;   // all-NaN column
  if (lane == 0) {
    idx[n] = best_k;
    if (count) atomicAdd(count + best_k, 1.0f);
  }
  double err = 0.0;
  for (int i = lane; i < d; i += 32) {
    float wv = W[(int64_t)best_k * d + i];
    float zv = myz[i];
    e[((int64_t)b * d + i) * T + t] = wv;
    float df = zv - wv;
    err += (double)df * (double)df;
    if (zsum) atomicAdd(zsum + (int64_t)best_k * d + i, zv);
  }
  if (sqerr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) err += __shfl_xor_sync(0xffffffffu, err, off);
    if (lane == 0) atomicAdd(sqerr, err);
  }
}

// gW[c,:] = sum over columns n with idx[n]==c of gy[b,:,t] in float64 (the reference's
// eye(k)[idx].T.dot(gy) is a float64 GEMM, utils.py:227-228).  One block per code:
//   1. the whole block compacts the matching columns into shared memory (order preserving:
//      ballot + prefix per warp, warps in column order);
//   2. warp w sums the hits h = w, w + 8, ... (features strided over the lanes, four independent
//      loads in flight), 3. the eight partial sums are added in warp order.
// Deterministic: the partition and the order of every addition are fixed.  (A random-init
// codebook sends most columns to a handful of codes -- hundreds of hits in one block; round 2's
// first version walked them one dependent load at a time: 0.32 ms.)
constexpr int VQB_WARPS = 8, VQB_CH = 2048;
__global__ void __launch_bounds__(VQB_WARPS * 32)
vq_backward_w_kernel(const float* __restrict__ gy, const int32_t* __restrict__ idx,
                     float* __restrict__ gW, int B, int d, int T, int k) {
  const int c = blockIdx.x;
  const int64_t N = (int64_t)B * T;
  __shared__ int hits[VQB_CH];
  __shared__ int wcount[VQB_WARPS];
  extern __shared__ double part[];          // [VQB_WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < VQB_WARPS * d; i += blockDim.x) part[i] = 0.0;
  for (int64_t n0 = 0; n0 < N; n0 += VQB_CH) {
    const int cn = (int)((N - n0 < VQB_CH) ? (N - n0) : VQB_CH);
    __syncthreads();
    // ---- compaction: warp w scans columns [w*seg, (w+1)*seg) of the chunk ----
    const int seg = (cn + VQB_WARPS - 1) / VQB_WARPS;
    const int s0 = warp * seg, s1 = min(cn, s0 + seg);
    int cnt = 0;
    for (int j0 = s0; j0 < s1; j0 += 32) {
      const int j = j0 + lane;
      const bool hit = j < s1 && idx[n0 + j] == c;
      cnt += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0) wcount[warp] = cnt;
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < VQB_WARPS; ++w) {
      if (w < warp) base += wcount[w];
      total += wcount[w];
    }
    for (int j0 = s0; j0 < s1; j0 += 32) {
      const int j = j0 + lane;
      const bool hit = j < s1 && idx[n0 + j] == c;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) hits[base + __popc(m & ((1u << lane) - 1u))] = j;
      base += __popc(m);
    }
    __syncthreads();
    // ---- warp w: hits w, w + 8, ...; lane: features lane, lane + 32, ... ----
    for (int i0 = lane; i0 < d; i0 += 32) {
      double acc = part[warp * d + i0];
      int h = warp;
      for (; h + 3 * VQB_WARPS < total; h += 4 * VQB_WARPS) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t n = n0 + hits[h + u * VQB_WARPS];
          const int bb = (int)(n / T), t = (int)(n % T);
          v[u] = __ldg(gy + ((int64_t)bb * d + i0) * T + t);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += (double)v[u];
      }
      for (; h < total; h += VQB_WARPS) {
        const int64_t n = n0 + hits[h];
        const int bb = (int)(n / T), t = (int)(n % T);
        acc += (double)__ldg(gy + ((int64_t)bb * d + i0) * T + t);
      }
      part[warp * d + i0] = acc;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    double acc = 0.0;
    for (int w = 0; w < VQB_WARPS; ++w) acc += part[w * d + i];
    gW[(int64_t)c * d + i] = (float)acc;
  }
}

}  // namespace vqw

extern "C" int vqw_vq_forward(const float* z, const float* W, int32_t* idx, float* e, float* count,
                              float* zsum, double* sqerr, int B, int d, int T, int k,
                              vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(B >= 0 && T >= 0 && d > 0 && k > 0, "vqw_vq_forward: bad sizes B=%d d=%d T=%d k=%d",
              B, d, T, k);
  if ((int64_t)B * T == 0) return 0;   // empty batch: nothing to do (pointers may be null)
  VQW_REQUIRE(z && W && idx && e, "vqw_vq_forward: null pointer");
  int kt_cap = (int)((96 * 1024) / (sizeof(float) * (d + 1)));
  int KT = (kt_cap / 32) * 32;
  VQW_REQUIRE(KT >= 32, "vqw_vq_forward: d=%d too large for the shared-memory codebook tile", d);
  int kround = ((k + 31) / 32) * 32;
  if (KT > kround) KT = kround;
  size_t smem = sizeof(float) * ((size_t)KT * (d + 1) + (size_t)VQ_WARPS * d);
  VQW_CHECK_CUDA(cudaFuncSetAttribute(vq_forward_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t N = (int64_t)B * T;
  int grid = (int)ceil_div64(N, VQ_WARPS);
  vq_forward_kernel<<<grid, VQ_WARPS * 32, smem, (cudaStream_t)stream>>>(z, W, idx, e, count, zsum,
                                                                        sqerr, B, d, T, k, KT);
  VQW_CHECK_LAUNCH("vq_forward_kernel");
  return 0;
}

extern "C" int vqw_vq_backward_w(const float* gy, const int32_t* idx, float* gW, int B, int d,
                                 int T, int k, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(B >= 0 && T >= 0 && d > 0 && k > 0, "vqw_vq_backward_w: bad sizes");
  VQW_REQUIRE(gW && (((int64_t)B * T == 0) || (gy && idx)), "vqw_vq_backward_w: null pointer");
  VQW_REQUIRE(d <= 2048, "vqw_vq_backward_w: d=%d > 2048 unsupported", d);
  const size_t smem = sizeof(double) * VQB_WARPS * d;
  if (smem > 40 * 1024)
    VQW_CHECK_CUDA(cudaFuncSetAttribute(vq_backward_w_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vq_backward_w_kernel<<<k, VQB_WARPS * 32, smem, (cudaStream_t)stream>>>(gy, idx, gW, B, d, T, k);
  VQW_CHECK_LAUNCH("vq_backward_w_kernel");
  return 0;
}
