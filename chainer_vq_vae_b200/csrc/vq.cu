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

// gW[c,:] = sum over columns n with idx[n]==c of gy[b,:,t]; one block per code, columns are
// visited in increasing n and accumulated in float64 (the reference's eye(k)[idx].T.dot(gy)
// is a float64 GEMM, utils.py:227-228), so the result is deterministic.  The index list is
// staged in shared memory once per chunk and scanned by the whole block: the test idx[n] == c is
// block-uniform, so the scan is a branch per column and the gather runs only on the ~N/k hits
// (round 1 compacted the hits with a single warp between two barriers: 0.19-0.36 ms per launch).
__global__ void __launch_bounds__(128)
vq_backward_w_kernel(const float* __restrict__ gy, const int32_t* __restrict__ idx,
                     float* __restrict__ gW, int B, int d, int T, int k) {
  const int c = blockIdx.x;
  const int64_t N = (int64_t)B * T;
  extern __shared__ int sidx[];   // one chunk of the index list
  const int CH = 4096;
  // every thread owns feature rows i = threadIdx.x, +blockDim.x, ... (d <= 8 * blockDim.x)
  double acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.0;
  for (int64_t n0 = 0; n0 < N; n0 += CH) {
    const int cn = (int)((N - n0 < CH) ? (N - n0) : CH);
    __syncthreads();
    for (int j = threadIdx.x; j < cn; j += blockDim.x) sidx[j] = idx[n0 + j];
    __syncthreads();
    for (int j = 0; j < cn; ++j) {
      if (sidx[j] != c) continue;                     // block-uniform
      const int64_t n = n0 + j;
      const int b = (int)(n / T), t = (int)(n % T);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = threadIdx.x + r * blockDim.x;
        if (i < d) acc[r] += (double)__ldg(gy + ((int64_t)b * d + i) * T + t);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int i = threadIdx.x + r * blockDim.x;
    if (i < d) gW[(int64_t)c * d + i] = (float)acc[r];
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
  VQW_REQUIRE(d <= 8 * 128, "vqw_vq_backward_w: d=%d > 1024 unsupported", d);
  vq_backward_w_kernel<<<k, 128, 4096 * sizeof(int), (cudaStream_t)stream>>>(gy, idx, gW, B, d, T,
                                                                           k);
  VQW_CHECK_LAUNCH("vq_backward_w_kernel");
  return 0;
}
