// Generic strided / dilated 1-D convolution family in fp32 on CUDA cores (sm_100a).
//
// Serves every convolution of the path that is NOT the fused residual block: the Encoder's
// six stride-2 convs (net.py:12-26), ConditionEmbed's five dilated convs (net.py:34-53), the
// WaveNet embed/proj1/proj2 convs (modules.py:127-141,151-159), and the unfused pieces of the
// residual-block backward.  One kernel computes
//     out[b,m,t] = post(bias[m] + sum_s sum_k w_s[m,k] * pre_s(in_s[b,k,ti_s(t)]))
// as a tiled SGEMM (64 output channels x 64 time steps per CTA, 4x4 register tile, K staged
// through shared memory in chunks of 16) with the activation / mask / residual / gate-backward
// epilogues fused; a second kernel computes the weight and bias gradients of the same family
// as a split-K GEMM over (b,t) with atomic accumulation.
#include "common.cuh"

namespace vqw {

constexpr int BM = 64, BT = 64, KC = 32, NT = 256;
constexpr int WPITCH = BM + 4;

__device__ __forceinline__ float conv_fetch(const vqw_conv_src& s, int b, int k, int t) {
  if (k >= s.K) return 0.0f;
  int num = t * s.mul + s.shift;
  if (num < 0) return 0.0f;
  int ti = num;
  if (s.div > 1) {
    ti = num / s.div;
    if (ti * s.div != num) return 0.0f;
  }
  if (ti >= s.Tin) return 0.0f;
  int64_t off = ((int64_t)b * s.K + k) * s.Tin + ti;
  float v = __ldg(s.in + off);
  if (s.relu_in) v = fmaxf(v, 0.0f);
  if (s.in_mask) v = (__ldg(s.in_mask + off) > 0.0f) ? v : 0.0f;
  return v;
}

__global__ void __launch_bounds__(NT)
conv_sum_kernel(const __grid_constant__ vqw_conv_desc D, float* __restrict__ out) {
  // K chunks of KC = 32 channels, all sources flattened into one chunk sequence; the operands of
  // chunk c+1 are fetched into registers while chunk c is multiplied out of shared memory (round
  // 2: the small encoder / condition layers were bound by the load -> barrier -> FMA -> barrier
  // latency chain of 16-channel chunks: 40-110 us per launch for ~1 us of arithmetic)
  __shared__ __align__(16) float Xs[KC][BT];
  __shared__ __align__(16) float Ws[KC][WPITCH];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * BT, m0 = blockIdx.y * BM, b = blockIdx.z;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  constexpr int XR = (KC * BT) / NT, WR = (KC * BM) / NT;
  float xr[XR], wr[WR];
  int chunk_src[VQW_MAX_SRC + 1];
  int nchunks = 0;
  for (int si = 0; si < D.nsrc; ++si) {
    chunk_src[si] = nchunks;
    nchunks += (D.src[si].K + KC - 1) / KC;
  }
  chunk_src[D.nsrc] = nchunks;
  auto fetch = [&](int c) {
    int si = 0;
    while (c >= chunk_src[si + 1]) ++si;
    const vqw_conv_src& s = D.src[si];
    const int k0 = (c - chunk_src[si]) * KC;
    const bool k_fast = s.wk <= s.wm;
#pragma unroll
    for (int r = 0; r < XR; ++r) {
      const int e = tid + r * NT;
      const int kc = e / BT, tt = e % BT;
      const int t = t0 + tt;
      xr[r] = (t < D.T) ? conv_fetch(s, b, k0 + kc, t) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const int e = tid + r * NT;
      int kc, mm;
      if (k_fast) { kc = e % KC; mm = e / KC; } else { mm = e % BM; kc = e / BM; }
      const int m = m0 + mm, k = k0 + kc;
      wr[r] = (m < D.M && k < s.K) ? __ldg(s.w + (int64_t)m * s.wm + (int64_t)k * s.wk) : 0.0f;
    }
  };
  auto stash = [&](int c) {
    int si = 0;
    while (c >= chunk_src[si + 1]) ++si;
    const bool k_fast = D.src[si].wk <= D.src[si].wm;
#pragma unroll
    for (int r = 0; r < XR; ++r) {
      const int e = tid + r * NT;
      Xs[e / BT][e % BT] = xr[r];
    }
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const int e = tid + r * NT;
      int kc, mm;
      if (k_fast) { kc = e % KC; mm = e / KC; } else { mm = e % BM; kc = e / BM; }
      Ws[kc][mm] = wr[r];
    }
  };
  if (nchunks > 0) fetch(0);
  for (int c = 0; c < nchunks; ++c) {
    __syncthreads();                 // the previous chunk has been multiplied out
    stash(c);
    __syncthreads();
    if (c + 1 < nchunks) fetch(c + 1);   // in flight during the FMAs below
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      float4 xv = *reinterpret_cast<const float4*>(&Xs[kc][tx * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[kc][ty * 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], xa[j], acc[i][j]);
    }
  }

  const bool gate = D.gate_tanh != nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= D.M) continue;
    float bias = D.bias ? __ldg(D.bias + m) : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int t = t0 + tx * 4 + j;
      if (t >= D.T) continue;
      float v = acc[i][j] + bias;
      int64_t off = ((int64_t)b * D.M + m) * D.T + t;
      if (gate) {
        float th = __ldg(D.gate_tanh + off), sg = __ldg(D.gate_sig + off);
        int64_t o1 = ((int64_t)b * 2 * D.M + m) * D.T + t;
        int64_t o2 = ((int64_t)b * 2 * D.M + D.M + m) * D.T + t;
        out[o1] = v * sg * (1.0f - th * th);
        out[o2] = v * th * sg * (1.0f - sg);
      } else {
        if (D.addend) v += __ldg(D.addend + off);
        if (D.relu_out) v = fmaxf(v, 0.0f);
        if (D.out_mask) v = (__ldg(D.out_mask + off) > 0.0f) ? v : 0.0f;
        if (D.accumulate) v += out[off];
        out[off] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// weight / bias gradient: gw[m,k] += sum_{b,t} A[b,m,t] * Bv[b,k,ti(t)]
// ---------------------------------------------------------------------------------------
constexpr int GK = 64, TC = 32;
constexpr int GPITCH = 68;

__global__ void __launch_bounds__(NT)
conv_wgrad_kernel(const __grid_constant__ vqw_wgrad_desc D, float* __restrict__ gw,
                  float* __restrict__ gb, int chunks_per_b, int total_chunks, int ktiles) {
  // blockIdx.y = tap * ktiles + k-tile: every tap of the filter in one launch; the next chunk's
  // operands are fetched into registers while the current one is multiplied out of shared memory
  __shared__ __align__(16) float As[TC][GPITCH];
  __shared__ __align__(16) float Bs[TC][GPITCH];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tap = blockIdx.y / ktiles;
  const int m0 = blockIdx.x * BM, k0 = (blockIdx.y - tap * ktiles) * GK;
  const int shift = D.shift + tap * D.tap_dshift;
  gw += (int64_t)tap * D.tap_gw;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bsum = 0.0f;
  const bool do_bias = (gb != nullptr) && (blockIdx.y == 0) && (tid < BM);
  constexpr int AR = (TC * BM) / NT, BR = (TC * GK) / NT;
  float ar[AR], br[BR];
  auto fetch = [&](int c) {
    const int b = c / chunks_per_b;
    const int tbase = (c % chunks_per_b) * TC;
#pragma unroll
    for (int r = 0; r < AR; ++r) {
      const int e = tid + r * NT;
      const int tt = e % TC, mm = e / TC;
      const int m = m0 + mm, t = tbase + tt;
      float v = 0.0f;
      if (m < D.M && t < D.T) {
        const int64_t off = ((int64_t)b * D.M + m) * D.T + t;
        v = __ldg(D.a + off);
        if (D.a_mask) v = (__ldg(D.a_mask + off) > 0.0f) ? v : 0.0f;
      }
      ar[r] = v;
    }
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const int e = tid + r * NT;
      const int tt = e % TC, kk = e / TC;
      const int k = k0 + kk, t = tbase + tt;
      float v = 0.0f;
      if (k < D.K && t < D.T) {
        const int num = t * D.mul + shift;
        int ti = num;
        bool ok = num >= 0;
        if (ok && D.div > 1) { ti = num / D.div; ok = (ti * D.div == num); }
        if (ok && ti < D.Tin) {
          const int64_t off = ((int64_t)b * D.K + k) * D.Tin + ti;
          v = __ldg(D.in + off);
          if (D.relu_in) v = fmaxf(v, 0.0f);
          if (D.in_mul) v *= __ldg(D.in_mul + off);
        }
      }
      br[r] = v;
    }
  };
  int c = blockIdx.z;
  if (c < total_chunks) fetch(c);
  for (; c < total_chunks; c += gridDim.z) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < AR; ++r) {
      const int e = tid + r * NT;
      As[e % TC][e / TC] = ar[r];
    }
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const int e = tid + r * NT;
      Bs[e % TC][e / TC] = br[r];
    }
    __syncthreads();
    if (c + (int)gridDim.z < total_chunks) fetch(c + gridDim.z);
#pragma unroll
    for (int tt = 0; tt < TC; ++tt) {
      float4 av = *reinterpret_cast<const float4*>(&As[tt][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[tt][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float ba[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], ba[j], acc[i][j]);
    }
    if (do_bias) {
#pragma unroll
      for (int tt = 0; tt < TC; ++tt) bsum += As[tt][tid];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= D.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k >= D.K) continue;
      atomicAdd(gw + (int64_t)m * D.gm + (int64_t)k * D.gk, acc[i][j]);
    }
  }
  if (do_bias && (m0 + tid) < D.M) atomicAdd(gb + m0 + tid, bsum);
}

// Weight gradient of a convolution with ONE input channel (the embed layer of the mixture-of-
// logistics decoder reads the raw waveform, modules.py:127-128 with input_dim = 1): gw[m, j] =
// sum_{b,t} a[b,m,t] * x[b, ti_j(t)] is a reduction, not a GEMM -- CTA = (output channel, slice of
// the (b,t) axis), coalesced rows of a, the taps from L1/L2, one atomic per tap and CTA.
constexpr int K1_MAXTAPS = 8;
__global__ void __launch_bounds__(256)
conv_wgrad_k1_kernel(const __grid_constant__ vqw_wgrad_desc D, float* __restrict__ gw,
                     float* __restrict__ gb, int tparts) {
  // blockIdx.y = item * tparts + slice of the time axis: no division in the loop, 4 rows of loads
  // in flight per thread
  const int m = blockIdx.x;
  const int ntaps = D.ntaps > 1 ? D.ntaps : 1;
  const int b = blockIdx.y / tparts, tp = blockIdx.y - b * tparts;
  const int per = (D.T + tparts - 1) / tparts;
  const int t0 = tp * per, t1 = (t0 + per < D.T) ? t0 + per : D.T;
  const float* __restrict__ arow = D.a + ((int64_t)b * D.M + m) * D.T;
  const float* __restrict__ mrow = D.a_mask ? D.a_mask + ((int64_t)b * D.M + m) * D.T : nullptr;
  const float* __restrict__ xrow = D.in + (int64_t)b * D.Tin;
  const float* __restrict__ xmul = D.in_mul ? D.in_mul + (int64_t)b * D.Tin : nullptr;
  float acc[K1_MAXTAPS], bsum = 0.0f;
#pragma unroll
  for (int j = 0; j < K1_MAXTAPS; ++j) acc[j] = 0.0f;
  for (int tb = t0 + threadIdx.x; tb < t1; tb += 4 * blockDim.x) {
    float av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = tb + u * blockDim.x;
      av[u] = (t < t1) ? __ldg(arow + t) : 0.0f;
      if (mrow && t < t1) av[u] = (__ldg(mrow + t) > 0.0f) ? av[u] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = tb + u * blockDim.x;
      if (t >= t1) break;
      const float a = av[u];
      bsum += a;
#pragma unroll
      for (int j = 0; j < K1_MAXTAPS; ++j) {
        if (j >= ntaps) break;
        const int num = t * D.mul + D.shift + j * D.tap_dshift;
        int ti = num;
        bool ok = num >= 0;
        if (ok && D.div > 1) { ti = num / D.div; ok = (ti * D.div == num); }
        if (ok && ti < D.Tin) {
          float v = __ldg(xrow + ti);
          if (D.relu_in) v = fmaxf(v, 0.0f);
          if (xmul) v *= __ldg(xmul + ti);
          acc[j] = fmaf(a, v, acc[j]);
        }
      }
    }
  }
  __shared__ float red[8][K1_MAXTAPS + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j <= K1_MAXTAPS; ++j) {
    float v = (j < K1_MAXTAPS) ? acc[j] : bsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x <= K1_MAXTAPS) {
    const int j = threadIdx.x;
    float v = 0.0f;
    for (int w = 0; w < 8; ++w) v += red[w][j];
    if (j < ntaps) atomicAdd(gw + (int64_t)j * D.tap_gw + (int64_t)m * D.gm, v);
    else if (j == K1_MAXTAPS && gb != nullptr) atomicAdd(gb + m, v);
  }
}

int launch_conv(const vqw_conv_desc& d, float* out, cudaStream_t stream) {
  VQW_REQUIRE(out != nullptr, "vqw_conv_forward: out is null");
  VQW_REQUIRE(d.B >= 0 && d.M > 0 && d.T >= 0, "vqw_conv_forward: bad sizes B=%d M=%d T=%d", d.B,
              d.M, d.T);
  VQW_REQUIRE(d.nsrc >= 1 && d.nsrc <= VQW_MAX_SRC, "vqw_conv_forward: nsrc=%d out of range",
              d.nsrc);
  for (int i = 0; i < d.nsrc; ++i) {
    const vqw_conv_src& s = d.src[i];
    VQW_REQUIRE(s.in && s.w, "vqw_conv_forward: source %d has a null pointer", i);
    VQW_REQUIRE(s.K > 0 && s.Tin >= 0 && s.div >= 1 && s.mul >= 1,
                "vqw_conv_forward: source %d bad K/Tin/div/mul", i);
  }
  VQW_REQUIRE((d.gate_tanh == nullptr) == (d.gate_sig == nullptr),
              "vqw_conv_forward: gate_tanh and gate_sig must be given together");
  if (d.B == 0 || d.T == 0) return 0;
  VQW_REQUIRE(d.B <= 65535, "vqw_conv_forward: B > 65535");
  dim3 grid(ceil_div(d.T, BT), ceil_div(d.M, BM), d.B);
  conv_sum_kernel<<<grid, NT, 0, stream>>>(d, out);
  VQW_CHECK_LAUNCH("conv_sum_kernel");
  return 0;
}

int launch_wgrad(const vqw_wgrad_desc& d, float* gw, float* gb, cudaStream_t stream) {
  VQW_REQUIRE(gw && d.a && d.in, "vqw_conv_wgrad: null pointer");
  VQW_REQUIRE(d.B >= 0 && d.M > 0 && d.T >= 0 && d.K > 0 && d.div >= 1 && d.mul >= 1,
              "vqw_conv_wgrad: bad sizes");
  if (d.B == 0 || d.T == 0) return 0;
  const int ntaps = d.ntaps > 1 ? d.ntaps : 1;
  if (d.K == 1 && ntaps <= K1_MAXTAPS && d.M <= 65535) {
    // about eight CTAs per SM; a slice is at least 1024 time steps (4 per thread)
    int tparts = ceil_div(148 * 8, d.M * d.B);
    if (tparts > ceil_div(d.T, 1024)) tparts = ceil_div(d.T, 1024);
    if (tparts < 1) tparts = 1;
    if ((int64_t)d.B * tparts > 65535) tparts = 65535 / d.B > 0 ? 65535 / d.B : 1;
    VQW_REQUIRE((int64_t)d.B * tparts <= 65535, "vqw_conv_wgrad: batch too large for the one-channel kernel");
    conv_wgrad_k1_kernel<<<dim3(d.M, d.B * tparts), 256, 0, stream>>>(d, gw, gb, tparts);
    VQW_CHECK_LAUNCH("conv_wgrad_k1_kernel");
    return 0;
  }
  VQW_REQUIRE(ntaps * ceil_div(d.K, GK) <= 65535, "vqw_conv_wgrad: too many taps / input channels");
  int chunks_per_b = ceil_div(d.T, TC);
  int total = chunks_per_b * d.B;
  const int ktiles = ceil_div(d.K, GK);
  int tiles = ceil_div(d.M, BM) * ktiles * ntaps;
  // enough CTAs to fill the chip about twice, no more: every CTA ends with a tile of atomics
  int splits = ceil_div(148 * 2, tiles);
  if (splits > total) splits = total;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  dim3 grid(ceil_div(d.M, BM), ktiles * ntaps, splits);
  conv_wgrad_kernel<<<grid, NT, 0, stream>>>(d, gw, gb, chunks_per_b, total, ktiles);
  VQW_CHECK_LAUNCH("conv_wgrad_kernel");
  return 0;
}

}  // namespace vqw

extern "C" int vqw_conv_forward(const vqw_conv_desc* desc, float* out, vqw_stream_t stream) {
  VQW_REQUIRE(desc != nullptr, "vqw_conv_forward: desc is null");
  return vqw::launch_conv(*desc, out, (cudaStream_t)stream);
}

extern "C" int vqw_conv_wgrad(const vqw_wgrad_desc* desc, float* gw, float* gb,
                              vqw_stream_t stream) {
  VQW_REQUIRE(desc != nullptr, "vqw_conv_wgrad: desc is null");
  return vqw::launch_wgrad(*desc, gw, gb, (cudaStream_t)stream);
}
