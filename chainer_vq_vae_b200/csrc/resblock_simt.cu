// Fused WaveNet residual block, fp32 CUDA-core path (sm_100a).
//
// Replaces ResidualBlock.__call__ (modules.py:30-56) + the skip accumulation of
// ResidualNet.__call__ (modules.py:92-95) with ONE kernel per block:
//   h   = conv_b + cond_b + sum_j Wc[:,:,j] x[t - dil*(fs-1-j)] + Wp cond[t]      (:40-44)
//   z   = tanh(h[:Ch]) * sigmoid(h[Ch:])                                           (:47-48)
//   res = Wr z + br + x ;  skip (+)= Ws z + bs                                     (:51-55)
// This is the path for small channel counts (the CPU-parity config, 32 channels, where the
// contraction is far too small for a tensor-core tile) and the fp32 reference mode for any
// shape; the 512-channel training config runs the tcgen05 kernel in resblock_tc.cu.
//
// CTA = (one batch item, 32 consecutive time steps); lane = time step, so every global access
// is a coalesced 128-byte row segment of the (B,C,T) tensors.  Warp w owns gate pairs
// [w*PPW, (w+1)*PPW): the tanh row p and the sigmoid row Ch+p accumulate in registers of the
// same thread, so the gate needs no exchange.  Weights stream through shared memory in
// 8-input-channel slabs and are read as warp-broadcast float4.  z is parked in shared memory
// and the second contraction (res | skip rows) runs in passes of 256 rows.
#include "common.cuh"

namespace vqw {

constexpr int RB_WARPS = 8;
constexpr int RB_KC = 8;

template <int PPW>
__global__ void __launch_bounds__(RB_WARPS * 32)
resblock_fwd_simt_kernel(const __grid_constant__ vqw_resblock_desc D, const float* __restrict__ x,
                         const float* __restrict__ cond, const vqw_resblock_weights W,
                         float* __restrict__ residual, float* __restrict__ skip,
                         float* __restrict__ gate_tanh, float* __restrict__ gate_sig) {
  constexpr int NP = RB_WARPS * PPW;   // pairs covered by the CTA
  constexpr int PITCH = NP + 4;
  constexpr int RPW = 32;              // rows per warp per pass of the second contraction
  constexpr int NR = RB_WARPS * RPW;   // rows per pass
  constexpr int PITCH2 = NR + 4;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                    // [RB_KC][32]
  float* Wt = Xs + RB_KC * 32;         // [RB_KC][PITCH]  (also reused as W2s [RB_KC][PITCH2])
  float* Wg = Wt + RB_KC * PITCH;      // [RB_KC][PITCH]
  float* Zs = smem + RB_KC * 32 + RB_KC * ((2 * PITCH > PITCH2) ? 2 * PITCH : PITCH2);  // [NP][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, t0 = blockIdx.x * 32, t = t0 + lane;
  const int Ch = D.Cd / 2;
  const bool t_ok = t < D.T;

  float acc_t[PPW], acc_g[PPW];
#pragma unroll
  for (int i = 0; i < PPW; ++i) {
    int p = warp * PPW + i;
    if (p < Ch) {
      acc_t[i] = __ldg(W.conv_b + p) + __ldg(W.cond_b + p);
      acc_g[i] = __ldg(W.conv_b + Ch + p) + __ldg(W.cond_b + Ch + p);
    } else {
      acc_t[i] = 0.0f;
      acc_g[i] = 0.0f;
    }
  }

  // ---- first contraction: fs causal taps of x, then the 1x1 condition projection ----
  for (int s = 0; s <= D.fs; ++s) {
    const bool is_cond = (s == D.fs);
    const float* in = is_cond ? cond : x;
    const int K = is_cond ? D.Cc : D.Cr;
    const int shift = is_cond ? 0 : -D.dilation * (D.fs - 1 - s);
    const float* wbase = is_cond ? W.cond_w : (W.conv_w + s);
    const int wm = is_cond ? D.Cc : D.Cr * D.fs;
    const int wk = is_cond ? 1 : D.fs;
    for (int k0 = 0; k0 < K; k0 += RB_KC) {
      __syncthreads();
      {
        int k = k0 + warp, ti = t + shift;
        float v = 0.0f;
        if (k < K && ti >= 0 && ti < D.T) v = __ldg(in + ((int64_t)b * K + k) * D.T + ti);
        Xs[warp * 32 + lane] = v;
      }
      for (int e = tid; e < RB_KC * NP; e += RB_WARPS * 32) {
        int kc = e % RB_KC, p = e / RB_KC;
        int k = k0 + kc;
        float vt = 0.0f, vg = 0.0f;
        if (p < Ch && k < K) {
          vt = __ldg(wbase + (int64_t)p * wm + (int64_t)k * wk);
          vg = __ldg(wbase + (int64_t)(Ch + p) * wm + (int64_t)k * wk);
        }
        Wt[kc * PITCH + p] = vt;
        Wg[kc * PITCH + p] = vg;
      }
      __syncthreads();
#pragma unroll
      for (int kc = 0; kc < RB_KC; ++kc) {
        float xv = Xs[kc * 32 + lane];
        const float4* wt4 = reinterpret_cast<const float4*>(Wt + kc * PITCH + warp * PPW);
        const float4* wg4 = reinterpret_cast<const float4*>(Wg + kc * PITCH + warp * PPW);
#pragma unroll
        for (int i = 0; i < PPW / 4; ++i) {
          float4 a = wt4[i], g = wg4[i];
          acc_t[4 * i + 0] = fmaf(a.x, xv, acc_t[4 * i + 0]);
          acc_t[4 * i + 1] = fmaf(a.y, xv, acc_t[4 * i + 1]);
          acc_t[4 * i + 2] = fmaf(a.z, xv, acc_t[4 * i + 2]);
          acc_t[4 * i + 3] = fmaf(a.w, xv, acc_t[4 * i + 3]);
          acc_g[4 * i + 0] = fmaf(g.x, xv, acc_g[4 * i + 0]);
          acc_g[4 * i + 1] = fmaf(g.y, xv, acc_g[4 * i + 1]);
          acc_g[4 * i + 2] = fmaf(g.z, xv, acc_g[4 * i + 2]);
          acc_g[4 * i + 3] = fmaf(g.w, xv, acc_g[4 * i + 3]);
        }
      }
    }
  }

  // ---- gate ----
#pragma unroll
  for (int i = 0; i < PPW; ++i) {
    int p = warp * PPW + i;
    float th = tanhf(acc_t[i]);
    float sg = 1.0f / (1.0f + expf(-acc_g[i]));
    float zv = (p < Ch) ? th * sg : 0.0f;
    Zs[p * 32 + lane] = zv;
    if (p < Ch && t_ok && gate_tanh) {
      int64_t off = ((int64_t)b * Ch + p) * D.T + t;
      gate_tanh[off] = th;
      gate_sig[off] = sg;
    }
  }

  // ---- second contraction: rows [0,Cr) = res, [Cr,Cr+Cs) = skip ----
  const int R = D.Cr + D.Cs;
  const int r_begin = D.write_residual ? 0 : D.Cr;
  float* W2s = Wt;
  for (int rbase = r_begin; rbase < R; rbase += NR) {
    float acc[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) acc[i] = 0.0f;
    for (int k0 = 0; k0 < Ch; k0 += RB_KC) {
      __syncthreads();
      for (int e = tid; e < RB_KC * NR; e += RB_WARPS * 32) {
        int kc = e % RB_KC, rr = e / RB_KC;
        int r = rbase + rr, k = k0 + kc;
        float v = 0.0f;
        if (r < R && k < Ch)
          v = (r < D.Cr) ? __ldg(W.res_w + (int64_t)r * Ch + k)
                         : __ldg(W.skip_w + (int64_t)(r - D.Cr) * Ch + k);
        W2s[kc * PITCH2 + rr] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kc = 0; kc < RB_KC; ++kc) {
        float zv = (k0 + kc < NP) ? Zs[(k0 + kc) * 32 + lane] : 0.0f;
        const float4* w4 = reinterpret_cast<const float4*>(W2s + kc * PITCH2 + warp * RPW);
#pragma unroll
        for (int i = 0; i < RPW / 4; ++i) {
          float4 a = w4[i];
          acc[4 * i + 0] = fmaf(a.x, zv, acc[4 * i + 0]);
          acc[4 * i + 1] = fmaf(a.y, zv, acc[4 * i + 1]);
          acc[4 * i + 2] = fmaf(a.z, zv, acc[4 * i + 2]);
          acc[4 * i + 3] = fmaf(a.w, zv, acc[4 * i + 3]);
        }
      }
    }
    if (t_ok) {
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        int r = rbase + warp * RPW + i;
        if (r >= R) continue;
        if (r < D.Cr) {
          int64_t off = ((int64_t)b * D.Cr + r) * D.T + t;
          residual[off] = acc[i] + __ldg(W.res_b + r) + __ldg(x + off);
        } else {
          int sidx = r - D.Cr;
          int64_t off = ((int64_t)b * D.Cs + sidx) * D.T + t;
          float v = acc[i] + __ldg(W.skip_b + sidx);
          if (D.skip_accumulate) v += skip[off];
          skip[off] = v;
        }
      }
    }
  }
}

template <int PPW>
static int launch_fwd(const vqw_resblock_desc& d, const float* x, const float* cond,
                      const vqw_resblock_weights& w, float* residual, float* skip, float* gt,
                      float* gs, cudaStream_t stream) {
  constexpr int NP = RB_WARPS * PPW, PITCH = NP + 4, NR = RB_WARPS * 32, PITCH2 = NR + 4;
  constexpr int WS = (2 * PITCH > PITCH2) ? 2 * PITCH : PITCH2;
  size_t smem = sizeof(float) * (RB_KC * 32 + RB_KC * WS + (size_t)NP * 32);
  auto kern = resblock_fwd_simt_kernel<PPW>;
  VQW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(d.T, 32), d.B);
  kern<<<grid, RB_WARPS * 32, smem, stream>>>(d, x, cond, w, residual, skip, gt, gs);
  VQW_CHECK_LAUNCH("resblock_fwd_simt_kernel");
  return 0;
}

static int validate(const vqw_resblock_desc& d, const char* who) {
  VQW_REQUIRE(d.B >= 0 && d.T >= 0, "%s: bad B/T", who);
  VQW_REQUIRE(d.Cr > 0 && d.Cd > 0 && d.Cs > 0 && d.Cc > 0, "%s: channel counts must be > 0", who);
  VQW_REQUIRE(d.Cd % 2 == 0, "%s: dilated_channels must be even (split_axis, modules.py:47)", who);
  VQW_REQUIRE(d.fs >= 1 && d.fs <= 8 && d.dilation >= 1, "%s: bad filter_size/dilation", who);
  VQW_REQUIRE(d.B <= 65535, "%s: B > 65535", who);
  return 0;
}

}  // namespace vqw

extern "C" int vqw_resblock_forward(const vqw_resblock_desc* desc, const float* x,
                                    const float* cond, const vqw_resblock_weights* w,
                                    float* residual, float* skip, float* gate_tanh,
                                    float* gate_sig, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(desc && w, "vqw_resblock_forward: null descriptor");
  const vqw_resblock_desc& d = *desc;
  if (int rc = validate(d, "vqw_resblock_forward")) return rc;
  VQW_REQUIRE(x && cond && skip, "vqw_resblock_forward: null tensor");
  VQW_REQUIRE(residual || !d.write_residual, "vqw_resblock_forward: residual is null");
  VQW_REQUIRE((gate_tanh == nullptr) == (gate_sig == nullptr),
              "vqw_resblock_forward: gate_tanh and gate_sig must be given together");
  VQW_REQUIRE(w->conv_w && w->conv_b && w->cond_w && w->cond_b && w->res_w && w->res_b &&
                  w->skip_w && w->skip_b, "vqw_resblock_forward: null weight");
  if (d.B == 0 || d.T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  VQW_REQUIRE(d.mode == VQW_MODE_FP32,
              "vqw_resblock_forward: only VQW_MODE_FP32 here; the tensor-core modes run through "
              "vqw_resnet_forward (mode %d)", d.mode);
  int Ch = d.Cd / 2;
  VQW_REQUIRE(Ch <= 256, "vqw_resblock_forward: fp32 path supports dilated_channels <= 512");
  if (Ch <= 32) return launch_fwd<4>(d, x, cond, *w, residual, skip, gate_tanh, gate_sig, st);
  if (Ch <= 64) return launch_fwd<8>(d, x, cond, *w, residual, skip, gate_tanh, gate_sig, st);
  if (Ch <= 128) return launch_fwd<16>(d, x, cond, *w, residual, skip, gate_tanh, gate_sig, st);
  return launch_fwd<32>(d, x, cond, *w, residual, skip, gate_tanh, gate_sig, st);
}

extern "C" int64_t vqw_resblock_backward_workspace(const vqw_resblock_desc* desc) {
  if (!desc) return -1;
  return (int64_t)sizeof(float) * desc->B * desc->Cd * desc->T;
}

// Backward = gate-gradient GEMM (gz -> gh), data-gradient convs (gx, gcond) and the weight
// gradients, following SURVEY.md appendix B; fp32 path composes the generic conv family.
extern "C" int vqw_resblock_backward(const vqw_resblock_desc* desc, const float* g_res,
                                     const float* g_skip, const float* x, const float* cond,
                                     const float* gate_tanh, const float* gate_sig,
                                     const vqw_resblock_weights* w, float* gx, float* gcond,
                                     const vqw_resblock_wgrads* gw, void* workspace,
                                     vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(desc && w, "vqw_resblock_backward: null descriptor");
  const vqw_resblock_desc& d = *desc;
  if (int rc = validate(d, "vqw_resblock_backward")) return rc;
  VQW_REQUIRE(g_skip && x && cond && gate_tanh && gate_sig && workspace,
              "vqw_resblock_backward: null tensor");
  if (d.B == 0 || d.T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int Ch = d.Cd / 2;
  float* gh = reinterpret_cast<float*>(workspace);

  // gz = Wr^T g_res + Ws^T g_skip, then the gate derivative -> gh (B,Cd,T)
  {
    vqw_conv_desc c = {};
    c.B = d.B; c.M = Ch; c.T = d.T;
    int n = 0;
    if (g_res) {
      c.src[n] = {g_res, w->res_w, nullptr, d.Cr, d.T, 1, Ch, 1, 0, 1, 0};
      ++n;
    }
    c.src[n] = {g_skip, w->skip_w, nullptr, d.Cs, d.T, 1, Ch, 1, 0, 1, 0};
    ++n;
    c.nsrc = n;
    c.gate_tanh = gate_tanh;
    c.gate_sig = gate_sig;
    if (int rc = launch_conv(c, gh, st)) return rc;
  }
  // gx[t] = g_res[t] + sum_j Wc[:,:,j]^T gh[t + dil*(fs-1-j)]
  if (gx) {
    for (int j0 = 0; j0 < d.fs; j0 += VQW_MAX_SRC) {
      vqw_conv_desc c = {};
      c.B = d.B; c.M = d.Cr; c.T = d.T;
      int n = 0;
      for (int j = j0; j < d.fs && n < VQW_MAX_SRC; ++j, ++n)
        c.src[n] = {gh, w->conv_w + j, nullptr, d.Cd, d.T, d.fs, d.Cr * d.fs, 1,
                    d.dilation * (d.fs - 1 - j), 1, 0};
      c.nsrc = n;
      c.addend = (j0 == 0) ? g_res : nullptr;
      c.accumulate = (j0 == 0) ? 0 : 1;
      if (int rc = launch_conv(c, gx, st)) return rc;
    }
  }
  // gcond += Wp^T gh
  if (gcond) {
    vqw_conv_desc c = {};
    c.B = d.B; c.M = d.Cc; c.T = d.T;
    c.nsrc = 1;
    c.src[0] = {gh, w->cond_w, nullptr, d.Cd, d.T, 1, d.Cc, 1, 0, 1, 0};
    c.accumulate = 1;
    if (int rc = launch_conv(c, gcond, st)) return rc;
  }
  if (gw) {
    for (int j = 0; j < d.fs; ++j) {
      vqw_wgrad_desc g = {};
      g.B = d.B; g.M = d.Cd; g.T = d.T; g.a = gh; g.in = x; g.K = d.Cr; g.Tin = d.T;
      g.mul = 1; g.shift = -d.dilation * (d.fs - 1 - j); g.div = 1;
      g.gm = d.Cr * d.fs; g.gk = d.fs;
      if (int rc = launch_wgrad(g, gw->conv_w + j, (j == 0) ? gw->conv_b : nullptr, st)) return rc;
    }
    {
      vqw_wgrad_desc g = {};
      g.B = d.B; g.M = d.Cd; g.T = d.T; g.a = gh; g.in = cond; g.K = d.Cc; g.Tin = d.T;
      g.mul = 1; g.shift = 0; g.div = 1; g.gm = d.Cc; g.gk = 1;
      if (int rc = launch_wgrad(g, gw->cond_w, gw->cond_b, st)) return rc;
    }
    if (g_res) {
      vqw_wgrad_desc g = {};
      g.B = d.B; g.M = d.Cr; g.T = d.T; g.a = g_res; g.in = gate_tanh; g.in_mul = gate_sig;
      g.K = Ch; g.Tin = d.T; g.mul = 1; g.shift = 0; g.div = 1; g.gm = Ch; g.gk = 1;
      if (int rc = launch_wgrad(g, gw->res_w, gw->res_b, st)) return rc;
    }
    {
      vqw_wgrad_desc g = {};
      g.B = d.B; g.M = d.Cs; g.T = d.T; g.a = g_skip; g.in = gate_tanh; g.in_mul = gate_sig;
      g.K = Ch; g.Tin = d.T; g.mul = 1; g.shift = 0; g.div = 1; g.gm = Ch; g.gk = 1;
      if (int rc = launch_wgrad(g, gw->skip_w, gw->skip_b, st)) return rc;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// whole stack (modules.py:89-96)
// ---------------------------------------------------------------------------------------
extern "C" int64_t vqw_resnet_forward_workspace(const vqw_resnet_desc* desc) {
  if (!desc) return -1;
  if (desc->mode == VQW_MODE_FP32)
    return 2 * (int64_t)sizeof(float) * desc->B * desc->Cr * desc->T + 512;
  return vqw::resnet_tc_workspace(*desc);
}

extern "C" int vqw_resnet_forward(const vqw_resnet_desc* desc, const float* x, const float* cond,
                                  const vqw_resblock_weights* weights, float* const* residuals,
                                  float* skip, float* const* gate_tanh, float* const* gate_sig,
                                  void* workspace, void* saved, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(desc && weights, "vqw_resnet_forward: null descriptor");
  const vqw_resnet_desc& d = *desc;
  VQW_REQUIRE(d.n_blocks >= 1 && d.dilations, "vqw_resnet_forward: n_blocks/dilations");
  VQW_REQUIRE(d.B >= 0 && d.T >= 0, "vqw_resnet_forward: bad B/T");
  if (d.B == 0 || d.T == 0) return 0;
  VQW_REQUIRE(x && cond && skip, "vqw_resnet_forward: null tensor");
  VQW_REQUIRE((gate_tanh == nullptr) == (gate_sig == nullptr),
              "vqw_resnet_forward: gate_tanh and gate_sig must be given together");
  cudaStream_t st = (cudaStream_t)stream;
  if (vqw_mode_tc(d.mode))
    return resnet_forward_tc(d, x, cond, weights, residuals, skip, gate_tanh, gate_sig, workspace,
                             saved, st);
  VQW_REQUIRE(d.mode == VQW_MODE_FP32, "vqw_resnet_forward: unknown mode %d", d.mode);
  VQW_REQUIRE(d.Cg == 0, "vqw_resnet_forward: the hoisted global condition (Cg > 0) is a tensor-core "
                         "mode feature; fp32 mode takes the concatenated condition");
  float* pp[2] = {nullptr, nullptr};
  if (workspace) {
    uintptr_t a = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    pp[0] = reinterpret_cast<float*>(a);
    pp[1] = pp[0] + (int64_t)d.B * d.Cr * d.T;
  }
  const float* cur = x;
  for (int i = 0; i < d.n_blocks; ++i) {
    const bool last = i == d.n_blocks - 1;
    vqw_resblock_desc b = {};
    b.B = d.B; b.T = d.T; b.Cr = d.Cr; b.Cd = d.Cd; b.Cs = d.Cs; b.Cc = d.Cc; b.fs = d.fs;
    b.dilation = d.dilations[i];
    b.skip_accumulate = i > 0;
    b.write_residual = (!last || d.keep_last_residual) ? 1 : 0;
    b.mode = VQW_MODE_FP32;
    float* out = nullptr;
    if (b.write_residual) {
      out = residuals ? residuals[i] : nullptr;
      if (!out) out = pp[i & 1];
      VQW_REQUIRE(out != nullptr, "vqw_resnet_forward: no buffer for the residual of block %d "
                                  "(pass residuals[] or a workspace)", i);
    }
    if (int rc = vqw_resblock_forward(&b, cur, cond, &weights[i], out, skip,
                                      gate_tanh ? gate_tanh[i] : nullptr,
                                      gate_sig ? gate_sig[i] : nullptr, stream))
      return rc;
    cur = out;
  }
  (void)st;
  return 0;
}

extern "C" int64_t vqw_resnet_saved_bytes(const vqw_resnet_desc* desc) {
  if (!desc) return -1;
  if (desc->mode == VQW_MODE_FP32) return 0;
  return vqw::resnet_tc_saved_bytes(*desc);
}

extern "C" int64_t vqw_resnet_backward_workspace(const vqw_resnet_desc* desc) {
  if (!desc) return -1;
  if (desc->mode == VQW_MODE_FP32)
    return (int64_t)sizeof(float) * desc->B * desc->T * ((int64_t)desc->Cd + 2 * desc->Cr) + 1024;
  return vqw::resnet_backward_tc_workspace(*desc);
}

extern "C" int vqw_resnet_backward(const vqw_resnet_desc* desc, const float* g_skip,
                                   const float* g_last_res, const float* x, const float* cond,
                                   float* const* residuals, float* const* gate_tanh,
                                   float* const* gate_sig, const vqw_resblock_weights* weights,
                                   float* gx, float* gcond, const vqw_resblock_wgrads* wgrads,
                                   void* workspace, const void* saved, vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(desc && weights && wgrads, "vqw_resnet_backward: null descriptor");
  const vqw_resnet_desc& d = *desc;
  VQW_REQUIRE(d.n_blocks >= 1 && d.dilations, "vqw_resnet_backward: n_blocks/dilations");
  if (d.B == 0 || d.T == 0) return 0;
  VQW_REQUIRE(g_skip && gate_tanh && gate_sig && workspace && gcond,
              "vqw_resnet_backward: null tensor");
  if (vqw_mode_tc(d.mode)) {
    VQW_REQUIRE(saved, "vqw_resnet_backward: the tensor-core modes need the `saved` buffer that "
                       "vqw_resnet_forward filled");
    return resnet_backward_tc(d, g_skip, g_last_res, gate_tanh, gate_sig, weights, gx, gcond,
                              wgrads, workspace, saved, (cudaStream_t)stream);
  }
  VQW_REQUIRE(d.mode == VQW_MODE_FP32, "vqw_resnet_backward: unknown mode %d", d.mode);
  VQW_REQUIRE(d.Cg == 0, "vqw_resnet_backward: Cg > 0 is a tensor-core mode feature");
  VQW_REQUIRE(x && cond, "vqw_resnet_backward: null tensor");
  VQW_REQUIRE(d.n_blocks == 1 || residuals, "vqw_resnet_backward: residuals[] is required");
  uintptr_t a = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  float* gh = reinterpret_cast<float*>(a);
  float* pp[2] = {gh + (int64_t)d.B * d.Cd * d.T, gh + (int64_t)d.B * (d.Cd + d.Cr) * d.T};
  const float* g_res = g_last_res;
  for (int i = d.n_blocks - 1; i >= 0; --i) {
    vqw_resblock_desc b = {};
    b.B = d.B; b.T = d.T; b.Cr = d.Cr; b.Cd = d.Cd; b.Cs = d.Cs; b.Cc = d.Cc; b.fs = d.fs;
    b.dilation = d.dilations[i];
    b.mode = VQW_MODE_FP32;
    const float* xin = (i == 0) ? x : residuals[i - 1];
    VQW_REQUIRE(xin && gate_tanh[i] && gate_sig[i], "vqw_resnet_backward: block %d saved tensors", i);
    float* out = (i == 0) ? gx : pp[i & 1];
    if (int rc = vqw_resblock_backward(&b, g_res, g_skip, xin, cond, gate_tanh[i], gate_sig[i],
                                       &weights[i], out, gcond, &wgrads[i], gh, stream))
      return rc;
    g_res = out;
    if (i == 0) break;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// output head (modules.py:155-159), tensor-core modes
// ---------------------------------------------------------------------------------------
static int head_check(const vqw_head_desc* desc, const char* who) {
  using namespace vqw;
  VQW_REQUIRE(desc != nullptr, "%s: null descriptor", who);
  VQW_REQUIRE(vqw_mode_tc(desc->mode),
              "%s: tensor-core modes only (fp32 runs through vqw_conv_forward)", who);
  VQW_REQUIRE(head_tc_supported(*desc),
              "%s: needs skip_channels %% 256 == 0, T >= 128 and T %% 8 == 0 (Cs=%d T=%d)", who,
              desc->Cs, desc->T);
  return 0;
}
extern "C" int64_t vqw_head_workspace(const vqw_head_desc* desc) {
  return desc ? vqw::head_tc_workspace(*desc) : -1;
}
extern "C" int64_t vqw_head_saved_bytes(const vqw_head_desc* desc) {
  return desc ? vqw::head_tc_saved_bytes(*desc) : -1;
}
extern "C" int vqw_head_forward(const vqw_head_desc* desc, const float* skip, const float* W1,
                                const float* b1, const float* W2, const float* b2, float* y,
                                void* workspace, void* saved, vqw_stream_t stream) {
  if (int rc = head_check(desc, "vqw_head_forward")) return rc;
  return vqw::head_forward_tc(*desc, skip, W1, b1, W2, b2, y, workspace, saved, (cudaStream_t)stream);
}
extern "C" int vqw_head_backward(const vqw_head_desc* desc, const float* gy, const float* W1,
                                 const float* W2, float* gskip, float* gW1, float* gb1, float* gW2,
                                 float* gb2, void* workspace, const void* saved,
                                 vqw_stream_t stream) {
  if (int rc = head_check(desc, "vqw_head_backward")) return rc;
  VQW_REQUIRE(gy != nullptr, "vqw_head_backward: null argument");
  return vqw::head_backward_tc(*desc, gy, nullptr, W1, W2, gskip, gW1, gb1, gW2, gb2, workspace, saved,
                               (cudaStream_t)stream);
}
extern "C" int vqw_head_loss_forward(const vqw_head_desc* desc, const float* skip, const float* W1,
                                     const float* b1, const float* W2, const float* b2,
                                     const int32_t* t_labels, const float* t_values, int quantize,
                                     float log_scale_min, double* loss, float* y_opt, void* workspace,
                                     void* saved, vqw_stream_t stream) {
  if (int rc = head_check(desc, "vqw_head_loss_forward")) return rc;
  return vqw::head_loss_forward_tc(*desc, skip, W1, b1, W2, b2, t_labels, t_values, quantize,
                                   log_scale_min, loss, y_opt, workspace, saved, (cudaStream_t)stream);
}
extern "C" int vqw_head_loss_backward(const vqw_head_desc* desc, const float* g_loss, const float* W1,
                                      const float* W2, float* gskip, float* gW1, float* gb1,
                                      float* gW2, float* gb2, void* workspace, const void* saved,
                                      vqw_stream_t stream) {
  using namespace vqw;
  if (int rc = head_check(desc, "vqw_head_loss_backward")) return rc;
  VQW_REQUIRE(g_loss != nullptr, "vqw_head_loss_backward: null upstream gradient");
  return head_backward_tc(*desc, nullptr, g_loss, W1, W2, gskip, gW1, gb1, gW2, gb2, workspace, saved,
                          (cudaStream_t)stream);
}
