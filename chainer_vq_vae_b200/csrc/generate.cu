// Persistent autoregressive generation kernel (sm_100a).
//
// Replaces the reference's per-sample Python loop, generate.py:109-145, together with the
// queue-based incremental WaveNet it drives (WaveNet.initialize / generate, modules.py:232-255;
// ResidualNet.generate :102-110; ResidualBlock.push / pop :58-74): ~500 Chainer calls, 16.8 MB
// of concat-shift queue copies and a device->host sync per generated sample become ONE
// cooperative kernel launch for the whole utterance.
//
//   * The kernel stays resident on every SM and walks the samples; phases that depend on each
//     other are separated by grid-wide barriers (the layer chain of one sample is strictly
//     sequential; only utterances are independent -- "replicas only").
//   * The dilation queues are ring buffers in global memory (L2 resident): a push is one
//     512-float row write, a tap read is an index computation -- nothing is shifted.
//   * Sampling happens on the device from a host-supplied uniform stream with
//     numpy.random.choice's rule (float64 cumsum of the float32 softmax, normalise,
//     searchsorted(side='right')), so a run is comparable with generate.py:139-141 given the
//     same draws.  The first input is all-zeros, not a one-hot (generate.py:51).
//
// One phase + one grid barrier per block.  Block l's gate needs x_l, which block l-1 produces in
// the same step; to avoid a second barrier the current-tap product is re-associated:
//   Wc_cur x_l = Wc_cur (x_{l-1} + br_{l-1}) + (Wc_cur Wr_{l-1}) z_{l-1}
// with M_l = Wc_cur_l Wr_{l-1} (Cd x Cd/2) and c_l = Wc_cur_l br_{l-1} computed once per
// utterance (fp32, vqw conv kernel).  Phase l then only depends on z_{l-1} and x_{l-1}:
//   T1  (2 gate pairs x 4 K chunks per CTA, one warp each; partials reduced in shared memory)
//       h_l = b + Wp cond[t] + past taps of ring_l + Wc_cur x_{l-1} + M_l z_{l-1} + c_l -> z_l
//   T2  (one output row per warp)   x_l = Wr_{l-1} z_{l-1} + br_{l-1} + x_{l-1}  -> ring_l,
//       skip += Ws_{l-1} z_{l-1} + bs_{l-1}
// then the head (relu, proj1, relu, proj2), softmax and the draw.
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace vqw {

constexpr int GEN_THREADS = 256;
constexpr int GEN_WARPS = GEN_THREADS / 32;

struct GenBlock {
  const float *conv_w, *conv_b, *cond_w, *cond_b, *res_w, *res_b, *skip_w, *skip_b;
  int dilation;
  int qlen;        // dil*(fs-1)+1 ring slots
  long long qoff;  // offset (floats) of this block's ring in the queue area
  const float* mmat;   // M_l = Wc_cur_l Wr_{l-1}  (Cd, Cd/2), null for l = 0
  const float* cvec;   // c_l = Wc_cur_l br_{l-1}  (Cd), null for l = 0
};

struct GenParams {
  int n_blocks, fs, Cr, Cd, Cs, Cc, Q;
  int T_total, n_steps, t_start, cond_t0;
  int mol;                     // mixture-of-logistics decoder: scalar input, Q = 3 * nr_mix outputs
  float log_scale_min;
  const GenBlock* blocks;      // device array
  const float *embed_w, *embed_b, *proj1_w, *proj1_b, *proj2_w, *proj2_b;
  const float* cond;           // (Cc, T_total)
  const double* uniforms;      // n_steps (mol: n_steps * nr_mix)
  const int32_t* forced;       // n_steps or null: teacher forcing (sample still recorded)
  int32_t* samples;            // n_steps (mol: the float32 values, as bits)
  float* logits;               // n_steps * Q or null
  float* queues;               // rings
  float* zbuf;                 // [2][Cd/2] gated activations (ping-pong over blocks)
  float* xbuf;                 // [2][Cr] block inputs (ping-pong over blocks)
  float* skipacc;              // [Cs]
  float* skiplast;             // [Cs] skip rows of the last block (tagged-exchange mode)
  uint2* xtag;                 // [2][Cr] {x_l + br_l bits, tag}: tagged exchange between phases
  uint2* ztag;                 // [2][Cd/2] {z_l bits, tag}
  float* h1;                   // [Cs] relu(proj1(relu(skip)))
  float* logit_buf;            // [Q]
  int32_t* state;              // [0] = sample(t-1), [1] = sample(t-2)  (-1 = none; mol: float bits)
  unsigned int* barrier;       // grid barrier counter (zeroed by the host before the launch)
  int use_tags;                // tagged exchange between phases instead of grid barriers
  long long* dbg;              // VQW_GEN_TIMELINE=1: per-phase clock64 stamps of one CTA (or null)
  int dbg_cta, dbg_step;
};

// Grid-wide barrier (all CTAs are co-resident: cooperative launch): one atomic arrival per CTA
// on a monotonically increasing counter, then a polling acquire load.  ~3x cheaper than
// cooperative_groups' grid.sync() at one CTA per SM.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    // release-add without a return value: the CTA's writes (ordered before this thread by the
    // bar.sync above) become visible to whoever acquires the counter; no round trip to wait for
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - epoch) < 0);
  }
  __syncthreads();
}

// Tagged exchange.  The values one phase hands to the next (x_l + br_l and z_l: 768 floats) are
// written as 8-byte {value, tag} words and the consumers poll the words themselves until the
// tag is the one of (step, phase): an 8-byte aligned store is single-copy atomic, so a matching
// tag means the value next to it is the new one.  This replaces "grid barrier, then load" (two
// dependent round trips plus the arrival skew, ~3.5 k cycles of an ~8.7 k-cycle phase) by one
// store -> load propagation.  Ping-pong over two buffers is enough: a CTA can only be two phases
// ahead of another after consuming something that CTA produced after its own reads.
__device__ __forceinline__ void st_tagged(uint2* p, float v, uint32_t tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ float ld_tagged(const uint2* p, uint32_t tag) {
  uint32_t v, g;
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {   // bounded: never hang the GPU
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(g) : "l"(p) : "memory");
    if (g == tag) return __uint_as_float(v);
  }
  __trap();
  return 0.0f;
}
__device__ __forceinline__ uint32_t phase_tag(int t, int l) { return (uint32_t)t * 512u + (uint32_t)l + 1u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Dot products of contiguous weight row segments (global) with a shared-memory vector.  The
// kernel is latency bound (one dependent DRAM/L2 round trip per phase), so all loads of a call
// are issued before the first FMA: up to NV float4 per lane per row are in flight at once.
template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ w, int n, int lane, float4* r) {
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int i = lane * 4 + u * 128;
    r[u] = (i + 3 < n) ? __ldg(reinterpret_cast<const float4*>(w + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NV>
__device__ __forceinline__ float fma_row(const float4* r, const float* v, int n, int lane) {
  float acc = 0.0f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int i = lane * 4 + u * 128;
    if (i + 3 < n) {
      acc = fmaf(r[u].x, v[i], acc);
      acc = fmaf(r[u].y, v[i + 1], acc);
      acc = fmaf(r[u].z, v[i + 2], acc);
      acc = fmaf(r[u].w, v[i + 3], acc);
    }
  }
  return acc;
}
// generic (any length / alignment) fallback and tail handling
__device__ __forceinline__ float dot_tail(const float* __restrict__ w, const float* v, int from, int n,
                                          int lane) {
  float acc = 0.0f;
  for (int j = from + lane; j < n; j += 32) acc = fmaf(__ldg(w + j), v[j], acc);
  return acc;
}
constexpr int NVMAX = 4;   // 4 float4 per lane = 512 floats per row segment per batch

// two rows (the tanh row and the sigmoid row of a gate pair) against the same vector
__device__ __forceinline__ void dot2(const float* __restrict__ w0, const float* __restrict__ w1,
                                     const float* v, int n, int lane, float& o0, float& o1) {
  float a0 = 0.0f, a1 = 0.0f;
  const bool aligned = ((((uintptr_t)w0) | ((uintptr_t)w1)) & 15) == 0;
  int done = 0;
  if (aligned) {
    for (; done + 3 < n; done += NVMAX * 128) {
      float4 r0[NVMAX], r1[NVMAX];
      load_row<NVMAX>(w0 + done, n - done, lane, r0);
      load_row<NVMAX>(w1 + done, n - done, lane, r1);
      a0 += fma_row<NVMAX>(r0, v + done, n - done, lane);
      a1 += fma_row<NVMAX>(r1, v + done, n - done, lane);
    }
    done = n & ~3;
  }
  a0 += dot_tail(w0, v, done, n, lane);
  a1 += dot_tail(w1, v, done, n, lane);
  o0 = warp_sum(a0);
  o1 = warp_sum(a1);
}
__device__ __forceinline__ float dot_row(const float* __restrict__ w, const float* v, int n, int lane) {
  float a0 = 0.0f;
  int done = 0;
  if ((((uintptr_t)w) & 15) == 0) {
    for (; done + 3 < n; done += NVMAX * 128) {
      float4 r0[NVMAX];
      load_row<NVMAX>(w + done, n - done, lane, r0);
      a0 += fma_row<NVMAX>(r0, v + done, n - done, lane);
    }
    done = n & ~3;
  }
  a0 += dot_tail(w, v, done, n, lane);
  return warp_sum(a0);
}
// pull the lines of a weight row segment into L2 ahead of the phase that reads them
__device__ __forceinline__ void prefetch_row(const float* w, int n, int lane) {
  for (int i = lane * 32; i < n; i += 32 * 32)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(w + i));
}

__global__ void __launch_bounds__(GEN_THREADS, 1)
generate_kernel(const GenParams P) {
  unsigned int epoch = 0;
  extern __shared__ __align__(16) float sm[];
  const int Ch = P.Cd / 2;
  const int KX = P.fs * P.Cr;           // interleaved tap vector length
  const int KXp = (KX + 3) & ~3, Ccp = (P.Cc + 3) & ~3;
  float* vx2 = sm;                      // [2][KX] v[c*fs + j] = x[t - s_j][c]; current tap = x_{l-1}+br
  float* vc2 = vx2 + 2 * KXp;           // [2][Cc] cond[:, t]
  float* zs = vc2 + 2 * Ccp;            // [Ch]  z_{l-1}
  float* xs = zs + ((Ch + 3) & ~3);     // [max(Cr, Cs, Q)] scratch vector
  float* ps = xs + ((max(max(P.Cr, P.Cs), P.Q) + 3) & ~3);   // [Q] softmax numerators
  float* part = ps + ((P.Q + 3) & ~3);  // [2 pairs][4 chunks][2] partial sums
  GenBlock* sblk = reinterpret_cast<GenBlock*>(part + 16);   // block table copy
  __shared__ int s_sample;
  __shared__ int s_slot[256];   // t % qlen of every block, refreshed once per step (<= 256 blocks)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gwarp = blockIdx.x * GEN_WARPS + warp;
  const int nwarps = gridDim.x * GEN_WARPS;
  // K chunks of a gate row: 3 over the tap vector, the 4th = condition + M_l z_{l-1}
  const int c3 = (((KX + 2) / 3) + 3) & ~3;
  const int npb = (Ch + 1) / 2;         // pair blocks (2 pairs each)

  for (int i = tid; i < P.n_blocks * (int)(sizeof(GenBlock) / 8); i += GEN_THREADS)
    reinterpret_cast<long long*>(sblk)[i] = reinterpret_cast<const long long*>(P.blocks)[i];
  __syncthreads();

  // Register prefetch of the NEXT phase's weight rows: the rows a warp reads in phase l depend
  // only on (l, warp), never on data, so their loads are issued before the barrier that ends
  // phase l-1 and the DRAM latency hides behind it.  Fast path: every row segment fits 4 (tap
  // chunk, condition) or 2 (M_l, Wr/Ws rows) float4 per lane and one pair block per CTA.
  const bool fastp = (c3 <= 512) && (P.Cc <= 512) && (Ch <= 256) && (npb <= (int)gridDim.x) &&
                     ((KX & 3) == 0) && ((P.Cc & 3) == 0) && ((Ch & 3) == 0);
  float4 r0[4], r1[4], m0[2], m1[2], q4[2];
  // prefetched biases (gate pair: conv + cond parts, summed at use so the loads stay in flight)
  float pb_t = 0.0f, pb_t2 = 0.0f, pb_g = 0.0f, pb_g2 = 0.0f, pb_row = 0.0f;
  const int t2_first = (npb * GEN_WARPS) % nwarps;
  // loop invariants of the per-phase code (integer division and modulo are ~25 dependent
  // instructions each; measured: they were most of the ~3 k cycles between T2 and the barrier)
  const int r_t2 = (gwarp - t2_first + nwarps) % nwarps;      // this warp's T2 row
  // A warp owns the same skip row in every phase but the last one, so the running skip sum of a
  // step lives in a register of its lane 0 and reaches memory once (instead of a dependent global
  // read-modify-write per block)
  const bool skreg = (P.Cr + P.Cs) <= nwarps && P.n_blocks >= 2;
  float skip_reg = 0.0f;
  const bool tagx = skreg && P.use_tags && P.n_blocks < 500;   // tagged exchange instead of barriers
  // a CTA without gate pairs and without T2 rows neither produces nor polls the exchanged vectors
  const bool cta_works = __syncthreads_or((int)blockIdx.x < npb || r_t2 < P.Cr + P.Cs) != 0;
  const int p_t1 = 2 * blockIdx.x + (warp >> 2), kc_t1 = warp & 3;
  const int k0_t1 = kc_t1 * c3, n_t1 = max(0, min(c3, KX - k0_t1));
  // past taps and condition of phase l into buffer l&1 (known before phase l-1 ends)
  auto gather_past = [&](int l, int t) {
    if (l >= P.n_blocks) return;
    const GenBlock& nb = sblk[l];
    float* vx = vx2 + (l & 1) * KXp;
    float* vc = vc2 + (l & 1) * Ccp;
    const float* ring = P.queues + nb.qoff;
    for (int i0 = tid; i0 < KX; i0 += 8 * GEN_THREADS) {
      float tmp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * GEN_THREADS;
        float v = 0.0f;
        if (i < KX) {
          const int c = i / P.fs, j = i - c * P.fs;
          const int s = nb.dilation * (P.fs - 1 - j);
          if (s > 0 && t - s >= 0) v = __ldcg(ring + (long long)((t - s) % nb.qlen) * P.Cr + c);
        }
        tmp[u] = v;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * GEN_THREADS;
        if (i < KX && (i % P.fs) != P.fs - 1) vx[i] = tmp[u];
      }
    }
    for (int i = tid; i < P.Cc; i += GEN_THREADS)
      vc[i] = __ldg(P.cond + (long long)i * P.T_total + (t - P.cond_t0));
  };
  // Software-pipelined form of the same gather (measured: the blocking gather above was ~45 % of a
  // phase -- two dependent HBM round trips, the rings are evicted from L2 by the 175 MB weight
  // stream).  The past taps of EVERY layer at step t were written in earlier steps, so the loads
  // for phase l+2 are issued during phase l and only stored to shared memory one phase later;
  // the condition column is the same for all layers and is loaded once per step.
  constexpr int GQ = 4;
  const int npast = (P.fs - 1) * P.Cr;
  const bool gfast = npast <= GQ * GEN_THREADS;
  float gt[GQ];
  int gj[GQ], gc[GQ];          // tap and channel of this thread's elements (step independent)
#pragma unroll
  for (int u = 0; u < GQ; ++u) {
    const int e = tid + u * GEN_THREADS;
    gj[u] = e / P.Cr;
    gc[u] = e - gj[u] * P.Cr;
  }
  auto gather_issue = [&](int l, int t) {
    if (l >= P.n_blocks) return;
    const GenBlock& nb = sblk[l];
    const float* ring = P.queues + nb.qoff;
#pragma unroll
    for (int u = 0; u < GQ; ++u) {
      const int e = tid + u * GEN_THREADS;
      float v = 0.0f;
      if (e < npast) {
        const int s = nb.dilation * (P.fs - 1 - gj[u]);       // 1 <= s < qlen
        if (t - s >= 0) {
          int slot = s_slot[l] - s;                            // (t - s) % qlen without dividing
          if (slot < 0) slot += nb.qlen;
          v = __ldcg(ring + (long long)slot * P.Cr + gc[u]);
        }
      }
      gt[u] = v;
    }
  };
  auto gather_commit = [&](int l) {
    if (l >= P.n_blocks) return;
    float* vx = vx2 + (l & 1) * KXp;
#pragma unroll
    for (int u = 0; u < GQ; ++u) {
      const int e = tid + u * GEN_THREADS;
      if (e < npast) vx[gc[u] * P.fs + gj[u]] = gt[u];
    }
  };
  // part 1 = the gate rows (r0, r1, m0, m1: dead as soon as this phase's T1 products are done, so
  // they are re-filled right there and the loads overlap the reduction, T2 and the exchange);
  // part 2 = biases, the T2 row and the rest, after T2
  auto prefetch_phase = [&](int l, bool part1, bool part2) {
    if (l > P.n_blocks) return;
    if (part1 && fastp && l < P.n_blocks) {
      const GenBlock& nb = sblk[l];
      const int p = p_t1, kc = kc_t1;
      if ((int)blockIdx.x < npb && p < Ch) {
        if (kc < 3) {
          const int k0 = k0_t1, n = n_t1;
          load_row<4>(nb.conv_w + (long long)p * KX + k0, n, lane, r0);
          load_row<4>(nb.conv_w + (long long)(Ch + p) * KX + k0, n, lane, r1);
        } else {
          load_row<4>(nb.cond_w + (long long)p * P.Cc, P.Cc, lane, r0);
          load_row<4>(nb.cond_w + (long long)(Ch + p) * P.Cc, P.Cc, lane, r1);
          if (l >= 1) {
            load_row<2>(nb.mmat + (long long)p * Ch, Ch, lane, m0);
            load_row<2>(nb.mmat + (long long)(Ch + p) * Ch, Ch, lane, m1);
          }
        }
      }
    }
    if (!part2) return;
    if (l < P.n_blocks && tid < 2 && 2 * (int)blockIdx.x + tid < Ch) {
      const GenBlock& nb = sblk[l];
      const int pp = 2 * blockIdx.x + tid;
      pb_t = __ldg(nb.conv_b + pp);
      pb_t2 = __ldg(nb.cond_b + pp);
      pb_g = __ldg(nb.conv_b + Ch + pp);
      pb_g2 = __ldg(nb.cond_b + Ch + pp);
    }
    if (l >= 1 && lane == 0) {
      const bool t1 = l < P.n_blocks;
      const int R = (t1 ? P.Cr : 0) + P.Cs, r_off = t1 ? 0 : P.Cr;
      const int r = r_t2;
      if (r < R) {
        const int rr = r + r_off;
        // x_l rows carry the NEXT block's residual bias (see T2), skip rows their own bias
        pb_row = rr < P.Cr ? __ldg(sblk[l].res_b + rr) : __ldg(sblk[l - 1].skip_b + (rr - P.Cr));
      }
    }
    if (!fastp) return;
    if (l >= 1) {
      const GenBlock& pbk = sblk[l - 1];
      const bool t1 = l < P.n_blocks;
      const int R = (t1 ? P.Cr : 0) + P.Cs, r_off = t1 ? 0 : P.Cr;
      const int r = r_t2;
      if (r < R) {
        const int rr = r + r_off;
        load_row<2>(rr < P.Cr ? pbk.res_w + (long long)rr * Ch : pbk.skip_w + (long long)(rr - P.Cr) * Ch,
                    Ch, lane, q4);
      }
    }
  };

  // the last two inputs (modules.py:246: the embed queue).  Every CTA draws the same sample from the
  // same logits, so each keeps its own copy instead of handing the state over through global
  // memory behind two grid barriers per sample
  int st1 = P.state[0], st2 = P.state[1];
  for (int step = 0; step < P.n_steps; ++step) {
    const int t = P.t_start + step;
    const long long step_start = P.dbg ? clock64() : 0;
    // ---- embed: x_0 = b + W[:, s(t-2), 0] + W[:, s(t-1), 1]   (modules.py:246-247; zeros at start)
    if (tid < P.n_blocks) s_slot[tid] = t % sblk[tid].qlen;
    __syncthreads();
    {
      const int s1 = st1, s2 = st2;
      float* ring0 = P.queues + sblk[0].qoff + (long long)s_slot[0] * P.Cr;
      for (int c = blockIdx.x * GEN_THREADS + tid; c < P.Cr; c += gridDim.x * GEN_THREADS) {
        float v = __ldg(P.embed_b + c);
        if (P.mol) {   // one input channel carrying the previous values (generate.py:137)
          const float* wr = P.embed_w + (long long)c * 2;
          v += __ldg(wr) * __int_as_float(s2) + __ldg(wr + 1) * __int_as_float(s1);
        } else {
          const float* wr = P.embed_w + (long long)c * P.Q * 2;
          if (s2 >= 0) v += __ldg(wr + 2 * s2);
          if (s1 >= 0) v += __ldg(wr + 2 * s1 + 1);
        }
        // slot 0 holds x_0 + br_0 (see T2)
        if (tagx) st_tagged(P.xtag + c, v + __ldg(sblk[0].res_b + c), phase_tag(t, 0));
        else P.xbuf[c] = v + __ldg(sblk[0].res_b + c);
        ring0[c] = v;                               // push (modules.py:72)
      }
      for (int c = blockIdx.x * GEN_THREADS + tid; c < P.Cs; c += gridDim.x * GEN_THREADS)
        P.skipacc[c] = 0.0f;
    }
    if (gfast) {
      gather_issue(0, t);
      for (int i = tid; i < P.Cc; i += GEN_THREADS)      // cond[:, t]: once per step, both buffers
        vc2[i] = vc2[Ccp + i] = __ldg(P.cond + (long long)i * P.T_total + (t - P.cond_t0));
      gather_commit(0);
      gather_issue(1, t);
    } else {
      gather_past(0, t);
    }
    prefetch_phase(0, true, true);
    grid_barrier(P.barrier, epoch);

    // phase l = 0..n-1 computes z_l (T1) and, for l >= 1, x_l and the skip rows of block l-1
    // (T2); phase n only runs T2 for the last block's skip rows.
    const bool rec = P.dbg != nullptr && (int)blockIdx.x == P.dbg_cta && step == P.dbg_step && tid == 0;
    if (rec) P.dbg[105] = step_start;
    for (int l = 0; l <= P.n_blocks; ++l) {
      const bool has_t1 = l < P.n_blocks, has_t2 = l >= 1;
      if (rec && l < 12) P.dbg[8 * l] = clock64();
      const GenBlock& blk = sblk[has_t1 ? l : l - 1];     // T1's block
      const GenBlock& pb = sblk[has_t2 ? l - 1 : 0];      // T2's block (l-1)
      const float* xprev = P.xbuf + (has_t2 ? ((l - 1) & 1) : 0) * P.Cr;   // x_{l-1} (x_0 for l=0)
      float* xout = P.xbuf + (l & 1) * P.Cr;                               // x_l
      const float* zprev = P.zbuf + ((l - 1) & 1) * Ch;                    // z_{l-1}
      float* zout = P.zbuf + (l & 1) * Ch;
      // ---- stage what depended on the previous phase: the current-tap entries and z_{l-1} ----
      float* vx = vx2 + (l & 1) * KXp;
      float* vc = vc2 + (l & 1) * Ccp;
      if (tagx && has_t2 && !cta_works) {
        // nothing to read: this CTA computes nothing in any phase
      } else if (tagx && has_t2) {
        // poll the producers' {value, tag} words of phase l-1 directly (no barrier in between)
        const uint32_t want = phase_tag(t, l - 1);
        const uint2* xt = P.xtag + ((l - 1) & 1) * P.Cr;
        const uint2* zt = P.ztag + ((l - 1) & 1) * Ch;
        // all of this thread's words are requested before the first one is looked at: one L2 round
        // trip for the lot instead of one per word (only words whose tag is still old are re-polled)
        constexpr int PQ = 6;
        const uint2* pp[PQ];
        float* dst[PQ];
        int np = 0;
        if (has_t1)
          for (int c = tid; c < P.Cr && np < PQ - 2; c += GEN_THREADS) {
            pp[np] = xt + c;
            dst[np++] = vx + c * P.fs + P.fs - 1;
          }
        const int nx = np;
        for (int i = tid; i < Ch && np < PQ; i += GEN_THREADS) {
          pp[np] = zt + i;
          dst[np++] = zs + i;
        }
        uint32_t pv[PQ], pg[PQ];
#pragma unroll
        for (int k = 0; k < PQ; ++k)
          if (k < np)
            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];"
                         : "=r"(pv[k]), "=r"(pg[k]) : "l"(pp[k]) : "memory");
#pragma unroll
        for (int k = 0; k < PQ; ++k)
          if (k < np) *dst[k] = (pg[k] == want) ? __uint_as_float(pv[k]) : ld_tagged(pp[k], want);
        // vectors longer than the batch (Cr > 4 * 256 or Ch > 2 * 256): the rest one by one
        if (has_t1)
          for (int c = tid + nx * GEN_THREADS; c < P.Cr; c += GEN_THREADS)
            vx[c * P.fs + P.fs - 1] = ld_tagged(xt + c, want);
        for (int i = tid + (np - nx) * GEN_THREADS; i < Ch; i += GEN_THREADS) zs[i] = ld_tagged(zt + i, want);
      } else {
        if (has_t1) {
          // l = 0: plain x_0 (just pushed into ring_0); l >= 1: x_{l-1} + br_{l-1} (xbuf)
          const float* cur = has_t2 ? xprev : P.queues + blk.qoff + (long long)s_slot[l] * P.Cr;
          for (int c = tid; c < P.Cr; c += GEN_THREADS) vx[c * P.fs + P.fs - 1] = __ldcg(cur + c);
        }
        if (has_t2)
          for (int i = tid; i < Ch; i += GEN_THREADS) zs[i] = __ldcg(zprev + i);
      }
      __syncthreads();
      if (rec && l < 12) P.dbg[8 * l + 1] = clock64();

      // ---- T1: gate pairs ----
      if (has_t1) {
        for (int pblk = blockIdx.x; pblk < npb; pblk += gridDim.x) {
          const int p = 2 * pblk + (warp >> 2), kc = warp & 3;
          float at = 0.0f, ag = 0.0f;
          if (fastp && p < Ch) {
            if (kc < 3) {
              const int k0 = kc * c3, n = max(0, min(c3, KX - k0));
              at = fma_row<4>(r0, vx + k0, n, lane);
              ag = fma_row<4>(r1, vx + k0, n, lane);
            } else {
              at = fma_row<4>(r0, vc, P.Cc, lane);
              ag = fma_row<4>(r1, vc, P.Cc, lane);
              if (has_t2) {
                at += fma_row<2>(m0, zs, Ch, lane);
                ag += fma_row<2>(m1, zs, Ch, lane);
              }
            }
            prefetch_phase(l + 1, true, false);   // r0/r1/m0/m1 are free again: next phase's rows
            at = warp_sum(at);
            ag = warp_sum(ag);
          } else if (p < Ch) {
            if (kc < 3) {
              const int k0 = kc * c3;
              const int n = max(0, min(c3, KX - k0));
              if (n > 0)
                dot2(blk.conv_w + (long long)p * KX + k0, blk.conv_w + (long long)(Ch + p) * KX + k0,
                     vx + k0, n, lane, at, ag);
            } else {
              dot2(blk.cond_w + (long long)p * P.Cc, blk.cond_w + (long long)(Ch + p) * P.Cc, vc, P.Cc,
                   lane, at, ag);
              if (has_t2) {
                float mt, mg;
                dot2(blk.mmat + (long long)p * Ch, blk.mmat + (long long)(Ch + p) * Ch, zs, Ch, lane,
                     mt, mg);
                at += mt;
                ag += mg;
              }
            }
          }
          if (lane == 0) {
            part[((warp >> 2) * 4 + kc) * 2 + 0] = at;
            part[((warp >> 2) * 4 + kc) * 2 + 1] = ag;
          }
          __syncthreads();
          if (tid < 2 && 2 * pblk + tid < Ch) {
            const int pp = 2 * pblk + tid;
            float ht = pb_t + pb_t2, hg = pb_g + pb_g2;
            if (pblk != (int)blockIdx.x) {     // only the first pair block of a CTA is prefetched
              ht = __ldg(blk.conv_b + pp) + __ldg(blk.cond_b + pp);
              hg = __ldg(blk.conv_b + Ch + pp) + __ldg(blk.cond_b + Ch + pp);
            }
            // the current tap saw x_{l-1} + br_{l-1}; c_l is NOT added (it is inside that sum)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ht += part[(tid * 4 + k) * 2 + 0];
              hg += part[(tid * 4 + k) * 2 + 1];
            }
            const float zv = tanhf(ht) * (1.0f / (1.0f + expf(-hg)));
            if (tagx) st_tagged(P.ztag + (l & 1) * Ch + pp, zv, phase_tag(t, l));
            else zout[pp] = zv;
          }
          __syncthreads();
        }
      }
      if (rec && l < 12) P.dbg[8 * l + 2] = clock64();
      // ---- T2: x_l = Wr z + br + x_{l-1} (pushed into ring_l), skip += Ws z + bs ----
      if (has_t2) {
        const int R = (has_t1 ? P.Cr : 0) + P.Cs;      // the last block's residual is unused
        const int r_off = has_t1 ? 0 : P.Cr;
        for (int r = r_t2; r < R; r += nwarps) {
          const int rr = r + r_off;
          const bool pre = fastp && r < nwarps;        // the first row of a warp was prefetched
          if (rr < P.Cr) {
            const float v = pre ? warp_sum(fma_row<2>(q4, zs, Ch, lane))
                                : dot_row(pb.res_w + (long long)rr * Ch, zs, Ch, lane);
            if (lane == 0) {
              // xbuf holds x_{l-1} + br_{l-1}; it gets x_l + br_l for the next phase
              const float xv = v + vx[rr * P.fs + P.fs - 1];   // x_{l-1} + br_{l-1}, staged above
              const float bnext = (r < nwarps) ? pb_row : __ldg(blk.res_b + rr);
              if (tagx) st_tagged(P.xtag + (l & 1) * P.Cr + rr, xv + bnext, phase_tag(t, l));
              else xout[rr] = xv + bnext;
              P.queues[blk.qoff + (long long)s_slot[l] * P.Cr + rr] = xv;
            }
          } else {
            const int sidx = rr - P.Cr;
            const float v = pre ? warp_sum(fma_row<2>(q4, zs, Ch, lane))
                                : dot_row(pb.skip_w + (long long)sidx * Ch, zs, Ch, lane);
            if (lane == 0) {
              const float add = v + ((r < nwarps) ? pb_row : __ldg(pb.skip_b + sidx));
              if (skreg && has_t1) {
                skip_reg += add;
                if (l == P.n_blocks - 1) {        // last phase in which this warp owns the row
                  P.skipacc[sidx] = skip_reg;
                  skip_reg = 0.0f;
                }
              } else if (tagx) {
                P.skiplast[sidx] = add;          // last block's rows: summed by the head
              } else {
                P.skipacc[sidx] += add;
              }
            }
          }
        }
      }
      if (rec && l < 12) P.dbg[8 * l + 3] = clock64();
      if (gfast) {
        gather_commit(l + 1);          // loaded during the previous phase
        gather_issue(l + 2, t);
      } else {
        gather_past(l + 1, t);
      }
      prefetch_phase(l + 1, false, true);
      if (rec && l < 12) P.dbg[8 * l + 4] = clock64();
      if (!tagx || l == P.n_blocks) grid_barrier(P.barrier, epoch);
      if (rec && l < 12) P.dbg[8 * l + 5] = clock64();
    }

    // ---- head: relu -> proj1 -> relu -> proj2 (modules.py:248-254) ----
    if (rec) P.dbg[100] = clock64();
    for (int i = tid; i < P.Cs; i += GEN_THREADS)
      xs[i] = fmaxf(__ldcg(P.skipacc + i) + (tagx ? __ldcg(P.skiplast + i) : 0.0f), 0.0f);
    __syncthreads();
    for (int r = gwarp; r < P.Cs; r += nwarps) {
      float v = dot_row(P.proj1_w + (long long)r * P.Cs, xs, P.Cs, lane);
      if (lane == 0) P.h1[r] = fmaxf(v + __ldg(P.proj1_b + r), 0.0f);
    }
    grid_barrier(P.barrier, epoch);
    if (rec) P.dbg[101] = clock64();
    for (int i = tid; i < P.Cs; i += GEN_THREADS) xs[i] = __ldcg(P.h1 + i);
    __syncthreads();
    for (int r = gwarp; r < P.Q; r += nwarps) {
      float v = dot_row(P.proj2_w + (long long)r * P.Cs, xs, P.Cs, lane);
      if (lane == 0) P.logit_buf[r] = v + __ldg(P.proj2_b + r);
    }
    grid_barrier(P.barrier, epoch);

    // ---- softmax + draw: every CTA does it redundantly (identical inputs -> identical result)
    if (rec) P.dbg[102] = clock64();
    for (int i = tid; i < P.Q; i += GEN_THREADS) xs[i] = __ldcg(P.logit_buf + i);
    __syncthreads();
    if (warp == 0 && P.mol) {
      // generate.py:116-137: one logistic draw from EVERY component, mixed by the softmax weights
      // (float32 softmax, float64 draw and sum, cast to float32, / 127.5, clip to [-1, 1])
      const int nr = P.Q / 3;
      float m = -INFINITY;
      for (int i = lane; i < nr; i += 32) m = fmaxf(m, xs[i]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      float s = 0.0f;
      for (int i = lane; i < nr; i += 32) {
        const float e = expf(xs[i] - m);
        ps[i] = e;
        s += e;
      }
      s = warp_sum(s);
      __syncwarp();
      // the float64 logarithms of the components are evaluated by one lane each; the terms are then
      // added in component order (the order of the reference's sum), so the result is unchanged
      double acc = 0.0;
      for (int k0 = 0; k0 < nr; k0 += 32) {
        const int k = k0 + lane;
        double term = 0.0;
        if (k < nr) {
          const float sc = expf(fmaxf(xs[2 * nr + k], P.log_scale_min));
          const double u = P.uniforms[(long long)step * nr + k];
          const double r = (double)xs[nr + k] + (double)sc * (log(u) - log(1.0 - u));
          term = r * (double)(ps[k] / s);
        }
        const int nk = min(32, nr - k0);
        for (int j = 0; j < nk; ++j) acc += __shfl_sync(0xffffffffu, term, j);
      }
      if (lane == 0) {
        float v = (float)acc;
        v = v / 127.5f;
        v = fminf(fmaxf(v, -1.0f), 1.0f);
        s_sample = __float_as_int(v);
      }
    } else if (warp == 0) {
      float m = -INFINITY;
      for (int i = lane; i < P.Q; i += 32) m = fmaxf(m, xs[i]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      float s = 0.0f;
      for (int i = lane; i < P.Q; i += 32) {
        float e = expf(xs[i] - m);
        ps[i] = e;
        s += e;
      }
      s = warp_sum(s);
      __syncwarp();
      for (int i = lane; i < P.Q; i += 32) ps[i] = ps[i] / s;     // the float32 softmax, in parallel
      __syncwarp();
      {
        // numpy.random.choice: cdf = cumsum(p as float64); cdf /= cdf[-1]; searchsorted(u, 'right'),
        // i.e. the first i with run_i / tot > u.  One lane doing it element by element cost 65 k
        // cycles per sample (a float64 division per element, then still 42 k: a dependent float64
        // add is ~64 cycles here), a fifth of the whole step.  Now every lane sums a contiguous run
        // of Q/32 elements, a warp scan gives the prefixes, and each lane tests its own elements;
        // run_i <= u * tot * (1 - 2^-40) rules an element out without dividing.  The float64 sums
        // are associated differently from NumPy's sequential cumsum (last-ulp differences, ~1e-16):
        // the draw can only differ when u lies within that distance of a cdf step.
        const int per = (P.Q + 31) / 32, e0 = lane * per;
        double loc = 0.0;
        for (int k = 0; k < per; ++k)
          if (e0 + k < P.Q) loc += (double)ps[e0 + k];
        double inc = loc;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, inc, off);
          if (lane >= off) inc += o;
        }
        const double tot = __shfl_sync(0xffffffffu, inc, 31);
        double run = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) run = 0.0;
        const double u = P.uniforms[step];
        const double lo = u * tot * (1.0 - 9.094947017729282e-13);
        int mine = P.Q;
        for (int k = 0; k < per; ++k) {
          if (e0 + k >= P.Q) break;
          run += (double)ps[e0 + k];
          if (run > lo && run / tot > u) { mine = e0 + k; break; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, off));
        if (lane == 0) s_sample = mine < P.Q ? mine : P.Q - 1;
      }
    }
    __syncthreads();
    if (rec) P.dbg[103] = clock64();
    const int pick = s_sample;
    const int fed = P.forced ? P.forced[step] : pick;
    if (blockIdx.x == 0) {
      if (P.logits)
        for (int i = tid; i < P.Q; i += GEN_THREADS) P.logits[(long long)step * P.Q + i] = xs[i];
      if (tid == 0) P.samples[step] = pick;
    }
    // logit_buf, h1 and skipacc are next written behind at least two grid barriers of the next step
    // (initial barrier, phase-n barrier), which every CTA only passes after the reads above
    st2 = st1;
    st1 = fed;
    if (rec) P.dbg[104] = clock64();
  }
  if (blockIdx.x == 0 && tid == 0) {   // for the step-wise API (WaveNetState.step)
    P.state[0] = st1;
    P.state[1] = st2;
  }
}

}  // namespace vqw

static inline int64_t gen_align(int64_t v) { return (v + 255) / 256 * 256; }

struct GenLayout {
  int64_t blocks, queues, zbuf, mmat, xbuf, skipacc, skiplast, xtag, ztag, h1, logit, state, barrier, total;
};

static GenLayout gen_layout(const vqw_generate_desc& d) {
  GenLayout L;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += gen_align(bytes); return o; };
  L.blocks = take((int64_t)d.n_blocks * sizeof(vqw::GenBlock));
  int64_t q = 0;
  for (int i = 0; i < d.n_blocks; ++i) q += (int64_t)(d.dilations[i] * (d.fs - 1) + 1) * d.Cr;
  L.queues = take(q * 4);
  L.zbuf = take((int64_t)2 * (d.Cd / 2) * 4);
  L.mmat = take((int64_t)d.n_blocks * d.Cd * (d.Cd / 2) * 4);
  L.xbuf = take((int64_t)2 * d.Cr * 4);
  L.skipacc = take((int64_t)d.Cs * 4);
  L.skiplast = take((int64_t)d.Cs * 4);
  L.xtag = take((int64_t)2 * d.Cr * 8);
  L.ztag = take((int64_t)2 * (d.Cd / 2) * 8);
  L.h1 = take((int64_t)d.Cs * 4);
  L.logit = take((int64_t)d.Q * 4);
  L.state = take(16);
  L.barrier = take(16);
  L.total = off + 256;
  return L;
}

extern "C" int64_t vqw_generate_workspace(const vqw_generate_desc* desc) {
  if (!desc || !desc->dilations || desc->n_blocks < 1) return -1;
  return gen_layout(*desc).total;
}

extern "C" int vqw_generate(const vqw_generate_desc* desc, const vqw_resblock_weights* blocks,
                            const float* embed_w, const float* embed_b, const float* proj1_w,
                            const float* proj1_b, const float* proj2_w, const float* proj2_b,
                            const float* cond, const double* uniforms, const int32_t* forced,
                            int32_t* samples, float* logits, void* workspace,
                            vqw_stream_t stream) {
  using namespace vqw;
  VQW_REQUIRE(desc && blocks, "vqw_generate: null descriptor");
  const vqw_generate_desc& d = *desc;
  VQW_REQUIRE(d.n_blocks >= 1 && d.dilations, "vqw_generate: n_blocks/dilations");
  VQW_REQUIRE(d.fs >= 1 && d.Cr > 0 && d.Cd > 0 && d.Cd % 2 == 0 && d.Cs > 0 && d.Cc > 0 && d.Q > 0,
              "vqw_generate: bad channel counts");
  VQW_REQUIRE(d.n_steps >= 0 && d.t_start >= 0 && d.t_start - d.cond_t0 >= 0 &&
                  d.t_start - d.cond_t0 + d.n_steps <= d.T_total,
              "vqw_generate: steps [%d, %d) are outside the condition columns [%d, %d)", d.t_start,
              d.t_start + d.n_steps, d.cond_t0, d.cond_t0 + d.T_total);
  if (d.n_steps == 0) return 0;
  VQW_REQUIRE(embed_w && embed_b && proj1_w && proj1_b && proj2_w && proj2_b && cond && uniforms &&
                  samples && workspace, "vqw_generate: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const GenLayout L = gen_layout(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(gen_align((int64_t)(uintptr_t)workspace));

  // block table (device copy in the workspace)
  static thread_local GenBlock host_blocks[256];
  VQW_REQUIRE(d.n_blocks <= 256, "vqw_generate: more than 256 blocks");
  long long qoff = 0;
  for (int i = 0; i < d.n_blocks; ++i) {
    const vqw_resblock_weights& w = blocks[i];
    VQW_REQUIRE(w.conv_w && w.conv_b && w.cond_w && w.cond_b && w.res_w && w.res_b && w.skip_w &&
                    w.skip_b, "vqw_generate: block %d has a null weight", i);
    GenBlock& g = host_blocks[i];
    g.conv_w = w.conv_w; g.conv_b = w.conv_b; g.cond_w = w.cond_w; g.cond_b = w.cond_b;
    g.res_w = w.res_w; g.res_b = w.res_b; g.skip_w = w.skip_w; g.skip_b = w.skip_b;
    g.dilation = d.dilations[i];
    g.qlen = d.dilations[i] * (d.fs - 1) + 1;
    g.qoff = qoff;
    qoff += (long long)g.qlen * d.Cr;
    g.mmat = (i == 0) ? nullptr
                      : reinterpret_cast<const float*>(ws + L.mmat) + (long long)i * d.Cd * (d.Cd / 2);
    g.cvec = nullptr;
  }
  VQW_CHECK_CUDA(cudaMemcpyAsync(ws + L.blocks, host_blocks, sizeof(GenBlock) * d.n_blocks,
                                 cudaMemcpyHostToDevice, st));
  if (d.t_start == 0) {
    // M_l = Wc_cur_l Wr_{l-1}: a (Cd x Cr) by (Cr x Cd/2) product per block boundary, through the
    // fp32 conv kernel (Wr viewed as a (Cr channels, Cd/2 steps) signal, Wc_cur as a 1x1 filter)
    for (int i = 1; i < d.n_blocks; ++i) {
      vqw_conv_desc c = {};
      c.B = 1; c.M = d.Cd; c.T = d.Cd / 2;
      c.nsrc = 1;
      c.src[0] = {blocks[i - 1].res_w, blocks[i].conv_w + (d.fs - 1), nullptr, d.Cr, d.Cd / 2,
                  d.Cr * d.fs, d.fs, 1, 0, 1, 0};
      if (int rc = launch_conv(c, reinterpret_cast<float*>(ws + L.mmat) + (long long)i * d.Cd * (d.Cd / 2),
                               st))
        return rc;
    }
    // WaveNet.initialize(): zero queues (modules.py:59-66,236-243); no previous samples
    VQW_CHECK_CUDA(cudaMemsetAsync(ws + L.queues, 0, (size_t)qoff * 4, st));
    // tag 0 is never expected: stale words of a previous utterance must not match
    VQW_CHECK_CUDA(cudaMemsetAsync(ws + L.xtag, 0, (size_t)2 * d.Cr * 8, st));
    VQW_CHECK_CUDA(cudaMemsetAsync(ws + L.ztag, 0, (size_t)2 * (d.Cd / 2) * 8, st));
    // no previous samples: index -1 (categorical) / value 0.0 (mixture of logistics), generate.py:51
    VQW_CHECK_CUDA(cudaMemsetAsync(ws + L.state, d.use_logistic ? 0 : 0xff, 16, st));
  }
  if (d.set_state) {
    static thread_local int32_t hs[4];
    hs[0] = d.s1; hs[1] = d.s2; hs[2] = hs[3] = -1;
    VQW_CHECK_CUDA(cudaMemcpyAsync(ws + L.state, hs, 16, cudaMemcpyHostToDevice, st));
  }

  GenParams P;
  P.n_blocks = d.n_blocks; P.fs = d.fs; P.Cr = d.Cr; P.Cd = d.Cd; P.Cs = d.Cs; P.Cc = d.Cc; P.Q = d.Q;
  P.T_total = d.T_total; P.n_steps = d.n_steps; P.t_start = d.t_start; P.cond_t0 = d.cond_t0;
  P.mol = d.use_logistic ? 1 : 0;
  P.log_scale_min = d.log_scale_min;
  VQW_REQUIRE(!d.use_logistic || (d.Q % 3 == 0 && d.Q >= 3),
              "vqw_generate: use_logistic needs Q = 3 * n_mixtures output channels (got %d)", d.Q);
  P.blocks = reinterpret_cast<const GenBlock*>(ws + L.blocks);
  P.embed_w = embed_w; P.embed_b = embed_b; P.proj1_w = proj1_w; P.proj1_b = proj1_b;
  P.proj2_w = proj2_w; P.proj2_b = proj2_b;
  P.cond = cond; P.uniforms = uniforms; P.forced = forced; P.samples = samples; P.logits = logits;
  P.queues = reinterpret_cast<float*>(ws + L.queues);
  P.zbuf = reinterpret_cast<float*>(ws + L.zbuf);
  P.xbuf = reinterpret_cast<float*>(ws + L.xbuf);
  P.skipacc = reinterpret_cast<float*>(ws + L.skipacc);
  P.skiplast = reinterpret_cast<float*>(ws + L.skiplast);
  P.xtag = reinterpret_cast<uint2*>(ws + L.xtag);
  P.ztag = reinterpret_cast<uint2*>(ws + L.ztag);
  P.h1 = reinterpret_cast<float*>(ws + L.h1);
  P.logit_buf = reinterpret_cast<float*>(ws + L.logit);
  P.state = reinterpret_cast<int32_t*>(ws + L.state);
  P.barrier = reinterpret_cast<unsigned int*>(ws + L.barrier);
  VQW_CHECK_CUDA(cudaMemsetAsync(ws + L.barrier, 0, 16, st));
  static long long* gen_dbg = nullptr;
  static const bool gen_timeline = getenv("VQW_GEN_TIMELINE") && getenv("VQW_GEN_TIMELINE")[0] == '1';
  P.dbg = nullptr; P.dbg_cta = 0; P.dbg_step = 0;
  if (gen_timeline && d.n_steps > 8) {
    if (!gen_dbg) cudaMalloc(&gen_dbg, 128 * sizeof(long long));
    cudaMemsetAsync(gen_dbg, 0, 128 * sizeof(long long), st);
    P.dbg = gen_dbg;
    P.dbg_cta = getenv("VQW_GEN_TIMELINE_CTA") ? atoi(getenv("VQW_GEN_TIMELINE_CTA")) : 0;
    P.dbg_step = 8;
  }

  const int KX = d.fs * d.Cr, Ch = d.Cd / 2;
  int mx = d.Cr > d.Cs ? d.Cr : d.Cs;
  if (d.Q > mx) mx = d.Q;
  size_t smem = sizeof(float) * (2 * ((KX + 3) & ~3) + 2 * ((d.Cc + 3) & ~3) + ((Ch + 3) & ~3) +
                                 ((mx + 3) & ~3) + ((d.Q + 3) & ~3) + 16 + 8) +
                sizeof(GenBlock) * d.n_blocks;
  VQW_REQUIRE(smem <= 200 * 1024, "vqw_generate: channel counts too large for shared memory");
  VQW_CHECK_CUDA(cudaFuncSetAttribute(generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  int dev = 0, sms = 0, per_sm = 0, coop = 0;
  VQW_CHECK_CUDA(cudaGetDevice(&dev));
  VQW_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  VQW_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  VQW_REQUIRE(coop, "vqw_generate: device does not support cooperative launch");
  VQW_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generate_kernel, GEN_THREADS,
                                                              smem));
  VQW_REQUIRE(per_sm >= 1, "vqw_generate: kernel does not fit on an SM");
  int grid = sms;   // one CTA per SM
  {
    // Tagged exchange needs every CTA that CONSUMES the exchanged vectors in a phase to also
    // PRODUCE a tagged word in it (that is what bounds how far CTAs can drift apart, see
    // st_tagged): true when no CTA holds skip rows only.  Otherwise: grid barriers.
    const int Ch = d.Cd / 2, nw = grid * (GEN_THREADS / 32), npb = (Ch + 1) / 2;
    const int t2_first = (npb * (GEN_THREADS / 32)) % nw;
    bool ok = !(getenv("VQW_GEN_TAGS") && getenv("VQW_GEN_TAGS")[0] == '0') &&
              (d.Cr + d.Cs) <= nw && d.n_blocks >= 2 && d.n_blocks < 500 && npb <= grid;
    for (int b = 0; b < grid && ok; ++b) {
      bool t1 = b < npb, xrow = false, srow = false;
      for (int w = 0; w < GEN_THREADS / 32; ++w) {
        const int r = (b * (GEN_THREADS / 32) + w - t2_first + nw) % nw;
        if (r < d.Cr) xrow = true;
        else if (r < d.Cr + d.Cs) srow = true;
      }
      if (srow && !t1 && !xrow) ok = false;
    }
    P.use_tags = ok ? 1 : 0;
  }
  // Queue residency.  The dilation queues of the configs[4] decoder are 16.8 MB (40 blocks x up to
  // 1025 columns x 512 channels fp32): they cannot live in shared memory / registers (148 SMs x
  // 227 KB = 33 MB in total, but every CTA of a phase needs the WHOLE 512-channel past column, and
  // distributed shared memory only spans a cluster of <= 16 CTAs = 3.6 MB).  They are kept resident
  // ON CHIP instead: an L2 access-policy window marks the queues and the re-associated matrices
  // M_l (38 MB together, < the 126 MB L2) persisting, while the 175 MB weight stream -- which
  // would otherwise evict them every sample -- is left evict-first.  VQW_GEN_L2PERSIST=0 disables it.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  int nattr = 1;
  const bool persist = !(getenv("VQW_GEN_L2PERSIST") && getenv("VQW_GEN_L2PERSIST")[0] == '0');
  if (persist) {
    int max_win = 0, max_persist = 0;
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    size_t want = (size_t)(L.xbuf - L.queues);            // queues, z buffer, M_l: contiguous
    if (max_win > 0 && max_persist > 0) {
      if (want > (size_t)max_win) want = (size_t)max_win;
      static size_t limit_set = 0;                        // device-wide set-aside, grown on demand
      const size_t lim = want < (size_t)max_persist ? want : (size_t)max_persist;
      if (lim > limit_set && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess)
        limit_set = lim;
      attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[1].val.accessPolicyWindow.base_ptr = ws + L.queues;
      attr[1].val.accessPolicyWindow.num_bytes = want;
      attr[1].val.accessPolicyWindow.hitRatio = limit_set >= want ? 1.0f : (float)limit_set / (float)want;
      attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      nattr = 2;
    }
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  VQW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, generate_kernel, P));
  VQW_CHECK_LAUNCH("generate_kernel");
  if (P.dbg) {
    long long h[128];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, gen_dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[vqw generate timeline] CTA %d, step %d (cycles rel. to the start of phase 0)\n",
            P.dbg_cta, P.dbg_step);
    for (int l = 0; l < 12; ++l)
      fprintf(stderr, "  phase %2d: start %6lld staged %6lld T1 %6lld T2 %6lld prefetch issued %6lld "
                      "barrier passed %6lld\n", l, h[8 * l] - h[0], h[8 * l + 1] - h[0],
              h[8 * l + 2] - h[0], h[8 * l + 3] - h[0], h[8 * l + 4] - h[0], h[8 * l + 5] - h[0]);
    fprintf(stderr, "  step start %lld | phases done / head start %lld, proj1 + barrier %lld, proj2 + barrier %lld, "
                    "softmax + draw %lld, sample hand-over %lld\n", h[105] - h[0], h[100] - h[0],
            h[101] - h[100], h[102] - h[101], h[103] - h[102], h[104] - h[103]);
  }
  return 0;
}
