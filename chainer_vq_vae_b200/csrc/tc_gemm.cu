// Tensor-core (tcgen05 / TMEM / TMA) GEMM kernels of the residual-block BACKWARD pass
// (the reference leaves this to Chainer autograd over modules.py:30-56).
//
// One kernel template, four epilogues.  Every operand is a TIME-major (B, T, C) bf16 hi/lo plane
// (the layout the forward writes, see resblock_tc.cu); every product is 3 MMAs in bf16x3 mode.
//
//   "time" flavour -- tile rows = 128 time steps of one batch item, K = channels (K-major tiles,
//   64-byte swizzle, exactly like the forward):
//     EPI_GATE_BWD  gz = Wr^T g_res + Ws^T g_skip (K = Cr + Cs), then the gate derivative with
//                   the saved tanh/sigmoid -> gh planes
//     EPI_GX        gx[t] = g_res[t] + sum_j Wc_j^T gh[t + dil*(fs-1-j)]  (anti-causal taps are
//                   row shifts of the TMA box; rows past T are out of bounds = zero) -> planes of
//                   the next g_res (and fp32 (B,Cr,T) for the gradient handed back to autograd)
//     EPI_ACCUM     gcond += Wp^T gh   (fp32 (B,Cc,T), read-modify-write)
//     EPI_HEAD      out = [mask] [relu] (acc + bias) -> planes and/or fp32 (B,C,T): the 1x1
//                   projections of the WaveNet head (modules.py:155-159) and their data gradients
//   "wgrad" flavour -- tile rows = 128 output channels, K = time.  Time is the ROW axis of the
//   planes, so both operands are MN-major tiles (128-byte swizzle, TMA boxes {64 channels, 32
//   steps}); the per-tap delay is a row coordinate.  All weight gradients of a block -- fs conv
//   taps, condition projection, res and skip 1x1 -- are ONE grouped launch: a job table maps
//   blockIdx.x to (operand planes, tile origin, delay, destination), blockIdx.z splits K over
//   batch items; fp32 atomics accumulate into the gradient tensors.
//
// CTA layout as in the forward kernel (warp 8 = TMA producer, warp 9 = MMA issuer, warps 0-7 =
// epilogue, two per TMEM lane quadrant) with a 2-stage ring and a 256-column accumulator so
// that TWO CTAs are resident per SM: one CTA's epilogue overlaps the other's MMAs.
#include "tc_common.cuh"
#include "mol.cuh"
#include <stdlib.h>

namespace vqw {
namespace tc {

constexpr int G_STAGES = 2;
constexpr int G_EPI_WARPS = 8;
constexpr int G_THREADS = (G_EPI_WARPS + 2) * 32;
constexpr int GW_TMA = G_EPI_WARPS, GW_MMA = G_EPI_WARPS + 1;
constexpr int MAX_SEG = 4;
constexpr int MAX_JOBS = 48;
constexpr int NMAPS = 12;
enum { EPI_GATE_BWD = 0, EPI_GX = 1, EPI_ACCUM = 2, EPI_WGRAD = 3, EPI_HEAD = 4, EPI_HEAD_LOSS = 5 };

struct Seg {
  int a_map, b_map;   // tensor-map pair index: hi plane = maps[2*i], lo plane = maps[2*i+1]
  int nslabs;         // K / 32
  int a_c0, b_c0;     // coordinate along the contiguous (K) axis of slab 0
  int a_shift;        // added to the time-row coordinate of A
  int b_row0;         // first B row; + 256 * blockIdx.y
};

struct Job {          // one 128 x 256 tile of one weight-gradient GEMM
  int a_map, b_map;   // plane pairs: A rows = output channels of the gradient, B rows = its inputs
  int m0, n0;         // tile origin (channels)
  int shift;          // B is read at time t + shift (shift = -delay of the tap)
  int M, N;           // valid extent of the gradient matrix
  float* out;         // gW base (already offset to the tap)
  long long gm, gk;   // strides along m / n
  // optional column col_n of the product (a constant-one channel of B, see cond_pitch): the sum
  // over time of A's rows for THIS batch item, accumulated into col_out[b*col_stride + m] (zeroed
  // by the caller: the time axis may be split over several tiles)
  float* col_out;
  int col_n;
  long long col_stride;
};

struct GemmParams {
  int nseg;
  Seg seg[MAX_SEG];
  int x3;
  int f16;                     // planes hold IEEE fp16 (VQW_MODE_FP16: single pass) instead of bf16
  int add_lo;                  // GX: the g_res stream keeps a lo plane (addend read / result write)
  const float* scale;          // fp16 backward: device {s, 1/s}, the power-of-two gradient scale the
                               // operand planes carry; fp32 outputs are multiplied by 1/s (or null)
  int B, T;
  int b_exact;                 // wgrad: B has no lo plane (exact bf16 values, e.g. a one-hot)
  int a_exact;                 // wgrad: A's lo plane is not contracted (hi-plane passes, see wgrad_passes)
  int wg_sub;                  // persistent wgrad: slabs of 32 time steps per ring stage (1 or 2);
                               // slabs_per_item counts stages
  int dbg_mode;                // persistent wgrad, timing experiments only (VQW_WGRAD_DEBUG): 1 = no
                               // operand loads, 2 = no MMAs, 3 = MMAs without barriers -- garbage results
  unsigned wide_maps;          // persistent wgrad: bit i = plane pair i has 4-D maps (make_map_mn4):
                               // one copy brings both 64-channel groups of a 128-channel operand half
  float* wg_partial;           // persistent wgrad: every (item, time chunk, job pair) tile is stored
                               // here as [256 rows][256 cols] fp32 instead of fp32 atomics into gW;
                               // wgrad_reduce_kernel sums them in a fixed order afterwards
  const float* plane_s;        // head + loss: device scalar the gradient planes are multiplied by
                               // (fp16 planes: a power of two near the label count), or null
  int slabs_per_item;          // wgrad: K slabs per (batch item, time chunk) work item
  int chunks_per_b;            // wgrad: work items per batch item
  const float* f0;             // GATE_BWD: tanh, time-major (B,T,256) fp32, or null: tanh = z / sigmoid
  const __nv_bfloat16* z_hi;   // GATE_BWD with f0 == null: the saved z planes (B,T,256)
  const __nv_bfloat16* z_lo;
  const float* f1;             // GATE_BWD: sigmoid
  const __nv_bfloat16* a_hi;   // GX: addend planes (B,T,Cout) (g_res) or null
  const __nv_bfloat16* a_lo;
  float* o0;                   // GX: fp32 (B,Cout,T) output or null | ACCUM: gcond (B,Cout,T)
  __nv_bfloat16* p_hi;         // GATE_BWD / GX: output planes (B,T,C) or null
  __nv_bfloat16* p_lo;
  int Cout;
  const float* bias;           // HEAD: per-output-channel bias or null
  int relu;                    // HEAD: relu on the result
  const __nv_bfloat16* mask_hi;  // HEAD: (B,T,Cout) plane; result is zeroed where mask <= 0
  // GX launches can carry the ACCUM tiles of the same block as extra blockIdx.y values
  // (>= alt_y): those CTAs stream gh once more (from L2: the GX tiles of the same rows read the
  // same slabs) and are bound by that stream, so they overlap the MMA-bound GX tiles instead of
  // running as a separate launch
  // HEAD_LOSS: the loss and d loss / d y in the epilogue of the proj2 GEMM (Cout = Q <= 256: the
  // whole row of logits is one accumulator tile) -- the logits never go to HBM unless o0 is set
  const int32_t* tgt_i;        // (B,T) class labels (softmax cross entropy), or null
  const float* tgt_f;          // (B,T) waveform values (mixture of logistics), or null
  double* loss;                // {loss (accumulated), number of valid labels (set before the launch)}
  int Qv;                      // real number of output channels (<= Cout = padded plane pitch)
  float mol_half, mol_lsmin;   // 127.5 / (quantize - 1), log_scale_min
  float* colsum;               // GX: column sums of the result (+= , the bias gradient
                               // sum_t g_res of the NEXT block to run) or null
  int alt_y;                   // 0 = no ACCUM tiles in this launch
  Seg alt_seg;
  float* alt_out;              // gcond (B, alt_Cout, T)
  int alt_Cout;
  int stage_epi;               // GX / GATE_BWD: epilogue through shared-memory tiles + TMA (see below)
  long long* dbg;              // optional per-CTA timestamps (VQW_GEMM_TIMELINE=<epilogue id>)
  int njobs;
  Job jobs[MAX_JOBS];
};

struct Maps {
  CUtensorMap m[NMAPS];
};

// PAIR = 1: the same GEMM on a cluster of two CTAs (two consecutive time tiles, or the two
// 128-row halves of a 256-row weight-gradient tile): cta_group::2 MMAs with M = 256, each CTA
// stages its own A rows and HALF of the B tile -- 32 KB instead of 48 KB of L2 -> SM traffic per K
// slab (the fill, not the tensor pipe, bounded these kernels: see resblock_tc.cu), in a 3-stage
// ring of the same shared-memory footprint.  Protocol as in resblock_tc_pair_kernel.
__device__ __forceinline__ long long gtime_ns() {   // chip-wide clock (ns): comparable across SMs
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int EPI, int PAIR>
__global__ void __launch_bounds__(G_THREADS, 2)
tc_gemm_kernel(const __grid_constant__ Maps maps, const __grid_constant__ GemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  constexpr int NST = PAIR ? 3 : G_STAGES;              // ring stages
  constexpr int BPL = PAIR ? B_PLANE / 2 : B_PLANE;     // bytes of one B plane this CTA stages per slab
  constexpr int STG = 2 * A_PLANE + 2 * BPL;            // 32 KB / 48 KB
  constexpr int BROWS = PAIR ? TN / 2 : TN;             // B rows (time flavour) this CTA stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * STG);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NST);
  const uint32_t acc_full = smem_u32(bars + 2 * NST);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);
  float* csum = reinterpret_cast<float*>(bars + 2 * NST + 2);   // [TN] column sums (GX)
  const uint32_t afull0 = smem_u32(csum + TN);                   // [2] staged-epilogue tiles landed
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader of the pair

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool WG = (EPI == EPI_WGRAD);
  // timeline of 64 CTAs: blockIdx.x in [0,32) of (y,z) = (0,0) and (0,1): {entry, setup, first slab
  // full, MMAs issued, accumulator full seen by warp 0, epilogue done, smid}
  long long* dbg = nullptr;
  if (P.dbg != nullptr && blockIdx.x < 32 && blockIdx.y == 0 && blockIdx.z < 2)
    dbg = P.dbg + (blockIdx.z * 32 + blockIdx.x) * 8;
  if (dbg && threadIdx.x == 0) {
    dbg[0] = gtime_ns();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    dbg[6] = smid;
  }
  if (EPI == EPI_GX && P.colsum != nullptr)
    for (int i = threadIdx.x; i < TN; i += G_THREADS) csum[i] = 0.0f;
  const int nplanes = P.x3 ? 2 : 1;

  const bool alt = (EPI == EPI_GX) && P.alt_y > 0 && (int)blockIdx.y >= P.alt_y;
  const int by = alt ? (int)blockIdx.y - P.alt_y : (int)blockIdx.y;   // N tile within its GEMM
  int total_slabs = 0;
  if (alt) {
    total_slabs = P.alt_seg.nslabs;
  } else if (WG) {
    const int items = P.B * P.chunks_per_b;
    int n_items = 0;
    for (int c = blockIdx.z; c < items; c += gridDim.z) ++n_items;
    total_slabs = n_items * P.slabs_per_item;
  } else {
    for (int s = 0; s < P.nseg; ++s) total_slabs += P.seg[s].nslabs;
  }

  if (warp == GW_TMA && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(afull0, 1);
    mbar_init(afull0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GW_MMA) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_slot)),
                   "r"(256)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_slot)),
                   "r"(256)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) dbg[1] = gtime_ns();
  // TMA copy into this CTA's stage; in a pair every copy completes on the LEADER's barrier
  auto ld3 = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    if (PAIR) tma2_load_3d(dst, m, bar, c0, c1, c2);
    else tma_load_3d(dst, m, bar, c0, c1, c2);
  };

  if (warp == GW_TMA) {
    // =============================== TMA producer ===============================
    // whole warp in uniform control flow, one elected lane issues (see tc_time_persistent_kernel)
    if (total_slabs > 0) {
      const bool el = elect_one_sync();
      int stage = 0;
      uint32_t ph = 0;
      if (WG) {
        // a stage = 2 (A) + 4 (B) boxes of {64 channels, 32 time steps} per plane
        const Job& jb = P.jobs[blockIdx.x];
        const CUtensorMap* ma = &maps.m[2 * jb.a_map];
        const CUtensorMap* mb = &maps.m[2 * jb.b_map];
        prefetch_tmap(ma); prefetch_tmap(mb);
        const int items = P.B * P.chunks_per_b;
        for (int c = blockIdx.z; c < items; c += gridDim.z) {
          const int bb = c / P.chunks_per_b;
          const int tk = (c % P.chunks_per_b) * P.slabs_per_item * BK;
          for (int i = 0; i < P.slabs_per_item; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = PAIR ? mapa_u32(full0 + 8 * stage, 0) : full0 + 8 * stage;
            const uint32_t sa = base + stage * STG;
            const bool blo = P.x3 && !P.b_exact, alo = P.x3 && !P.a_exact;
            if (rank == 0 && el)
              mbar_expect_tx(full0 + 8 * stage,
                             (PAIR ? 2 : 1) * ((alo ? 2 : 1) * A_PLANE + (blo ? 2 : 1) * BPL));
            const int ta = tk + i * BK, tb = ta + jb.shift;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (el) ld3(sa + h * 4096, ma, fb, jb.m0 + 64 * h, ta, bb);
              if (alo && el) ld3(sa + A_PLANE + h * 4096, ma + 1, fb, jb.m0 + 64 * h, ta, bb);
            }
            // the B tile is 256 input channels; a CTA of a pair stages its 128-channel half
            const int nb0 = jb.n0 + (PAIR ? (int)rank * (TN / 2) : 0);
#pragma unroll
            for (int h = 0; h < BROWS / 64; ++h) {
              if (el) ld3(sa + 2 * A_PLANE + h * 4096, mb, fb, nb0 + 64 * h, tb, bb);
              if (blo && el) ld3(sa + 2 * A_PLANE + BPL + h * 4096, mb + 1, fb, nb0 + 64 * h, tb, bb);
            }
            if (++stage == NST) { stage = 0; ph ^= 1; }
          }
        }
      } else {
        const int t0 = blockIdx.x * TM, bb = blockIdx.z;
        const int nseg = alt ? 1 : P.nseg;
        for (int s = 0; s < nseg; ++s) {
          const Seg& sg = alt ? P.alt_seg : P.seg[s];
          const CUtensorMap* ma = &maps.m[2 * sg.a_map];
          const CUtensorMap* mb = &maps.m[2 * sg.b_map];
          if (s == 0) { prefetch_tmap(ma); prefetch_tmap(mb); }
          for (int i = 0; i < sg.nslabs; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = PAIR ? mapa_u32(full0 + 8 * stage, 0) : full0 + 8 * stage;
            const uint32_t sa = base + stage * STG;
            if (rank == 0 && el) mbar_expect_tx(full0 + 8 * stage, (PAIR ? 2 : 1) * nplanes * (A_PLANE + BPL));
            // this CTA's half (pair) of the 256 weight rows of the N tile
            const int brow = sg.b_row0 + TN * by + (PAIR ? (int)rank * (TN / 2) : 0);
            if (el) ld3(sa, ma, fb, sg.a_c0 + i * BK, t0 + sg.a_shift, bb);
            if (el) ld3(sa + 2 * A_PLANE, mb, fb, sg.b_c0 + i * BK, brow, 0);
            if (P.x3) {
              if (el) ld3(sa + A_PLANE, ma + 1, fb, sg.a_c0 + i * BK, t0 + sg.a_shift, bb);
              if (el) ld3(sa + 2 * A_PLANE + BPL, mb + 1, fb, sg.b_c0 + i * BK, brow, 0);
            }
            if (++stage == NST) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == GW_MMA) {
    // =============================== MMA issuer =================================
    if (total_slabs > 0 && rank == 0) {   // in a pair: ONE (elected) thread issues for both SMs
      const bool el = elect_one_sync();
      int stage = 0;
      uint32_t ph = 0;
      // M = 256 over the pair: the M field of the instruction descriptor (bits [24,29)) is M >> 4
      const uint32_t ID = idesc_for(WG ? IDESC_MN : IDESC, P.f16) +
                          (PAIR ? ((uint32_t)(TM >> 4) << 24) : 0u);
      auto mma = [&](uint64_t a, uint64_t b, uint32_t acc) {
        if (!el) return;
        if (PAIR) mma2_ss(tmem_base, a, b, ID, acc);
        else mma_ss(tmem_base, a, b, ID, acc);
      };
      for (int i = 0; i < total_slabs; ++i) {
        mbar_wait(full0 + 8 * stage, ph);
        tc_fence_after();
        if (el && dbg && i == 0) dbg[2] = gtime_ns();
        const uint32_t sa = base + stage * STG;
#pragma unroll
        for (int ks = 0; ks < BK / UK; ++ks) {
          uint64_t a_hi, b_hi, a_lo, b_lo;
          if (WG) {   // 16 K-rows = two 1024-byte swizzle atoms per step
            a_hi = smem_desc_sw128_mn(sa + ks * 2048);
            b_hi = smem_desc_sw128_mn(sa + 2 * A_PLANE + ks * 2048);
            a_lo = smem_desc_sw128_mn(sa + A_PLANE + ks * 2048);
            b_lo = smem_desc_sw128_mn(sa + 2 * A_PLANE + BPL + ks * 2048);
          } else {
            a_hi = smem_desc_sw64(sa + ks * UK * 2);
            b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            a_lo = smem_desc_sw64(sa + A_PLANE + ks * UK * 2);
            b_lo = smem_desc_sw64(sa + 2 * A_PLANE + BPL + ks * UK * 2);
          }
          mma(a_hi, b_hi, (i | ks) ? 1u : 0u);
          if (P.x3) {
            if (!(WG && P.a_exact)) mma(a_lo, b_hi, 1u);
            if (!(WG && P.b_exact)) mma(a_hi, b_lo, 1u);
          }
        }
        if (el) {
          if (PAIR) tc_commit2(empty0 + 8 * stage);
          else tc_commit(empty0 + 8 * stage);
        }
        if (++stage == NST) { stage = 0; ph ^= 1; }
      }
      if (el) {
        if (PAIR) tc_commit2(acc_full);
        else tc_commit(acc_full);
        if (dbg) dbg[3] = gtime_ns();
      }
    }
  } else if (total_slabs > 0) {
    // =============================== epilogue (warps 0-7) =======================
    // warp e: TMEM lane quadrant e%4, column group e/4; 16-column chunks dealt round-robin.
    const int quad = warp & 3, grp = warp >> 2;
    constexpr int NG = G_EPI_WARPS / 4;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float inv = P.scale ? P.scale[1] : 1.0f;   // undoes the gradient scale on fp32 outputs

    if (WG) {
      const Job& jb = P.jobs[blockIdx.x];
      const int m = jb.m0 + row;
      mbar_wait(acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int q = grp; q < TN / 16; q += NG) {
        float o[16];
        tmem_ld16(lane_base + 16 * q, o);
        if (m < jb.M) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = jb.n0 + 16 * q + i;
            if (n < jb.N) atomicAdd(jb.out + (long long)m * jb.gm + (long long)n * jb.gk, o[i] * inv);
            else if (jb.col_out != nullptr && n == jb.col_n)
              atomicAdd(jb.col_out + (long long)blockIdx.z * jb.col_stride + m, o[i] * inv);
          }
        }
      }
    } else {
      const int b = blockIdx.z;
      const int t = blockIdx.x * TM + row;
      const bool t_ok = t < P.T;
      if (EPI == EPI_GATE_BWD) {
        // gz -> gh_t = gz*sig*(1-tanh^2), gh_s = gz*tanh*sig*(1-sig)   (modules.py:47-48 differentiated)
        const int CHh = TN;   // 256 gate pairs
        // saved tanh / sigmoid: time-major (B,T,256) fp32, written by the forward gate epilogue
        const bool from_z = P.f0 == nullptr;
        const float* tp = P.f0 + ((int64_t)b * P.T + t) * CHh;
        const float* sp = P.f1 + ((int64_t)b * P.T + t) * CHh;
        const __nv_bfloat16* zh = P.z_hi + ((int64_t)b * P.T + t) * CHh;
        const __nv_bfloat16* zl = P.z_lo + ((int64_t)b * P.T + t) * CHh;
        uint32_t pt[16], ps[16];       // tanh (or the z hi / lo pair words) and sigmoid of 16 channels
        auto fetch = [&](int q) {
          if (t_ok) {
            if (from_z) {
              ld256(zh + 16 * q, pt);
              ld256(zl + 16 * q, pt + 8);
            } else {
              ld256(tp + 16 * q, pt);
              ld256(tp + 16 * q + 8, pt + 8);
            }
            ld256(sp + 16 * q, ps);
            ld256(sp + 16 * q + 8, ps + 8);
          }
        };
        fetch(grp);
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int q = grp; q < TN / 16; q += NG) {
          float gz[16], th[16], sg[16];
          tmem_ld16(lane_base + 16 * q, gz);
#pragma unroll
          for (int i = 0; i < 16; ++i) sg[i] = t_ok ? __uint_as_float(ps[i]) : 0.0f;
          if (from_z) {
            // tanh = z / sigmoid with z = hi + lo (2^-17); sigmoid == 0 (pre-activation < -87)
            // makes both gate derivatives vanish, whatever tanh is
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float z0, z1, e0, e1;
              unpack_pair_f(pt[i], P.f16, z0, z1);
              unpack_pair_f(pt[8 + i], P.f16, e0, e1);
              const float s0 = sg[2 * i], s1 = sg[2 * i + 1];
              th[2 * i] = (t_ok && s0 > 0.0f) ? __fdividef(z0 + e0, s0) : 0.0f;
              th[2 * i + 1] = (t_ok && s1 > 0.0f) ? __fdividef(z1 + e1, s1) : 0.0f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) th[i] = t_ok ? __uint_as_float(pt[i]) : 0.0f;
          }
          if (q + NG < TN / 16) fetch(q + NG);
          if (!t_ok) continue;
          uint32_t th_hi[8], th_lo[8], sg_hi[8], sg_lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float g0 = gz[2 * i], a0 = th[2 * i], s0 = sg[2 * i];
            const float g1 = gz[2 * i + 1], a1 = th[2 * i + 1], s1 = sg[2 * i + 1];
            const float ht0 = g0 * s0 * (1.0f - a0 * a0), ht1 = g1 * s1 * (1.0f - a1 * a1);
            const float hs0 = g0 * a0 * s0 * (1.0f - s0), hs1 = g1 * a1 * s1 * (1.0f - s1);
            if (P.x3) {
              split_pair_f(ht0, ht1, th_hi[i], th_lo[i], P.f16);
              split_pair_f(hs0, hs1, sg_hi[i], sg_lo[i], P.f16);
            } else {
              th_hi[i] = pack_pair_f(ht0, ht1, P.f16);
              sg_hi[i] = pack_pair_f(hs0, hs1, P.f16);
            }
          }
          const int64_t poff = ((int64_t)b * P.T + t) * (2 * CHh) + 16 * q;
          st256(P.p_hi + poff, th_hi);
          st256(P.p_hi + poff + CHh, sg_hi);
          if (P.x3) {
            st256(P.p_lo + poff, th_lo);
            st256(P.p_lo + poff + CHh, sg_lo);
          }
        }
      } else if (EPI == EPI_GX && !alt && P.stage_epi) {
        // ---- staged: gx = acc + g_res through shared-memory tiles and TMA.  Round 2 measured the
        // direct version LSU bound: a thread owns one time row, so each of its 32-byte plane
        // accesses is its own wavefront (8192 per tile, 17 us of a 63 us CTA with the tensor pipe
        // idle).  Here the addend quarter (128 rows x 64 channels, hi / lo) is TMA-loaded into the
        // ring (idle once the accumulator is full), every thread updates ITS row in place, and
        // one thread TMA-stores the quarter to the next block's g_res planes; two buffers. ----
        const int cbase = TN * blockIdx.y, t0 = blockIdx.x * TM;
        const bool has_add = P.a_hi != nullptr, lo2 = P.add_lo != 0;
        const bool leader = threadIdx.x == 0;
        auto tile = [&](int bf, int pl) -> uint32_t { return base + (uint32_t)(2 * bf + pl) * 16384u; };
        auto issue = [&](int bf, int c) {
          const uint32_t bar = afull0 + 8 * bf;
          mbar_expect_tx(bar, (lo2 ? 2 : 1) * 16384);
          tma_load_3d(tile(bf, 0), &maps.m[6], bar, cbase + 64 * c, t0, b);
          if (lo2) tma_load_3d(tile(bf, 1), &maps.m[7], bar, cbase + 64 * c, t0, b);
        };
        mbar_wait(acc_full, 0);      // every MMA of the pair is done: the ring is free
        tc_fence_after();
        if (leader && has_add) { issue(0, 0); issue(1, 1); }
        const uint32_t rsw = (uint32_t)(row & 7), rowb = (uint32_t)row * 128u;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int bf = c & 1;
          if (has_add) {
            if (leader && c >= 1 && c + 1 < 4) {   // buffer (c+1)&1 was the source of quarter c-1's store
              tma_store_wait_read();
              issue((c + 1) & 1, c + 1);
            }
            mbar_wait(afull0 + 8 * bf, (c >> 1) & 1);
          } else if (c >= 2) {
            if (leader) tma_store_wait_read();
            asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          }
#pragma unroll 1
          for (int u = 0; u < 2; ++u) {
            const int cq = 2 * grp + u;            // 16-column chunk of this quarter
            const int q = 4 * c + cq;              // ... and of the accumulator
            float o[16];
            tmem_ld16(lane_base + 16 * q, o);
            const uint32_t a0 = tile(bf, 0) + rowb + (((uint32_t)(2 * cq) ^ rsw) << 4);
            const uint32_t a1 = tile(bf, 0) + rowb + (((uint32_t)(2 * cq + 1) ^ rsw) << 4);
            const uint32_t l0 = a0 + 16384u, l1 = a1 + 16384u;
            if (has_add) {
              const uint4 h0 = lds128(a0), h1 = lds128(a1);
              const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float v0, v1;
                unpack_pair_f(hw[i], P.f16, v0, v1);
                o[2 * i] += v0;
                o[2 * i + 1] += v1;
              }
              if (lo2) {
                const uint4 w0 = lds128(l0), w1 = lds128(l1);
                const uint32_t lw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float v0, v1;
                  unpack_pair_f(lw[i], P.f16, v0, v1);
                  o[2 * i] += v0;
                  o[2 * i + 1] += v1;
                }
              }
            }
            if (P.colsum != nullptr) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = t_ok ? o[i] : 0.0f;
#pragma unroll
              for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
                const bool up = lane & bit;
#pragma unroll
                for (int i = 0; i < w; ++i) {
                  const float send = up ? v[i] : v[i + w];
                  const float keep = up ? v[i + w] : v[i];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
              }
              v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
              if ((lane & 1) == 0) atomicAdd(csum + 16 * q + (lane >> 1), v[0]);
            }
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (lo2) split_pair_f(o[2 * i], o[2 * i + 1], vh[i], vl[i], P.f16);
              else vh[i] = pack_pair_f(o[2 * i], o[2 * i + 1], P.f16);
            }
            sts128(a0, vh[0], vh[1], vh[2], vh[3]);
            sts128(a1, vh[4], vh[5], vh[6], vh[7]);
            if (lo2) {
              sts128(l0, vl[0], vl[1], vl[2], vl[3]);
              sts128(l1, vl[4], vl[5], vl[6], vl[7]);
            }
          }
          fence_async_smem();
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          if (leader) {
            tma_store_3d(&maps.m[8], tile(bf, 0), cbase + 64 * c, t0, b);
            if (lo2) tma_store_3d(&maps.m[9], tile(bf, 1), cbase + 64 * c, t0, b);
            tma_store_commit();
          }
        }
        if (leader) tma_store_wait_all();
        if (P.colsum != nullptr) {
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          for (int i = threadIdx.x; i < TN; i += G_EPI_WARPS * 32)
            atomicAdd(P.colsum + cbase + i, csum[i] * inv);
        }
      } else if (EPI == EPI_GX && !alt) {
        // gx = acc + g_res: the addend and the result are time-major planes (16-byte accesses);
        // the fp32 (B,Cout,T) copy is only written for the gradient handed back to autograd
        const int cbase = TN * blockIdx.y;
        uint32_t hw[8], lw[8];
        auto fetch = [&](int q) {
          if (P.a_hi != nullptr && t_ok) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cout + cbase + 16 * q;
            ld256(P.a_hi + poff, hw);
            if (P.add_lo) ld256(P.a_lo + poff, lw);
          }
        };
        fetch(grp);
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int q = grp; q < TN / 16; q += NG) {
          float o[16];
          tmem_ld16(lane_base + 16 * q, o);
          if (P.a_hi != nullptr && t_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v0, v1;
              unpack_pair_f(hw[i], P.f16, v0, v1);
              o[2 * i] += v0;
              o[2 * i + 1] += v1;
              if (P.add_lo) {
                unpack_pair_f(lw[i], P.f16, v0, v1);
                o[2 * i] += v0;
                o[2 * i + 1] += v1;
              }
            }
          }
          if (q + NG < TN / 16) fetch(q + NG);
          const int ch0 = cbase + 16 * q;
          if (P.colsum != nullptr) {
            // column sums over the warp's 32 rows by recursive halving: every step exchanges
            // half of the remaining columns with the lane whose index differs in one bit
            // (16 shuffles for 16 columns), then one shared-memory atomic per column
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = t_ok ? o[i] : 0.0f;
#pragma unroll
            for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
              const bool up = lane & bit;
#pragma unroll
              for (int i = 0; i < w; ++i) {
                const float send = up ? v[i] : v[i + w];
                const float keep = up ? v[i + w] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
              }
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
            if ((lane & 1) == 0) atomicAdd(csum + 16 * q + (lane >> 1), v[0]);
          }
          if (!t_ok) continue;
          if (P.o0 != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) P.o0[((int64_t)b * P.Cout + ch0 + i) * P.T + t] = o[i] * inv;
          }
          if (P.p_hi != nullptr) {
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (P.add_lo) split_pair_f(o[2 * i], o[2 * i + 1], vh[i], vl[i], P.f16);
              else vh[i] = pack_pair_f(o[2 * i], o[2 * i + 1], P.f16);
            }
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cout + ch0;
            st256(P.p_hi + poff, vh);
            if (P.add_lo) st256(P.p_lo + poff, vl);
          }
        }
        if (P.colsum != nullptr) {
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");   // epilogue warps only
          for (int i = threadIdx.x; i < TN; i += G_EPI_WARPS * 32)
            atomicAdd(P.colsum + cbase + i, csum[i] * inv);
        }
      } else if (EPI == EPI_HEAD_LOSS) {
        // ---- y = acc + bias (one N tile = the whole row), loss and gy = d loss / d y -> planes ----
        float* red = csum;                                  // [2][128] floats of row exchange
        const int Q = P.Qv;
        const float inv_n = (float)(P.loss[1] > 0.0 ? 1.0 / P.loss[1] : 0.0);
        const float gmul = inv_n * (P.plane_s ? *P.plane_s : 1.0f);   // what the gradient planes hold
        mbar_wait(acc_full, 0);
        tc_fence_after();
        double nll = 0.0;
        if (P.tgt_i != nullptr) {
          // softmax cross entropy (train.py:95): two warps share a row (column chunks dealt
          // round-robin), so max and sum are combined through shared memory
          const int k = t_ok ? P.tgt_i[(int64_t)b * P.T + t] : -1;
          const bool valid = k >= 0 && k < Q;               // ignore_label = -1
          float mx = -INFINITY;
#pragma unroll 1
          for (int q = grp; 16 * q < Q; q += NG) {
            float o[16];
            tmem_ld16(lane_base + 16 * q, o);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (16 * q + i < Q) mx = fmaxf(mx, o[i] + __ldg(P.bias + 16 * q + i));
          }
          red[grp * TM + row] = mx;
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          mx = fmaxf(red[row], red[TM + row]);
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          float sum = 0.0f, vt = 0.0f;
#pragma unroll 1
          for (int q = grp; 16 * q < Q; q += NG) {
            float o[16];
            tmem_ld16(lane_base + 16 * q, o);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (16 * q + i < Q) {
                const float v = o[i] + __ldg(P.bias + 16 * q + i);
                sum += expf(v - mx);
                if (16 * q + i == k) vt = v;
              }
            }
          }
          red[grp * TM + row] = sum;
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          const float lse = mx + logf(red[row] + red[TM + row]);
          if (valid && (k >> 4) % NG == grp) nll = (double)(lse - vt);   // the warp that owns column k
#pragma unroll 1
          for (int q = grp; 16 * q < P.Cout; q += NG) {
            float o[16];
            tmem_ld16(lane_base + 16 * q, o);
            if (!t_ok) continue;
            float g[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = 16 * q + i;
              const float v = (c < Q) ? o[i] + __ldg(P.bias + c) : 0.0f;
              o[i] = v;
              g[i] = (c < Q && valid) ? (expf(v - lse) - (c == k ? 1.0f : 0.0f)) * gmul : 0.0f;
            }
            if (P.o0 != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (16 * q + i < Q) P.o0[((int64_t)b * Q + 16 * q + i) * P.T + t] = o[i];
            }
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (P.x3) split_pair_f(g[2 * i], g[2 * i + 1], vh[i], vl[i], P.f16);
              else vh[i] = pack_pair_f(g[2 * i], g[2 * i + 1], P.f16);
            }
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cout + 16 * q;
            st256(P.p_hi + poff, vh);
            if (P.x3) st256(P.p_lo + poff, vl);
          }
        } else {
          // mixture of logistics (modules.py:169-230): 3 * nr <= 32 outputs, one thread per row
          const int nr = Q / 3;
          if (grp == 0) {
            float yv[32], gv[32];
            tmem_ld16(lane_base, yv);
            tmem_ld16(lane_base + 16, yv + 16);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              yv[i] = (i < Q) ? yv[i] + __ldg(P.bias + i) : 0.0f;
              gv[i] = 0.0f;
            }
            if (t_ok) {
              nll = (double)mol_position(yv, 1, P.tgt_f[(int64_t)b * P.T + t], nr, P.mol_half,
                                         P.mol_lsmin, gmul, gv, 1);
              if (P.o0 != nullptr)
                for (int i = 0; i < Q; ++i) P.o0[((int64_t)b * Q + i) * P.T + t] = yv[i];
              for (int q = 0; 16 * q < P.Cout; ++q) {
                uint32_t vh[8], vl[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float g0 = (q < 2) ? gv[16 * q + 2 * i] : 0.0f;
                  const float g1 = (q < 2) ? gv[16 * q + 2 * i + 1] : 0.0f;
                  if (P.x3) split_pair_f(g0, g1, vh[i], vl[i], P.f16);
                  else vh[i] = pack_pair_f(g0, g1, P.f16);
                }
                const int64_t poff = ((int64_t)b * P.T + t) * P.Cout + 16 * q;
                st256(P.p_hi + poff, vh);
                if (P.x3) st256(P.p_lo + poff, vl);
              }
            }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) nll += __shfl_xor_sync(0xffffffffu, nll, off);
        if (lane == 0 && nll != 0.0) atomicAdd(P.loss, nll * (double)inv_n);
      } else if (EPI == EPI_HEAD) {
        const int cbase = TN * blockIdx.y;
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int q = grp; q < TN / 16; q += NG) {
          const int ch0 = cbase + 16 * q;
          if (ch0 >= P.Cout) break;            // Cout is a multiple of 16 for plane outputs
          float o[16];
          tmem_ld16(lane_base + 16 * q, o);
          if (!t_ok) continue;
          uint32_t mk[8];
          if (P.mask_hi != nullptr) {
            ld256(P.mask_hi + ((int64_t)b * P.T + t) * P.Cout + ch0, mk);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = o[i];
            if (P.bias != nullptr && ch0 + i < P.Cout) v += __ldg(P.bias + ch0 + i);
            if (P.relu) v = fmaxf(v, 0.0f);
            if (P.mask_hi != nullptr) {
              float m0, m1;
              unpack_pair_f(mk[i >> 1], P.f16, m0, m1);
              v = (((i & 1) ? m1 : m0) > 0.0f) ? v : 0.0f;
            }
            o[i] = v;
          }
          if (P.o0 != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ch0 + i < P.Cout) P.o0[((int64_t)b * P.Cout + ch0 + i) * P.T + t] = o[i] * inv;
          }
          if (P.p_hi != nullptr) {
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (P.x3) split_pair_f(o[2 * i], o[2 * i + 1], vh[i], vl[i], P.f16);
              else vh[i] = pack_pair_f(o[2 * i], o[2 * i + 1], P.f16);
            }
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cout + ch0;
            st256(P.p_hi + poff, vh);
            if (P.x3) st256(P.p_lo + poff, vl);
          }
        }
      } else {
        // EPI_ACCUM (or an ACCUM tile of a GX launch): gcond[b, ch, t] += acc for
        // ch = 256 * by + col < Cout (fp32, lanes = t)
        const int cbase = TN * by;
        const int Cout = alt ? P.alt_Cout : P.Cout;
        float* op = (alt ? P.alt_out : P.o0) + ((int64_t)b * Cout + cbase) * P.T + t;
        const int nq = (Cout - cbase + 15) / 16 < TN / 16 ? (Cout - cbase + 15) / 16 : TN / 16;
        float pre[16];
        auto fetch = [&](int q) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pre[i] = (t_ok && cbase + 16 * q + i < Cout) ? __ldcs(op + (int64_t)(16 * q + i) * P.T) : 0.0f;
        };
        fetch(grp);
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int q = grp; q < nq; q += NG) {      // only the column chunks that exist
          float o[16], add[16];
          tmem_ld16(lane_base + 16 * q, o);
#pragma unroll
          for (int i = 0; i < 16; ++i) add[i] = pre[i];
          if (q + NG < nq) fetch(q + NG);
          if (!t_ok) continue;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cbase + 16 * q + i < Cout) op[(int64_t)(16 * q + i) * P.T] = fmaf(o[i], inv, add[i]);
        }
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 0) dbg[5] = gtime_ns();
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // neither CTA leaves while the pair's MMAs can still touch it
  if (warp == GW_MMA) {
    __syncwarp();
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256)
                   : "memory");
  }
}

static size_t gemm_smem() {   // 2 x 48 KB (single CTAs) = 3 x 32 KB (pairs)
  return 1024 + (size_t)G_STAGES * STAGE_BYTES + 8 * (2 * 3 + 2) + sizeof(float) * TN + 16 + 16;
}
// VQW_TC_GEMM_PAIR=0 keeps every backward / head GEMM on single CTAs
static bool gemm_pair_enabled() {
  const char* e = getenv("VQW_TC_GEMM_PAIR");
  return !(e && e[0] == '0');
}

// weight rows per TMA box of a time-flavour GEMM = the rows one CTA stages per slab
static uint32_t WROWS() { return gemm_pair_enabled() ? tc::TN / 2 : tc::TN; }

// `pair_ok`: consecutive blockIdx.x values form valid pairs (same B tile, adjacent A row blocks);
// the grid's x extent is rounded up to even -- the extra CTA of a time-flavour launch is all
// padding (its loads are out of bounds = zero, its stores are masked)
template <int EPI>
static int launch_gemm(const Maps& maps, const GemmParams& P_in, dim3 grid, cudaStream_t stream,
                       bool pair_ok = true) {
  const size_t smem = gemm_smem();
  GemmParams P = P_in;
  static long long* dbg_buf = nullptr;
  const char* tl = getenv("VQW_GEMM_TIMELINE");
  const bool timeline = tl && atoi(tl) == EPI && tl[0] >= '0' && tl[0] <= '9';
  if (timeline) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * 8 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 64 * 8 * sizeof(long long), stream);
    P.dbg = dbg_buf;
  }
  struct Dump {
    bool on; cudaStream_t st; long long* buf; int epi;
    ~Dump() {
      if (!on) return;
      long long h[64 * 8];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
      long long t0 = 0;
      for (int c = 0; c < 64; ++c) if (h[8 * c] && (!t0 || h[8 * c] < t0)) t0 = h[8 * c];
      fprintf(stderr, "[vqw gemm timeline] epilogue %d: CTA(x,z) sm | entry setup first-slab mma-issued epi-done"
                      " (ns rel. to the earliest entry)\n", epi);
      for (int c = 0; c < 64; ++c)
        if (h[8 * c])
          fprintf(stderr, "  (%2d,%d) sm %3lld | %7lld %7lld %7lld %7lld %7lld\n", c % 32, c / 32, h[8 * c + 6],
                  h[8 * c] - t0, h[8 * c + 1] - t0, h[8 * c + 2] ? h[8 * c + 2] - t0 : -1,
                  h[8 * c + 3] ? h[8 * c + 3] - t0 : -1, h[8 * c + 5] - t0);
    }
  } dump{timeline, stream, dbg_buf, EPI};
  if (pair_ok && gemm_pair_enabled() && (EPI != EPI_WGRAD || grid.x % 2 == 0)) {
    auto kern = tc_gemm_kernel<EPI, 1>;
    VQW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((grid.x + 1) / 2 * 2, grid.y, grid.z);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VQW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, P));
    VQW_CHECK_LAUNCH("tc_gemm_kernel(pair)");
    return 0;
  }
  auto kern = tc_gemm_kernel<EPI, 0>;
  VQW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, G_THREADS, smem, stream>>>(maps, P);
  static const char* names[6] = {"tc_gemm_kernel<GATE_BWD>", "tc_gemm_kernel<GX>",
                                 "tc_gemm_kernel<ACCUM>", "tc_gemm_kernel<WGRAD>",
                                 "tc_gemm_kernel<HEAD>", "tc_gemm_kernel<HEAD_LOSS>"};
  VQW_CHECK_LAUNCH(names[EPI]);
  return 0;
}

// ------------------------------------------------------------------ persistent time-flavour kernel
// GX (+ the ACCUM tiles of the same block) and GATE_BWD as ONE persistent launch: a cluster of two
// CTAs per SM pair walks the tile list, the accumulator is double buffered in TMEM (2 x 256
// columns) and every epilogue goes through shared-memory tiles + TMA, so the epilogue of tile i
// runs under the MMAs of tile i + 1.  Why (round 2 timelines, profiles/r2_gemm_timeline_*.log):
// with one tile per CTA and two CTAs per SM both CTAs of an SM run in lockstep -- they share the
// tensor pipe during their MMA phases and then both sit in their epilogues with the pipe idle
// (GX: 43 us of MMA + 12-23 us of epilogue; GATE_BWD: 9 us + 43 us), and the direct epilogues were
// LSU bound (one 32-byte access per thread and row = one wavefront each).
constexpr int P_NST = 4;                         // ring stages of 32 KB (A 16 KB + this CTA's half of B)
constexpr int P_STG = 2 * A_PLANE + B_PLANE;     // 32 KB
constexpr int P_STAGING = 64 * 1024;             // epilogue tiles (two buffers of 32 KB)
// weight-gradient tiles do not use the staging tiles: their ring spans ring + staging (192 KB) in
// stages of 8 KB per operand plane actually loaded -- 6 stages of 32 KB with all four planes, 12
// of 16 KB with the hi planes only (a 4-stage ring of 16 KB loads was bound by the TMA latency:
// ~700 cycles per 256-cycle slab)
constexpr int P_MAXST = 12;
constexpr int P_WG_RING = P_NST * P_STG + P_STAGING;
constexpr int P_NBAR = 2 * P_MAXST + 6;          // ring + acc_full[2] + acc_empty[2] + afull[2]

static size_t persist_smem() {
  return 1024 + (size_t)P_NST * P_STG + P_STAGING + 8 * P_NBAR + 16 + sizeof(float) * TN + 16;
}

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
tc_time_persistent_kernel(const __grid_constant__ Maps maps, const __grid_constant__ GemmParams P,
                          int n_x, int n_y, int n_z) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stg0 = base + P_NST * P_STG;                       // staging tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_NST * P_STG + P_STAGING);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + P_MAXST);
  const uint32_t acc_full0 = smem_u32(bars + 2 * P_MAXST), acc_empty0 = acc_full0 + 16, afull0 = acc_full0 + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + P_NBAR);
  float* csum = reinterpret_cast<float*>(bars + P_NBAR + 2);       // [TN] column sums (GX)
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nplanes = P.x3 ? 2 : 1;
  constexpr bool WG = (EPI == EPI_WGRAD);   // weight gradients: tile = (256-row job pair, batch item), K = time
  // ring geometry: stage size / count and the plane offsets inside a stage
  const bool wg_alo = P.x3 && !P.a_exact, wg_blo = P.x3 && !P.b_exact;
  // a weight-gradient stage holds wg_sub slabs of 32 time steps (2 with the hi planes only: the
  // ~300 cycles of barrier wait + fence + commit per stage are then spread over 4 MMAs instead of 2)
  const uint32_t stg1 = (uint32_t)(((wg_alo ? 2 : 1) + (wg_blo ? 2 : 1)) * A_PLANE);
  const int wg_sub = (WG && P.wg_sub > 1) ? P.wg_sub : 1;
  const uint32_t stg = WG ? stg1 * (uint32_t)wg_sub : (uint32_t)P_STG;
  const int nst = WG ? min(P_MAXST, (int)(P_WG_RING / stg)) : P_NST;
  const uint32_t wg_boff = (uint32_t)((wg_alo ? 2 : 1) * A_PLANE);
  const int n_tiles = n_x * n_y * n_z;                             // pair tiles
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == GW_TMA && lane == 0) {
    for (int s = 0; s < P_MAXST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full0 + 8 * s, 1);
      mbar_init(acc_empty0 + 8 * s, 2 * G_EPI_WARPS);   // one elected lane per epilogue warp, both CTAs
      mbar_init(afull0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GW_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile id -> (time-tile pair, N tile, batch item); x runs fastest, so the clusters working at
  // the same time share their weight tile and neighbouring activation rows in L2
  auto decode = [&](int tile, int& bx, int& y, int& z) {
    const int xp = tile % n_x;
    y = (tile / n_x) % n_y;
    z = tile / (n_x * n_y);
    bx = 2 * xp + (int)rank;
  };

  if (warp == GW_TMA) {
    // =============================== TMA producer ===============================
    // the WHOLE warp walks the loop (uniform control flow: counters, addresses and coordinates live
    // in uniform registers) and one elected lane issues; inside `if (lane == 0) { loop }` every
    // operand of a UTMALDG / UTCHMMA came from a per-thread register through an ELECT / R2UR
    // waterfall: ~350 cycles of issue per K = 16 step, measured (profiles/r2_summary.md)
    {
      const bool el = elect_one_sync();
      int stage = 0;
      uint32_t ph = 0;
      bool free_seen = false;   // (weight-gradient path) the empty barrier of `stage` has been seen complete
      for (int tile = cid; tile < n_tiles; tile += ncl) {
        int bx, y, bb;
        decode(tile, bx, y, bb);
        if (WG) {
          // a stage = 2 (A) + 2 (this CTA's half of B) boxes of {64 channels, 32 time steps} per plane
          const Job& jb = P.jobs[bx];
          const CUtensorMap* ma = &maps.m[2 * jb.a_map];
          const CUtensorMap* mb = &maps.m[2 * jb.b_map];
          const bool blo = wg_blo, alo = wg_alo;
          const int nb0 = jb.n0 + (int)rank * (TN / 2);
          for (int i = 0; i < P.slabs_per_item; ++i) {
            if (P.dbg_mode >= 3) continue;
            if (!free_seen) mbar_wait(empty0 + 8 * stage, ph ^ 1);
            // probe the next stage's empty barrier; the answer arrives while this stage is issued
            const int nstage = (stage + 1 == nst) ? 0 : stage + 1;
            const uint32_t nph = (stage + 1 == nst) ? ph ^ 1 : ph;
            const bool nfree = mbar_try(empty0 + 8 * nstage, nph ^ 1);
            const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
            const uint32_t sa0 = base + stage * stg;
            if (rank == 0 && el)
              mbar_expect_tx(full0 + 8 * stage, P.dbg_mode == 1 ? 0 :
                             2 * wg_sub * ((alo ? 2 : 1) * A_PLANE + (blo ? 2 : 1) * (B_PLANE / 2)));
            if (P.dbg_mode == 1) {
              stage = nstage; ph = nph; free_seen = nfree;
              continue;
            }
            for (int u = 0; u < wg_sub; ++u) {
              const uint32_t sa = sa0 + u * stg1;
              // y = time chunk; steps past T are zero-filled by TMA
              const int ta = ((y * P.slabs_per_item + i) * wg_sub + u) * BK, tb = ta + jb.shift;
              if ((P.wide_maps >> jb.a_map) & 1u) {
                if (el) tma2_load_4d(sa, ma, fb, 0, ta, jb.m0 >> 6, bb);
                if (alo && el) tma2_load_4d(sa + A_PLANE, ma + 1, fb, 0, ta, jb.m0 >> 6, bb);
              } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  if (el) tma2_load_3d(sa + h * 4096, ma, fb, jb.m0 + 64 * h, ta, bb);
                  if (alo && el) tma2_load_3d(sa + A_PLANE + h * 4096, ma + 1, fb, jb.m0 + 64 * h, ta, bb);
                }
              }
              if ((P.wide_maps >> jb.b_map) & 1u) {
                if (el) tma2_load_4d(sa + wg_boff, mb, fb, 0, tb, nb0 >> 6, bb);
                if (blo && el) tma2_load_4d(sa + wg_boff + B_PLANE / 2, mb + 1, fb, 0, tb, nb0 >> 6, bb);
              } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  if (el) tma2_load_3d(sa + wg_boff + h * 4096, mb, fb, nb0 + 64 * h, tb, bb);
                  if (blo && el)
                    tma2_load_3d(sa + wg_boff + B_PLANE / 2 + h * 4096, mb + 1, fb, nb0 + 64 * h, tb, bb);
                }
              }
            }
            stage = nstage; ph = nph; free_seen = nfree;
          }
          continue;
        }
        const bool alt = (EPI == EPI_GX) && P.alt_y > 0 && y >= P.alt_y;
        const int by = alt ? y - P.alt_y : y;
        const int t0 = bx * TM;
        const int nseg = alt ? 1 : P.nseg;
        for (int s = 0; s < nseg; ++s) {
          const Seg& sg = alt ? P.alt_seg : P.seg[s];
          const CUtensorMap* ma = &maps.m[2 * sg.a_map];
          const CUtensorMap* mb = &maps.m[2 * sg.b_map];
          for (int i = 0; i < sg.nslabs; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
            const uint32_t sa = base + stage * P_STG;
            if (rank == 0 && el) mbar_expect_tx(full0 + 8 * stage, 2 * nplanes * (A_PLANE + B_PLANE / 2));
            const int brow = sg.b_row0 + TN * by + (int)rank * (TN / 2);
            if (el) tma2_load_3d(sa, ma, fb, sg.a_c0 + i * BK, t0 + sg.a_shift, bb);
            if (el) tma2_load_3d(sa + 2 * A_PLANE, mb, fb, sg.b_c0 + i * BK, brow, 0);
            if (P.x3) {
              if (el) tma2_load_3d(sa + A_PLANE, ma + 1, fb, sg.a_c0 + i * BK, t0 + sg.a_shift, bb);
              if (el) tma2_load_3d(sa + 2 * A_PLANE + B_PLANE / 2, mb + 1, fb, sg.b_c0 + i * BK, brow, 0);
            }
            if (++stage == nst) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == GW_MMA) {
    // =============================== MMA issuer (leader CTA) ====================
    if (rank == 0) {
      const bool el = elect_one_sync();
      int stage = 0;
      uint32_t ph = 0;
      const uint32_t ID = idesc_for(WG ? IDESC_MN : IDESC, P.f16) + ((uint32_t)(TM >> 4) << 24);   // M = 256
      bool ready = false;      // the full barrier of `stage` has already been seen complete
      int it = 0;
      for (int tile = cid; tile < n_tiles; tile += ncl, ++it) {
        int bx, y, bb;
        decode(tile, bx, y, bb);
        const bool alt = (EPI == EPI_GX) && P.alt_y > 0 && y >= P.alt_y;
        int total_slabs = 0;
        if (WG) total_slabs = P.slabs_per_item;
        else if (alt) total_slabs = P.alt_seg.nslabs;
        else for (int s = 0; s < P.nseg; ++s) total_slabs += P.seg[s].nslabs;
        const int buf = it & 1;
        if (it >= 2) {                       // the epilogue of tile it-2 has drained this accumulator
          mbar_wait(acc_empty0 + 8 * buf, ((it >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + 256 * buf;
        if (el && P.dbg && cid == 0 && it < 8) P.dbg[4 * it] = gtime_ns();
        for (int i = 0; i < total_slabs; ++i) {
          if (!(WG && P.dbg_mode >= 3) && !ready) mbar_wait(full0 + 8 * stage, ph);
          if (!(WG && P.dbg_mode >= 3)) tc_fence_after();
          // probe the NEXT stage now: the answer arrives while this stage's MMAs are being issued
          const int nstage = (stage + 1 == nst) ? 0 : stage + 1;
          const uint32_t nph = (stage + 1 == nst) ? ph ^ 1 : ph;
          const bool nready = mbar_try(full0 + 8 * nstage, nph);
          const uint32_t sa0 = base + stage * stg;
          const int nks = (BK / UK) * wg_sub;
#pragma unroll 4
          for (int kq = 0; kq < nks; ++kq) {
            const int ks = kq & (BK / UK - 1);
            const uint32_t sa = sa0 + (uint32_t)(kq / (BK / UK)) * stg1;   // sub-slab of this K step
            if (WG && P.dbg_mode == 2) continue;
            if (WG) {   // MN-major tiles: 16 K-rows = two 1024-byte swizzle atoms per step
              const uint32_t a_hi = desc_lo_sw128_mn(sa + ks * 2048);
              const uint32_t b_hi = desc_lo_sw128_mn(sa + wg_boff + ks * 2048);
              if (el) mma2_ss_w<DESC_HI_SW128_MN>(acc, a_hi, b_hi, ID, (i | kq) ? 1u : 0u);
              if (wg_alo && el)
                mma2_ss_w<DESC_HI_SW128_MN>(acc, desc_lo_sw128_mn(sa + A_PLANE + ks * 2048), b_hi, ID, 1u);
              if (wg_blo && el)
                mma2_ss_w<DESC_HI_SW128_MN>(acc, a_hi, desc_lo_sw128_mn(sa + wg_boff + B_PLANE / 2 + ks * 2048),
                                            ID, 1u);
            } else {
              const uint32_t a_hi = desc_lo_sw64(sa + ks * UK * 2);
              const uint32_t b_hi = desc_lo_sw64(sa + 2 * A_PLANE + ks * UK * 2);
              if (el) mma2_ss_w<DESC_HI_SW64>(acc, a_hi, b_hi, ID, (i | kq) ? 1u : 0u);
              if (P.x3 && el) {
                mma2_ss_w<DESC_HI_SW64>(acc, desc_lo_sw64(sa + A_PLANE + ks * UK * 2), b_hi, ID, 1u);
                mma2_ss_w<DESC_HI_SW64>(acc, a_hi, desc_lo_sw64(sa + 2 * A_PLANE + B_PLANE / 2 + ks * UK * 2),
                                        ID, 1u);
              }
            }
          }
          if (!(WG && P.dbg_mode >= 3) && el) tc_commit2(empty0 + 8 * stage);
          stage = nstage; ph = nph; ready = nready;
        }
        if (el) tc_commit2(acc_full0 + 8 * buf);
        if (el && P.dbg && cid == 0 && it < 8) P.dbg[4 * it + 1] = gtime_ns();
      }
    }
  } else {
    // =============================== epilogue (warps 0-7) =======================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;
    const float inv = P.scale ? P.scale[1] : 1.0f;
    const bool leader = threadIdx.x == 0;
    const bool rec = P.dbg != nullptr && cid == 0 && rank == 0 && threadIdx.x == 0;
    const uint32_t acc_empty_l = mapa_u32(acc_empty0, 0);
    uint32_t nld[2] = {0, 0};                 // loads issued so far into each staging buffer
    int it = 0;
    for (int tile = cid; tile < n_tiles; tile += ncl, ++it) {
      int bx, y, b;
      decode(tile, bx, y, b);
      const bool alt = (EPI == EPI_GX) && P.alt_y > 0 && y >= P.alt_y;
      const int by = alt ? y - P.alt_y : y;
      const int buf = it & 1;
      const int t0 = bx * TM, t = t0 + row;
      const bool t_ok = t < P.T;
      const uint32_t lane_base = tmem_base + 256 * buf + ((uint32_t)(quad * 32) << 16);
      mbar_wait(acc_full0 + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      if (rec && it < 8) P.dbg[4 * it + 2] = gtime_ns();
      if (WG && P.wg_partial != nullptr) {
        // ---- weight-gradient tile -> its own [256 x 256] fp32 slot (plain 64-byte stores per
        // thread; 18-36 M scattered fp32 atomics per launch were what bounded this kernel) ----
        constexpr int NG = G_EPI_WARPS / 4;
        float* pt = P.wg_partial + ((size_t)(b * n_y + y) * n_x + (bx >> 1)) * 65536 +
                    (size_t)((int)rank * TM + row) * TN;
#pragma unroll 1
        for (int q = grp; q < TN / 16; q += NG) {
          uint32_t o[16];
          tmem_ld16(lane_base + 16 * q, reinterpret_cast<float*>(o));
          st256(pt + 16 * q, o);
          st256(pt + 16 * q + 8, o + 8);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_l + 8 * buf);
      } else if (WG) {
        // ---- weight-gradient tile: fp32 atomics into gW (K is split over the batch items) ----
        constexpr int NG = G_EPI_WARPS / 4;
        const Job& jb = P.jobs[bx];
        const int m = jb.m0 + row;
#pragma unroll 1
        for (int q = grp; q < TN / 16; q += NG) {
          float o[16];
          tmem_ld16(lane_base + 16 * q, o);
          if (m < jb.M) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = jb.n0 + 16 * q + i;
              if (n < jb.N) atomicAdd(jb.out + (long long)m * jb.gm + (long long)n * jb.gk, o[i] * inv);
              else if (jb.col_out != nullptr && n == jb.col_n)
                atomicAdd(jb.col_out + (long long)b * jb.col_stride + m, o[i] * inv);   // over time chunks
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_l + 8 * buf);
      } else if (EPI == EPI_GX && !alt) {
        // ---- gx = acc + g_res, quarter by quarter through [128 rows x 64 ch] hi / lo tiles ----
        const int cbase = TN * y;
        const bool has_add = P.a_hi != nullptr, lo2 = P.add_lo != 0;
        auto tl = [&](int bf, int pl) -> uint32_t { return stg0 + (uint32_t)(2 * bf + pl) * 16384u; };
        auto issue = [&](int bf, int c) {
          const uint32_t bar = afull0 + 8 * bf;
          mbar_expect_tx(bar, (lo2 ? 2 : 1) * 16384);
          tma_load_3d(tl(bf, 0), &maps.m[6], bar, cbase + 64 * c, t0, b);
          if (lo2) tma_load_3d(tl(bf, 1), &maps.m[7], bar, cbase + 64 * c, t0, b);
        };
        if (P.colsum != nullptr) {
          for (int i = threadIdx.x; i < TN; i += G_EPI_WARPS * 32) csum[i] = 0.0f;
        }
        if (leader) {
          tma_store_wait_read();             // the previous tile's stores have left the buffers
          if (has_add) { issue(0, 0); issue(1, 1); }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
        const uint32_t rsw = (uint32_t)(row & 7), rowb = (uint32_t)row * 128u;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int bf = c & 1;
          if (has_add) {
            if (leader && c >= 1 && c + 1 < 4) {
              tma_store_wait_read();
              issue((c + 1) & 1, c + 1);
            }
            mbar_wait(afull0 + 8 * bf, nld[bf] & 1);
            ++nld[bf];
          } else if (c >= 2) {
            if (leader) tma_store_wait_read();
            asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          }
#pragma unroll 1
          for (int u = 0; u < 2; ++u) {
            const int cq = 2 * grp + u, q = 4 * c + cq;
            float o[16];
            tmem_ld16(lane_base + 16 * q, o);
            const uint32_t a0 = tl(bf, 0) + rowb + (((uint32_t)(2 * cq) ^ rsw) << 4);
            const uint32_t a1 = tl(bf, 0) + rowb + (((uint32_t)(2 * cq + 1) ^ rsw) << 4);
            const uint32_t l0 = a0 + 16384u, l1 = a1 + 16384u;
            if (has_add) {
              const uint4 h0 = lds128(a0), h1 = lds128(a1);
              const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float v0, v1;
                unpack_pair_f(hw[i], P.f16, v0, v1);
                o[2 * i] += v0;
                o[2 * i + 1] += v1;
              }
              if (lo2) {
                const uint4 w0 = lds128(l0), w1 = lds128(l1);
                const uint32_t lw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float v0, v1;
                  unpack_pair_f(lw[i], P.f16, v0, v1);
                  o[2 * i] += v0;
                  o[2 * i + 1] += v1;
                }
              }
            }
            if (P.colsum != nullptr) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = t_ok ? o[i] : 0.0f;
#pragma unroll
              for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
                const bool up = lane & bit;
#pragma unroll
                for (int i = 0; i < w; ++i) {
                  const float send = up ? v[i] : v[i + w];
                  const float keep = up ? v[i + w] : v[i];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
              }
              v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
              if ((lane & 1) == 0) atomicAdd(csum + 16 * q + (lane >> 1), v[0]);
            }
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (lo2) split_pair_f(o[2 * i], o[2 * i + 1], vh[i], vl[i], P.f16);
              else vh[i] = pack_pair_f(o[2 * i], o[2 * i + 1], P.f16);
            }
            sts128(a0, vh[0], vh[1], vh[2], vh[3]);
            sts128(a1, vh[4], vh[5], vh[6], vh[7]);
            if (lo2) {
              sts128(l0, vl[0], vl[1], vl[2], vl[3]);
              sts128(l1, vl[4], vl[5], vl[6], vl[7]);
            }
          }
          if (c == 3) {                      // every TMEM read of this tile is done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_l + 8 * buf);
          }
          fence_async_smem();
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          if (leader) {
            tma_store_3d(&maps.m[8], tl(bf, 0), cbase + 64 * c, t0, b);
            if (lo2) tma_store_3d(&maps.m[9], tl(bf, 1), cbase + 64 * c, t0, b);
            tma_store_commit();
          }
        }
        if (P.colsum != nullptr) {           // csum is complete after the last quarter's barrier
          for (int i = threadIdx.x; i < TN; i += G_EPI_WARPS * 32)
            atomicAdd(P.colsum + cbase + i, csum[i] * inv);
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");   // before the next tile zeroes it
        }
      } else if (EPI == EPI_GX) {
        // ---- ACCUM tile: gcond[b, ch, t] += acc (fp32 (B,Cout,T): lanes = consecutive t) ----
        constexpr int NG = G_EPI_WARPS / 4;
        const int cbase = TN * by, Cout = P.alt_Cout;
        float* op = P.alt_out + ((int64_t)b * Cout + cbase) * P.T + t;
        const int nq = (Cout - cbase + 15) / 16 < TN / 16 ? (Cout - cbase + 15) / 16 : TN / 16;
#pragma unroll 1
        for (int q = grp; q < nq; q += NG) {
          float o[16], pre[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)        // 16 independent loads in flight, then the accumulator
            pre[i] = (t_ok && cbase + 16 * q + i < Cout) ? __ldcs(op + (int64_t)(16 * q + i) * P.T) : 0.0f;
          tmem_ld16(lane_base + 16 * q, o);
          if (!t_ok) continue;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cbase + 16 * q + i < Cout) op[(int64_t)(16 * q + i) * P.T] = fmaf(o[i], inv, pre[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_l + 8 * buf);
      } else {
        // ---- GATE_BWD: gz -> gh_t = gz sig (1 - tanh^2), gh_s = gz tanh sig (1 - sig) with
        // tanh = z / sig (bf16x3: the forward saves sig and the z planes), eighth by eighth:
        // 32 pairs = [128 rows x 32 ch] tiles (64-byte rows, 64-byte swizzle); z hi / lo arrive by
        // TMA and are overwritten in place by gh_t hi / lo, gh_s hi / lo go to two more tiles, all
        // four leave by TMA; sig is read directly (fp32, 64 contiguous bytes per thread) ----
        constexpr int CHh = TN;
        auto tl = [&](int bf, int k) -> uint32_t { return stg0 + (uint32_t)bf * 32768u + (uint32_t)k * 8192u; };
        auto issue = [&](int bf, int e) {
          const uint32_t bar = afull0 + 8 * bf;
          mbar_expect_tx(bar, 2 * 8192);
          tma_load_3d(tl(bf, 0), &maps.m[6], bar, 32 * e, t0, b);
          tma_load_3d(tl(bf, 1), &maps.m[7], bar, 32 * e, t0, b);
        };
        const float* sp = P.f1 + ((int64_t)b * P.T + t) * CHh + 16 * grp;
        uint32_t ps[16];
        auto fetch_sig = [&](int e) {
          if (t_ok) {
            ld256(sp + 32 * e, ps);
            ld256(sp + 32 * e + 8, ps + 8);
          }
        };
        if (leader) {
          tma_store_wait_read();
          issue(0, 0);
          issue(1, 1);
        }
        fetch_sig(0);
        asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
        // this thread's 16 pairs of an eighth: 32 bytes = 16-byte chunks 2 grp, 2 grp + 1 of its 64-byte row
        const uint32_t sw = (uint32_t)((row >> 1) & 3), rowb = (uint32_t)row * 64u;
        const uint32_t o0 = rowb + (((uint32_t)(2 * grp) ^ sw) << 4), o1 = rowb + (((uint32_t)(2 * grp + 1) ^ sw) << 4);
#pragma unroll 1
        for (int e = 0; e < 8; ++e) {
          const int bf = e & 1;
          if (leader && e >= 1 && e + 1 < 8) {
            tma_store_wait_read();
            issue((e + 1) & 1, e + 1);
          }
          mbar_wait(afull0 + 8 * bf, nld[bf] & 1);
          ++nld[bf];
          float gz[16], sg[16];
          tmem_ld16(lane_base + 32 * e + 16 * grp, gz);
#pragma unroll
          for (int i = 0; i < 16; ++i) sg[i] = t_ok ? __uint_as_float(ps[i]) : 0.0f;
          if (e + 1 < 8) fetch_sig(e + 1);
          const uint4 zh0 = lds128(tl(bf, 0) + o0), zh1 = lds128(tl(bf, 0) + o1);
          const uint4 zl0 = lds128(tl(bf, 1) + o0), zl1 = lds128(tl(bf, 1) + o1);
          const uint32_t zh[8] = {zh0.x, zh0.y, zh0.z, zh0.w, zh1.x, zh1.y, zh1.z, zh1.w};
          const uint32_t zl[8] = {zl0.x, zl0.y, zl0.z, zl0.w, zl1.x, zl1.y, zl1.z, zl1.w};
          uint32_t th_hi[8], th_lo[8], sg_hi[8], sg_lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float z0, z1, e0, e1;
            unpack_pair_f(zh[i], P.f16, z0, z1);
            unpack_pair_f(zl[i], P.f16, e0, e1);
            const float s0 = sg[2 * i], s1 = sg[2 * i + 1];
            const float a0 = (t_ok && s0 > 0.0f) ? __fdividef(z0 + e0, s0) : 0.0f;
            const float a1 = (t_ok && s1 > 0.0f) ? __fdividef(z1 + e1, s1) : 0.0f;
            const float g0 = gz[2 * i], g1 = gz[2 * i + 1];
            const float ht0 = g0 * s0 * (1.0f - a0 * a0), ht1 = g1 * s1 * (1.0f - a1 * a1);
            const float hs0 = g0 * a0 * s0 * (1.0f - s0), hs1 = g1 * a1 * s1 * (1.0f - s1);
            split_pair_f(ht0, ht1, th_hi[i], th_lo[i], P.f16);
            split_pair_f(hs0, hs1, sg_hi[i], sg_lo[i], P.f16);
          }
          sts128(tl(bf, 0) + o0, th_hi[0], th_hi[1], th_hi[2], th_hi[3]);
          sts128(tl(bf, 0) + o1, th_hi[4], th_hi[5], th_hi[6], th_hi[7]);
          sts128(tl(bf, 1) + o0, th_lo[0], th_lo[1], th_lo[2], th_lo[3]);
          sts128(tl(bf, 1) + o1, th_lo[4], th_lo[5], th_lo[6], th_lo[7]);
          sts128(tl(bf, 2) + o0, sg_hi[0], sg_hi[1], sg_hi[2], sg_hi[3]);
          sts128(tl(bf, 2) + o1, sg_hi[4], sg_hi[5], sg_hi[6], sg_hi[7]);
          sts128(tl(bf, 3) + o0, sg_lo[0], sg_lo[1], sg_lo[2], sg_lo[3]);
          sts128(tl(bf, 3) + o1, sg_lo[4], sg_lo[5], sg_lo[6], sg_lo[7]);
          if (e == 7) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_l + 8 * buf);
          }
          fence_async_smem();
          asm volatile("bar.sync 1, %0;" ::"n"(G_EPI_WARPS * 32) : "memory");
          if (leader) {
            tma_store_3d(&maps.m[8], tl(bf, 0), 32 * e, t0, b);            // gh_t hi
            tma_store_3d(&maps.m[9], tl(bf, 1), 32 * e, t0, b);            // gh_t lo
            tma_store_3d(&maps.m[8], tl(bf, 2), CHh + 32 * e, t0, b);      // gh_s hi
            tma_store_3d(&maps.m[9], tl(bf, 3), CHh + 32 * e, t0, b);      // gh_s lo
            tma_store_commit();
          }
        }
      }
      if (rec && it < 8) P.dbg[4 * it + 3] = gtime_ns();
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == GW_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// time-flavour launch on the persistent kernel: grid (x, y, z) as for launch_gemm
// Fixed-order sum of the weight-gradient tiles the persistent kernel stored (wg_partial): pair
// tile `blockIdx.x`, 8 of its 256 rows per CTA, thread = column.  gW += inv * sum over (item, time
// chunk); the constant-one column of the condition job goes to col_out per item.  Every output
// element has exactly one owner, so the result is deterministic (no atomics).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const __grid_constant__ GemmParams P, const float* __restrict__ partial,
                    int n_pairs, int n_chunks, int B) {
  // CTA = 4 rows of pair tile blockIdx.x, thread = 4 consecutive columns of one row
  const float inv = P.scale ? P.scale[1] : 1.0f;
  const int pair = blockIdx.x;
  const int r = blockIdx.y * 4 + (threadIdx.x >> 6), n4 = (threadIdx.x & 63) * 4;
  const size_t ustride = (size_t)n_pairs * 65536;
  const Job& jb = P.jobs[2 * pair + (r >> 7)];
  const int m = jb.m0 + (r & 127);
  if (m >= jb.M) return;
  const int na = jb.n0 + n4;
  const bool has_col = jb.col_out != nullptr && jb.col_n >= na && jb.col_n < na + 4;
  if (na >= jb.N && !has_col) return;
  const float* p = partial + (size_t)pair * 65536 + (size_t)r * TN + n4;
  const int units = B * n_chunks;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int u0 = 0; u0 < units; u0 += 8) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = (u0 + k < units) ? __ldcs(reinterpret_cast<const float4*>(p + (size_t)(u0 + k) * ustride))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
  }
  const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (na + i < jb.N) jb.out[(long long)m * jb.gm + (long long)(na + i) * jb.gk] += sv[i] * inv;
  if (has_col) {   // the constant-one column: per item, summed over its time chunks
    const float* pc = p + (jb.col_n - na);
    for (int b = 0; b < B; ++b) {
      float c = 0.0f;
      for (int k = 0; k < n_chunks; ++k) c += pc[(size_t)(b * n_chunks + k) * ustride];
      jb.col_out[(long long)b * jb.col_stride + m] += c * inv;
    }
  }
}

template <int EPI>
static int launch_persistent(const Maps& maps, const GemmParams& P, dim3 grid, cudaStream_t stream) {
  auto kern = tc_time_persistent_kernel<EPI>;
  const size_t smem = persist_smem();
  VQW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  VQW_CHECK_CUDA(cudaGetDevice(&dev));
  VQW_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_x = (int)(grid.x + 1) / 2, n_y = (int)grid.y, n_z = (int)grid.z;
  const long long n_tiles = (long long)n_x * n_y * n_z;
  int ncl = sms / 2;
  if (n_tiles < ncl) ncl = (int)n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ncl);
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GemmParams Pd = P;
  static long long* dbg_buf = nullptr;
  const char* tl = getenv("VQW_GEMM_TIMELINE");
  const bool timeline = tl && atoi(tl) == 10 + EPI;
  if (timeline) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 64 * sizeof(long long), stream);
    Pd.dbg = dbg_buf;
  }
  VQW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, Pd, n_x, n_y, n_z));
  VQW_CHECK_LAUNCH(EPI == EPI_GX ? "tc_time_persistent_kernel<GX>"
                                 : (EPI == EPI_WGRAD ? "tc_time_persistent_kernel<WGRAD>"
                                                     : "tc_time_persistent_kernel<GATE_BWD>"));
  if (timeline) {
    long long h[64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[vqw persistent timeline] epilogue %d, cluster 0: tile | mma [start, issued] epilogue [start, done] (ns)\n", EPI);
    for (int i = 0; i < 8; ++i)
      if (h[4 * i])
        fprintf(stderr, "  %d | %7lld %7lld | %7lld %7lld\n", i, h[4 * i] - h[0], h[4 * i + 1] - h[0],
                h[4 * i + 2] - h[0], h[4 * i + 3] - h[0]);
  }
  return 0;
}

// ------------------------------------------------------------------ operand preparation ----
// out[r][k] (rows x K, K contiguous) = src(r, k) for the three transposed weight operands
//   kind 0: W2T [Ch rows][Cr + Cs]   = [Wr ; Ws]^T
//   kind 1: WcT [Cr rows][fs * Cd]   : WcT[cr][j*Cd + cd] = conv_w[cd][cr][j]
//   kind 2: WpT [rows >= Cl][Cd]     : WpT[cc][cd] = cond_w[cd][cc] for the Cl time-varying
//           condition channels, zero rows beyond
__device__ __forceinline__ void
pack_wt(int kind, const float* __restrict__ w0, const float* __restrict__ w1,
        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int rows, int K,
        int Cr, int Cs, int Cd, int Cc, int Cl, int fs, int f16) {
  const int Ch = Cd / 2;
  const int64_t n = (int64_t)rows * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / K), k = (int)(e % K);
    float v;
    if (kind == 0) {
      v = (k < Cr) ? w0[(int64_t)k * Ch + r] : w1[(int64_t)(k - Cr) * Ch + r];
    } else if (kind == 1) {
      const int j = k / Cd, cd = k % Cd;
      v = w0[((int64_t)cd * Cr + r) * fs + j];
    } else {
      v = (r < Cl) ? w0[(int64_t)k * Cc + r] : 0.0f;   // only the Cl time-varying condition rows
    }
    __nv_bfloat16 h, l;
    split_16(v, f16, h, l);
    hi[e] = h;
    if (lo) lo[e] = l;
  }
}

// the three transposed operands of up to PW_MAX blocks in one launch: grid (x, blocks)
constexpr int PW_MAX = 32;
struct PackWtArgs {
  const float* conv_w[PW_MAX];
  const float* cond_w[PW_MAX];
  const float* res_w[PW_MAX];
  const float* skip_w[PW_MAX];
};
__global__ void __launch_bounds__(256)
pack_wt_kernel(const __grid_constant__ PackWtArgs A, __nv_bfloat16* __restrict__ w2t_hi,
               __nv_bfloat16* __restrict__ w2t_lo, __nv_bfloat16* __restrict__ wct_hi,
               __nv_bfloat16* __restrict__ wct_lo, __nv_bfloat16* __restrict__ wpt_hi,
               __nv_bfloat16* __restrict__ wpt_lo, int64_t stride, int wp_rows, int Cr, int Cs,
               int Cd, int Cc, int Cl, int fs, int f16) {
  const int i = blockIdx.y;
  const int64_t o = (int64_t)i * stride;   // elements
  pack_wt(0, A.res_w[i], A.skip_w[i], w2t_hi + o, w2t_lo ? w2t_lo + o : nullptr, Cd / 2, Cr + Cs, Cr,
          Cs, Cd, Cc, Cl, fs, f16);
  pack_wt(1, A.conv_w[i], nullptr, wct_hi + o, wct_lo ? wct_lo + o : nullptr, Cr, fs * Cd, Cr, Cs, Cd,
          Cc, Cl, fs, f16);
  pack_wt(2, A.cond_w[i], nullptr, wpt_hi + o, wpt_lo ? wpt_lo + o : nullptr, wp_rows, Cd, Cr, Cs, Cd,
          Cc, Cl, fs, f16);
}

// bias gradients: g0[c] (and g1[c]) += sum over all (b,t) rows of a time-major plane pair
__global__ void __launch_bounds__(256)
colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                     float* __restrict__ g0, float* __restrict__ g1, int C, int64_t rows,
                     int rows_per_block, int valid, int f16, const float* __restrict__ scale) {
  const float inv = scale ? scale[1] : 1.0f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  for (int c2 = threadIdx.x; c2 < C / 2; c2 += blockDim.x) {
    float s0 = 0.0f, s1 = 0.0f;
    for (int64_t r = r0; r < r1; ++r) {
      float v0, v1;
      unpack_pair_f(*reinterpret_cast<const uint32_t*>(hi + r * C + 2 * c2), f16, v0, v1);
      s0 += v0;
      s1 += v1;
      if (lo) {
        unpack_pair_f(*reinterpret_cast<const uint32_t*>(lo + r * C + 2 * c2), f16, v0, v1);
        s0 += v0;
        s1 += v1;
      }
    }
    s0 *= inv;
    s1 *= inv;
    if (2 * c2 < valid) {
      atomicAdd(g0 + 2 * c2, s0);
      if (g1) atomicAdd(g1 + 2 * c2, s0);
    }
    if (2 * c2 + 1 < valid) {
      atomicAdd(g0 + 2 * c2 + 1, s1);
      if (g1) atomicAdd(g1 + 2 * c2 + 1, s1);
    }
  }
}

__global__ void add_vec_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// ---- gradients that flow through the per-(block, item) gate bias (resblock_tc.cu, gbias_kernel) ----
// S[i][b][m] = sum_t gh_i[b, t, m] (the constant-one column of the weight-gradient GEMM).
//   conv_b_i[m], cond_b_i[m]  += sum_b S[i][b][m]
//   cond_w_i[m][Cl + g]       += sum_b S[i][b][m] * glob[b][g]
//   g_glob[b][g]              += sum_i sum_m cond_w_i[m][Cl + g] * S[i][b][m]
constexpr int TAIL_MAX = 32;
struct TailArgs {
  const float* cond_w[TAIL_MAX];
  float* g_conv_b[TAIL_MAX];
  float* g_cond_b[TAIL_MAX];
  float* g_cond_w[TAIL_MAX];
};
// grid (blocks, Cd), Cg' = max(Cg, 1) threads: thread g of CTA (i, m)
__global__ void __launch_bounds__(256)
bias_tail_kernel(const __grid_constant__ TailArgs A, const float* __restrict__ S,
                 const float* __restrict__ glob, int B, int Cd, int Cc, int Cg, int blk0) {
  const int i = blockIdx.x, m = blockIdx.y;
  const float* Sb = S + ((int64_t)(blk0 + i) * B) * Cd + m;        // S[i][b][m], stride Cd over b
  if (threadIdx.x == 0) {
    float sb = 0.0f;
    for (int b = 0; b < B; ++b) sb += Sb[(int64_t)b * Cd];
    A.g_conv_b[i][m] += sb;
    A.g_cond_b[i][m] += sb;
  }
  float* gw = A.g_cond_w[i] + (int64_t)m * Cc + (Cc - Cg);
  for (int g = threadIdx.x; g < Cg; g += blockDim.x) {
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) acc = fmaf(Sb[(int64_t)b * Cd], __ldg(glob + (int64_t)b * Cg + g), acc);
    gw[g] += acc;
  }
}
// grid (Cd / 32, blocks), 256 threads: CTA (mc, i) owns the 32 dilated channels m0 = 32 mc .. of
// block i; thread -> (item b, global channel g) pairs; g_glob[b][g] += sum_m cond_w_i[m][Cl + g] *
// S[i][b][m] over its 32 channels (weights coalesced over g, S broadcast from shared memory), one
// atomic per output and CTA.  (A first version looped one thread over all 512 channels: 97 us.)
__global__ void __launch_bounds__(256)
gglob_tail_kernel(const __grid_constant__ TailArgs A, const float* __restrict__ S,
                  float* __restrict__ g_glob, int B, int Cd, int Cc, int Cg, int blk0) {
  extern __shared__ float Ss[];                          // [B][32]
  const int m0 = blockIdx.x * 32, i = blockIdx.y;
  for (int e = threadIdx.x; e < B * 32; e += blockDim.x) {
    const int b = e >> 5, mm = e & 31;
    Ss[e] = (m0 + mm < Cd) ? S[((int64_t)(blk0 + i) * B + b) * Cd + m0 + mm] : 0.0f;
  }
  __syncthreads();
  const float* w = A.cond_w[i] + (int64_t)m0 * Cc + (Cc - Cg);
  for (int e = threadIdx.x; e < B * Cg; e += blockDim.x) {
    const int b = e / Cg, g = e - b * Cg;
    float acc = 0.0f;
#pragma unroll 8
    for (int mm = 0; mm < 32; ++mm)
      if (m0 + mm < Cd) acc = fmaf(__ldg(w + (int64_t)mm * Cc + g), Ss[b * 32 + mm], acc);
    atomicAdd(g_glob + (int64_t)b * Cg + g, acc);
  }
}

int pack_act_launch(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int T,
                    int f16, const float* scale, cudaStream_t stream);   // resblock_tc.cu
int pack_act_launch_ex(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int T,
                       int pitch, int relu, int f16, const float* scale,
                       cudaStream_t stream);   // resblock_tc.cu

// ---- power-of-two gradient scale of the fp16 backward ------------------------------------
// fp16 planes have 5 exponent bits; gradients of a mean loss over B*T samples are ~1e-5 and
// smaller, i.e. subnormal in fp16.  Every backward entry point therefore measures max|g| of its
// incoming gradient on the device and scales the planes by s = 2^k so that the maximum lands in
// [2^5, 2^6): exact (power of two), 2^10 of headroom against growth through the stack, 2^-30 of
// the maximum still representable.  The backward is linear in g, so every fp32 result is simply
// multiplied by 1/s in the epilogue that writes it.  No host synchronisation.
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, int64_t n, unsigned* __restrict__ out) {
  float m = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}
__global__ void grad_scale_kernel(const unsigned* __restrict__ amax, float* __restrict__ scale) {
  const float a = __uint_as_float(*amax);
  float s = 1.0f;
  if (a > 0.0f && a < 3.0e38f) {
    int e;
    frexpf(a, &e);            // a = m * 2^e with m in [0.5, 1)
    int k = 6 - e;
    k = k < -60 ? -60 : (k > 60 ? 60 : k);
    s = ldexpf(1.0f, k);
  }
  scale[0] = s;
  scale[1] = 1.0f / s;
}
// scale_mem: 4 floats of workspace {amax bits, s, 1/s, pad}; returns the {s, 1/s} pointer
static int make_grad_scale(const float* g0, int64_t n0, const float* g1, int64_t n1, float* scale_mem,
                           cudaStream_t stream, const float** out) {
  unsigned* amax = reinterpret_cast<unsigned*>(scale_mem);
  VQW_CHECK_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned), stream));
  amax_kernel<<<148 * 8, 256, 0, stream>>>(g0, n0, amax);
  VQW_CHECK_LAUNCH("amax_kernel");
  if (g1 != nullptr) {
    amax_kernel<<<148 * 8, 256, 0, stream>>>(g1, n1, amax);
    VQW_CHECK_LAUNCH("amax_kernel");
  }
  grad_scale_kernel<<<1, 1, 0, stream>>>(amax, scale_mem + 1);
  VQW_CHECK_LAUNCH("grad_scale_kernel");
  *out = scale_mem + 1;
  return 0;
}

// out[r][k] (out_rows x out_cols bf16 planes) = src[r][k] or src[k][r] (transpose), zero padded
__global__ void __launch_bounds__(256)
pack_mat_kernel(const float* __restrict__ src, int src_rows, int src_cols, int transpose,
                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int out_rows,
                int out_cols, int f16) {
  const int64_t n = (int64_t)out_rows * out_cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / out_cols), k = (int)(e % out_cols);
    const int sr = transpose ? k : r, sc = transpose ? r : k;
    const float v = (sr < src_rows && sc < src_cols) ? src[(int64_t)sr * src_cols + sc] : 0.0f;
    __nv_bfloat16 h, l;
    split_16(v, f16, h, l);
    hi[e] = h;
    if (lo) lo[e] = l;
  }
}

}  // namespace tc

static inline int64_t al(int64_t v) { return (v + 1023) / 1024 * 1024; }
static inline int pad256(int v) { return (v + 255) / 256 * 256; }   // TMA boxes never exceed the tensor

// workspace: g_skip planes, gh planes, g_res planes (ping-pong), transposed weight planes per block
struct BwdLayout {
  int64_t total;
  int64_t gs_p[2], gh_p[2], gr_p[2][2];
  int64_t gs_sum;   // column sum of g_skip (the same bias gradient for every block)
  int64_t colS;     // S[i][b][m]: per-item column sums of gh (n_blocks * B * Cd floats)
  int64_t scale;    // 4 floats: gradient-scale scratch of the fp16 mode
  int64_t wg_partial;   // [B * chunks][job pairs][256][256] fp32 weight-gradient tiles of one block
  int64_t w2t[2], wct[2], wpt[2];
  int64_t wstride;
};

// time chunks of the persistent weight-gradient GEMM (work units = job pairs x items x chunks)
static int wgrad_chunks(int T) {
  int chunks = getenv("VQW_WGRAD_CHUNKS") ? atoi(getenv("VQW_WGRAD_CHUNKS")) : 1;
  if (chunks < 1) chunks = 1;
  if (chunks > 8) chunks = 8;
  while (chunks > 1 && ceil_div(T, tc::BK) / chunks < 16) chunks /= 2;
  return chunks;
}

static BwdLayout bwd_layout(const vqw_resnet_desc& d) {
  BwdLayout L;
  int64_t off = 0;
  const int64_t N = (int64_t)d.B * d.T;
  auto take = [&](int64_t bytes) { int64_t o = off; off += al(bytes); return o; };
  for (int p = 0; p < 2; ++p) L.gs_p[p] = take(N * d.Cs * 2);
  L.gs_sum = take((int64_t)d.Cs * 4);
  L.colS = take((int64_t)d.n_blocks * d.B * d.Cd * 4);
  L.scale = take(16);
  for (int p = 0; p < 2; ++p) L.gh_p[p] = take(N * d.Cd * 2);
  for (int q = 0; q < 2; ++q)
    for (int p = 0; p < 2; ++p) L.gr_p[q][p] = take(N * d.Cr * 2);
  const int64_t wbase = off;
  for (int p = 0; p < 2; ++p) L.w2t[p] = take((int64_t)(d.Cd / 2) * (d.Cr + d.Cs) * 2);
  for (int p = 0; p < 2; ++p) L.wct[p] = take((int64_t)d.Cr * d.fs * d.Cd * 2);
  for (int p = 0; p < 2; ++p) L.wpt[p] = take((int64_t)pad256(cond_local(d)) * d.Cd * 2);
  L.wstride = off - wbase;
  off = wbase + L.wstride * d.n_blocks;
  L.wg_partial = take((int64_t)d.B * wgrad_chunks(d.T) * (tc::MAX_JOBS / 2) * 65536 * 4);
  L.total = off + 1024;
  return L;
}

int64_t resnet_backward_tc_workspace(const vqw_resnet_desc& d) { return bwd_layout(d).total; }

int resnet_backward_tc(const vqw_resnet_desc& d, const float* g_skip, const float* g_last_res,
                       float* const* gate_tanh, float* const* gate_sig,
                       const vqw_resblock_weights* weights, float* gx0, float* gcond,
                       const vqw_resblock_wgrads* wgrads, void* workspace, const void* saved,
                       cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(resnet_tc_supported(d), "tcgen05 backward: unsupported channel counts");
  VQW_REQUIRE(workspace && saved && g_skip && gate_tanh && gate_sig && weights && wgrads && gcond,
              "vqw_resnet_backward: null argument");
  VQW_REQUIRE(d.fs <= MAX_SEG, "tcgen05 backward: filter_size > %d unsupported", MAX_SEG);
  const bool x3 = vqw_mode_x3(d.mode);
  const int f16 = vqw_mode_f16(d.mode) ? 1 : 0;
  const bool xlo = x3 || f16;   // the g_res stream keeps hi + lo (addend of GX), like x in the forward
  const BwdLayout L = bwd_layout(d);
  const TcSaved S = tc_saved_layout(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(al((int64_t)(uintptr_t)workspace));
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(al((int64_t)(uintptr_t)saved));
  auto P16 = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto LO = [&](int64_t off) { return x3 ? P16(off) : nullptr; };
  auto LOX = [&](int64_t off) { return xlo ? P16(off) : nullptr; };
  const int B = d.B, T = d.T, Cr = d.Cr, Cd = d.Cd, Cs = d.Cs, Cc = d.Cc, Ch = d.Cd / 2, fs = d.fs;
  const int Cl = cond_local(d), CP = cond_pitch(d), Cg = d.Cg;
  VQW_REQUIRE(Cg == 0 || (d.cond_global && d.g_cond_global),
              "vqw_resnet_backward: Cg > 0 needs cond_global and g_cond_global");
  float* colS = reinterpret_cast<float*>(ws + L.colS);
  VQW_CHECK_CUDA(cudaMemsetAsync(colS, 0, sizeof(float) * (size_t)d.n_blocks * B * Cd, stream));
  // persistent double-buffered GX / GATE_BWD (needs the CTA-pair weight boxes); VQW_GEMM_PERSIST=0: off
  const bool persist = gemm_pair_enabled() &&
                       !(getenv("VQW_GEMM_PERSIST") && getenv("VQW_GEMM_PERSIST")[0] == '0');
  const int64_t NROWS = (int64_t)B * T;
  const int RPB = 256;   // rows per block of the bias column sums
  const int CS_GRID = (int)((NROWS + RPB - 1) / RPB);

  // ---- once per call: g_skip planes, transposed weight planes of every block ----
  const float* gscale = nullptr;   // device {s, 1/s} (fp16 mode) -- see make_grad_scale
  if (f16) {
    if (int rc = make_grad_scale(g_skip, (int64_t)B * Cs * T, g_last_res, (int64_t)B * Cr * T,
                                 reinterpret_cast<float*>(ws + L.scale), stream, &gscale))
      return rc;
  }
  if (int rc = pack_act_launch(g_skip, P16(L.gs_p[0]), LO(L.gs_p[1]), B, Cs, T, f16, gscale, stream))
    return rc;
  float* gs_sum = reinterpret_cast<float*>(ws + L.gs_sum);
  VQW_CHECK_CUDA(cudaMemsetAsync(gs_sum, 0, sizeof(float) * Cs, stream));
  colsum_planes_kernel<<<CS_GRID, 256, 0, stream>>>(P16(L.gs_p[0]), LO(L.gs_p[1]), gs_sum, nullptr, Cs,
                                                   NROWS, RPB, Cs, f16, gscale);
  VQW_CHECK_LAUNCH("colsum_planes_kernel(g_skip)");
  for (int i0 = 0; i0 < d.n_blocks; i0 += PW_MAX) {
    PackWtArgs A = {};
    const int nb = d.n_blocks - i0 < PW_MAX ? d.n_blocks - i0 : PW_MAX;
    for (int i = 0; i < nb; ++i) {
      const vqw_resblock_weights& w = weights[i0 + i];
      A.conv_w[i] = w.conv_w; A.cond_w[i] = w.cond_w; A.res_w[i] = w.res_w; A.skip_w[i] = w.skip_w;
    }
    const int64_t wo = i0 * L.wstride;
    pack_wt_kernel<<<dim3(74, nb), 256, 0, stream>>>(
        A, P16(L.w2t[0] + wo), LO(L.w2t[1] + wo), P16(L.wct[0] + wo), LO(L.wct[1] + wo),
        P16(L.wpt[0] + wo), LO(L.wpt[1] + wo), L.wstride / 2, pad256(Cl), Cr, Cs, Cd, Cc, Cl, fs, f16);
    VQW_CHECK_LAUNCH("pack_wt_kernel");
  }

  // K-major (64-byte swizzle) map pair of a time-major plane / weight plane
  auto mapk = [&](CUtensorMap* m, const void* hi, const void* lo, uint64_t inner, uint64_t rows,
                  uint64_t batch, uint32_t box_rows) -> int {
    if (int rc = make_map(&m[0], hi, 3, inner, rows, batch, box_rows)) return rc;
    return make_map(&m[1], x3 ? lo : hi, 3, inner, rows, batch, box_rows);
  };
  // MN-major (128-byte swizzle) map pair of a time-major plane
  auto mapmn = [&](CUtensorMap* m, const void* hi, const void* lo, uint64_t C) -> int {
    if (int rc = make_map_mn(&m[0], hi, C, C, T, B)) return rc;
    return make_map_mn(&m[1], x3 ? lo : hi, C, C, T, B);
  };

  int cur = 0;
  bool have_gres = false;
  if (g_last_res != nullptr) {
    if (int rc = pack_act_launch(g_last_res, P16(L.gr_p[0][0]), LOX(L.gr_p[0][1]), B, Cr, T, f16,
                                 gscale, stream))
      return rc;
    have_gres = true;
  }

  for (int i = d.n_blocks - 1; i >= 0; --i) {
    const vqw_resblock_wgrads& gw = wgrads[i];
    const int64_t wo = i * L.wstride;
    VQW_REQUIRE((x3 || gate_tanh[i]) && gate_sig[i], "vqw_resnet_backward: block %d saved gates", i);
    const int nxt = cur ^ 1;
    const int dil = d.dilations[i];
    const uint8_t* xp_hi = sv + S.x0 + i * S.x_stride;
    const uint8_t* xp_lo = xp_hi + S.x_plane;
    const uint8_t* zp_hi = sv + S.z0 + i * S.z_stride;
    const uint8_t* zp_lo = zp_hi + S.z_plane;

    // ---- A1: gz GEMM + gate derivative -> gh planes ----
    {
      Maps maps;
      GemmParams P = {};
      if (int rc = mapk(&maps.m[0], ws + L.gr_p[cur][0], ws + L.gr_p[cur][1], Cr, T, B, TM)) return rc;
      if (int rc = mapk(&maps.m[2], ws + L.gs_p[0], ws + L.gs_p[1], Cs, T, B, TM)) return rc;
      if (int rc = mapk(&maps.m[4], ws + L.w2t[0] + wo, ws + L.w2t[1] + wo, Cr + Cs, Ch, 1, WROWS())) return rc;
      for (int k = 6; k < NMAPS; ++k) maps.m[k] = maps.m[k % 6];
      int n = 0;
      if (have_gres) P.seg[n++] = Seg{0, 2, Cr / BK, 0, 0, 0, 0};
      P.seg[n++] = Seg{1, 2, Cs / BK, 0, Cr, 0, 0};
      P.nseg = n;
      P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T;
      P.f0 = x3 ? nullptr : gate_tanh[i]; P.f1 = gate_sig[i];
      P.z_hi = reinterpret_cast<const __nv_bfloat16*>(zp_hi);
      P.z_lo = reinterpret_cast<const __nv_bfloat16*>(zp_lo);
      P.p_hi = P16(L.gh_p[0]); P.p_lo = LO(L.gh_p[1]);
      P.Cout = Cd;
      if (x3 && persist) {
        // persistent, double-buffered, TMA-staged epilogue (tc_time_persistent_kernel)
        if (int rc = make_map(&maps.m[6], zp_hi, 3, Ch, T, B, TM)) return rc;
        if (int rc = make_map(&maps.m[7], zp_lo, 3, Ch, T, B, TM)) return rc;
        if (int rc = make_map(&maps.m[8], ws + L.gh_p[0], 3, Cd, T, B, TM)) return rc;
        if (int rc = make_map(&maps.m[9], ws + L.gh_p[1], 3, Cd, T, B, TM)) return rc;
        if (int rc = launch_persistent<EPI_GATE_BWD>(maps, P, dim3(ceil_div(T, TM), 1, B), stream)) return rc;
      } else {
        if (int rc = launch_gemm<EPI_GATE_BWD>(maps, P, dim3(ceil_div(T, TM), 1, B), stream)) return rc;
      }
    }
    // ---- A2: gx = g_res + sum_j Wc_j^T gh[t + s_j] ; gcond += Wp^T gh ----
    {
      Maps maps;
      if (int rc = mapk(&maps.m[0], ws + L.gh_p[0], ws + L.gh_p[1], Cd, T, B, TM)) return rc;
      if (int rc = mapk(&maps.m[2], ws + L.wct[0] + wo, ws + L.wct[1] + wo, (uint64_t)fs * Cd, Cr, 1, WROWS())) return rc;
      if (int rc = mapk(&maps.m[4], ws + L.wpt[0] + wo, ws + L.wpt[1] + wo, Cd, pad256(Cl), 1, WROWS())) return rc;
      for (int k = 6; k < NMAPS; ++k) maps.m[k] = maps.m[k % 6];
      const bool run_gx = i > 0 || gx0 != nullptr;
      if (run_gx) {
        GemmParams P = {};
        int n = 0;
        for (int j = 0; j < fs; ++j) P.seg[n++] = Seg{0, 1, Cd / BK, 0, j * Cd, dil * (fs - 1 - j), 0};
        P.nseg = n;
        P.x3 = x3; P.f16 = f16; P.add_lo = xlo; P.scale = gscale; P.B = B; P.T = T;
        if (have_gres) { P.a_hi = P16(L.gr_p[cur][0]); P.a_lo = LOX(L.gr_p[cur][1]); }
        if (i > 0) { P.p_hi = P16(L.gr_p[nxt][0]); P.p_lo = LOX(L.gr_p[nxt][1]); }
        else P.o0 = gx0;
        P.Cout = Cr;
        // this gx is the g_res of block i-1: its time sum is that block's res_b gradient
        if (i > 0) P.colsum = wgrads[i - 1].res_b;
        // staged epilogue (shared-memory tiles + TMA) whenever the result goes to planes
        if (i > 0 && !(getenv("VQW_GEMM_STAGE") && getenv("VQW_GEMM_STAGE")[0] == '0')) {
          P.stage_epi = 1;
          const void* ah = have_gres ? (const void*)(ws + L.gr_p[cur][0]) : (const void*)(ws + L.gr_p[nxt][0]);
          const void* al_ = (have_gres && xlo) ? (const void*)(ws + L.gr_p[cur][1]) : ah;
          if (int rc = make_map_tile(&maps.m[6], ah, Cr, T, B)) return rc;
          if (int rc = make_map_tile(&maps.m[7], al_, Cr, T, B)) return rc;
          if (int rc = make_map_tile(&maps.m[8], ws + L.gr_p[nxt][0], Cr, T, B)) return rc;
          if (int rc = make_map_tile(&maps.m[9], xlo ? ws + L.gr_p[nxt][1] : ws + L.gr_p[nxt][0], Cr, T, B))
            return rc;
        }
        // the ACCUM tiles (gcond += Wp^T gh, K = Cd) ride in the same launch as extra N tiles
        P.alt_y = Cr / TN;
        P.alt_seg = Seg{0, 2, Cd / BK, 0, 0, 0, 0};
        P.alt_out = gcond;
        P.alt_Cout = Cl;
        const dim3 gxgrid(ceil_div(T, TM), Cr / TN + ceil_div(Cl, TN), B);
        if (P.stage_epi && persist) {
          if (int rc = launch_persistent<EPI_GX>(maps, P, gxgrid, stream)) return rc;
        } else {
          if (int rc = launch_gemm<EPI_GX>(maps, P, gxgrid, stream)) return rc;
        }
      } else {
        GemmParams P = {};
        P.nseg = 1;
        P.seg[0] = Seg{0, 2, Cd / BK, 0, 0, 0, 0};
        P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T;
        P.o0 = gcond;                  // (B, Cl, T): the time-varying condition channels
        P.Cout = Cl;
        if (int rc = launch_gemm<EPI_ACCUM>(maps, P, dim3(ceil_div(T, TM), ceil_div(Cl, TN), B), stream))
          return rc;
      }
    }
    // ---- all weight gradients of the block: one grouped launch ----
    {
      Maps maps;
      GemmParams P = {};
      // the persistent kernel runs when every job has an even number of 128-row tiles (CTA pairs);
      // its producer then takes 4-D maps: one copy per 128-channel operand half instead of two
      const bool even_tiles = (ceil_div(Cd, TM) & 1) == 0 && (ceil_div(Cs, TM) & 1) == 0 &&
                              (!have_gres || (ceil_div(Cr, TM) & 1) == 0);
      static const bool wide_on = !(getenv("VQW_WGRAD_WIDE") && getenv("VQW_WGRAD_WIDE")[0] == '0');
      const bool wide = persist && even_tiles && wide_on;
      auto mapmn_w = [&](int pair, const void* hi, const void* lo, uint64_t C) -> int {
        if (wide && C % 128 == 0) {
          P.wide_maps |= 1u << pair;
          if (int rc = make_map_mn4(&maps.m[2 * pair], hi, C, T, B)) return rc;
          return make_map_mn4(&maps.m[2 * pair + 1], x3 ? lo : hi, C, T, B);
        }
        return mapmn(&maps.m[2 * pair], hi, lo, C);
      };
      // pairs: 0 gh, 1 x_i, 2 cond, 3 z_i, 4 g_res, 5 g_skip
      if (int rc = mapmn_w(0, ws + L.gh_p[0], ws + L.gh_p[1], Cd)) return rc;
      if (int rc = mapmn_w(1, xp_hi, xp_lo, Cr)) return rc;
      if (int rc = mapmn(&maps.m[4], sv + S.cond[0], sv + S.cond[1], CP)) return rc;
      if (int rc = mapmn_w(3, zp_hi, zp_lo, Ch)) return rc;
      if (int rc = mapmn_w(4, ws + L.gr_p[cur][0], ws + L.gr_p[cur][1], Cr)) return rc;
      if (int rc = mapmn_w(5, ws + L.gs_p[0], ws + L.gs_p[1], Cs)) return rc;
      P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T;
      // MMA passes per product of the weight-gradient GEMMs (split modes): 3 = hi*hi + lo*hi + hi*lo,
      // 2 = the activations' lo plane is left out, 1 = hi planes only.  One contraction over time
      // per weight, so the rounding of a hi plane (2^-12 per operand in fp16) enters each gradient
      // once and does not accumulate over the depth of the stack.  Default: 3 for bf16 planes
      // (2^-9 per operand is outside the parity bar), 1 for fp16 planes.
      {
        static const int env_passes = getenv("VQW_WGRAD_PASSES") ? atoi(getenv("VQW_WGRAD_PASSES")) : 0;
        const int passes = env_passes >= 1 && env_passes <= 3 ? env_passes : ((x3 && f16) ? 1 : 3);
        P.b_exact = (x3 && passes <= 2) ? 1 : 0;
        P.a_exact = (x3 && passes <= 1) ? 1 : 0;
      }
      P.slabs_per_item = ceil_div(T, BK);
      P.chunks_per_b = 1;
      int nj = 0;
      bool wg_pair = true;
      // `ncols` >= N columns of the product are computed; column `col_n` (if col_out) is the
      // per-item column sum that the constant-one channel of the condition planes produces
      auto add_jobs = [&](int a_map, int M, int b_map, int N, int shift, float* out, long long gm,
                          long long gk, int ncols = 0, float* col_out = nullptr, int col_n = -1) -> int {
        if (ncols < N) ncols = N;
        // m0 runs fastest: consecutive jobs are the two 128-row halves of a 256-row tile (a CTA pair)
        if ((ceil_div(M, TM) & 1) != 0) wg_pair = false;
        for (int n0 = 0; n0 < ncols; n0 += TN)
          for (int m0 = 0; m0 < M; m0 += TM) {
            VQW_REQUIRE(nj < MAX_JOBS, "tcgen05 backward: too many weight-gradient tiles");
            P.jobs[nj++] = Job{a_map, b_map, m0, n0, shift, M, N, out, gm, gk, col_out, col_n,
                               (long long)Cd};
          }
        return 0;
      };
      for (int j = 0; j < fs; ++j)
        if (int rc = add_jobs(0, Cd, 1, Cr, -dil * (fs - 1 - j), gw.conv_w + j, (long long)Cr * fs, fs))
          return rc;
      if (int rc = add_jobs(0, Cd, 2, Cl, 0, gw.cond_w, Cc, 1, Cl + 1, colS + (int64_t)i * B * Cd, Cl))
        return rc;
      if (have_gres)
        if (int rc = add_jobs(4, Cr, 3, Ch, 0, gw.res_w, Ch, 1)) return rc;
      if (int rc = add_jobs(5, Cs, 3, Ch, 0, gw.skip_w, Ch, 1)) return rc;
      P.njobs = nj;
      VQW_REQUIRE(!wide || wg_pair, "tcgen05 backward: weight-gradient tiles are not pairable");
      if (wg_pair && persist) {
        // every (job pair, item, time chunk) tile goes to its own slot of wg_partial and a second
        // kernel sums the slots in a fixed order (round 2: with fp32 atomics straight into gW this
        // launch was bound by 18 M scattered atomics, not by its MMAs; VQW_WGRAD_ATOMICS=1 restores it)
        static const bool use_atomics = getenv("VQW_WGRAD_ATOMICS") && getenv("VQW_WGRAD_ATOMICS")[0] == '1';
        const int chunks = wgrad_chunks(T);
        // hi planes only: several 32-step slabs per ring stage (VQW_WGRAD_SUB = 1..4, default 2)
        static const int sub_env = getenv("VQW_WGRAD_SUB") ? atoi(getenv("VQW_WGRAD_SUB")) : 0;
        const int sub = (sub_env >= 1 && sub_env <= 4) ? sub_env : 2;
        P.wg_sub = (x3 && P.a_exact && P.b_exact) ? sub : 1;
        P.chunks_per_b = chunks;
        P.slabs_per_item = ceil_div(ceil_div(ceil_div(T, BK), chunks), P.wg_sub);
        float* partial = reinterpret_cast<float*>(ws + L.wg_partial);
        P.wg_partial = use_atomics ? nullptr : partial;
        P.dbg_mode = getenv("VQW_WGRAD_DEBUG") ? atoi(getenv("VQW_WGRAD_DEBUG")) : 0;
        if (int rc = launch_persistent<EPI_WGRAD>(maps, P, dim3(nj, chunks, B), stream)) return rc;
        if (!use_atomics) {
          wgrad_reduce_kernel<<<dim3(nj / 2, 256 / 4), 256, 0, stream>>>(P, partial, nj / 2, chunks, B);
          VQW_CHECK_LAUNCH("wgrad_reduce_kernel");
        }
      } else {
        if (int rc = launch_gemm<EPI_WGRAD>(maps, P, dim3(nj, 1, B), stream, wg_pair)) return rc;
      }
    }
    // ---- bias gradients: gh's column sums come out of the grouped launch above (column Cl of
    // the condition job, per item) and are folded in after the loop; g_res, g_skip here ----
    if (have_gres && i == d.n_blocks - 1) {   // g_last_res: the only g_res no GX epilogue produced
      colsum_planes_kernel<<<CS_GRID, 256, 0, stream>>>(P16(L.gr_p[cur][0]), LOX(L.gr_p[cur][1]),
                                                       gw.res_b, nullptr, Cr, NROWS, RPB, Cr, f16,
                                                       gscale);
      VQW_CHECK_LAUNCH("colsum_planes_kernel(g_res)");
    }
    add_vec_kernel<<<ceil_div(Cs, 256), 256, 0, stream>>>(gw.skip_b, gs_sum, Cs);
    VQW_CHECK_LAUNCH("add_vec_kernel(skip_b)");
    // ---- conv_b / cond_b, the global-condition columns of cond_w, g_cond_global: per block, so
    // that every gradient of block i is final here (block_events) ----
    {
      TailArgs A = {};
      A.cond_w[0] = weights[i].cond_w;
      A.g_conv_b[0] = gw.conv_b;
      A.g_cond_b[0] = gw.cond_b;
      A.g_cond_w[0] = gw.cond_w;
      bias_tail_kernel<<<dim3(1, Cd), Cg > 128 ? 256 : 128, 0, stream>>>(A, colS, d.cond_global, B, Cd, Cc,
                                                                       Cg, i);
      VQW_CHECK_LAUNCH("bias_tail_kernel");
      if (Cg > 0) {
        gglob_tail_kernel<<<dim3(ceil_div(Cd, 32), 1), 256, sizeof(float) * B * 32, stream>>>(
            A, colS, d.g_cond_global, B, Cd, Cc, Cg, i);
        VQW_CHECK_LAUNCH("gglob_tail_kernel");
      }
    }
    if (d.block_events != nullptr && d.block_events[i] != nullptr)
      VQW_CHECK_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(d.block_events[i]), stream));
    have_gres = true;
    cur = nxt;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// WaveNet output head on the tensor cores: y = proj2(relu(proj1(relu(skip)))), modules.py:155-159
// ---------------------------------------------------------------------------------------
static inline int qpad(int Q) { int v = (Q + 31) / 32 * 32; return v < 64 ? 64 : v; }

bool head_tc_supported(const vqw_head_desc& d) {
  return d.Cs % tc::TN == 0 && d.Q >= 1 && d.T >= tc::TM && d.T % 8 == 0 && d.B >= 1 && d.B <= 65535;
}

struct HeadLayout {   // workspace (both directions)
  int64_t w1[2], w2[2], w1t[2], w2t[2], gy[2], gh1[2], scale, total;
};
static HeadLayout head_layout(const vqw_head_desc& d) {
  HeadLayout L;
  int64_t off = 0;
  const int64_t N = (int64_t)d.B * d.T;
  auto take = [&](int64_t bytes) { int64_t o = off; off += al(bytes); return o; };
  for (int p = 0; p < 2; ++p) L.w1[p] = take((int64_t)d.Cs * d.Cs * 2);
  for (int p = 0; p < 2; ++p) L.w2[p] = take((int64_t)pad256(d.Q) * d.Cs * 2);
  for (int p = 0; p < 2; ++p) L.w1t[p] = take((int64_t)d.Cs * d.Cs * 2);
  for (int p = 0; p < 2; ++p) L.w2t[p] = take((int64_t)d.Cs * qpad(d.Q) * 2);
  for (int p = 0; p < 2; ++p) L.gy[p] = take(N * qpad(d.Q) * 2);
  for (int p = 0; p < 2; ++p) L.gh1[p] = take(N * d.Cs * 2);
  L.scale = take(16);
  L.total = off + 1024;
  return L;
}
struct HeadSaved { int64_t s[2], h1[2], gy[2], scale, total; };
static HeadSaved head_saved(const vqw_head_desc& d) {
  HeadSaved S;
  const int64_t plane = al((int64_t)d.B * d.T * d.Cs * 2);
  const int64_t gplane = al((int64_t)d.B * d.T * qpad(d.Q) * 2);
  S.s[0] = 0; S.s[1] = plane; S.h1[0] = 2 * plane; S.h1[1] = 3 * plane;
  S.gy[0] = 4 * plane; S.gy[1] = 4 * plane + gplane;     // d loss / d y planes (fused loss)
  S.scale = 4 * plane + 2 * gplane;                      // {1/g, g}: upstream gradient of the loss
  S.total = S.scale + 1024 + 1024;
  return S;
}

int64_t head_tc_workspace(const vqw_head_desc& d) { return head_layout(d).total; }
int64_t head_tc_saved_bytes(const vqw_head_desc& d) { return head_saved(d).total; }

namespace tc {
__global__ void set_double_kernel(double* p, double v) { *p = v; }
// scale[2] = 1 (bf16 planes) or, for fp16 planes, a power of two near n * 2^up for the n labels the
// loss is averaged over: d loss / d y = (p - onehot) / n would be subnormal in fp16; with up = 7 the
// softmax gradient planes hold values up to 256 (hi AND lo plane in the normal range), up = 0 leaves
// the mixture-of-logistics gradient (|g| up to ~1e3 per position) four decades of headroom
__global__ void plane_scale_kernel(const double* __restrict__ count, float* __restrict__ scale, int f16,
                                   int up) {
  const double n = *count;
  scale[2] = (f16 && n >= 1.0) ? exp2f(floorf(log2f((float)n)) + (float)up) : 1.0f;
}
__global__ void __launch_bounds__(256)
count_labels_kernel(const int32_t* __restrict__ tgt, int64_t n, int Q, double* __restrict__ count) {
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
}  // namespace tc

// Loss description of the fused head: exactly one of tgt_i (softmax cross entropy, train.py:95) and
// tgt_f (mixture of logistics, modules.py:169-230) is set; loss = {loss, valid count} (zeroed by
// the caller); the logits go to `y` only if it is non-null.
struct HeadLoss {
  const int32_t* tgt_i;
  const float* tgt_f;
  double* loss;
  int quantize;
  float log_scale_min;
};

static int head_forward_impl(const vqw_head_desc& d, const float* skip, const float* W1, const float* b1,
                             const float* W2, const float* b2, float* y, void* workspace, void* saved,
                             const HeadLoss* hl, cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(head_tc_supported(d), "tcgen05 head: needs skip_channels %% 256 == 0, T >= 128, T %% 8 == 0");
  VQW_REQUIRE(skip && W1 && b1 && W2 && b2 && (y || hl) && workspace, "vqw_head_forward: null argument");
  VQW_REQUIRE(!hl || (saved && d.Q <= TN && (hl->tgt_i != nullptr) != (hl->tgt_f != nullptr) && hl->loss),
              "vqw_head_loss_forward: needs `saved`, Q <= 256 and exactly one target");
  VQW_REQUIRE(!hl || !hl->tgt_f || (d.Q % 3 == 0 && d.Q <= 32),
              "vqw_head_loss_forward: the mixture-of-logistics head has 3 * n_mix <= 32 outputs");
  const bool x3 = vqw_mode_x3(d.mode);
  const int f16 = vqw_mode_f16(d.mode) ? 1 : 0;
  const HeadLayout L = head_layout(d);
  const HeadSaved S = head_saved(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(al((int64_t)(uintptr_t)workspace));
  // without a `saved` buffer (inference) the activation planes live in the gradient scratch
  uint8_t* sv = saved ? reinterpret_cast<uint8_t*>(al((int64_t)(uintptr_t)saved)) : nullptr;
  auto W16 = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  __nv_bfloat16* s_hi = sv ? reinterpret_cast<__nv_bfloat16*>(sv + S.s[0]) : W16(L.gh1[0]);
  __nv_bfloat16* s_lo = sv ? reinterpret_cast<__nv_bfloat16*>(sv + S.s[1]) : W16(L.gh1[1]);
  __nv_bfloat16* h_hi = sv ? reinterpret_cast<__nv_bfloat16*>(sv + S.h1[0]) : W16(L.gy[0]);
  __nv_bfloat16* h_lo = sv ? reinterpret_cast<__nv_bfloat16*>(sv + S.h1[1]) : W16(L.gy[1]);
  VQW_REQUIRE(sv || qpad(d.Q) >= d.Cs, "vqw_head_forward: inference scratch too small; pass `saved`");
  const int B = d.B, T = d.T, Cs = d.Cs, Q = d.Q;
  auto LOW = [&](__nv_bfloat16* p) { return x3 ? p : nullptr; };
  // relu(skip) planes, weight planes
  if (int rc = pack_act_launch_ex(skip, s_hi, LOW(s_lo), B, Cs, T, Cs, 1, f16, nullptr, stream)) return rc;
  pack_mat_kernel<<<148, 256, 0, stream>>>(W1, Cs, Cs, 0, W16(L.w1[0]), LOW(W16(L.w1[1])), Cs, Cs, f16);
  VQW_CHECK_LAUNCH("pack_mat_kernel(W1)");
  pack_mat_kernel<<<148, 256, 0, stream>>>(W2, Q, Cs, 0, W16(L.w2[0]), LOW(W16(L.w2[1])), pad256(Q), Cs,
                                           f16);
  VQW_CHECK_LAUNCH("pack_mat_kernel(W2)");
  auto mapk = [&](CUtensorMap* m, const void* hi, const void* lo, uint64_t inner, uint64_t rows,
                  uint64_t batch, uint32_t box_rows) -> int {
    if (int rc = make_map(&m[0], hi, 3, inner, rows, batch, box_rows)) return rc;
    return make_map(&m[1], x3 ? lo : hi, 3, inner, rows, batch, box_rows);
  };
  {   // h1 = relu(W1 s + b1) -> planes
    Maps maps;
    if (int rc = mapk(&maps.m[0], s_hi, s_lo, Cs, T, B, TM)) return rc;
    if (int rc = mapk(&maps.m[2], ws + L.w1[0], ws + L.w1[1], Cs, Cs, 1, WROWS())) return rc;
    for (int k = 4; k < NMAPS; ++k) maps.m[k] = maps.m[k % 4];
    GemmParams P = {};
    P.nseg = 1; P.seg[0] = Seg{0, 1, Cs / BK, 0, 0, 0, 0};
    P.x3 = x3; P.f16 = f16; P.B = B; P.T = T; P.Cout = Cs; P.bias = b1; P.relu = 1;
    P.p_hi = h_hi; P.p_lo = LOW(h_lo);
    if (int rc = launch_gemm<EPI_HEAD>(maps, P, dim3(ceil_div(T, TM), Cs / TN, B), stream)) return rc;
  }
  {   // y = W2 h1 + b2 -> fp32 (B,Q,T), or straight into the loss
    Maps maps;
    if (int rc = mapk(&maps.m[0], h_hi, h_lo, Cs, T, B, TM)) return rc;
    if (int rc = mapk(&maps.m[2], ws + L.w2[0], ws + L.w2[1], Cs, pad256(Q), 1, WROWS())) return rc;
    for (int k = 4; k < NMAPS; ++k) maps.m[k] = maps.m[k % 4];
    GemmParams P = {};
    P.nseg = 1; P.seg[0] = Seg{0, 1, Cs / BK, 0, 0, 0, 0};
    P.x3 = x3; P.f16 = f16; P.B = B; P.T = T; P.Cout = Q; P.bias = b2; P.o0 = y;
    if (hl == nullptr) {
      if (int rc = launch_gemm<EPI_HEAD>(maps, P, dim3(ceil_div(T, TM), ceil_div(Q, TN), B), stream)) return rc;
    } else {
      // number of labels the mean runs over (normalize=True, ignore_label=-1), then the GEMM whose
      // epilogue turns each row of logits into its loss term and its gradient planes
      if (hl->tgt_i) {
        count_labels_kernel<<<148, 256, 0, stream>>>(hl->tgt_i, (int64_t)B * T, Q, hl->loss + 1);
        VQW_CHECK_LAUNCH("count_labels_kernel");
      } else {
        set_double_kernel<<<1, 1, 0, stream>>>(hl->loss + 1, (double)B * (double)T);
        VQW_CHECK_LAUNCH("set_double_kernel");
      }
      float* sc = reinterpret_cast<float*>(sv + S.scale);
      plane_scale_kernel<<<1, 1, 0, stream>>>(hl->loss + 1, sc, f16, hl->tgt_i ? 7 : 0);
      VQW_CHECK_LAUNCH("plane_scale_kernel");
      P.plane_s = sc + 2;
      P.Cout = qpad(Q); P.Qv = Q;
      P.tgt_i = hl->tgt_i; P.tgt_f = hl->tgt_f; P.loss = hl->loss;
      P.mol_half = (float)(127.5 / (hl->quantize - 1)); P.mol_lsmin = hl->log_scale_min;
      P.p_hi = reinterpret_cast<__nv_bfloat16*>(sv + S.gy[0]);
      P.p_lo = LOW(reinterpret_cast<__nv_bfloat16*>(sv + S.gy[1]));
      if (int rc = launch_gemm<EPI_HEAD_LOSS>(maps, P, dim3(ceil_div(T, TM), 1, B), stream)) return rc;
    }
  }
  return 0;
}

int head_forward_tc(const vqw_head_desc& d, const float* skip, const float* W1, const float* b1,
                    const float* W2, const float* b2, float* y, void* workspace, void* saved,
                    cudaStream_t stream) {
  VQW_REQUIRE(y != nullptr, "vqw_head_forward: null argument");
  return head_forward_impl(d, skip, W1, b1, W2, b2, y, workspace, saved, nullptr, stream);
}

int head_loss_forward_tc(const vqw_head_desc& d, const float* skip, const float* W1, const float* b1,
                         const float* W2, const float* b2, const int32_t* tgt_i, const float* tgt_f,
                         int quantize, float log_scale_min, double* loss, float* y, void* workspace,
                         void* saved, cudaStream_t stream) {
  HeadLoss hl = {tgt_i, tgt_f, loss, quantize, log_scale_min};
  return head_forward_impl(d, skip, W1, b1, W2, b2, y, workspace, saved, &hl, stream);
}

namespace tc {
// {s/g, g/s} from the upstream gradient g of the scalar loss (device memory, no host sync) and the
// factor s = scale[2] the fused loss multiplied its gradient planes by
__global__ void loss_scale_kernel(const float* __restrict__ g, float* __restrict__ scale) {
  const float v = *g, s = scale[2];
  scale[0] = v != 0.0f ? s / v : 0.0f;
  scale[1] = v / s;
}

}  // namespace tc

// gy == nullptr: the gradient planes were written by head_loss_forward_tc (in `saved`); g_loss is
// the upstream gradient of the scalar loss (device pointer), applied to every fp32 result
int head_backward_tc(const vqw_head_desc& d, const float* gy, const float* g_loss, const float* W1,
                     const float* W2, float* gskip, float* gW1, float* gb1, float* gW2, float* gb2,
                     void* workspace, const void* saved, cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(head_tc_supported(d), "tcgen05 head: unsupported shape");
  VQW_REQUIRE((gy || g_loss) && W1 && W2 && gskip && gW1 && gb1 && gW2 && gb2 && workspace && saved,
              "vqw_head_backward: null argument");
  const bool x3 = vqw_mode_x3(d.mode);
  const int f16 = vqw_mode_f16(d.mode) ? 1 : 0;
  const HeadLayout L = head_layout(d);
  const HeadSaved S = head_saved(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(al((int64_t)(uintptr_t)workspace));
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(al((int64_t)(uintptr_t)saved));
  auto W16 = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto LOW = [&](__nv_bfloat16* p) { return x3 ? p : nullptr; };
  const int B = d.B, T = d.T, Cs = d.Cs, Q = d.Q, Qp = qpad(Q);
  const int64_t NROWS = (int64_t)B * T;
  const int RPB = 256, CS_GRID = (int)((NROWS + RPB - 1) / RPB);
  const float* gscale = nullptr;
  if (f16 && gy != nullptr) {
    if (int rc = make_grad_scale(gy, (int64_t)B * Q * T, nullptr, 0,
                                 reinterpret_cast<float*>(ws + L.scale), stream, &gscale))
      return rc;
  }
  const uint8_t* gyp[2] = {ws + L.gy[0], ws + L.gy[1]};
  if (gy != nullptr) {
    if (int rc = pack_act_launch_ex(gy, W16(L.gy[0]), LOW(W16(L.gy[1])), B, Q, T, Qp, 0, f16, gscale, stream))
      return rc;
  } else {
    gyp[0] = sv + S.gy[0];
    gyp[1] = sv + S.gy[1];
    float* sc = reinterpret_cast<float*>(const_cast<uint8_t*>(sv) + S.scale);
    loss_scale_kernel<<<1, 1, 0, stream>>>(g_loss, sc);
    VQW_CHECK_LAUNCH("loss_scale_kernel");
    gscale = sc;
  }
  pack_mat_kernel<<<148, 256, 0, stream>>>(W2, Q, Cs, 1, W16(L.w2t[0]), LOW(W16(L.w2t[1])), Cs, Qp, f16);
  VQW_CHECK_LAUNCH("pack_mat_kernel(W2T)");
  pack_mat_kernel<<<148, 256, 0, stream>>>(W1, Cs, Cs, 1, W16(L.w1t[0]), LOW(W16(L.w1t[1])), Cs, Cs, f16);
  VQW_CHECK_LAUNCH("pack_mat_kernel(W1T)");
  auto mapk = [&](CUtensorMap* m, const void* hi, const void* lo, uint64_t inner, uint64_t rows,
                  uint64_t batch, uint32_t box_rows) -> int {
    if (int rc = make_map(&m[0], hi, 3, inner, rows, batch, box_rows)) return rc;
    return make_map(&m[1], x3 ? lo : hi, 3, inner, rows, batch, box_rows);
  };
  auto mapmn = [&](CUtensorMap* m, const void* hi, const void* lo, uint64_t C) -> int {
    if (int rc = make_map_mn(&m[0], hi, C, C, T, B)) return rc;
    return make_map_mn(&m[1], x3 ? lo : hi, C, C, T, B);
  };
  {   // g_h1 = (W2^T g_y) * (h1 > 0) -> planes
    Maps maps;
    if (int rc = mapk(&maps.m[0], gyp[0], gyp[1], Qp, T, B, TM)) return rc;
    if (int rc = mapk(&maps.m[2], ws + L.w2t[0], ws + L.w2t[1], Qp, Cs, 1, WROWS())) return rc;
    for (int k = 4; k < NMAPS; ++k) maps.m[k] = maps.m[k % 4];
    GemmParams P = {};
    P.nseg = 1; P.seg[0] = Seg{0, 1, Qp / BK, 0, 0, 0, 0};
    P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T; P.Cout = Cs;
    P.mask_hi = reinterpret_cast<const __nv_bfloat16*>(sv + S.h1[0]);
    P.p_hi = W16(L.gh1[0]); P.p_lo = LOW(W16(L.gh1[1]));
    if (int rc = launch_gemm<EPI_HEAD>(maps, P, dim3(ceil_div(T, TM), Cs / TN, B), stream)) return rc;
  }
  {   // g_skip = (W1^T g_h1) * (skip > 0) -> fp32 (B,Cs,T)
    Maps maps;
    if (int rc = mapk(&maps.m[0], ws + L.gh1[0], ws + L.gh1[1], Cs, T, B, TM)) return rc;
    if (int rc = mapk(&maps.m[2], ws + L.w1t[0], ws + L.w1t[1], Cs, Cs, 1, WROWS())) return rc;
    for (int k = 4; k < NMAPS; ++k) maps.m[k] = maps.m[k % 4];
    GemmParams P = {};
    P.nseg = 1; P.seg[0] = Seg{0, 1, Cs / BK, 0, 0, 0, 0};
    P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T; P.Cout = Cs;
    P.mask_hi = reinterpret_cast<const __nv_bfloat16*>(sv + S.s[0]);
    P.o0 = gskip;
    if (int rc = launch_gemm<EPI_HEAD>(maps, P, dim3(ceil_div(T, TM), Cs / TN, B), stream)) return rc;
  }
  {   // gW2 = g_y (x) h1, gW1 = g_h1 (x) relu(skip): one grouped launch
    Maps maps;
    if (int rc = mapmn(&maps.m[0], gyp[0], gyp[1], Qp)) return rc;
    if (int rc = mapmn(&maps.m[2], sv + S.h1[0], sv + S.h1[1], Cs)) return rc;
    if (int rc = mapmn(&maps.m[4], ws + L.gh1[0], ws + L.gh1[1], Cs)) return rc;
    if (int rc = mapmn(&maps.m[6], sv + S.s[0], sv + S.s[1], Cs)) return rc;
    for (int k = 8; k < NMAPS; ++k) maps.m[k] = maps.m[k % 8];
    GemmParams P = {};
    P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T;
    P.slabs_per_item = ceil_div(T, BK);
    P.chunks_per_b = 1;
    int nj = 0;
    // m0 fastest: consecutive jobs = the halves of a 256-row tile (CTA pair); Q < 256 has an odd tile
    const bool wg_pair = (ceil_div(Q, TM) & 1) == 0 && (ceil_div(Cs, TM) & 1) == 0;
    for (int n0 = 0; n0 < Cs; n0 += TN)
      for (int m0 = 0; m0 < Q; m0 += TM) {
        VQW_REQUIRE(nj < MAX_JOBS, "tcgen05 head: too many weight-gradient tiles");
        P.jobs[nj++] = Job{0, 1, m0, n0, 0, Q, Cs, gW2, Cs, 1};
      }
    for (int n0 = 0; n0 < Cs; n0 += TN)
      for (int m0 = 0; m0 < Cs; m0 += TM) {
        VQW_REQUIRE(nj < MAX_JOBS, "tcgen05 head: too many weight-gradient tiles");
        P.jobs[nj++] = Job{2, 3, m0, n0, 0, Cs, Cs, gW1, Cs, 1};
      }
    P.njobs = nj;
    if (int rc = launch_gemm<EPI_WGRAD>(maps, P, dim3(nj, 1, B), stream, wg_pair)) return rc;
  }
  colsum_planes_kernel<<<CS_GRID, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(gyp[0]),
      x3 ? reinterpret_cast<const __nv_bfloat16*>(gyp[1]) : nullptr, gb2, nullptr, Qp, NROWS, RPB, Q, f16, gscale);
  VQW_CHECK_LAUNCH("colsum_planes_kernel(gy)");
  colsum_planes_kernel<<<CS_GRID, 256, 0, stream>>>(W16(L.gh1[0]), LOW(W16(L.gh1[1])), gb1, nullptr,
                                                   Cs, NROWS, RPB, Cs, f16, gscale);
  VQW_CHECK_LAUNCH("colsum_planes_kernel(gh1)");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Embed weight gradient on the tensor cores (modules.py:151-152 differentiated):
//   gW[c, q, j] = sum_{b,t} g[b, c, t] * onehot(q[b, t - 1 + j])[q]
// i.e. two K=time GEMMs of the gradient planes against a one-hot plane (exact in bf16, so it
// has no lo plane), through the grouped weight-gradient kernel; gb = column sums of g.
// ---------------------------------------------------------------------------------------
namespace tc {
__global__ void __launch_bounds__(256)
onehot_planes_kernel(const int32_t* __restrict__ q, __nv_bfloat16* __restrict__ oh, int64_t n, int Qp,
                     int f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int k = q[i];
    if (k >= 0 && k < Qp)   // 1.0 as bf16 or fp16 bits
      oh[i * Qp + k] = __ushort_as_bfloat16((unsigned short)(f16 ? 0x3C00 : 0x3F80));
  }
}
}  // namespace tc

bool embed_bwd_tc_supported(int B, int T, int Cr, int Q) {
  return B >= 1 && B <= 65535 && T >= tc::TM && T % 8 == 0 && Cr % 64 == 0 && Q >= 1;
}
int64_t embed_bwd_tc_workspace(int B, int T, int Cr, int Q) {
  const int64_t N = (int64_t)B * T;
  return 2 * al(N * Cr * 2) + al(N * pad256(Q) * 2) + 1024 + 2048;
}
int embed_backward_tc(const int32_t* q, const float* g, float* gW, float* gb, int B, int T, int Cr,
                      int Q, int mode, void* workspace, cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(embed_bwd_tc_supported(B, T, Cr, Q), "tcgen05 embed backward: unsupported shape");
  VQW_REQUIRE(q && g && gW && workspace, "vqw_embed_gather_backward_tc: null pointer");
  const bool x3 = vqw_mode_x3(mode);
  const int f16 = vqw_mode_f16(mode) ? 1 : 0;
  const int Qp = pad256(Q);
  const int64_t N = (int64_t)B * T;
  uint8_t* ws = reinterpret_cast<uint8_t*>(al((int64_t)(uintptr_t)workspace));
  __nv_bfloat16* g_hi = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* g_lo = reinterpret_cast<__nv_bfloat16*>(ws + al(N * Cr * 2));
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(ws + 2 * al(N * Cr * 2));
  const float* gscale = nullptr;
  if (f16) {
    float* scale_mem = reinterpret_cast<float*>(ws + 2 * al(N * Cr * 2) + al(N * Qp * 2));
    if (int rc = make_grad_scale(g, N * Cr, nullptr, 0, scale_mem, stream, &gscale)) return rc;
  }
  if (int rc = pack_act_launch(g, g_hi, x3 ? g_lo : nullptr, B, Cr, T, f16, gscale, stream)) return rc;
  VQW_CHECK_CUDA(cudaMemsetAsync(oh, 0, (size_t)N * Qp * 2, stream));
  onehot_planes_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(q, oh, N, Qp, f16);
  VQW_CHECK_LAUNCH("onehot_planes_kernel");
  Maps maps;
  if (int rc = make_map_mn(&maps.m[0], g_hi, Cr, Cr, T, B)) return rc;
  if (int rc = make_map_mn(&maps.m[1], x3 ? g_lo : g_hi, Cr, Cr, T, B)) return rc;
  if (int rc = make_map_mn(&maps.m[2], oh, Qp, Qp, T, B)) return rc;
  maps.m[3] = maps.m[2];
  for (int k = 4; k < NMAPS; ++k) maps.m[k] = maps.m[k % 4];
  GemmParams P = {};
  P.x3 = x3; P.f16 = f16; P.scale = gscale; P.B = B; P.T = T;
  P.b_exact = 1;
  P.slabs_per_item = ceil_div(T, BK);
  P.chunks_per_b = 1;
  int nj = 0;
  const bool wg_pair = (ceil_div(Cr, TM) & 1) == 0;
  for (int j = 0; j < 2; ++j)
    for (int n0 = 0; n0 < Q; n0 += TN)
      for (int m0 = 0; m0 < Cr; m0 += TM) {
        VQW_REQUIRE(nj < MAX_JOBS, "tcgen05 embed backward: too many tiles");
        // tap j reads the index at t - 1 + j  (j = 0: previous sample, j = 1: current sample)
        P.jobs[nj++] = Job{0, 1, m0, n0, j - 1, Cr, Q, gW + j, (long long)Q * 2, 2};
      }
  P.njobs = nj;
  if (int rc = launch_gemm<EPI_WGRAD>(maps, P, dim3(nj, 1, B), stream, wg_pair)) return rc;
  if (gb) {
    const int RPB = 256;
    colsum_planes_kernel<<<(int)((N + RPB - 1) / RPB), 256, 0, stream>>>(g_hi, x3 ? g_lo : nullptr, gb,
                                                                        nullptr, Cr, N, RPB, Cr, f16,
                                                                        gscale);
    VQW_CHECK_LAUNCH("colsum_planes_kernel(embed)");
  }
  return 0;
}

}  // namespace vqw
