// Fused WaveNet residual block on the 5th-generation tensor cores (sm_100a: tcgen05 + TMEM + TMA).
//
// Replaces ResidualBlock.__call__ (modules.py:30-56) and ResidualNet's skip accumulation
// (modules.py:92-95) for the 512-channel training config, where the block is two dense
// contractions per time step (K = fs*Cr + Cc = 1728 then K = Cd/2 = 256): 2.16 MFLOP per audio
// sample against ~13 KB of HBM traffic, i.e. tensor-core bound.
//
// Precision.  The reference computes in fp32.  tcgen05 has no fp32 operand type and
// kind::tf32 truncates fp32 operands (a -1e-3 systematic bias), so operands are SPLIT into two
// 16-bit planes (hi = r(v), lo = r(v - hi): bf16 pairs = 16 significant bits, "bf16x3"; IEEE fp16
// pairs = 22 bits, "fp16x3") and every product is issued as three MMAs  hi*hi + lo*hi + hi*lo
// accumulated in fp32 in TMEM (parity modes).  VQW_MODE_BF16 / VQW_MODE_FP16 issue only hi*hi
// (throughput modes, not parity).
//
// Data layout.  Between blocks activations live in a packed (B, T, C) layout, channel
// contiguous, one bf16 plane for hi and one for lo: the K axis (channels) of both MMA operands
// is then contiguous ("K-major"), and the dilated causal taps are plain row shifts of a TMA
// box -- rows with t < 0 are out of bounds and TMA zero-fills them, which IS the causal pad.
// Weights are packed once per step as [out-channel rows][K] bf16 planes.
//
// One CTA = one (batch item, 128 time steps) tile, 192 threads:
//   warp 4 lane 0 : TMA producer, 4-stage mbarrier ring of {A_hi, A_lo, B_hi, B_lo} K=32 slabs
//   warp 5 lane 0 : tcgen05.mma issuer (M=128 time rows x N=256 channels, fp32 accum in TMEM)
//   warps 0-3     : epilogue, one thread per time row (TMEM lane)
// TMEM (512 columns): [0,256) accumulator, [256,384) z hi plane, [384,512) z lo plane.
// Phases per tile:   H_a (tanh 0..127 | sigmoid 0..127) -> gate -> z[0:128] into TMEM
//                    H_b (tanh 128..255 | sigmoid 128..255) -> gate -> z[128:256] into TMEM
//                    O_0, O_1 = Wr z (A operand read straight from TMEM) -> + bias + x -> residual
//                    O_2 = Ws z -> skip (+)=
// so the gated activation never leaves the SM and the block is one kernel.
#include "tc_common.cuh"
#include <stdlib.h>

namespace vqw {
namespace tc {

constexpr int STAGES = 4;
constexpr int FWD_EPI_WARPS = 16;                       // 4 warps per TMEM lane quadrant
constexpr int FWD_THREADS = (FWD_EPI_WARPS + 2) * 32;   // + TMA producer warp + MMA warp
constexpr int W_TMA = FWD_EPI_WARPS, W_MMA = FWD_EPI_WARPS + 1;
constexpr int W_ST = FWD_EPI_WARPS + 2;                 // pair kernel: staging-tile store / addend warp
constexpr int PAIR_THREADS = FWD_THREADS + 32;
constexpr int CD = 512, CH = 256, HALF = 128;             // dilated channels handled by this kernel

struct Params {
  int B, T, Cr, Cs, Cc, fs, dilation;
  int x3;               // 1: bf16x3 (three MMAs per product), 0: single pass over the hi planes
  int f16;              // planes hold IEEE fp16 instead of bf16 (VQW_MODE_FP16, single pass)
  int xlo;              // the residual stream keeps its lo plane (residual-add operand): x3 or fp16
  int skip_accumulate;
  int write_residual;   // 0 for the last block: O_0/O_1 are skipped entirely
  int pf_dist;          // L2 prefetch distance (K slabs) of the activation operand in H_a, 0 = off
  int stage_res;        // residual chunks go through the shared-memory tiles + TMA (needs res_hi)
  int stage_gate;       // pair kernel: sigmoid / z of the second gate phase leave as TMA stores
  const __nv_bfloat16* xp_hi;   // packed (B,T,Cr) planes of the block input (residual-add operand)
  const __nv_bfloat16* xp_lo;
  const float* gbias;   // (B,512): conv_b + cond_b + W_p[:, Cl:] . global condition of the item
  const float* res_b;  const float* skip_b;
  float* res_f32;       // (B,Cr,T) fp32 or null (saved block input of the next block / API output)
  __nv_bfloat16* res_hi;  // packed (B,T,Cr) planes for the next block (null for the last block)
  __nv_bfloat16* res_lo;
  float* skip;          // (B,Cs,T) fp32 in/out
  // saved for the backward, TIME-major (B,T,Ch) fp32 or null (thread = time row in both directions).
  // bf16x3 keeps only the sigmoid: z = tanh * sigmoid is saved anyway as hi/lo planes (16
  // significant bits), so the backward recovers tanh = z / sigmoid and the forward writes a third
  // fewer bytes in its gate phases (which are store bound).  The single-pass modes keep both.
  float* gate_tanh;
  float* gate_sig;
  __nv_bfloat16* zp_hi; // packed (B,T,Ch) planes of z = tanh*sigmoid, saved for the backward (or null)
  __nv_bfloat16* zp_lo;
  long long* dbg;       // optional phase timestamps of one CTA (VQW_TC_TIMELINE=1)
  int dbg_x, dbg_y;
};

// ------------------------------------------------------------------ the kernel -------------
// TMEM plan (512 columns, fp32 accumulators; one CTA per SM):
//   [  0,256) accumulator of H_a   (tanh 0..127 | sigmoid 0..127)
//   [256,512) accumulator of H_b   (tanh 128..255 | sigmoid 128..255)
// so the MMAs of H_b run while the 16 epilogue warps gate H_a.  The gate writes z IN PLACE: the 16
// tanh columns of a chunk are overwritten by the 8 + 8 columns of its bf16 hi / lo z pairs (each
// thread only ever touches the columns it has just read), i.e. z_a ends up in [0,128), z_b in
// [256,384), interleaved {hi(16 ch), lo(16 ch)} per K = 16 step -- exactly the A-operand tiles of
// the second contraction.  The two drained sigmoid halves [128,256) and [384,512) become the
// ping-pong accumulators of the output phase, which runs as N = 128 chunks of [Wr ; Ws] (only the
// weights stream, z stays in TMEM): the epilogue of chunk c overlaps the MMAs of chunk c + 1.
// Exposed epilogues per tile: the gate of H_b and the last output chunk (round 1: all five).
constexpr int ON = 128;                                   // output chunk = UMMA N of the 2nd contraction
constexpr int OB_PLANE = ON * BK * 2;                     // 8 KB per weight plane and K slab
constexpr int OACC0 = 128, OACC1 = 384;                   // ping-pong output accumulators
constexpr uint32_t IDESC_ON = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ON >> 3) << 17) |
                              ((uint32_t)(TM >> 4) << 24);
constexpr int NBAR = 2 * STAGES + 10;                     // ring + hfull[2] zready[2] ofull[2] oempty[2] afull[2]
// Residual epilogue staging.  In the output phase a ring stage only holds two 8 KB weight planes
// ([16 K, 32 K) of its 48 KB), so its activation area [0, 16 K) and its tail [32 K, 48 K) are free:
// eight 16 KB regions = two buffers of four [128 rows x 64 channels] bf16 tiles (128-byte swizzle).
// The addend x of a residual chunk is TMA-loaded into a buffer, every thread updates ITS row in
// place (x + Wr z + br, re-split into hi / lo), and one thread TMA-stores the tiles to the next
// block's input planes.  Round 2 measured why: a thread owns one time row, so every per-thread
// 32-byte global access is its own LSU wavefront (4096 per chunk, ~8 k cycles); through TMA the
// epilogue issues none.
constexpr int REGION_BYTES = 128 * 128;

// X3: three MMAs per product over hi/lo planes; F16: the planes hold IEEE fp16 (single pass)
template <int X3, int F16>
__global__ void __launch_bounds__(FWD_THREADS, 1)
resblock_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi,
                   const __grid_constant__ CUtensorMap map_x_lo,
                   const __grid_constant__ CUtensorMap map_c_hi,
                   const __grid_constant__ CUtensorMap map_c_lo,
                   const __grid_constant__ CUtensorMap map_w1_hi,
                   const __grid_constant__ CUtensorMap map_w1_lo,
                   const __grid_constant__ CUtensorMap map_w2_hi,
                   const __grid_constant__ CUtensorMap map_w2_lo,
                   const __grid_constant__ CUtensorMap map_xa_hi,   // addend tiles of x (staged epilogue)
                   const __grid_constant__ CUtensorMap map_xa_lo,
                   const __grid_constant__ CUtensorMap map_r_hi,    // residual output planes
                   const __grid_constant__ CUtensorMap map_r_lo, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the dynamic window is only guaranteed 16-byte aligned: round up to the swizzle period
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* b1s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);   // [512] conv_b + cond_b
  float* brs = b1s + CD;                                                // [Cr]
  float* bss = brs + P.Cr;                                              // [Cs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bss + P.Cs);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t hfull0 = smem_u32(bars + 2 * STAGES), zready0 = hfull0 + 16;
  const uint32_t ofull0 = hfull0 + 32, oempty0 = hfull0 + 48, afull0 = hfull0 + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * TM;
  constexpr int XLO = X3 | F16;   // the residual stream keeps its lo plane (residual-add operand)
  const int nplanes = X3 ? 2 : 1;
  const bool rec_cta = P.dbg != nullptr && blockIdx.x == P.dbg_x && blockIdx.y == P.dbg_y;
  if (rec_cta && threadIdx.x == 0) P.dbg[40] = clock64();
  const int chunks_per_tap = P.Cr / BK;
  const int nk1 = P.fs * chunks_per_tap + P.Cc / BK;     // K slabs of the first contraction
  const int nk2 = CH / BK;                               // K slabs of the second contraction
  const int n_res = P.Cr / ON;                           // output chunks: residual rows, then skip rows
  const int o_begin = P.write_residual ? 0 : n_res;
  const int o_end = n_res + P.Cs / ON;

  if (warp == W_TMA && lane == 0) {
    prefetch_tmap(&map_x_hi); prefetch_tmap(&map_c_hi); prefetch_tmap(&map_w1_hi);
    prefetch_tmap(&map_w2_hi);
    if (X3) {
      prefetch_tmap(&map_x_lo); prefetch_tmap(&map_c_lo); prefetch_tmap(&map_w1_lo);
      prefetch_tmap(&map_w2_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(hfull0 + 8 * g, 1);
      mbar_init(zready0 + 8 * g, FWD_EPI_WARPS * 32);
      mbar_init(ofull0 + 8 * g, 1);
      mbar_init(oempty0 + 8 * g, FWD_EPI_WARPS * 32);
      mbar_init(afull0 + 8 * g, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < FWD_EPI_WARPS) {
    for (int i = threadIdx.x; i < CD; i += FWD_EPI_WARPS * 32) b1s[i] = P.gbias[(int64_t)b * CD + i];
    for (int i = threadIdx.x; i < P.Cr; i += FWD_EPI_WARPS * 32) brs[i] = P.res_b[i];
    for (int i = threadIdx.x; i < P.Cs; i += FWD_EPI_WARPS * 32) bss[i] = P.skip_b[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (rec_cta && threadIdx.x == 0) P.dbg[41] = clock64();

  if (warp == W_TMA) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      // activation slab i of the first contraction: tap slabs of x (row shift = causal delay,
      // negative rows are zero-filled by TMA), then the condition slabs
      auto a_src = [&](int i, const CUtensorMap*& mh, const CUtensorMap*& ml, int& c0, int& tt) {
        const int tap = i / chunks_per_tap;
        if (tap < P.fs) {
          mh = &map_x_hi; ml = &map_x_lo;
          c0 = (i - tap * chunks_per_tap) * BK;
          tt = t0 - P.dilation * (P.fs - 1 - tap);
        } else {
          mh = &map_c_hi; ml = &map_c_lo;
          c0 = (i - P.fs * chunks_per_tap) * BK;
          tt = t0;
        }
      };
      auto a_prefetch = [&](int i) {
        const CUtensorMap *mh, *ml;
        int c0, tt;
        a_src(i, mh, ml, c0, tt);
        tma_prefetch_3d(mh, c0, tt, b);
        if (X3) tma_prefetch_3d(ml, c0, tt, b);
      };
      // H_a reads its activation slabs for the first time (DRAM); H_b re-reads them from L2.
      // An L2 prefetch runs pf_dist slabs ahead of the ring so that the DRAM latency is not paid
      // once per ring round trip (round 1: H_a 8-22 k cycles slower than H_b).
      const int pf = P.pf_dist < nk1 ? P.pf_dist : nk1;
      for (int i = 0; i < pf; ++i) a_prefetch(i);
      for (int gp = 0; gp < 2; ++gp) {
        for (int i = 0; i < nk1; ++i) {
          if (gp == 0 && pf > 0 && i + pf < nk1) a_prefetch(i + pf);
          mbar_wait(empty0 + 8 * stage, ph ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          const uint32_t sa = base + stage * STAGE_BYTES;
          mbar_expect_tx(fb, nplanes * (A_PLANE + B_PLANE));
          const CUtensorMap *mh, *ml;
          int c0, tt;
          a_src(i, mh, ml, c0, tt);
          tma_load_3d(sa, mh, fb, c0, tt, b);
          if (X3) tma_load_3d(sa + A_PLANE, ml, fb, c0, tt, b);
          tma_load_2d(sa + 2 * A_PLANE, &map_w1_hi, fb, i * BK, gp * TN);
          if (X3) tma_load_2d(sa + 2 * A_PLANE + B_PLANE, &map_w1_lo, fb, i * BK, gp * TN);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
      for (int oc = o_begin; oc < o_end; ++oc) {
        for (int i = 0; i < nk2; ++i) {
          mbar_wait(empty0 + 8 * stage, ph ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          const uint32_t sa = base + stage * STAGE_BYTES;
          mbar_expect_tx(fb, nplanes * OB_PLANE);
          tma_load_2d(sa + 2 * A_PLANE, &map_w2_hi, fb, i * BK, oc * ON);
          if (X3) tma_load_2d(sa + 2 * A_PLANE + OB_PLANE, &map_w2_lo, fb, i * BK, oc * ON);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      const bool rec = rec_cta;
      const uint32_t idesc = idesc_for(IDESC, F16);
      const uint32_t idesc_o = idesc_for(IDESC_ON, F16);
      for (int gp = 0; gp < 2; ++gp) {
        const uint32_t acc = tmem_base + 256 * gp;
        if (rec) P.dbg[2 * gp] = clock64();
        for (int i = 0; i < nk1; ++i) {
          mbar_wait(full0 + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            const uint64_t a_hi = smem_desc_sw64(sa + ks * UK * 2);
            const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            mma_ss(acc, a_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
            if (X3) {
              const uint64_t a_lo = smem_desc_sw64(sa + A_PLANE + ks * UK * 2);
              const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + B_PLANE + ks * UK * 2);
              mma_ss(acc, a_lo, b_hi, idesc, 1u);
              mma_ss(acc, a_hi, b_lo, idesc, 1u);
            }
          }
          tc_commit(empty0 + 8 * stage);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        tc_commit(hfull0 + 8 * gp);
        if (rec) P.dbg[2 * gp + 1] = clock64();
      }
      // both halves of z are in TMEM (and both sigmoid halves are drained) from here on
      mbar_wait(zready0, 0);
      mbar_wait(zready0 + 8, 0);
      tc_fence_after();
      for (int oc = o_begin, j = 0; oc < o_end; ++oc, ++j) {
        const int buf = j & 1, use = j >> 1;
        if (use > 0) {
          mbar_wait(oempty0 + 8 * buf, (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (buf ? OACC1 : OACC0);
        if (rec && j < 6) P.dbg[4 + 2 * j] = clock64();
        for (int i = 0; i < nk2; ++i) {
          mbar_wait(full0 + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            // z channels [32 i + 16 ks, +16): hi pairs in 8 columns, lo pairs in the next 8
            const int kstep = 2 * i + ks;
            const uint32_t z_hi = tmem_base + (kstep < 8 ? 0 : 256) + 16 * (kstep & 7);
            const uint32_t z_lo = z_hi + 8;
            const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            mma_ts(acc, z_hi, b_hi, idesc_o, (i | ks) ? 1u : 0u);
            if (X3) {
              const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + OB_PLANE + ks * UK * 2);
              mma_ts(acc, z_lo, b_hi, idesc_o, 1u);
              mma_ts(acc, z_hi, b_lo, idesc_o, 1u);
            }
          }
          tc_commit(empty0 + 8 * stage);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        tc_commit(ofull0 + 8 * buf);
        if (rec && j < 6) P.dbg[4 + 2 * j + 1] = clock64();
      }
    }
  } else {
    // =============================== epilogue (warps 0-15) ======================
    // warp e: TMEM lane quadrant e%4 (hardware rule: a warp reaches lanes 32*(warp%4)..+31),
    // column group e/4 -- the 16-column chunks of a phase are dealt round-robin to the 4 groups,
    // so every SM sub-partition has 4 resident epilogue warps to hide latencies with.
    const int quad = warp & 3, grp = warp >> 2;
    constexpr int NG = FWD_EPI_WARPS / 4;
    const int row = quad * 32 + lane;
    const int t = t0 + row;
    const bool t_ok = t < P.T;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool rec = rec_cta && threadIdx.x == 0;
    const bool save_gates = P.gate_sig != nullptr && t_ok;
    const bool leader = threadIdx.x == 0;
    // tile h (0 = channels [0,64), 1 = [64,128)) of plane pl (0 = hi, 1 = lo) of staging buffer bf
    auto tile_addr = [&](int bf, int pl, int h) -> uint32_t {
      return base + (uint32_t)(2 * bf + pl) * STAGE_BYTES + (h ? (uint32_t)(2 * A_PLANE + 2 * OB_PLANE) : 0u);
    };
    // TMA-load the addend x[t0 .. t0+127][128 oc .. +127] (hi, lo) of residual chunk oc
    auto issue_addend = [&](int bf, int oc) {
      const uint32_t bar = afull0 + 8 * bf;
      mbar_expect_tx(bar, (XLO ? 4 : 2) * REGION_BYTES);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(tile_addr(bf, 0, h), &map_xa_hi, bar, oc * ON + 64 * h, t0, b);
        if (XLO) tma_load_3d(tile_addr(bf, 1, h), &map_xa_lo, bar, oc * ON + 64 * h, t0, b);
      }
    };
    // ---- gate phases: z = tanh(h_t) * sigmoid(h_s), written over the tanh columns ----
    for (int gp = 0; gp < 2; ++gp) {
      mbar_wait(hfull0 + 8 * gp, 0);
      tc_fence_after();
      if (rec) P.dbg[16 + 2 * gp] = clock64();
      if (gp == 1 && leader && P.stage_res) {
        // every MMA of the first contraction is done: the activation areas and tails of the ring
        // are free from here on.  The addends of the first two residual chunks land during E_b.
        for (int j = 0; j < 2 && o_begin + j < n_res; ++j) issue_addend(j, o_begin + j);
      }
      const uint32_t accb = lane_base + 256 * gp;
#pragma unroll 1
      for (int q = grp; q < HALF / 16; q += NG) {
        uint32_t ar[16], gr[16];
        tmem_ld16_issue(accb + 16 * q, ar);
        tmem_ld16_issue(accb + HALF + 16 * q, gr);
        tmem_ld_wait(ar, gr);
        uint32_t zh[8], zl[8];
        const int ch0 = gp * HALF + 16 * q;
        float bt[16], bs[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          *reinterpret_cast<float4*>(bt + 4 * i) = *reinterpret_cast<const float4*>(b1s + ch0 + 4 * i);
          *reinterpret_cast<float4*>(bs + 4 * i) = *reinterpret_cast<const float4*>(b1s + CH + ch0 + 4 * i);
        }
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float z2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float th, sg;
            gate_pair(__uint_as_float(ar[i + u]) + bt[i + u], __uint_as_float(gr[i + u]) + bs[i + u],
                      th, sg);
            z2[u] = th * sg;
            ar[i + u] = __float_as_uint(th);     // kept for the backward (gate derivative)
            gr[i + u] = __float_as_uint(sg);
          }
          if (X3) split_pair_f(z2[0], z2[1], zh[i >> 1], zl[i >> 1], F16);
          else zh[i >> 1] = pack_pair_f(z2[0], z2[1], F16);
        }
        if (save_gates) {   // time-major (B,T,Ch) fp32: 64 contiguous bytes per thread and array
          const int64_t goff = ((int64_t)b * P.T + t) * CH + ch0;
          if (!X3) {
            st256(P.gate_tanh + goff, ar);
            st256(P.gate_tanh + goff + 8, ar + 8);
          }
          st256(P.gate_sig + goff, gr);
          st256(P.gate_sig + goff + 8, gr + 8);
        }
        tmem_st8(accb + 16 * q, zh);
        if (X3) tmem_st8(accb + 16 * q + 8, zl);
        if (P.zp_hi != nullptr && t_ok) {
          const int64_t zoff = ((int64_t)b * P.T + t) * CH + ch0;
          st256(P.zp_hi + zoff, zh);
          if (X3) st256(P.zp_lo + zoff, zl);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(zready0 + 8 * gp);
      if (rec) P.dbg[16 + 2 * gp + 1] = clock64();
    }
    // ---- output chunks: residual rows then skip rows, 128 per chunk ----
    // residual = Wr z + br + x with x read back from the packed hi/lo planes (x = hi + lo to
    // 2^-17: two 32-byte loads per 16 channels instead of 16 strided fp32 loads); the running
    // skip sum is fp32 (B,Cs,T): lanes are consecutive t, so every access is a coalesced 128-byte
    // row segment.  Operands of the NEXT 16-column piece are fetched before the current one is
    // processed.
    for (int oc = o_begin, j = 0; oc < o_end; ++oc, ++j) {
      const int buf = j & 1, use = j >> 1;
      const bool is_res = oc < n_res;
      const int cbase = (is_res ? oc : oc - n_res) * ON;
      const uint32_t accb = lane_base + (buf ? OACC1 : OACC0);
      if (is_res && P.stage_res) {
        // ---------- staged residual chunk: addend and result live in shared-memory tiles ----------
        if (leader && j >= 1 && oc + 1 < n_res) {
          // buffer (j+1)&1 was the source of chunk j-1's stores: wait until they have read it,
          // then fetch the addend of chunk j+1 into it
          tma_store_wait_read();
          issue_addend((j + 1) & 1, oc + 1);
        }
        mbar_wait(afull0 + 8 * buf, use & 1);
        mbar_wait(ofull0 + 8 * buf, use & 1);
        tc_fence_after();
        if (rec && j < 6) P.dbg[20 + 2 * j] = clock64();
        const uint32_t rsw = (uint32_t)(row & 7);
#pragma unroll 1
        for (int q = grp; q < ON / 16; q += NG) {
          float o[16];
          tmem_ld16(accb + 16 * q, o);
          const int ch0 = cbase + 16 * q;
          // 16 channels of this thread's row: two 16-byte chunks (k0, k0+1) of a 128-byte tile row,
          // XOR-swizzled with the row index (CU_TENSOR_MAP_SWIZZLE_128B) -> conflict-free
          const int h = (16 * q) >> 6, k0 = ((16 * q) & 63) >> 3;
          const uint32_t rowb = (uint32_t)row * 128u;
          const uint32_t a0 = tile_addr(buf, 0, h) + rowb + (((uint32_t)k0 ^ rsw) << 4);
          const uint32_t a1 = tile_addr(buf, 0, h) + rowb + (((uint32_t)(k0 + 1) ^ rsw) << 4);
          const uint32_t l0 = a0 + (uint32_t)STAGE_BYTES, l1 = a1 + (uint32_t)STAGE_BYTES;   // lo tile: next stage
          const uint4 h0 = lds128(a0), h1 = lds128(a1);
          uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
          if (XLO) { w0 = lds128(l0); w1 = lds128(l1); }
          const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          const uint32_t lw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float bb[16];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(bb + 4 * i) = *reinterpret_cast<const float4*>(brs + ch0 + 4 * i);
          uint32_t rh[8], rl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0, v1;
            unpack_pair_f(hw[i], F16, v0, v1);
            if (XLO) {
              float e0, e1;
              unpack_pair_f(lw[i], F16, e0, e1);
              v0 += e0;
              v1 += e1;
            }
            v0 += o[2 * i] + bb[2 * i];
            v1 += o[2 * i + 1] + bb[2 * i + 1];
            if (P.res_f32 != nullptr && t_ok) {
              __stcs(P.res_f32 + ((int64_t)b * P.Cr + ch0 + 2 * i) * P.T + t, v0);
              __stcs(P.res_f32 + ((int64_t)b * P.Cr + ch0 + 2 * i + 1) * P.T + t, v1);
            }
            if (XLO) split_pair_sat_f(v0, v1, rh[i], rl[i], F16);
            else rh[i] = pack_pair_f(v0, v1, F16);
          }
          sts128(a0, rh[0], rh[1], rh[2], rh[3]);
          sts128(a1, rh[4], rh[5], rh[6], rh[7]);
          if (XLO) {
            sts128(l0, rl[0], rl[1], rl[2], rl[3]);
            sts128(l1, rl[4], rl[5], rl[6], rl[7]);
          }
        }
        tc_fence_before();
        mbar_arrive(oempty0 + 8 * buf);          // the accumulator is drained
        fence_async_smem();                      // my tile writes -> visible to the TMA store
        asm volatile("bar.sync 1, %0;" ::"n"(FWD_EPI_WARPS * 32) : "memory");
        if (leader) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_store_3d(&map_r_hi, tile_addr(buf, 0, h), oc * ON + 64 * h, t0, b);
            if (XLO) tma_store_3d(&map_r_lo, tile_addr(buf, 1, h), oc * ON + 64 * h, t0, b);
          }
          tma_store_commit();
        }
        if (rec && j < 6) P.dbg[20 + 2 * j + 1] = clock64();
        continue;
      }
      float addf[16];
      uint32_t hw[8], lw[8];
      auto fetch = [&](int q) {
        const int ch0 = cbase + 16 * q;
        if (is_res) {
          const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
          if (t_ok) {
            ld256(P.xp_hi + poff, hw);
            if (XLO) ld256(P.xp_lo + poff, lw);
          }
        } else if (P.skip_accumulate && t_ok) {
          const float* sp = P.skip + ((int64_t)b * P.Cs + ch0) * P.T + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) addf[i] = __ldcs(sp + (int64_t)i * P.T);
        }
      };
      fetch(grp);
      mbar_wait(ofull0 + 8 * buf, use & 1);
      tc_fence_after();
      if (rec && j < 6) P.dbg[20 + 2 * j] = clock64();
#pragma unroll 1
      for (int q = grp; q < ON / 16; q += NG) {
        float o[16], add[16];
        tmem_ld16(accb + 16 * q, o);
        if (is_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0, v1;
            unpack_pair_f(hw[i], F16, v0, v1);
            if (XLO) {
              float l0, l1;
              unpack_pair_f(lw[i], F16, l0, l1);
              v0 += l0;
              v1 += l1;
            }
            add[2 * i] = t_ok ? v0 : 0.0f;
            add[2 * i + 1] = t_ok ? v1 : 0.0f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) add[i] = (P.skip_accumulate && t_ok) ? addf[i] : 0.0f;
        }
        if (q + NG < ON / 16) fetch(q + NG);
        const int ch0 = cbase + 16 * q;
        if (is_res) {
          uint32_t rh[8], rl[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float v2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ch = ch0 + i + u;
              const float v = o[i + u] + brs[ch] + add[i + u];
              if (t_ok && P.res_f32 != nullptr)
                __stcs(P.res_f32 + ((int64_t)b * P.Cr + ch) * P.T + t, v);
              v2[u] = v;
            }
            if (XLO) split_pair_sat_f(v2[0], v2[1], rh[i >> 1], rl[i >> 1], F16);
            else rh[i >> 1] = pack_pair_f(v2[0], v2[1], F16);
          }
          if (t_ok && P.res_hi != nullptr) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
            st256(P.res_hi + poff, rh);
            if (XLO) st256(P.res_lo + poff, rl);
          }
        } else if (t_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ch = ch0 + i;
            P.skip[((int64_t)b * P.Cs + ch) * P.T + t] = o[i] + bss[ch] + add[i];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(oempty0 + 8 * buf);
      if (rec && j < 6) P.dbg[20 + 2 * j + 1] = clock64();
    }
    if (leader && P.stage_res) tma_store_wait_all();   // the tiles must outlive their stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
    if (rec_cta && lane == 0) P.dbg[42] = clock64();
  }
}

// ------------------------------------------------------------------ the kernel on a CTA pair
// Same schedule on a cluster of two CTAs = two consecutive time tiles of one item.  Every MMA is a
// cta_group::2 instruction with M = 256 (128 rows per SM); each SM stages only HALF of every
// weight slab (the tensor core reads the other half from the peer's shared memory), so a K slab
// costs an SM 32 KB of L2 -> SM traffic instead of 48 KB.  Why: round 2 measured the H phases
// fill bound -- 48 KB per 768 MMA cycles = 62.5 B/clk per SM is asked for, ~52 B/clk is what the
// L2 delivers when all 148 SMs stream at once (H phase 50-52 k cycles against 41.5 k of MMA).
// Protocol (validated by round 1's experimental kernel): the TMA copies of both CTAs complete on
// the LEADER's "full" barrier; one thread of the leader issues the MMAs; "empty", "H full" and
// "O full" are tcgen05.commit multicasts to both CTAs; one elected lane per epilogue warp of both
// CTAs arrives on the leader's "z ready" / "O empty" barriers.
constexpr int PST = 4;                                    // ring stages
constexpr int PB_PLANE = (TN / 2) * BK * 2;               // 8 KB: this CTA's 128 rows of a W1 slab
constexpr int POB_PLANE = (ON / 2) * BK * 2;              // 4 KB: this CTA's 64 rows of a W2 slab
constexpr int PSTAGE = 2 * A_PLANE + 2 * PB_PLANE;        // 32 KB
constexpr uint32_t IDESC_P = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) |
                             ((uint32_t)((2 * TM) >> 4) << 24);
constexpr uint32_t IDESC_PON = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ON >> 3) << 17) |
                               ((uint32_t)((2 * TM) >> 4) << 24);

template <int X3, int F16>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
resblock_tc_pair_kernel(const __grid_constant__ CUtensorMap map_x_hi,
                   const __grid_constant__ CUtensorMap map_x_lo,
                   const __grid_constant__ CUtensorMap map_c_hi,
                   const __grid_constant__ CUtensorMap map_c_lo,
                   const __grid_constant__ CUtensorMap map_w1_hi,
                   const __grid_constant__ CUtensorMap map_w1_lo,
                   const __grid_constant__ CUtensorMap map_w2_hi,
                   const __grid_constant__ CUtensorMap map_w2_lo,
                   const __grid_constant__ CUtensorMap map_xa_hi,   // addend tiles of x (staged epilogue)
                   const __grid_constant__ CUtensorMap map_xa_lo,
                   const __grid_constant__ CUtensorMap map_r_hi,    // residual output planes
                   const __grid_constant__ CUtensorMap map_r_lo,
                   const __grid_constant__ CUtensorMap map_z_hi,    // saved z planes / sigmoid (staged
                   const __grid_constant__ CUtensorMap map_z_lo,    // gate of H_b, P.stage_gate)
                   const __grid_constant__ CUtensorMap map_sig, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the dynamic window is only guaranteed 16-byte aligned: round up to the swizzle period
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* b1s = reinterpret_cast<float*>(smem + PST * PSTAGE + 4 * REGION_BYTES);   // [512] gate bias
  float* brs = b1s + CD;                                                // [Cr]
  float* bss = brs + P.Cr;                                              // [Cs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bss + P.Cs);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + PST);
  const uint32_t hfull0 = smem_u32(bars + 2 * PST), zready0 = hfull0 + 16;
  const uint32_t ofull0 = hfull0 + 32, oempty0 = hfull0 + 48, afull0 = hfull0 + 64;
  const uint32_t sready0 = hfull0 + 80, gready0 = hfull0 + 96;   // staging tiles written (-> store warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PST + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader of the pair (blockIdx.x even)
  const int b = blockIdx.y, t0 = blockIdx.x * TM;
  constexpr int XLO = X3 | F16;   // the residual stream keeps its lo plane (residual-add operand)
  const int nplanes = X3 ? 2 : 1;
  const bool rec_cta = P.dbg != nullptr && blockIdx.x == P.dbg_x && blockIdx.y == P.dbg_y;
  if (rec_cta && threadIdx.x == 0) P.dbg[40] = clock64();
  const int chunks_per_tap = P.Cr / BK;
  const int nk1 = P.fs * chunks_per_tap + P.Cc / BK;     // K slabs of the first contraction
  const int nk2 = CH / BK;                               // K slabs of the second contraction
  const int n_res = P.Cr / ON;                           // output chunks: residual rows, then skip rows
  const int o_begin = P.write_residual ? 0 : n_res;
  const int o_end = n_res + P.Cs / ON;

  if (warp == W_TMA && lane == 0) {
    prefetch_tmap(&map_x_hi); prefetch_tmap(&map_c_hi); prefetch_tmap(&map_w1_hi);
    prefetch_tmap(&map_w2_hi);
    if (X3) {
      prefetch_tmap(&map_x_lo); prefetch_tmap(&map_c_lo); prefetch_tmap(&map_w1_lo);
      prefetch_tmap(&map_w2_lo);
    }
    for (int s = 0; s < PST; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(hfull0 + 8 * g, 1);
      mbar_init(zready0 + 8 * g, 2 * FWD_EPI_WARPS);   // one elected lane per epilogue warp, both CTAs
      mbar_init(ofull0 + 8 * g, 1);
      mbar_init(oempty0 + 8 * g, 2 * FWD_EPI_WARPS);
      mbar_init(afull0 + 8 * g, 1);
      mbar_init(sready0 + 8 * g, FWD_EPI_WARPS);       // one elected lane per epilogue warp (this CTA)
    }
    mbar_init(gready0, FWD_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp < FWD_EPI_WARPS) {
    for (int i = threadIdx.x; i < CD; i += FWD_EPI_WARPS * 32) b1s[i] = P.gbias[(int64_t)b * CD + i];
    for (int i = threadIdx.x; i < P.Cr; i += FWD_EPI_WARPS * 32) brs[i] = P.res_b[i];
    for (int i = threadIdx.x; i < P.Cs; i += FWD_EPI_WARPS * 32) bss[i] = P.skip_b[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (rec_cta && threadIdx.x == 0) P.dbg[41] = clock64();
  // tile h (0 = channels [0,64), 1 = [64,128)) of plane pl (0 = hi, 1 = lo) of staging buffer bf
  // buffer 0 = the activation areas of the four ring stages (free in the output phase),
  // buffer 1 = the dedicated 64 KB behind the ring
  auto tile_addr = [&](int bf, int pl, int h) -> uint32_t {
    return bf ? base + (uint32_t)(PST * PSTAGE) + (uint32_t)(2 * pl + h) * REGION_BYTES
              : base + (uint32_t)(2 * pl + h) * PSTAGE;
  };
  // TMA-load the addend x[t0 .. t0+127][128 oc .. +127] (hi, lo) of residual chunk oc
  auto issue_addend = [&](int bf, int oc) {
    const uint32_t bar = afull0 + 8 * bf;
    mbar_expect_tx(bar, ((X3 | F16) ? 4 : 2) * REGION_BYTES);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tma_load_3d(tile_addr(bf, 0, h), &map_xa_hi, bar, oc * ON + 64 * h, t0, b);
      if (X3 | F16) tma_load_3d(tile_addr(bf, 1, h), &map_xa_lo, bar, oc * ON + 64 * h, t0, b);
    }
  };

  if (warp == W_TMA) {
    // =============================== TMA producer ===============================
    // the whole warp walks the loop in uniform control flow and one elected lane issues (inside an
    // `if (lane == 0)` region every UTMALDG / UTCHMMA operand came out of per-thread registers
    // through an ELECT / R2UR waterfall: ~100+ cycles of issue per instruction)
    {
      const bool el = elect_one_sync();
      int stage = 0;
      uint32_t ph = 0;
      // activation slab i of the first contraction: tap slabs of x (row shift = causal delay,
      // negative rows are zero-filled by TMA), then the condition slabs
      auto a_src = [&](int i, const CUtensorMap*& mh, const CUtensorMap*& ml, int& c0, int& tt) {
        const int tap = i / chunks_per_tap;
        if (tap < P.fs) {
          mh = &map_x_hi; ml = &map_x_lo;
          c0 = (i - tap * chunks_per_tap) * BK;
          tt = t0 - P.dilation * (P.fs - 1 - tap);
        } else {
          mh = &map_c_hi; ml = &map_c_lo;
          c0 = (i - P.fs * chunks_per_tap) * BK;
          tt = t0;
        }
      };
      // H_a reads its activation slabs for the first time (HBM); with four 32 KB stages in flight
      // and ~4 k cycles of loaded HBM latency the ring covers 3/4 of what the MMAs consume (H_a
      // 49 k cycles against 41 k for H_b, which re-reads the slabs from L2).  An L2 prefetch runs
      // pf slabs ahead of the ring so that the ring's own loads find the lines in L2.
      const int pf = P.pf_dist < nk1 ? P.pf_dist : nk1;
      auto a_prefetch = [&](int i) {
        const CUtensorMap *mh, *ml;
        int c0, tt;
        a_src(i, mh, ml, c0, tt);
        if (el) tma_prefetch_3d(mh, c0, tt, b);
        if (X3 && el) tma_prefetch_3d(ml, c0, tt, b);
      };
      for (int i = 0; i < pf; ++i) a_prefetch(i);
      for (int gp = 0; gp < 2; ++gp) {
        for (int i = 0; i < nk1; ++i) {
          if (gp == 0 && pf > 0 && i + pf < nk1) a_prefetch(i + pf);
          mbar_wait(empty0 + 8 * stage, ph ^ 1);
          // both CTAs' copies complete on the LEADER's barrier, which expects the bytes of both
          const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
          const uint32_t sa = base + stage * PSTAGE;
          if (rank == 0 && el) mbar_expect_tx(full0 + 8 * stage, 2 * nplanes * (A_PLANE + PB_PLANE));
          const CUtensorMap *mh, *ml;
          int c0, tt;
          a_src(i, mh, ml, c0, tt);
          if (el) tma2_load_3d(sa, mh, fb, c0, tt, b);
          if (X3 && el) tma2_load_3d(sa + A_PLANE, ml, fb, c0, tt, b);
          // this CTA stages ITS half (128 rows) of the phase's 256 weight rows
          const int wr = gp * TN + (int)rank * (TN / 2);
          if (el) tma2_load_2d(sa + 2 * A_PLANE, &map_w1_hi, fb, i * BK, wr);
          if (X3 && el) tma2_load_2d(sa + 2 * A_PLANE + PB_PLANE, &map_w1_lo, fb, i * BK, wr);
          if (++stage == PST) { stage = 0; ph ^= 1; }
        }
      }
      for (int oc = o_begin; oc < o_end; ++oc) {
        for (int i = 0; i < nk2; ++i) {
          mbar_wait(empty0 + 8 * stage, ph ^ 1);
          const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
          const uint32_t sa = base + stage * PSTAGE;
          if (rank == 0 && el) mbar_expect_tx(full0 + 8 * stage, 2 * nplanes * POB_PLANE);
          const int wr = oc * ON + (int)rank * (ON / 2);
          if (el) tma2_load_2d(sa + 2 * A_PLANE, &map_w2_hi, fb, i * BK, wr);
          if (X3 && el) tma2_load_2d(sa + 2 * A_PLANE + POB_PLANE, &map_w2_lo, fb, i * BK, wr);
          if (++stage == PST) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer =================================
    if (rank == 0) {   // ONE (elected) thread of the pair issues the M = 256 MMAs of both SMs
      const bool el = elect_one_sync();
      int stage = 0;
      uint32_t ph = 0;
      const bool rec = rec_cta && el;
      const uint32_t idesc = idesc_for(IDESC_P, F16);
      const uint32_t idesc_o = idesc_for(IDESC_PON, F16);
      for (int gp = 0; gp < 2; ++gp) {
        const uint32_t acc = tmem_base + 256 * gp;
        if (rec) P.dbg[2 * gp] = clock64();
        for (int i = 0; i < nk1; ++i) {
          mbar_wait(full0 + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + stage * PSTAGE;
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            const uint32_t a_hi = desc_lo_sw64(sa + ks * UK * 2);
            const uint32_t b_hi = desc_lo_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            if (el) mma2_ss_w<DESC_HI_SW64>(acc, a_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
            if (X3 && el) {
              mma2_ss_w<DESC_HI_SW64>(acc, desc_lo_sw64(sa + A_PLANE + ks * UK * 2), b_hi, idesc, 1u);
              mma2_ss_w<DESC_HI_SW64>(acc, a_hi, desc_lo_sw64(sa + 2 * A_PLANE + PB_PLANE + ks * UK * 2),
                                      idesc, 1u);
            }
          }
          if (el) tc_commit2(empty0 + 8 * stage);
          if (++stage == PST) { stage = 0; ph ^= 1; }
        }
        if (el) tc_commit2(hfull0 + 8 * gp);
        if (rec) P.dbg[2 * gp + 1] = clock64();
      }
      // z_a is in TMEM and the sigmoid half of H_a (= the first output accumulator) is drained:
      // the K steps of output chunk 0 that contract z_a run under the gate of H_b; everything
      // else needs z_b in TMEM and the sigmoid half of H_b drained
      mbar_wait(zready0, 0);
      tc_fence_after();
      bool zb_ready = false;
      for (int oc = o_begin, j = 0; oc < o_end; ++oc, ++j) {
        const int buf = j & 1, use = j >> 1;
        if (j > 0 && !zb_ready) {
          mbar_wait(zready0 + 8, 0);
          tc_fence_after();
          zb_ready = true;
        }
        if (use > 0) {
          mbar_wait(oempty0 + 8 * buf, (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (buf ? OACC1 : OACC0);
        if (rec && j < 6) P.dbg[4 + 2 * j] = clock64();
        for (int i = 0; i < nk2; ++i) {
          if (2 * i >= 8 && !zb_ready) {      // first K slab of z_b
            mbar_wait(zready0 + 8, 0);
            tc_fence_after();
            zb_ready = true;
          }
          mbar_wait(full0 + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + stage * PSTAGE;
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            // z channels [32 i + 16 ks, +16): hi pairs in 8 columns, lo pairs in the next 8
            const int kstep = 2 * i + ks;
            const uint32_t z_hi = tmem_base + (kstep < 8 ? 0 : 256) + 16 * (kstep & 7);
            const uint32_t z_lo = z_hi + 8;
            const uint32_t b_hi = desc_lo_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            if (el) mma2_ts_w<DESC_HI_SW64>(acc, z_hi, b_hi, idesc_o, (i | ks) ? 1u : 0u);
            if (X3 && el) {
              mma2_ts_w<DESC_HI_SW64>(acc, z_lo, b_hi, idesc_o, 1u);
              mma2_ts_w<DESC_HI_SW64>(acc, z_hi, desc_lo_sw64(sa + 2 * A_PLANE + POB_PLANE + ks * UK * 2),
                                      idesc_o, 1u);
            }
          }
          if (el) tc_commit2(empty0 + 8 * stage);
          if (++stage == PST) { stage = 0; ph ^= 1; }
        }
        if (el) tc_commit2(ofull0 + 8 * buf);
        if (rec && j < 6) P.dbg[4 + 2 * j + 1] = clock64();
      }
    }
  } else if (warp == W_ST) {
    // =============================== store warp =================================
    // Everything that moves staging tiles: the TMA stores of finished tiles, the wait until a
    // store has read its tile, and the TMA load of the next addend into the freed buffer.  With
    // the epilogue's own leader thread doing this, every residual chunk waited ~3 k cycles for
    // `cp.async.bulk.wait_group.read` before 512 threads could start (timeline: the gap between
    // the end of a chunk's MMAs and the start of its epilogue).
    if (P.stage_res || P.stage_gate) {
      const bool el = elect_one_sync();
      if (P.stage_gate) {
        mbar_wait(gready0, 0);                 // sigma / z tiles of the second gate phase are written
        if (el) {
#pragma unroll
          for (int st = 0; st < 4; ++st)
            tma_store_3d(&map_sig, tile_addr(0, st >> 1, st & 1), HALF + 32 * st, t0, b);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_store_3d(&map_z_hi, tile_addr(1, 0, h), HALF + 64 * h, t0, b);
            if (X3) tma_store_3d(&map_z_lo, tile_addr(1, 1, h), HALF + 64 * h, t0, b);
          }
          tma_store_commit();
        }
      } else {
        mbar_wait(hfull0 + 8, 0);              // every MMA of the first contraction is done: the
        tc_fence_after();                      // activation areas of the ring are free
      }
      if (P.stage_res) {
        if (el) {
          if (P.stage_gate) tma_store_wait_read();
          // the addends of the first two residual chunks land under the MMAs of chunk 0
          for (int j = 0; j < 2 && j < n_res; ++j) issue_addend(j, j);
        }
        for (int j = 0; j < n_res; ++j) {      // stage_res implies write_residual: chunk j = rows 128 j..
          const int buf = j & 1;
          mbar_wait(sready0 + 8 * buf, (j >> 1) & 1);
          if (el) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_store_3d(&map_r_hi, tile_addr(buf, 0, h), j * ON + 64 * h, t0, b);
              if (X3 | F16) tma_store_3d(&map_r_lo, tile_addr(buf, 1, h), j * ON + 64 * h, t0, b);
            }
            tma_store_commit();
            if (j + 2 < n_res) {               // the buffer's next user: chunk j + 2
              tma_store_wait_read();
              issue_addend(buf, j + 2);
            }
          }
        }
      }
      if (el) tma_store_wait_all();            // the tiles must outlive their stores
    }
  } else {
    // =============================== epilogue (warps 0-15) ======================
    // warp e: TMEM lane quadrant e%4 (hardware rule: a warp reaches lanes 32*(warp%4)..+31),
    // column group e/4 -- the 16-column chunks of a phase are dealt round-robin to the 4 groups,
    // so every SM sub-partition has 4 resident epilogue warps to hide latencies with.
    const int quad = warp & 3, grp = warp >> 2;
    constexpr int NG = FWD_EPI_WARPS / 4;
    const int row = quad * 32 + lane;
    const int t = t0 + row;
    const bool t_ok = t < P.T;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool rec = rec_cta && threadIdx.x == 0;
    const bool save_gates = P.gate_sig != nullptr && t_ok;
    const uint32_t zready_l = mapa_u32(zready0, 0), oempty_l = mapa_u32(oempty0, 0);   // the leader's
    // ---- gate phases: z = tanh(h_t) * sigmoid(h_s), written over the tanh columns ----
    for (int gp = 0; gp < 2; ++gp) {
      mbar_wait(hfull0 + 8 * gp, 0);
      tc_fence_after();
      if (rec) P.dbg[16 + 2 * gp] = clock64();
      // staged gate of H_b: sigmoid and the z planes of this phase go through the two staging
      // buffers (sigma: four [128 x 32] fp32 tiles in buffer 0, z hi/lo: four [128 x 64] bf16 tiles in
      // buffer 1) and leave as TMA stores -- 128 KB of per-thread 32-byte stores were the whole
      // exposed cost of this phase
      const bool stage_gate = gp == 1 && P.stage_gate;
      const uint32_t rswg = (uint32_t)(row & 7);
      const uint32_t accb = lane_base + 256 * gp;
#pragma unroll 1
      for (int q = grp; q < HALF / 16; q += NG) {
        uint32_t ar[16], gr[16];
        tmem_ld16_issue(accb + 16 * q, ar);
        tmem_ld16_issue(accb + HALF + 16 * q, gr);
        tmem_ld_wait(ar, gr);
        uint32_t zh[8], zl[8];
        const int ch0 = gp * HALF + 16 * q;
        float bt[16], bs[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          *reinterpret_cast<float4*>(bt + 4 * i) = *reinterpret_cast<const float4*>(b1s + ch0 + 4 * i);
          *reinterpret_cast<float4*>(bs + 4 * i) = *reinterpret_cast<const float4*>(b1s + CH + ch0 + 4 * i);
        }
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float z2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float th, sg;
            gate_pair(__uint_as_float(ar[i + u]) + bt[i + u], __uint_as_float(gr[i + u]) + bs[i + u],
                      th, sg);
            z2[u] = th * sg;
            ar[i + u] = __float_as_uint(th);     // kept for the backward (gate derivative)
            gr[i + u] = __float_as_uint(sg);
          }
          if (X3) split_pair_f(z2[0], z2[1], zh[i >> 1], zl[i >> 1], F16);
          else zh[i >> 1] = pack_pair_f(z2[0], z2[1], F16);
        }
        if (stage_gate) {
          // 16-byte chunks of a 128-byte tile row, XOR-swizzled with the row (SWIZZLE_128B)
          const uint32_t rowb = (uint32_t)row * 128u;
          const int st = (16 * q) >> 5, c0 = ((16 * q) & 31) >> 2;
          const uint32_t sb = tile_addr(0, st >> 1, st & 1) + rowb;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts128(sb + (((uint32_t)(c0 + c) ^ rswg) << 4), gr[4 * c], gr[4 * c + 1], gr[4 * c + 2],
                   gr[4 * c + 3]);
          const int h = (16 * q) >> 6, k0 = ((16 * q) & 63) >> 3;
          const uint32_t zb = tile_addr(1, 0, h) + rowb;
          const uint32_t o0 = ((uint32_t)k0 ^ rswg) << 4, o1 = ((uint32_t)(k0 + 1) ^ rswg) << 4;
          sts128(zb + o0, zh[0], zh[1], zh[2], zh[3]);
          sts128(zb + o1, zh[4], zh[5], zh[6], zh[7]);
          if (X3) {
            const uint32_t lb = tile_addr(1, 1, h) + rowb;
            sts128(lb + o0, zl[0], zl[1], zl[2], zl[3]);
            sts128(lb + o1, zl[4], zl[5], zl[6], zl[7]);
          }
        } else if (save_gates) {   // time-major (B,T,Ch) fp32: 64 contiguous bytes per thread and array
          const int64_t goff = ((int64_t)b * P.T + t) * CH + ch0;
          if (!X3) {
            st256(P.gate_tanh + goff, ar);
            st256(P.gate_tanh + goff + 8, ar + 8);
          }
          st256(P.gate_sig + goff, gr);
          st256(P.gate_sig + goff + 8, gr + 8);
        }
        tmem_st8(accb + 16 * q, zh);
        if (X3) tmem_st8(accb + 16 * q + 8, zl);
        if (!stage_gate && P.zp_hi != nullptr && t_ok) {
          const int64_t zoff = ((int64_t)b * P.T + t) * CH + ch0;
          st256(P.zp_hi + zoff, zh);
          if (X3) st256(P.zp_lo + zoff, zl);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(zready_l + 8 * gp);
      if (stage_gate) {
        fence_async_smem();                      // tile writes -> visible to the TMA stores
        __syncwarp();
        if (lane == 0) mbar_arrive(gready0);     // the store warp takes it from here
      }
      if (rec) P.dbg[16 + 2 * gp + 1] = clock64();
    }
    // ---- output chunks: residual rows then skip rows, 128 per chunk ----
    // residual = Wr z + br + x with x read back from the packed hi/lo planes (x = hi + lo to
    // 2^-17: two 32-byte loads per 16 channels instead of 16 strided fp32 loads); the running
    // skip sum is fp32 (B,Cs,T): lanes are consecutive t, so every access is a coalesced 128-byte
    // row segment.  Operands of the NEXT 16-column piece are fetched before the current one is
    // processed.
    for (int oc = o_begin, j = 0; oc < o_end; ++oc, ++j) {
      const int buf = j & 1, use = j >> 1;
      const bool is_res = oc < n_res;
      const int cbase = (is_res ? oc : oc - n_res) * ON;
      const uint32_t accb = lane_base + (buf ? OACC1 : OACC0);
      if (is_res && P.stage_res) {
        // ---------- staged residual chunk: addend and result live in shared-memory tiles ----------
        mbar_wait(afull0 + 8 * buf, use & 1);
        mbar_wait(ofull0 + 8 * buf, use & 1);
        tc_fence_after();
        if (rec && j < 6) P.dbg[20 + 2 * j] = clock64();
        const uint32_t rsw = (uint32_t)(row & 7);
#pragma unroll 1
        for (int q = grp; q < ON / 16; q += NG) {
          float o[16];
          tmem_ld16(accb + 16 * q, o);
          const int ch0 = cbase + 16 * q;
          // 16 channels of this thread's row: two 16-byte chunks (k0, k0+1) of a 128-byte tile row,
          // XOR-swizzled with the row index (CU_TENSOR_MAP_SWIZZLE_128B) -> conflict-free
          const int h = (16 * q) >> 6, k0 = ((16 * q) & 63) >> 3;
          const uint32_t rowb = (uint32_t)row * 128u;
          const uint32_t a0 = tile_addr(buf, 0, h) + rowb + (((uint32_t)k0 ^ rsw) << 4);
          const uint32_t a1 = tile_addr(buf, 0, h) + rowb + (((uint32_t)(k0 + 1) ^ rsw) << 4);
          const uint32_t lob = tile_addr(buf, 1, h) - tile_addr(buf, 0, h);
          const uint32_t l0 = a0 + lob, l1 = a1 + lob;
          const uint4 h0 = lds128(a0), h1 = lds128(a1);
          uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
          if (XLO) { w0 = lds128(l0); w1 = lds128(l1); }
          const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          const uint32_t lw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float bb[16];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(bb + 4 * i) = *reinterpret_cast<const float4*>(brs + ch0 + 4 * i);
          uint32_t rh[8], rl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0, v1;
            unpack_pair_f(hw[i], F16, v0, v1);
            if (XLO) {
              float e0, e1;
              unpack_pair_f(lw[i], F16, e0, e1);
              v0 += e0;
              v1 += e1;
            }
            v0 += o[2 * i] + bb[2 * i];
            v1 += o[2 * i + 1] + bb[2 * i + 1];
            if (P.res_f32 != nullptr && t_ok) {
              __stcs(P.res_f32 + ((int64_t)b * P.Cr + ch0 + 2 * i) * P.T + t, v0);
              __stcs(P.res_f32 + ((int64_t)b * P.Cr + ch0 + 2 * i + 1) * P.T + t, v1);
            }
            if (XLO) split_pair_sat_f(v0, v1, rh[i], rl[i], F16);
            else rh[i] = pack_pair_f(v0, v1, F16);
          }
          sts128(a0, rh[0], rh[1], rh[2], rh[3]);
          sts128(a1, rh[4], rh[5], rh[6], rh[7]);
          if (XLO) {
            sts128(l0, rl[0], rl[1], rl[2], rl[3]);
            sts128(l1, rl[4], rl[5], rl[6], rl[7]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(oempty_l + 8 * buf);   // the accumulator is drained
        fence_async_smem();                      // my tile writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(sready0 + 8 * buf);   // the store warp stores the tile
        if (rec && j < 6) P.dbg[20 + 2 * j + 1] = clock64();
        continue;
      }
      float addf[16];
      uint32_t hw[8], lw[8];
      auto fetch = [&](int q) {
        const int ch0 = cbase + 16 * q;
        if (is_res) {
          const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
          if (t_ok) {
            ld256(P.xp_hi + poff, hw);
            if (XLO) ld256(P.xp_lo + poff, lw);
          }
        } else if (P.skip_accumulate && t_ok) {
          const float* sp = P.skip + ((int64_t)b * P.Cs + ch0) * P.T + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) addf[i] = __ldcs(sp + (int64_t)i * P.T);
        }
      };
      fetch(grp);
      mbar_wait(ofull0 + 8 * buf, use & 1);
      tc_fence_after();
      if (rec && j < 6) P.dbg[20 + 2 * j] = clock64();
#pragma unroll 1
      for (int q = grp; q < ON / 16; q += NG) {
        float o[16], add[16];
        tmem_ld16(accb + 16 * q, o);
        if (is_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0, v1;
            unpack_pair_f(hw[i], F16, v0, v1);
            if (XLO) {
              float l0, l1;
              unpack_pair_f(lw[i], F16, l0, l1);
              v0 += l0;
              v1 += l1;
            }
            add[2 * i] = t_ok ? v0 : 0.0f;
            add[2 * i + 1] = t_ok ? v1 : 0.0f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) add[i] = (P.skip_accumulate && t_ok) ? addf[i] : 0.0f;
        }
        if (q + NG < ON / 16) fetch(q + NG);
        const int ch0 = cbase + 16 * q;
        if (is_res) {
          uint32_t rh[8], rl[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float v2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ch = ch0 + i + u;
              const float v = o[i + u] + brs[ch] + add[i + u];
              if (t_ok && P.res_f32 != nullptr)
                __stcs(P.res_f32 + ((int64_t)b * P.Cr + ch) * P.T + t, v);
              v2[u] = v;
            }
            if (XLO) split_pair_sat_f(v2[0], v2[1], rh[i >> 1], rl[i >> 1], F16);
            else rh[i >> 1] = pack_pair_f(v2[0], v2[1], F16);
          }
          if (t_ok && P.res_hi != nullptr) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
            st256(P.res_hi + poff, rh);
            if (XLO) st256(P.res_lo + poff, rl);
          }
        } else if (t_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ch = ch0 + i;
            P.skip[((int64_t)b * P.Cs + ch) * P.T + t] = o[i] + bss[ch] + add[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(oempty_l + 8 * buf);
      if (rec && j < 6) P.dbg[20 + 2 * j + 1] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair's MMAs can still touch it
  if (warp == W_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
    if (rec_cta && lane == 0) P.dbg[42] = clock64();
  }
}

// ------------------------------------------------------------------ packing kernels --------
// (B,C,T) fp32 -> (B,T,C) bf16 hi/lo planes: 32x32 transpose through shared memory
__global__ void __launch_bounds__(256)
pack_act_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                __nv_bfloat16* __restrict__ lo, int C, int T, int pitch, int relu, int f16,
                const float* __restrict__ scale, int ones_ch) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float sc = scale ? scale[0] : 1.0f;   // power-of-two gradient scale (fp16 backward)
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r, t = t0 + tx;
    tile[r][tx] = (c < C && t < T) ? in[((int64_t)b * C + c) * T + t] * sc : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int t = t0 + r, c = c0 + tx;
    if (t < T && c < pitch) {          // channels in [C, pitch) are zero padding ...
      __nv_bfloat16 h, l;
      float v = tile[tx][r];
      if (c == ones_ch) v = 1.0f;      // ... except the constant-one channel (bias-gradient column)
      if (relu) v = fmaxf(v, 0.0f);
      split_16(v, f16, h, l);
      const int64_t off = ((int64_t)b * T + t) * pitch + c;
      hi[off] = h;
      if (lo) lo[off] = l;
    }
  }
}

int pack_act_launch_ex(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int T,
                       int pitch, int relu, int f16, const float* scale, cudaStream_t stream) {
  dim3 g(ceil_div(T, 32), ceil_div(pitch, 32), B);
  pack_act_kernel<<<g, 256, 0, stream>>>(in, hi, lo, C, T, pitch, relu, f16, scale, -1);
  VQW_CHECK_LAUNCH("pack_act_kernel");
  return 0;
}
int pack_act_launch(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int T,
                    int f16, const float* scale, cudaStream_t stream) {
  return pack_act_launch_ex(in, hi, lo, B, C, T, C, 0, f16, scale, stream);
}

// W1 packed [512 rows in chunk order][K1 = fs*Cr + Cc]: row r -> original row; `half` gate pairs
// per accumulator chunk (v1: 128, v2: 64):
//   half = 128: [0,128) tanh 0..127 | [128,256) sigmoid 0..127 | [256,384) tanh 128..255 | ...
//   half = 64 : [0,64) tanh 0..63 | [64,128) sigmoid 0..63 | [128,192) tanh 64..127 | ...
__device__ __forceinline__ void
pack_w1(const float* __restrict__ conv_w, const float* __restrict__ cond_w,
        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cr, int Cc,
        int Cl, int fs, int f16, int half) {
  const int K1 = fs * Cr + Cl;         // only the Cl time-varying condition columns are contracted
  const int64_t n = (int64_t)CD * K1;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / K1), k = (int)(e % K1);
    const int quad = r / half, within = r % half;
    const int orig = ((quad & 1) ? CH : 0) + (quad >> 1) * half + within;
    float v;
    if (k < fs * Cr) {
      const int j = k / Cr, c = k % Cr;
      v = conv_w[((int64_t)orig * Cr + c) * fs + j];
    } else {
      v = cond_w[(int64_t)orig * Cc + (k - fs * Cr)];
    }
    __nv_bfloat16 h, l;
    split_16(v, f16, h, l);
    hi[e] = h;
    if (lo) lo[e] = l;
  }
}

// W2 packed [(Cr + Cs) rows][Ch]: residual rows then skip rows
__device__ __forceinline__ void
pack_w2(const float* __restrict__ res_w, const float* __restrict__ skip_w,
        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cr, int Cs,
        int f16) {
  const int64_t n = (int64_t)(Cr + Cs) * CH;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / CH), k = (int)(e % CH);
    const float v = (r < Cr) ? res_w[(int64_t)r * CH + k] : skip_w[(int64_t)(r - Cr) * CH + k];
    __nv_bfloat16 h, l;
    split_16(v, f16, h, l);
    hi[e] = h;
    if (lo) lo[e] = l;
  }
}

// both weight operands of up to GB_MAX blocks in one launch: grid (x, blocks); block i's planes
// sit at i * stride elements behind the given ones (w1 hi, w1 lo, w2 hi, w2 lo: see the layout)
struct PackWArgs {
  const float* conv_w[32];
  const float* cond_w[32];
  const float* res_w[32];
  const float* skip_w[32];
};
__global__ void __launch_bounds__(256)
pack_w_kernel(const __grid_constant__ PackWArgs A, __nv_bfloat16* __restrict__ w1h,
              __nv_bfloat16* __restrict__ w1l, __nv_bfloat16* __restrict__ w2h,
              __nv_bfloat16* __restrict__ w2l, int64_t stride, int Cr, int Cs, int Cc, int Cl, int fs,
              int f16, int half) {
  const int i = blockIdx.y;
  const int64_t o = (int64_t)i * stride;
  pack_w1(A.conv_w[i], A.cond_w[i], w1h + o, w1l ? w1l + o : nullptr, Cr, Cc, Cl, fs, f16, half);
  pack_w2(A.res_w[i], A.skip_w[i], w2h + o, w2l ? w2l + o : nullptr, Cr, Cs, f16);
}

// Per-(block, item) gate bias: conv_b + cond_b + W_p[:, Cl:] . g_b -- the condition projection of
// the time-constant (speaker) channels, modules.py:17-18,44 applied to net.py:59-61's broadcast.
constexpr int GB_MAX = 32;
struct GbiasArgs {
  const float* conv_b[GB_MAX];
  const float* cond_b[GB_MAX];
  const float* cond_w[GB_MAX];
};
// grid (blocks, 512 / 8): warp w of CTA (i, y) owns dilated channel 8 y + w; lanes stride the Cg
// global channels (coalesced weight row), one shuffle reduction per item
__global__ void __launch_bounds__(256)
gbias_kernel(const __grid_constant__ GbiasArgs A, const float* __restrict__ glob,
             float* __restrict__ out, int B, int Cc, int Cg, int blk0) {
  const int i = blockIdx.x, ch = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const float base = A.conv_b[i][ch] + A.cond_b[i][ch];
  const float* w = A.cond_w[i] + (int64_t)ch * Cc + (Cc - Cg);
  for (int b = 0; b < B; ++b) {
    float acc = 0.0f;
    for (int k = lane; k < Cg; k += 32) acc = fmaf(__ldg(w + k), __ldg(glob + (int64_t)b * Cg + k), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[((int64_t)(blk0 + i) * B + b) * CD + ch] = base + acc;
  }
}

// ------------------------------------------------------------------ host side --------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor (rank 2 or 3), innermost extent `inner` contiguous; box = {BK, box_rows, 1}
int make_map(CUtensorMap* m, const void* ptr, int rank, uint64_t inner, uint64_t rows,
                    uint64_t batch, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  VQW_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {inner * 2, inner * rows * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VQW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int make_map_tile(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t T, uint64_t B) {
  EncodeTiledFn enc = get_encode();
  VQW_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[3] = {C, T, B};
  cuuint64_t strides[2] = {C * 2, C * T * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VQW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tile) failed with CUresult %d", (int)r);
  return 0;
}

// MN-major operand, two 64-channel groups per copy: the plane as {64 ch, T, C/64 groups, B}, box
// {64, 32, 2, 1} = two consecutive [32 steps x 128 B] swizzle tiles (8 KB) from ONE TMA instruction
// (the weight-gradient producer is bound by its TMA issue rate, ~130-180 cycles per copy)
int make_map_mn4(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t T, uint64_t B) {
  EncodeTiledFn enc = get_encode();
  VQW_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  VQW_REQUIRE(C % 128 == 0, "make_map_mn4: channel count must be a multiple of 128");
  cuuint64_t dims[4] = {64, T, C / 64, B};
  cuuint64_t strides[3] = {C * 2, 128, C * T * 2};
  cuuint32_t box[4] = {64, 32, 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VQW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(mn4) failed with CUresult %d", (int)r);
  return 0;
}

// [128 rows x 32 channels] fp32 tiles of a time-major (B,T,C) fp32 tensor, 128-byte rows
int make_map_tile_f32(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t T, uint64_t B) {
  EncodeTiledFn enc = get_encode();
  VQW_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[3] = {C, T, B};
  cuuint64_t strides[2] = {C * 4, C * T * 4};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VQW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tile f32) failed with CUresult %d", (int)r);
  return 0;
}

int make_map_mn(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t pitch, uint64_t T,
                uint64_t B) {
  EncodeTiledFn enc = get_encode();
  VQW_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[3] = {C, T, B};
  cuuint64_t strides[2] = {pitch * 2, pitch * T * 2};
  cuuint32_t box[3] = {64, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VQW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(mn) failed with CUresult %d", (int)r);
  return 0;
}

size_t smem_bytes_pair(int Cr, int Cs) {
  return 1024 + (size_t)PST * PSTAGE + 4 * REGION_BYTES + sizeof(float) * (CD + Cr + Cs) +
         8 * (2 * PST + 13) + 16;
}
size_t smem_bytes(int Cr, int Cs) {
  return 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(float) * (CD + Cr + Cs) + 8 * NBAR + 16;
}
}  // namespace tc

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

bool resnet_tc_supported(const vqw_resnet_desc& d) {
  const int Cl = d.Cc - d.Cg;
  return d.Cd == tc::CD && d.Cr % tc::TN == 0 && d.Cs % tc::TN == 0 && d.Cg >= 0 && Cl >= tc::BK &&
         Cl % tc::BK == 0 && d.Cr >= tc::TN && d.Cs >= tc::TN && d.fs >= 1 && d.T >= tc::TM &&
         d.T % 8 == 0;
}

// workspace: [cond hi|lo] [x ping hi|lo] [x pong hi|lo] [per block: w1 hi|lo, w2 hi|lo]
struct TcWorkspace {
  int64_t cond_plane, x_plane, w1_plane, w2_plane, block_stride, total;
  int64_t off_cond, off_x[2], off_w, off_gbias;
};
static TcWorkspace tc_layout(const vqw_resnet_desc& d) {
  TcWorkspace w;
  const int K1 = d.fs * d.Cr + cond_local(d);
  w.cond_plane = align_up((int64_t)d.B * d.T * cond_pitch(d) * 2, 1024);
  w.x_plane = align_up((int64_t)d.B * d.T * d.Cr * 2, 1024);
  w.w1_plane = align_up((int64_t)tc::CD * K1 * 2, 1024);
  w.w2_plane = align_up((int64_t)(d.Cr + d.Cs) * tc::CH * 2, 1024);
  w.block_stride = 2 * w.w1_plane + 2 * w.w2_plane;
  w.off_cond = 0;
  w.off_x[0] = 2 * w.cond_plane;
  w.off_x[1] = w.off_x[0] + 2 * w.x_plane;
  w.off_w = w.off_x[1] + 2 * w.x_plane;
  w.off_gbias = w.off_w + (int64_t)d.n_blocks * w.block_stride;
  w.total = w.off_gbias + align_up((int64_t)d.n_blocks * d.B * tc::CD * 4, 1024);
  return w;
}

int64_t resnet_tc_workspace(const vqw_resnet_desc& d) { return tc_layout(d).total + 1024; }
int64_t resnet_tc_saved_bytes(const vqw_resnet_desc& d) { return tc_saved_layout(d).total; }

// vqw_probe_forward_kernels: events recorded around the block kernels of the NEXT forward call
static void* g_probe_events[2] = {nullptr, nullptr};

int resnet_forward_tc(const vqw_resnet_desc& d, const float* x, const float* cond,
                      const vqw_resblock_weights* weights, float* const* residuals, float* skip,
                      float* const* gate_tanh, float* const* gate_sig, void* workspace,
                      void* saved, cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(resnet_tc_supported(d),
              "tcgen05 path needs dilated_channels=512, residual/skip channels multiples of 256, "
              "time-varying condition channels a multiple of 32, T >= 128 and T %% 8 == 0 "
              "(got Cr=%d Cd=%d Cs=%d Cc=%d Cg=%d T=%d)", d.Cr, d.Cd, d.Cs, d.Cc, d.Cg, d.T);
  VQW_REQUIRE(d.Cg == 0 || d.cond_global != nullptr, "vqw_resnet_forward: Cg > 0 needs cond_global");
  VQW_REQUIRE(workspace != nullptr, "vqw_resnet_forward: workspace is null");
  VQW_REQUIRE(d.B <= 65535, "vqw_resnet_forward: B > 65535");
  const bool x3 = vqw_mode_x3(d.mode);
  const int f16 = vqw_mode_f16(d.mode) ? 1 : 0;
  const bool xlo = x3 || f16;   // residual stream hi + lo (only the hi plane feeds the MMAs in fp16)
  // The CTA-pair kernel (cta_group::2, each SM stages half of every weight slab) is the default;
  // VQW_TC_FWD_PAIR=0 selects the single-CTA kernel.
  const char* penv = getenv("VQW_TC_FWD_PAIR");
  const bool pair = !(penv && penv[0] == '0');
  // weight rows per TMA box = the rows of a slab one CTA stages
  const int wrows = pair ? TN / 2 : TN, wrows2 = pair ? ON / 2 : ON;
  const TcWorkspace L = tc_layout(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up((int64_t)(uintptr_t)workspace, 1024));
  auto plane = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  // activation planes: in the caller's `saved` buffer when training (every block's input and
  // z are kept for the backward), else ping-pong in the workspace
  const TcSaved S = tc_saved_layout(d);
  uint8_t* sv = saved ? reinterpret_cast<uint8_t*>(align_up((int64_t)(uintptr_t)saved, 1024)) : nullptr;
  auto splane = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(sv + off); };
  __nv_bfloat16* c_hi = sv ? splane(S.cond[0]) : plane(L.off_cond);
  __nv_bfloat16* c_lo = sv ? splane(S.cond[1]) : plane(L.off_cond + L.cond_plane);
  auto xin_hi = [&](int i) { return sv ? splane(S.x0 + i * S.x_stride) : plane(L.off_x[i & 1]); };
  auto xin_lo = [&](int i) {
    return sv ? splane(S.x0 + i * S.x_stride + S.x_plane) : plane(L.off_x[i & 1] + L.x_plane);
  };
  __nv_bfloat16* x_hi[2] = {xin_hi(0), nullptr};
  __nv_bfloat16* x_lo[2] = {xin_lo(0), nullptr};
  const int Cl = cond_local(d), CP = cond_pitch(d);
  const int K1 = d.fs * d.Cr + Cl;

  float* gbias = reinterpret_cast<float*>(ws + L.off_gbias);   // (n_blocks, B, 512) gate biases
  // pack the two inputs and every block's weights
  {
    dim3 g1(ceil_div(d.T, 32), ceil_div(d.Cr, 32), d.B), g2(ceil_div(d.T, 32), ceil_div(CP, 32), d.B);
    pack_act_kernel<<<g1, 256, 0, stream>>>(x, x_hi[0], xlo ? x_lo[0] : nullptr, d.Cr, d.T, d.Cr, 0,
                                            f16, nullptr, -1);
    VQW_CHECK_LAUNCH("pack_act_kernel(x)");
    // condition planes: Cl channels, the constant-one channel, zero padding (see cond_pitch)
    pack_act_kernel<<<g2, 256, 0, stream>>>(cond, c_hi, x3 ? c_lo : nullptr, Cl, d.T, CP, 0, f16,
                                            nullptr, Cl);
    VQW_CHECK_LAUNCH("pack_act_kernel(cond)");
    for (int i0 = 0; i0 < d.n_blocks; i0 += GB_MAX) {
      GbiasArgs A = {};
      const int nb = d.n_blocks - i0 < GB_MAX ? d.n_blocks - i0 : GB_MAX;
      for (int i = 0; i < nb; ++i) {
        const vqw_resblock_weights& w = weights[i0 + i];
        VQW_REQUIRE(w.conv_b && w.cond_b && w.cond_w, "vqw_resnet_forward: block %d has a null weight",
                    i0 + i);
        A.conv_b[i] = w.conv_b; A.cond_b[i] = w.cond_b; A.cond_w[i] = w.cond_w;
      }
      gbias_kernel<<<dim3(nb, CD / 8), 256, 0, stream>>>(A, d.cond_global, gbias, d.B, d.Cc, d.Cg, i0);
      VQW_CHECK_LAUNCH("gbias_kernel");
    }
    for (int i = 0; i < d.n_blocks; ++i) {
      const vqw_resblock_weights& w = weights[i];
      VQW_REQUIRE(w.conv_w && w.conv_b && w.cond_w && w.cond_b && w.res_w && w.res_b && w.skip_w &&
                      w.skip_b, "vqw_resnet_forward: block %d has a null weight", i);
    }
    for (int i0 = 0; i0 < d.n_blocks; i0 += 32) {
      PackWArgs A = {};
      const int nb = d.n_blocks - i0 < 32 ? d.n_blocks - i0 : 32;
      for (int i = 0; i < nb; ++i) {
        const vqw_resblock_weights& w = weights[i0 + i];
        A.conv_w[i] = w.conv_w; A.cond_w[i] = w.cond_w; A.res_w[i] = w.res_w; A.skip_w[i] = w.skip_w;
      }
      const int64_t o = L.off_w + i0 * L.block_stride;
      pack_w_kernel<<<dim3(74, nb), 256, 0, stream>>>(
          A, plane(o), x3 ? plane(o + L.w1_plane) : nullptr, plane(o + 2 * L.w1_plane),
          x3 ? plane(o + 2 * L.w1_plane + L.w2_plane) : nullptr, L.block_stride / 2, d.Cr, d.Cs, d.Cc,
          Cl, d.fs, f16, HALF);
      VQW_CHECK_LAUNCH("pack_w_kernel");
    }
  }

  const size_t smem = pair ? smem_bytes_pair(d.Cr, d.Cs) : smem_bytes(d.Cr, d.Cs);
  auto kern1 = x3 ? (f16 ? resblock_tc_kernel<1, 1> : resblock_tc_kernel<1, 0>)
                  : (f16 ? resblock_tc_kernel<0, 1> : resblock_tc_kernel<0, 0>);
  auto kern2 = x3 ? (f16 ? resblock_tc_pair_kernel<1, 1> : resblock_tc_pair_kernel<1, 0>)
                  : (f16 ? resblock_tc_pair_kernel<0, 1> : resblock_tc_pair_kernel<0, 0>);
  if (pair) {
    VQW_CHECK_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  } else {
    VQW_CHECK_CUDA(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  CUtensorMap m_c_hi, m_c_lo;
  if (int rc = make_map(&m_c_hi, c_hi, 3, CP, d.T, d.B, TM)) return rc;
  if (int rc = make_map(&m_c_lo, x3 ? c_lo : c_hi, 3, CP, d.T, d.B, TM)) return rc;

  void* probe[2] = {g_probe_events[0], g_probe_events[1]};
  g_probe_events[0] = g_probe_events[1] = nullptr;
  if (probe[0]) VQW_CHECK_CUDA(cudaEventRecord((cudaEvent_t)probe[0], stream));
  for (int i = 0; i < d.n_blocks; ++i) {
    const bool last = (i == d.n_blocks - 1);
    const bool write_res = !last || d.keep_last_residual;
    const int cur = 0, nxt = 1;
    x_hi[0] = xin_hi(i); x_lo[0] = xin_lo(i);
    x_hi[1] = last ? nullptr : xin_hi(i + 1);
    x_lo[1] = last ? nullptr : xin_lo(i + 1);
    const vqw_resblock_weights& w = weights[i];
    __nv_bfloat16* w1h = plane(L.off_w + i * L.block_stride);
    __nv_bfloat16* w1l = plane(L.off_w + i * L.block_stride + L.w1_plane);
    __nv_bfloat16* w2h = plane(L.off_w + i * L.block_stride + 2 * L.w1_plane);
    __nv_bfloat16* w2l = plane(L.off_w + i * L.block_stride + 2 * L.w1_plane + L.w2_plane);
    CUtensorMap m_x_hi, m_x_lo, m_w1_hi, m_w1_lo, m_w2_hi, m_w2_lo;
    if (int rc = make_map(&m_x_hi, x_hi[cur], 3, d.Cr, d.T, d.B, TM)) return rc;
    if (int rc = make_map(&m_x_lo, x3 ? x_lo[cur] : x_hi[cur], 3, d.Cr, d.T, d.B, TM)) return rc;
    if (int rc = make_map(&m_w1_hi, w1h, 2, K1, CD, 1, wrows)) return rc;
    if (int rc = make_map(&m_w1_lo, x3 ? w1l : w1h, 2, K1, CD, 1, wrows)) return rc;
    if (int rc = make_map(&m_w2_hi, w2h, 2, CH, d.Cr + d.Cs, 1, wrows2)) return rc;
    if (int rc = make_map(&m_w2_lo, x3 ? w2l : w2h, 2, CH, d.Cr + d.Cs, 1, wrows2)) return rc;
    // staged residual epilogue: addend tiles of this block's input, output tiles of the next one's
    const bool stage_res = write_res && !last;
    CUtensorMap m_xa_hi, m_xa_lo, m_r_hi, m_r_lo;
    if (int rc = make_map_tile(&m_xa_hi, x_hi[cur], d.Cr, d.T, d.B)) return rc;
    if (int rc = make_map_tile(&m_xa_lo, xlo ? x_lo[cur] : x_hi[cur], d.Cr, d.T, d.B)) return rc;
    if (int rc = make_map_tile(&m_r_hi, stage_res ? x_hi[nxt] : x_hi[cur], d.Cr, d.T, d.B)) return rc;
    if (int rc = make_map_tile(&m_r_lo, stage_res ? (xlo ? x_lo[nxt] : x_hi[nxt]) : x_hi[cur], d.Cr, d.T, d.B))
      return rc;
    Params P;
    P.stage_res = stage_res ? 1 : 0;
    P.stage_gate = 0;
    P.B = d.B; P.T = d.T; P.Cr = d.Cr; P.Cs = d.Cs; P.Cc = Cl; P.fs = d.fs;   // Cc: contracted channels
    P.dilation = d.dilations[i];
    P.x3 = x3 ? 1 : 0;
    P.f16 = f16;
    P.xlo = xlo ? 1 : 0;
    P.skip_accumulate = i > 0;
    P.write_residual = write_res ? 1 : 0;
    // L2 prefetch distance of H_a's activation slabs: 6 slabs ahead on CTA pairs (measured -1.3 % per
    // launch in the step, two A/B pairs), off on single CTAs (no gain there)
    P.pf_dist = getenv("VQW_TC_PREFETCH") ? atoi(getenv("VQW_TC_PREFETCH")) : (pair ? 6 : 0);
    P.xp_hi = x_hi[cur];
    P.xp_lo = x_lo[cur];
    P.gbias = gbias + (int64_t)i * d.B * CD;
    P.res_b = w.res_b; P.skip_b = w.skip_b;
    P.res_f32 = (write_res && residuals) ? residuals[i] : nullptr;
    P.res_hi = last ? nullptr : x_hi[nxt];
    P.res_lo = last ? nullptr : x_lo[nxt];
    P.skip = skip;
    P.gate_tanh = (gate_tanh && !x3) ? gate_tanh[i] : nullptr;
    P.gate_sig = gate_sig ? gate_sig[i] : nullptr;
    P.zp_hi = sv ? splane(S.z0 + i * S.z_stride) : nullptr;
    P.zp_lo = sv ? splane(S.z0 + i * S.z_stride + S.z_plane) : nullptr;
    VQW_REQUIRE(x3 || (P.gate_tanh == nullptr) == (P.gate_sig == nullptr),
                "vqw_resnet_forward: gate_tanh/gate_sig of block %d must be given together", i);
    VQW_REQUIRE(P.gate_sig == nullptr || sv != nullptr,
                "vqw_resnet_forward: saving the gates needs the `saved` buffer (z planes)");
    static long long* dbg_buf = nullptr;
    const bool timeline = getenv("VQW_TC_TIMELINE") && getenv("VQW_TC_TIMELINE")[0] == '1';
    P.dbg = nullptr; P.dbg_x = P.dbg_y = 0;
    if (timeline) {
      if (!dbg_buf) cudaMalloc(&dbg_buf, 128 * sizeof(long long));
      cudaMemsetAsync(dbg_buf, 0, 128 * sizeof(long long), stream);
      P.dbg = dbg_buf;
      P.dbg_x = getenv("VQW_TC_TIMELINE_X") ? atoi(getenv("VQW_TC_TIMELINE_X")) : 1;
      P.dbg_y = getenv("VQW_TC_TIMELINE_Y") ? atoi(getenv("VQW_TC_TIMELINE_Y")) : 0;
    }
    CUtensorMap m_z_hi = m_xa_hi, m_z_lo = m_xa_hi, m_sig = m_xa_hi;
    static const bool stage_gate_on = !(getenv("VQW_TC_STAGE_GATE") && getenv("VQW_TC_STAGE_GATE")[0] == '0');
    if (pair && x3 && stage_gate_on && P.gate_sig != nullptr && P.zp_hi != nullptr) {
      if (int rc = make_map_tile(&m_z_hi, P.zp_hi, CH, d.T, d.B)) return rc;
      if (int rc = make_map_tile(&m_z_lo, P.zp_lo, CH, d.T, d.B)) return rc;
      if (int rc = make_map_tile_f32(&m_sig, P.gate_sig, CH, d.T, d.B)) return rc;
      P.stage_gate = 1;
    }
    if (pair) {
      // clusters of two CTAs = two consecutive time tiles (an odd tile count gets one all-padding tile)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * ceil_div(ceil_div(d.T, TM), 2), d.B);
      cfg.blockDim = dim3(PAIR_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      VQW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern2, m_x_hi, m_x_lo, m_c_hi, m_c_lo, m_w1_hi, m_w1_lo,
                                        m_w2_hi, m_w2_lo, m_xa_hi, m_xa_lo, m_r_hi, m_r_lo, m_z_hi,
                                        m_z_lo, m_sig, P));
    } else {
      dim3 grid(ceil_div(d.T, TM), d.B);
      kern1<<<grid, FWD_THREADS, smem, stream>>>(m_x_hi, m_x_lo, m_c_hi, m_c_lo, m_w1_hi, m_w1_lo,
                                                m_w2_hi, m_w2_lo, m_xa_hi, m_xa_lo, m_r_hi, m_r_lo, P);
    }
    VQW_CHECK_LAUNCH("resblock_tc_kernel");
    if (timeline) {
      long long h[64];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[vqw timeline] block %d CTA (%d,%d): entry %lld, setup done %lld, exit %lld "
                      "(cycles rel. to first MMA phase start)\n", i, P.dbg_x, P.dbg_y, h[40] - h[0],
              h[41] - h[0], h[42] - h[0]);
      {
        for (int n = 0; n < 2; ++n)
          fprintf(stderr, "  H_%c : mma [%lld, %lld]  gate epilogue [%lld, %lld]\n", 'a' + n,
                  h[2 * n] - h[0], h[2 * n + 1] - h[0], h[16 + 2 * n] - h[0], h[16 + 2 * n + 1] - h[0]);
        for (int n = 0; n < 6; ++n)
          fprintf(stderr, "  O_%d : mma [%lld, %lld]  epilogue [%lld, %lld]\n", n, h[4 + 2 * n] - h[0],
                  h[5 + 2 * n] - h[0], h[20 + 2 * n] - h[0], h[21 + 2 * n] - h[0]);
      }
    }
  }
  if (probe[1]) VQW_CHECK_CUDA(cudaEventRecord((cudaEvent_t)probe[1], stream));
  return 0;
}

}  // namespace vqw

extern "C" int vqw_probe_forward_kernels(void* start_event, void* end_event) {
  vqw::g_probe_events[0] = start_event;
  vqw::g_probe_events[1] = end_event;
  return 0;
}
