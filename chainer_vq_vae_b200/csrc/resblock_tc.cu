// Fused WaveNet residual block on the 5th-generation tensor cores (sm_100a: tcgen05 + TMEM + TMA).
//
// Replaces ResidualBlock.__call__ (modules.py:30-56) and ResidualNet's skip accumulation
// (modules.py:92-95) for the 512-channel training config, where the block is two dense
// contractions per time step (K = fs*Cr + Cc = 1728 then K = Cd/2 = 256): 2.16 MFLOP per audio
// sample against ~13 KB of HBM traffic, i.e. tensor-core bound.
//
// Precision.  The reference computes in fp32.  tcgen05 has no fp32 operand type and
// kind::tf32 truncates fp32 operands (a -1e-3 systematic bias), so operands are SPLIT into two
// bf16 planes (hi = bf16(v), lo = bf16(v - hi), 16 significant bits) and every product is
// issued as three MMAs  hi*hi + lo*hi + hi*lo  accumulated in fp32 in TMEM ("bf16x3", error
// ~2^-16: parity mode).  VQW_MODE_BF16 issues only hi*hi (throughput mode, not parity).
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
constexpr int ACC_COL = 0, ZHI_COL = 256, ZLO_COL = 384;
constexpr int CD = 512, CH = 256, HALF = 128;             // dilated channels handled by this kernel

struct Params {
  int B, T, Cr, Cs, Cc, fs, dilation;
  int x3;               // 1: bf16x3 (three MMAs per product), 0: single pass over the hi planes
  int f16;              // planes hold IEEE fp16 instead of bf16 (VQW_MODE_FP16, single pass)
  int xlo;              // the residual stream keeps its lo plane (residual-add operand): x3 or fp16
  int skip_accumulate;
  int write_residual;   // 0 for the last block: O_0/O_1 are skipped entirely
  int pf_dist;          // L2 prefetch distance (K slabs) of the activation operand in H_a, 0 = off
  const __nv_bfloat16* xp_hi;   // packed (B,T,Cr) planes of the block input (residual-add operand)
  const __nv_bfloat16* xp_lo;
  const float* conv_b;  const float* cond_b;  const float* res_b;  const float* skip_b;
  float* res_f32;       // (B,Cr,T) fp32 or null (saved block input of the next block / API output)
  __nv_bfloat16* res_hi;  // packed (B,T,Cr) planes for the next block (null for the last block)
  __nv_bfloat16* res_lo;
  float* skip;          // (B,Cs,T) fp32 in/out
  float* gate_tanh;     // (B,Ch,T) fp32 or null
  float* gate_sig;
  __nv_bfloat16* zp_hi; // packed (B,T,Ch) planes of z = tanh*sigmoid, saved for the backward (or null)
  __nv_bfloat16* zp_lo;
  long long* dbg;       // optional phase timestamps of one CTA (VQW_TC_TIMELINE=1)
  int dbg_x, dbg_y;
};

// ------------------------------------------------------------------ the kernel -------------
__global__ void __launch_bounds__(FWD_THREADS, 1)
resblock_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi,
                   const __grid_constant__ CUtensorMap map_x_lo,
                   const __grid_constant__ CUtensorMap map_c_hi,
                   const __grid_constant__ CUtensorMap map_c_lo,
                   const __grid_constant__ CUtensorMap map_w1_hi,
                   const __grid_constant__ CUtensorMap map_w1_lo,
                   const __grid_constant__ CUtensorMap map_w2_hi,
                   const __grid_constant__ CUtensorMap map_w2_lo, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the dynamic window is only guaranteed 16-byte aligned: round up to the swizzle period
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* b1s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);   // [512] conv_b + cond_b
  float* brs = b1s + CD;                                                // [Cr]
  float* bss = brs + P.Cr;                                              // [Cs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bss + P.Cs);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t acc_full = smem_u32(bars + 2 * STAGES), acc_empty = smem_u32(bars + 2 * STAGES + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * TM;
  const int nplanes = P.x3 ? 2 : 1;
  const bool rec_cta = P.dbg != nullptr && blockIdx.x == P.dbg_x && blockIdx.y == P.dbg_y;
  if (rec_cta && threadIdx.x == 0) P.dbg[40] = clock64();
  const int chunks_per_tap = P.Cr / BK;
  const int nk1 = P.fs * chunks_per_tap + P.Cc / BK;     // K slabs of the first contraction
  const int nk2 = CH / BK;                               // K slabs of the second contraction
  const int o_begin = P.write_residual ? 0 : P.Cr / TN;  // first N chunk of [Wr ; Ws]
  const int o_end = P.Cr / TN + P.Cs / TN;

  if (warp == W_TMA && lane == 0) {
    prefetch_tmap(&map_x_hi); prefetch_tmap(&map_c_hi); prefetch_tmap(&map_w1_hi);
    prefetch_tmap(&map_w2_hi);
    if (P.x3) {
      prefetch_tmap(&map_x_lo); prefetch_tmap(&map_c_lo); prefetch_tmap(&map_w1_lo);
      prefetch_tmap(&map_w2_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, FWD_EPI_WARPS * 32);
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
    for (int i = threadIdx.x; i < CD; i += FWD_EPI_WARPS * 32) b1s[i] = P.conv_b[i] + P.cond_b[i];
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
        if (P.x3) tma_prefetch_3d(ml, c0, tt, b);
      };
      // H_a reads its activation slabs for the first time (DRAM); H_b re-reads them from L2.
      // Run an L2 prefetch pf_dist slabs ahead of the ring so that the DRAM latency is not paid
      // once per ring round trip (measured: H_a 8-22 k cycles slower than H_b without it).
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
          if (P.x3) tma_load_3d(sa + A_PLANE, ml, fb, c0, tt, b);
          tma_load_2d(sa + 2 * A_PLANE, &map_w1_hi, fb, i * BK, gp * TN);
          if (P.x3) tma_load_2d(sa + 2 * A_PLANE + B_PLANE, &map_w1_lo, fb, i * BK, gp * TN);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
      for (int oc = o_begin; oc < o_end; ++oc) {
        for (int i = 0; i < nk2; ++i) {
          mbar_wait(empty0 + 8 * stage, ph ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          const uint32_t sa = base + stage * STAGE_BYTES;
          mbar_expect_tx(fb, nplanes * B_PLANE);
          tma_load_2d(sa + 2 * A_PLANE, &map_w2_hi, fb, i * BK, oc * TN);
          if (P.x3) tma_load_2d(sa + 2 * A_PLANE + B_PLANE, &map_w2_lo, fb, i * BK, oc * TN);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      int nphase = 0;
      const uint32_t acc = tmem_base + ACC_COL;
      const bool rec = rec_cta;
      const uint32_t idesc = idesc_for(IDESC, P.f16);
      for (int gp = 0; gp < 2; ++gp, ++nphase) {
        if (nphase > 0) {
          mbar_wait(acc_empty, (nphase - 1) & 1);
          tc_fence_after();
        }
        if (rec) P.dbg[2 * nphase] = clock64();
        for (int i = 0; i < nk1; ++i) {
          mbar_wait(full0 + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            const uint64_t a_hi = smem_desc_sw64(sa + ks * UK * 2);
            const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            mma_ss(acc, a_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
            if (P.x3) {
              const uint64_t a_lo = smem_desc_sw64(sa + A_PLANE + ks * UK * 2);
              const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + B_PLANE + ks * UK * 2);
              mma_ss(acc, a_lo, b_hi, idesc, 1u);
              mma_ss(acc, a_hi, b_lo, idesc, 1u);
            }
          }
          tc_commit(empty0 + 8 * stage);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        tc_commit(acc_full);
        if (rec) P.dbg[2 * nphase + 1] = clock64();
      }
      for (int oc = o_begin; oc < o_end; ++oc, ++nphase) {
        mbar_wait(acc_empty, (nphase - 1) & 1);   // accumulator drained AND z complete in TMEM
        tc_fence_after();
        if (rec) P.dbg[2 * nphase] = clock64();
        for (int i = 0; i < nk2; ++i) {
          mbar_wait(full0 + 8 * stage, ph);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            // z channel k sits in column k/2 of its plane: slab i, step ks -> 16*i + 8*ks
            const uint32_t z_hi = tmem_base + ZHI_COL + (BK / 2) * i + (UK / 2) * ks;
            const uint32_t z_lo = tmem_base + ZLO_COL + (BK / 2) * i + (UK / 2) * ks;
            const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            mma_ts(acc, z_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
            if (P.x3) {
              const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + B_PLANE + ks * UK * 2);
              mma_ts(acc, z_lo, b_hi, idesc, 1u);
              mma_ts(acc, z_hi, b_lo, idesc, 1u);
            }
          }
          tc_commit(empty0 + 8 * stage);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        tc_commit(acc_full);
        if (rec) P.dbg[2 * nphase + 1] = clock64();
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
    int nphase = 0;
    const bool rec = rec_cta && threadIdx.x == 0;
    // ---- gate phases: z = tanh(h_t) * sigmoid(h_s), kept in TMEM as bf16 hi/lo planes ----
    for (int gp = 0; gp < 2; ++gp, ++nphase) {
      mbar_wait(acc_full, nphase & 1);
      tc_fence_after();
      if (rec) P.dbg[16 + 2 * nphase] = clock64();
#pragma unroll 1
      for (int q = grp; q < HALF / 16; q += NG) {
        float a[16], g[16];
        tmem_ld16(lane_base + ACC_COL + 16 * q, a);
        tmem_ld16(lane_base + ACC_COL + HALF + 16 * q, g);
        uint32_t zh[8], zl[8];
        const int ch0 = gp * HALF + 16 * q;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float z2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ch = ch0 + i + u;
            const float th = tanh_fast(a[i + u] + b1s[ch]);
            const float sg = sigmoid_fast(g[i + u] + b1s[CH + ch]);
            z2[u] = th * sg;
            if (P.gate_tanh != nullptr && t_ok) {
              const int64_t off = ((int64_t)b * CH + ch) * P.T + t;
              __stcs(P.gate_tanh + off, th);
              __stcs(P.gate_sig + off, sg);
            }
          }
          if (P.x3) split_pair_f(z2[0], z2[1], zh[i >> 1], zl[i >> 1], P.f16);
          else zh[i >> 1] = pack_pair_f(z2[0], z2[1], P.f16);
        }
        tmem_st8(lane_base + ZHI_COL + (ch0 >> 1), zh);
        if (P.x3) tmem_st8(lane_base + ZLO_COL + (ch0 >> 1), zl);
        if (P.zp_hi != nullptr && t_ok) {
          const int64_t zoff = ((int64_t)b * P.T + t) * CH + ch0;
          st256(P.zp_hi + zoff, zh);
          if (P.x3) st256(P.zp_lo + zoff, zl);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(acc_empty);
      if (rec) P.dbg[16 + 2 * nphase + 1] = clock64();
    }
    // ---- output phases: residual chunks then skip chunks ----
    // residual = Wr z + br + x with x read back from the packed hi/lo planes (x = hi + lo to
    // 2^-17: two 16-byte loads per plane per 16 channels instead of 16 strided fp32 loads);
    // the running skip sum is fp32 (B,Cs,T): lanes are consecutive t, so every access is a
    // coalesced 128-byte row segment.  Operands of the NEXT chunk are fetched before the
    // current one is processed.
    for (int oc = o_begin; oc < o_end; ++oc, ++nphase) {
      const bool is_res = oc < P.Cr / TN;
      const int cbase = is_res ? oc * TN : (oc - P.Cr / TN) * TN;
      float addf[16];
      uint32_t hw[8], lw[8];
      auto fetch = [&](int q) {
        const int ch0 = cbase + 16 * q;
        if (is_res) {
          const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
          if (t_ok) {
            ld256(P.xp_hi + poff, hw);
            if (P.xlo) ld256(P.xp_lo + poff, lw);
          }
        } else if (P.skip_accumulate && t_ok) {
          const float* sp = P.skip + ((int64_t)b * P.Cs + ch0) * P.T + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) addf[i] = __ldcs(sp + (int64_t)i * P.T);
        }
      };
      fetch(grp);
      mbar_wait(acc_full, nphase & 1);
      tc_fence_after();
      if (rec) P.dbg[16 + 2 * nphase] = clock64();
#pragma unroll 1
      for (int q = grp; q < TN / 16; q += NG) {
        float o[16], add[16];
        tmem_ld16(lane_base + ACC_COL + 16 * q, o);
        if (is_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0, v1;
            unpack_pair_f(hw[i], P.f16, v0, v1);
            if (P.xlo) {
              float l0, l1;
              unpack_pair_f(lw[i], P.f16, l0, l1);
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
        if (q + NG < TN / 16) fetch(q + NG);
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
            if (P.xlo) split_pair_f(v2[0], v2[1], rh[i >> 1], rl[i >> 1], P.f16);
            else rh[i >> 1] = pack_pair_f(v2[0], v2[1], P.f16);
          }
          if (t_ok && P.res_hi != nullptr) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
            st256(P.res_hi + poff, rh);
            if (P.xlo) st256(P.res_lo + poff, rl);
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
      mbar_arrive(acc_empty);
      if (rec) P.dbg[16 + 2 * nphase + 1] = clock64();
    }
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


// ------------------------------------------------------------------ the kernel, v2 ---------
// Persistent variant (EXPERIMENTAL, VQW_TC_FWD_V2=1; measured SLOWER than v1 on B200: 1.02 vs
// 0.89 ms per block in bf16x3).  What it achieves is the epilogue overlap: its per-chunk timeline
// shows the gate / output epilogues fully hidden.  What it loses: an H chunk takes 33-35 k cycles
// (in bf16x3 AND in the single-pass mode) against 20.7 k / 6.9 k of MMA time -- ~620 cycles per
// K = 32 slab whatever the slab's size.  Ruled out by measurement: the issue rate of the TMA
// producer (three producer warps: same time), the L2 port (the CTA-pair version below stages 25 %
// fewer bytes per SM and is slower still), any saturated unit (ncu: tensor pipe 47 %, L2 48 %,
// xbar->SM 30 %, L2 hit rate 89 %).  What is left is latency: ~3.7 k cycles from "stage free" to
// "stage full" under this load, i.e. 6 stages x 32 KB in flight sustain one slab per 620 cycles, and
// feeding N = 128 chunks at the MMA rate would need ~300 KB in flight per SM.  v1 (48 KB per 768
// MMA cycles) needs ~230 KB and has 192: it is close to, but not at, the MMA rate in its first phase.
// One CTA per SM loops over (batch item, 128 time steps)
// tiles, and the work of a tile is cut into N = 128 output-channel CHUNKS that ping-pong between
// two 128-column TMEM accumulators, so the epilogue of chunk n runs while the tensor core works
// on chunk n+1 (v1 above has ONE 256-column accumulator: its 59 k cycles of epilogue per tile
// are all exposed, a third of the tile).  Chunks of a tile, in order:
//   H_0..H_3   h rows {tanh 64c..64c+63 | sigmoid 64c..64c+63} (K = fs*Cr + Cc) -> gate -> z
//              channels 64c..64c+63 into the TMEM z planes
//   R_0..R_3   residual channels 128r..128r+127 = Wr z + br + x      (A operand = z in TMEM)
//   S_0..S_1   skip channels 128s..128s+127 (+)= Ws z + bs
// The TMA producer streams K = 32 slabs for this chunk sequence without regard to tile
// boundaries (6-stage ring), so the first slabs of the next tile are in flight during the last
// epilogues of the current one.  TMEM: [0,128) acc 0, [128,256) acc 1, [256,384) z hi,
// [384,512) z lo.  Ordering: MMAs execute in issue order, so (a) waiting for the epilogue of
// H_3 before issuing R_0 guarantees every z column is written, and (b) the epilogue of the NEXT
// tile's H_0 -- which overwrites z -- can only start after that chunk's commit, i.e. after
// every R/S MMA of this tile has read z.
// v2 / v3 use THREE TMA producer warps ({A_hi, A_lo}, B_hi, B_lo): measured,
// one thread issues a tensor copy every ~170 cycles whatever its size, so with N = 128 chunks (384
// MMA cycles per K = 32 slab in bf16x3, 128 in the single-pass modes) a single producer thread
// issuing four copies per slab is the bottleneck (v2 measured 640 cycles per slab, v3 740).
constexpr int PW = 3;   // 20 warps in all: 96 registers per thread without spills
constexpr int FWD2_THREADS = (FWD_EPI_WARPS + PW + 1) * 32;
constexpr int W2_TMA0 = FWD_EPI_WARPS, W2_MMA = FWD_EPI_WARPS + PW;
constexpr int V2_STAGES = 6;
constexpr int V2_TN = 128;
constexpr int V2_B_PLANE = V2_TN * BK * 2;                  // 8 KB
constexpr int V2_STAGE_BYTES = 2 * A_PLANE + 2 * V2_B_PLANE;   // 32 KB
constexpr int V2_HALF = 64;                                 // gate pairs per H chunk
constexpr uint32_t IDESC_N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(V2_TN >> 3) << 17) |
                                ((uint32_t)(TM >> 4) << 24);

__global__ void __launch_bounds__(FWD2_THREADS, 1)
resblock_tc2_kernel(const __grid_constant__ CUtensorMap map_x_hi,
                    const __grid_constant__ CUtensorMap map_x_lo,
                    const __grid_constant__ CUtensorMap map_c_hi,
                    const __grid_constant__ CUtensorMap map_c_lo,
                    const __grid_constant__ CUtensorMap map_w1_hi,
                    const __grid_constant__ CUtensorMap map_w1_lo,
                    const __grid_constant__ CUtensorMap map_w2_hi,
                    const __grid_constant__ CUtensorMap map_w2_lo, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* b1s = reinterpret_cast<float*>(smem + V2_STAGES * V2_STAGE_BYTES);   // [512] conv_b + cond_b
  float* brs = b1s + CD;                                                      // [Cr]
  float* bss = brs + P.Cr;                                                    // [Cs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bss + P.Cs);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + V2_STAGES);
  const uint32_t acc_full0 = smem_u32(bars + 2 * V2_STAGES);        // [2]
  const uint32_t acc_empty0 = smem_u32(bars + 2 * V2_STAGES + 2);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * V2_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nplanes = P.x3 ? 2 : 1;
  const int chunks_per_tap = P.Cr / BK;
  const int nk1 = P.fs * chunks_per_tap + P.Cc / BK;     // K slabs of an H chunk
  const int nk2 = CH / BK;                               // K slabs of an R / S chunk
  const int n_h = CD / V2_TN;                            // 4
  const int n_r = P.Cr / V2_TN;
  const int o_begin = P.write_residual ? 0 : n_r;        // first chunk of [Wr ; Ws]
  const int o_end = n_r + P.Cs / V2_TN;
  const int tiles_per_item = (P.T + TM - 1) / TM;
  const int n_tiles = P.B * tiles_per_item;

  if (warp == W2_TMA0 && lane == 0) {
    prefetch_tmap(&map_x_hi); prefetch_tmap(&map_c_hi); prefetch_tmap(&map_w1_hi);
    prefetch_tmap(&map_w2_hi);
    if (P.x3) {
      prefetch_tmap(&map_x_lo); prefetch_tmap(&map_c_lo); prefetch_tmap(&map_w1_lo);
      prefetch_tmap(&map_w2_lo);
    }
    for (int s = 0; s < V2_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full0 + 8 * s, 1);
      mbar_init(acc_empty0 + 8 * s, FWD_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W2_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < FWD_EPI_WARPS) {
    for (int i = threadIdx.x; i < CD; i += FWD_EPI_WARPS * 32) b1s[i] = P.conv_b[i] + P.cond_b[i];
    for (int i = threadIdx.x; i < P.Cr; i += FWD_EPI_WARPS * 32) brs[i] = P.res_b[i];
    for (int i = threadIdx.x; i < P.Cs; i += FWD_EPI_WARPS * 32) bss[i] = P.skip_b[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= W2_TMA0 && warp < W2_MMA) {
    // =============================== TMA producers ==============================
    // producer warp 0 loads the activation planes, 1 the weight hi plane, 2 the weight lo plane; all
    // three walk the same stage ring, warp 0 also registers the stage's byte count
    const int pw = warp - W2_TMA0;
    if (lane == 0 && (pw < 2 || P.x3)) {
      int stage = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_item, t0 = (tile - b * tiles_per_item) * TM;
        for (int c = 0; c < n_h; ++c) {
          for (int i = 0; i < nk1; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t sa = base + stage * V2_STAGE_BYTES;
            if (pw == 0) mbar_expect_tx(fb, nplanes * (A_PLANE + V2_B_PLANE));
            const int tap = i / chunks_per_tap;
            if (tap < P.fs) {
              const int c0 = (i - tap * chunks_per_tap) * BK;
              const int tt = t0 - P.dilation * (P.fs - 1 - tap);   // negative rows -> zero fill
              if (pw == 0) tma_load_3d(sa, &map_x_hi, fb, c0, tt, b);
              if (pw == 0 && P.x3) tma_load_3d(sa + A_PLANE, &map_x_lo, fb, c0, tt, b);
            } else {
              const int c0 = (i - P.fs * chunks_per_tap) * BK;
              if (pw == 0) tma_load_3d(sa, &map_c_hi, fb, c0, t0, b);
              if (pw == 0 && P.x3) tma_load_3d(sa + A_PLANE, &map_c_lo, fb, c0, t0, b);
            }
            if (pw == 1) tma_load_2d(sa + 2 * A_PLANE, &map_w1_hi, fb, i * BK, c * V2_TN);
            if (pw == 2) tma_load_2d(sa + 2 * A_PLANE + V2_B_PLANE, &map_w1_lo, fb, i * BK, c * V2_TN);
            if (++stage == V2_STAGES) { stage = 0; ph ^= 1; }
          }
        }
        for (int oc = o_begin; oc < o_end; ++oc) {
          for (int i = 0; i < nk2; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t sa = base + stage * V2_STAGE_BYTES;
            if (pw == 0) mbar_expect_tx(fb, nplanes * V2_B_PLANE);
            if (pw == 1) tma_load_2d(sa + 2 * A_PLANE, &map_w2_hi, fb, i * BK, oc * V2_TN);
            if (pw == 2) tma_load_2d(sa + 2 * A_PLANE + V2_B_PLANE, &map_w2_lo, fb, i * BK, oc * V2_TN);
            if (++stage == V2_STAGES) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == W2_MMA) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      uint32_t n = 0;   // running chunk counter: accumulator n & 1, its k-th use k = n >> 1
      const uint32_t idesc = idesc_for(IDESC_N128, P.f16);
      const bool rec = P.dbg != nullptr && (int)blockIdx.x == P.dbg_x;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int c = 0; c < n_h; ++c, ++n) {
          const uint32_t buf = n & 1;
          mbar_wait(acc_empty0 + 8 * buf, ((n >> 1) & 1) ^ 1);   // epilogue of chunk n-2 is done
          tc_fence_after();
          if (rec && n < 24) P.dbg[4 * n] = clock64();
          const uint32_t acc = tmem_base + buf * V2_TN;
          for (int i = 0; i < nk1; ++i) {
            mbar_wait(full0 + 8 * stage, ph);
            tc_fence_after();
            const uint32_t sa = base + stage * V2_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < BK / UK; ++ks) {
              const uint64_t a_hi = smem_desc_sw64(sa + ks * UK * 2);
              const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
              mma_ss(acc, a_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
              if (P.x3) {
                const uint64_t a_lo = smem_desc_sw64(sa + A_PLANE + ks * UK * 2);
                const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + V2_B_PLANE + ks * UK * 2);
                mma_ss(acc, a_lo, b_hi, idesc, 1u);
                mma_ss(acc, a_hi, b_lo, idesc, 1u);
              }
            }
            tc_commit(empty0 + 8 * stage);
            if (++stage == V2_STAGES) { stage = 0; ph ^= 1; }
          }
          tc_commit(acc_full0 + 8 * buf);
          if (rec && n < 24) P.dbg[4 * n + 1] = clock64();
        }
        {   // every z column of this tile must be in TMEM: wait for the epilogue of H_3 (chunk n-1)
          const uint32_t m = n - 1;
          mbar_wait(acc_empty0 + 8 * (m & 1), (m >> 1) & 1);
          tc_fence_after();
        }
        for (int oc = o_begin; oc < o_end; ++oc, ++n) {
          const uint32_t buf = n & 1;
          mbar_wait(acc_empty0 + 8 * buf, ((n >> 1) & 1) ^ 1);
          tc_fence_after();
          if (rec && n < 24) P.dbg[4 * n] = clock64();
          const uint32_t acc = tmem_base + buf * V2_TN;
          for (int i = 0; i < nk2; ++i) {
            mbar_wait(full0 + 8 * stage, ph);
            tc_fence_after();
            const uint32_t sa = base + stage * V2_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < BK / UK; ++ks) {
              // z channel k sits in column k/2 of its plane: slab i, step ks -> 16*i + 8*ks
              const uint32_t z_hi = tmem_base + ZHI_COL + (BK / 2) * i + (UK / 2) * ks;
              const uint32_t z_lo = tmem_base + ZLO_COL + (BK / 2) * i + (UK / 2) * ks;
              const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
              mma_ts(acc, z_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
              if (P.x3) {
                const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + V2_B_PLANE + ks * UK * 2);
                mma_ts(acc, z_lo, b_hi, idesc, 1u);
                mma_ts(acc, z_hi, b_lo, idesc, 1u);
              }
            }
            tc_commit(empty0 + 8 * stage);
            if (++stage == V2_STAGES) { stage = 0; ph ^= 1; }
          }
          tc_commit(acc_full0 + 8 * buf);
          if (rec && n < 24) P.dbg[4 * n + 1] = clock64();
        }
      }
    }
  } else {
    // =============================== epilogue (warps 0-15) ======================
    // warp e: TMEM lane quadrant e%4, column group e/4 (16-column units dealt round-robin)
    const bool rec = P.dbg != nullptr && (int)blockIdx.x == P.dbg_x && threadIdx.x == 0;
    const int quad = warp & 3, grp = warp >> 2;
    constexpr int NG = FWD_EPI_WARPS / 4;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_item, t0 = (tile - b * tiles_per_item) * TM;
      const int t = t0 + row;
      const bool t_ok = t < P.T;
      // ---- gate chunks: z = tanh(h_t) * sigmoid(h_s) -> TMEM z planes (+ saved tensors) ----
      for (int c = 0; c < n_h; ++c, ++n) {
        const uint32_t buf = n & 1;
        const uint32_t acc = lane_base + buf * V2_TN;
        mbar_wait(acc_full0 + 8 * buf, (n >> 1) & 1);
        tc_fence_after();
        if (rec && n < 24) P.dbg[4 * n + 2] = clock64();
#pragma unroll 1
        for (int q = grp; q < V2_HALF / 16; q += NG) {
          float a[16], g[16];
          tmem_ld16(acc + 16 * q, a);
          tmem_ld16(acc + V2_HALF + 16 * q, g);
          uint32_t zh[8], zl[8];
          const int ch0 = c * V2_HALF + 16 * q;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float z2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ch = ch0 + i + u;
              const float th = tanh_fast(a[i + u] + b1s[ch]);
              const float sg = sigmoid_fast(g[i + u] + b1s[CH + ch]);
              z2[u] = th * sg;
              if (P.gate_tanh != nullptr && t_ok) {
                const int64_t off = ((int64_t)b * CH + ch) * P.T + t;
                __stcs(P.gate_tanh + off, th);
                __stcs(P.gate_sig + off, sg);
              }
            }
            if (P.x3) split_pair_f(z2[0], z2[1], zh[i >> 1], zl[i >> 1], P.f16);
            else zh[i >> 1] = pack_pair_f(z2[0], z2[1], P.f16);
          }
          tmem_st8(lane_base + ZHI_COL + (ch0 >> 1), zh);
          if (P.x3) tmem_st8(lane_base + ZLO_COL + (ch0 >> 1), zl);
          if (P.zp_hi != nullptr && t_ok) {
            const int64_t zoff = ((int64_t)b * P.T + t) * CH + ch0;
            st256(P.zp_hi + zoff, zh);
            if (P.x3) st256(P.zp_lo + zoff, zl);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(acc_empty0 + 8 * buf);
        if (rec && n < 24) P.dbg[4 * n + 3] = clock64();
      }
      // ---- output chunks: residual channels, then skip channels ----
      for (int oc = o_begin; oc < o_end; ++oc, ++n) {
        const uint32_t buf = n & 1;
        const uint32_t acc = lane_base + buf * V2_TN;
        const bool is_res = oc < n_r;
        const int cbase = is_res ? oc * V2_TN : (oc - n_r) * V2_TN;
        float addf[16];
        uint32_t hw[8], lw[8];
        auto fetch = [&](int q) {
          const int ch0 = cbase + 16 * q;
          if (is_res) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
            if (t_ok) {
              ld256(P.xp_hi + poff, hw);
              if (P.xlo) ld256(P.xp_lo + poff, lw);
            }
          } else if (P.skip_accumulate && t_ok) {
            const float* sp = P.skip + ((int64_t)b * P.Cs + ch0) * P.T + t;
#pragma unroll
            for (int i = 0; i < 16; ++i) addf[i] = __ldcs(sp + (int64_t)i * P.T);
          }
        };
        fetch(grp);
        mbar_wait(acc_full0 + 8 * buf, (n >> 1) & 1);
        tc_fence_after();
        if (rec && n < 24) P.dbg[4 * n + 2] = clock64();
#pragma unroll 1
        for (int q = grp; q < V2_TN / 16; q += NG) {
          float o[16], add[16];
          tmem_ld16(acc + 16 * q, o);
          if (is_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v0, v1;
              unpack_pair_f(hw[i], P.f16, v0, v1);
              if (P.xlo) {
                float l0, l1;
                unpack_pair_f(lw[i], P.f16, l0, l1);
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
          if (q + NG < V2_TN / 16) fetch(q + NG);
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
              if (P.xlo) split_pair_f(v2[0], v2[1], rh[i >> 1], rl[i >> 1], P.f16);
              else rh[i >> 1] = pack_pair_f(v2[0], v2[1], P.f16);
            }
            if (t_ok && P.res_hi != nullptr) {
              const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
              st256(P.res_hi + poff, rh);
              if (P.xlo) st256(P.res_lo + poff, rl);
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
        mbar_arrive(acc_empty0 + 8 * buf);
        if (rec && n < 24) P.dbg[4 * n + 3] = clock64();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W2_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// ------------------------------------------------------------------ the kernel, v3 ---------
// The chunk pipeline of v2 on a CTA PAIR (EXPERIMENTAL, VQW_TC_FWD_V3=1; parity green, measured
// 1.08 ms per block): a cluster of two CTAs covers 256 time steps, every MMA is a cta_group::2
// instruction with M = 256 (128 rows per SM) and N = 128, and each SM stages only HALF of a weight
// slab (64 rows): 24 KB per slab and SM instead of 32.  One thread of the leader CTA issues the
// MMAs of both SMs; the TMA copies of both CTAs complete on the leader's "full" barrier; "empty"
// and "accumulator full" are tcgen05.commit multicasts to both CTAs; one elected lane per epilogue
// warp of both CTAs arrives on the leader's "accumulator empty" barriers.  It validates the whole
// 2-SM protocol, but the slab interval is latency bound like v2's (740 cycles with 8 stages), so
// the next step is the pair WITHOUT the N = 128 chunking: v1's N = 256 phases with the weight slab
// split over the pair need 32 KB per 768 MMA cycles, ~150 KB in flight, which fits.
constexpr int V3_STAGES = 8;
constexpr int V3_B_PLANE = (V2_TN / 2) * BK * 2;               // 4 KB: this CTA's 64 weight rows
constexpr int V3_STAGE_BYTES = 2 * A_PLANE + 2 * V3_B_PLANE;   // 24 KB
constexpr uint32_t IDESC_M256_N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(V2_TN >> 3) << 17) |
                                     ((uint32_t)((2 * TM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                             int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far are done
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          bar),
      "h"((unsigned short)3)
      : "memory");
}

__global__ void __launch_bounds__(FWD2_THREADS, 1)
resblock_tc3_kernel(const __grid_constant__ CUtensorMap map_x_hi,
                    const __grid_constant__ CUtensorMap map_x_lo,
                    const __grid_constant__ CUtensorMap map_c_hi,
                    const __grid_constant__ CUtensorMap map_c_lo,
                    const __grid_constant__ CUtensorMap map_w1_hi,
                    const __grid_constant__ CUtensorMap map_w1_lo,
                    const __grid_constant__ CUtensorMap map_w2_hi,
                    const __grid_constant__ CUtensorMap map_w2_lo, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* b1s = reinterpret_cast<float*>(smem + V3_STAGES * V3_STAGE_BYTES);   // [512] conv_b + cond_b
  float* brs = b1s + CD;                                                      // [Cr]
  float* bss = brs + P.Cr;                                                    // [Cs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bss + P.Cs);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + V3_STAGES);
  const uint32_t acc_full0 = smem_u32(bars + 2 * V3_STAGES);        // [2]
  const uint32_t acc_empty0 = smem_u32(bars + 2 * V3_STAGES + 2);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * V3_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nplanes = P.x3 ? 2 : 1;
  const uint32_t rank = cluster_ctarank();             // 0 = leader of the CTA pair
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int chunks_per_tap = P.Cr / BK;
  const int nk1 = P.fs * chunks_per_tap + P.Cc / BK;     // K slabs of an H chunk
  const int nk2 = CH / BK;                               // K slabs of an R / S chunk
  const int n_h = CD / V2_TN;                            // 4
  const int n_r = P.Cr / V2_TN;
  const int o_begin = P.write_residual ? 0 : n_r;        // first chunk of [Wr ; Ws]
  const int o_end = n_r + P.Cs / V2_TN;
  const int tiles_per_item = (P.T + 2 * TM - 1) / (2 * TM);   // a pair covers 256 time steps
  const int n_tiles = P.B * tiles_per_item;

  if (warp == W2_TMA0 && lane == 0) {
    prefetch_tmap(&map_x_hi); prefetch_tmap(&map_c_hi); prefetch_tmap(&map_w1_hi);
    prefetch_tmap(&map_w2_hi);
    if (P.x3) {
      prefetch_tmap(&map_x_lo); prefetch_tmap(&map_c_lo); prefetch_tmap(&map_w1_lo);
      prefetch_tmap(&map_w2_lo);
    }
    for (int s = 0; s < V3_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full0 + 8 * s, 1);
      mbar_init(acc_empty0 + 8 * s, 2 * FWD_EPI_WARPS);   // one elected lane per epilogue warp, both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W2_MMA) {   // the same warp in BOTH CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp < FWD_EPI_WARPS) {
    for (int i = threadIdx.x; i < CD; i += FWD_EPI_WARPS * 32) b1s[i] = P.conv_b[i] + P.cond_b[i];
    for (int i = threadIdx.x; i < P.Cr; i += FWD_EPI_WARPS * 32) brs[i] = P.res_b[i];
    for (int i = threadIdx.x; i < P.Cs; i += FWD_EPI_WARPS * 32) bss[i] = P.skip_b[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= W2_TMA0 && warp < W2_MMA) {
    // =============================== TMA producers ==============================
    // producer warp 0 loads the activation planes, 1 the weight hi plane, 2 the weight lo plane; all
    // three walk the same stage ring, warp 0 also registers the stage's byte count
    const int pw = warp - W2_TMA0;
    if (lane == 0 && (pw < 2 || P.x3)) {
      int stage = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int b = tile / tiles_per_item, t0 = (tile - b * tiles_per_item) * 2 * TM + (int)rank * TM;
        for (int c = 0; c < n_h; ++c) {
          for (int i = 0; i < nk1; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            // both CTAs' copies complete on the LEADER's barrier, which expects the bytes of both
            const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
            const uint32_t sa = base + stage * V3_STAGE_BYTES;
            if (rank == 0 && pw == 0) mbar_expect_tx(full0 + 8 * stage, 2 * nplanes * (A_PLANE + V3_B_PLANE));
            const int tap = i / chunks_per_tap;
            if (tap < P.fs) {
              const int c0 = (i - tap * chunks_per_tap) * BK;
              const int tt = t0 - P.dilation * (P.fs - 1 - tap);   // negative rows -> zero fill
              if (pw == 0) tma2_load_3d(sa, &map_x_hi, fb, c0, tt, b);
              if (pw == 0 && P.x3) tma2_load_3d(sa + A_PLANE, &map_x_lo, fb, c0, tt, b);
            } else {
              const int c0 = (i - P.fs * chunks_per_tap) * BK;
              if (pw == 0) tma2_load_3d(sa, &map_c_hi, fb, c0, t0, b);
              if (pw == 0 && P.x3) tma2_load_3d(sa + A_PLANE, &map_c_lo, fb, c0, t0, b);
            }
            // this CTA stages its half (64 rows) of the chunk's 128 weight rows
            if (pw == 1)
              tma2_load_2d(sa + 2 * A_PLANE, &map_w1_hi, fb, i * BK, c * V2_TN + (int)rank * (V2_TN / 2));
            if (pw == 2)
              tma2_load_2d(sa + 2 * A_PLANE + V3_B_PLANE, &map_w1_lo, fb, i * BK,
                           c * V2_TN + (int)rank * (V2_TN / 2));
            if (++stage == V3_STAGES) { stage = 0; ph ^= 1; }
          }
        }
        for (int oc = o_begin; oc < o_end; ++oc) {
          for (int i = 0; i < nk2; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = mapa_u32(full0 + 8 * stage, 0);
            const uint32_t sa = base + stage * V3_STAGE_BYTES;
            if (rank == 0 && pw == 0) mbar_expect_tx(full0 + 8 * stage, 2 * nplanes * V3_B_PLANE);
            if (pw == 1)
              tma2_load_2d(sa + 2 * A_PLANE, &map_w2_hi, fb, i * BK, oc * V2_TN + (int)rank * (V2_TN / 2));
            if (pw == 2)
              tma2_load_2d(sa + 2 * A_PLANE + V3_B_PLANE, &map_w2_lo, fb, i * BK,
                           oc * V2_TN + (int)rank * (V2_TN / 2));
            if (++stage == V3_STAGES) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == W2_MMA) {
    // =============================== MMA issuer =================================
    if (lane == 0 && rank == 0) {   // ONE thread of the pair issues the M = 256 MMAs of both SMs
      int stage = 0;
      uint32_t ph = 0;
      uint32_t n = 0;   // running chunk counter: accumulator n & 1, its k-th use k = n >> 1
      const uint32_t idesc = idesc_for(IDESC_M256_N128, P.f16);
      const bool rec = P.dbg != nullptr && (int)blockIdx.x == P.dbg_x;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        for (int c = 0; c < n_h; ++c, ++n) {
          const uint32_t buf = n & 1;
          mbar_wait(acc_empty0 + 8 * buf, ((n >> 1) & 1) ^ 1);   // epilogue of chunk n-2 is done
          tc_fence_after();
          if (rec && n < 24) P.dbg[4 * n] = clock64();
          const uint32_t acc = tmem_base + buf * V2_TN;
          for (int i = 0; i < nk1; ++i) {
            mbar_wait(full0 + 8 * stage, ph);
            tc_fence_after();
            const uint32_t sa = base + stage * V3_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < BK / UK; ++ks) {
              const uint64_t a_hi = smem_desc_sw64(sa + ks * UK * 2);
              const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
              mma2_ss(acc, a_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
              if (P.x3) {
                const uint64_t a_lo = smem_desc_sw64(sa + A_PLANE + ks * UK * 2);
                const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + V3_B_PLANE + ks * UK * 2);
                mma2_ss(acc, a_lo, b_hi, idesc, 1u);
                mma2_ss(acc, a_hi, b_lo, idesc, 1u);
              }
            }
            tc_commit2(empty0 + 8 * stage);
            if (++stage == V3_STAGES) { stage = 0; ph ^= 1; }
          }
          tc_commit2(acc_full0 + 8 * buf);
          if (rec && n < 24) P.dbg[4 * n + 1] = clock64();
        }
        {   // every z column of this tile must be in TMEM: wait for the epilogue of H_3 (chunk n-1)
          const uint32_t m = n - 1;
          mbar_wait(acc_empty0 + 8 * (m & 1), (m >> 1) & 1);
          tc_fence_after();
        }
        for (int oc = o_begin; oc < o_end; ++oc, ++n) {
          const uint32_t buf = n & 1;
          mbar_wait(acc_empty0 + 8 * buf, ((n >> 1) & 1) ^ 1);
          tc_fence_after();
          if (rec && n < 24) P.dbg[4 * n] = clock64();
          const uint32_t acc = tmem_base + buf * V2_TN;
          for (int i = 0; i < nk2; ++i) {
            mbar_wait(full0 + 8 * stage, ph);
            tc_fence_after();
            const uint32_t sa = base + stage * V3_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < BK / UK; ++ks) {
              // z channel k sits in column k/2 of its plane: slab i, step ks -> 16*i + 8*ks
              const uint32_t z_hi = tmem_base + ZHI_COL + (BK / 2) * i + (UK / 2) * ks;
              const uint32_t z_lo = tmem_base + ZLO_COL + (BK / 2) * i + (UK / 2) * ks;
              const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
              mma2_ts(acc, z_hi, b_hi, idesc, (i | ks) ? 1u : 0u);
              if (P.x3) {
                const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + V3_B_PLANE + ks * UK * 2);
                mma2_ts(acc, z_lo, b_hi, idesc, 1u);
                mma2_ts(acc, z_hi, b_lo, idesc, 1u);
              }
            }
            tc_commit2(empty0 + 8 * stage);
            if (++stage == V3_STAGES) { stage = 0; ph ^= 1; }
          }
          tc_commit2(acc_full0 + 8 * buf);
          if (rec && n < 24) P.dbg[4 * n + 1] = clock64();
        }
      }
    }
  } else {
    // =============================== epilogue (warps 0-15) ======================
    // warp e: TMEM lane quadrant e%4, column group e/4 (16-column units dealt round-robin)
    const bool rec = P.dbg != nullptr && (int)blockIdx.x == P.dbg_x && threadIdx.x == 0;
    const int quad = warp & 3, grp = warp >> 2;
    constexpr int NG = FWD_EPI_WARPS / 4;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t n = 0;
    const uint32_t acc_empty_leader0 = mapa_u32(acc_empty0, 0);   // the leader's barriers
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int b = tile / tiles_per_item, t0 = (tile - b * tiles_per_item) * 2 * TM + (int)rank * TM;
      const int t = t0 + row;
      const bool t_ok = t < P.T;
      // ---- gate chunks: z = tanh(h_t) * sigmoid(h_s) -> TMEM z planes (+ saved tensors) ----
      for (int c = 0; c < n_h; ++c, ++n) {
        const uint32_t buf = n & 1;
        const uint32_t acc = lane_base + buf * V2_TN;
        mbar_wait(acc_full0 + 8 * buf, (n >> 1) & 1);
        tc_fence_after();
        if (rec && n < 24) P.dbg[4 * n + 2] = clock64();
#pragma unroll 1
        for (int q = grp; q < V2_HALF / 16; q += NG) {
          float a[16], g[16];
          tmem_ld16(acc + 16 * q, a);
          tmem_ld16(acc + V2_HALF + 16 * q, g);
          uint32_t zh[8], zl[8];
          const int ch0 = c * V2_HALF + 16 * q;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float z2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ch = ch0 + i + u;
              const float th = tanh_fast(a[i + u] + b1s[ch]);
              const float sg = sigmoid_fast(g[i + u] + b1s[CH + ch]);
              z2[u] = th * sg;
              if (P.gate_tanh != nullptr && t_ok) {
                const int64_t off = ((int64_t)b * CH + ch) * P.T + t;
                __stcs(P.gate_tanh + off, th);
                __stcs(P.gate_sig + off, sg);
              }
            }
            if (P.x3) split_pair_f(z2[0], z2[1], zh[i >> 1], zl[i >> 1], P.f16);
            else zh[i >> 1] = pack_pair_f(z2[0], z2[1], P.f16);
          }
          tmem_st8(lane_base + ZHI_COL + (ch0 >> 1), zh);
          if (P.x3) tmem_st8(lane_base + ZLO_COL + (ch0 >> 1), zl);
          if (P.zp_hi != nullptr && t_ok) {
            const int64_t zoff = ((int64_t)b * P.T + t) * CH + ch0;
            st256(P.zp_hi + zoff, zh);
            if (P.x3) st256(P.zp_lo + zoff, zl);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader0 + 8 * buf);
        if (rec && n < 24) P.dbg[4 * n + 3] = clock64();
      }
      // ---- output chunks: residual channels, then skip channels ----
      for (int oc = o_begin; oc < o_end; ++oc, ++n) {
        const uint32_t buf = n & 1;
        const uint32_t acc = lane_base + buf * V2_TN;
        const bool is_res = oc < n_r;
        const int cbase = is_res ? oc * V2_TN : (oc - n_r) * V2_TN;
        float addf[16];
        uint32_t hw[8], lw[8];
        auto fetch = [&](int q) {
          const int ch0 = cbase + 16 * q;
          if (is_res) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
            if (t_ok) {
              ld256(P.xp_hi + poff, hw);
              if (P.xlo) ld256(P.xp_lo + poff, lw);
            }
          } else if (P.skip_accumulate && t_ok) {
            const float* sp = P.skip + ((int64_t)b * P.Cs + ch0) * P.T + t;
#pragma unroll
            for (int i = 0; i < 16; ++i) addf[i] = __ldcs(sp + (int64_t)i * P.T);
          }
        };
        fetch(grp);
        mbar_wait(acc_full0 + 8 * buf, (n >> 1) & 1);
        tc_fence_after();
        if (rec && n < 24) P.dbg[4 * n + 2] = clock64();
#pragma unroll 1
        for (int q = grp; q < V2_TN / 16; q += NG) {
          float o[16], add[16];
          tmem_ld16(acc + 16 * q, o);
          if (is_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v0, v1;
              unpack_pair_f(hw[i], P.f16, v0, v1);
              if (P.xlo) {
                float l0, l1;
                unpack_pair_f(lw[i], P.f16, l0, l1);
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
          if (q + NG < V2_TN / 16) fetch(q + NG);
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
              if (P.xlo) split_pair_f(v2[0], v2[1], rh[i >> 1], rl[i >> 1], P.f16);
              else rh[i >> 1] = pack_pair_f(v2[0], v2[1], P.f16);
            }
            if (t_ok && P.res_hi != nullptr) {
              const int64_t poff = ((int64_t)b * P.T + t) * P.Cr + ch0;
              st256(P.res_hi + poff, rh);
              if (P.xlo) st256(P.res_lo + poff, rl);
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
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader0 + 8 * buf);
        if (rec && n < 24) P.dbg[4 * n + 3] = clock64();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair's MMAs can still touch it
  if (warp == W2_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}


// ------------------------------------------------------------------ packing kernels --------
// (B,C,T) fp32 -> (B,T,C) bf16 hi/lo planes: 32x32 transpose through shared memory
__global__ void __launch_bounds__(256)
pack_act_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                __nv_bfloat16* __restrict__ lo, int C, int T, int pitch, int relu, int f16,
                const float* __restrict__ scale) {
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
    if (t < T && c < pitch) {          // channels in [C, pitch) are zero padding
      __nv_bfloat16 h, l;
      float v = tile[tx][r];
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
  pack_act_kernel<<<g, 256, 0, stream>>>(in, hi, lo, C, T, pitch, relu, f16, scale);
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
__global__ void __launch_bounds__(256)
pack_w1_kernel(const float* __restrict__ conv_w, const float* __restrict__ cond_w,
               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cr, int Cc,
               int fs, int f16, int half) {
  const int K1 = fs * Cr + Cc;
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
__global__ void __launch_bounds__(256)
pack_w2_kernel(const float* __restrict__ res_w, const float* __restrict__ skip_w,
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

size_t smem_bytes(int Cr, int Cs) {
  return 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(float) * (CD + Cr + Cs) + 8 * (2 * STAGES + 2) + 16;
}
size_t smem_bytes_v3(int Cr, int Cs) {
  return 1024 + (size_t)V3_STAGES * V3_STAGE_BYTES + sizeof(float) * (CD + Cr + Cs) +
         8 * (2 * V3_STAGES + 4) + 16;
}
size_t smem_bytes_v2(int Cr, int Cs) {
  return 1024 + (size_t)V2_STAGES * V2_STAGE_BYTES + sizeof(float) * (CD + Cr + Cs) +
         8 * (2 * V2_STAGES + 4) + 16;
}

}  // namespace tc

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

bool resnet_tc_supported(const vqw_resnet_desc& d) {
  return d.Cd == tc::CD && d.Cr % tc::TN == 0 && d.Cs % tc::TN == 0 && d.Cc % tc::BK == 0 &&
         d.Cr >= tc::TN && d.Cs >= tc::TN && d.fs >= 1 && d.T >= tc::TM && d.T % 8 == 0 &&
         d.Cc % 16 == 0;
}

// workspace: [cond hi|lo] [x ping hi|lo] [x pong hi|lo] [per block: w1 hi|lo, w2 hi|lo]
struct TcWorkspace {
  int64_t cond_plane, x_plane, w1_plane, w2_plane, block_stride, total;
  int64_t off_cond, off_x[2], off_w;
};
static TcWorkspace tc_layout(const vqw_resnet_desc& d) {
  TcWorkspace w;
  const int K1 = d.fs * d.Cr + d.Cc;
  w.cond_plane = align_up((int64_t)d.B * d.T * d.Cc * 2, 1024);
  w.x_plane = align_up((int64_t)d.B * d.T * d.Cr * 2, 1024);
  w.w1_plane = align_up((int64_t)tc::CD * K1 * 2, 1024);
  w.w2_plane = align_up((int64_t)(d.Cr + d.Cs) * tc::CH * 2, 1024);
  w.block_stride = 2 * w.w1_plane + 2 * w.w2_plane;
  w.off_cond = 0;
  w.off_x[0] = 2 * w.cond_plane;
  w.off_x[1] = w.off_x[0] + 2 * w.x_plane;
  w.off_w = w.off_x[1] + 2 * w.x_plane;
  w.total = w.off_w + (int64_t)d.n_blocks * w.block_stride;
  return w;
}

int64_t resnet_tc_workspace(const vqw_resnet_desc& d) { return tc_layout(d).total + 1024; }
int64_t resnet_tc_saved_bytes(const vqw_resnet_desc& d) { return tc_saved_layout(d).total; }

int resnet_forward_tc(const vqw_resnet_desc& d, const float* x, const float* cond,
                      const vqw_resblock_weights* weights, float* const* residuals, float* skip,
                      float* const* gate_tanh, float* const* gate_sig, void* workspace,
                      void* saved, cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(resnet_tc_supported(d),
              "tcgen05 path needs dilated_channels=512, residual/skip channels multiples of 256, "
              "condition channels a multiple of 32, T >= 128 and T %% 8 == 0 "
              "(got Cr=%d Cd=%d Cs=%d Cc=%d T=%d)", d.Cr, d.Cd, d.Cs, d.Cc, d.T);
  VQW_REQUIRE(workspace != nullptr, "vqw_resnet_forward: workspace is null");
  VQW_REQUIRE(d.B <= 65535, "vqw_resnet_forward: B > 65535");
  const bool x3 = d.mode == VQW_MODE_BF16X3;
  const int f16 = d.mode == VQW_MODE_FP16 ? 1 : 0;
  const bool xlo = x3 || f16;   // residual stream hi + lo (only the hi plane feeds the MMAs in fp16)
  // VQW_TC_FWD_V2=1 selects the experimental persistent kernel (see resblock_tc2_kernel)
  const char* v2env = getenv("VQW_TC_FWD_V2");
  const bool use_v2 = v2env && v2env[0] == '1';
  // VQW_TC_FWD_V3=1: the same chunk pipeline on CTA pairs (cta_group::2), see resblock_tc3_kernel
  const char* v3env = getenv("VQW_TC_FWD_V3");
  const bool v3 = v3env && v3env[0] == '1' && d.Cr % V2_TN == 0 && d.Cs % V2_TN == 0;
  const bool v2 = (use_v2 || v3) && d.Cr % V2_TN == 0 && d.Cs % V2_TN == 0;   // v3 shares v2's packing
  // weight rows per TMA box: a whole accumulator chunk, or this CTA's half of it (v3)
  const int wrows = v3 ? V2_TN / 2 : (v2 ? V2_TN : TN);
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
  const int K1 = d.fs * d.Cr + d.Cc;

  // pack the two inputs and every block's weights
  {
    dim3 g1(ceil_div(d.T, 32), ceil_div(d.Cr, 32), d.B), g2(ceil_div(d.T, 32), ceil_div(d.Cc, 32), d.B);
    pack_act_kernel<<<g1, 256, 0, stream>>>(x, x_hi[0], xlo ? x_lo[0] : nullptr, d.Cr, d.T, d.Cr, 0,
                                            f16, nullptr);
    VQW_CHECK_LAUNCH("pack_act_kernel(x)");
    pack_act_kernel<<<g2, 256, 0, stream>>>(cond, c_hi, x3 ? c_lo : nullptr, d.Cc, d.T, d.Cc, 0, f16,
                                            nullptr);
    VQW_CHECK_LAUNCH("pack_act_kernel(cond)");
    for (int i = 0; i < d.n_blocks; ++i) {
      const vqw_resblock_weights& w = weights[i];
      VQW_REQUIRE(w.conv_w && w.conv_b && w.cond_w && w.cond_b && w.res_w && w.res_b && w.skip_w &&
                      w.skip_b, "vqw_resnet_forward: block %d has a null weight", i);
      __nv_bfloat16* w1h = plane(L.off_w + i * L.block_stride);
      __nv_bfloat16* w1l = plane(L.off_w + i * L.block_stride + L.w1_plane);
      __nv_bfloat16* w2h = plane(L.off_w + i * L.block_stride + 2 * L.w1_plane);
      __nv_bfloat16* w2l = plane(L.off_w + i * L.block_stride + 2 * L.w1_plane + L.w2_plane);
      pack_w1_kernel<<<296, 256, 0, stream>>>(w.conv_w, w.cond_w, w1h, x3 ? w1l : nullptr, d.Cr,
                                               d.Cc, d.fs, f16, v2 ? V2_HALF : HALF);
      VQW_CHECK_LAUNCH("pack_w1_kernel");
      pack_w2_kernel<<<148, 256, 0, stream>>>(w.res_w, w.skip_w, w2h, x3 ? w2l : nullptr, d.Cr, d.Cs,
                                               f16);
      VQW_CHECK_LAUNCH("pack_w2_kernel");
    }
  }

  const size_t smem = v3 ? smem_bytes_v3(d.Cr, d.Cs) : v2 ? smem_bytes_v2(d.Cr, d.Cs) : smem_bytes(d.Cr, d.Cs);
  if (v3)
    VQW_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc3_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (v2)
    VQW_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc2_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    VQW_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int n_sm = 148;
  {
    int dev = 0;
    VQW_CHECK_CUDA(cudaGetDevice(&dev));
    VQW_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap m_c_hi, m_c_lo;
  if (int rc = make_map(&m_c_hi, c_hi, 3, d.Cc, d.T, d.B, TM)) return rc;
  if (int rc = make_map(&m_c_lo, x3 ? c_lo : c_hi, 3, d.Cc, d.T, d.B, TM)) return rc;

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
    if (int rc = make_map(&m_w2_hi, w2h, 2, CH, d.Cr + d.Cs, 1, wrows)) return rc;
    if (int rc = make_map(&m_w2_lo, x3 ? w2l : w2h, 2, CH, d.Cr + d.Cs, 1, wrows)) return rc;
    Params P;
    P.B = d.B; P.T = d.T; P.Cr = d.Cr; P.Cs = d.Cs; P.Cc = d.Cc; P.fs = d.fs;
    P.dilation = d.dilations[i];
    P.x3 = x3 ? 1 : 0;
    P.f16 = f16;
    P.xlo = xlo ? 1 : 0;
    P.skip_accumulate = i > 0;
    P.write_residual = write_res ? 1 : 0;
    {
      static const int pf_env = getenv("VQW_TC_PREFETCH") ? atoi(getenv("VQW_TC_PREFETCH")) : 12;
      P.pf_dist = pf_env;
    }
    P.xp_hi = x_hi[cur];
    P.xp_lo = x_lo[cur];
    P.conv_b = w.conv_b; P.cond_b = w.cond_b; P.res_b = w.res_b; P.skip_b = w.skip_b;
    P.res_f32 = (write_res && residuals) ? residuals[i] : nullptr;
    P.res_hi = last ? nullptr : x_hi[nxt];
    P.res_lo = last ? nullptr : x_lo[nxt];
    P.skip = skip;
    P.gate_tanh = gate_tanh ? gate_tanh[i] : nullptr;
    P.gate_sig = gate_sig ? gate_sig[i] : nullptr;
    P.zp_hi = sv ? splane(S.z0 + i * S.z_stride) : nullptr;
    P.zp_lo = sv ? splane(S.z0 + i * S.z_stride + S.z_plane) : nullptr;
    VQW_REQUIRE((P.gate_tanh == nullptr) == (P.gate_sig == nullptr),
                "vqw_resnet_forward: gate_tanh/gate_sig of block %d must be given together", i);
    static long long* dbg_buf = nullptr;
    static const bool timeline = getenv("VQW_TC_TIMELINE") && getenv("VQW_TC_TIMELINE")[0] == '1';
    P.dbg = nullptr; P.dbg_x = P.dbg_y = 0;
    if (timeline) {
      if (!dbg_buf) cudaMalloc(&dbg_buf, 128 * sizeof(long long));
      cudaMemsetAsync(dbg_buf, 0, 128 * sizeof(long long), stream);
      P.dbg = dbg_buf;
      P.dbg_x = getenv("VQW_TC_TIMELINE_X") ? atoi(getenv("VQW_TC_TIMELINE_X")) : 1;
      P.dbg_y = getenv("VQW_TC_TIMELINE_Y") ? atoi(getenv("VQW_TC_TIMELINE_Y")) : 0;
    }
    if (v3) {
      const int n_pair_tiles = d.B * ceil_div(d.T, 2 * TM);
      const int n_pairs = n_pair_tiles < n_sm / 2 ? n_pair_tiles : n_sm / 2;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * n_pairs);
      cfg.blockDim = dim3(FWD2_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      VQW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, resblock_tc3_kernel, m_x_hi, m_x_lo, m_c_hi, m_c_lo, m_w1_hi,
                                        m_w1_lo, m_w2_hi, m_w2_lo, P));
      VQW_CHECK_LAUNCH("resblock_tc3_kernel");
    } else if (v2) {
      const int n_tiles = d.B * ceil_div(d.T, TM);
      resblock_tc2_kernel<<<n_tiles < n_sm ? n_tiles : n_sm, FWD2_THREADS, smem, stream>>>(
          m_x_hi, m_x_lo, m_c_hi, m_c_lo, m_w1_hi, m_w1_lo, m_w2_hi, m_w2_lo, P);
      VQW_CHECK_LAUNCH("resblock_tc2_kernel");
    } else {
      dim3 grid(ceil_div(d.T, TM), d.B);
      resblock_tc_kernel<<<grid, FWD_THREADS, smem, stream>>>(m_x_hi, m_x_lo, m_c_hi, m_c_lo, m_w1_hi,
                                                          m_w1_lo, m_w2_hi, m_w2_lo, P);
      VQW_CHECK_LAUNCH("resblock_tc_kernel");
    }
    if (timeline && v2) {
      long long h[128];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[vqw timeline v2] block %d CTA %d (cycles rel. to the first MMA issue)\n", i, P.dbg_x);
      for (int n = 0; n < 24; ++n)
        fprintf(stderr, "  chunk %2d: mma issue [%7lld, %7lld]  epilogue [%7lld, %7lld]\n", n,
                h[4 * n] - h[0], h[4 * n + 1] - h[0], h[4 * n + 2] - h[0], h[4 * n + 3] - h[0]);
    }
    if (timeline && !v2) {
      long long h[64];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[vqw timeline] block %d CTA (%d,%d): entry %lld, setup done %lld, exit %lld "
                      "(cycles rel. to first MMA phase start)\n", i, P.dbg_x, P.dbg_y, h[40] - h[0],
              h[41] - h[0], h[42] - h[0]);
      for (int n = 0; n < 5; ++n)
        fprintf(stderr, "  phase %d: mma [%lld, %lld]  epilogue [%lld, %lld]\n", n, h[2 * n] - h[0],
                h[2 * n + 1] - h[0], h[16 + 2 * n] - h[0], h[16 + 2 * n + 1] - h[0]);
    }
  }
  return 0;
}

}  // namespace vqw
