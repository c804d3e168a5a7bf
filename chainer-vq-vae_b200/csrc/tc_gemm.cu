// Tensor-core (tcgen05 / TMEM / TMA) GEMM kernels of the residual-block BACKWARD pass
// (the reference leaves this to Chainer autograd over modules.py:30-56).
//
// One kernel template, four epilogues.  Operands are K-major bf16 hi/lo planes (see
// resblock_tc.cu for the split-precision scheme); every product is 3 MMAs in bf16x3 mode.
//
//   "time" flavour (rows of the tile = 128 time steps of one batch item, like the forward):
//     EPI_GATE_BWD  gz = Wr^T g_res + Ws^T g_skip (K = Cr + Cs), then the gate derivative with
//                   the saved tanh/sigmoid -> gh, written both time-major (B,T,Cd) for the
//                   data-gradient GEMMs and channel-major (B,Cd,T) for the weight-gradient GEMMs
//     EPI_GX        gx[t] = g_res[t] + sum_j Wc_j^T gh[t + dil*(fs-1-j)]  (anti-causal taps are
//                   row shifts of the TMA box; rows past T are out of bounds = zero)
//     EPI_ACCUM     gcond += Wp^T gh
//   "wgrad" flavour (rows = 128 output channels, K = time, split over batch items):
//     EPI_WGRAD     gW[m,n] += sum_t A[m,t] * B[n,t - shift]   (atomic accumulation; a row of
//                   ones appended to B yields the bias gradient in the same MMA)
//
// CTA layout as in the forward kernel (warp 4 = TMA producer, warp 5 = MMA issuer, warps 0-3 =
// epilogue, one thread per TMEM lane) but with a 2-stage ring and a 256-column accumulator so
// that TWO CTAs are resident per SM: one CTA's epilogue overlaps the other's MMAs.
#include "tc_common.cuh"
#include <stdlib.h>

namespace vqw {
namespace tc {

constexpr int G_STAGES = 2;
constexpr int G_EPI_WARPS = 8;                         // 2 warps per TMEM lane quadrant
constexpr int G_THREADS = (G_EPI_WARPS + 2) * 32;
constexpr int GW_TMA = G_EPI_WARPS, GW_MMA = G_EPI_WARPS + 1;
constexpr int MAX_SEG = 4;
enum { EPI_GATE_BWD = 0, EPI_GX = 1, EPI_ACCUM = 2, EPI_WGRAD = 3, EPI_WGRAD_MN = 4 };

struct Seg {
  int a_map, b_map;   // tensor-map pair index: hi plane = maps[2*i], lo plane = maps[2*i+1]
  int nslabs;         // K / 32 (time flavour)
  int a_c0, b_c0;     // coordinate along the contiguous (K) axis of slab 0
  int a_shift;        // time flavour: added to the time-row coordinate of A
  int b_row0;         // time flavour: first B row; + 256 * blockIdx.y
};

struct GemmParams {
  int nseg;
  Seg seg[MAX_SEG];
  int x3;
  int B, T;
  int M, N;              // wgrad: valid rows / cols of the whole output
  int slabs_per_item;    // wgrad: K slabs per (batch item, time chunk) work item
  int chunks_per_b;      // wgrad: work items per batch item
  const float* f0;       // GATE_BWD: tanh (B,256,T);  GX: g_res addend (B,Cout,T) or null
  const float* f1;       // GATE_BWD: sigmoid
  float* o0;             // GX: gx fp32 (B,Cout,T) | ACCUM: gcond (B,Cout,T) | WGRAD: gW
  __nv_bfloat16* p_hi;   // time-major (B,T,C) output planes (or null)
  __nv_bfloat16* p_lo;
  __nv_bfloat16* c_hi;   // channel-major (B,C,T) output planes (or null)
  __nv_bfloat16* c_lo;
  int Cout;              // channels of the output tensor
  long long gm, gk;      // WGRAD: strides of gW along m / n
  float* gb;             // WGRAD: bias gradient (or null)
  float* gb2;            // WGRAD: second bias receiving the same gradient (conv_b and cond_b)
  int ones_col;          // WGRAD: column of B that is the appended row of ones (-1: none)
};

struct Maps {
  CUtensorMap m[8];
};

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 2)
tc_gemm_kernel(const __grid_constant__ Maps maps, const __grid_constant__ GemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + G_STAGES);
  const uint32_t acc_full = smem_u32(bars + 2 * G_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool WG = (EPI == EPI_WGRAD || EPI == EPI_WGRAD_MN);
  constexpr bool MN = (EPI == EPI_WGRAD_MN);   // operands read from time-major planes (MN-major tiles)
  const int nplanes = P.x3 ? 2 : 1;

  // total number of K slabs this CTA contracts over
  int total_slabs = 0;
  int n_items = 0;
  if (WG) {
    const int items = P.B * P.chunks_per_b;
    for (int c = blockIdx.z; c < items; c += gridDim.z) ++n_items;
    total_slabs = n_items * P.slabs_per_item;
  } else {
    for (int s = 0; s < P.nseg; ++s) total_slabs += P.seg[s].nslabs;
  }

  if (warp == GW_TMA && lane == 0) {
    for (int i = 0; i < 8; ++i) prefetch_tmap(&maps.m[i]);
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GW_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == GW_TMA) {
    // =============================== TMA producer ===============================
    if (lane == 0 && total_slabs > 0) {
      int stage = 0;
      uint32_t ph = 0;
      auto issue = [&](const Seg& sg, int a0, int a1, int a2, int b0, int b1, int b2) {
        mbar_wait(empty0 + 8 * stage, ph ^ 1);
        const uint32_t fb = full0 + 8 * stage;
        const uint32_t sa = base + stage * STAGE_BYTES;
        mbar_expect_tx(fb, nplanes * (A_PLANE + B_PLANE));
        tma_load_3d(sa, &maps.m[2 * sg.a_map], fb, a0, a1, a2);
        tma_load_3d(sa + 2 * A_PLANE, &maps.m[2 * sg.b_map], fb, b0, b1, b2);
        if (P.x3) {
          tma_load_3d(sa + A_PLANE, &maps.m[2 * sg.a_map + 1], fb, a0, a1, a2);
          tma_load_3d(sa + 2 * A_PLANE + B_PLANE, &maps.m[2 * sg.b_map + 1], fb, b0, b1, b2);
        }
        if (++stage == G_STAGES) { stage = 0; ph ^= 1; }
      };
      if (MN) {
        // K = time is the ROW axis of the time-major planes: a stage is 2 (A) + 4 (B) boxes of
        // {64 channels, 32 time steps}; the tap delay is a row coordinate (no alignment rule)
        const Seg& sg = P.seg[0];
        const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
        const int items = P.B * P.chunks_per_b;
        for (int c = blockIdx.z; c < items; c += gridDim.z) {
          const int bb = c / P.chunks_per_b;
          const int tk = (c % P.chunks_per_b) * P.slabs_per_item * BK;
          for (int i = 0; i < P.slabs_per_item; ++i) {
            mbar_wait(empty0 + 8 * stage, ph ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t sa = base + stage * STAGE_BYTES;
            mbar_expect_tx(fb, nplanes * (A_PLANE + B_PLANE));
            const int ta = tk + sg.a_c0 + i * BK, tb = tk + sg.b_c0 + i * BK;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_load_3d(sa + h * 4096, &maps.m[2 * sg.a_map], fb, m0 + 64 * h, ta, bb);
              if (P.x3) tma_load_3d(sa + A_PLANE + h * 4096, &maps.m[2 * sg.a_map + 1], fb, m0 + 64 * h, ta, bb);
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              tma_load_3d(sa + 2 * A_PLANE + h * 4096, &maps.m[2 * sg.b_map], fb, n0 + 64 * h, tb, bb);
              if (P.x3)
                tma_load_3d(sa + 2 * A_PLANE + B_PLANE + h * 4096, &maps.m[2 * sg.b_map + 1], fb,
                            n0 + 64 * h, tb, bb);
            }
            if (++stage == G_STAGES) { stage = 0; ph ^= 1; }
          }
        }
      } else if (WG) {
        const Seg& sg = P.seg[0];
        const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
        const int items = P.B * P.chunks_per_b;
        for (int c = blockIdx.z; c < items; c += gridDim.z) {
          const int bb = c / P.chunks_per_b;
          const int tk = (c % P.chunks_per_b) * P.slabs_per_item * BK;
          for (int i = 0; i < P.slabs_per_item; ++i)
            issue(sg, tk + sg.a_c0 + i * BK, m0, bb, tk + sg.b_c0 + i * BK, n0, bb);
        }
      } else {
        const int t0 = blockIdx.x * TM, bb = blockIdx.z;
        for (int s = 0; s < P.nseg; ++s) {
          const Seg& sg = P.seg[s];
          for (int i = 0; i < sg.nslabs; ++i)
            issue(sg, sg.a_c0 + i * BK, t0 + sg.a_shift, bb, sg.b_c0 + i * BK,
                  sg.b_row0 + TN * blockIdx.y, 0);
        }
      }
    }
  } else if (warp == GW_MMA) {
    // =============================== MMA issuer =================================
    if (lane == 0 && total_slabs > 0) {
      int stage = 0;
      uint32_t ph = 0;
      for (int i = 0; i < total_slabs; ++i) {
        mbar_wait(full0 + 8 * stage, ph);
        tc_fence_after();
        const uint32_t sa = base + stage * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / UK; ++ks) {
          if (MN) {
            // 16 K-rows = two 1024-byte swizzle atoms per step
            const uint64_t a_hi = smem_desc_sw128_mn(sa + ks * 2048);
            const uint64_t b_hi = smem_desc_sw128_mn(sa + 2 * A_PLANE + ks * 2048);
            mma_ss(tmem_base, a_hi, b_hi, IDESC_MN, (i | ks) ? 1u : 0u);
            if (P.x3) {
              const uint64_t a_lo = smem_desc_sw128_mn(sa + A_PLANE + ks * 2048);
              const uint64_t b_lo = smem_desc_sw128_mn(sa + 2 * A_PLANE + B_PLANE + ks * 2048);
              mma_ss(tmem_base, a_lo, b_hi, IDESC_MN, 1u);
              mma_ss(tmem_base, a_hi, b_lo, IDESC_MN, 1u);
            }
          } else {
            const uint64_t a_hi = smem_desc_sw64(sa + ks * UK * 2);
            const uint64_t b_hi = smem_desc_sw64(sa + 2 * A_PLANE + ks * UK * 2);
            mma_ss(tmem_base, a_hi, b_hi, IDESC, (i | ks) ? 1u : 0u);
            if (P.x3) {
              const uint64_t a_lo = smem_desc_sw64(sa + A_PLANE + ks * UK * 2);
              const uint64_t b_lo = smem_desc_sw64(sa + 2 * A_PLANE + B_PLANE + ks * UK * 2);
              mma_ss(tmem_base, a_lo, b_hi, IDESC, 1u);
              mma_ss(tmem_base, a_hi, b_lo, IDESC, 1u);
            }
          }
        }
        tc_commit(empty0 + 8 * stage);
        if (++stage == G_STAGES) { stage = 0; ph ^= 1; }
      }
      tc_commit(acc_full);
    }
  } else if (total_slabs > 0) {
    // =============================== epilogue (warps 0-7) =======================
    // warp e: TMEM lane quadrant e%4, column group e/4; 16-column chunks dealt round-robin.
    const int quad = warp & 3, grp = warp >> 2;
    constexpr int NG = G_EPI_WARPS / 4;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);

    if (WG) {
      const int m = blockIdx.x * TM + row;
      const int n0 = blockIdx.y * TN;
      mbar_wait(acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int q = grp; q < TN / 16; q += NG) {
        float o[16];
        tmem_ld16(lane_base + 16 * q, o);
        if (m < P.M) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + 16 * q + i;
            if (n < P.N)
              atomicAdd(P.o0 + (long long)m * P.gm + (long long)n * P.gk, o[i]);
            else if (n == P.ones_col && P.gb != nullptr) {
              atomicAdd(P.gb + m, o[i]);
              if (P.gb2 != nullptr) atomicAdd(P.gb2 + m, o[i]);
            }
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
        const float* tp = P.f0 + ((int64_t)b * CHh) * P.T + t;
        const float* sp = P.f1 + ((int64_t)b * CHh) * P.T + t;
        float pt[16], ps[16];
        auto fetch = [&](int q) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            pt[i] = t_ok ? __ldcs(tp + (int64_t)(16 * q + i) * P.T) : 0.0f;
            ps[i] = t_ok ? __ldcs(sp + (int64_t)(16 * q + i) * P.T) : 0.0f;
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
          for (int i = 0; i < 16; ++i) { th[i] = pt[i]; sg[i] = ps[i]; }
          if (q + NG < TN / 16) fetch(q + NG);
          uint32_t th_hi[8], th_lo[8], sg_hi[8], sg_lo[8];
          const int zc0 = 16 * q;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            __nv_bfloat16 h[2][2], l[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float g = gz[i + u], a = th[i + u], s = sg[i + u];
              const float ght = g * s * (1.0f - a * a);
              const float ghs = g * a * s * (1.0f - s);
              split_bf16(ght, h[0][u], l[0][u]);
              split_bf16(ghs, h[1][u], l[1][u]);
              if (t_ok && P.c_hi != nullptr) {
                const int64_t o1 = ((int64_t)b * 2 * CHh + zc0 + i + u) * P.T + t;
                const int64_t o2 = o1 + (int64_t)CHh * P.T;
                P.c_hi[o1] = h[0][u];
                P.c_hi[o2] = h[1][u];
                if (P.x3) { P.c_lo[o1] = l[0][u]; P.c_lo[o2] = l[1][u]; }
              }
            }
            th_hi[i >> 1] = pack2(h[0][0], h[0][1]);
            th_lo[i >> 1] = pack2(l[0][0], l[0][1]);
            sg_hi[i >> 1] = pack2(h[1][0], h[1][1]);
            sg_lo[i >> 1] = pack2(l[1][0], l[1][1]);
          }
          if (t_ok && P.p_hi != nullptr) {
            const int64_t poff = ((int64_t)b * P.T + t) * (2 * CHh) + zc0;
            uint4* d0 = reinterpret_cast<uint4*>(P.p_hi + poff);
            uint4* d1 = reinterpret_cast<uint4*>(P.p_hi + poff + CHh);
            d0[0] = make_uint4(th_hi[0], th_hi[1], th_hi[2], th_hi[3]);
            d0[1] = make_uint4(th_hi[4], th_hi[5], th_hi[6], th_hi[7]);
            d1[0] = make_uint4(sg_hi[0], sg_hi[1], sg_hi[2], sg_hi[3]);
            d1[1] = make_uint4(sg_hi[4], sg_hi[5], sg_hi[6], sg_hi[7]);
            if (P.x3) {
              uint4* e0 = reinterpret_cast<uint4*>(P.p_lo + poff);
              uint4* e1 = reinterpret_cast<uint4*>(P.p_lo + poff + CHh);
              e0[0] = make_uint4(th_lo[0], th_lo[1], th_lo[2], th_lo[3]);
              e0[1] = make_uint4(th_lo[4], th_lo[5], th_lo[6], th_lo[7]);
              e1[0] = make_uint4(sg_lo[0], sg_lo[1], sg_lo[2], sg_lo[3]);
              e1[1] = make_uint4(sg_lo[4], sg_lo[5], sg_lo[6], sg_lo[7]);
            }
          }
        }
      } else {
        // EPI_GX / EPI_ACCUM: out[b, ch, t] = acc (+ addend) for ch = 256*blockIdx.y + col < Cout
        const int cbase = TN * blockIdx.y;
        const float* addsrc = (EPI == EPI_GX) ? P.f0 : P.o0;
        const float* addp = addsrc ? addsrc + ((int64_t)b * P.Cout + cbase) * P.T + t : nullptr;
        float pre[16];
        auto fetch = [&](int q) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pre[i] = (addp && t_ok && cbase + 16 * q + i < P.Cout)
                         ? __ldcs(addp + (int64_t)(16 * q + i) * P.T) : 0.0f;
        };
        fetch(grp);
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int q = grp; q < TN / 16; q += NG) {
          float o[16], add[16];
          tmem_ld16(lane_base + 16 * q, o);
#pragma unroll
          for (int i = 0; i < 16; ++i) add[i] = pre[i];
          if (q + NG < TN / 16) fetch(q + NG);
          const int ch0 = cbase + 16 * q;
          if (ch0 >= P.Cout) continue;      // Cout is a multiple of 16
          uint32_t vh[8], vl[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            __nv_bfloat16 h[2], l[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float v = o[i + u] + add[i + u];
              const int64_t off = ((int64_t)b * P.Cout + ch0 + i + u) * P.T + t;
              if (t_ok) P.o0[off] = v;
              split_bf16(v, h[u], l[u]);
              if (EPI == EPI_GX && t_ok && P.c_hi != nullptr) {
                P.c_hi[off] = h[u];
                if (P.x3) P.c_lo[off] = l[u];
              }
            }
            vh[i >> 1] = pack2(h[0], h[1]);
            vl[i >> 1] = pack2(l[0], l[1]);
          }
          if (EPI == EPI_GX && t_ok && P.p_hi != nullptr) {
            const int64_t poff = ((int64_t)b * P.T + t) * P.Cout + ch0;
            uint4* d0 = reinterpret_cast<uint4*>(P.p_hi + poff);
            d0[0] = make_uint4(vh[0], vh[1], vh[2], vh[3]);
            d0[1] = make_uint4(vh[4], vh[5], vh[6], vh[7]);
            if (P.x3) {
              uint4* e0 = reinterpret_cast<uint4*>(P.p_lo + poff);
              e0[0] = make_uint4(vl[0], vl[1], vl[2], vl[3]);
              e0[1] = make_uint4(vl[4], vl[5], vl[6], vl[7]);
            }
          }
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == GW_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256)
                 : "memory");
  }
}

static size_t gemm_smem() { return 1024 + (size_t)G_STAGES * STAGE_BYTES + 8 * (2 * G_STAGES + 1) + 16; }

template <int EPI>
static int launch_gemm(const Maps& maps, const GemmParams& P, dim3 grid, cudaStream_t stream) {
  auto kern = tc_gemm_kernel<EPI>;
  const size_t smem = gemm_smem();
  VQW_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, G_THREADS, smem, stream>>>(maps, P);
  static const char* names[5] = {"tc_gemm_kernel<GATE_BWD>", "tc_gemm_kernel<GX>",
                                 "tc_gemm_kernel<ACCUM>", "tc_gemm_kernel<WGRAD>",
                                 "tc_gemm_kernel<WGRAD_MN>"};
  VQW_CHECK_LAUNCH(names[EPI]);
  return 0;
}

// ------------------------------------------------------------------ operand preparation ----
// (B,C,T) fp32 [* mul], delayed by `shift` steps -> (B,rows,T) bf16 hi/lo planes; row C (if any)
// is ones, rows > C zero.  (TMA wants 16-byte aligned starts along the contiguous axis, so the
// per-tap time shift of the weight-gradient operand is applied here, not as a box coordinate.)
__global__ void __launch_bounds__(256)
cvt_planes_kernel(const float* __restrict__ in, const float* __restrict__ mul,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int C, int rows,
                  int T, int B, int shift) {
  // one thread = 8 consecutive time steps of one (b, row): 2 x float4 in, 16-byte stores out
  const int T8 = T >> 3;
  const int64_t n = (int64_t)B * rows * T8;
  const bool vec = (shift & 3) == 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int t0 = (int)(e % T8) * 8;
    const int64_t r = e / T8;
    const int c = (int)(r % rows), b = (int)(r / rows);
    float v[8];
    if (c >= C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (c == C) ? 1.0f : 0.0f;
    } else {
      const int64_t rowoff = ((int64_t)b * C + c) * T;
      const int ts0 = t0 - shift;          // out[t] = in[t - shift] (zero before the start)
      if (vec && ts0 >= 0 && ts0 + 7 < T) {
        const float4 a0 = __ldcs(reinterpret_cast<const float4*>(in + rowoff + ts0));
        const float4 a1 = __ldcs(reinterpret_cast<const float4*>(in + rowoff + ts0 + 4));
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
        v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        if (mul) {
          const float4 m0 = __ldcs(reinterpret_cast<const float4*>(mul + rowoff + ts0));
          const float4 m1 = __ldcs(reinterpret_cast<const float4*>(mul + rowoff + ts0 + 4));
          v[0] *= m0.x; v[1] *= m0.y; v[2] *= m0.z; v[3] *= m0.w;
          v[4] *= m1.x; v[5] *= m1.y; v[6] *= m1.z; v[7] *= m1.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ts = ts0 + i;
          float x = 0.0f;
          if (ts >= 0 && ts < T) {
            x = in[rowoff + ts];
            if (mul) x *= mul[rowoff + ts];
          }
          v[i] = x;
        }
      }
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * i], h0, l0);
      split_bf16(v[2 * i + 1], h1, l1);
      ph[i] = pack2(h0, h1);
      pl[i] = pack2(l0, l1);
    }
    const int64_t o = (r * T + t0);
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// out[r][k] (rows x K, K contiguous) = src(r, k) for the three transposed weight operands
//   kind 0: W2T [Ch rows][Cr + Cs]   = [Wr ; Ws]^T
//   kind 1: WcT [Cr rows][fs * Cd]   : WcT[cr][j*Cd + cd] = conv_w[cd][cr][j]
//   kind 2: WpT [Cc rows][Cd]        : WpT[cc][cd] = cond_w[cd][cc]
__global__ void __launch_bounds__(256)
pack_wt_kernel(int kind, const float* __restrict__ w0, const float* __restrict__ w1,
               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int rows, int K,
               int Cr, int Cs, int Cd, int Cc, int fs) {
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
      v = (r < Cc) ? w0[(int64_t)k * Cc + r] : 0.0f;
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[e] = h;
    if (lo) lo[e] = l;
  }
}

// gb[c] += sum_{b,t} in[b,c,t]
__global__ void __launch_bounds__(256)
rowsum_kernel(const float* __restrict__ in, float* __restrict__ gb, int C, int T) {
  const int c = blockIdx.x, b = blockIdx.y;
  const float* p = in + ((int64_t)b * C + c) * T;
  float s = 0.0f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) s += p[t];
  __shared__ float red[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    atomicAdd(gb + c, tot);
  }
}

int pack_act_launch(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int C, int T,
                    cudaStream_t stream);   // resblock_tc.cu

}  // namespace tc

static inline int64_t al(int64_t v) { return (v + 1023) / 1024 * 1024; }
static inline int pad256(int v) { return (v + 255) / 256 * 256; }   // TMA boxes never exceed the tensor

struct BwdLayout {
  int64_t total;
  int64_t gs_p[2], gs_c[2];          // g_skip: time-major / channel-major planes
  int64_t cond_c[2];                 // cond channel-major planes with the ones row (Cc+1 rows)
  int64_t x_c[2], x_s[2], z_c[2];    // per-block: x, x delayed by an unaligned tap shift, z = tanh*sig
  int64_t gh_p[2], gh_c[2];          // per-block: gh both layouts
  int64_t gr_f[2], gr_p[2][2], gr_c[2][2];   // g_res ping-pong: fp32, time-major, channel-major
  int64_t w2t[2], wct[2], wpt[2];    // per-block transposed weight planes (base; + i*wstride)
  int64_t wstride;
};

static BwdLayout bwd_layout(const vqw_resnet_desc& d) {
  BwdLayout L;
  int64_t off = 0;
  const int64_t N = (int64_t)d.B * d.T;
  auto take = [&](int64_t bytes) { int64_t o = off; off += al(bytes); return o; };
  for (int p = 0; p < 2; ++p) L.gs_p[p] = take(N * d.Cs * 2);
  for (int p = 0; p < 2; ++p) L.gs_c[p] = take(N * d.Cs * 2);
  for (int p = 0; p < 2; ++p) L.cond_c[p] = take(N * pad256(d.Cc + 1) * 2);
  for (int p = 0; p < 2; ++p) L.x_c[p] = take(N * d.Cr * 2);
  for (int p = 0; p < 2; ++p) L.x_s[p] = take(N * d.Cr * 2);
  for (int p = 0; p < 2; ++p) L.z_c[p] = take(N * (d.Cd / 2) * 2);
  for (int p = 0; p < 2; ++p) L.gh_p[p] = take(N * d.Cd * 2);
  for (int p = 0; p < 2; ++p) L.gh_c[p] = take(N * d.Cd * 2);
  for (int q = 0; q < 2; ++q) {
    L.gr_f[q] = take(N * d.Cr * 4);
    for (int p = 0; p < 2; ++p) L.gr_p[q][p] = take(N * d.Cr * 2);
    for (int p = 0; p < 2; ++p) L.gr_c[q][p] = take(N * d.Cr * 2);
  }
  const int64_t wbase = off;
  for (int p = 0; p < 2; ++p) L.w2t[p] = take((int64_t)(d.Cd / 2) * (d.Cr + d.Cs) * 2);
  for (int p = 0; p < 2; ++p) L.wct[p] = take((int64_t)d.Cr * d.fs * d.Cd * 2);
  for (int p = 0; p < 2; ++p) L.wpt[p] = take((int64_t)pad256(d.Cc) * d.Cd * 2);
  L.wstride = off - wbase;
  off = wbase + L.wstride * d.n_blocks;
  L.total = off + 1024;
  return L;
}

int64_t resnet_backward_tc_workspace(const vqw_resnet_desc& d) { return bwd_layout(d).total; }

int resnet_backward_tc(const vqw_resnet_desc& d, const float* g_skip, const float* g_last_res,
                       const float* x0, const float* cond, float* const* residuals,
                       float* const* gate_tanh, float* const* gate_sig,
                       const vqw_resblock_weights* weights, float* gx0, float* gcond,
                       const vqw_resblock_wgrads* wgrads, void* workspace, cudaStream_t stream) {
  using namespace tc;
  VQW_REQUIRE(resnet_tc_supported(d), "tcgen05 backward: unsupported channel counts");
  VQW_REQUIRE(d.Cc % 16 == 0 && d.Cr % 16 == 0, "tcgen05 backward: channels must be multiples of 16");
  VQW_REQUIRE(workspace && g_skip && x0 && cond && gate_tanh && gate_sig && weights && wgrads && gcond,
              "vqw_resnet_backward: null argument");
  const bool x3 = d.mode == VQW_MODE_BF16X3;
  const BwdLayout L = bwd_layout(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(al((int64_t)(uintptr_t)workspace));
  auto P16 = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto LO = [&](int64_t off) { return x3 ? P16(off) : nullptr; };
  const int B = d.B, T = d.T, Cr = d.Cr, Cd = d.Cd, Cs = d.Cs, Cc = d.Cc, Ch = d.Cd / 2, fs = d.fs;
  const int GRID1D = 148 * 8;

  // ---- once per call: g_skip in both layouts, cond channel-major with the ones row ----
  if (int rc = pack_act_launch(g_skip, P16(L.gs_p[0]), LO(L.gs_p[1]), B, Cs, T, stream)) return rc;
  cvt_planes_kernel<<<GRID1D, 256, 0, stream>>>(g_skip, nullptr, P16(L.gs_c[0]), LO(L.gs_c[1]), Cs,
                                                Cs, T, B, 0);
  VQW_CHECK_LAUNCH("cvt_planes_kernel(g_skip)");
  const int CcP = pad256(Cc + 1);   // rows: Cc condition channels, one row of ones, zero padding
  cvt_planes_kernel<<<GRID1D, 256, 0, stream>>>(cond, nullptr, P16(L.cond_c[0]), LO(L.cond_c[1]),
                                                Cc, CcP, T, B, 0);
  VQW_CHECK_LAUNCH("cvt_planes_kernel(cond)");
  for (int i = 0; i < d.n_blocks; ++i) {
    const vqw_resblock_weights& w = weights[i];
    const int64_t wo = i * L.wstride;
    pack_wt_kernel<<<148, 256, 0, stream>>>(0, w.res_w, w.skip_w, P16(L.w2t[0] + wo),
                                            LO(L.w2t[1] + wo), Ch, Cr + Cs, Cr, Cs, Cd, Cc, fs);
    VQW_CHECK_LAUNCH("pack_wt_kernel(0)");
    pack_wt_kernel<<<296, 256, 0, stream>>>(1, w.conv_w, nullptr, P16(L.wct[0] + wo),
                                            LO(L.wct[1] + wo), Cr, fs * Cd, Cr, Cs, Cd, Cc, fs);
    VQW_CHECK_LAUNCH("pack_wt_kernel(1)");
    pack_wt_kernel<<<148, 256, 0, stream>>>(2, w.cond_w, nullptr, P16(L.wpt[0] + wo),
                                            LO(L.wpt[1] + wo), pad256(Cc), Cd, Cr, Cs, Cd, Cc, fs);
    VQW_CHECK_LAUNCH("pack_wt_kernel(2)");
  }
  // gb_skip is the same for every block (g_skip is shared): computed per block below.

  auto map3 = [&](CUtensorMap* m, int64_t off_hi, int64_t off_lo, uint64_t inner, uint64_t rows,
                  uint64_t batch, uint32_t box_rows) -> int {
    if (int rc = make_map(&m[0], ws + off_hi, 3, inner, rows, batch, box_rows)) return rc;
    return make_map(&m[1], ws + (x3 ? off_lo : off_hi), 3, inner, rows, batch, box_rows);
  };

  // g_res of the last block: zero unless the caller kept the last residual
  int cur = 0;
  bool have_gres = false;
  if (g_last_res != nullptr) {
    VQW_CHECK_CUDA(cudaMemcpyAsync(ws + L.gr_f[0], g_last_res, (size_t)B * Cr * T * 4,
                                   cudaMemcpyDeviceToDevice, stream));
    if (int rc = pack_act_launch(g_last_res, P16(L.gr_p[0][0]), LO(L.gr_p[0][1]), B, Cr, T, stream)) return rc;
    cvt_planes_kernel<<<GRID1D, 256, 0, stream>>>(g_last_res, nullptr, P16(L.gr_c[0][0]),
                                                  LO(L.gr_c[0][1]), Cr, Cr, T, B, 0);
    VQW_CHECK_LAUNCH("cvt_planes_kernel(g_last_res)");
    have_gres = true;
  }

  for (int i = d.n_blocks - 1; i >= 0; --i) {
    const vqw_resblock_weights& w = weights[i];
    const vqw_resblock_wgrads& gw = wgrads[i];
    const int64_t wo = i * L.wstride;
    const float* xin = (i == 0) ? x0 : residuals[i - 1];
    VQW_REQUIRE(xin && gate_tanh[i] && gate_sig[i], "vqw_resnet_backward: block %d saved tensors", i);
    const int nxt = cur ^ 1;
    const int dil = d.dilations[i];
    const float* gres_f = have_gres ? reinterpret_cast<const float*>(ws + L.gr_f[cur]) : nullptr;

    // operands of the weight gradients: x and z = tanh*sig, channel-major planes
    cvt_planes_kernel<<<GRID1D, 256, 0, stream>>>(gate_tanh[i], gate_sig[i], P16(L.z_c[0]),
                                                  LO(L.z_c[1]), Ch, Ch, T, B, 0);
    VQW_CHECK_LAUNCH("cvt_planes_kernel(z)");

    // ---- A1: gz GEMM + gate derivative -> gh (both layouts) ----
    {
      Maps maps;
      GemmParams P = {};
      if (int rc = map3(&maps.m[0], L.gr_p[cur][0], L.gr_p[cur][1], Cr, T, B, TM)) return rc;
      if (int rc = map3(&maps.m[2], L.gs_p[0], L.gs_p[1], Cs, T, B, TM)) return rc;
      if (int rc = map3(&maps.m[4], L.w2t[0] + wo, L.w2t[1] + wo, Cr + Cs, Ch, 1, TN)) return rc;
      maps.m[6] = maps.m[4]; maps.m[7] = maps.m[5];
      int n = 0;
      if (have_gres) P.seg[n++] = Seg{0, 2, Cr / BK, 0, 0, 0, 0};
      P.seg[n++] = Seg{1, 2, Cs / BK, 0, Cr, 0, 0};
      P.nseg = n;
      P.x3 = x3; P.B = B; P.T = T;
      P.f0 = gate_tanh[i]; P.f1 = gate_sig[i];
      P.p_hi = P16(L.gh_p[0]); P.p_lo = LO(L.gh_p[1]);
      P.c_hi = P16(L.gh_c[0]); P.c_lo = LO(L.gh_c[1]);
      P.Cout = Cd;
      if (int rc = launch_gemm<EPI_GATE_BWD>(maps, P, dim3(ceil_div(T, TM), 1, B), stream)) return rc;
    }
    // ---- A2: gx = g_res + sum_j Wc_j^T gh[t + s_j] ; gcond += Wp^T gh ----
    {
      Maps maps;
      if (int rc = map3(&maps.m[0], L.gh_p[0], L.gh_p[1], Cd, T, B, TM)) return rc;
      if (int rc = map3(&maps.m[2], L.wct[0] + wo, L.wct[1] + wo, (uint64_t)fs * Cd, Cr, 1, TN)) return rc;
      if (int rc = map3(&maps.m[4], L.wpt[0] + wo, L.wpt[1] + wo, Cd, pad256(Cc), 1, TN)) return rc;
      maps.m[6] = maps.m[4]; maps.m[7] = maps.m[5];
      float* gx_out = (i == 0) ? gx0 : reinterpret_cast<float*>(ws + L.gr_f[nxt]);
      if (gx_out != nullptr) {
        for (int j0 = 0; j0 < fs; j0 += MAX_SEG) {
          GemmParams P = {};
          int n = 0;
          for (int j = j0; j < fs && n < MAX_SEG; ++j)
            P.seg[n++] = Seg{0, 1, Cd / BK, 0, j * Cd, dil * (fs - 1 - j), 0};
          P.nseg = n;
          P.x3 = x3; P.B = B; P.T = T;
          P.f0 = (j0 == 0) ? gres_f : gx_out;     // later tap groups accumulate onto the output
          P.o0 = gx_out;
          const bool lastgrp = j0 + MAX_SEG >= fs;
          if (i > 0 && lastgrp) {
            P.p_hi = P16(L.gr_p[nxt][0]); P.p_lo = LO(L.gr_p[nxt][1]);
            P.c_hi = P16(L.gr_c[nxt][0]); P.c_lo = LO(L.gr_c[nxt][1]);
          }
          P.Cout = Cr;
          if (int rc = launch_gemm<EPI_GX>(maps, P, dim3(ceil_div(T, TM), Cr / TN, B), stream)) return rc;
        }
      }
      {
        GemmParams P = {};
        P.nseg = 1;
        P.seg[0] = Seg{0, 2, Cd / BK, 0, 0, 0, 0};
        P.x3 = x3; P.B = B; P.T = T;
        P.o0 = gcond;
        P.Cout = Cc;
        if (int rc = launch_gemm<EPI_ACCUM>(maps, P, dim3(ceil_div(T, TM), ceil_div(Cc, TN), B), stream))
          return rc;
      }
    }
    // ---- weight gradients: K = time, split over batch items ----
    {
      const int slabs = ceil_div(T, BK);
      auto wgrad = [&](int64_t a_hi, int64_t a_lo, int M, int64_t b_hi, int64_t b_lo, int Nrows,
                       int Nvalid, int shift, float* out, long long gm, long long gk, float* gb,
                       float* gb2, int ones_col) -> int {
        Maps maps;
        if (int rc = map3(&maps.m[0], a_hi, a_lo, T, M, B, TM)) return rc;
        if (int rc = map3(&maps.m[2], b_hi, b_lo, T, Nrows, B, TN)) return rc;
        for (int k = 4; k < 8; ++k) maps.m[k] = maps.m[k - 4];
        GemmParams P = {};
        P.nseg = 1;
        P.seg[0] = Seg{0, 1, 0, 0, shift, 0, 0};
        P.x3 = x3; P.B = B; P.T = T;
        P.M = M; P.N = Nvalid;
        P.slabs_per_item = slabs; P.chunks_per_b = 1;
        P.o0 = out; P.gm = gm; P.gk = gk; P.gb = gb; P.gb2 = gb2; P.ones_col = ones_col;
        dim3 grid(ceil_div(M, TM), ceil_div(Nrows, TN), B);
        return launch_gemm<EPI_WGRAD>(maps, P, grid, stream);
      };
      // x as channel-major planes; a tap delay that is a multiple of 8 samples (16 bytes) is a
      // TMA box coordinate, any other delay needs its own shifted copy (TMA alignment rule)
      cvt_planes_kernel<<<GRID1D, 256, 0, stream>>>(xin, nullptr, P16(L.x_c[0]), LO(L.x_c[1]), Cr, Cr,
                                                    T, B, 0);
      VQW_CHECK_LAUNCH("cvt_planes_kernel(x)");
      static const bool use_mn = getenv("VQW_WGRAD_MN") && getenv("VQW_WGRAD_MN")[0] == '1';
      if (use_mn) {
        // MN-major validation path: gh and x as TIME-major planes, tap delay = row coordinate
        if (int rc = pack_act_launch(xin, P16(L.x_s[0]), LO(L.x_s[1]), B, Cr, T, stream)) return rc;
        for (int j = 0; j < fs; ++j) {
          Maps maps;
          if (int rc = make_map_mn(&maps.m[0], ws + L.gh_p[0], Cd, Cd, T, B)) return rc;
          if (int rc = make_map_mn(&maps.m[1], ws + (x3 ? L.gh_p[1] : L.gh_p[0]), Cd, Cd, T, B)) return rc;
          if (int rc = make_map_mn(&maps.m[2], ws + L.x_s[0], Cr, Cr, T, B)) return rc;
          if (int rc = make_map_mn(&maps.m[3], ws + (x3 ? L.x_s[1] : L.x_s[0]), Cr, Cr, T, B)) return rc;
          for (int k = 4; k < 8; ++k) maps.m[k] = maps.m[k - 4];
          GemmParams P = {};
          P.nseg = 1;
          P.seg[0] = Seg{0, 1, 0, 0, -dil * (fs - 1 - j), 0, 0};
          P.x3 = x3; P.B = B; P.T = T;
          P.M = Cd; P.N = Cr;
          P.slabs_per_item = slabs; P.chunks_per_b = 1;
          P.o0 = gw.conv_w + j; P.gm = (long long)Cr * fs; P.gk = fs; P.ones_col = -1;
          dim3 grid(ceil_div(Cd, TM), ceil_div(Cr, TN), B);
          if (int rc = launch_gemm<EPI_WGRAD_MN>(maps, P, grid, stream)) return rc;
        }
      } else
      for (int j = 0; j < fs; ++j) {
        const int sh = dil * (fs - 1 - j);
        if (sh % 8 == 0) {
          if (int rc = wgrad(L.gh_c[0], L.gh_c[1], Cd, L.x_c[0], L.x_c[1], Cr, Cr, -sh,
                             gw.conv_w + j, (long long)Cr * fs, fs, nullptr, nullptr, -1))
            return rc;
        } else {
          cvt_planes_kernel<<<GRID1D, 256, 0, stream>>>(xin, nullptr, P16(L.x_s[0]), LO(L.x_s[1]), Cr,
                                                        Cr, T, B, sh);
          VQW_CHECK_LAUNCH("cvt_planes_kernel(x shifted)");
          if (int rc = wgrad(L.gh_c[0], L.gh_c[1], Cd, L.x_s[0], L.x_s[1], Cr, Cr, 0, gw.conv_w + j,
                             (long long)Cr * fs, fs, nullptr, nullptr, -1))
            return rc;
        }
      }
      // cond projection; the appended row of ones gives sum_t gh = gb_conv = gb_cond
      if (int rc = wgrad(L.gh_c[0], L.gh_c[1], Cd, L.cond_c[0], L.cond_c[1], CcP, Cc, 0,
                         gw.cond_w, Cc, 1, gw.cond_b, gw.conv_b, Cc))
        return rc;
      if (have_gres) {
        if (int rc = wgrad(L.gr_c[cur][0], L.gr_c[cur][1], Cr, L.z_c[0], L.z_c[1], Ch, Ch, 0,
                           gw.res_w, Ch, 1, nullptr, nullptr, -1))
          return rc;
        rowsum_kernel<<<dim3(Cr, B), 256, 0, stream>>>(gres_f, gw.res_b, Cr, T);
        VQW_CHECK_LAUNCH("rowsum_kernel(g_res)");
      }
      if (int rc = wgrad(L.gs_c[0], L.gs_c[1], Cs, L.z_c[0], L.z_c[1], Ch, Ch, 0, gw.skip_w, Ch, 1,
                         nullptr, nullptr, -1))
        return rc;
      rowsum_kernel<<<dim3(Cs, B), 256, 0, stream>>>(g_skip, gw.skip_b, Cs, T);
      VQW_CHECK_LAUNCH("rowsum_kernel(g_skip)");
    }
    have_gres = true;
    cur = nxt;
  }
  return 0;
}

}  // namespace vqw
