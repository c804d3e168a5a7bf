// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and tensor-map helpers shared by the
// tensor-core kernels (resblock_tc.cu, tc_gemm.cu).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace vqw {
namespace tc {

constexpr int TM = 128;     // rows per CTA tile = UMMA M
constexpr int TN = 256;     // UMMA N
constexpr int BK = 32;      // K elements per pipeline stage (one 64-byte swizzle row)
constexpr int UK = 16;      // UMMA K for 16-bit operands
constexpr int A_PLANE = TM * BK * 2;                      // 8 KB
constexpr int B_PLANE = TN * BK * 2;                      // 16 KB
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;    // 48 KB: {A_hi, A_lo, B_hi, B_lo}
constexpr int NTHREADS = 192;

// ------------------------------------------------------------------ PTX wrappers ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug must surface as a launch error, never as a hung GPU.
// one non-blocking probe of a barrier phase: issued early, its result is consumed after the work
// that does not depend on it (the instruction itself has ~150 cycles of latency)
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of one TMA box (no shared-memory stage, no barrier): decouples the DRAM latency of
// operands that are read for the first time from the depth of the shared-memory ring
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// shared -> global tensor store of one box (bulk async-group completion), and its group fences
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes become visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the same load without the wait: issue several, then ONE tmem_ld_wait that also carries the
// destination registers as in/out operands, so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t* a, uint32_t* b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]),
                 "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]),
                 "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]),
                 "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]),
                 "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2) ---
// A cluster of two CTAs runs M = 256 MMAs over both SMs; each SM stages its own 128 rows of A
// and HALF of B (the tensor core reads the other half from the peer's shared memory).
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
// one lane of a converged warp (elect.sync): the issuing lane of the TMA / MMA roles
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
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
// The same MMAs with the shared-memory descriptors passed as their LOW words only: the high word
// (SBO, version, swizzle mode) is a compile-time constant that ptxas materialises in a uniform
// register for free, which halves the R2UR traffic in front of every UTCHMMA (the issue loop of
// one elected thread is what bounds the K = 16 steps of the thin GEMMs).
constexpr uint32_t DESC_HI_SW64 = 0x80004020u;      // smem_desc_sw64() >> 32
constexpr uint32_t DESC_HI_SW128_MN = 0x40004040u;  // smem_desc_sw128_mn() >> 32
__device__ __forceinline__ uint32_t desc_lo_sw64(uint32_t saddr) {
  return ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
}
__device__ __forceinline__ uint32_t desc_lo_sw128_mn(uint32_t saddr) {
  return ((saddr & 0x3FFFFu) >> 4) | ((4096u >> 4) << 16);
}
template <uint32_t HI>
__device__ __forceinline__ void mma2_ss_w(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(HI)
      : "memory");
}
template <uint32_t HI>
__device__ __forceinline__ void mma2_ts_w(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(HI)
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

// K-major, 64-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bits):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 8 rows * 64 B,
// [46,48) version = 1, [61,64) layout = 4 (SWIZZLE_64B).
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// MN-major, 128-byte swizzle descriptor for an operand whose M/N axis is contiguous in shared
// memory: canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in bf16 elements, i.e. rows of
// 64 contiguous M/N elements (128 B), 8 K-rows per 1024-byte swizzle atom.  LBO = byte distance
// between successive 64-element M/N groups, SBO = byte distance between successive 8-row K
// groups.  Tiles are written by TMA boxes {64 elements, 32 rows} with CU_TENSOR_MAP_SWIZZLE_128B,
// one 4096-byte box per 64-element group: LBO = 4096, SBO = 1024.
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);
// same, both operands MN-major (bit 15 = A, bit 16 = B)
constexpr uint32_t IDESC_MN = IDESC | (1u << 15) | (1u << 16);
// A = B = IEEE fp16: format code 0 in both operand fields (VQW_MODE_FP16)
constexpr uint32_t IDESC_F16BITS = (1u << 7) | (1u << 10);
__host__ __device__ __forceinline__ constexpr uint32_t idesc_for(uint32_t base, int f16) {
  return f16 ? (base & ~IDESC_F16BITS) : base;
}

// Gate non-linearities.  The epilogue is special-function-unit bound (16 MUFU lanes/cycle/SM):
// exp stays on the MUFU (ex2), the two reciprocals run on the FMA pipe -- the argument of
// each is 1 + e with e in (0, 1], so a linear minimax seed (|err| <= 1/17) followed by three
// Newton steps r <- r * (2 - y r) gives 1/y to ~1e-7 relative without touching the MUFU.
__device__ __forceinline__ float rcp_1to2(float y) {
  float r = fmaf(-0.47058824f, y, 1.41176471f);     // 24/17 - 8/17 * y
  r = r * fmaf(-y, r, 2.0f);
  r = r * fmaf(-y, r, 2.0f);
  r = r * fmaf(-y, r, 2.0f);
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float v) {
  const float e = __expf(-fabsf(v));                // in (0, 1]
  const float s = rcp_1to2(1.0f + e);               // sigmoid(|v|)
  return v >= 0.0f ? s : e * s;                     // sigmoid(-|v|) = e / (1 + e)
}
__device__ __forceinline__ float tanh_fast(float v) {
  const float e = __expf(-2.0f * fabsf(v));
  return copysignf((1.0f - e) * rcp_1to2(1.0f + e), v);
}

// Both gate non-linearities of one pair with ONE reciprocal and two bare MUFU.EX2:
//   ea = e^{-2|a|}, eg = e^{-|g|};  tanh(a) = sgn(a)(1 - ea)/(1 + ea);  sigmoid(g) = (g >= 0 ? 1 : eg)/(1 + eg)
//   r = 1 / ((1 + ea)(1 + eg))  ->  tanh = sgn(a)(1 - ea)(1 + eg) r,  sigmoid = (g >= 0 ? 1 : eg)(1 + ea) r.
// ~19 issue slots and 3 MUFU operations per pair against 62 / 2 for tanh_fast + sigmoid_fast with
// __expf: round 2 measured the gate epilogue ISSUE bound (997 SASS instructions per 16 pairs).
// ex2.approx.ftz / rcp.approx.ftz are accurate to ~2 ulp; everything else is exact fp32.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gate_pair(float a, float g, float& th, float& sg) {
  const float ea = ex2_ftz(fabsf(a) * -2.8853900817779268f);   // -2 log2(e)
  const float eg = ex2_ftz(fabsf(g) * -1.4426950408889634f);   // -log2(e)
  const float pa = 1.0f + ea, pg = 1.0f + eg;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pa * pg));
  th = copysignf((1.0f - ea) * pg * r, a);
  sg = (g >= 0.0f ? 1.0f : eg) * pa * r;
}

// 256-bit global accesses (sm_100: LDG/STG.256): one instruction per 32-byte sector, i.e. per
// 16 bf16 channels of one time-major plane row
__device__ __forceinline__ void ld256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// two values -> packed bf16x2 hi and lo words (low half = first value): one F2FP pack
// instruction per plane instead of one conversion per value
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16),
                                                 v1 - __uint_as_float(hb & 0xffff0000u));
  hi = hb;
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// ---- format-generic variants: the planes hold bf16 (f16 = 0) or IEEE fp16 (f16 = 1) bits.
// fp16 has 11 significant bits (vs 8), so ONE pass over fp16 planes gives ~2e-4 relative per
// product -- inside the 1e-3 parity bar -- at a third of the bf16x3 tensor work; its narrow
// exponent range is handled by power-of-two gradient scaling in the backward (tc_gemm.cu).
__device__ __forceinline__ uint32_t pack_pair_f(float v0, float v1, int f16) {
  if (f16) {
    // one F2FP.SATFINITE: values beyond +-65504 saturate instead of becoming inf (and then, in a
    // hi/lo split, hi = inf, lo = -inf = NaN on recombination)
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v1), "f"(v0));
    return r;
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_pair_f(uint32_t w, int f16, float& v0, float& v1) {
  if (f16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    v0 = f.x;
    v1 = f.y;
  } else {
    v0 = __uint_as_float(w << 16);
    v1 = __uint_as_float(w & 0xffff0000u);
  }
}
__device__ __forceinline__ void split_pair_f(float v0, float v1, uint32_t& hi, uint32_t& lo, int f16) {
  hi = pack_pair_f(v0, v1, f16);
  float b0, b1;
  unpack_pair_f(hi, f16, b0, b1);
  lo = pack_pair_f(v0 - b0, v1 - b1, f16);
}
// the same for values that are not bounded by construction (the residual stream, packed inputs):
// fp16 planes saturate at +-65504 instead of producing hi = inf, lo = -inf (a NaN on recombination)
__device__ __forceinline__ void split_pair_sat_f(float v0, float v1, uint32_t& hi, uint32_t& lo, int f16) {
  split_pair_f(v0, v1, hi, lo, f16);      // pack_pair_f saturates by itself
}
// one value -> raw 16-bit patterns of its hi and lo parts
__device__ __forceinline__ void split_16(float v, int f16, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  if (f16) {
    v = fminf(fmaxf(v, -65504.0f), 65504.0f);          // saturate instead of inf / NaN
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = __ushort_as_bfloat16(__half_as_ushort(h));
    lo = __ushort_as_bfloat16(__half_as_ushort(l));
  } else {
    split_bf16(v, hi, lo);
  }
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}


// bf16 tensor map of rank 2 or 3 (innermost extent `inner` contiguous); box = {BK, box_rows, 1},
// 64-byte swizzle, out-of-bounds elements read as zero.
int make_map(CUtensorMap* m, const void* ptr, int rank, uint64_t inner, uint64_t rows,
             uint64_t batch, uint32_t box_rows);
// time-major plane (C channels contiguous with row pitch `pitch` elements, T rows, B items):
// box = {64 channels, 32 rows, 1}, 128-byte swizzle (MN-major operand tiles of the weight-
// gradient GEMMs)
int make_map_mn(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t pitch, uint64_t T,
                uint64_t B);
int make_map_mn4(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t T, uint64_t B);
// time-major plane (B,T,C): box = {64 channels, 128 rows, 1}, 128-byte swizzle -- the staging tiles of
// the forward's residual epilogue (TMA load of the addend, TMA store of the result)
int make_map_tile(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t T, uint64_t B);

}  // namespace tc

// Layout of the caller-owned `saved` buffer the tensor-core forward fills for the backward:
// time-major (B,T,C) bf16 hi/lo planes of the condition, of every block's input x_i and of
// every block's gated activation z_i (all offsets 1024-byte aligned).  (Saving the gate
// derivative as planes too was measured: the extra bf16 splits cost the SFU-bound forward
// gate epilogue more than the backward saved, so tanh/sigmoid stay fp32 tensors.)
// Condition planes hold the Cl = Cc - Cg time-varying channels, then ONE channel that is 1.0 at
// every time step, zero padded to a pitch of Cl + 32: the weight-gradient GEMM against these
// planes (N tile = 256 columns, Cl + 1 of them used) then yields the gate-bias gradient -- the
// column sum of gh -- as column Cl for free, per batch item (tc_gemm.cu, EPI_WGRAD).
inline int cond_local(const vqw_resnet_desc& d) { return d.Cc - d.Cg; }
inline int cond_pitch(const vqw_resnet_desc& d) { return d.Cc - d.Cg + 32; }
struct TcSaved {
  int64_t cond[2];
  int64_t x0, x_plane, x_stride;   // block i: hi at x0 + i*x_stride, lo at + x_plane
  int64_t z0, z_plane, z_stride;
  int64_t total;
};
inline TcSaved tc_saved_layout(const vqw_resnet_desc& d) {
  auto al = [](int64_t v) { return (v + 1023) / 1024 * 1024; };
  TcSaved L;
  const int64_t N = (int64_t)d.B * d.T;
  const int64_t cplane = al(N * cond_pitch(d) * 2);
  L.cond[0] = 0;
  L.cond[1] = cplane;
  L.x_plane = al(N * d.Cr * 2);
  L.x_stride = 2 * L.x_plane;
  L.x0 = 2 * cplane;
  L.z_plane = al(N * (d.Cd / 2) * 2);
  L.z_stride = 2 * L.z_plane;
  L.z0 = L.x0 + (int64_t)d.n_blocks * L.x_stride;
  L.total = L.z0 + (int64_t)d.n_blocks * L.z_stride + 1024;
  return L;
}
}  // namespace vqw
