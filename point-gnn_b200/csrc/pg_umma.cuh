// Thin inline-PTX layer over the Blackwell (sm_100a) primitives the tensor-core kernels use:
// mbarrier, cp.async.bulk, tcgen05 alloc / mma / commit / ld, clusters.  No CUTLASS dependency;
// descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pg {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- cluster -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}

// One lane of the (fully converged) warp; the same lane every time.  Code that issues tcgen05.mma /
// tcgen05.commit runs warp-converged and predicates only the instruction itself with this, so that
// descriptors computed from warp-uniform values stay in uniform registers (a divergent
// `if (lane == 0)` region makes the compiler wrap every UTCHMMA in an ELECT / R2UR.BROADCAST
// waterfall loop: measured ~90 issue cycles per MMA, more than the MMA itself takes).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster (rank may be our own).
// Default .release.cta semantics on purpose: an explicit .cluster scope compiles to MEMBAR.ALL.GPU
// per arrive (measured: 16x slowdown of the producer loop); the data handed over here lives in the
// arriving CTA's own shared memory and is published to the async proxy by fence.proxy.async.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  const uint32_t remote = mapa(smem_u32(bar), rank);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// Relaxed arrive (no memory ordering, compiles to a bare SYNCS.ARRIVE).  Used where the hand-over is
// ordered by tcgen05 fences instead of the memory model: returning TMEM to the MMA warp.  A release
// arrive there would wait for the epilogue's global reductions still in flight (~2 us, measured).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  const uint32_t remote = mapa(smem_u32(bar), rank);
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// (remote CTAs may arrive on this barrier; the plain .acquire.cta wait is what CUTLASS uses for
// cluster barriers too - an .acquire.cluster wait adds a CCTL.IVALL L1 invalidate per wait)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }

// generic-proxy smem writes -> visible to the async proxy (tensor core / bulk copy engines)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk copy global -> shared, completion on an mbarrier -------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tensor memory ---------------------------------------------------------------------------
template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t num_cols) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(num_cols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(num_cols)
                 : "memory");
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t num_cols) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(num_cols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(num_cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, no swizzle ("interleave"): the operand is a grid
// of 8-row x 16-byte core matrices (128 contiguous bytes each, row r of the core at +16*r).
//   lbo = byte distance between the two core matrices that are adjacent in K,
//   sbo = byte distance between core matrices adjacent in M (A) / N (B).
// bits [0,14) addr>>4 | [16,30) lbo>>4 | [32,46) sbo>>4 | [46,48) version=1 | [61,64) layout=0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3fffu);
  d |= uint64_t((lbo >> 4) & 0x3fffu) << 16;
  d |= uint64_t((sbo >> 4) & 0x3fffu) << 32;
  d |= uint64_t(1) << 46;
  return d;
}
// Instruction descriptor, kind::f16: BF16 x BF16 -> FP32, both operands K-major.
// bits [4,6) D fmt (1 = F32) | [7,10) A fmt (1 = BF16) | [10,13) B fmt | [17,23) N>>3 | [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
template <int kCtaGroup>
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  if constexpr (kCtaGroup == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

// Same with the A operand in TENSOR MEMORY: lane i of the issuing CTA's (and, for cta_group::2, of the peer's)
// TMEM holds row i of its M-half, K-major, two 16-bit elements per 32-bit column (element k in column
// k / 2, low half first) - i.e. a K = 16 MMA reads 8 columns starting at `tmem_a`.
template <int kCtaGroup>
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  if constexpr (kCtaGroup == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

// All previously issued MMAs of this thread complete -> one arrival on `bar` (1-CTA form), or on the
// barrier at the same offset in every CTA of `cta_mask` (2-CTA form).  Implies fence::before_thread_sync.
__device__ __forceinline__ void mma_commit_1cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// 32 lanes x 32 bit x 16 columns: thread t of the warp gets columns [c, c+16) of TMEM lane (base + t);
// the lane field of `tmem_addr` must be 32 * (warp_id % 4).
__device__ __forceinline__ void tmem_ld16(uint32_t tmem_addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(tmem_addr)
      : "memory");
}
// same, 32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t tmem_addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(tmem_addr)
      : "memory");
}
// registers -> TMEM: thread t of the warp writes v[0..7] to columns [c, c+8) of TMEM lane (base + t)
__device__ __forceinline__ void tmem_st8(uint32_t tmem_addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tmem_addr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- BF16 three-way split ------------------------------------------------------------------
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 significant bits, so the three products
// hi*hi' + lo*hi' + hi*lo' reproduce an fp32 product to ~2^-16 relative.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  *hi = *reinterpret_cast<const uint32_t*>(&h);
  *lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Cheaper split for the hot producers: hi = x TRUNCATED to BF16 (its fp32 image is one LOP3, the packed pair
// one PRMT), lo = bf16_rn(x - hi).  |lo| can be twice as large as with a rounded hi, so the split keeps ~2^-17
// instead of ~2^-18 relative - below the 2^-16 of the dropped lo*lo' term; 6 instead of 8 SASS instructions
// per pair of values.
__device__ __forceinline__ void split_bf16x2_trunc(float a, float b, uint32_t* hi, uint32_t* lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  *hi = __byte_perm(ua, ub, 0x7632);                       // {hi16(a), hi16(b)}: a in the low half
  const float la = a - __uint_as_float(ua & 0xffff0000u);
  const float lb = b - __uint_as_float(ub & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(la, lb);
  *lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace umma
}  // namespace pg
