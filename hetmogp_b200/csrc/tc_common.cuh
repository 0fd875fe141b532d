// tcgen05 / TMEM / mbarrier / bulk-copy PTX helpers shared by the tensor-core kernels (sm_100a only).
//
// Operand format of every tensor-core product in this library: "split fp16".  A value v (pre-scaled by an
// exact power of two into the fp16 range) is carried as hi = fp16(v), lo = fp16(v - hi)  (22 mantissa bits), and a
// product A.B is issued as three tcgen05.mma (kind::f16, fp32 accumulate in TMEM):  Ah.Bh + Ah.Bl + Al.Bh.
// bf16 hi/lo (16 bits) misses the 1e-4 ELBO tolerance and a single fp16/bf16/tf32 pass produces negative
// posterior variances on the benchmark inputs ("operand precision").
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// try_wait with a long suspend-time hint: the thread sleeps in hardware until the phase completes (wake-up ~60
// cycles after the arrive) instead of spinning -- pollers must not steal issue slots / MIO bandwidth from producers.
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
// non-blocking phase test
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    long long t0 = 0;
    for (uint32_t it = 1;; ++it) {
        if (mbar_try(bar, parity)) return;
        if ((it & 1023u) == 0) {                       // sleeping waiters wake on every barrier event: keep the loop lean
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 8000000000LL) __trap();  // ~4 s at 2 GHz
        }
    }
}
// Warp-collective wait: every lane executes the (hardware-suspended) try_wait, so the warp stays converged.
// (Polling from one lane + __syncwarp() does NOT reconverge on sm_100: lane 0 and lanes 1-31 then run the whole
// consumer body as two divergent groups -- measured: 2x the issued instructions, ncu "avg threads executed" = 16.)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }

// ---------------------------------------------------------------- thread-block cluster (CTA pair, cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
// wait on a barrier that peers of the cluster arrive on (acquire at cluster scope).  HM_ISSUER_WAIT 0: hardware-suspended
// try_wait; 1: no suspend hint -- the one issuing thread spins (0.1-0.2 ms per kernel at cfg3)
#ifndef HM_ISSUER_WAIT
#define HM_ISSUER_WAIT 1
#endif
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t it = 1;; ++it) {
        uint32_t ok;
#if HM_ISSUER_WAIT == 0
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
#else
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
#endif
        if (ok) return;
        if ((it & 1023u) == 0) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 8000000000LL) __trap();
        }
    }
}

// one non-blocking probe of such a barrier
__device__ __forceinline__ bool mbar_probe_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

// One lane of a converged warp (elect.sync).  The MMA-issuing warps run their loops with all 32 lanes and predicate only
// the tcgen05 instructions on this: the shared-memory descriptors and TMEM addresses then stay provably warp-uniform and
// ptxas keeps them in uniform registers.  Issued from inside an `if (lane == 0)` region instead, every UTCHMMA is preceded
// by seven R2UR moves (~120 cycles of issue per MMA: as long as the MMA itself runs, so the tensor pipe idles at every
// hiccup).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- async proxy / bulk copy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// cta_group::2 variants: one thread of the leader CTA issues for the pair (M = 256: 128 rows per CTA; B split along N)
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> one row (lane) per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one row (lane) per thread: 32 consecutive fp32 columns registers -> TMEM
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// 16-column variants (lower register pressure in the flush loops)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30)
// (unused for swizzled K-major, 1), SBO>>4 [32,46) = bytes between 8-row groups, version 1 [46,48), layout [61,64).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // rows of 128 B, 8-row atoms of 1024 B
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {    // rows of 64 B, 8-row atoms of 512 B
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (bit 4), A=B=f16 (0), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __host__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- arithmetic helpers
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): halves the issue slots of the generator loops
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
// (v0, v1) -> packed fp16 hi pair and fp16 residual pair
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 r = __fadd2_rn(make_float2(v0, v1), make_float2(-__low2float(h), -__high2float(h)));
    const __half2 l = __floats2half2_rn(r.x, r.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// (v0, v1) -> packed fp16 pair (round to nearest), and back
__device__ __forceinline__ uint32_t pack_h2(float v0, float v1) {
    const __half2 h = __floats2half2_rn(v0, v1);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
    const __half2 h = *reinterpret_cast<const __half2*>(&u);
    return make_float2(__low2float(h), __high2float(h));
}
// x*s = hi + lo in fp32 (input differences of nearby fp64 coordinates must survive fp32)
__device__ __forceinline__ void split_scaled(double x, double s, float& hi, float& lo) {
    const double v = x * s;
    hi = (float)v;
    lo = (float)(v - (double)hi);
}
// exact power of two as float, exponent clamped to the normal range
__device__ __forceinline__ float pow2i(int e) {
    e = e < -126 ? -126 : (e > 127 ? 127 : e);
    return __int_as_float((e + 127) << 23);
}

}  // namespace tc
