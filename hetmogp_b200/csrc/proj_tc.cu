// N-sized projection kernels on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Same contraction as proj_simt.cu (reference: /root/reference/hetmogp/svmogp_inf.py:212-218 restated):
//     P = K_tq C_q   (128-row tile x M) ,  c_tq[n] = sum_j P[n,j] K[n,j] ,  a_tq[n] = K[n,:] . alpha_q
// but the product runs as tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM) with both operands
// split into bf16 hi + lo parts and three products  A_hi B_hi + A_hi B_lo + A_lo B_hi  (~2^-17 relative operand
// error, fp32-class results; a single bf16 product misses the 1e-4 ELBO tolerance).
//
// Warp roles (320 threads, 1 CTA/SM, persistent over row tiles; grid.y = latent q):
//   warps 0-3  generators : build the K_fu tile of this stage (128 rows x 64 inducing points) from (x, Z_q) with one
//                           MUFU ex2 per entry, split hi/lo, store it straight into the SWIZZLE_128B K-major smem
//                           image the MMA reads (the tile never exists in HBM; reference: util.py:145-164)
//   warps 4-7  epilogue   : tcgen05.ld the 128x256 fp32 accumulator, multiply by the regenerated K entries and reduce
//                           along the row (thread = row = TMEM lane), accumulate a, c
//   warp  8    MMA issuer : one thread issues 12 tcgen05.mma (M128 N256 K16) per stage, tcgen05.commit -> mbarriers
//   warp  9    TMA        : cp.async.bulk of the pre-swizzled C_q operand image (hi+lo, 64 KB per stage) from L2
// Rings: 2 smem stages (96 KB each) full/empty, 2 TMEM accumulator buffers (2 x 256 columns) full/empty.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kRows = 128;                 // UMMA M
constexpr int kNB = 256;                   // UMMA N (output columns per job)
constexpr int kKB = 64;                    // inducing points per stage (= 128 B of bf16 = one swizzle atom row)
constexpr int kStages = 2;
constexpr int kAHalf = kRows * 128;        // 16 KB : A hi (or lo) image of one stage
constexpr int kBHalf = kNB * 128;          // 32 KB : B hi (or lo) image of one stage
constexpr int kStageBytes = 2 * kAHalf + 2 * kBHalf;   // 96 KB
constexpr int kThreads = 320;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO | SBO=1024B | v1 | SW128
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNB >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);

struct TileRef { int t; int64_t row0; int nrows; };
__device__ __forceinline__ TileRef find_tile(const HmTasks& tk, int64_t tile) {
    TileRef r; r.t = 0; r.row0 = 0; r.nrows = 0;
    for (int t = 0; t < tk.T; ++t) {
        const int64_t nt = (tk.count[t] + kRows - 1) / kRows;
        if (tile < nt) {
            r.t = t; r.row0 = tile * kRows;
            const int64_t rem = tk.count[t] - r.row0;
            r.nrows = rem < kRows ? (int)rem : kRows;
            return r;
        }
        tile -= nt;
    }
    return r;
}

__device__ __forceinline__ void split_scaled(double x, double s, float& hi, float& lo) {
    const double v = x * s;
    hi = (float)v;
    lo = (float)(v - (double)hi);
}

// ------------------------------------------------------------------------------------------- operand image of C_q
// Cb layout: [q][h = column block of 256][kb = k block of 64] { hi image (256 rows x 128 B, SW128), lo image }.
// Row j of an image holds B[j][k] = C_q[h*256 + j][kb*64 + k]; 16-byte chunk c of row j sits at chunk (c ^ (j & 7)).
__global__ void tc_prepare_kernel(const double* __restrict__ C, uint16_t* __restrict__ Cb, int Mp, int Mc) {
    const int q = blockIdx.z;
    const int nkb = Mc / kKB;
    const int h = blockIdx.y / nkb, kb = blockIdx.y % nkb;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;   // (row j, chunk c)
    if (e >= kNB * 8) return;
    const int j = e >> 3, c = e & 7;
    const double* src = C + ((size_t)q * Mp + (size_t)(h * kNB + j)) * Mp + kb * kKB + c * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float v0 = (float)src[2 * p], v1 = (float)src[2 * p + 1];
        const __nv_bfloat162 hb = __floats2bfloat162_rn(v0, v1);
        const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hb);
        const float r0 = v0 - __uint_as_float(hu << 16), r1 = v1 - __uint_as_float(hu & 0xFFFF0000u);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(r0, r1);
        hi[p] = hu;
        lo[p] = *reinterpret_cast<const uint32_t*>(&lb);
    }
    uint8_t* img = reinterpret_cast<uint8_t*>(Cb) + ((size_t)(q * (Mc / kNB) + h) * nkb + kb) * (2 * kBHalf);
    const int off = j * 128 + ((c ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + kBHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------- forward kernel
struct TcSmem {
    uint64_t full[kStages], empty[kStages], tfull[2], tempty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_fwd_kernel(HmTasks tk, HmProjArgs pa, const uint16_t* __restrict__ Cb, int64_t ntiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operand images need a 1024-byte aligned base: align by hand (the launch reserves the slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int Mc = pa.Mc, Mp = pa.Mp, M = pa.M, Xd = pa.Xdim, Q = pa.Q;
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nhalf = Mc / kNB, nkb = Mc / kKB;

    uint8_t* stage_base = smem;                                              // kStages * 96 KB, 1024-aligned
    float* zh = reinterpret_cast<float*>(smem + kStages * kStageBytes);      // [Mc][Xd]
    float* zl = zh + (size_t)Mc * Xd;                                        // [Mc][Xd]
    float* bias = zl + (size_t)Mc * Xd;                                      // [Mc] log2(sigma^2) | -1e30 (padded column)
    float* als = bias + Mc;                                                  // [Mc] alpha_q
    TcSmem* sb = reinterpret_cast<TcSmem*>(als + Mc);

    const HmConsts* __restrict__ cs = pa.consts;
    const double sscale = sqrt(0.5 * 1.4426950408889634 * cs->inv_l2[q]);    // 2^(-(s d)^2) = exp(-d^2 / (2 l^2))
    for (int m = threadIdx.x; m < Mc; m += kThreads) {
        for (int i = 0; i < Xd; ++i) {
            const double z = (m < M) ? pa.Zp[((size_t)q * Mp + m) * Xd + i] : 0.0;
            split_scaled(z, sscale, zh[m * Xd + i], zl[m * Xd + i]);
        }
        bias[m] = (m < M) ? (float)log2(cs->var[q]) : -1.0e30f;
        als[m] = (m < M) ? (float)pa.alpha[(size_t)q * Mp + m] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&sb->full[s], 128 + 1); mbar_init(&sb->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sb->tfull[b], 1); mbar_init(&sb->tempty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    if (warp < 4) {
        // ======================================================= generators: thread = row of the tile
        const int r = threadIdx.x;
        int stage = 0; uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileRef tr = find_tile(tk, tile);
            float xh[HM_MAXXD], xl[HM_MAXXD];
            for (int i = 0; i < Xd; ++i) {
                const double x = (r < tr.nrows) ? tk.X[tr.t][(tk.begin[tr.t] + tr.row0 + r) * Xd + i] : 0.0;
                split_scaled(x, sscale, xh[i], xl[i]);
            }
            for (int h = 0; h < nhalf; ++h) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&sb->empty[stage], phase ^ 1);
                    uint8_t* a_hi = stage_base + (size_t)stage * kStageBytes + r * 128;
                    uint8_t* a_lo = a_hi + kAHalf;
#pragma unroll 2
                    for (int c = 0; c < 8; ++c) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            float kv[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int m = kb * kKB + c * 8 + 2 * p + u;
                                float arg = bias[m];
                                for (int i = 0; i < Xd; ++i) {
                                    const float d = (xh[i] - zh[m * Xd + i]) + (xl[i] - zl[m * Xd + i]);
                                    arg = fmaf(-d, d, arg);
                                }
                                kv[u] = ex2_approx(arg);
                            }
                            const __nv_bfloat162 hb = __floats2bfloat162_rn(kv[0], kv[1]);
                            const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hb);
                            const __nv_bfloat162 lb = __floats2bfloat162_rn(kv[0] - __uint_as_float(hu << 16),
                                                                            kv[1] - __uint_as_float(hu & 0xFFFF0000u));
                            hi[p] = hu;
                            lo[p] = *reinterpret_cast<const uint32_t*>(&lb);
                        }
                        const int off = (c ^ (r & 7)) << 4;
                        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    fence_async_smem();            // generic-proxy stores -> visible to the tensor-core (async) proxy
                    mbar_arrive(&sb->full[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 8) {
        // ======================================================= epilogue: thread = row = TMEM lane
        const int e = warp - 4, r = e * 32 + lane;
        uint32_t jc = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileRef tr = find_tile(tk, tile);
            float xh[HM_MAXXD], xl[HM_MAXXD];
            for (int i = 0; i < Xd; ++i) {
                const double x = (r < tr.nrows) ? tk.X[tr.t][(tk.begin[tr.t] + tr.row0 + r) * Xd + i] : 0.0;
                split_scaled(x, sscale, xh[i], xl[i]);
            }
            float a_acc = 0.f, c_acc = 0.f;
            for (int h = 0; h < nhalf; ++h, ++jc) {
                const uint32_t buf = jc & 1u;
                mbar_wait(&sb->tfull[buf], (jc >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + buf * kNB;
                for (int cc = 0; cc < kNB / 32; ++cc) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr + cc * 32));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int m = h * kNB + cc * 32 + jj;
                        float arg = bias[m];
                        for (int i = 0; i < Xd; ++i) {
                            const float d = (xh[i] - zh[m * Xd + i]) + (xl[i] - zl[m * Xd + i]);
                            arg = fmaf(-d, d, arg);
                        }
                        const float kv = ex2_approx(arg);
                        c_acc = fmaf(__uint_as_float(v[jj]), kv, c_acc);
                        a_acc = fmaf(kv, als[m], a_acc);
                    }
                }
                tc_fence_before();
                mbar_arrive(&sb->tempty[buf]);
            }
            if (r < tr.nrows) {
                float* ac = reinterpret_cast<float*>(tk.AC[tr.t]) + (tr.row0 + r) * 2 * Q;
                ac[q] = a_acc;
                ac[Q + q] = c_acc;
            }
        }
    } else if (warp == 8) {
        // ======================================================= MMA issuer (one thread)
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, jc = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int h = 0; h < nhalf; ++h, ++jc) {
                    const uint32_t buf = jc & 1u;
                    mbar_wait(&sb->tempty[buf], ((jc >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kNB;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&sb->full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes);
                        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + kAHalf);
                        const uint64_t b_hi = make_desc(sa + 2 * kAHalf), b_lo = make_desc(sa + 2 * kAHalf + kBHalf);
#pragma unroll
                        for (int ks = 0; ks < kKB / 16; ++ks) {
                            const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per K=16 step, in 16-byte units
                            tc_mma_bf16(d_tmem, a_hi + adv, b_hi + adv, kIdesc, (kb | ks) ? 1u : 0u);
                            tc_mma_bf16(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
                            tc_mma_bf16(d_tmem, a_lo + adv, b_hi + adv, kIdesc, 1u);
                        }
                        tc_commit(&sb->empty[stage]);      // frees the smem stage when these MMAs have read it
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&sb->tfull[buf]);            // accumulator of this job complete
                }
            }
        }
    } else {
        // ======================================================= TMA producer (one thread)
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int h = 0; h < nhalf; ++h) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&sb->empty[stage], phase ^ 1);
                        uint8_t* dst = stage_base + (size_t)stage * kStageBytes + 2 * kAHalf;
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(Cb) + ((size_t)(q * nhalf + h) * nkb + kb) * (2 * kBHalf);
                        mbar_expect_tx(&sb->full[stage], 2 * kBHalf);
#pragma unroll
                        for (int part = 0; part < 4; ++part)
                            bulk_g2s(dst + part * (kBHalf / 2), src + part * (kBHalf / 2), kBHalf / 2, &sb->full[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

size_t tc_smem_bytes(int Mc, int Xd) {
    return (size_t)kStages * kStageBytes + sizeof(float) * ((size_t)2 * Mc * Xd + 2 * Mc) + sizeof(TcSmem) + 64 + 1024;
}

}  // namespace

int hm_tc_available() { return 1; }

int hm_tc_prepare(cudaStream_t s, const double* C, const double* alpha, void* Cb, int Mp, int Mc, int Q) {
    (void)alpha;
    dim3 grid((unsigned)hm_cdiv(kNB * 8, 256), (unsigned)((Mc / kNB) * (Mc / kKB)), (unsigned)Q);
    tc_prepare_kernel<<<grid, 256, 0, s>>>(C, reinterpret_cast<uint16_t*>(Cb), Mp, Mc);
    HM_CUDA(cudaGetLastError());
    return 0;
}

int hm_tc_proj_fwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb) {
    int64_t ntiles = 0;
    for (int t = 0; t < tk.T; ++t) ntiles += hm_cdiv(tk.count[t], kRows);
    if (ntiles == 0) return 0;
    const size_t smem = tc_smem_bytes(a.Mc, a.Xdim);
    if (smem > 227 * 1024) {
        hm_set_error("tensor-core projection: M=%d (padded %d) with Xdim=%d needs %zu B of shared memory", a.M, a.Mc, a.Xdim, smem);
        return HMOGP_ERR_ARG;
    }
    HM_CUDA(cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nw = a.nworkers;
    if (ntiles < nw) nw = (int)ntiles;
    dim3 grid((unsigned)nw, (unsigned)a.Q);
    tc_fwd_kernel<<<grid, kThreads, smem, s>>>(tk, a, reinterpret_cast<const uint16_t*>(Cb), ntiles);
    HM_CUDA(cudaGetLastError());
    return 0;
}
