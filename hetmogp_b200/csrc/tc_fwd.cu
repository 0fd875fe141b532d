// Forward projection on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Contraction (reference: /root/reference/hetmogp/svmogp_inf.py:212-218 restated), per latent q and
// 128-row tile of task t, with K = k_q(X_t, Z_q) generated on the fly (reference materialises it: util.py:145-164):
//     P = K C_q                       tcgen05.mma, split-fp16 operands (3 products), fp32 accumulators in TMEM
//     c_tq[n] = sum_j P[n,j] K[n,j]   a_tq[n] = K[n,:] . alpha_q                                   (epilogue)
//     b_tq[n] = sum_j alpha_j K[n,j] |x_n - z_j|^2     e_tq[n] = sum_j P[n,j] K[n,j] |x_n - z_j|^2   (hyper only)
// b, e are the per-row scalars from which the lengthscale gradient of the K_mn chain (svmogp.py:139-141, GPy
// RBF.update_gradients_full) follows without a second contraction:  d l_q = (1/l^3) sum_n mu^c b + 2 omega^c e.
//
// Warp roles (576 threads, 1 CTA/SM, persistent over row tiles; grid.y = latent q):
//   warps 0-7   generators : K tile of this stage (128 rows x 64 inducing points; warp = 32 rows x 32 columns): one
//                            MUFU ex2 per entry, split into fp16 hi/lo, stored straight into the SWIZZLE_128B K-major
//                            smem image the MMA reads
//   warps 8-15  epilogue   : tcgen05.ld the 128x256 fp32 accumulator (thread = row = TMEM lane, warp = 32 lanes x 128
//                            columns), multiply by the regenerated K entries, reduce along the row
//   warp  16    MMA issuer : one thread, 12 tcgen05.mma (M128 N256 K16) per stage, tcgen05.commit -> mbarriers
//   warp  17    bulk copy  : cp.async.bulk of the pre-swizzled C_q operand image (hi+lo, 64 KB per stage) from L2
// Rings: 2 smem stages (96 KB each) full/empty, 2 TMEM accumulators (2 x 256 columns) full/empty.
#include <string.h>

#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kRows = 128;   // UMMA M
constexpr int kNB = 256;     // UMMA N (output columns per job)
constexpr int kKB = 64;      // inducing points per stage (128 B of fp16 = one swizzle-atom row)
constexpr int kAHalf = kRows * 128;                    // 16 KB : A hi (or lo) image of one stage
constexpr int kBHalf = kNB * 128;                      // 32 KB : B hi (or lo) image of one stage (all 256 rows)
constexpr int kMaxStages = 3;
// per-CTA stage: NCTA = 1: A (32 KB) + whole B tile (64 KB), 2 stages;  NCTA = 2 (cta_group::2 pair): A + this CTA's
// half of the B rows (32 KB), 3 stages
template <int NCTA> struct FwdCfg {
    static constexpr int bhalf = kBHalf / NCTA;
    static constexpr int stage_bytes = 2 * kAHalf + 2 * bhalf;
};
constexpr int kGenWarps = 8, kEpiWarps = 8, kMmaWarp = 16;   // + bulk-copy warp 17
constexpr int kThreads = 576;


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

// ------------------------------------------------------------------------------------------- scales
// max |C_q| over the M x M block (grid = (row blocks, Q); float bits are monotone for non-negative values), then
// cexp from it and kexp from sigma_q^2.
__global__ void tc_absmax_kernel(const double* __restrict__ C, unsigned* __restrict__ cmax, int M, int Mp) {
    const int q = blockIdx.y;
    float mx = 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = blockIdx.x * nwarp + warp; i < M; i += gridDim.x * nwarp)
        for (int j = lane; j < M; j += 32) mx = fmaxf(mx, fabsf((float)C[((size_t)q * Mp + i) * Mp + j]));
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 0 && mx == mx) atomicMax(&cmax[q], __float_as_uint(fminf(mx, 3.0e38f)));
}
__global__ void tc_scale_kernel(const unsigned* __restrict__ cmax, const HmConsts* __restrict__ cs, HmTcInfo* info) {
    const int q = threadIdx.x;
    const float mx = __uint_as_float(cmax[q]);
    int e = 0;
    if (mx > 0.f && isfinite(mx)) frexpf(mx, &e);          // mx < 2^e
    info->cexp[q] = (mx > 0.f && isfinite(mx)) ? 14 - e : 0;
    int ev = 0;
    const float v = (float)cs->var[q];
    if (v > 0.f && isfinite(v)) frexpf(v, &ev);
    info->kexp[q] = (v > 0.f && isfinite(v)) ? 12 - ev : 0;
}

// ------------------------------------------------------------------------------------------- operand image of C_q
// Cb layout: [q][h = column block of 256][kb = k block of 64] { hi image (256 rows x 128 B, SW128), lo image }.
// Row j of an image holds B[j][k] = 2^cexp C_q[h*256 + j][kb*64 + k]; 16-byte chunk c of row j sits at chunk (c ^ (j & 7)).
__global__ void tc_image_kernel(const double* __restrict__ C, const HmTcInfo* __restrict__ info, uint16_t* __restrict__ Cb,
                                int Mp, int Mc) {
    const int q = blockIdx.z;
    const int nkb = Mc / kKB;
    const int h = blockIdx.y / nkb, kb = blockIdx.y % nkb;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;   // (row j, chunk c)
    if (e >= kNB * 8) return;
    const int j = e >> 3, c = e & 7;
    const double sc = ldexp(1.0, info->cexp[q]);
    const double* src = C + ((size_t)q * Mp + (size_t)(h * kNB + j)) * Mp + kb * kKB + c * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) split2((float)(src[2 * p] * sc), (float)(src[2 * p + 1] * sc), hi[p], lo[p]);
    uint8_t* img = reinterpret_cast<uint8_t*>(Cb) + ((size_t)(q * (Mc / kNB) + h) * nkb + kb) * (2 * kBHalf);
    const int off = j * 128 + ((c ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + kBHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------- forward kernel
struct FwdBars {
    uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
    uint32_t tmem_base;
};

// compensated fp32 accumulation (TwoSum): the fp64 pipe is too slow to sit in the epilogue loop
__device__ __forceinline__ void two_sum_add(float& hi, float& lo, float x) {
    const float s = hi + x;
    const float bp = s - hi;
    lo += (hi - (s - bp)) + (x - bp);
    hi = s;
}

// Per-column constants, grouped by 8 columns so that one thread fetches them with 128-bit loads:
//   tab[chunk8][row][8]   rows: 2i = -s z_i (hi), 2i+1 = -s z_i (lo), 2XD = alpha_q
// (negated so that the packed loops are pure FADD2 / FFMA2:  K = ex2(-(d.d - bias))).  The biases are the same for every
// column (-(log2 sigma^2 + kexp) in the generator, -log2 sigma^2 in the epilogue) and travel in registers; a padded
// column has -s z = -1e18 on its first coordinate, so d.d = 1e36 and ex2 gives exactly 0 (shared-memory wavefronts, not
// issue slots, bound this kernel: every table row dropped is 12 % of its LSU traffic).
template <int XD> struct FwdTab { static constexpr int R = 2 * XD + 1; };

template <int XD, int NCTA, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tc_fwd_kernel(HmTasks tk, HmProjArgs pa, const uint16_t* __restrict__ Cb, const HmTcInfo* __restrict__ info, int64_t ntiles,
              int hyper, int npass) {
    constexpr int R = FwdTab<XD>::R;
    constexpr int kStages = STAGES, kStageBytes = FwdCfg<NCTA>::stage_bytes, kBH = FwdCfg<NCTA>::bhalf;
    constexpr uint32_t kIdesc = idesc_f16(kRows * NCTA, kNB);
    const uint32_t rank = (NCTA == 2) ? cluster_ctarank() : 0u;          // 0 = leader of the CTA pair
    const int64_t nsteps = (ntiles + NCTA - 1) / NCTA;                   // row tiles are taken NCTA at a time
    const int64_t step0 = blockIdx.x / NCTA, dstep = gridDim.x / NCTA;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operand images need a 1024-byte aligned base: align by hand (the launch reserves the slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int Mc = pa.Mc, Mp = pa.Mp, M = pa.M, Q = pa.Q;
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // k blocks of 64 inducing points: `nkbf` in the operand image (padded M), `nkb` worth contracting over -- K and C are
    // exactly zero on padded inducing points, so the blocks beyond M are skipped (M = 64: one of four)
    const int nhalf = Mc / kNB, nkbf = Mc / kKB, nkb = (M + kKB - 1) / kKB;

    uint8_t* stage_base = smem;                                                   // kStages * 96 KB, 1024-aligned
    float* tab = reinterpret_cast<float*>(smem + kStages * kStageBytes);          // [Mc/8][R][8]
    float* xch = tab + (size_t)Mc * R;                                            // [2][4][128] epilogue half-row exchange
    FwdBars* sb = reinterpret_cast<FwdBars*>(xch + 2 * 4 * 128);

    const HmConsts* __restrict__ cs = pa.consts;
    const double s2 = 0.5 * 1.4426950408889634 * cs->inv_l2[q];                   // 2^(-s2 d^2) = exp(-d^2 / (2 l^2))
    const double sscale = sqrt(s2);
    const int kexp = info->kexp[q], cexp = info->cexp[q];
    for (int m = threadIdx.x; m < Mc; m += kThreads) {
        float* t8 = tab + (size_t)(m >> 3) * R * 8 + (m & 7);
        for (int i = 0; i < XD; ++i) {
            const double z = (m < M) ? pa.Zp[((size_t)q * Mp + m) * XD + i] : 0.0;
            float h, l;
            split_scaled(z, sscale, h, l);
            t8[(2 * i) * 8] = (m < M || i > 0) ? -h : -1.0e18f;
            t8[(2 * i + 1) * 8] = (m < M) ? -l : 0.f;
        }
        t8[(2 * XD) * 8] = (m < M) ? (float)pa.alpha[(size_t)q * Mp + m] : 0.f;
    }
    const float lv_ = (float)log2(cs->var[q]);
    const float2 gen_bias = dup2(-(lv_ + (float)kexp)), epi_bias = dup2(-lv_);
    if (threadIdx.x == 0) {
        // full: generator warps + bulk-copy expect_tx (+ on the leader of a pair: the peer's relay)
        for (int s = 0; s < kStages; ++s) { mbar_init(&sb->full[s], kGenWarps + 1 + ((NCTA == 2 && rank == 0) ? 1 : 0)); mbar_init(&sb->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sb->tfull[b], 1); mbar_init(&sb->tempty[b], kEpiWarps * NCTA); }
        mbar_fence_init();
    }
    if (warp == kMmaWarp) { if (NCTA == 2) tmem_alloc2(&sb->tmem_base, 512u); else tmem_alloc(&sb->tmem_base, 512u); }
    fence_before();
    if (NCTA == 2) cluster_sync(); else __syncthreads();
    fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    if (warp < kGenWarps) {
        // ======================================================= generators: thread = 2 rows (lane, lane + 32 of a 64-row
        // half), warp = (row half, column quarter): the per-column table values are loaded once for both rows
        const int rg = warp & 1, cq = warp >> 1;
        const int r0 = rg * 64 + lane, r1 = r0 + 32;
        int stage = 0; uint32_t phase = 0;
        // the rows' inputs of the NEXT tile are fetched while this one is generated (an exposed HBM round trip per tile
        // was ~4 % of the kernel)
        double xa_n[XD], xb_n[XD];
        auto fetch_x = [&](int64_t st_) {
#pragma unroll
            for (int i = 0; i < XD; ++i) { xa_n[i] = 0.0; xb_n[i] = 0.0; }
            if (st_ >= nsteps) return;
            const TileRef tr = find_tile(tk, st_ * NCTA + rank);
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                if (r0 < tr.nrows) xa_n[i] = tk.X[tr.t][(tk.begin[tr.t] + tr.row0 + r0) * XD + i];
                if (r1 < tr.nrows) xb_n[i] = tk.X[tr.t][(tk.begin[tr.t] + tr.row0 + r1) * XD + i];
            }
        };
        fetch_x(step0);
        for (int64_t st_ = step0; st_ < nsteps; st_ += dstep) {
            float2 xh0[XD], xl0[XD], xh1[XD], xl1[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                float h, l;
                split_scaled(xa_n[i], sscale, h, l); xh0[i] = dup2(h); xl0[i] = dup2(l);
                split_scaled(xb_n[i], sscale, h, l); xh1[i] = dup2(h); xl1[i] = dup2(l);
            }
            fetch_x(st_ + dstep);
            for (int h = 0; h < nhalf; ++h) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_warp(&sb->empty[stage], phase ^ 1);
                    uint8_t* a0_hi = stage_base + (size_t)stage * kStageBytes + r0 * 128;
                    uint8_t* a1_hi = stage_base + (size_t)stage * kStageBytes + r1 * 128;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = cq * 2 + cc;
                        const float4* t4 = reinterpret_cast<const float4*>(tab + (size_t)(kb * 8 + c) * R * 8);
                        float2 e0[4], e1[4];   // d.d - bias for the 8 columns (4 pairs), rows r0 / r1
#pragma unroll
                        for (int p = 0; p < 4; ++p) e0[p] = e1[p] = gen_bias;
#pragma unroll
                        for (int i = 0; i < XD; ++i) {
                            const float4 h0 = t4[(2 * i) * 2], h1 = t4[(2 * i) * 2 + 1];
                            const float4 l0 = t4[(2 * i + 1) * 2], l1 = t4[(2 * i + 1) * 2 + 1];
                            const float2 nzh[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w)};
                            const float2 nzl[4] = {make_float2(l0.x, l0.y), make_float2(l0.z, l0.w), make_float2(l1.x, l1.y), make_float2(l1.z, l1.w)};
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const float2 da = add2(add2(xh0[i], nzh[p]), add2(xl0[i], nzl[p]));
                                const float2 db = add2(add2(xh1[i], nzh[p]), add2(xl1[i], nzl[p]));
                                e0[p] = fma2(da, da, e0[p]);
                                e1[p] = fma2(db, db, e1[p]);
                            }
                        }
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) split2(ex2(-e0[p].x), ex2(-e0[p].y), hi[p], lo[p]);
                        int off = (c ^ (r0 & 7)) << 4;
                        *reinterpret_cast<uint4*>(a0_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(a0_hi + kAHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                        for (int p = 0; p < 4; ++p) split2(ex2(-e1[p].x), ex2(-e1[p].y), hi[p], lo[p]);
                        off = (c ^ (r1 & 7)) << 4;
                        *reinterpret_cast<uint4*>(a1_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(a1_hi + kAHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    fence_async_smem();            // generic-proxy stores -> visible to the tensor-core (async) proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sb->full[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < kGenWarps + kEpiWarps) {
        // ======================================================= epilogue: thread = row = TMEM lane, warp = (lane quadrant, column half)
        const int ew = warp - kGenWarps, lq = ew & 3, ch = ew >> 2, r = lq * 32 + lane;
        const float inv_pc = pow2i(-(kexp + cexp));      // undo the operand scales of P
        const float inv_s2 = (float)(1.0 / s2);           // scaled squared distance -> |x - z|^2
        uint32_t jc = 0, tcount = 0;
        double x_n[XD];
        auto fetch_x = [&](int64_t st_) {
#pragma unroll
            for (int i = 0; i < XD; ++i) x_n[i] = 0.0;
            if (st_ >= nsteps) return;
            const TileRef tn = find_tile(tk, st_ * NCTA + rank);
#pragma unroll
            for (int i = 0; i < XD; ++i)
                if (r < tn.nrows) x_n[i] = tk.X[tn.t][(tk.begin[tn.t] + tn.row0 + r) * XD + i];
        };
        fetch_x(step0);
        for (int64_t st_ = step0; st_ < nsteps; st_ += dstep, ++tcount) {
            const TileRef tr = find_tile(tk, st_ * NCTA + rank);
            float xh[XD], xl[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) split_scaled(x_n[i], sscale, xh[i], xl[i]);
            fetch_x(st_ + dstep);
            float ah = 0.f, al = 0.f, chh = 0.f, cl = 0.f, bh = 0.f, bl = 0.f, eh = 0.f, el = 0.f;
            for (int h = 0; h < nhalf; ++h, ++jc) {
                const uint32_t buf = jc & 1u;
                mbar_wait_warp(&sb->tfull[buf], (jc >> 1) & 1u);
                fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + buf * kNB + ch * (kNB / 2);
#pragma unroll 1
                for (int cc = 0; cc < kNB / 64; ++cc) {
                    if (h * kNB + ch * (kNB / 2) + cc * 32 >= M) break;   // padded columns: K = 0 there, every sum gets exactly 0
                    uint32_t v[32];
                    tmem_ld32(taddr + cc * 32, v);
                    tmem_ld_wait();
                    float2 a2 = dup2(0.f), c2 = dup2(0.f), b2 = dup2(0.f), e2 = dup2(0.f);
                    const int m0 = h * kNB + ch * (kNB / 2) + cc * 32;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4* t4 = reinterpret_cast<const float4*>(tab + (size_t)((m0 >> 3) + g) * R * 8);
                        float2 u[4];
#pragma unroll
                        for (int i = 0; i < XD; ++i) {
                            const float4 h0 = t4[(2 * i) * 2], h1 = t4[(2 * i) * 2 + 1];
                            const float4 l0 = t4[(2 * i + 1) * 2], l1 = t4[(2 * i + 1) * 2 + 1];
                            const float2 nzh[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w)};
                            const float2 nzl[4] = {make_float2(l0.x, l0.y), make_float2(l0.z, l0.w), make_float2(l1.x, l1.y), make_float2(l1.z, l1.w)};
                            const float2 xh2 = dup2(xh[i]), xl2 = dup2(xl[i]);
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const float2 d = add2(add2(xh2, nzh[p]), add2(xl2, nzl[p]));
                                u[p] = (i == 0) ? mul2(d, d) : fma2(d, d, u[p]);
                            }
                        }
                        const float4 q0 = t4[(2 * XD) * 2], q1 = t4[(2 * XD) * 2 + 1];
                        const float2 aa[4] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), make_float2(q1.z, q1.w)};
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float2 ea = add2(u[p], epi_bias);                    // d.d - log2 sigma^2
                            const float2 kv = make_float2(ex2(-ea.x), ex2(-ea.y));
                            const float2 pk = mul2(make_float2(__uint_as_float(v[g * 8 + 2 * p]), __uint_as_float(v[g * 8 + 2 * p + 1])), kv);
                            const float2 ak = mul2(kv, aa[p]);
                            c2 = add2(c2, pk);
                            a2 = add2(a2, ak);
                            if (hyper) {
                                e2 = fma2(pk, u[p], e2);
                                b2 = fma2(ak, u[p], b2);
                            }
                        }
                    }
                    const float a32 = a2.x + a2.y, c32 = c2.x + c2.y, b32 = b2.x + b2.y, e32 = e2.x + e2.y;
                    two_sum_add(ah, al, a32);
                    two_sum_add(chh, cl, c32);
                    if (hyper) { two_sum_add(bh, bl, b32); two_sum_add(eh, el, e32); }
                }
                fence_before();
                __syncwarp();
                if (lane == 0) { if (NCTA == 2 && rank != 0) mbar_arrive_remote(&sb->tempty[buf], 0); else mbar_arrive(&sb->tempty[buf]); }
            }
            // combine the two column halves of each row: half 1 -> smem -> half 0 -> HBM
            float* xb = xch + (tcount & 1u) * 4 * 128;
            if (ch == 1) {
                xb[r] = ah + al; xb[128 + r] = chh + cl; xb[256 + r] = bh + bl; xb[384 + r] = eh + el;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            if (ch == 0 && r < tr.nrows) {
                float* ac = reinterpret_cast<float*>(tk.AC[tr.t]) + tr.row0 + r;   // SoA: array k at k * cap
                const size_t cap = (size_t)tk.cap[tr.t];
                ac[(size_t)q * cap] = (ah + al) + xb[r];
                ac[(size_t)(Q + q) * cap] = ((chh + cl) + xb[128 + r]) * inv_pc;
                if (hyper) {
                    ac[(size_t)(2 * Q + q) * cap] = ((bh + bl) + xb[256 + r]) * inv_s2;
                    ac[(size_t)(3 * Q + q) * cap] = ((eh + el) + xb[384 + r]) * inv_pc * inv_s2;
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ======================================================= MMA issuer (leader CTA of a pair): the whole warp runs the
        // loop, one elected lane issues (elect_one: the descriptors stay in uniform registers)
        if (rank == 0) {
            int stage = 0; uint32_t phase = 0, jc = 0;
            for (int64_t st_ = step0; st_ < nsteps; st_ += dstep) {
                for (int h = 0; h < nhalf; ++h, ++jc) {
                    const uint32_t buf = jc & 1u;
                    if (NCTA == 2) mbar_wait_cluster(&sb->tempty[buf], ((jc >> 1) & 1u) ^ 1u);
                    else mbar_wait(&sb->tempty[buf], ((jc >> 1) & 1u) ^ 1u);
                    fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kNB;
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (NCTA == 2) mbar_wait_cluster(&sb->full[stage], phase);
                        else mbar_wait(&sb->full[stage], phase);
                        fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes);
                        const uint64_t a_hi = desc_sw128(sa), a_lo = desc_sw128(sa + kAHalf);
                        const uint64_t b_hi = desc_sw128(sa + 2 * kAHalf), b_lo = desc_sw128(sa + 2 * kAHalf + kBH);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < kKB / 16; ++ks) {
                                const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per K=16 step, in 16-byte units
                                if (NCTA == 2) {
                                    mma2_f16(d_tmem, a_hi + adv, b_hi + adv, kIdesc, (kb | ks) ? 1u : 0u);
                                    if (npass >= 2) mma2_f16(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
                                    if (npass >= 3) mma2_f16(d_tmem, a_lo + adv, b_hi + adv, kIdesc, 1u);
                                } else {
                                    mma_f16(d_tmem, a_hi + adv, b_hi + adv, kIdesc, (kb | ks) ? 1u : 0u);
                                    if (npass >= 2) mma_f16(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
                                    if (npass >= 3) mma_f16(d_tmem, a_lo + adv, b_hi + adv, kIdesc, 1u);
                                }
                            }
                            if (NCTA == 2) commit2(&sb->empty[stage]); else commit(&sb->empty[stage]);   // frees the smem stage
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) { if (NCTA == 2) commit2(&sb->tfull[buf]); else commit(&sb->tfull[buf]); }   // accumulator complete
                    __syncwarp();
                }
            }
        } else if (NCTA == 2 && lane == 0) {
            // peer CTA of the pair: relay "my operand tiles of this stage are in shared memory" to the leader
            int stage = 0; uint32_t phase = 0;
            for (int64_t st_ = step0; st_ < nsteps; st_ += dstep)
                for (int h = 0; h < nhalf; ++h)
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&sb->full[stage], phase);
                        mbar_arrive_remote(&sb->full[stage], 0);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
        }
    } else {
        // ======================================================= bulk-copy producer (one thread): this CTA's rows of B
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t st_ = step0; st_ < nsteps; st_ += dstep) {
                for (int h = 0; h < nhalf; ++h) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&sb->empty[stage], phase ^ 1);
                        uint8_t* dst = stage_base + (size_t)stage * kStageBytes + 2 * kAHalf;
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(Cb) + ((size_t)(q * nhalf + h) * nkbf + kb) * (2 * kBHalf) + rank * kBH;
                        mbar_expect_tx(&sb->full[stage], 2 * kBH);
                        bulk_g2s(dst, src, kBH / 2, &sb->full[stage]);                                   // hi
                        bulk_g2s(dst + kBH / 2, src + kBH / 2, kBH / 2, &sb->full[stage]);
                        bulk_g2s(dst + kBH, src + kBHalf, kBH / 2, &sb->full[stage]);                    // lo
                        bulk_g2s(dst + kBH + kBH / 2, src + kBHalf + kBH / 2, kBH / 2, &sb->full[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    }
    // ---- teardown
    fence_before();
    if (NCTA == 2) cluster_sync(); else __syncthreads();
    if (warp == kMmaWarp) { if (NCTA == 2) tmem_dealloc2(tmem_base, 512u); else tmem_dealloc(tmem_base, 512u); }
}

template <int NCTA> size_t fwd_smem_bytes(int Mc, int Xd, int stages) {
    return (size_t)stages * FwdCfg<NCTA>::stage_bytes + sizeof(float) * ((size_t)Mc * (2 * Xd + 1) + 2 * 4 * 128) +
           sizeof(FwdBars) + 64 + 1024;
}

template <int XD, int NCTA, int STAGES>
int launch_fwd3(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, int64_t ntiles,
                bool hyper, int npass) {
    const size_t smem = fwd_smem_bytes<NCTA>(a.Mc, XD, STAGES);
    if (smem > 227 * 1024) {
        hm_set_error("tensor-core projection: M=%d (padded %d) with Xdim=%d needs %zu B of shared memory", a.M, a.Mc, XD, smem);
        return HMOGP_ERR_ARG;
    }
    auto kern = tc_fwd_kernel<XD, NCTA, STAGES>;
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nw = a.nworkers;
    const int64_t nsteps = (ntiles + NCTA - 1) / NCTA;
    if (nsteps * NCTA < nw) nw = (int)(nsteps * NCTA);
    nw -= nw % NCTA;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)nw, (unsigned)a.Q);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (NCTA == 2) ? 1 : 0;
    HM_CUDA(cudaLaunchKernelEx(&cfg, kern, tk, a, reinterpret_cast<const uint16_t*>(Cb), info, ntiles, hyper ? 1 : 0, npass));
    HM_CUDA(cudaGetLastError());
    return 0;
}

template <int XD>
int launch_fwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, int64_t ntiles,
               bool hyper, int npass, int ncta) {
    // the per-column tables grow with M: a pair drops from 3 to 2 operand stages when they no longer fit (M > 1536 at Xdim 1)
    if (ncta == 2 && fwd_smem_bytes<2>(a.Mc, XD, 3) <= 227 * 1024) return launch_fwd3<XD, 2, 3>(s, tk, a, Cb, info, ntiles, hyper, npass);
    if (ncta == 2) return launch_fwd3<XD, 2, 2>(s, tk, a, Cb, info, ntiles, hyper, npass);
    return launch_fwd3<XD, 1, 2>(s, tk, a, Cb, info, ntiles, hyper, npass);
}

}  // namespace

int hm_tc_available() { return 1; }

size_t hm_tc_image_elems(int Mc, int Q) { return (size_t)Q * Mc * Mc * 2; }   // fp16 elements (hi + lo)

int hm_tc_prepare(cudaStream_t s, const double* C, const HmConsts* consts, HmTcInfo* info, void* Cb, int M, int Mp, int Mc, int Q) {
    HM_CUDA(cudaMemsetAsync(&info->cmax[0], 0, sizeof(unsigned) * HM_MAXQ, s));
    tc_absmax_kernel<<<dim3(48, (unsigned)Q), 256, 0, s>>>(C, &info->cmax[0], M, Mp);
    HM_CUDA(cudaGetLastError());
    tc_scale_kernel<<<1, Q, 0, s>>>(&info->cmax[0], consts, info);
    HM_CUDA(cudaGetLastError());
    dim3 grid((unsigned)hm_cdiv(kNB * 8, 256), (unsigned)((Mc / kNB) * (Mc / kKB)), (unsigned)Q);
    tc_image_kernel<<<grid, 256, 0, s>>>(C, info, reinterpret_cast<uint16_t*>(Cb), Mp, Mc);
    HM_CUDA(cudaGetLastError());
    return 0;
}

int hm_tc_proj_fwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, bool hyper,
                   int npass, int ncta) {
    int64_t ntiles = 0;
    for (int t = 0; t < tk.T; ++t) ntiles += hm_cdiv(tk.count[t], kRows);
    if (ntiles == 0) return 0;
    switch (a.Xdim) {
        case 1: return launch_fwd<1>(s, tk, a, Cb, info, ntiles, hyper, npass, ncta);
        case 2: return launch_fwd<2>(s, tk, a, Cb, info, ntiles, hyper, npass, ncta);
        case 3: return launch_fwd<3>(s, tk, a, Cb, info, ntiles, hyper, npass, ncta);
        case 4: return launch_fwd<4>(s, tk, a, Cb, info, ntiles, hyper, npass, ncta);
    }
    hm_set_error("Xdim=%d unsupported", a.Xdim);
    return HMOGP_ERR_ARG;
}
