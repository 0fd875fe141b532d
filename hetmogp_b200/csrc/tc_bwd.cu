// Hyper-parameter column statistics of the backward pass on the tensor cores: the TRANSPOSED projection.
//
// With P = K C_q (K = k_q(X_t, Z_q) on the fly) and the row weights mu, mu^c, omega^c of the likelihood pass, per latent q:
//     GK[n, m] = K[n,m] (mu^c[n] alpha_m + 2 omega^c[n] P[n,m])        = sum_d W'_dq dL_dKmn_d[m,n] K[n,m]
//     dz[m, i] = sum_n GK[n,m] (x_ni - z_mi)     -> inducing-input gradient of the K_mn chain (GPy RBF.gradients_X,
//                                                  /root/reference/hetmogp/svmogp.py:153-156; dL_dKmn svmogp_inf.py:157-161)
//     g1[m]    = sum_n K[n,m] mu[n]              -> dVE/dm_q = K_uu^-1 g1 (svmogp_inf.py:144)
// These are sums over DATA ROWS per INDUCING POINT, so the product is formed transposed, P^T = C_q K^T:
//     D[m (TMEM lane), n (column)] = sum_k C_q[m,k] K[n,k]      tcgen05.mma.cta_group::2, M = 256 (two 128-blocks of m,
//                                                                 one per CTA of the pair), N = 256 data rows, K = 64/stage
// and the epilogue thread that owns lane m walks the 256 rows of the super-tile: no cross-lane reduction, the sums stay
// in registers for a whole pass over the data and are written once per (CTA, m-block).
//   A operand: this CTA's 128 rows of the split-fp16 C_q image (cp.async.bulk, same image as tc_fwd.cu)
//   B operand: K tile of this CTA's 128 data rows of the super-tile, generated exactly like the A operand of tc_fwd.cu
// Roles as in tc_fwd.cu (576 threads): warps 0-7 generators, 8-15 epilogue, 16 MMA issuer (leader) / stage relay (peer),
// 17 bulk copy.  Replaces the distance-weighted Gram launch of tc_gram.cu at ~0.6x its cost (same MMA work as the
// forward pass instead of a generator-bound Gram).
#include <string.h>

#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kRows = 128;      // rows of K per CTA and super-tile half; also the m-block
constexpr int kSuper = 256;     // data rows per super-tile (pair)
constexpr int kKB = 64;
constexpr int kStages = 3;
constexpr int kTile = kRows * 128;               // 16 KB: hi (or lo) image of a 128 x 64 fp16 operand tile
constexpr int kStageBytes = 4 * kTile;           // generated K hi/lo + C block hi/lo = 64 KB
constexpr int kCHalf = 256 * 128;                // C image: 256 rows x 128 B per (h, kb), hi then lo
constexpr int kGenWarps = 8, kEpiWarps = 8, kMmaWarp = 16;
constexpr int kThreads = 576;
constexpr uint32_t kIdesc = idesc_f16(256, kSuper);

struct BwdBars {
    uint64_t full[kStages], empty[kStages], tfull[2], tempty[2];
    uint32_t tmem_base;
};

struct SuperRef { int t; int64_t row0; int nrows; };
__device__ __forceinline__ SuperRef find_super(const HmTasks& tk, int64_t st) {
    SuperRef r; r.t = 0; r.row0 = 0; r.nrows = 0;
    for (int t = 0; t < tk.T; ++t) {
        const int64_t nt = (tk.count[t] + kSuper - 1) / kSuper;
        if (st < nt) {
            r.t = t; r.row0 = st * kSuper;
            const int64_t rem = tk.count[t] - r.row0;
            r.nrows = rem < kSuper ? (int)rem : kSuper;
            return r;
        }
        st -= nt;
    }
    return r;
}

template <int XD>
__global__ void __launch_bounds__(kThreads, 1)
tc_bwd_kernel(HmTasks tk, HmProjArgs pa, const uint16_t* __restrict__ Cb, const HmTcInfo* __restrict__ info, int64_t nsuper,
              double* __restrict__ colpart, int npass) {
    constexpr int R = 2 * XD;         // generator table rows per 8 columns: -s z (hi, lo) per dim (the bias is a register; a
                                      // padded column has -s z = -1e18: d.d = 1e36 and ex2 gives exactly 0)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int Mc = pa.Mc, Mp = pa.Mp, M = pa.M, Q = pa.Q;
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // k blocks of 64 inducing points: `nkbf` in the operand image (padded M), `nkb` worth contracting over (K is exactly
    // zero on padded inducing points); pair jobs: m-blocks (2 jm, 2 jm + 1)
    const int nkbf = Mc / kKB, nkb = (M + kKB - 1) / kKB, njobs = Mc / 256;
    const uint32_t rank = cluster_ctarank();
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    uint8_t* stage_base = smem;
    float* tab = reinterpret_cast<float*>(smem + kStages * kStageBytes);          // [Mc/8][R][8] generator constants
    constexpr int kRowArr = 2 * XD + 3;            // per-row arrays of the epilogue: xh[XD] xl[XD] | mu^c | omega^c | mu
    float* rowdat = tab + (size_t)Mc * R;                                         // [kRowArr][kSuper] epilogue row data
    BwdBars* sb = reinterpret_cast<BwdBars*>(rowdat + kRowArr * kSuper);

    const HmConsts* __restrict__ cs = pa.consts;
    const double s2 = 0.5 * 1.4426950408889634 * cs->inv_l2[q];
    const double sscale = sqrt(s2);
    const int kexp = info->kexp[q], cexp = info->cexp[q];
    const float lv = (float)log2(cs->var[q]);
    for (int m = threadIdx.x; m < Mc; m += kThreads) {
        float* t8 = tab + (size_t)(m >> 3) * R * 8 + (m & 7);
        for (int i = 0; i < XD; ++i) {
            const double z = (m < M) ? pa.Zp[((size_t)q * Mp + m) * XD + i] : 0.0;
            float h, l;
            split_scaled(z, sscale, h, l);
            t8[(2 * i) * 8] = (m < M || i > 0) ? -h : -1.0e18f;
            t8[(2 * i + 1) * 8] = (m < M) ? -l : 0.f;
        }
    }
    const float2 gen_bias = dup2(-(lv + (float)kexp));
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&sb->full[s], kGenWarps + 1 + (rank == 0 ? 1 : 0)); mbar_init(&sb->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sb->tfull[b], 1); mbar_init(&sb->tempty[b], kEpiWarps * 2); }
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc2(&sb->tmem_base, 512u);
    fence_before();
    cluster_sync();
    fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    if (warp < kGenWarps) {
        // ======================================================= generators: K tile of this CTA's 128 rows (2 rows / thread)
        const int rg = warp & 1, cq = warp >> 1;
        const int r0 = rg * 64 + lane, r1 = r0 + 32;
        int stage = 0; uint32_t phase = 0;
        const int ra = (int)rank * kRows + r0, rb_ = (int)rank * kRows + r1;      // rows within the super-tile
        // the rows' inputs of the NEXT super-tile are fetched while this one is generated
        double xa_n[XD], xb_n[XD];
        auto fetch_x = [&](int64_t st) {
#pragma unroll
            for (int i = 0; i < XD; ++i) { xa_n[i] = 0.0; xb_n[i] = 0.0; }
            if (st >= nsuper) st = pair;                 // wraps to the first super-tile of the next m-block pass
            if (st >= nsuper) return;
            const SuperRef sr = find_super(tk, st);
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                if (ra < sr.nrows) xa_n[i] = tk.X[sr.t][(tk.begin[sr.t] + sr.row0 + ra) * XD + i];
                if (rb_ < sr.nrows) xb_n[i] = tk.X[sr.t][(tk.begin[sr.t] + sr.row0 + rb_) * XD + i];
            }
        };
        fetch_x(pair);
        for (int jm = 0; jm < njobs; ++jm) {
            for (int64_t st = pair; st < nsuper; st += npairs) {
                float2 xh0[XD], xl0[XD], xh1[XD], xl1[XD];
#pragma unroll
                for (int i = 0; i < XD; ++i) {
                    float h, l;
                    split_scaled(xa_n[i], sscale, h, l); xh0[i] = dup2(h); xl0[i] = dup2(l);
                    split_scaled(xb_n[i], sscale, h, l); xh1[i] = dup2(h); xl1[i] = dup2(l);
                }
                fetch_x(st + npairs);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_warp(&sb->empty[stage], phase ^ 1);
                    uint8_t* k0_hi = stage_base + (size_t)stage * kStageBytes + r0 * 128;
                    uint8_t* k1_hi = stage_base + (size_t)stage * kStageBytes + r1 * 128;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = cq * 2 + cc;
                        const float4* t4 = reinterpret_cast<const float4*>(tab + (size_t)(kb * 8 + c) * R * 8);
                        float2 e0[4], e1[4];
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
                        *reinterpret_cast<uint4*>(k0_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(k0_hi + kTile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                        for (int p = 0; p < 4; ++p) split2(ex2(-e1[p].x), ex2(-e1[p].y), hi[p], lo[p]);
                        off = (c ^ (r1 & 7)) << 4;
                        *reinterpret_cast<uint4*>(k1_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(k1_hi + kTile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sb->full[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < kGenWarps + kEpiWarps) {
        // ======================================================= epilogue: thread = TMEM lane = inducing point m of this job
        const int ew = warp - kGenWarps, lq = ew & 3, ch = ew >> 2, ml = lq * 32 + lane;
        const int et = (int)threadIdx.x - kGenWarps * 32;                  // 0..255: row of the super-tile this thread stages
        const float inv_pc = pow2i(-(kexp + cexp));
        uint32_t jc = 0, scount = 0;
        // The row this thread stages for the NEXT super-tile is fetched from HBM while the current one is processed: the
        // round trip (~1 us) used to sit between two named barriers at the head of every super-tile's epilogue, on the path
        // that decides when the MMAs get their accumulator buffer back.
        double x_n[XD];
        float mc_n = 0.f, oc_n = 0.f, mu_n = 0.f;
        auto fetch_row = [&](int64_t st) {
#pragma unroll
            for (int i = 0; i < XD; ++i) x_n[i] = 0.0;
            mc_n = oc_n = mu_n = 0.f;
            if (st >= nsuper) st = pair;                 // wraps to the first super-tile of the next m-block pass
            if (st >= nsuper) return;
            const SuperRef sn = find_super(tk, st);
            if (et >= sn.nrows) return;
            const int64_t row = sn.row0 + et;
#pragma unroll
            for (int i = 0; i < XD; ++i) x_n[i] = tk.X[sn.t][(tk.begin[sn.t] + row) * XD + i];
            const float* mw = reinterpret_cast<const float*>(tk.MW[sn.t]) + row;
            const size_t cap = (size_t)tk.cap[sn.t];
            mc_n = mw[(size_t)(2 * Q + q) * cap];        // mu^c
            oc_n = mw[(size_t)(3 * Q + q) * cap];        // omega^c
            mu_n = mw[(size_t)q * cap];                  // mu
        };
        fetch_row(pair);
        for (int jm = 0; jm < njobs; ++jm) {
            const int m = (2 * jm + (int)rank) * kRows + ml;
            float2 nzh[XD], nzl[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                const double z = (m < M) ? pa.Zp[((size_t)q * Mp + m) * XD + i] : 0.0;
                float h, l;
                split_scaled(z, sscale, h, l);
                nzh[i] = dup2(-h); nzl[i] = dup2(-l);
            }
            const float2 nb0 = dup2((m < M) ? -lv : 1.0e30f);
            const float2 al2 = dup2((m < M) ? (float)pa.alpha[(size_t)q * Mp + m] : 0.f);
            double g1 = 0.0, dz[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) dz[i] = 0.0;
            for (int64_t st = pair; st < nsuper; st += npairs, ++jc, ++scount) {
                const SuperRef sr = find_super(tk, st);
                // ---- stage the 256 rows' inputs and weights (SoA; single buffer, fenced by two named barriers)
                float* rd = rowdat;
                if (scount > 0) asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // previous super-tile fully read
                {
#pragma unroll
                    for (int i = 0; i < XD; ++i) {
                        float h, l;
                        split_scaled(x_n[i], sscale, h, l);
                        rd[i * kSuper + et] = h;
                        rd[(XD + i) * kSuper + et] = l;
                    }
                    rd[(2 * XD) * kSuper + et] = mc_n;
                    rd[(2 * XD + 1) * kSuper + et] = oc_n;
                    rd[(2 * XD + 2) * kSuper + et] = mu_n;
                    fetch_row(st + npairs);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
                const uint32_t buf = jc & 1u;
                mbar_wait_warp(&sb->tfull[buf], (jc >> 1) & 1u);
                fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + buf * kSuper + ch * (kSuper / 2);
                float2 g2 = dup2(0.f), dz2[XD];
#pragma unroll
                for (int i = 0; i < XD; ++i) dz2[i] = dup2(0.f);
#pragma unroll 1
                for (int cc = 0; cc < kSuper / 64; ++cc) {
                    uint32_t v[32];
                    tmem_ld32(taddr + cc * 32, v);
                    tmem_ld_wait();
                    const int n0 = ch * (kSuper / 2) + cc * 32;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {            // 4 rows (2 pairs) per step
                        const int n = n0 + g * 4;
                        float2 d[XD][2], u[2];
#pragma unroll
                        for (int i = 0; i < XD; ++i) {
                            const float4 h4 = *reinterpret_cast<const float4*>(rd + i * kSuper + n);
                            const float4 l4 = *reinterpret_cast<const float4*>(rd + (XD + i) * kSuper + n);
                            d[i][0] = add2(add2(make_float2(h4.x, h4.y), nzh[i]), add2(make_float2(l4.x, l4.y), nzl[i]));
                            d[i][1] = add2(add2(make_float2(h4.z, h4.w), nzh[i]), add2(make_float2(l4.z, l4.w), nzl[i]));
                            u[0] = (i == 0) ? mul2(d[i][0], d[i][0]) : fma2(d[i][0], d[i][0], u[0]);
                            u[1] = (i == 0) ? mul2(d[i][1], d[i][1]) : fma2(d[i][1], d[i][1], u[1]);
                        }
                        const float4 mc4 = *reinterpret_cast<const float4*>(rd + (2 * XD) * kSuper + n);
                        const float4 oc4 = *reinterpret_cast<const float4*>(rd + (2 * XD + 1) * kSuper + n);
                        const float4 mu4 = *reinterpret_cast<const float4*>(rd + (2 * XD + 2) * kSuper + n);
                        const float2 mc[2] = {make_float2(mc4.x, mc4.y), make_float2(mc4.z, mc4.w)};
                        const float2 oc[2] = {make_float2(oc4.x, oc4.y), make_float2(oc4.z, oc4.w)};
                        const float2 mu[2] = {make_float2(mu4.x, mu4.y), make_float2(mu4.z, mu4.w)};
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            const float2 ea = add2(u[p], nb0);
                            const float2 kv = make_float2(ex2(-ea.x), ex2(-ea.y));
                            const float2 P = mul2(make_float2(__uint_as_float(v[g * 4 + 2 * p]), __uint_as_float(v[g * 4 + 2 * p + 1])), dup2(inv_pc));
                            // GK = K (mu^c alpha + 2 omega^c P)
                            const float2 gk = mul2(kv, fma2(add2(oc[p], oc[p]), P, mul2(mc[p], al2)));
                            g2 = fma2(kv, mu[p], g2);
#pragma unroll
                            for (int i = 0; i < XD; ++i) dz2[i] = fma2(gk, d[i][p], dz2[i]);
                        }
                    }
                }
                fence_before();
                __syncwarp();
                if (lane == 0) { if (rank != 0) mbar_arrive_remote(&sb->tempty[buf], 0); else mbar_arrive(&sb->tempty[buf]); }
                g1 += (double)(g2.x + g2.y);
#pragma unroll
                for (int i = 0; i < XD; ++i) dz[i] += (double)(dz2[i].x + dz2[i].y);
            }
            // ---- one partial per (CTA pair, column half): colpart[q][slot][ (1 + XD) * Mc + 1 ]
            if (m < Mc) {
                const int ncol = (1 + XD) * Mc + 1;
                double* out = colpart + ((size_t)q * (npairs * 2) + (size_t)pair * 2 + ch) * ncol;
                out[m] = g1;
#pragma unroll
                for (int i = 0; i < XD; ++i) out[(1 + i) * Mc + m] = dz[i] / sscale;
            }
        }
    } else if (warp == kMmaWarp) {
        if (rank == 0) {
            // ======================================================= MMA issuer (leader CTA): the whole warp runs the loop, one
            // elected lane issues (elect_one: the descriptors stay in uniform registers)
            int stage = 0; uint32_t phase = 0, jc = 0;
            for (int jm = 0; jm < njobs; ++jm)
                for (int64_t st = pair; st < nsuper; st += npairs, ++jc) {
                    const uint32_t buf = jc & 1u;
                    mbar_wait_cluster(&sb->tempty[buf], ((jc >> 1) & 1u) ^ 1u);
                    fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kSuper;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait_cluster(&sb->full[stage], phase);
                        fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes);
                        const uint64_t k_hi = desc_sw128(sa), k_lo = desc_sw128(sa + kTile);                 // B: K rows (N)
                        const uint64_t c_hi = desc_sw128(sa + 2 * kTile), c_lo = desc_sw128(sa + 3 * kTile); // A: C block (M)
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < kKB / 16; ++ks) {
                                const uint64_t adv = (uint64_t)(ks * 2);
                                mma2_f16(d_tmem, c_hi + adv, k_hi + adv, kIdesc, (kb | ks) ? 1u : 0u);
                                if (npass >= 2) mma2_f16(d_tmem, c_hi + adv, k_lo + adv, kIdesc, 1u);
                                if (npass >= 3) mma2_f16(d_tmem, c_lo + adv, k_hi + adv, kIdesc, 1u);
                            }
                            commit2(&sb->empty[stage]);
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) commit2(&sb->tfull[buf]);
                    __syncwarp();
                }
        } else if (lane == 0) {
            // peer CTA: relay stage readiness to the leader
            int stage = 0; uint32_t phase = 0;
            for (int jm = 0; jm < njobs; ++jm)
                for (int64_t st = pair; st < nsuper; st += npairs)
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&sb->full[stage], phase);
                        mbar_arrive_remote(&sb->full[stage], 0);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
        }
    } else {
        // ======================================================= bulk-copy producer: this CTA's 128 rows of the C_q image
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const int nhalf = Mc / 256;
            for (int jm = 0; jm < njobs; ++jm) {
                // m-block 2 jm + rank = rows [rank * 128, +128) of column block h = jm of the image
                for (int64_t st = pair; st < nsuper; st += npairs)
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&sb->empty[stage], phase ^ 1);
                        uint8_t* dst = stage_base + (size_t)stage * kStageBytes + 2 * kTile;
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(Cb) + ((size_t)(q * nhalf + jm) * nkbf + kb) * (2 * kCHalf) + rank * kTile;
                        mbar_expect_tx(&sb->full[stage], 2 * kTile);
                        bulk_g2s(dst, src, kTile, &sb->full[stage]);                    // hi
                        bulk_g2s(dst + kTile, src + kCHalf, kTile, &sb->full[stage]);   // lo
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    }
    fence_before();
    cluster_sync();
    if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512u);
}

size_t bwd_smem_bytes(int Mc, int Xd) {
    return (size_t)kStages * kStageBytes + sizeof(float) * ((size_t)Mc * (2 * Xd) + (2 * Xd + 3) * kSuper) + sizeof(BwdBars) + 64 + 1024;
}

template <int XD>
int launch_bwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, int64_t nsuper,
               double* colpart, int nslots, int npass) {
    const size_t smem = bwd_smem_bytes(a.Mc, XD);
    if (smem > 227 * 1024) {
        hm_set_error("tensor-core backward projection: M=%d (padded %d) with Xdim=%d needs %zu B of shared memory", a.M, a.Mc, XD, smem);
        return HMOGP_ERR_ARG;
    }
    auto kern = tc_bwd_kernel<XD>;
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)nslots, (unsigned)a.Q);     // nslots = 2 * pairs (even); every slot is written
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HM_CUDA(cudaLaunchKernelEx(&cfg, kern, tk, a, reinterpret_cast<const uint16_t*>(Cb), info, nsuper, colpart, npass));
    HM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// colpart: [Q][nslots][(1 + Xdim) * Mc + 1] doubles (the last entry of a slot is unused here), nslots even
int hm_tc_proj_bwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, double* colpart,
                   int nslots, int npass) {
    int64_t nsuper = 0;
    for (int t = 0; t < tk.T; ++t) nsuper += hm_cdiv(tk.count[t], kSuper);
    switch (a.Xdim) {
        case 1: return launch_bwd<1>(s, tk, a, Cb, info, nsuper, colpart, nslots, npass);
        case 2: return launch_bwd<2>(s, tk, a, Cb, info, nsuper, colpart, nslots, npass);
        case 3: return launch_bwd<3>(s, tk, a, Cb, info, nsuper, colpart, nslots, npass);
        case 4: return launch_bwd<4>(s, tk, a, Cb, info, nsuper, colpart, nslots, npass);
    }
    hm_set_error("Xdim=%d unsupported", a.Xdim);
    return HMOGP_ERR_ARG;
}
