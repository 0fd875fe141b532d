// Backward statistics on the 5th-generation tensor cores: weighted Grams of the on-the-fly RBF cross-covariance.
//
// For latent q, with K = k_q(X_t, Z_q) regenerated per 32-row chunk (never in HBM) and a row weight w[n]:
//     H_q[i, j] = sum_t sum_n w[n] K[n,i] K[n,j]            (M x M; lower block-triangle computed)
//     g^v_q[i]  = sum_t sum_n v_v[n] K[n,i]
// w = omega_tq  -> H^1, from which dVE/dS_q = K_uu^-1 H^1 K_uu^-1 (reference: A^T diag(dv) A per output function,
// /root/reference/hetmogp/svmogp_inf.py:145-148, summed over d with W_dq^2 folded into omega);
// g^mu = K^T mu gives dVE/dm_q (svmogp_inf.py:144); in a full step it comes from tc_bwd.cu together with the
// inducing-input statistic, so the Gram launch then carries no g-vector.  (The kernel can also weight the A operand
// by the signed distance s (x_ni - z_mi) of the row it owns -- template flag DIST -- which an earlier revision used to
// obtain dZ from a second Gram; the transposed projection of tc_bwd.cu replaced it at 0.6x the cost.)
//
// MMA: D[i (128 TMEM lanes), j (<=256 columns)] += A[i][n] . B[j][n]^T over n = 32 data rows per stage,
//   A = 2^wexp w[n] (d_i) K[n, I-block]   B = 2^kexp K[n, J-block]   (split fp16, 3 products).
// Accumulation is three-level, because tcgen05 accumulates with truncation and white noise in H is amplified
//   ~cond(K_uu)^1.4 by K_uu^-1 . K_uu^-1:
//   level 1  TMEM columns [0,256): tensor-core fp32 accumulator over `f1` chunks (512 rows)
//   level 2  TMEM columns [256,512): fp32, round-to-nearest SIMT adds of level 1 (tcgen05.ld / add / tcgen05.st)
//   level 3  fp64 partial tile of this CTA in HBM/L2, every `f2` level-1 windows (16 K rows); summed in a fixed order
//            by tc_gram_reduce_kernel -> deterministic.
//
// Warp roles (608 threads, 1 CTA/SM, persistent over a host-built plan of (q, tile, row-range) segments):
//   warps 0-7   B generators (thread = column j of the tile, all 32 rows of the chunk)
//   warps 8-15  A generators (thread = column i; the A operand costs ~1.5x per element -- weights, g-vectors -- and gates
//               every stage, so each A column group is split over two warps, 16 rows each)
//   warp 16     MMA issuer (one thread)                          warps 17-18 row loaders (x, weights -> smem ring)
//   Generators also run the level-2/3 flushes (4 TMEM lane quadrants x 4 column groups).
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kGC = HM_GRAM_CHUNK;                       // data rows per chunk = MMA K extent per stage (2 x K16)
constexpr int kGStages = 4;
constexpr int kGBHalf = 256 * 64;                        // 16 KB : B hi (or lo), 256 rows x 64 B (SW64)
constexpr int kGAHalf = 128 * 64;                        //  8 KB : A hi (or lo)
constexpr int kGStageBytes = 2 * kGBHalf + 2 * kGAHalf;  // 48 KB
constexpr int kRowSlots = 8;
constexpr int kRowArrays = 16;   // arrays per slot, each [kGC] floats (SoA): xh[XD] | xl[XD] | w | v0..v4
constexpr int kSplitA = HM_GRAM_ROWSPLIT;      // A-generator warps per column group: each takes kGC / kSplitA rows of a chunk
constexpr int kGenWarps = 8 + 4 * kSplitA;     // 8 B column groups (all rows) + 4 A column groups x row parts
constexpr int kMmaWarp = kGenWarps;            // + 2 row-loader warps
constexpr int kGThreads = (kGenWarps + 3) * 32;
constexpr uint32_t kAcc2 = 256;  // TMEM column of the level-2 accumulator

struct GramBars {
    uint64_t full[kGStages], empty[kGStages], rowfull[kRowSlots], rowempty[kRowSlots], accfull, accempty;
    uint32_t tmem_base;
};

__device__ __forceinline__ float weight_scale(const HmTcInfo* info, const HmConsts* cs, int q, int base) {
    // distance-weighted operands: |s d| K <= 0.43 sigma^2, so the plain bound max|w| sigma^2 holds for them too
    const float amax = __uint_as_float(info->wmax[base == 3 ? 1 : 0][q]) * (float)cs->var[q];
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e = 0;
    frexpf(amax, &e);   // amax < 2^e
    return pow2i(14 - e);
}

template <int XD> struct RowX { float2 h[XD][4], l[XD][4]; };   // 8 rows of scaled split inputs, as 4 row pairs

template <int XD>
__device__ __forceinline__ void load_rowx(RowX<XD>& r, const float* rb, int n8) {
#pragma unroll
    for (int i = 0; i < XD; ++i) {
        const float4 h0 = *reinterpret_cast<const float4*>(rb + i * kGC + n8 * 8);
        const float4 h1 = *reinterpret_cast<const float4*>(rb + i * kGC + n8 * 8 + 4);
        const float4 l0 = *reinterpret_cast<const float4*>(rb + (XD + i) * kGC + n8 * 8);
        const float4 l1 = *reinterpret_cast<const float4*>(rb + (XD + i) * kGC + n8 * 8 + 4);
        r.h[i][0] = make_float2(h0.x, h0.y); r.h[i][1] = make_float2(h0.z, h0.w);
        r.h[i][2] = make_float2(h1.x, h1.y); r.h[i][3] = make_float2(h1.z, h1.w);
        r.l[i][0] = make_float2(l0.x, l0.y); r.l[i][1] = make_float2(l0.z, l0.w);
        r.l[i][2] = make_float2(l1.x, l1.y); r.l[i][3] = make_float2(l1.z, l1.w);
    }
}
__device__ __forceinline__ void load8(float2 (&o)[4], const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = make_float2(a.x, a.y); o[1] = make_float2(a.z, a.w); o[2] = make_float2(b.x, b.y); o[3] = make_float2(b.z, b.w);
}

template <int XD, int NV, bool DIST>
__global__ void __launch_bounds__(kGThreads, 1)
tc_gram_kernel(HmTasks tk, HmProjArgs pa, const HmTcInfo* __restrict__ info, const HmGramSeg* __restrict__ segs,
               const int* __restrict__ seg_off, HmGramWeights gw, double* __restrict__ slots, int f1, int f2, int npass) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    float* rowbuf = reinterpret_cast<float*>(smem + kGStages * kGStageBytes);   // [kRowSlots][kRowArrays][kGC]
    GramBars* sb = reinterpret_cast<GramBars*>(rowbuf + kRowSlots * kGC * kRowArrays);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Mp = pa.Mp, M = pa.M, Q = pa.Q;
    const HmConsts* __restrict__ cs = pa.consts;
    const int seg_begin = seg_off[blockIdx.x], seg_end = seg_off[blockIdx.x + 1];

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGStages; ++s) { mbar_init(&sb->full[s], kGenWarps); mbar_init(&sb->empty[s], 1); }
        for (int s = 0; s < kRowSlots; ++s) { mbar_init(&sb->rowfull[s], 1); mbar_init(&sb->rowempty[s], kGenWarps); }
        mbar_init(&sb->accfull, 1);
        mbar_init(&sb->accempty, kGenWarps);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&sb->tmem_base, 512u);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    if (warp < kGenWarps) {
        // ======================================================= generators (+ level-2/3 flushes)
        const bool isA = warp >= 8;                                        // warps 0-7: B column groups; 8..: A (column group, row part)
        const int cg = isA ? 8 + (warp - 8) % 4 : warp, rh = isA ? (warp - 8) / 4 : 0;
        const int col = (isA ? cg - 8 : cg) * 32 + lane;                   // row of the operand tile this thread writes
        constexpr int kN8A = kGC / 8 / kSplitA;                            // groups of 8 rows per A thread and chunk
        const int swz = (col >> 1) & 3;                                    // SW64: chunk ^= (row >> 1) & 3
        const int wdim = gw.wdim[0];
        uint32_t cc_ = 0;     // chunk counter (stage / row-slot rings)
        uint32_t iv = 0;      // level-1 window counter (accumulator barriers)
        struct Pending { bool on, to_l3, slot_fresh, acc2_fresh; int ncc; double* slot; double inv_sc; uint32_t parity; };
        Pending pend;
        pend.on = false;
        // fold one finished level-1 window (TMEM cols [0,256)) into level 2 (TMEM cols [256,512), fp32 round-to-nearest)
        // or, every f2 windows / at the end of a segment, level 1 + level 2 into the fp64 partial tile
        auto flush_window = [&](const Pending& pd) {
            mbar_wait_warp(&sb->accfull, pd.parity);
            fence_after();
            const int lq = warp & 3, wq = warp >> 2;
            const int i = lq * 32 + lane;
            const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16);
            for (int c16 = wq; c16 < 2 * pd.ncc; c16 += kGenWarps / 4) {   // units of 16 columns
                uint32_t v[16];
                tmem_ld16(tl + c16 * 16, v);
                if (!pd.acc2_fresh) {
                    uint32_t u[16];
                    tmem_ld16(tl + kAcc2 + c16 * 16, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int p = 0; p < 16; ++p) v[p] = __float_as_uint(__uint_as_float(v[p]) + __uint_as_float(u[p]));
                } else {
                    tmem_ld_wait();
                }
                if (!pd.to_l3) {
                    tmem_st16(tl + kAcc2 + c16 * 16, v);
                } else {
                    double2* dst = reinterpret_cast<double2*>(pd.slot + ((size_t)i * 256 + c16 * 16));
                    if (pd.slot_fresh) {
#pragma unroll
                        for (int p = 0; p < 8; ++p)
                            dst[p] = make_double2((double)__uint_as_float(v[2 * p]) * pd.inv_sc, (double)__uint_as_float(v[2 * p + 1]) * pd.inv_sc);
                    } else {
#pragma unroll
                        for (int p = 0; p < 8; ++p) {
                            double2 o = dst[p];
                            o.x += (double)__uint_as_float(v[2 * p]) * pd.inv_sc;
                            o.y += (double)__uint_as_float(v[2 * p + 1]) * pd.inv_sc;
                            dst[p] = o;
                        }
                    }
                }
            }
            if (!pd.to_l3) tmem_st_wait();
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sb->accempty);
        };
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const double sscale = sqrt(0.5 * 1.4426950408889634 * cs->inv_l2[q]);
            const int gcol = isA ? sg.I * 128 + col : sg.j0 + col;
            const bool active = isA || col < sg.nw;
            float2 nzh[XD], nzl[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                const double z = (gcol < M) ? pa.Zp[((size_t)q * Mp + gcol) * XD + i] : 0.0;
                float zh, zl;
                split_scaled(z, sscale, zh, zl);
                nzh[i] = dup2(-zh); nzl[i] = dup2(-zl);
            }
            const float lv = (float)log2(cs->var[q]);
            // K = ex2(-(d.d - bias)); A uses the unscaled kernel, B carries 2^kexp; padded columns give 0
            const float2 nb2 = dup2((gcol < M) ? -(isA ? lv : lv + (float)info->kexp[q]) : 1.0e30f);
            double g64[NV > 0 ? NV : 1];
#pragma unroll
            for (int v = 0; v < NV; ++v) g64[v] = 0.0;
            const double inv_sc = 1.0 / ((double)weight_scale(info, cs, q, gw.wbase[0]) * (double)pow2i(info->kexp[q]));
            double* slot = slots + (size_t)sg.slot * HM_GRAM_SLOT_DOUBLES;
            bool slot_fresh = true;    // level 3: first flush of the segment stores, later ones add
            bool acc2_fresh = true;    // level 2: holds nothing yet
            int win = 0;

            for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += f1, ++win) {
                const int c1 = min(sg.chunk_end, c0 + f1);
                for (int c = c0; c < c1; ++c, ++cc_) {
                    const int stage = cc_ % kGStages, rs = cc_ % kRowSlots;
                    if (pend.on) {   // fold the previous window as soon as its MMAs are done, or before we would block on them
                        if (mbar_test(&sb->accfull, pend.parity) || !mbar_test(&sb->empty[stage], ((cc_ / kGStages) & 1u) ^ 1u)) {
                            flush_window(pend);
                            pend.on = false;
                        }
                    }
                    mbar_wait_warp(&sb->rowfull[rs], (cc_ / kRowSlots) & 1u);
                    mbar_wait_warp(&sb->empty[stage], ((cc_ / kGStages) & 1u) ^ 1u);
                    const float* rb = rowbuf + (size_t)rs * kGC * kRowArrays;   // SoA: array a at rb + a * kGC
                    uint8_t* st = stage_base + (size_t)stage * kGStageBytes;
                    if (!isA) {
                        if (active) {
                            uint8_t* b_hi = st + col * 64;
                            uint8_t* b_lo = b_hi + kGBHalf;
                            auto gen_b = [&](const RowX<XD>& r, int n8) {
                                float2 e[4] = {nb2, nb2, nb2, nb2};   // d.d - bias, 4 row pairs
#pragma unroll
                                for (int i = 0; i < XD; ++i)
#pragma unroll
                                    for (int p = 0; p < 4; ++p) {
                                        const float2 d = add2(add2(r.h[i][p], nzh[i]), add2(r.l[i][p], nzl[i]));
                                        e[p] = fma2(d, d, e[p]);
                                    }
                                uint32_t hi[4], lo[4];
#pragma unroll
                                for (int p = 0; p < 4; ++p) split2(ex2(-e[p].x), ex2(-e[p].y), hi[p], lo[p]);
                                const int off = (n8 ^ swz) << 4;
                                *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            };
                            static_assert(kGC == 32, "generator pipeline is written for 4 groups of 8 rows");
                            RowX<XD> r0, r1;          // software pipeline: the next 8 rows are in flight while 8 are processed
                            load_rowx<XD>(r0, rb, 0);
                            load_rowx<XD>(r1, rb, 1);
                            gen_b(r0, 0);
                            load_rowx<XD>(r0, rb, 2);
                            gen_b(r1, 1);
                            load_rowx<XD>(r1, rb, 3);
                            gen_b(r0, 2);
                            gen_b(r1, 3);
                        }
                    } else {
                        uint8_t* a_hi = st + 2 * kGBHalf + col * 64;
                        float2 g2[NV > 0 ? NV : 1];
#pragma unroll
                        for (int v = 0; v < NV; ++v) g2[v] = dup2(0.f);
                        auto gen_a = [&](const RowX<XD>& r, int n8) {
                            float2 w[4];
                            load8(w, rb + (2 * XD) * kGC + n8 * 8);
                            float2 e[4] = {nb2, nb2, nb2, nb2};
                            float2 dd[XD][4];   // signed scaled distances (distance-weighted operands)
#pragma unroll
                            for (int i = 0; i < XD; ++i)
#pragma unroll
                                for (int p = 0; p < 4; ++p) {
                                    dd[i][p] = add2(add2(r.h[i][p], nzh[i]), add2(r.l[i][p], nzl[i]));
                                    e[p] = fma2(dd[i][p], dd[i][p], e[p]);
                                }
                            float2 kv[4], a0[4];
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                kv[p] = make_float2(ex2(-e[p].x), ex2(-e[p].y));
                                float2 f = kv[p];
                                if (DIST) {
                                    float2 ds = dd[0][p];
#pragma unroll
                                    for (int i = 1; i < XD; ++i) ds = (wdim == i) ? dd[i][p] : ds;
                                    f = mul2(kv[p], ds);
                                }
                                a0[p] = mul2(f, w[p]);
                            }
                            uint32_t hi[4], lo[4];
                            const int off = (n8 ^ swz) << 4;
#pragma unroll
                            for (int p = 0; p < 4; ++p) split2(a0[p].x, a0[p].y, hi[p], lo[p]);
                            *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(a_hi + kGAHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            if (NV > 0) {
                                if (sg.has_g) {
#pragma unroll
                                    for (int v = 0; v < NV; ++v) {
                                        float2 vv[4];
                                        load8(vv, rb + (2 * XD + 1 + v) * kGC + n8 * 8);
#pragma unroll
                                        for (int p = 0; p < 4; ++p)   // v = 0: g^mu (plain); v = 1 + i: distance-weighted in dim i
                                            g2[v] = fma2(v == 0 ? kv[p] : mul2(kv[p], dd[v > 0 ? v - 1 : 0][p]), vv[p], g2[v]);
                                    }
                                }
                            }
                        };
                        RowX<XD> r0, r1;
                        static_assert(kN8A == 2 || kN8A == 4, "A generator pipeline: 2 or 4 groups of 8 rows per thread");
                        const int nb = rh * kN8A;
                        load_rowx<XD>(r0, rb, nb);
                        load_rowx<XD>(r1, rb, nb + 1);
                        gen_a(r0, nb);
                        if (kN8A == 4) load_rowx<XD>(r0, rb, nb + 2);
                        gen_a(r1, nb + 1);
                        if (kN8A == 4) {
                            load_rowx<XD>(r1, rb, nb + 3);
                            gen_a(r0, nb + 2);
                            gen_a(r1, nb + 3);
                        }
                        if (NV > 0 && sg.has_g) {
#pragma unroll
                            for (int v = 0; v < NV; ++v) g64[v] += (double)(g2[v].x + g2[v].y);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&sb->full[stage]);
                        mbar_arrive(&sb->rowempty[rs]);
                    }
                }
                // ---- end of a level-1 window: its fold into level 2 / 3 is deferred (see flush_window) so that the
                //      generators keep the smem ring full while the MMAs of the window drain
                if (pend.on) flush_window(pend);
                pend.on = true;
                pend.to_l3 = ((win + 1) % f2 == 0) || (c1 == sg.chunk_end);
                pend.slot_fresh = slot_fresh; pend.acc2_fresh = acc2_fresh;
                pend.ncc = sg.nw / 32; pend.slot = slot; pend.inv_sc = inv_sc; pend.parity = iv & 1u;
                if (pend.to_l3) { slot_fresh = false; acc2_fresh = true; } else acc2_fresh = false;
                ++iv;
            }
            if (NV > 0 && isA && sg.has_g) {   // one partial per row part; tc_gram_reduce_kernel adds them
                double* gdst = slot + (size_t)128 * 256 + (size_t)rh * HM_GRAM_MAXV * 128;
#pragma unroll
                for (int v = 0; v < NV; ++v) gdst[v * 128 + col] = g64[v];
            }
        }
        if (pend.on) flush_window(pend);
    } else if (warp == kMmaWarp) {
        // ======================================================= MMA issuer (one thread)
        if (lane == 0) {
            uint32_t cc_ = 0, iv = 0;
            for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
                const HmGramSeg sg = segs[sgi];
                const uint32_t idesc = idesc_f16(128, sg.nw);
                for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += f1) {
                    const int c1 = min(sg.chunk_end, c0 + f1);
                    mbar_wait(&sb->accempty, (iv & 1u) ^ 1u);
                    fence_after();
                    for (int c = c0; c < c1; ++c, ++cc_) {
                        const int stage = cc_ % kGStages;
                        mbar_wait(&sb->full[stage], (cc_ / kGStages) & 1u);
                        fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kGStageBytes);
                        const uint64_t b_hi = desc_sw64(sa), b_lo = desc_sw64(sa + kGBHalf);
                        const uint64_t a_hi = desc_sw64(sa + 2 * kGBHalf), a_lo = desc_sw64(sa + 2 * kGBHalf + kGAHalf);
#pragma unroll
                        for (int ks = 0; ks < kGC / 16; ++ks) {
                            const uint64_t adv = (uint64_t)(ks * 2);
                            mma_f16(tmem_base, a_hi + adv, b_hi + adv, idesc, (c > c0 || ks > 0) ? 1u : 0u);
                            if (npass >= 2) mma_f16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                            if (npass >= 3) mma_f16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
                        }
                        commit(&sb->empty[stage]);
                    }
                    commit(&sb->accfull);
                    ++iv;
                }
            }
        }
    } else {
        // ======================================================= row loaders (2 warps, alternating groups of 4 chunks)
        // Each iteration issues the global loads of 4 chunks (128 rows) before touching the ring: memory-level
        // parallelism instead of one exposed HBM/L2 latency per chunk.
        const int rw = warp - (kMmaWarp + 1);
        constexpr int kGrp = 4;
        int nch[HM_MAXT];
        for (int t = 0; t < HM_MAXT; ++t) nch[t] = (t < tk.T) ? (int)((tk.count[t] + kGC - 1) / kGC) : 0;
        uint32_t cc_ = 0, grp = 0;
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const double sscale = sqrt(0.5 * 1.4426950408889634 * cs->inv_l2[q]);
            const float wsc = weight_scale(info, cs, q, gw.wbase[0]);
            // (task, chunk-in-task) of the segment's first chunk
            int t = 0, ct = sg.chunk_begin;
            while (t < tk.T && ct >= nch[t]) { ct -= nch[t]; ++t; }
            for (int c = sg.chunk_begin; c < sg.chunk_end; c += kGrp, ++grp) {
                const int ng = min(kGrp, sg.chunk_end - c);
                const bool mine = (int)(grp & 1u) == rw;
                double xv[kGrp][XD];
                float wv[kGrp], vv[kGrp][NV > 0 ? NV : 1];
#pragma unroll
                for (int j = 0; j < kGrp; ++j) {
                    if (j < ng) {
                        const int64_t row = (int64_t)ct * kGC + lane;
                        const bool valid = t < tk.T && row < tk.count[t];
                        if (mine) {
#pragma unroll
                            for (int i = 0; i < XD; ++i) xv[j][i] = valid ? tk.X[t][(tk.begin[t] + row) * XD + i] : 0.0;
                            const float* mw = reinterpret_cast<const float*>(tk.MW[t]) + row;   // SoA: array k at k * cap
                            const size_t cap = (size_t)tk.cap[t < tk.T ? t : 0];
                            wv[j] = valid ? mw[(size_t)(gw.wbase[0] * Q + q) * cap] : 0.f;
#pragma unroll
                            for (int v = 0; v < NV; ++v) vv[j][v] = valid ? mw[(size_t)(gw.vbase[v] * Q + q) * cap] : 0.f;
                        }
                        if (++ct >= nch[t]) { ct = 0; ++t; while (t < tk.T && nch[t] == 0) ++t; }
                    }
                }
                if (mine) {
#pragma unroll
                    for (int j = 0; j < kGrp; ++j) {
                        if (j < ng) {
                            const uint32_t cj = cc_ + j;
                            const int rs = cj % kRowSlots;
                            float xh_[XD], xl_[XD];
#pragma unroll
                            for (int i = 0; i < XD; ++i) split_scaled(xv[j][i], sscale, xh_[i], xl_[i]);
                            mbar_wait_warp(&sb->rowempty[rs], ((cj / kRowSlots) & 1u) ^ 1u);
                            float* dst = rowbuf + (size_t)rs * kGC * kRowArrays + lane;   // SoA: array a, row = lane
#pragma unroll
                            for (int i = 0; i < XD; ++i) { dst[i * kGC] = xh_[i]; dst[(XD + i) * kGC] = xl_[i]; }
                            dst[(2 * XD) * kGC] = wv[j] * wsc;
#pragma unroll
                            for (int v = 0; v < NV; ++v) dst[(2 * XD + 1 + v) * kGC] = vv[j][v];
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sb->rowfull[rs]);
                        }
                    }
                }
                cc_ += ng;
            }
        }
    }
    // ---- teardown
    fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512u);
}

// Sum the partial tiles of every (q, tile job) in slot order; write H (lower from the tile; mirrored if symmetric) and g^v.
__global__ void tc_gram_reduce_kernel(const double* __restrict__ slots, const HmGramJob* __restrict__ jobs,
                                      const int2* __restrict__ jobslots, int njobs, int symmetric, int nV, double* H,
                                      double* g0, int64_t gstride, int M, int Mp) {
    const int job = blockIdx.x, q = blockIdx.y;
    const HmGramJob jb = jobs[job];
    const int2 sr = jobslots[q * njobs + job];
    for (int e = blockIdx.z * 8 * jb.nw + threadIdx.x; e < (blockIdx.z + 1) * 8 * jb.nw; e += blockDim.x) {   // 8 rows per CTA
        const int i = e / jb.nw, j = e % jb.nw;
        const int gr = jb.I * 128 + i, gc = jb.j0 + j;
        if (gc > gr || gr >= M) continue;
        double s = 0.0;
        for (int sl = sr.x; sl < sr.y; ++sl) s += slots[(size_t)sl * HM_GRAM_SLOT_DOUBLES + ((size_t)i * 256 + j)];
        H[((size_t)q * Mp + gr) * Mp + gc] = s;
        if (symmetric) H[((size_t)q * Mp + gc) * Mp + gr] = s;   // plain Grams are symmetric; D^i keeps its lower triangle
    }
    if (jb.j0 == 0 && blockIdx.z == 0) {
        for (int e = threadIdx.x; e < nV * 128; e += blockDim.x) {
            const int v = e / 128, i = e % 128;
            if (jb.I * 128 + i >= M) continue;
            double s = 0.0;
            for (int sl = sr.x; sl < sr.y; ++sl)
                for (int rh = 0; rh < HM_GRAM_ROWSPLIT; ++rh)
                    s += slots[(size_t)sl * HM_GRAM_SLOT_DOUBLES + (size_t)128 * 256 + (size_t)(rh * HM_GRAM_MAXV + v) * 128 + i];
            g0[(size_t)v * gstride + (size_t)q * Mp + jb.I * 128 + i] = s;
        }
    }
}

size_t gram_smem_bytes() {
    return (size_t)kGStages * kGStageBytes + sizeof(float) * kRowSlots * kGC * kRowArrays + sizeof(GramBars) + 64 + 1024;
}

template <int XD, int NV, bool DIST>
int launch_gram3(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                 const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int f1, int f2, int npass) {
    const size_t smem = gram_smem_bytes();
    HM_CUDA(cudaFuncSetAttribute(tc_gram_kernel<XD, NV, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gram_kernel<XD, NV, DIST><<<nctas, kGThreads, smem, s>>>(tk, a, info, segs, seg_off, gw, slots, f1, f2, npass);
    HM_CUDA(cudaGetLastError());
    return 0;
}

// launches the engine issues: (NV = 1) VE step (H^1 and g^mu); (NV = 0) full step (H^1 only)
template <int XD>
int launch_gram(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int f1, int f2, int npass) {
    const bool dist = gw.wdim[0] >= 0;
    if (!dist && gw.nV == 1) return launch_gram3<XD, 1, false>(s, tk, a, info, segs, seg_off, gw, slots, nctas, f1, f2, npass);
    if (!dist && gw.nV == 0) return launch_gram3<XD, 0, false>(s, tk, a, info, segs, seg_off, gw, slots, nctas, f1, f2, npass);
    hm_set_error("gram launch: unsupported (dist=%d, nV=%d) for Xdim=%d", (int)dist, gw.nV, XD);
    return HMOGP_ERR_ARG;
}

}  // namespace

int hm_tc_gram(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
               const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int f1, int f2, int npass) {
    if (gw.nW != 1) { hm_set_error("gram launch: one weight per launch"); return HMOGP_ERR_ARG; }
    switch (a.Xdim) {
        case 1: return launch_gram<1>(s, tk, a, info, segs, seg_off, gw, slots, nctas, f1, f2, npass);
        case 2: return launch_gram<2>(s, tk, a, info, segs, seg_off, gw, slots, nctas, f1, f2, npass);
        case 3: return launch_gram<3>(s, tk, a, info, segs, seg_off, gw, slots, nctas, f1, f2, npass);
        case 4: return launch_gram<4>(s, tk, a, info, segs, seg_off, gw, slots, nctas, f1, f2, npass);
    }
    hm_set_error("Xdim=%d unsupported", a.Xdim);
    return HMOGP_ERR_ARG;
}

int hm_tc_gram_reduce(cudaStream_t s, const double* slots, const HmGramJob* jobs, const int2* jobslots, int njobs, int Q,
                      const HmGramWeights& gw, double* H, double* g0, int64_t gstride, int M, int Mp) {
    dim3 grid((unsigned)njobs, (unsigned)Q, 16u);
    tc_gram_reduce_kernel<<<grid, 256, 0, s>>>(slots, jobs, jobslots, njobs, gw.wdim[0] < 0 ? 1 : 0, gw.nV, H, g0, gstride, M, Mp);
    HM_CUDA(cudaGetLastError());
    return 0;
}
