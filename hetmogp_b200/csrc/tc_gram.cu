// Backward statistics on the 5th-generation tensor cores: weighted Grams of the on-the-fly RBF cross-covariance.
//
// For latent q, with K = k_q(X_t, Z_q) regenerated per 32-row chunk (never in HBM) and row weights w_k[n]:
//     H^k_q[i, j] = sum_t sum_n w_k[n] K[n,i] K[n,j]        (M x M symmetric; lower block-triangle computed)
//     g^v_q[i]    = sum_t sum_n v_v[n] K[n,i]
// k = 0: w = omega_tq  -> H^1, from which dVE/dS_q = K_uu^-1 H^1 K_uu^-1 (reference: A^T diag(dv) A per output
// function, /root/reference/hetmogp/svmogp_inf.py:145-148, summed over d with W_dq^2 folded into omega);
// k >= 1: distance-weighted D^i_q[m, j] = sum_n omega^c[n] s (x_ni - z_mi) K[n,m] K[n,j]  (the weight depends on the
// output ROW m, which the A-operand generator -- thread = column m of K -- applies for free and without the
// cancellation of H^{x_i} - z_mi H^1); with H^1 it gives the inducing-input gradient of the K_mn chain
// (svmogp.py:153-156, GPy RBF.gradients_X) without ever forming dL_dKmn (M x N per (q,d) in the reference,
// svmogp_inf.py:157-161).  D^i is not symmetric, but D[m,j] - D[j,m] = s (z_j - z_m) H^1[m,j], so the lower
// block-triangle suffices.  g^mu gives dVE/dm_q (svmogp_inf.py:144).
//
// MMA: D[i (128 TMEM lanes), j (<=256 columns)] += A_k[i][n] . B[j][n]^T over n = 32 data rows per stage,
//   A_k = 2^wexp_k w_k[n] K[n, I-block]   B = 2^kexp K[n, J-block]   (split fp16, 3 products), up to two weights
//   per launch share one generated B tile; accumulators: 2 x 256 fp32 TMEM columns.
// Accumulation: TMEM fp32 over `flush_chunks` chunks, then added in fp64 to this CTA's private partial tile
//   (white rounding noise in H is amplified ~cond(K_uu)^1.4 by K_uu^-1 . K_uu^-1, so the fp32 window is kept short);
//   partials are summed in a fixed order by tc_gram_reduce_kernel -> deterministic.
//
// Warp roles (480 threads, 1 CTA/SM, persistent over a host-built plan of (q, tile, row-range) segments):
//   warps 0-7   B generators (thread = column j of the tile)     warps 8-11  A generators (thread = column i)
//   warp 12     MMA issuer (one thread)                          warps 13-14 row loaders (x, weights -> smem ring)
//   generators also drain TMEM at flush points (12 warps: 4 lane quadrants x 3 column groups).
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kGC = HM_GRAM_CHUNK;                       // data rows per chunk = MMA K extent per stage (2 x K16)
constexpr int kGStages = 3;
constexpr int kGBHalf = 256 * 64;                        // 16 KB : B hi (or lo), 256 rows x 64 B (SW64)
constexpr int kGAHalf = 128 * 64;                        //  8 KB : A_k hi (or lo)
constexpr int kGStageBytes = 2 * kGBHalf + 4 * kGAHalf;  // 64 KB
constexpr int kRowSlots = 8;
constexpr int kRowFloats = 16;   // arrays per slot, each [kGC] floats (SoA): xh[XD] | xl[XD] | w0 w1 | v0..v5
constexpr int kGThreads = 480;
constexpr int kGenWarps = 12;

struct GramBars {
    uint64_t full[kGStages], empty[kGStages], rowfull[kRowSlots], rowempty[kRowSlots], accfull, accempty;
    uint32_t tmem_base;
};

struct ChunkRef { int t; int64_t row0; };
__device__ __forceinline__ ChunkRef find_chunk(const HmTasks& tk, int c) {
    ChunkRef r; r.t = 0; r.row0 = 0;
    for (int t = 0; t < tk.T; ++t) {
        const int nc = (int)((tk.count[t] + kGC - 1) / kGC);
        if (c < nc) { r.t = t; r.row0 = (int64_t)c * kGC; return r; }
        c -= nc;
    }
    r.t = -1;
    return r;
}

__device__ __forceinline__ float weight_scale(const HmTcInfo* info, const HmTasks& tk, const HmConsts* cs, int q, int base, int dim) {
    (void)tk; (void)dim;   // distance-weighted operands: |s d| K <= 0.43 sigma^2, the plain bound holds
    float amax = __uint_as_float(info->wmax[base == 3 ? 1 : 0][q]) * (float)cs->var[q];
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e = 0;
    frexpf(amax, &e);   // amax < 2^e
    return pow2i(14 - e);
}

template <int XD, int NW, int NV>
__global__ void __launch_bounds__(kGThreads, 1)
tc_gram_kernel(HmTasks tk, HmProjArgs pa, const HmTcInfo* __restrict__ info, const HmGramSeg* __restrict__ segs,
               const int* __restrict__ seg_off, HmGramWeights gw, double* __restrict__ slots, int flush_chunks, int npass) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    float* rowbuf = reinterpret_cast<float*>(smem + kGStages * kGStageBytes);   // [kRowSlots][kGC][kRowFloats]
    GramBars* sb = reinterpret_cast<GramBars*>(rowbuf + kRowSlots * kGC * kRowFloats);

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
    if (warp == 12) tmem_alloc(&sb->tmem_base, 512u);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    if (warp < kGenWarps) {
        // ======================================================= generators (+ TMEM drain at flush points)
        const bool isA = warp >= 8;
        const int col = isA ? (int)threadIdx.x - 256 : (int)threadIdx.x;   // row of the operand tile this thread writes
        const int swz = (col >> 1) & 3;                                    // SW64: chunk ^= (row >> 1) & 3
        uint32_t cc_ = 0;     // chunk counter (stage / row-slot rings)
        uint32_t iv = 0;      // flush-interval counter (accumulator barriers)
        const int wd0 = gw.wdim[0], wd1 = (NW > 1) ? gw.wdim[1] : -1;   // distance-weighted operand? (uniform)
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const double s2 = 0.5 * 1.4426950408889634 * cs->inv_l2[q];
            const double sscale = sqrt(s2);
            const int gcol = isA ? sg.I * 128 + col : sg.j0 + col;
            const bool active = isA || col < sg.nw;
            float zh[XD], zl[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                const double z = (gcol < M) ? pa.Zp[((size_t)q * Mp + gcol) * XD + i] : 0.0;
                split_scaled(z, sscale, zh[i], zl[i]);
            }
            const float lv = (float)log2(cs->var[q]);
            const float bias = (gcol < M) ? (isA ? lv : lv + (float)info->kexp[q]) : -1.0e30f;
            float2 nzh[XD], nzl[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) { nzh[i] = dup2(-zh[i]); nzl[i] = dup2(-zl[i]); }
            const float2 nb2 = dup2(-bias);   // K = ex2(-(d.d - bias))
            double g64[HM_GRAM_MAXV];
#pragma unroll
            for (int v = 0; v < HM_GRAM_MAXV; ++v) g64[v] = 0.0;
            // inverse operand scales of the accumulators (drain)
            float inv_k[2];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                inv_k[k] = (k < NW) ? 1.f / (weight_scale(info, tk, cs, q, gw.wbase[k], gw.wdim[k]) * pow2i(info->kexp[q])) : 0.f;
            double* slot = slots + (size_t)sg.slot * HM_GRAM_SLOT_DOUBLES;
            bool first_flush = true;

            for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += flush_chunks) {
                const int c1 = min(sg.chunk_end, c0 + flush_chunks);
                for (int c = c0; c < c1; ++c, ++cc_) {
                    const int stage = cc_ % kGStages, rs = cc_ % kRowSlots;
                    mbar_wait_warp(&sb->rowfull[rs], (cc_ / kRowSlots) & 1u);
                    mbar_wait_warp(&sb->empty[stage], ((cc_ / kGStages) & 1u) ^ 1u);
                    const float* rb = rowbuf + (size_t)rs * kGC * kRowFloats;   // SoA: array a at rb + a * kGC
                    uint8_t* st = stage_base + (size_t)stage * kGStageBytes;
                    if (!isA) {
                        if (active) {
                            uint8_t* b_hi = st + col * 64;
                            uint8_t* b_lo = b_hi + kGBHalf;
#pragma unroll
                            for (int n8 = 0; n8 < kGC / 8; ++n8) {
                                float2 e[4] = {nb2, nb2, nb2, nb2};   // d.d - bias for rows n8*8 .. +7, as 4 pairs
#pragma unroll
                                for (int i = 0; i < XD; ++i) {
                                    const float4 h0 = *reinterpret_cast<const float4*>(rb + i * kGC + n8 * 8);
                                    const float4 h1 = *reinterpret_cast<const float4*>(rb + i * kGC + n8 * 8 + 4);
                                    const float4 l0 = *reinterpret_cast<const float4*>(rb + (XD + i) * kGC + n8 * 8);
                                    const float4 l1 = *reinterpret_cast<const float4*>(rb + (XD + i) * kGC + n8 * 8 + 4);
                                    const float2 xh2[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w)};
                                    const float2 xl2[4] = {make_float2(l0.x, l0.y), make_float2(l0.z, l0.w), make_float2(l1.x, l1.y), make_float2(l1.z, l1.w)};
#pragma unroll
                                    for (int p = 0; p < 4; ++p) {
                                        const float2 d = add2(add2(xh2[p], nzh[i]), add2(xl2[p], nzl[i]));
                                        e[p] = fma2(d, d, e[p]);
                                    }
                                }
                                uint32_t hi[4], lo[4];
#pragma unroll
                                for (int p = 0; p < 4; ++p) split2(ex2(-e[p].x), ex2(-e[p].y), hi[p], lo[p]);
                                const int off = (n8 ^ swz) << 4;
                                *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                    } else {
                        uint8_t* a_hi = st + 2 * kGBHalf + col * 64;
                        float2 g2[NV > 0 ? NV : 1];
#pragma unroll
                        for (int v = 0; v < NV; ++v) g2[v] = dup2(0.f);
#pragma unroll
                        for (int n8 = 0; n8 < kGC / 8; ++n8) {
                            float2 e[4] = {nb2, nb2, nb2, nb2};
                            float2 dsel0[4], dsel1[4];   // signed distance of the weighted dim (distance-weighted operands)
                            float2 dd[XD][4];
#pragma unroll
                            for (int i = 0; i < XD; ++i) {
                                const float4 h0 = *reinterpret_cast<const float4*>(rb + i * kGC + n8 * 8);
                                const float4 h1 = *reinterpret_cast<const float4*>(rb + i * kGC + n8 * 8 + 4);
                                const float4 l0 = *reinterpret_cast<const float4*>(rb + (XD + i) * kGC + n8 * 8);
                                const float4 l1 = *reinterpret_cast<const float4*>(rb + (XD + i) * kGC + n8 * 8 + 4);
                                const float2 xh2[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w)};
                                const float2 xl2[4] = {make_float2(l0.x, l0.y), make_float2(l0.z, l0.w), make_float2(l1.x, l1.y), make_float2(l1.z, l1.w)};
#pragma unroll
                                for (int p = 0; p < 4; ++p) {
                                    dd[i][p] = add2(add2(xh2[p], nzh[i]), add2(xl2[p], nzl[i]));
                                    e[p] = fma2(dd[i][p], dd[i][p], e[p]);
                                }
                            }
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                dsel0[p] = dup2(1.f); dsel1[p] = dup2(1.f);
#pragma unroll
                                for (int i = 0; i < XD; ++i) {
                                    if (wd0 == i) dsel0[p] = dd[i][p];
                                    if (wd1 == i) dsel1[p] = dd[i][p];
                                }
                            }
                            const float4 wa0 = *reinterpret_cast<const float4*>(rb + (2 * XD) * kGC + n8 * 8);
                            const float4 wa1 = *reinterpret_cast<const float4*>(rb + (2 * XD) * kGC + n8 * 8 + 4);
                            const float2 w0[4] = {make_float2(wa0.x, wa0.y), make_float2(wa0.z, wa0.w), make_float2(wa1.x, wa1.y), make_float2(wa1.z, wa1.w)};
                            float2 kv[4], a0[4], a1[4];
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                kv[p] = make_float2(ex2(-e[p].x), ex2(-e[p].y));
                                a0[p] = mul2(wd0 >= 0 ? mul2(kv[p], dsel0[p]) : kv[p], w0[p]);
                            }
                            uint32_t hi[4], lo[4];
                            const int off = (n8 ^ swz) << 4;
#pragma unroll
                            for (int p = 0; p < 4; ++p) split2(a0[p].x, a0[p].y, hi[p], lo[p]);
                            *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(a_hi + kGAHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            if (NW > 1) {
                                const float4 wb0 = *reinterpret_cast<const float4*>(rb + (2 * XD + 1) * kGC + n8 * 8);
                                const float4 wb1 = *reinterpret_cast<const float4*>(rb + (2 * XD + 1) * kGC + n8 * 8 + 4);
                                const float2 w1[4] = {make_float2(wb0.x, wb0.y), make_float2(wb0.z, wb0.w), make_float2(wb1.x, wb1.y), make_float2(wb1.z, wb1.w)};
#pragma unroll
                                for (int p = 0; p < 4; ++p) {
                                    a1[p] = mul2(wd1 >= 0 ? mul2(kv[p], dsel1[p]) : kv[p], w1[p]);
                                    split2(a1[p].x, a1[p].y, hi[p], lo[p]);
                                }
                                *reinterpret_cast<uint4*>(a_hi + 2 * kGAHalf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(a_hi + 3 * kGAHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                            if (NV > 0) {
                                if (sg.has_g) {
#pragma unroll
                                    for (int v = 0; v < NV; ++v) {
                                        const float4 va = *reinterpret_cast<const float4*>(rb + (2 * XD + 2 + v) * kGC + n8 * 8);
                                        const float4 vb = *reinterpret_cast<const float4*>(rb + (2 * XD + 2 + v) * kGC + n8 * 8 + 4);
                                        const float2 vv[4] = {make_float2(va.x, va.y), make_float2(va.z, va.w), make_float2(vb.x, vb.y), make_float2(vb.z, vb.w)};
#pragma unroll
                                        for (int p = 0; p < 4; ++p)   // v = 0: g^mu (plain); v = 1 + i: distance-weighted in dim i
                                            g2[v] = fma2(v == 0 ? kv[p] : mul2(kv[p], dd[v > 0 ? v - 1 : 0][p]), vv[p], g2[v]);
                                    }
                                }
                            }
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
                // ---- flush: TMEM fp32 window -> fp64 partial tile of this segment
                mbar_wait_warp(&sb->accfull, iv & 1u);
                fence_after();
                {
                    const int lq = warp & 3, wq = warp >> 2;
                    const int ncc = sg.nw / 32, total = NW * ncc;
                    const int i = lq * 32 + lane;
                    for (int idx = wq; idx < total; idx += 3) {
                        const int k = idx / ncc, cc = idx % ncc;
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + k * 256 + cc * 32, v);
                        tmem_ld_wait();
                        double2* dst = reinterpret_cast<double2*>(slot + ((size_t)(k * 128 + i) * 256 + cc * 32));
                        const double sc = (double)inv_k[k];
                        if (first_flush) {
#pragma unroll
                            for (int p = 0; p < 16; ++p)
                                dst[p] = make_double2((double)__uint_as_float(v[2 * p]) * sc, (double)__uint_as_float(v[2 * p + 1]) * sc);
                        } else {
#pragma unroll
                            for (int p = 0; p < 16; ++p) {
                                double2 o = dst[p];
                                o.x += (double)__uint_as_float(v[2 * p]) * sc;
                                o.y += (double)__uint_as_float(v[2 * p + 1]) * sc;
                                dst[p] = o;
                            }
                        }
                    }
                }
                first_flush = false;
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sb->accempty);
                ++iv;
            }
            if (isA && sg.has_g) {
                double* gdst = slot + (size_t)2 * 128 * 256;
#pragma unroll
                for (int v = 0; v < NV; ++v) gdst[v * 128 + col] = g64[v];
            }
        }
    } else if (warp == 12) {
        // ======================================================= MMA issuer (one thread)
        if (lane == 0) {
            uint32_t cc_ = 0, iv = 0;
            for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
                const HmGramSeg sg = segs[sgi];
                const uint32_t idesc = idesc_f16(128, sg.nw);
                for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += flush_chunks) {
                    const int c1 = min(sg.chunk_end, c0 + flush_chunks);
                    mbar_wait(&sb->accempty, (iv & 1u) ^ 1u);
                    fence_after();
                    for (int c = c0; c < c1; ++c, ++cc_) {
                        const int stage = cc_ % kGStages;
                        mbar_wait(&sb->full[stage], (cc_ / kGStages) & 1u);
                        fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kGStageBytes);
                        const uint64_t b_hi = desc_sw64(sa), b_lo = desc_sw64(sa + kGBHalf);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if (k < NW) {
                                const uint64_t a_hi = desc_sw64(sa + 2 * kGBHalf + k * 2 * kGAHalf);
                                const uint64_t a_lo = desc_sw64(sa + 2 * kGBHalf + k * 2 * kGAHalf + kGAHalf);
                                const uint32_t d_tmem = tmem_base + k * 256;
#pragma unroll
                                for (int ks = 0; ks < kGC / 16; ++ks) {
                                    const uint64_t adv = (uint64_t)(ks * 2);
                                    mma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, (c > c0 || ks > 0) ? 1u : 0u);
                                    if (npass >= 2) mma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                                    if (npass >= 3) mma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                                }
                            }
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
        const int rw = warp - 13;
        constexpr int kGrp = 4;
        int nch[HM_MAXT];
        for (int t = 0; t < HM_MAXT; ++t) nch[t] = (t < tk.T) ? (int)((tk.count[t] + kGC - 1) / kGC) : 0;
        uint32_t cc_ = 0;
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const double sscale = sqrt(0.5 * 1.4426950408889634 * cs->inv_l2[q]);
            float wsc[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) wsc[k] = (k < NW) ? weight_scale(info, tk, cs, q, gw.wbase[k], gw.wdim[k]) : 0.f;
            // (task, chunk-in-task) of the segment's first chunk
            int t = 0, ct = sg.chunk_begin;
            while (t < tk.T && ct >= nch[t]) { ct -= nch[t]; ++t; }
            for (int c = sg.chunk_begin; c < sg.chunk_end; c += kGrp, cc_ += kGrp) {
                const int ng = min(kGrp, sg.chunk_end - c);
                const bool mine = (int)((cc_ / kGrp) & 1u) == rw;
                double xv[kGrp][XD];
                float wv[kGrp][2], vv[kGrp][HM_GRAM_MAXV];
#pragma unroll
                for (int j = 0; j < kGrp; ++j) {
                    if (j < ng) {
                        const int64_t row = (int64_t)ct * kGC + lane;
                        const bool valid = t < tk.T && row < tk.count[t];
                        if (mine) {
#pragma unroll
                            for (int i = 0; i < XD; ++i) xv[j][i] = valid ? tk.X[t][(tk.begin[t] + row) * XD + i] : 0.0;
                            const float* mw = reinterpret_cast<const float*>(tk.MW[t]) + row * 4 * Q;
#pragma unroll
                            for (int k = 0; k < 2; ++k) wv[j][k] = (valid && k < NW) ? mw[gw.wbase[k] * Q + q] : 0.f;
#pragma unroll
                            for (int v = 0; v < HM_GRAM_MAXV; ++v) vv[j][v] = (valid && v < NV) ? mw[gw.vbase[v] * Q + q] : 0.f;
                        }
                        if (++ct >= nch[t]) { ct = 0; ++t; while (t < tk.T && nch[t] == 0) ++t; }
                    }
                }
                if (!mine) continue;
#pragma unroll
                for (int j = 0; j < kGrp; ++j) {
                    if (j < ng) {
                        const uint32_t cj = cc_ + j;
                        const int rs = cj % kRowSlots;
                        float xh_[XD], xl_[XD];
#pragma unroll
                        for (int i = 0; i < XD; ++i) split_scaled(xv[j][i], sscale, xh_[i], xl_[i]);
                        mbar_wait_warp(&sb->rowempty[rs], ((cj / kRowSlots) & 1u) ^ 1u);
                        float* dst = rowbuf + (size_t)rs * kGC * kRowFloats + lane;   // SoA: array a, row = lane
#pragma unroll
                        for (int i = 0; i < XD; ++i) { dst[i * kGC] = xh_[i]; dst[(XD + i) * kGC] = xl_[i]; }
                        dst[(2 * XD) * kGC] = wv[j][0] * wsc[0];
                        if (NW > 1) dst[(2 * XD + 1) * kGC] = wv[j][1] * wsc[1];
#pragma unroll
                        for (int v = 0; v < NV; ++v) dst[(2 * XD + 2 + v) * kGC] = vv[j][v];
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sb->rowfull[rs]);
                    }
                }
            }
            cc_ -= 0;   // cc_ advanced by whole groups; realign to the true chunk count of this segment
            {
                const int n = sg.chunk_end - sg.chunk_begin;
                const int adv = ((n + kGrp - 1) / kGrp) * kGrp;
                cc_ = cc_ - adv + n;
            }
        }
    }
    // ---- teardown
    fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 512u);
}

// Sum the partial tiles of every (q, tile job) in slot order; write H^k (lower from the tile, mirrored) and g^v.
__global__ void tc_gram_reduce_kernel(const double* __restrict__ slots, const HmGramJob* __restrict__ jobs,
                                      const int2* __restrict__ jobslots, int njobs, HmGramWeights gw, double* H0, double* H1,
                                      double* g0, int64_t gstride, int M, int Mp) {
    const int job = blockIdx.x, q = blockIdx.y, k = blockIdx.z;
    const HmGramJob jb = jobs[job];
    const int2 sr = jobslots[q * njobs + job];
    double* H = (k == 0) ? H0 : H1;
    for (int e = threadIdx.x; e < 128 * jb.nw; e += blockDim.x) {
        const int i = e / jb.nw, j = e % jb.nw;
        const int gr = jb.I * 128 + i, gc = jb.j0 + j;
        if (gc > gr || gr >= M) continue;
        double s = 0.0;
        for (int sl = sr.x; sl < sr.y; ++sl) s += slots[(size_t)sl * HM_GRAM_SLOT_DOUBLES + ((size_t)(k * 128 + i) * 256 + j)];
        H[((size_t)q * Mp + gr) * Mp + gc] = s;
        if (gw.wdim[k] < 0) H[((size_t)q * Mp + gc) * Mp + gr] = s;   // plain Grams are symmetric; D^i keeps its lower triangle
    }
    if (k == 0 && jb.j0 == 0) {
        for (int e = threadIdx.x; e < gw.nV * 128; e += blockDim.x) {
            const int v = e / 128, i = e % 128;
            if (jb.I * 128 + i >= M) continue;
            double s = 0.0;
            for (int sl = sr.x; sl < sr.y; ++sl) s += slots[(size_t)sl * HM_GRAM_SLOT_DOUBLES + (size_t)2 * 128 * 256 + v * 128 + i];
            g0[(size_t)v * gstride + (size_t)q * Mp + jb.I * 128 + i] = s;
        }
    }
}

size_t gram_smem_bytes() {
    return (size_t)kGStages * kGStageBytes + sizeof(float) * kRowSlots * kGC * kRowFloats + sizeof(GramBars) + 64 + 1024;
}

template <int XD, int NW, int NV>
int launch_gram3(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                 const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int flush_chunks, int npass) {
    const size_t smem = gram_smem_bytes();
    HM_CUDA(cudaFuncSetAttribute(tc_gram_kernel<XD, NW, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gram_kernel<XD, NW, NV><<<nctas, kGThreads, smem, s>>>(tk, a, info, segs, seg_off, gw, slots, flush_chunks, npass);
    HM_CUDA(cudaGetLastError());
    return 0;
}

// (NW, NV) combinations the engine issues: (1,1) VE step; (2,1+XD) first launch of a full step; (1,0)/(2,0) the rest
template <int XD>
int launch_gram(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int flush_chunks, int npass) {
    if (gw.nW == 1 && gw.nV == 1) return launch_gram3<XD, 1, 1>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
    if (gw.nW == 1 && gw.nV == 0) return launch_gram3<XD, 1, 0>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
    if (gw.nW == 2 && gw.nV == 0) return launch_gram3<XD, 2, 0>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
    if (gw.nW == 2 && gw.nV == 1 + XD) return launch_gram3<XD, 2, 1 + XD>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
    hm_set_error("gram launch: unsupported (nW=%d, nV=%d) for Xdim=%d", gw.nW, gw.nV, XD);
    return HMOGP_ERR_ARG;
}

}  // namespace

int hm_tc_gram(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
               const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int flush_chunks, int npass) {
    switch (a.Xdim) {
        case 1: return launch_gram<1>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
        case 2: return launch_gram<2>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
        case 3: return launch_gram<3>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
        case 4: return launch_gram<4>(s, tk, a, info, segs, seg_off, gw, slots, nctas, flush_chunks, npass);
    }
    hm_set_error("Xdim=%d unsupported", a.Xdim);
    return HMOGP_ERR_ARG;
}

int hm_tc_gram_reduce(cudaStream_t s, const double* slots, const HmGramJob* jobs, const int2* jobslots, int njobs, int Q,
                      const HmGramWeights& gw, double* H0, double* H1, double* g0, int64_t gstride, int M, int Mp) {
    dim3 grid((unsigned)njobs, (unsigned)Q, (unsigned)gw.nW);
    tc_gram_reduce_kernel<<<grid, 256, 0, s>>>(slots, jobs, jobslots, njobs, gw, H0, H1, g0, gstride, M, Mp);
    HM_CUDA(cudaGetLastError());
    return 0;
}
