// Weighted Gram of the backward pass on CTA pairs (tcgen05 cta_group::2).
//
//     H_q[i, j] = sum_t sum_n omega[n] K[n,i] K[n,j]          (M x M; lower triangle in blocks of 256 x 256)
//     g_q[i]    = sum_t sum_n mu[n] K[n,i]                    (VE steps; a full step takes it from tc_bwd.cu)
// from which dVE/dS_q = K_uu^-1 H K_uu^-1 (reference: A^T diag(dv) A per output function,
// /root/reference/hetmogp/svmogp_inf.py:145-148, summed over d with W_dq^2 folded into omega) and
// dVE/dm_q (svmogp_inf.py:144).
//
// Operand economy.  The weight is split symmetrically: with V[n, m] = 2^kexp K[n,m] sqrt|omega_n| 2^se (K 2^kexp < 2^12,
// sqrt|omega| 2^se <= 4),
//     H = 2^-(2 kexp + 2 se) (sgn(omega) V)^T V,
// so both operands come from ONE generated and split value (V = Vh + Vl in fp16).  When all weights of a 64-row chunk have
// the same sign (the usual case: omega <= 0 for log-concave likelihoods) the MMA applies the sign through the negate-A bit
// of its instruction descriptor, and on a diagonal block the A descriptor simply points at the B tile; in a mixed-sign
// chunk the A operand is the B operand with the sign bit of the row flipped (an XOR on the packed fp16 pairs).  A CTA pair
// owns a 256 x 256 output block: CTA r generates its 128 rows of A and its 128-column half of B per 64-row chunk
// (cta_group::2 takes the other half from the peer's shared memory), so a 128 x 256 MMA tile costs 256 generated columns
// off the diagonal and 128 on it, against 384 in the one-CTA kernel.  Off the diagonal a stage takes three products
// (Vh^T Vh + Vh^T Vl + Vl^T Vh); on it two: Y = Vh^T Vh + Vh^T (2 Vl), whose symmetric part is the same sum -- the reduce
// kernel symmetrises diagonal blocks (HM_G2_SYMDIAG).
//
// Accumulation.  tcgen05 adds into its fp32 accumulator with truncation, so a long accumulation drifts by a bias that grows
// with the window length and is then amplified by K_uu^-1 . K_uu^-1 in the finish chain ("Gram accuracy").  The
// 512 TMEM columns therefore hold TWO 128 x 256 accumulators; windows of `f1` chunks alternate between them.  While the
// MMA issuer fills one, four fold warps (one per TMEM lane quadrant) drain the other: tcgen05.ld -> fp64 add into this
// CTA's partial tile in L2 (slot_index: every warp-wide access is 32 consecutive 16-byte pieces) -> write
// -HM_G2_CARRY x value back as the start of that buffer's next window.  The carry keeps the accumulator centred around
// zero, where truncation errors of growing and shrinking magnitudes cancel to first order; the bookkeeping is exact:
// sum_w S_w = sum_w (1 + carry_w) v_w, with carry_w = HM_G2_CARRY when the buffer has another window in the segment, else 0,
// which is what the fold adds.  Generators never stop for a fold; a fold has one window's time to finish.  What a fold
// costs is its fp64 accumulation traffic through the L2 (512 KB of read-modify-write per CTA and window; the chip sustains
// ~3.5 TB/s of it): with the adds taken out the kernel runs at the generator-bound 11 ms down to 256-row windows, and
// cp.reduce.async.bulk .add.f64 from a staging ring (adds at the L2, nothing read back) costs the same.
//
// MMA: D[i (2 x 128 TMEM lanes), j (256 columns)] += A[i][n] . B[j][n]^T over n = 64 data rows per stage (SWIZZLE_128B).
// Warp roles (768 threads per CTA at 80 registers, one CTA per SM, pairs persistent over a host-built plan of segments):
//   warps 0-15  generators: (column group of 32) x (16 of the 64 rows)
//   warp 16     leader: MMA issuer (one thread); peer: relays "my stage is written" to the leader's barrier
//   warps 17-19 row loaders (x, sqrt|omega|, sign words, mu -> smem ring), alternating groups of 2 chunks
//   warps 20-23 fold warps (above)
#include "tc_common.cuh"

using namespace tc;

namespace {

#ifndef HM_G2_FORCE_MIXED
#define HM_G2_FORCE_MIXED 0   // test hook: treat every chunk as mixed-sign
#endif
#ifndef HM_G2_LOADERS
#define HM_G2_LOADERS 3        // row-loader warps (16 generator + 1 MMA + 3 loader warps = 20: the register file is allocated per 4 warps)
#endif
#ifndef HM_G2_ALIAS
#define HM_G2_ALIAS 1          // diagonal blocks, uniform sign: A descriptor = B tile (no A stores)
#endif
#ifndef HM_G2_CORR
#define HM_G2_CORR 1      // diagonal lo x lo correction
#endif
#ifndef HM_G2_SYMDIAG
#define HM_G2_SYMDIAG 1   // diagonal blocks: Y = Vh^T Vh + Vh^T (2 Vl), symmetrised by the reduce kernel (2 products instead of 3)
#endif
#ifndef HM_G2_CENTRE
#define HM_G2_CENTRE 1    // carry-centred windows (header comment)
#endif
constexpr int kC = HM_GRAM2_CHUNK;                 // data rows per chunk = MMA K extent per stage (4 x K16)
constexpr int kStages = 3;
constexpr int kHalf = 128 * 128;                   // 16 KB: 128 operand rows (inducing points) x 64 data rows fp16, SW128
constexpr int kStageBytes = 4 * kHalf;             // A hi | A lo | B hi | B lo
constexpr int kRowSlots = 8;                       // row-data ring (decoupled from the operand stages: the loader runs ahead)
constexpr int kRowArrays = 12;                     // per slot, [kC] floats each (SoA): xh[XD] | xl[XD] | sw | sign words | mu
constexpr int kGenWarps = 16;
constexpr int kMmaWarp = kGenWarps;
#ifndef HM_G2_PHASES
#define HM_G2_PHASES 1
#endif
constexpr int kPhases = HM_G2_PHASES;               // generator warp groups on alternate chunks (2 was measured slower: the ring is 3 deep)
constexpr int kPhaseWarps = kGenWarps / kPhases;    // 4 column groups x (4 / kPhases) row parts
constexpr int kGroups = 2 * kPhases;                // groups of 8 rows per thread and chunk
constexpr int kLoaders = HM_G2_LOADERS;
constexpr int kFoldWarps = 4;                       // one per TMEM lane quadrant (a warp reaches lanes 32 (warp % 4) .. + 31)
constexpr int kFoldWarp0 = kGenWarps + 1 + kLoaders;
constexpr int kThreads = (kGenWarps + 1 + kLoaders + kFoldWarps) * 32;
static_assert(kFoldWarp0 % 4 == 0, "fold warp w serves TMEM lane quadrant w % 4");
constexpr uint32_t kAccCols = 256;                 // TMEM columns per accumulator buffer (two buffers: windows alternate)
#ifndef HM_G2_LGRP
#define HM_G2_LGRP 2        // chunks per row-loader group
#endif
#ifndef HM_G2_TRACE
#define HM_G2_TRACE 0       // 1: CTA 0 records clock64() at the hand-offs of its first chunks (tools/g2_trace.py)
#endif
#ifndef HM_G2_KO
#define HM_G2_KO 0          // timing experiments (wrong results): 1 generators do nothing but the handshakes, 2 no operand stores,
                            // 3 no ex2, 4 no hi/lo split
#endif
#ifndef HM_G2_CARRY
#define HM_G2_CARRY 0.75f                          // a window starts at -HM_G2_CARRY times the buffer's previous final value
#endif
static_assert(kC == 64, "SWIZZLE_128B operand rows hold 64 fp16 values");

struct Bars {
    uint64_t full[kStages], empty[kStages], rowfull[kRowSlots], rowempty[kRowSlots], accfull[2], accempty[2];
    uint32_t sflag[kStages];   // sign class of the chunk in the stage (for the MMA issuer)
    uint32_t tmem_base;
};

// fp64 partial tile of a (segment, CTA): element (row i < 128, column j < 256) of the block.  Laid out so that the 32 lanes
// of a fold warp (consecutive rows, the same column pair) touch 512 consecutive bytes: [j / 16][(j % 16) / 2][i][j % 2].
__device__ __host__ __forceinline__ size_t slot_index(int i, int j) {
    return ((size_t)((j >> 4) * 8 + ((j & 15) >> 1)) * 128 + i) * 2 + (j & 1);
}

// exponent se with max sqrt|w| 2^se <= 4 (operands then stay below 2^14: K 2^kexp < 2^12)
__device__ __forceinline__ int weight_exp(const HmTcInfo* info, int q) {
    const float amax = __uint_as_float(info->wmax[0][q]);
    if (!(amax > 0.f) || !isfinite(amax)) return 0;
    int e = 0;
    frexpf(amax, &e);            // amax < 2^e
    return 2 - ((e + 1) >> 1);   // ceil(e / 2)
}

template <int XD> struct RowX { float2 h[XD][4], l[XD][4]; };   // 8 rows of scaled split inputs, as 4 row pairs

__device__ __forceinline__ void load8(float2 (&o)[4], const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = make_float2(a.x, a.y); o[1] = make_float2(a.z, a.w); o[2] = make_float2(b.x, b.y); o[3] = make_float2(b.z, b.w);
}
template <int XD>
__device__ __forceinline__ void load_rowx(RowX<XD>& r, const float* rb, int n8) {
#pragma unroll
    for (int i = 0; i < XD; ++i) {
        load8(r.h[i], rb + i * kC + n8 * 8);
        load8(r.l[i], rb + (XD + i) * kC + n8 * 8);
    }
}
// unit-variance kernel values of 8 rows against one inducing point (-z split in zh/zl; nb = 0, or 1e30 for padding)
template <int XD>
__device__ __forceinline__ void kgen(const RowX<XD>& r, const float (&zh)[XD], const float (&zl)[XD], float nb, float2 (&kv)[4]) {
    float2 e[4] = {dup2(nb), dup2(nb), dup2(nb), dup2(nb)};
#pragma unroll
    for (int i = 0; i < XD; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float2 d = add2(add2(r.h[i][p], dup2(zh[i])), add2(r.l[i][p], dup2(zl[i])));
            e[p] = fma2(d, d, e[p]);
        }
#pragma unroll
    for (int p = 0; p < 4; ++p) kv[p] = (HM_G2_KO == 3) ? e[p] : make_float2(ex2(-e[p].x), ex2(-e[p].y));
}

// split2 of tc_common.cuh that also returns the fp32 residual v - hi (the value the fp16 lo pair rounds)
__device__ __forceinline__ void split2r(float2 v, uint32_t& hi, uint32_t& lo, float2& r) {
    if (HM_G2_KO == 4) { hi = __float_as_uint(v.x); lo = __float_as_uint(v.y); r = v; return; }
    const __half2 h = __floats2half2_rn(v.x, v.y);
    r = __fadd2_rn(v, make_float2(-__low2float(h), -__high2float(h)));
    const __half2 l = __floats2half2_rn(r.x, r.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void G2_STORE(uint8_t* p, uint4 v) {
    if (HM_G2_KO != 2 || v.x == 0x12345678u) *reinterpret_cast<uint4*>(p) = v;
}
#if HM_G2_TRACE
__device__ long long g2_trace[8][2048];
#define G2_TR(k, c) do { if (blockIdx.x == 0 && (c) < 2048u) g2_trace[k][c] = clock64(); } while (0)
#else
#define G2_TR(k, c) do { } while (0)
#endif

template <int XD, int NV>
__global__ void __launch_bounds__(kThreads, 1)   // 24 warps: 80 registers per thread (the generators, rid of the folds, fit)
tc_gram2_kernel(HmTasks tk, HmProjArgs pa, const HmTcInfo* __restrict__ info, const HmGramSeg* __restrict__ segs,
                const int* __restrict__ seg_off, double* __restrict__ slots, int f1, int f2, int npass) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    float* rowbuf = reinterpret_cast<float*>(smem + kStages * kStageBytes);   // [kRowSlots][kRowArrays][kC]
    Bars* sb = reinterpret_cast<Bars*>(rowbuf + kRowSlots * kC * kRowArrays);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();      // 0 = leader of the pair
    const int pair = blockIdx.x >> 1;
    const int Mp = pa.Mp, M = pa.M, Q = pa.Q;
    const HmConsts* __restrict__ cs = pa.consts;
    const int seg_begin = seg_off[pair], seg_end = seg_off[pair + 1];

    if (threadIdx.x == 0) {
        // full: the phase group's warps wrote the operand tiles (+ on the leader: the peer's relay); empty: the MMAs that
        // read the stage have completed; rowfull / rowempty: row-data ring between the loader and the generators
        for (int s = 0; s < kStages; ++s) { mbar_init(&sb->full[s], kPhaseWarps + (rank == 0 ? 1 : 0)); mbar_init(&sb->empty[s], 1); }
        for (int s = 0; s < kRowSlots; ++s) { mbar_init(&sb->rowfull[s], 1); mbar_init(&sb->rowempty[s], kPhaseWarps); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sb->accfull[b], 1);
            mbar_init(&sb->accempty[b], 2 * kFoldWarps);  // used on the leader: both CTAs' fold warps
        }
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc2(&sb->tmem_base, 512u);
    fence_before();
    cluster_sync();
    fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    if (warp < kGenWarps) {
        // ======================================================= generators
        const int cgp = warp & 3, rh = (warp % kPhaseWarps) >> 2, rp = warp >> 2;   // column group, row part of the chunk, partial index
        const uint32_t phase = (uint32_t)warp / kPhaseWarps;
        const int col = cgp * 32 + lane;                  // operand row (inducing point within this CTA's 128) this thread writes
        const int swz = col & 7;                          // SW128: 16-byte chunk ^= row & 7
        uint32_t cc_ = 0;     // chunk counter (stage / row-slot rings)
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const double sscale = sqrt(0.5 * 1.4426950408889634 * cs->inv_l2[q]);
            const int acol = sg.I * 256 + (int)rank * 128 + col;       // A: this CTA's 128 output rows
            const int bcol = sg.j0 + (int)rank * 128 + col;            // B: this CTA's half of the 256 output columns
            const bool diag = sg.j0 == sg.I * 256;                     // then B is A without the sign
            float azh[XD], azl[XD], bzh[XD], bzl[XD];
#pragma unroll
            for (int i = 0; i < XD; ++i) {
                const double za = (acol < M) ? pa.Zp[((size_t)q * Mp + acol) * XD + i] : 0.0;
                const double zb = (bcol < M) ? pa.Zp[((size_t)q * Mp + bcol) * XD + i] : 0.0;
                float h, l;
                split_scaled(za, sscale, h, l); azh[i] = -h; azl[i] = -l;
                split_scaled(zb, sscale, h, l); bzh[i] = -h; bzl[i] = -l;
            }
            // the generated value is the unit-variance kernel ex2(-d.d): its argument is small where K is large, so the fp32
            // rounding of the argument stays below 1e-7 there; sigma^2 2^kexp rides on the row multiplier (loader)
            const float anb = (acol < M) ? 0.f : 1.0e30f, bnb = (bcol < M) ? 0.f : 1.0e30f;   // padded columns give 0
            double g64 = 0.0;
            float2 g2 = dup2(0.f);      // g partial of the current level-1 window (fp32), folded into g64 per window
            float2 cacc = dup2(0.f);    // sum of sgn . lo^2 of the diagonal entry over the segment (a 1e-7 correction: fp32 is plenty)
            const int se = weight_exp(info, q);
            const double ksc = (double)((float)cs->var[q] * pow2i(info->kexp[q]));              // as the loader rounds it
            const double inv_sc = (double)pow2i(-2 * se) * (cs->var[q] / ksc) * (cs->var[q] / ksc);   // D = (ksc / var)^2 2^(2 se) H
            const double inv_k = cs->var[q];
            double* slot = slots + (size_t)(2 * sg.slot + (int)rank) * HM_GRAM_SLOT_DOUBLES;

            for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += f1) {
                const int c1 = min(sg.chunk_end, c0 + f1);
                for (int c = c0; c < c1; ++c, ++cc_) {
                    if (kPhases > 1 && (cc_ % kPhases) != phase) continue;
                    const int stage = cc_ % kStages, rs = cc_ % kRowSlots;
                    mbar_wait_warp(&sb->rowfull[rs], (cc_ / kRowSlots) & 1u);
                    if (threadIdx.x == 0) G2_TR(0, cc_);
                    mbar_wait_warp(&sb->empty[stage], ((cc_ / kStages) & 1u) ^ 1u);
                    if (threadIdx.x == 0) G2_TR(1, cc_);
                    const float* rb = rowbuf + (size_t)rs * kC * kRowArrays;   // SoA: array a at rb + a * kC
                    uint8_t* a_hi = stage_base + (size_t)stage * kStageBytes + col * 128;
                    uint8_t* b_hi = a_hi + 2 * kHalf;
                    const uint32_t sflag = reinterpret_cast<const uint32_t*>(rb + (2 * XD + 1) * kC)[32];   // 0: all w >= 0, 1: all <= 0, 2: mixed
                    // Uniform sign (the usual case: omega <= 0 for log-concave likelihoods): the MMA negates A through its
                    // instruction descriptor and no per-element sign work is needed.
                    const bool mixed = sflag == 2u;
#pragma unroll 2
                    for (int g = 0; g < (HM_G2_KO == 1 ? 0 : kGroups); ++g) {
                        const int n8 = rh * kGroups + g;
                        const int off = (n8 ^ swz) << 4;
                        RowX<XD> r;
                        load_rowx<XD>(r, rb, n8);
                        float2 sw[4];
                        load8(sw, rb + (2 * XD) * kC + n8 * 8);
                        float2 kv[4];
                        kgen<XD>(r, azh, azl, anb, kv);
                        if (NV > 0) {
                            if (sg.has_g) {
                                float2 mu[4];
                                load8(mu, rb + (2 * XD + 2) * kC + n8 * 8);
#pragma unroll
                                for (int p = 0; p < 4; ++p) g2 = fma2(kv[p], mu[p], g2);
                            }
                        }
                        uint32_t hi[4], lo[4];
                        float2 res[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) split2r(mul2(kv[p], sw[p]), hi[p], lo[p], res[p]);
                        if (diag) {
                            G2_STORE(b_hi + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                            if (HM_G2_SYMDIAG) {
                                // A diagonal block is symmetric: hh + hl + lh = sym(hh + 2 hl).  The lo operand is stored
                                // doubled (exact) and only the two products A_hi.B_hi, A_hi.(2 B_lo) are issued; the reduce
                                // kernel averages Y[i][j] and Y[j][i].
#pragma unroll
                                for (int p = 0; p < 4; ++p) lo[p] = pack_h2(2.f * res[p].x, 2.f * res[p].y);
                            }
                            G2_STORE(b_hi + kHalf + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
                            // Ah.Bh + Ah.Bl + Al.Bh leaves out Al.Bl.  Between different inducing points the residuals are
                            // uncorrelated; on the diagonal entry both operands are the same value, the term is sgn . lo^2
                            // every row, and K_uu^-1 . K_uu^-1 amplifies that 1e-7-relative diagonal bias to several 1e-3
                            // of dL/dS (measured).  It is accumulated here per column and added by the reduce kernel.
                            if (HM_G2_CORR) {
                                if (sflag == 0u) {
#pragma unroll
                                    for (int p = 0; p < 4; ++p) cacc = fma2(res[p], res[p], cacc);
                                } else if (sflag == 1u) {
#pragma unroll
                                    for (int p = 0; p < 4; ++p) cacc = fma2(make_float2(-res[p].x, -res[p].y), res[p], cacc);
                                }
                            }
                        }
                        if (mixed) {
                            const uint4 sgn = *reinterpret_cast<const uint4*>(rb + (2 * XD + 1) * kC + n8 * 4);
                            if (HM_G2_CORR && diag) {
                                const uint32_t sw4[4] = {sgn.x, sgn.y, sgn.z, sgn.w};
#pragma unroll
                                for (int p = 0; p < 4; ++p) {
                                    const float2 t = make_float2(__uint_as_float(__float_as_uint(res[p].x) ^ ((sw4[p] << 16) & 0x80000000u)),
                                                                 __uint_as_float(__float_as_uint(res[p].y) ^ (sw4[p] & 0x80000000u)));
                                    cacc = fma2(t, res[p], cacc);
                                }
                            }
                            G2_STORE(a_hi + off, make_uint4(hi[0] ^ sgn.x, hi[1] ^ sgn.y, hi[2] ^ sgn.z, hi[3] ^ sgn.w));
                            if (!(HM_G2_SYMDIAG && diag))
                                G2_STORE(a_hi + kHalf + off, make_uint4(lo[0] ^ sgn.x, lo[1] ^ sgn.y, lo[2] ^ sgn.z, lo[3] ^ sgn.w));
                        } else if (!diag || !HM_G2_ALIAS) {
                            G2_STORE(a_hi + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                            G2_STORE(a_hi + kHalf + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
                        }
                        if (!diag) {
                            kgen<XD>(r, bzh, bzl, bnb, kv);
#pragma unroll
                            for (int p = 0; p < 4; ++p) split2r(mul2(kv[p], sw[p]), hi[p], lo[p], res[p]);
                            G2_STORE(b_hi + off, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                            G2_STORE(b_hi + kHalf + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if ((warp % kPhaseWarps) == 0) sb->sflag[stage] = sflag;   // ordered before the MMA issuer's read by the barrier
                        mbar_arrive(&sb->full[stage]);
                        mbar_arrive(&sb->rowempty[rs]);
                        if (threadIdx.x == 0) G2_TR(2, cc_);
                    }
                }
                if (NV > 0 && sg.has_g) { g64 += (double)(g2.x + g2.y); g2 = dup2(0.f); }   // fp32 partial per window
            }
            if (NV > 0 && sg.has_g) slot[(size_t)128 * 256 + rp * 128 + col] = g64 * inv_k;   // one partial per row part
            if (diag) slot[(size_t)128 * 256 + 512 + rp * 128 + col] = (double)(cacc.x + cacc.y) * inv_sc;
        }
    } else if (warp >= kFoldWarp0) {
        // ======================================================= fold warps: finished accumulator windows -> fp64 partial tile
        // The 512 TMEM columns hold TWO fp32 accumulators of the 256 x 256 block; windows of f1 chunks alternate between
        // them, so the MMAs of window w + 1 run while window w is moved out here -- one tcgen05.ld pass over the buffer and a
        // coalesced read-modify-write of the segment's (L2-resident) fp64 tile -- and the generators never stop.  A fold
        // has a whole window of MMA time to finish, so four warps with a few loads in flight are enough.
        // tcgen05 accumulates with truncation: every MMA loses up to one ulp of the running sum TOWARDS ZERO, a bias that
        // grows with the window length and that K_uu^-1 . K_uu^-1 amplifies.  The sums of a window all have one
        // sign (V^T V with the sign of omega), so a buffer's next window does not start at 0 but at -c times the value v the
        // buffer ended this one with (c = HM_G2_CARRY): the accumulator then runs from about -0.43 S to +0.57 S and most of
        // the bias cancels.  The bookkeeping is exact and needs no storage: with v_w = I_w + S_w and I_w = -c v_{w-2},
        // sum_w S_w = sum_w (1 + c [window w + 2 of this segment exists]) v_w, accumulated in fp64.
        const int lq = warp & 3;
        const int i = lq * 32 + lane;
        uint32_t iv = 0;
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const int se = weight_exp(info, q);
            const double ksc = (double)((float)cs->var[q] * pow2i(info->kexp[q]));              // as the loader rounds it
            const double inv_sc = (double)pow2i(-2 * se) * (cs->var[q] / ksc) * (cs->var[q] / ksc);   // D = (ksc / var)^2 2^(2 se) H
            double* slot = slots + (size_t)(2 * sg.slot + (int)rank) * HM_GRAM_SLOT_DOUBLES;
            for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += f1, ++iv) {
                const uint32_t buf = iv & 1u;
                const bool fresh = c0 == sg.chunk_begin;                                   // first fold of the segment stores
                const bool carry = HM_G2_CENTRE && (c0 + 2 * f1 < sg.chunk_end);           // the buffer's next window is of this segment
                const double sc = inv_sc * (carry ? 1.0 + (double)HM_G2_CARRY : 1.0);
                mbar_wait_warp(&sb->accfull[buf], (iv >> 1) & 1u);
                fence_after();
                const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16) + buf * kAccCols;
#pragma unroll 1
                for (int c16 = 0; c16 < 16; ++c16) {
                    uint32_t v[16];
                    tmem_ld16(tl + c16 * 16, v);
                    double2* dst = reinterpret_cast<double2*>(slot + slot_index(i, c16 * 16));   // + p * 128 double2 per column pair
                    double2 o[8];
                    if (!fresh) {
#pragma unroll
                        for (int p = 0; p < 8; ++p) o[p] = __ldcg(dst + p * 128);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const double x = (double)__uint_as_float(v[2 * p]) * sc, y = (double)__uint_as_float(v[2 * p + 1]) * sc;
                        __stcg(dst + p * 128, fresh ? make_double2(x, y) : make_double2(o[p].x + x, o[p].y + y));
                    }
                    if (carry) {
#pragma unroll
                        for (int p = 0; p < 16; ++p) v[p] = __float_as_uint(-HM_G2_CARRY * __uint_as_float(v[p]));
                        tmem_st16(tl + c16 * 16, v);                   // start value of this buffer's next window
                    }
                }
                if (carry) tmem_st_wait();
                fence_before();
                __syncwarp();
                if (lane == 0) { if (rank == 0) mbar_arrive(&sb->accempty[buf]); else mbar_arrive_remote(&sb->accempty[buf], 0); }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0 && rank == 0) {
            // ======================================================= MMA issuer (one thread of the leader CTA)
            constexpr uint32_t idesc = idesc_f16(256, 256);
            uint32_t cc_ = 0, iv = 0;
            for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
                const HmGramSeg sg = segs[sgi];
                const bool diag = sg.j0 == sg.I * 256;
                for (int c0 = sg.chunk_begin; c0 < sg.chunk_end; c0 += f1) {
                    const int c1 = min(sg.chunk_end, c0 + f1);
                    // the fold of this buffer's previous window pre-set the accumulator if that window was of this segment
                    const uint32_t preset = (HM_G2_CENTRE && c0 - 2 * f1 >= sg.chunk_begin) ? 1u : 0u;
                    const uint32_t buf = iv & 1u;
                    const uint32_t d_tmem = tmem_base + buf * kAccCols;
                    mbar_wait_cluster(&sb->accempty[buf], ((iv >> 1) & 1u) ^ 1u);
                    fence_after();
                    for (int c = c0; c < c1; ++c, ++cc_) {
                        const int stage = cc_ % kStages;
                        G2_TR(3, cc_);
                        mbar_wait_cluster(&sb->full[stage], (cc_ / kStages) & 1u);
                        G2_TR(4, cc_);
                        fence_after();
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * kStageBytes);
                        // sign class of the chunk's weights: uniform -> A is the unsigned operand, negated by the instruction if the
                        // weights are negative
                        const uint32_t sflag = *reinterpret_cast<const volatile uint32_t*>(&sb->sflag[stage]);
                        const bool mixed = sflag == 2u;
                        const uint32_t idc = idesc | (sflag == 1u ? (1u << 13) : 0u);   // bit 13: negate A
                        const uint64_t b_hi = desc_sw128(sa + 2 * kHalf), b_lo = desc_sw128(sa + 3 * kHalf);
                        const bool alias = HM_G2_ALIAS && diag && !mixed;
                        const uint64_t a_hi = alias ? b_hi : desc_sw128(sa), a_lo = alias ? b_lo : desc_sw128(sa + kHalf);
#pragma unroll
                        for (int ks = 0; ks < kC / 16; ++ks) {
                            const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per K=16 step, in 16-byte units
                            mma2_f16(d_tmem, a_hi + adv, b_hi + adv, idc, (c > c0 || ks > 0) ? 1u : preset);
                            if (npass >= 2) mma2_f16(d_tmem, a_hi + adv, b_lo + adv, idc, 1u);
                            if (npass >= 3 && !(HM_G2_SYMDIAG && diag)) mma2_f16(d_tmem, a_lo + adv, b_hi + adv, idc, 1u);
                            if (npass >= 4) mma2_f16(d_tmem, a_lo + adv, b_lo + adv, idc, 1u);   // diagnostic
                        }
                        G2_TR(5, cc_);
                        commit2(&sb->empty[stage]);
                        G2_TR(6, cc_);
                    }
                    commit2(&sb->accfull[buf]);
                    ++iv;
                }
            }
        } else if (lane == 0) {
            // peer CTA: relay "my operand tiles of this stage are in shared memory" to the leader
            uint32_t cc_ = 0;
            for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
                const HmGramSeg sg = segs[sgi];
                for (int c = sg.chunk_begin; c < sg.chunk_end; ++c, ++cc_) {
                    const int stage = cc_ % kStages;
                    mbar_wait(&sb->full[stage], (cc_ / kStages) & 1u);
                    mbar_arrive_remote(&sb->full[stage], 0);
                }
            }
        }
    } else {
        // ======================================================= row loaders (alternating groups of 2 chunks)
        // Each iteration issues the global loads of 2 chunks (128 rows) before touching the ring: memory-level
        // parallelism instead of one exposed HBM/L2 latency per chunk.  Both CTAs of a pair stage the same rows.
        // (One loader warp caps the kernel at ~3400 cycles per chunk: fp64 input splits, sqrt, sign words.)
        constexpr int kGrp = HM_G2_LGRP;
        static_assert(kGrp * kLoaders <= kRowSlots, "a loader must stay within one lap of the row ring (mbarrier parity)");
        const int rw = warp - (kMmaWarp + 1);
        uint32_t grp = 0;
        int nch[HM_MAXT];
        for (int t = 0; t < HM_MAXT; ++t) nch[t] = (t < tk.T) ? (int)((tk.count[t] + kC - 1) / kC) : 0;
        uint32_t cc_ = 0;
        for (int sgi = seg_begin; sgi < seg_end; ++sgi) {
            const HmGramSeg sg = segs[sgi];
            const int q = sg.q;
            const double sscale = sqrt(0.5 * 1.4426950408889634 * cs->inv_l2[q]);
            const float wsc = pow2i(2 * weight_exp(info, q));
            const float ksc = (float)cs->var[q] * pow2i(info->kexp[q]);
            // (task, chunk-in-task) of the segment's first chunk
            int t = 0, ct = sg.chunk_begin;
            while (t < tk.T && ct >= nch[t]) { ct -= nch[t]; ++t; }
            for (int c = sg.chunk_begin; c < sg.chunk_end; c += kGrp, ++grp) {
                const int ng = min(kGrp, sg.chunk_end - c);
                const bool mine = (int)(grp % kLoaders) == rw;
                double xv[kGrp][2][XD];
                float wv[kGrp][2], vv[kGrp][2];
#pragma unroll
                for (int j = 0; j < kGrp; ++j) {
                    if (j < ng) {
                        if (mine) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int64_t row = (int64_t)ct * kC + h * 32 + lane;
                                const bool valid = t < tk.T && row < tk.count[t];
#pragma unroll
                                for (int i = 0; i < XD; ++i) xv[j][h][i] = valid ? tk.X[t][(tk.begin[t] + row) * XD + i] : 0.0;
                                const float* mw = reinterpret_cast<const float*>(tk.MW[t < tk.T ? t : 0]) + row;   // SoA: array k at k * cap
                                const size_t cap = (size_t)tk.cap[t < tk.T ? t : 0];
                                wv[j][h] = valid ? mw[(size_t)(Q + q) * cap] : 0.f;                               // omega
                                vv[j][h] = (NV > 0 && valid) ? mw[(size_t)q * cap] : 0.f;                         // mu
                            }
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
                            mbar_wait_warp(&sb->rowempty[rs], ((cj / kRowSlots) & 1u) ^ 1u);
                            float* dst = rowbuf + (size_t)rs * kC * kRowArrays;   // SoA: array a, row = h * 32 + lane
                            {
                            const bool anyneg = __any_sync(0xffffffffu, wv[j][0] < 0.f || wv[j][1] < 0.f);
                            const bool anypos = __any_sync(0xffffffffu, wv[j][0] > 0.f || wv[j][1] > 0.f);
                            if (lane == 0) reinterpret_cast<uint32_t*>(dst + (2 * XD + 1) * kC)[32] = HM_G2_FORCE_MIXED ? 2u : (!anyneg ? 0u : (!anypos ? 1u : 2u));
                            }
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int r = h * 32 + lane;
#pragma unroll
                                for (int i = 0; i < XD; ++i) {
                                    float xh_, xl_;
                                    split_scaled(xv[j][h][i], sscale, xh_, xl_);
                                    dst[i * kC + r] = xh_; dst[(XD + i) * kC + r] = xl_;
                                }
                                const float w = wv[j][h];
                                dst[(2 * XD) * kC + r] = sqrtf(fabsf(w) * wsc) * ksc;
                                const uint32_t sbit = (w < 0.f) ? 0x8000u : 0u;
                                const uint32_t nxt = __shfl_down_sync(0xffffffffu, sbit, 1);
                                if (!(lane & 1)) reinterpret_cast<uint32_t*>(dst + (2 * XD + 1) * kC)[r >> 1] = sbit | (nxt << 16);
                                if (NV > 0) dst[(2 * XD + 2) * kC + r] = vv[j][h];
                            }
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
    cluster_sync();
    if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512u);
}

// Sum the partial tiles of every (q, block job) in slot order; write H (lower from the tile, mirrored) and g.
__global__ void tc_gram2_reduce_kernel(const double* __restrict__ slots, const HmGramJob* __restrict__ jobs,
                                       const int2* __restrict__ jobslots, int njobs, int nV, double* H, double* g0, int M, int Mp,
                                       int npass) {
    const int job = blockIdx.x >> 1, r = blockIdx.x & 1, q = blockIdx.y;
    const HmGramJob jb = jobs[job];
    const int2 sr = jobslots[q * njobs + job];
    const int row0 = jb.I * 256 + r * 128;
    for (int e = blockIdx.z * 8 * 256 + threadIdx.x; e < (blockIdx.z + 1) * 8 * 256; e += blockDim.x) {   // 8 rows per CTA
        const int i = e >> 8, j = e & 255;
        const int gr = row0 + i, gc = jb.j0 + j;
        if (gc > gr || gr >= M) continue;
        double s = 0.0;
        if (HM_G2_SYMDIAG && jb.j0 == 256 * jb.I && npass >= 3) {
            // symmetrise the diagonal block: H = (Y + Y^T) / 2; Y[lj][li] lives in the slot of CTA lj / 128
            const int li = r * 128 + i, lj = j;
            const int r2 = lj >> 7, i2 = lj & 127;
            for (int sl = sr.x; sl < sr.y; ++sl)
                s += 0.5 * (slots[(size_t)(2 * sl + r) * HM_GRAM_SLOT_DOUBLES + slot_index(i, j)] +
                            slots[(size_t)(2 * sl + r2) * HM_GRAM_SLOT_DOUBLES + slot_index(i2, li)]);
        } else
        for (int sl = sr.x; sl < sr.y; ++sl) s += slots[(size_t)(2 * sl + r) * HM_GRAM_SLOT_DOUBLES + slot_index(i, j)];
        if (gr == gc)   // the lo x lo term of the diagonal (see the generator)
            for (int sl = sr.x; sl < sr.y; ++sl)
                for (int rp = 0; rp < 4; ++rp) s += slots[(size_t)(2 * sl + r) * HM_GRAM_SLOT_DOUBLES + (size_t)128 * 256 + 512 + rp * 128 + i];
        H[((size_t)q * Mp + gr) * Mp + gc] = s;
        H[((size_t)q * Mp + gc) * Mp + gr] = s;
    }
    if (jb.j0 == 0 && nV > 0 && blockIdx.z == 0) {
        for (int i = threadIdx.x; i < 128; i += blockDim.x) {
            if (row0 + i >= M) continue;
            double s = 0.0;
            for (int sl = sr.x; sl < sr.y; ++sl)
                for (int rp = 0; rp < 4; ++rp) s += slots[(size_t)(2 * sl + r) * HM_GRAM_SLOT_DOUBLES + (size_t)128 * 256 + rp * 128 + i];
            g0[(size_t)q * Mp + row0 + i] = s;
        }
    }
}

size_t gram2_smem_bytes() {
    return (size_t)kStages * kStageBytes + sizeof(float) * kRowSlots * kC * kRowArrays + sizeof(Bars) + 64 + 1024;
}

template <int XD, int NV>
int launch_gram2(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                 const int* seg_off, double* slots, int npairs, int f1, int f2, int npass) {
    const size_t smem = gram2_smem_bytes();
    auto kern = tc_gram2_kernel<XD, NV>;
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * npairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HM_CUDA(cudaLaunchKernelEx(&cfg, kern, tk, a, info, segs, seg_off, slots, f1, f2, npass));
    HM_CUDA(cudaGetLastError());
    return 0;
}

template <int XD>
int launch_gram2v(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                  const int* seg_off, int nV, double* slots, int npairs, int f1, int f2, int npass) {
    if (nV == 0) return launch_gram2<XD, 0>(s, tk, a, info, segs, seg_off, slots, npairs, f1, f2, npass);
    if (nV == 1) return launch_gram2<XD, 1>(s, tk, a, info, segs, seg_off, slots, npairs, f1, f2, npass);
    hm_set_error("gram launch: nV=%d unsupported", nV);
    return HMOGP_ERR_ARG;
}

}  // namespace

// nV = 1: H and g = K^T mu (VE step); nV = 0: H only.  segs/seg_off: plan per CTA pair (engine.cu build_gram_plan).
int hm_tc_gram2(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                const int* seg_off, int nV, double* slots, int npairs, int f1, int f2, int npass) {
    static_assert(2 * 4 * 128 <= HM_GRAM_SLOT_DOUBLES - 128 * 256, "g and diagonal partials of the 4 row parts must fit the slot");
    if (a.Mc % 256 != 0) { hm_set_error("pair Gram needs the padded M to be a multiple of 256 (got %d)", a.Mc); return HMOGP_ERR_ARG; }
    switch (a.Xdim) {
        case 1: return launch_gram2v<1>(s, tk, a, info, segs, seg_off, nV, slots, npairs, f1, f2, npass);
        case 2: return launch_gram2v<2>(s, tk, a, info, segs, seg_off, nV, slots, npairs, f1, f2, npass);
        case 3: return launch_gram2v<3>(s, tk, a, info, segs, seg_off, nV, slots, npairs, f1, f2, npass);
        case 4: return launch_gram2v<4>(s, tk, a, info, segs, seg_off, nV, slots, npairs, f1, f2, npass);
    }
    hm_set_error("Xdim=%d unsupported", a.Xdim);
    return HMOGP_ERR_ARG;
}

int hm_tc_gram2_reduce(cudaStream_t s, const double* slots, const HmGramJob* jobs, const int2* jobslots, int njobs, int Q, int nV,
                       double* H, double* g0, int M, int Mp, int npass) {
    dim3 grid((unsigned)(2 * njobs), (unsigned)Q, 16u);
    tc_gram2_reduce_kernel<<<grid, 256, 0, s>>>(slots, jobs, jobslots, njobs, nV, H, g0, M, Mp, npass);
    HM_CUDA(cudaGetLastError());
    return 0;
}

#if HM_G2_TRACE
extern "C" int hmogp_debug_g2_trace(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g2_trace, sizeof(long long) * 8 * 2048);
}
#endif
