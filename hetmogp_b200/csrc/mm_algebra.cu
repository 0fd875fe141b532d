// fp64 M x M algebra of the HetMOGP hot path: batched GEMM, blocked Cholesky, triangular inverse, K_uu build.
//
// Replaces the LAPACK/BLAS calls the reference makes through GPy/scipy:
//   dpotrf via GPy linalg.jitchol   /root/reference/hetmogp/util.py:198
//   dpotri via GPy linalg.dpotri    util.py:199, svmogp_inf.py:124
//   dgemm  via numpy.dot            svmogp_inf.py:120,130-161,236,245-246
//   RBF.K(Z_q, Z_q)                 util.py:197
// All matrices are row-major with leading dimension Mp (M padded to 256*2^k with an identity block, so no
// kernel needs edge handling and the padded block factors/inverts to identity).
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ GEMM
// C = alpha * op(A) op(B) + beta * C ; 64x64x16 tiles, 256 threads = 8 warps of 32 x 16 outputs on the fp64 tensor
// cores (mma.sync m8n8k4: 4 x 2 blocks per warp, fragments read from the k-major shared-memory tiles with conflict-free
// 64-bit loads; the 4 x 4 DFMA register tile this replaces spent two shared-memory wavefronts per FMA-cycle and ran at
// 16 % of the fp64 rate).  The next k-tile is fetched from global memory (128-bit loads) into registers while the current
// one is multiplied.  M, N multiples of 4, K a multiple of 16 (every call site: padded M, 64 * 2^k blocks, 32-wide
// Cholesky panels).
// flags (HM_GEMM_*): LOWER  only tiles on or below the block diagonal;  MIRROR  LOWER + the transposed tile is written
// too (symmetric products);  K_GE / K_LE / KB_GE  restrict the k range of a tile to where triangular operands are non-zero.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                    int lda, int64_t sA, int64_t subA, const double* __restrict__ B,
                                                    int ldb, int64_t sB, int64_t subB, double beta, double* C, int ldc,
                                                    int64_t sC, int64_t subC, int nsub, int flags) {
    constexpr int BM = 64, BN = 64, BK = 16, LD = BM + 4;
    if ((flags & (HM_GEMM_LOWER | HM_GEMM_MIRROR)) && blockIdx.x > blockIdx.y) return;
    {
        const int b = blockIdx.z, q = b / nsub, j = b % nsub;
        A += q * sA + j * subA;
        B += q * sB + j * subB;
        C += q * sC + j * subC;
    }
    __shared__ __align__(16) double As[BK][LD];
    __shared__ __align__(16) double Bs[BK][LD];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    int kbeg = 0, kend = K;
    if (flags & HM_GEMM_K_GE) kbeg = row0 > col0 ? row0 : col0;
    if (flags & HM_GEMM_KB_GE) kbeg = col0;
    if (flags & HM_GEMM_K_LE) { const int m = (row0 < col0 ? row0 : col0) + BM; kend = m < K ? m : K; }

    // this thread's share of a k-tile: 4 elements of op(A) and 4 of op(B), contiguous in memory
    const int ai = TA ? (tid & 15) * 4 : (tid & 63), ak = TA ? (tid >> 4) : (tid >> 6) * 4;
    const int bj = TB ? (tid & 63) : (tid & 15) * 4, bk = TB ? (tid >> 6) * 4 : (tid >> 4);
    const bool a_ok = row0 + ai < M, b_ok = col0 + bj < N;
    const double* ap = TA ? A + (int64_t)ak * lda + row0 + ai : A + (int64_t)(row0 + ai) * lda + ak;
    const double* bp = TB ? B + (int64_t)(col0 + bj) * ldb + bk : B + (int64_t)bk * ldb + col0 + bj;
    const int64_t astep = TA ? (int64_t)BK * lda : BK, bstep = TB ? BK : (int64_t)BK * ldb;
    double2 ra0, ra1, rb0, rb1;
    auto fetch = [&](int k0) {
        ra0 = ra1 = rb0 = rb1 = make_double2(0.0, 0.0);
        if (a_ok) {
            const double2* p = reinterpret_cast<const double2*>(ap + (int64_t)(k0 / BK) * astep);
            ra0 = p[0]; ra1 = p[1];
        }
        if (b_ok) {
            const double2* p = reinterpret_cast<const double2*>(bp + (int64_t)(k0 / BK) * bstep);
            rb0 = p[0]; rb1 = p[1];
        }
    };
    auto stash = [&]() {
        if (TA) {
            *reinterpret_cast<double2*>(&As[ak][ai]) = ra0;
            *reinterpret_cast<double2*>(&As[ak][ai + 2]) = ra1;
        } else {
            As[ak][ai] = ra0.x; As[ak + 1][ai] = ra0.y; As[ak + 2][ai] = ra1.x; As[ak + 3][ai] = ra1.y;
        }
        if (TB) {
            Bs[bk][bj] = rb0.x; Bs[bk + 1][bj] = rb0.y; Bs[bk + 2][bj] = rb1.x; Bs[bk + 3][bj] = rb1.y;
        } else {
            *reinterpret_cast<double2*>(&Bs[bk][bj]) = rb0;
            *reinterpret_cast<double2*>(&Bs[bk][bj + 2]) = rb1;
        }
    };
    // warp tile: rows [wm * 32, +32) x columns [wn * 16, +16) of the CTA tile, as 4 x 2 blocks of m8n8
    const int warp = tid >> 5, lane = tid & 31, wm = warp & 1, wn = warp >> 1;
    const int fr = lane >> 2, fk = lane & 3;          // fragment coordinates: A[row fr][k fk], B[k fk][col fr]
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        stash();
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk + fk][wm * 32 + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Bs[kk + fk][wn * 16 + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                 : "d"(a[i]), "d"(b[j]));
        }
        __syncthreads();
    }
    const bool mirror = (flags & HM_GEMM_MIRROR) && blockIdx.x != blockIdx.y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = row0 + wm * 32 + i * 8 + fr;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = col0 + wn * 16 + j * 8 + 2 * fk + e;
                if (gj >= N) continue;
                double* c = C + (int64_t)gi * ldc + gj;
                double v = alpha * acc[i][j][e];
                if (beta != 0.0) v += beta * (*c);
                *c = v;
                if (mirror) C[(int64_t)gj * ldc + gi] = v;
            }
        }
    }
}

int hm_dgemm(cudaStream_t s, bool ta, bool tb, int M, int N, int K, double alpha, const double* A, int lda, int64_t sA,
             const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC, int batch, int nsub,
             int64_t subA, int64_t subB, int64_t subC, int flags) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    if ((M | N) % 4 != 0 || K % 16 != 0 || (lda | ldb) % 2 != 0 || ((sA | sB | subA | subB) & 1) ||
        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15)) {
        hm_set_error("hm_dgemm: unsupported shape / alignment (M=%d N=%d K=%d)", M, N, K);
        return HMOGP_ERR_ARG;
    }
    dim3 grid((unsigned)hm_cdiv(N, 64), (unsigned)hm_cdiv(M, 64), (unsigned)(batch * nsub));
#define HM_LAUNCH_GEMM(TA_, TB_)                                                                                      \
    dgemm_kernel<TA_, TB_><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, sA, subA, B, ldb, sB, subB, beta, C, ldc, sC, \
                                                subC, nsub, flags)
    if (!ta && !tb) HM_LAUNCH_GEMM(false, false);
    else if (!ta && tb) HM_LAUNCH_GEMM(false, true);
    else if (ta && !tb) HM_LAUNCH_GEMM(true, false);
    else HM_LAUNCH_GEMM(true, true);
#undef HM_LAUNCH_GEMM
    HM_CUDA(cudaGetLastError());
    return 0;
}

// -------------------------------------------------------------------------------------------- Cholesky
// Right-looking blocked lower Cholesky, NB = 32.  Per panel: (1) every CTA factors the 32x32 diagonal block
// redundantly in shared memory and solves its own 128 rows against it, (2) trailing update by dgemm.
// A non-positive (or NaN) pivot sets flags[q] and is replaced by 1 so the factorisation runs to the end; the
// host retries with jitter (GPy jitchol semantics, /root/reference/hetmogp/util.py:198).
__global__ void __launch_bounds__(128) chol_panel_kernel(double* A, int ld, int64_t sQ, int k0, int n, int* flags) {
    constexpr int NB = 32;
    A += blockIdx.y * sQ;
    __shared__ double D[NB][NB + 1];
    __shared__ double invd[NB];   // 1 / D[j][j]: fp64 division and square root are ~350-cycle dependent sequences, and the
                                  // panel is one long dependent chain of them; one rsqrt per pivot, multiplications elsewhere
    const int tid = threadIdx.x;
    // the panel row this thread solves is requested first: its HBM/L2 round trip overlaps the diagonal block's factorisation
    const int row = k0 + blockIdx.x * 128 + tid;
    double* arow = A + (int64_t)row * ld + k0;
    double x[NB];
    if (row < n && row >= k0 + NB) {
#pragma unroll
        for (int j = 0; j < NB; ++j) x[j] = arow[j];
    }
    for (int e = tid; e < NB * NB; e += 128) {
        const int i = e / NB, j = e % NB;
        D[i][j] = A[(int64_t)(k0 + i) * ld + k0 + j];
    }
    __syncthreads();
    if (tid < 32) {
        // Right-looking factorisation of the 32 x 32 diagonal block with lane = row and the row held in registers: per
        // pivot one shuffle (a_jj), one reciprocal square root, the scaled column through shared memory, and a rank-one
        // update of the trailing row entries (broadcast loads + FMAs, independent of each other).  The left-looking
        // version this replaces walked a dependent chain of shared-memory dot products per pivot (~25 us per panel,
        // 16 panels: the largest item of the M-sized serial chain).
        const int lane = tid;
        double a[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) a[j] = D[lane][j];
        __shared__ double colj[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double piv = __shfl_sync(0xffffffffu, a[j], j);
            if (!(piv > 0.0)) {
                if (lane == j && blockIdx.x == 0) atomicOr(&flags[blockIdx.y], 1);
                piv = 1.0;
            }
            const double r = rsqrt(piv);
            const double lij = (lane == j) ? piv * r : a[j] * r;
            a[j] = lij;
            if (lane == 0) invd[j] = r;
            colj[lane] = lij;
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < NB; ++k) a[k] -= lij * colj[k];
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) D[lane][j] = (j <= lane) ? a[j] : 0.0;
    }
    __syncthreads();
    if (row >= n) return;
    if (row < k0 + NB) {
        const int i = row - k0;
        for (int j = 0; j < NB; ++j) arow[j] = (j <= i) ? D[i][j] : 0.0;
    } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) {   // four independent partial sums per entry
            double s0 = x[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int l = 0; l < j; ++l) {
                const double p = x[l] * D[j][l];
                if ((l & 3) == 0) s0 -= p; else if ((l & 3) == 1) s1 -= p; else if ((l & 3) == 2) s2 -= p; else s3 -= p;
            }
            x[j] = ((s0 + s1) + (s2 + s3)) * invd[j];
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) arow[j] = x[j];
    }
}

int hm_cholesky(cudaStream_t s, double* A, int Mp, int64_t sQ, int Q, int* flags) {
    constexpr int NB = 32;
    for (int k0 = 0; k0 < Mp; k0 += NB) {
        dim3 grid((unsigned)hm_cdiv(Mp - k0, 128), (unsigned)Q);
        chol_panel_kernel<<<grid, 128, 0, s>>>(A, Mp, sQ, k0, Mp, flags);
        HM_CUDA(cudaGetLastError());
        const int rem = Mp - k0 - NB;
        if (rem > 0) {
            const double* P = A + (int64_t)(k0 + NB) * Mp + k0;
            double* C = A + (int64_t)(k0 + NB) * Mp + k0 + NB;
            HM_CHECK(hm_dgemm(s, false, true, rem, rem, NB, -1.0, P, Mp, sQ, P, Mp, sQ, 1.0, C, Mp, sQ, Q, 1, 0, 0, 0,
                              HM_GEMM_LOWER));
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------ triangular inverse
// X = L^-1 for lower-triangular L: invert the 64x64 diagonal blocks, then merge pairs of inverted blocks of
// size s = 64, 128, ... with X21 = -X22 (L21 X11) (two batched GEMMs per level).  X must be zero on entry.
__global__ void __launch_bounds__(64) triinv_diag_kernel(const double* __restrict__ L, double* X, int ld, int64_t sQ) {
    constexpr int NB = 64;
    L += blockIdx.y * sQ;
    X += blockIdx.y * sQ;
    const int o = blockIdx.x * NB;
    extern __shared__ double triinv_smem[];
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(triinv_smem);
    for (int e = threadIdx.x; e < NB * NB; e += 64) {
        const int i = e / NB, j = e % NB;
        Ls[i][j] = L[(int64_t)(o + i) * ld + o + j];
    }
    __syncthreads();
    const int c = threadIdx.x;
    __shared__ double rdiag[NB];   // reciprocal diagonal, computed in parallel: no division in the serial sweep below
    rdiag[c] = 1.0 / Ls[c][c];
    __syncthreads();
    // Thread c solves L x = e_c by forward substitution with x in REGISTERS (fully unrolled: 2016 FMAs on broadcast loads of
    // L; entries above the diagonal come out as exact zeros, so control flow is uniform).  The version this replaces kept x
    // in shared memory and walked dependent load -> FMA chains: 45-60 us for the 64 x 64 blocks, on the critical path of the
    // prepare chain.
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int l = 0; l < i; ++l) {
            const double p = Ls[i][l] * x[l];
            if ((l & 3) == 0) s0 -= p; else if ((l & 3) == 1) s1 -= p; else if ((l & 3) == 2) s2 -= p; else s3 -= p;
        }
        x[i] = ((s0 + s1) + (s2 + s3)) * rdiag[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) X[(int64_t)(o + i) * ld + o + c] = (i >= c) ? x[i] : 0.0;
}

int hm_tri_inverse(cudaStream_t s, const double* L, double* X, double* tmp, int Mp, int64_t sQ, int Q) {
    HM_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * sQ * Q, s));
    dim3 grid((unsigned)(Mp / 64), (unsigned)Q);
    const int triinv_bytes = 64 * 65 * (int)sizeof(double);
    HM_CUDA(cudaFuncSetAttribute(triinv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, triinv_bytes));
    triinv_diag_kernel<<<grid, 64, triinv_bytes, s>>>(L, X, Mp, sQ);
    HM_CUDA(cudaGetLastError());
    for (int sz = 64; sz < Mp; sz *= 2) {
        const int npairs = Mp / (2 * sz);
        const int64_t sub = (int64_t)2 * sz * (Mp + 1);
        const int64_t o21 = (int64_t)sz * Mp;           // block (1,0) of a pair
        const int64_t o22 = (int64_t)sz * Mp + sz;      // block (1,1)
        // tmp21 = L21 * X11
        HM_CHECK(hm_dgemm(s, false, false, sz, sz, sz, 1.0, L + o21, Mp, sQ, X, Mp, sQ, 0.0, tmp + o21, Mp, sQ, Q,
                          npairs, sub, sub, sub));
        // X21 = -X22 * tmp21
        HM_CHECK(hm_dgemm(s, false, false, sz, sz, sz, -1.0, X + o22, Mp, sQ, tmp + o21, Mp, sQ, 0.0, X + o21, Mp, sQ,
                          Q, npairs, sub, sub, sub));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- K_uu
// GPy RBF.K(Z_q, Z_q): sigma^2 exp(-r^2/2) with the diagonal of r^2 forced to 0;
// identity in the padded block; optional jitter on the diagonal (jitchol retry).
__global__ void build_kuu_kernel(const double* __restrict__ Zp, const HmConsts* __restrict__ c,
                                 const double* __restrict__ jitter, double* Kuu, int M, int Mp, int Xdim) {
    const int q = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= Mp) return;
    double v;
    if (i < M && j < M) {
        if (i == j) v = c->var[q] + (jitter ? jitter[q] : 0.0);
        else {
            const double* zi = Zp + ((int64_t)q * Mp + i) * Xdim;
            const double* zj = Zp + ((int64_t)q * Mp + j) * Xdim;
            double r2 = 0.0;
            for (int k = 0; k < Xdim; ++k) { const double d = zi[k] - zj[k]; r2 += d * d; }
            v = c->var[q] * exp(-0.5 * r2 * c->inv_l2[q]);
        }
    } else v = (i == j) ? 1.0 : 0.0;
    Kuu[((int64_t)q * Mp + i) * Mp + j] = v;
}

int hm_build_kuu(cudaStream_t s, const double* Zp, const HmConsts* c, const double* jitter, double* Kuu, int M, int Mp,
                 int Xdim, int Q) {
    dim3 grid((unsigned)hm_cdiv(Mp, 128), (unsigned)Mp, (unsigned)Q);
    build_kuu_kernel<<<grid, 128, 0, s>>>(Zp, c, jitter, Kuu, M, Mp, Xdim);
    HM_CUDA(cudaGetLastError());
    return 0;
}
