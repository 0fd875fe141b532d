// Weighted Gram statistic of the backward pass, SIMT version:
//     H1_q[m, m'] = sum_t sum_n K_tq[n, m] omega_tq[n] K_tq[n, m']            (M x M, symmetric)
// from which dVE/dS_q = K_uu^-1 H1_q K_uu^-1 (reference: A^T diag(dv) A per output function,
// /root/reference/hetmogp/svmogp_inf.py:145-148, summed over d with W_dq^2 folded into omega).
//
// The K tiles are regenerated from (X, Z_q) in shared memory, never read from HBM.  Grid = (lower-triangular
// output tile pairs) x Q x nsplit; each CTA owns one output tile for a contiguous range of 16-row blocks and
// accumulates in registers (fp32 mode: flushed to its private fp64 partial tile every 8192 rows), so the
// reduction over splits is deterministic (hm_gram_reduce in engine.cu).
#include "common.cuh"

namespace {

template <typename T> struct SplitX2 {
    static __device__ __forceinline__ void split(double x, T& hi, T& lo);
};
template <> __device__ __forceinline__ void SplitX2<double>::split(double x, double& hi, double& lo) { hi = x; lo = 0.0; }
template <> __device__ __forceinline__ void SplitX2<float>::split(double x, float& hi, float& lo) {
    hi = (float)x;
    lo = (float)(x - (double)hi);
}
template <typename T> __device__ __forceinline__ T exp_t2(T x);
template <> __device__ __forceinline__ double exp_t2<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float exp_t2<float>(float x) { return expf(x); }

constexpr int kThreads = 256;
constexpr int kBK = 16;        // rows per k-step
template <typename T> struct Chunk_ { static constexpr int v = sizeof(T) == 4 ? 32 : 8; };  // k-steps of row data staged at once
constexpr int kFlush = 512;    // k-steps between fp64 flushes (8192 rows)

template <typename T, int TT>
__global__ void __launch_bounds__(kThreads, (sizeof(T) == 4 ? 2 : 1))
gram_kernel(HmTasks tk, HmProjArgs pa, double* Hpart, int nsplit, int64_t nblk_total) {
    constexpr int BT = 16 * TT;  // output tile edge
    constexpr int kChunk = Chunk_<T>::v;
    const int q = blockIdx.y, split = blockIdx.z;
    const int Mc = pa.Mc, Mp = pa.Mp, M = pa.M, Xd = pa.Xdim, Q = pa.Q;
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    // decode the lower-triangular tile pair
    int ti = 0, tj = 0;
    {
        int p = blockIdx.x;
        while (p > ti) { p -= ti + 1; ++ti; }
        tj = p;
    }
    __shared__ __align__(16) T As[kBK][BT];
    __shared__ __align__(16) T Bs[kBK][BT];
    __shared__ T zah[BT * HM_MAXXD], zal[BT * HM_MAXXD], zbh[BT * HM_MAXXD], zbl[BT * HM_MAXXD];
    __shared__ T ksa[BT], ksb[BT];
    __shared__ T xh[kChunk * kBK * HM_MAXXD], xl[kChunk * kBK * HM_MAXXD], om[kChunk * kBK];

    const HmConsts* __restrict__ cs = pa.consts;
    const T var_q = T(cs->var[q]);
    const T nhalf_inv_l2 = T(-0.5 * cs->inv_l2[q]);
    for (int m = tid; m < BT; m += kThreads) {
        const int ga = ti * BT + m, gb = tj * BT + m;
        for (int i = 0; i < Xd; ++i) {
            const double za = (ga < M) ? pa.Zp[((size_t)q * Mp + ga) * Xd + i] : 0.0;
            const double zb = (gb < M) ? pa.Zp[((size_t)q * Mp + gb) * Xd + i] : 0.0;
            SplitX2<T>::split(za, zah[m * Xd + i], zal[m * Xd + i]);
            SplitX2<T>::split(zb, zbh[m * Xd + i], zbl[m * Xd + i]);
        }
        ksa[m] = (ga < M) ? var_q : T(0);
        ksb[m] = (gb < M) ? var_q : T(0);
    }

    T acc[TT][TT];
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
        for (int j = 0; j < TT; ++j) acc[i][j] = T(0);

    double* out = Hpart + (((size_t)split * Q + q) * Mc + (size_t)ti * BT) * Mc + (size_t)tj * BT;
    bool first = true;
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < TT; ++i) {
            const int r = (i < TT / 2 ? ty * (TT / 2) + i : BT / 2 + ty * (TT / 2) + (i - TT / 2));
#pragma unroll
            for (int j = 0; j < TT; ++j) {
                const int c = (j < TT / 2 ? tx * (TT / 2) + j : BT / 2 + tx * (TT / 2) + (j - TT / 2));
                double* o = out + (size_t)r * Mc + c;
                *o = first ? (double)acc[i][j] : (*o + (double)acc[i][j]);
                acc[i][j] = T(0);
            }
        }
        first = false;
    };

    const int64_t per = (nblk_total + nsplit - 1) / nsplit;
    int64_t vb = (int64_t)split * per;
    const int64_t vb_end = (vb + per < nblk_total) ? vb + per : nblk_total;
    int since_flush = 0;
    __syncthreads();

    while (vb < vb_end) {
        // locate the task of virtual block vb and how many blocks can be staged from it
        int t = 0;
        int64_t lb = vb;
        for (; t < tk.T; ++t) {
            const int64_t nb_t = (tk.count[t] + kBK - 1) / kBK;
            if (lb < nb_t) break;
            lb -= nb_t;
        }
        const int64_t nb_t = (tk.count[t] + kBK - 1) / kBK;
        int nb = kChunk;
        if (nb_t - lb < nb) nb = (int)(nb_t - lb);
        if (vb_end - vb < nb) nb = (int)(vb_end - vb);
        const int64_t row0 = lb * kBK;
        const T* mw = reinterpret_cast<const T*>(tk.MW[t]);
        for (int e = tid; e < nb * kBK; e += kThreads) {
            const int64_t row = row0 + e;
            const bool ok = row < tk.count[t];
            for (int i = 0; i < Xd; ++i) {
                const double x = ok ? tk.X[t][(tk.begin[t] + row) * Xd + i] : 0.0;
                SplitX2<T>::split(x, xh[e * Xd + i], xl[e * Xd + i]);
            }
            om[e] = ok ? mw[(size_t)(Q + q) * tk.cap[t] + row] : T(0);
        }
        __syncthreads();
        for (int kb = 0; kb < nb; ++kb) {
            // generate the two operand tiles for these 16 rows
            for (int e = tid; e < kBK * BT; e += kThreads) {
                const int kk = e / BT, m = e % BT;
                const int r = kb * kBK + kk;
                T d2 = T(0);
                for (int i = 0; i < Xd; ++i) {
                    const T d = (xh[r * Xd + i] - zbh[m * Xd + i]) + (xl[r * Xd + i] - zbl[m * Xd + i]);
                    d2 += d * d;
                }
                const T kbv = ksb[m] * exp_t2<T>(d2 * nhalf_inv_l2);
                Bs[kk][m] = kbv;
                T kav = kbv;
                if (ti != tj) {
                    T e2 = T(0);
                    for (int i = 0; i < Xd; ++i) {
                        const T d = (xh[r * Xd + i] - zah[m * Xd + i]) + (xl[r * Xd + i] - zal[m * Xd + i]);
                        e2 += d * d;
                    }
                    kav = ksa[m] * exp_t2<T>(e2 * nhalf_inv_l2);
                }
                As[kk][m] = kav * om[r];
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kBK; ++kk) {
                T a[TT], b[TT];
#pragma unroll
                for (int i = 0; i < TT / 2; ++i) {
                    a[i] = As[kk][ty * (TT / 2) + i];
                    a[TT / 2 + i] = As[kk][BT / 2 + ty * (TT / 2) + i];
                    b[i] = Bs[kk][tx * (TT / 2) + i];
                    b[TT / 2 + i] = Bs[kk][BT / 2 + tx * (TT / 2) + i];
                }
#pragma unroll
                for (int i = 0; i < TT; ++i)
#pragma unroll
                    for (int j = 0; j < TT; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
            if (sizeof(T) == 4 && ++since_flush >= kFlush) { flush(); since_flush = 0; }
        }
        vb += nb;
    }
    flush();
}

template <typename T, int TT>
int launch_gram(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, double* Hpart, int nsplit) {
    constexpr int BT = 16 * TT;
    const int ntile = a.Mc / BT;
    const int npairs = ntile * (ntile + 1) / 2;
    int64_t nblk = 0;
    for (int t = 0; t < tk.T; ++t) nblk += hm_cdiv(tk.count[t], kBK);
    dim3 grid((unsigned)npairs, (unsigned)a.Q, (unsigned)nsplit);
    gram_kernel<T, TT><<<grid, kThreads, 0, s>>>(tk, a, Hpart, nsplit, nblk);
    HM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int hm_gram_tile(int prec) { return prec == HMOGP_PREC_FP64 ? 64 : 128; }

int hm_gram_splits(int prec, int Mc, int Q) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int BT = hm_gram_tile(prec);
    const int ntile = Mc / BT, npairs = ntile * (ntile + 1) / 2;
    const int slots = sms * (prec == HMOGP_PREC_FP64 ? 1 : 2);
    int ns = slots / (npairs * Q);
    if (ns < 1) ns = 1;
    if (ns > 64) ns = 64;
    return ns;
}

int hm_gram(cudaStream_t s, int prec, const HmTasks& tk, const HmProjArgs& a, double* Hpart, int nsplit) {
    if (a.Xdim > HM_MAXXD) { hm_set_error("Xdim > %d unsupported", HM_MAXXD); return HMOGP_ERR_ARG; }
    if (prec == HMOGP_PREC_FP64) return launch_gram<double, 4>(s, tk, a, Hpart, nsplit);
    return launch_gram<float, 8>(s, tk, a, Hpart, nsplit);
}
