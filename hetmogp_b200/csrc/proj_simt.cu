// N-sized projection kernels, SIMT (CUDA-core) version, templated on the arithmetic type.
//
// For one latent q and a tile of BM data rows of task t the kernel builds the RBF cross-covariance tile
// K_tq = k_q(X_t, Z_q) on the fly in shared memory (never in HBM; reference materialises it: util.py:145-164)
// and contracts it with the M x M factor C_q = K_uu^-1 S_q K_uu^-1 - K_uu^-1:
//
//   forward  (svmogp_inf.py:212-218 restated):
//        a_tq[n] = K[n,:] . alpha_q            alpha_q = K_uu^-1 m_q                (:216)
//        c_tq[n] = K[n,:] C_q K[n,:]^T                                              (:217-218)
//   backward (svmogp_inf.py:144,157-161 and svmogp.py:139-141,153-156 restated):
//        g1[m]   = sum_n K[n,m] mu[n]                                -> dVE/dm_q = K_uu^-1 g1
//        GK[n,m] = K[n,m] (muc[n] alpha[m] + 2 omc[n] (K C)[n,m])     (= sum_d W'_dq dL_dKmn_d[m,n] K[n,m])
//        dls    += sum GK r^2 ;  dz[m,i] += sum_n GK[n,m] (x_ni - z_mi)
//
// Layout: CTA = 256 threads = 8 warps; warp w owns rows [w*TM, (w+1)*TM) of the tile, lane l owns columns
// {4l..4l+3} and {128+4l..128+4l+3} of the current 256-wide column block.  K tile lives transposed in shared
// memory (Ks[m][r]) so the row operands are broadcast vector loads and the C_q tile loads are conflict-free.
// C_q (L2-resident) is streamed through a 2-stage cp.async ring.  Per-CTA column accumulators are fp64 in
// shared memory and written once per CTA (deterministic two-level reduction, no global atomics).
#include "common.cuh"

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <typename T> struct SplitX {  // x = hi + lo so that differences of nearby fp64 inputs survive fp32
    static __device__ __forceinline__ void split(double x, T& hi, T& lo);
};
template <> __device__ __forceinline__ void SplitX<double>::split(double x, double& hi, double& lo) { hi = x; lo = 0.0; }
template <> __device__ __forceinline__ void SplitX<float>::split(double x, float& hi, float& lo) {
    hi = (float)x;
    lo = (float)(x - (double)hi);
}
template <typename T> __device__ __forceinline__ T exp_t(T x);
template <> __device__ __forceinline__ double exp_t<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float exp_t<float>(float x) { return expf(x); }

constexpr int kThreads = 256;
constexpr int kBN = 256;  // column block

template <typename T> struct BK_ { static constexpr int v = 16 / (sizeof(T) / 4); };  // 16 (fp32) / 8 (fp64)

struct TileRef { int t; int64_t row0; int nrows; };

__device__ __forceinline__ TileRef find_tile(const HmTasks& tk, int64_t tile, int BM) {
    TileRef r; r.t = -1; r.row0 = 0; r.nrows = 0;
    for (int t = 0; t < tk.T; ++t) {
        const int64_t nt = (tk.count[t] + BM - 1) / BM;
        if (tile < nt) {
            r.t = t; r.row0 = tile * BM;
            const int64_t rem = tk.count[t] - r.row0;
            r.nrows = rem < BM ? (int)rem : BM;
            return r;
        }
        tile -= nt;
    }
    return r;
}

template <typename T, int BM, bool BWD>
__global__ void __launch_bounds__(kThreads, 1) proj_kernel(HmTasks tk, HmProjArgs pa, int64_t ntiles, int hyper) {
    constexpr int TM = BM / 8;
    constexpr int BK = BK_<T>::v;
    const int q = blockIdx.y;
    const int Mc = pa.Mc, Mp = pa.Mp, M = pa.M, Xd = pa.Xdim, Q = pa.Q;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Ks = reinterpret_cast<T*>(smem_raw);                     // [Mc][BM]
    T* Bs = Ks + (size_t)Mc * BM;                               // [2][BK][kBN]
    T* zh = Bs + 2 * BK * kBN;                                  // [Mc][Xd]
    T* zl = zh + (size_t)Mc * Xd;                               // [Mc][Xd]
    T* ksc = zl + (size_t)Mc * Xd;                              // [Mc] sigma^2 or 0 for padded columns
    T* als = ksc + Mc;                                          // [Mc] alpha_q
    T* xh = als + Mc;                                           // [BM][Xd]
    T* xl = xh + BM * Xd;                                       // [BM][Xd]
    T* rw = xl + BM * Xd;                                       // [4][BM] mu, om, muc, omc
    double* racc = reinterpret_cast<double*>(rw + 4 * BM);      // [2][BM] a, c accumulators  (8-byte aligned: all counts even)
    double* colacc = racc + 2 * BM;                             // bwd: [(1+Xd)*Mc + 1]

    const HmConsts* __restrict__ cs = pa.consts;
    const T var_q = T(cs->var[q]);
    const T inv_l2 = T(cs->inv_l2[q]);
    const T* __restrict__ Cq = reinterpret_cast<const T*>(pa.C) + (size_t)q * Mp * Mp;

    for (int m = tid; m < Mc; m += kThreads) {
        for (int i = 0; i < Xd; ++i) {
            const double z = (m < M) ? pa.Zp[((size_t)q * Mp + m) * Xd + i] : 0.0;
            SplitX<T>::split(z, zh[m * Xd + i], zl[m * Xd + i]);
        }
        ksc[m] = (m < M) ? var_q : T(0);
        als[m] = (m < M) ? T(pa.alpha[(size_t)q * Mp + m]) : T(0);
    }
    const int ncol = (1 + Xd) * Mc + 1;
    if (BWD) for (int i = tid; i < ncol; i += kThreads) colacc[i] = 0.0;
    __syncthreads();

    double dls_thread = 0.0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileRef tr = find_tile(tk, tile, BM);
        const int t = tr.t;
        // ---- row data
        for (int e = tid; e < BM * Xd; e += kThreads) {
            const int r = e / Xd;
            const double x = (r < tr.nrows) ? tk.X[t][(tk.begin[t] + tr.row0) * Xd + e] : 0.0;
            SplitX<T>::split(x, xh[e], xl[e]);
        }
        if (BWD) {
            const T* mw = reinterpret_cast<const T*>(tk.MW[t]);
            for (int e = tid; e < 4 * BM; e += kThreads) {
                const int k = e / BM, r = e % BM;
                rw[e] = (r < tr.nrows) ? mw[(size_t)(k * Q + q) * tk.cap[t] + tr.row0 + r] : T(0);
            }
        }
        for (int e = tid; e < 2 * BM; e += kThreads) racc[e] = 0.0;
        __syncthreads();
        // ---- build the K tile (transposed) in shared memory; forward: a = K alpha on the fly
        {
            const int r = tid % BM;
            T apart = T(0);
            for (int m = tid / BM; m < Mc; m += kThreads / BM) {
                T d2 = T(0);
                for (int i = 0; i < Xd; ++i) {
                    const T d = (xh[r * Xd + i] - zh[m * Xd + i]) + (xl[r * Xd + i] - zl[m * Xd + i]);
                    d2 += d * d;
                }
                const T kv = ksc[m] * exp_t<T>(T(-0.5) * d2 * inv_l2);
                Ks[(size_t)m * BM + r] = kv;
                apart += kv * als[m];
            }
            if (!BWD) {  // deterministic reduction of the per-thread partials (scratch: the idle C_q stage buffer)
                Bs[tid] = apart;
                __syncthreads();
                if (tid < BM) {
                    double sacc = 0.0;
                    for (int g = 0; g < kThreads / BM; ++g) sacc += (double)Bs[g * BM + tid];
                    racc[tid] = sacc;
                }
            }
        }
        __syncthreads();

        if (!BWD || hyper) {
            for (int jb = 0; jb < Mc; jb += kBN) {
                T acc[TM][8];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
                // stage loader: BK rows x 256 columns of C_q
                auto load_stage = [&](int st, int k0) {
                    constexpr int per_row = kBN * sizeof(T) / 16;  // 16-byte chunks per row
                    constexpr int total = BK * per_row;
                    for (int e = tid; e < total; e += kThreads) {
                        const int kk = e / per_row, ch = e % per_row;
                        const T* src = Cq + (size_t)(k0 + kk) * Mp + jb + ch * (16 / sizeof(T));
                        cp_async16(Bs + ((size_t)st * BK + kk) * kBN + ch * (16 / sizeof(T)), src);
                    }
                    cp_async_commit();
                };
                load_stage(0, 0);
                const int nk = Mc / BK;
                for (int ks = 0; ks < nk; ++ks) {
                    cp_async_wait<0>();
                    __syncthreads();
                    if (ks + 1 < nk) load_stage((ks + 1) & 1, (ks + 1) * BK);
                    const T* bs = Bs + (size_t)(ks & 1) * BK * kBN;
                    const T* ka = Ks + (size_t)(ks * BK) * BM + w * TM;
#pragma unroll
                    for (int kk = 0; kk < BK; ++kk) {
                        T a[TM], b[8];
#pragma unroll
                        for (int i = 0; i < TM; ++i) a[i] = ka[(size_t)kk * BM + i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            b[j] = bs[kk * kBN + lane * 4 + j];
                            b[4 + j] = bs[kk * kBN + 128 + lane * 4 + j];
                        }
#pragma unroll
                        for (int i = 0; i < TM; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
                    }
                }
                __syncthreads();  // all warps done with Bs before the next block's stage 0 load
                // ---- epilogue for this column block
                if (!BWD) {
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        T cp = T(0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int col = jb + (j < 4 ? lane * 4 + j : 128 + lane * 4 + (j - 4));
                            cp += acc[i][j] * Ks[(size_t)col * BM + w * TM + i];
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) cp += __shfl_xor_sync(0xffffffffu, cp, off);
                        if (lane == 0) racc[BM + w * TM + i] += (double)cp;
                    }
                } else {
                    // column sums over the tile rows: per-warp partials -> fixed-order cross-warp sum through shared
                    // memory (scratch: the idle C_q stage buffer) -> single owner per column: deterministic
                    double* red = reinterpret_cast<double*>(Bs);   // [8 warps][256 columns]
                    for (int i2 = 0; i2 < Xd; ++i2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int slot = (j < 4 ? lane * 4 + j : 128 + lane * 4 + (j - 4));
                            const int col = jb + slot;
                            const T al = als[col];
                            T pz = T(0), dl = T(0);
#pragma unroll
                            for (int i = 0; i < TM; ++i) {
                                const int r = w * TM + i;
                                const T kv = Ks[(size_t)col * BM + r];
                                const T gk = kv * (rw[2 * BM + r] * al + T(2) * rw[3 * BM + r] * acc[i][j]);
                                const T d = (xh[r * Xd + i2] - zh[col * Xd + i2]) + (xl[r * Xd + i2] - zl[col * Xd + i2]);
                                pz += gk * d;
                                dl += gk * d * d;
                            }
                            dls_thread += (double)dl;   // sum over input dims of GK d_i^2 = GK |x - z|^2
                            red[w * kBN + slot] = (double)pz;
                        }
                        __syncthreads();
                        {
                            double sacc = 0.0;
#pragma unroll
                            for (int ww = 0; ww < kThreads / 32; ++ww) sacc += red[ww * kBN + tid];
                            colacc[(1 + i2) * Mc + jb + tid] += sacc;
                        }
                        __syncthreads();
                    }
                }
            }
        }
        if (BWD) {
            // g1[m] += sum_r K[r,m] mu[r]  (column sums of the tile; no contraction needed)
            for (int m = tid; m < Mc; m += kThreads) {
                T sacc = T(0);
                const T* kr = Ks + (size_t)m * BM;
#pragma unroll 8
                for (int r = 0; r < BM; ++r) sacc += kr[r] * rw[r];
                colacc[m] += (double)sacc;
            }
        } else {
            __syncthreads();
            T* ac = reinterpret_cast<T*>(tk.AC[t]);
            for (int r = tid; r < tr.nrows; r += kThreads) {
                ac[(size_t)q * tk.cap[t] + tr.row0 + r] = T(racc[r]);
                ac[(size_t)(Q + q) * tk.cap[t] + tr.row0 + r] = T(racc[BM + r]);
            }
        }
        __syncthreads();
    }
    if (BWD) {
        // dls: fixed-order block reduce
        double* red = reinterpret_cast<double*>(Bs);
        red[tid] = dls_thread;
        __syncthreads();
        if (tid == 0) {
            double sacc = 0.0;
            for (int i = 0; i < kThreads; ++i) sacc += red[i];
            colacc[ncol - 1] = sacc;
        }
        __syncthreads();
        double* out = pa.colpart + ((size_t)q * gridDim.x + blockIdx.x) * ncol;
        for (int i = tid; i < ncol; i += kThreads) out[i] = colacc[i];
    }
}

template <typename T, int BM> size_t proj_smem(int Mc, int Xd, bool bwd) {
    constexpr int BK = BK_<T>::v;
    size_t n = (size_t)Mc * BM + 2 * BK * kBN + 2 * (size_t)Mc * Xd + 2 * Mc + 2 * BM * Xd + 4 * BM;
    size_t bytes = n * sizeof(T);
    bytes = (bytes + 7) & ~(size_t)7;
    bytes += sizeof(double) * (2 * BM + (bwd ? ((1 + Xd) * (size_t)Mc + 1) : 0));
    return bytes;
}

template <typename T, int BM, bool BWD>
int launch_proj(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, bool hyper) {
    int64_t ntiles = 0;
    for (int t = 0; t < tk.T; ++t) ntiles += hm_cdiv(tk.count[t], BM);
    if (ntiles == 0 && !BWD) return 0;
    const size_t smem = proj_smem<T, BM>(a.Mc, a.Xdim, BWD);
    auto kern = proj_kernel<T, BM, BWD>;
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nw = a.nworkers;
    if (!BWD && ntiles < nw) nw = (int)ntiles;  // bwd keeps the full worker count: every partial slot is written
    dim3 grid((unsigned)nw, (unsigned)a.Q);
    kern<<<grid, kThreads, smem, s>>>(tk, a, ntiles, hyper ? 1 : 0);
    HM_CUDA(cudaGetLastError());
    return 0;
}

// largest row tile whose whole shared-memory footprint (K tile, C_q stages, tables, fp64 accumulators) fits one SM
template <typename T, bool BWD> int pick_bm(int Mc, int Xd) {
    const size_t limit = 227 * 1024;
    if constexpr (sizeof(T) == 4) {
        if (proj_smem<T, 64>(Mc, Xd, BWD) <= limit) return 64;
    }
    if (proj_smem<T, 32>(Mc, Xd, BWD) <= limit) return 32;
    if (proj_smem<T, 16>(Mc, Xd, BWD) <= limit) return 16;
    if (proj_smem<T, 8>(Mc, Xd, BWD) <= limit) return 8;
    return 0;
}

template <typename T, bool BWD> int dispatch_proj(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, bool hyper) {
    const int bm = (a.Xdim <= HM_MAXXD) ? pick_bm<T, BWD>(a.Mc, a.Xdim) : 0;
    if (bm == 0) {
        hm_set_error("projection kernel: M=%d (padded %d) or Xdim=%d exceeds the shared-memory tile budget", a.M, a.Mc, a.Xdim);
        return HMOGP_ERR_ARG;
    }
    if constexpr (sizeof(T) == 4) {
        if (bm == 64) return launch_proj<T, 64, BWD>(s, tk, a, hyper);
    }
    if (bm == 32) return launch_proj<T, 32, BWD>(s, tk, a, hyper);
    if (bm == 16) return launch_proj<T, 16, BWD>(s, tk, a, hyper);
    return launch_proj<T, 8, BWD>(s, tk, a, hyper);
}

}  // namespace

int hm_proj_workers(int prec, int Mc) {
    (void)prec; (void)Mc;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;  // one persistent CTA per SM and latent (grid.y = Q) -> Q waves of equal work
}

int hm_proj_fwd(cudaStream_t s, int prec, const HmTasks& tk, const HmProjArgs& a) {
    if (prec == HMOGP_PREC_FP64) return dispatch_proj<double, false>(s, tk, a, true);
    return dispatch_proj<float, false>(s, tk, a, true);
}

int hm_proj_bwd(cudaStream_t s, int prec, const HmTasks& tk, const HmProjArgs& a, bool hyper) {
    if (prec == HMOGP_PREC_FP64) return dispatch_proj<double, true>(s, tk, a, hyper);
    return dispatch_proj<float, true>(s, tk, a, hyper);
}
