// Shared declarations of the hetmogp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hetmogp_b200.h"

#define HM_MAXQ HMOGP_MAX_Q
#define HM_MAXT HMOGP_MAX_TASKS
#define HM_MAXJ HMOGP_MAX_J
#define HM_MAXF HMOGP_MAX_DIMF
#define HM_MAXXD 4

void hm_set_error(const char* fmt, ...);
// Every kernel launch in this library is followed by HM_CUDA(cudaGetLastError()); that convention doubles as the
// launch counter reported to bench.py (gpu_launches).
extern long long hm_launch_counter;
static inline void hm_note_call(const char* text) {
    if (text[0] == 'c' && text[4] == 'G' && text[7] == 'L' && text[11] == 'E') ++hm_launch_counter;  // "cudaGetLastError()"
}

#define HM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        hm_note_call(#call);                                                                   \
        if (_e != cudaSuccess) {                                                               \
            hm_set_error("%s:%d CUDA error: %s (%s)", __FILE__, __LINE__, cudaGetErrorString(_e), #call); \
            return HMOGP_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

#define HM_CHECK(call)                 \
    do {                               \
        int _r = (call);               \
        if (_r != 0) return _r;        \
    } while (0)

static inline int64_t hm_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Per-step scalar constants, filled on the device by hm_prep_consts (no host round trip).
struct HmConsts {
    double var[HM_MAXQ];     // sigma_q^2
    double ls[HM_MAXQ];      // l_q
    double inv_l2[HM_MAXQ];  // 1/l_q^2
    double W[HM_MAXJ][HM_MAXQ];
    double kappa[HM_MAXJ][HM_MAXQ];
    double Wc[HM_MAXJ][HM_MAXQ];  // chain multipliers (quirk C-5), = W by default
    double kc[HM_MAXJ][HM_MAXQ];
    double kdiag[HM_MAXJ];        // sum_q (W_dq^2 + kappa_dq) sigma_q^2   (util.py:178 diagonal)
    double bscale[HM_MAXT];
};

// Task table shared by the N-sized kernels.
struct HmTasks {
    int T, Q, Xdim, J;
    int kind[HM_MAXT], K[HM_MAXT], dimf[HM_MAXT], foff[HM_MAXT];
    double sigma[HM_MAXT];
    const double* X[HM_MAXT];  // [N_t, Xdim] resident fp64 rows
    const double* Y[HM_MAXT];  // [N_t]
    int64_t begin[HM_MAXT];    // active slice
    int64_t count[HM_MAXT];
    // per-row intermediates, structure-of-arrays so that every pass streams full sectors:  array k of task t starts at
    // base + k * cap[t], row r of the active slice at [r]
    void* AC[HM_MAXT];         // [acs][cap]  k = Q*0+q: a_tq, Q*1+q: c_tq (+ Q*2+q: b_tq, Q*3+q: e_tq on the tensor-core path)
    void* MW[HM_MAXT];         // [4Q][cap]   k = Q*0+q: mu, Q*1+q: omega, Q*2+q: mu_c, Q*3+q: omega_c
    int64_t cap[HM_MAXT];      // allocated rows per array
    int acs;                   // arrays in AC: 2Q (SIMT) or 4Q (tensor-core path)
};

// Per-step scale state of the tensor-core path (device resident; see tc_common.cuh "split fp16").
struct HmTcInfo {
    int cexp[HM_MAXQ];             // C_q is carried as C_q * 2^cexp   (|.| < 2^14)
    int kexp[HM_MAXQ];             // K_tq is carried as K_tq * 2^kexp (sigma_q^2 -> [2^11, 2^12))
    unsigned wmax[2][HM_MAXQ];     // float bits of max_n |omega_tq|, max_n |omega^c_tq| (likelihood kernel, atomicMax)
    unsigned cmax[HM_MAXQ];        // float bits of max |C_q|
};

// ---------------------------------------------------------------- M x M fp64 algebra (mm_algebra.cu)
// Row-major, leading dimension ld, batch of Q matrices with stride sQ.
int hm_dgemm(cudaStream_t s, bool ta, bool tb, int M, int N, int K, double alpha, const double* A, int lda, int64_t sA,
             const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC, int batch,
             int nsub = 1, int64_t subA = 0, int64_t subB = 0, int64_t subC = 0, int flags = 0);
enum {
    HM_GEMM_LOWER = 1,    // only the 64x64 tiles on or below the block diagonal are computed
    HM_GEMM_MIRROR = 2,   // symmetric product: lower tiles computed, their transposes written too
    HM_GEMM_K_GE = 4,     // op(A)[i,k] = 0 for k < i and op(B)[k,j] = 0 for k < j   (X^T X, X lower-triangular)
    HM_GEMM_K_LE = 8,     // op(A)[i,k] = 0 for k > i and op(B)[k,j] = 0 for k > j   (L L^T)
    HM_GEMM_KB_GE = 16    // op(B)[k,j] = 0 for k < j                                (. L, L lower-triangular)
};
int hm_cholesky(cudaStream_t s, double* A, int Mp, int64_t sQ, int Q, int* flags);       // in place, lower
int hm_tri_inverse(cudaStream_t s, const double* L, double* X, double* tmp, int Mp, int64_t sQ, int Q);
int hm_build_kuu(cudaStream_t s, const double* Zp, const HmConsts* c, const double* jitter, double* Kuu, int M, int Mp,
                 int Xdim, int Q);

// ---------------------------------------------------------------- likelihood kernels (lik_kernels.cu)
struct HmLikStatsLayout {
    int per_task;  // doubles per task in the partial/stat block: [VE, nneg, sdv[F], sma[F][Q], svc[F][Q]]
};
int hm_lik_rows(cudaStream_t s, int prec, const HmTasks& tk, const HmConsts* consts, int t, bool want_grads,
                bool has_chain, double* partials, int max_blocks, int* nblocks_out, double* rows_m, double* rows_v,
                double* rows_ve, double* rows_dm, double* rows_dv, HmTcInfo* tcinfo = nullptr, bool hyper = false);
int hm_lik_var_exp(cudaStream_t s, int prec, const hmogp_lik_desc& lik, int64_t N, const double* Y, const double* Mf,
                   const double* Vf, double* VE, double* dm, double* dv);
int hm_lik_pointwise(cudaStream_t s, const hmogp_lik_desc& lik, int64_t N, const double* F, const double* Y,
                     double* logp, double* dlogp, double* d2logp);
int hm_lik_predictive(cudaStream_t s, const hmogp_lik_desc& lik, int gh_tensor, int64_t N, const double* Mf, const double* Vf,
                      double* mean_pred, double* var_pred);
int hm_upload_gh_tables();

// ---------------------------------------------------------------- N-sized SIMT contractions (proj_simt.cu, gram_simt.cu)
struct HmProjArgs {
    int M, Mp, Mc, Q, Xdim;
    const double* Zp;      // [Q][Mp][Xdim] padded inducing inputs
    const double* alpha;   // [Q][Mp]
    const void* C;         // [Q][Mp][Mp] in the compute type
    const HmConsts* consts;
    double* colpart;       // bwd: [Q][nworkers][ (1+Xdim)*Mc + 1 ]
    int nworkers;
};
int hm_proj_fwd(cudaStream_t s, int prec, const HmTasks& tk, const HmProjArgs& a);
int hm_proj_bwd(cudaStream_t s, int prec, const HmTasks& tk, const HmProjArgs& a, bool hyper);
int hm_proj_workers(int prec, int Mc);
int hm_gram(cudaStream_t s, int prec, const HmTasks& tk, const HmProjArgs& a, double* Hpart, int nsplit);
int hm_gram_splits(int prec, int Mc, int Q);
int hm_gram_tile(int prec);

// ---------------------------------------------------------------- tensor-core path (tc_fwd.cu, tc_gram.cu)
#define HM_GRAM_CHUNK 32                                   // data rows per Gram stage
#define HM_GRAM_MAXV 6                                     // g-vectors per launch
#define HM_GRAM_ROWSPLIT 2                                 // generator warps per column group (partial g-vectors per slot)
#define HM_GRAM_SLOT_DOUBLES (128 * 256 + HM_GRAM_ROWSPLIT * HM_GRAM_MAXV * 128)
struct HmGramJob { int I, j0, nw; };                       // output tile: rows [128 I, 128 I + 128), columns [j0, j0 + nw)
struct HmGramSeg { int q, I, j0, nw, chunk_begin, chunk_end, slot, has_g; };
struct HmGramWeights {                                     // what one Gram launch accumulates (nW == 1: one weight per launch)
    int nW, wbase[2], wdim[2];                             // H^k: weight = MW[wbase] (1 omega | 3 omega^c) * (wdim >= 0 ? s (x - z_row)[wdim] : 1)
    int nV, vbase[HM_GRAM_MAXV], vdim[HM_GRAM_MAXV];       // g^v: weight = MW[vbase] (0 mu | 2 mu^c) * (vdim >= 0 ? s (x - z_row)[vdim] : 1)
};
int hm_tc_available();
size_t hm_tc_image_elems(int Mc, int Q);
int hm_tc_prepare(cudaStream_t s, const double* C, const HmConsts* consts, HmTcInfo* info, void* Cb, int M, int Mp, int Mc, int Q);
int hm_tc_proj_fwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, bool hyper,
                   int npass, int ncta);
int hm_tc_proj_bwd(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const void* Cb, const HmTcInfo* info, double* colpart,
                   int nslots, int npass);
int hm_tc_gram(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
               const int* seg_off, const HmGramWeights& gw, double* slots, int nctas, int f1, int f2, int npass);
// CTA-pair Gram (tc_gram2.cu): jobs are 256 x 256 blocks {I = row block, j0, nw = 256}; plan per pair; 2 slots per segment
#define HM_GRAM2_CHUNK 64
int hm_tc_gram2(cudaStream_t s, const HmTasks& tk, const HmProjArgs& a, const HmTcInfo* info, const HmGramSeg* segs,
                const int* seg_off, int nV, double* slots, int npairs, int f1, int f2, int npass);
int hm_tc_gram2_reduce(cudaStream_t s, const double* slots, const HmGramJob* jobs, const int2* jobslots, int njobs, int Q, int nV,
                       double* H, double* g0, int M, int Mp, int npass);
int hm_tc_gram_reduce(cudaStream_t s, const double* slots, const HmGramJob* jobs, const int2* jobslots, int njobs, int Q,
                      const HmGramWeights& gw, double* H, double* g0, int64_t gstride, int M, int Mp);
