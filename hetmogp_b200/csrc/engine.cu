// Engine: orchestration of one ELBO/gradient evaluation and the C-ABI of include/hetmogp_b200.h.
//
// One evaluation = SVMOGP.parameters_changed() (/root/reference/hetmogp/svmogp.py:85-166), i.e.
// SVMOGPInf.inference (svmogp_inf.py:23-109) + the hyper-parameter chain rule (svmogp.py:100-166,
// util.py:228-255), restated in the sufficient-statistics form of SURVEY.md App. B:
//
//   prepare   (fp64, M-sized)  K_uu, chol, K_uu^-1, S, alpha, C, S^-1, KL            [replicated on every rank]
//   forward   (N-sized)        a_tq, c_tq per row                                    proj_simt.cu / proj_tc.cu
//   lik       (N-sized)        W-mix, var_exp(+derivatives), row weights, scalars    lik_kernels.cu
//   backward  (N-sized)        g1, dz, dls column statistics; H1 Gram                proj_*.cu, gram_*.cu
//   reduce                     fp64 packed statistics buffer  <-- the ONE all-reduce of the multi-GPU path
//   finish    (fp64, M-sized)  dL_dmu, dL_dL, dL_dKmm, RBF / W / kappa / Z gradients
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"

// ------------------------------------------------------------------------------------------ error text
static thread_local char g_err[512] = "";
long long hm_launch_counter = 0;
void hm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#include <stdlib.h>

// ------------------------------------------------------------------------------------------ small kernels
namespace {

__global__ void prep_consts_kernel(HmConsts* c, const double* var, const double* ls, const double* W, const double* kappa,
                                   const double* Wc, const double* kc, const double* bscale, int Q, int J, int T) {
    const int tid = threadIdx.x;
    if (tid < Q) {
        c->var[tid] = var[tid];
        c->ls[tid] = ls[tid];
        c->inv_l2[tid] = 1.0 / (ls[tid] * ls[tid]);
    }
    for (int e = tid; e < J * Q; e += blockDim.x) {
        const int d = e / Q, q = e % Q;
        c->W[d][q] = W[e];
        c->kappa[d][q] = kappa[e];
        c->Wc[d][q] = Wc ? Wc[e] : W[e];
        c->kc[d][q] = kc ? kc[e] : kappa[e];
    }
    if (tid < T) c->bscale[tid] = bscale ? bscale[tid] : 1.0;
    __syncthreads();
    if (tid < J) {
        double s = 0.0;
        for (int q = 0; q < Q; ++q) s += (W[tid * Q + q] * W[tid * Q + q] + kappa[tid * Q + q]) * var[q];
        c->kdiag[tid] = s;
    }
}

// Z [M, Q*Xd] -> Zp [Q][Mp][Xd];  m_u [M,Q] -> mp [Q][Mp];  L_u packed [P,Q] -> Lu [Q][Mp][Mp] (identity padding)
__global__ void pad_inputs_kernel(const double* Z, const double* m_u, const double* L_u, double* Zp, double* mp,
                                  double* Lu, int M, int Mp, int Q, int Xd) {
    const int q = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Mp) return;
    double v;
    if (i < M && j < M) v = (j <= i) ? L_u[((int64_t)i * (i + 1) / 2 + j) * Q + q] : 0.0;
    else v = (i == j) ? 1.0 : 0.0;
    Lu[((int64_t)q * Mp + i) * Mp + j] = v;
    if (j == 0) {
        mp[(int64_t)q * Mp + i] = (i < M) ? m_u[(int64_t)i * Q + q] : 0.0;
        for (int k = 0; k < Xd; ++k) Zp[((int64_t)q * Mp + i) * Xd + k] = (i < M) ? Z[(int64_t)i * Q * Xd + q * Xd + k] : 0.0;
    }
}

// y[q][i] = sum_j A[q][i][j] x[q][j]   (one warp per row)
__global__ void dgemv_kernel(const double* __restrict__ A, const double* __restrict__ x, double* y, int Mp) {
    const int q = blockIdx.y;
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (row >= Mp) return;
    const double* a = A + ((int64_t)q * Mp + row) * Mp;
    const double* xv = x + (int64_t)q * Mp;
    double s = 0.0;
    for (int j = lane; j < Mp; j += 32) s += a[j] * xv[j];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[(int64_t)q * Mp + row] = s;
}

// C = KSK - Ki (fp64) and its fp32 copy
__global__ void make_c_kernel(const double* KSK, const double* Ki, double* C, float* Cf, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = KSK[i] - Ki[i];
    C[i] = v;
    if (Cf) Cf[i] = (float)v;
}

// KL_q = 0.5 sum(Ki o S) + 0.5 m.alpha - 0.5 M + sum log|diag Luu| - sum log|diag Lu|   (svmogp_inf.py:243-250)
// also flags inf in S^-1 (svmogp_inf.py:126).  grid = (row blocks, Q): per-block partials in KLpart[q][block]
// (summed in a fixed order by kl_finish_kernel -> deterministic).
__global__ void kl_kernel(const double* Ki, const double* S, const double* Sinv, const double* mp, const double* alpha,
                          const double* Luu, const double* Lu, double* KLpart, int* lu_singular, int M, int Mp) {
    const int q = blockIdx.y;
    const int64_t base = (int64_t)q * Mp * Mp;
    double s = 0.0;
    int bad = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = blockIdx.x * nwarp + warp; i < M; i += gridDim.x * nwarp) {
        const int64_t ro = base + (int64_t)i * Mp;
        for (int j = lane; j < M; j += 32) {
            s += 0.5 * Ki[ro + j] * S[ro + j];
            if (!isfinite(Sinv[ro + j])) bad = 1;   // inf (svmogp_inf.py:126) or the NaN an inf turns into downstream
        }
        if (lane == 0) {
            s += 0.5 * mp[(int64_t)q * Mp + i] * alpha[(int64_t)q * Mp + i];
            s += log(fabs(Luu[ro + i])) - log(fabs(Lu[ro + i]));
        }
    }
    __shared__ double sh[32];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) sh[warp] = s;
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&lu_singular[q], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < nwarp; ++w) tot += sh[w];
        KLpart[q * gridDim.x + blockIdx.x] = tot;
    }
}
__global__ void kl_finish_kernel(const double* KLpart, int nblk, double* KLq, int M) {
    const int q = threadIdx.x;
    double tot = 0.0;
    for (int b = 0; b < nblk; ++b) tot += KLpart[q * nblk + b];
    KLq[q] = tot - 0.5 * M;
}

// ---- statistic reducers (deterministic: fixed summation order)
// one block per statistic: strided partial sums, then a fixed-order tree (deterministic)
__global__ void reduce_lik_kernel(const double* partials, int nblocks, int nstat, double* stats, int t, int T, int J, int Q,
                                  int foff, int dimf, int off_sdv, int off_sma, int off_svc, int off_dls) {
    const int i = blockIdx.x;
    __shared__ double sh[128];
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 128) s += partials[(int64_t)b * nstat + i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    s = sh[0];
    const int nbase = 2 + dimf * (1 + 2 * Q);
    if (i == 0) stats[t] = s;
    else if (i == 1) stats[T + t] = s;
    else if (i >= nbase) stats[off_dls + (i - nbase)] += s;   // tensor-core path: K_mn lengthscale statistic (tasks run in stream order)
    else {
        const int k = i - 2, f = k / (1 + 2 * Q), r = k % (1 + 2 * Q), d = foff + f;
        if (r == 0) stats[off_sdv + d] = s;
        else if (r <= Q) stats[off_sma + d * Q + (r - 1)] = s;
        else stats[off_svc + d * Q + (r - 1 - Q)] = s;
    }
}

__global__ void reduce_col_kernel(const double* colpart, int nworkers, int ncol, int Mc, int Mp, int Xd, double* g1,
                                  double* dz, double* dls) {
    const int q = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncol) return;
    double s = 0.0;
    for (int w = 0; w < nworkers; ++w) s += colpart[((int64_t)q * nworkers + w) * ncol + i];
    if (i == ncol - 1) { if (dls) dls[q] = s; }
    else {
        const int k = i / Mc, m = i % Mc;
        if (k == 0) g1[(int64_t)q * Mp + m] = s;
        else dz[((int64_t)q * Xd + (k - 1)) * Mp + m] = s;
    }
}

// H[q][i][j] (ld Mp) = sum_split Hpart[split][q][max-tile-order(i,j)]  (lower tiles computed; mirror the rest)
__global__ void reduce_gram_kernel(const double* Hpart, int nsplit, int Q, int Mc, int Mp, int BT, double* H) {
    const int q = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Mp) return;
    double s = 0.0;
    if (i < Mc && j < Mc) {
        int a = i, b = j;
        if (a / BT < b / BT) { a = j; b = i; }
        for (int sp = 0; sp < nsplit; ++sp) s += Hpart[(((int64_t)sp * Q + q) * Mc + a) * Mc + b];
    }
    H[((int64_t)q * Mp + i) * Mp + j] = s;
}

// dL_dS = E - 0.5 (Ki - Sinv)
__global__ void dlds_kernel(const double* E, const double* Ki, const double* Sinv, double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = E[i] - 0.5 * (Ki[i] - Sinv[i]);
}

// dL_dK = sym(E - tmp - tmp^T - kg alpha^T) - (0.5 Ki - 0.5 KSK - 0.5 alpha alpha^T)   (svmogp_inf.py:132-133,151-154,166,170)
__global__ void dldk_kernel(const double* E, const double* tmp, const double* Ki, const double* KSK, const double* kg,
                            const double* alpha, double* out, int Mp) {
    const int q = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Mp) return;
    const int64_t b = (int64_t)q * Mp * Mp, ij = b + (int64_t)i * Mp + j, ji = b + (int64_t)j * Mp + i;
    const double* al = alpha + (int64_t)q * Mp;
    const double* g = kg + (int64_t)q * Mp;
    const double vij = E[ij] - tmp[ij] - tmp[ji] - g[i] * al[j];
    const double vji = E[ji] - tmp[ji] - tmp[ij] - g[j] * al[i];
    const double dve = 0.5 * (vij + vji);
    const double dkl = 0.5 * Ki[ij] - 0.5 * KSK[ij] - 0.5 * al[i] * al[j];
    out[ij] = dve - dkl;
}

// Per (q, row m): sums over j of KG = Kuu o dL_dK (K_uu recomputed, jitter-free as GPy's update_gradients_full does):
//   rowstat[q][m][0] = sum_j KG, [1] = sum_j KG r^2, dzmm[q][i][m] = -2 sum_j KG (z_mi - z_ji) / l^2   (svmogp.py:116,154)
__global__ void kmm_grad_kernel(const double* dLdK, const double* Zp, const HmConsts* c, double* rowstat, double* dzmm, int M,
                                int Mp, int Xd) {
    const int q = blockIdx.y;
    const int m = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (m >= M) return;
    const double var = c->var[q], il2 = c->inv_l2[q];
    const double* zm = Zp + ((int64_t)q * Mp + m) * Xd;
    double s0 = 0.0, s1 = 0.0, dz[HM_MAXXD] = {0.0, 0.0, 0.0, 0.0};
    for (int j = lane; j < M; j += 32) {
        const double* zj = Zp + ((int64_t)q * Mp + j) * Xd;
        double r2 = 0.0, d[HM_MAXXD];
        for (int k = 0; k < Xd; ++k) { d[k] = zm[k] - zj[k]; r2 += d[k] * d[k]; }
        r2 *= il2;
        const double kg = ((j == m) ? var : var * exp(-0.5 * r2)) * dLdK[((int64_t)q * Mp + m) * Mp + j];
        s0 += kg;
        s1 += kg * r2;
        for (int k = 0; k < Xd; ++k) dz[k] += kg * d[k];
    }
    for (int off = 16; off > 0; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        for (int k = 0; k < Xd; ++k) dz[k] += __shfl_xor_sync(0xffffffffu, dz[k], off);
    }
    if (lane == 0) {
        rowstat[((int64_t)q * Mp + m) * 2 + 0] = s0;
        rowstat[((int64_t)q * Mp + m) * 2 + 1] = s1;
        for (int k = 0; k < Xd; ++k) dzmm[((int64_t)q * Xd + k) * Mp + m] = -2.0 * dz[k] * il2;
    }
}

// Per latent q, fixed-order sums over the real M x M block (one CTA per q):
//   tr[q][0] = sum_ij H_ij Ki_ij = tr(H K_uu^-1),  [1] = tr(S K_uu^-1),  [2] = g1 . alpha,  [3] = m . alpha
// These give sum_ij K_uu o dL/dK_mm -- the K_mm part of the RBF variance gradient (svmogp.py:116) -- WITHOUT going through
// E = K_uu^-1 H K_uu^-1: with K_uu K_uu^-1 = I,
//   sum K_uu o dL/dK = -tr(H K^-1) - 2 tr(H C) - g1.alpha - M/2 + tr(S K^-1)/2 + m.alpha/2,   tr(H C) = sum_n omega_n c_n,
// where the last is a sum of per-row quantities the forward pass already produced.  H enters once against K_uu^-1 instead
// of twice: its error is amplified by cond(K_uu) instead of cond^1.4+ (tensor-core mode, cfg3: 3.5e-2 -> see DESIGN.md).
constexpr int kTraceSlices = 32;
__global__ void kmm_trace_kernel(const double* __restrict__ H, const double* __restrict__ Ki, const double* __restrict__ S,
                                 const double* __restrict__ g1, const double* __restrict__ alpha, const double* __restrict__ mp,
                                 double* tr, int M, int Mp) {
    const int q = blockIdx.y, sl = blockIdx.x, tid = threadIdx.x;    // slice sl: rows sl, sl + 32, ... (partials, summed in
    const int64_t b = (int64_t)q * Mp * Mp;                           // slice order by assemble_scalar_kernel)
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = sl; i < M; i += kTraceSlices)
        for (int j = tid; j < M; j += blockDim.x) {
            const int64_t ij = b + (int64_t)i * Mp + j;
            const double ki = Ki[ij];
            t[0] += H[ij] * ki;
            t[1] += S[ij] * ki;
        }
    if (sl == 0)
        for (int i = tid; i < M; i += blockDim.x) {
            const double al = alpha[(int64_t)q * Mp + i];
            t[2] += g1[(int64_t)q * Mp + i] * al;
            t[3] += mp[(int64_t)q * Mp + i] * al;
        }
    __shared__ double sh[4][256];
    for (int k = 0; k < 4; ++k) sh[k][tid] = t[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) for (int k = 0; k < 4; ++k) sh[k][tid] += sh[k][tid + w];
        __syncthreads();
    }
    if (tid < 4) tr[(q * kTraceSlices + sl) * 4 + tid] = sh[tid][0];
}

struct AssembleArgs {
    int M, Mp, Q, J, T, Xd, what;
    const double* stats;
    int off_sdv, off_sma, off_svc, off_dls, off_g1, off_dz;
    const double *KLq, *kg, *alpha, *dLdLfull, *dLdK, *rowstat, *dzmm;
    const double *tr, *jitter;   // kmm_trace_kernel sums (nullptr: not used) and the jitter K_uu^q was factored with
    const HmConsts* c;
    // outputs (device staging, reference layouts)
    double *log_marginal, *VE, *KL, *dmu, *dL, *dKmm, *drbf, *dW, *dkappa, *dZ;
};

// scalars + hyper-parameter gradients (single block)
__global__ void assemble_scalar_kernel(AssembleArgs a) {
    const int tid = threadIdx.x;
    const HmConsts* c = a.c;
    if (tid == 0) {
        double ve = 0.0, kl = 0.0;
        for (int t = 0; t < a.T; ++t) { a.VE[t] = a.stats[t]; ve += a.stats[t]; }
        for (int q = 0; q < a.Q; ++q) kl += a.KLq[q];
        a.KL[0] = kl;
        a.log_marginal[0] = ve - kl;  // svmogp_inf.py:84-88
    }
    if (a.what < HMOGP_WHAT_FULL) return;
    const double* sdv = a.stats + a.off_sdv;
    const double* sma = a.stats + a.off_sma;
    const double* svc = a.stats + a.off_svc;
    for (int e = tid; e < a.J * a.Q; e += blockDim.x) {
        const int d = e / a.Q, q = e % a.Q;
        // util.update_gradients_diag + update_gradients_Kmn (util.py:228-231,248-254; quirk C-4)
        a.dW[e] = c->W[d][q] * sdv[d] + sma[e] + 2.0 * c->W[d][q] * svc[e];
        a.dkappa[e] = sdv[d];
    }
    if (tid < 32 * a.Q) {   // one warp per latent (Q <= 8 = blockDim / 32): lane-strided sums + a fixed shuffle tree
        const int q = tid >> 5, lane = tid & 31;
        double s0 = 0.0, s1 = 0.0;
        for (int m = lane; m < a.M; m += 32) {
            s0 += a.rowstat[((int64_t)q * a.Mp + m) * 2 + 0];
            s1 += a.rowstat[((int64_t)q * a.Mp + m) * 2 + 1];
        }
        for (int w = 16; w > 0; w >>= 1) {
            s0 += __shfl_down_sync(0xffffffffu, s0, w);
            s1 += __shfl_down_sync(0xffffffffu, s1, w);
        }
        if (lane != 0) return;
        if (a.tr && a.jitter[q] == 0.0) {
            // sum K_uu o dL/dK_mm through traces (kmm_trace_kernel); with jitter K_uu K_uu^-1 != I and the direct sum stays
            double hc = 0.0;
            for (int d = 0; d < a.J; ++d) hc += c->W[d][q] * c->W[d][q] * svc[d * a.Q + q];   // tr(H C) = sum_n omega_n c_n
            double t[4] = {0.0, 0.0, 0.0, 0.0};
            for (int sl = 0; sl < kTraceSlices; ++sl)
                for (int k = 0; k < 4; ++k) t[k] += a.tr[(q * kTraceSlices + sl) * 4 + k];
            s0 = -t[0] - 2.0 * hc - t[2] - 0.5 * a.M + 0.5 * t[1] + 0.5 * t[3];
        }
        double dvar = s0 / c->var[q];   // update_gradients_full(dL_dKmm, Z_q)   svmogp.py:116
        double dls = s1 / c->ls[q];
        double kmn = 0.0, kd = 0.0;
        for (int d = 0; d < a.J; ++d) {
            kmn += c->Wc[d][q] * (sma[d * a.Q + q] + 2.0 * c->W[d][q] * svc[d * a.Q + q]);  // svmogp.py:140-141
            kd += (c->Wc[d][q] * c->Wc[d][q] + c->kc[d][q]) * sdv[d];                       // svmogp.py:142-143
        }
        dvar += kmn / c->var[q] + kd;
        dls += a.stats[a.off_dls + q] * c->inv_l2[q] / c->ls[q];
        a.drbf[q * 2 + 0] = dvar;
        a.drbf[q * 2 + 1] = dls;
    }
}

// M-sized outputs in the reference's layouts
__global__ void assemble_mat_kernel(AssembleArgs a) {
    const int q = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M || j >= a.M) return;
    const int M = a.M, Mp = a.Mp, Q = a.Q;
    if (j <= i && a.dL) a.dL[((int64_t)i * (i + 1) / 2 + j) * Q + q] = 2.0 * a.dLdLfull[((int64_t)q * Mp + i) * Mp + j];
    if (a.dKmm) a.dKmm[((int64_t)q * M + i) * M + j] = a.dLdK[((int64_t)q * Mp + i) * Mp + j];
    if (j == 0) {
        a.dmu[(int64_t)i * Q + q] = a.kg[(int64_t)q * Mp + i] - a.alpha[(int64_t)q * Mp + i];  // svmogp_inf.py:168
        if (a.what >= HMOGP_WHAT_FULL) {
            const double* dzs = a.stats + a.off_dz;
            for (int k = 0; k < a.Xd; ++k)
                a.dZ[(int64_t)i * Q * a.Xd + q * a.Xd + k] =
                    a.dzmm[((int64_t)q * a.Xd + k) * Mp + i] + dzs[((int64_t)q * a.Xd + k) * Mp + i] * a.c->inv_l2[q];
        }
    }
}

__global__ void flat_to_triang_kernel(const double* flat, double* dense, int M, int D) {
    const int d = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    dense[((int64_t)d * M + i) * M + j] = (j <= i) ? flat[((int64_t)i * (i + 1) / 2 + j) * D + d] : 0.0;
}
__global__ void triang_to_flat_kernel(const double* dense, double* flat, int M, int D) {
    const int d = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > i || j >= M) return;
    flat[((int64_t)i * (i + 1) / 2 + j) * D + d] = dense[((int64_t)d * M + i) * M + j];
}

// dense dL_dKmn for small N (tests): out[m][n] = alpha[m] dm[n] + 2 W dv[n] sum_m' C[m][m'] K[n][m']
__global__ void dense_dkmn_kernel(const double* X, int64_t N, const double* Zp, const double* Cq, const double* alpha,
                                  const HmConsts* c, int q, int d, const double* dm, const double* dv, int F, int f,
                                  double* out, int M, int Mp, int Xd) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (n >= N) return;
    double acc = 0.0;
    for (int mm = 0; mm < M; ++mm) {
        double r2 = 0.0;
        for (int k = 0; k < Xd; ++k) { const double dd = X[n * Xd + k] - Zp[((int64_t)q * Mp + mm) * Xd + k]; r2 += dd * dd; }
        acc += Cq[(int64_t)m * Mp + mm] * c->var[q] * exp(-0.5 * r2 * c->inv_l2[q]);
    }
    out[(int64_t)m * N + n] = alpha[m] * dm[n * F + f] + 2.0 * c->W[d][q] * dv[n * F + f] * acc;
}

__global__ void extract_mm_kernel(const double* src, double* dst, int M, int Mp, int lower_only) {
    const int q = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || j >= M) return;
    dst[((int64_t)q * M + i) * M + j] = (lower_only && j > i) ? 0.0 : src[((int64_t)q * Mp + i) * Mp + j];
}

// W-mix only (svmogp_inf.py:216-218 via SURVEY App. B): m_fd = sum_q W_dq a_q, v_fd = kdiag_d + sum_q W_dq^2 c_q for the F
// output functions of one task, from the per-row projections (SoA: array k at k * cap).  Prediction path: no likelihood.
template <typename T>
__global__ void mix_rows_kernel(const void* AC, int64_t cap, int64_t n, int Q, int foff, int F, const HmConsts* cs,
                                double* m_out, double* v_out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const T* ac = reinterpret_cast<const T*>(AC) + row;
    for (int f = 0; f < F; ++f) {
        const int d = foff + f;
        double mm = 0.0, vv = cs->kdiag[d];
        for (int q = 0; q < Q; ++q) {
            const double w = cs->W[d][q];
            mm += w * (double)ac[(size_t)q * cap];
            vv += w * w * (double)ac[(size_t)(Q + q) * cap];
        }
        m_out[row * F + f] = mm;
        v_out[row * F + f] = vv;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------ engine state
struct hmogp_engine {
    int M, Q, Xd, T, J, P, Mp, Mc, prec, device;
    hmogp_lik_desc liks[HM_MAXT];
    HmTasks tk;
    int64_t N[HM_MAXT];
    double *Xd_[HM_MAXT], *Yd_[HM_MAXT];
    int64_t cap[HM_MAXT];     // allocated rows for AC/MW
    cudaStream_t stream;
    bool own_stream;
    // parameter staging (device)
    double *pZ, *pm, *pL, *pvar, *pls, *pW, *pkappa, *pWc, *pkc, *pbs;
    HmConsts* consts;
    // M-sized buffers [Q][Mp][Mp] unless noted
    double *Zp, *mp, *alpha, *kg;  // [Q][Mp][Xd], [Q][Mp]
    double *Kuu, *Luu, *LuuInv, *Ki, *Lu, *LuInv, *Sinv, *S, *SK, *KSK, *C, *tmp, *T1, *E, *tmpE, *dLdS, *dLdLfull, *dLdK;
    float* Cf;
    // tensor-core path
    void* Cb;              // split-fp16 SW128 operand image of C
    HmTcInfo* tcinfo;
    int tc_npass, tc_f1, tc_f2;   // MMA passes; level-1 window (chunks); level-3 period (windows)
    int gram2_cost_diag;          // plan cost of a diagonal block job relative to 100 for an off-diagonal one
    bool gram2;                   // Gram on CTA pairs (tc_gram2.cu): padded M a multiple of 256
    int gram_chunk;               // data rows per plan chunk
    int tc_ncta;                  // forward kernel: 1 = one CTA per SM, 2 = cta_group::2 CTA pairs
    std::vector<HmGramJob> jobs_h;
    HmGramJob* jobs_d;
    HmGramSeg* segs_d; int* segoff_d; int2* jobslots_d;
    int max_segs, nslots_max;
    double* slots;         // [nslots][HM_GRAM_SLOT_DOUBLES] fp64 partial tiles
    double* gvec;          // [HM_GRAM_MAXV][Q][Mp]
    bool plan_dirty;
    double *KLq, *KLpart, *jitter_d, *rowstat, *dzmm, *trq;
    cudaStream_t s2;          // side stream of the prepare phase (S, S^-1 branch)
    cudaEvent_t ev_fork, ev_S, ev_Sinv;
    cudaGraphExec_t chol_graph;   // the 2 Mp / 32 panel + update launches of the blocked Cholesky, captured once
    // The M-sized chain replayed as CUDA graphs (fixed engine-owned buffers, so each is captured once):
    //   prepA: padding, S / S^-1 branch (forked), K_uu build and Cholesky -> host checks the pivot flags (jitchol)
    //   prepB: K_uu^-1, alpha, S K^-1, K^-1 S K^-1, C, KL, tensor-core operand image
    //   fin[k]: the finish chain for one (statistics buffer, what, dL_dKmm wanted) combination
    cudaGraphExec_t prepA_graph, prepB_graph, prepR_graph;   // R: the chain with K_uu's factorisation reused
    long long prepR_launches;
    // K_uu, its Cholesky factor, the inverses: valid for the (Z, sigma^2, l) in kuu_key_h (host callers) or, for device
    // callers, on the caller's word (hmogp_hint_hyper_unchanged)
    bool kuu_valid, kuu_key_ok, hint_unchanged, kuu_cache_off;
    bool trace_off;   // HMOGP_NO_KMM_TRACE=1: the RBF variance gradient sums K_uu o dL/dK_mm directly (diagnostic)
    std::vector<double> kuu_key_h;
    long long kuu_reused;
    long long prepA_launches, prepB_launches;
    struct FinGraph { cudaGraphExec_t exec; const double* stats; int what; bool dkmm; long long launches; };
    std::vector<FinGraph> fin_graphs;
    bool graphs_off;
    int prepare_calls;
    cudaStream_t sc;          // copy stream: host -> device data uploads overlap the M-sized prepare phase of the next step
    cudaEvent_t ev_cfence, ev_data, ev_dataY;   // inputs X uploaded (the forward needs them) / labels Y too (the likelihoods do)
    bool dataY_pending;
    bool data_pending;        // an upload on sc has not been ordered before the compute stream yet
    const double* up_X[HM_MAXT];   // deferred uploads from pinned host memory (queued behind the next step's parameter copies,
    const double* up_Y[HM_MAXT];   // so that the M-sized prepare phase is not stuck behind them in the copy engine)
    int *flags_d;  // [2][HM_MAXQ]: chol_fail, lu_singular
    // statistics
    int64_t stats_len;
    int off_nneg, off_sdv, off_sma, off_svc, off_dls, off_g1, off_dz, off_H;
    double* stats;
    double* lik_part; int lik_max_blocks;
    double* colpart; int nworkers;
    double* Hpart; int nsplit;
    // output staging
    double *o_lm, *o_VE, *o_KL, *o_dmu, *o_dL, *o_dKmm, *o_drbf, *o_dW, *o_dkappa, *o_dZ;
    // status
    double jitter_h[HM_MAXQ];
    int chol_fail_h[HM_MAXQ];   // factorisation attempts of K_uu^q that failed in the last prepare (jitchol retries)
    bool has_chain;
    int last_what;
    // timing
    bool timing;
    cudaEvent_t ev[7];
    float ms[6];
    long long launches, launch0;
    std::vector<void*> allocs;
};

namespace {

template <typename U> int dalloc(hmogp_engine* e, U** p, size_t count) {
    void* q = nullptr;
    cudaError_t err = cudaMalloc(&q, count * sizeof(U) > 0 ? count * sizeof(U) : 16);
    if (err != cudaSuccess) {
        hm_set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(U), cudaGetErrorString(err));
        return HMOGP_ERR_CUDA;
    }
    e->allocs.push_back(q);
    *p = (U*)q;
    return 0;
}

int lik_dims(const hmogp_lik_desc& l, int* dy, int* df, int* dp) {
    switch (l.kind) {
        case HMOGP_LIK_GAUSSIAN: case HMOGP_LIK_BERNOULLI: case HMOGP_LIK_POISSON: case HMOGP_LIK_EXPONENTIAL:
            *dy = 1; *df = 1; *dp = 1; return 0;
        case HMOGP_LIK_HETGAUSSIAN: case HMOGP_LIK_GAMMA: case HMOGP_LIK_BETA:
            *dy = 1; *df = 2; *dp = 1; return 0;
        case HMOGP_LIK_CATEGORICAL:
            if (l.K < 2 || l.K - 1 > HM_MAXF) { hm_set_error("Categorical K=%d unsupported (2..%d)", l.K, HM_MAXF + 1); return HMOGP_ERR_ARG; }
            *dy = 1; *df = l.K - 1; *dp = l.K - 1; return 0;  // categorical.py:287-291
        default: hm_set_error("unknown likelihood kind %d", l.kind); return HMOGP_ERR_ARG;
    }
}

size_t esize(int prec) { return prec == HMOGP_PREC_FP64 ? 8 : 4; }

int copy_in(hmogp_engine* e, double* dst, const double* src, size_t n, int mem_kind) {
    if (!src || n == 0) return 0;
    HM_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), mem_kind == HMOGP_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, e->stream));
    return 0;
}
int copy_out(hmogp_engine* e, double* dst, const double* src, size_t n, int mem_kind) {
    if (!dst || n == 0) return 0;
    HM_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), mem_kind == HMOGP_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, e->stream));
    return 0;
}

HmProjArgs proj_args(hmogp_engine* e) {
    HmProjArgs a;
    a.M = e->M; a.Mp = e->Mp; a.Mc = e->Mc; a.Q = e->Q; a.Xdim = e->Xd;
    a.Zp = e->Zp; a.alpha = e->alpha;
    a.C = (e->prec == HMOGP_PREC_FP64) ? (const void*)e->C : (const void*)e->Cf;
    a.consts = e->consts; a.colpart = e->colpart; a.nworkers = e->nworkers;
    return a;
}

int flush_uploads(hmogp_engine* e);

// The blocked Cholesky is 2 Mp / 32 short dependent launches with fixed arguments: replay them as one graph (the gaps
// between dependent launches are a visible part of the M-sized serial chain).  HMOGP_NO_GRAPH=1 issues them directly.
int cholesky_graphed(hmogp_engine* e) {
    static const bool off = [] { const char* v = getenv("HMOGP_NO_GRAPH"); return v && atoi(v) != 0; }();
    cudaStream_t s = e->stream;
    if (off) return hm_cholesky(s, e->Luu, e->Mp, (int64_t)e->Mp * e->Mp, e->Q, e->flags_d);
    if (!e->chol_graph) {
        cudaStream_t cs;
        HM_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        HM_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        const int rc = hm_cholesky(cs, e->Luu, e->Mp, (int64_t)e->Mp * e->Mp, e->Q, e->flags_d);
        const cudaError_t ce = cudaStreamEndCapture(cs, &g);
        cudaStreamDestroy(cs);
        if (rc || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            hm_set_error("Cholesky graph capture failed");
            return rc ? rc : HMOGP_ERR_CUDA;
        }
        const cudaError_t ie = cudaGraphInstantiate(&e->chol_graph, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { e->chol_graph = nullptr; hm_set_error("Cholesky graph instantiation failed"); return HMOGP_ERR_CUDA; }
    }
    HM_CUDA(cudaGraphLaunch(e->chol_graph, s));
    hm_launch_counter += 2 * (e->Mp / 32) - 1;   // kernels replayed by the graph (panel + trailing update per 32 columns)
    return 0;
}
#ifdef HM_DEBUG_SKIP
static int dbg_skip() { static int v = -1; if (v < 0) { const char* e = getenv("HMOGP_DEBUG_SKIP"); v = e ? atoi(e) : 0; } return v; }
#define HM_SKIP(bit) (dbg_skip() & (bit))
#else
#define HM_SKIP(bit) 0
#endif

// ---- prepare: everything M-sized that precedes the data pass.  Returns HMOGP_ERR_LINALG if jitchol gives up.
// part A on streams (s, s2): padded inputs; S = Lu Lu^T, Lu^-1, S^-1 on the side stream; K_uu (with the jitter in
// e->jitter_d), its copy into Luu, cleared flags, blocked Cholesky.  `graphed_chol`: replay the Cholesky's own graph
// (direct issue; inside a capture the launches are recorded individually).
int prepare_partA(hmogp_engine* e, cudaStream_t s, cudaStream_t s2, bool graphed_chol, bool reuse = false) {
    const int M = e->M, Mp = e->Mp, Q = e->Q, Xd = e->Xd;
    const int64_t sQ = (int64_t)Mp * Mp;
    {
        dim3 grid((unsigned)hm_cdiv(Mp, 128), (unsigned)Mp, (unsigned)Q);
        pad_inputs_kernel<<<grid, 128, 0, s>>>(e->pZ, e->pm, e->pL, e->Zp, e->mp, e->Lu, M, Mp, Q, Xd);
        HM_CUDA(cudaGetLastError());
    }
    // ---- side stream: S = Lu Lu^T (svmogp_inf.py:193-195) and S^-1 (svmogp_inf.py:124) do not depend on K_uu
    HM_CUDA(cudaEventRecord(e->ev_fork, s));
    HM_CUDA(cudaStreamWaitEvent(s2, e->ev_fork, 0));
    HM_CHECK(hm_dgemm(s2, false, true, Mp, Mp, Mp, 1.0, e->Lu, Mp, sQ, e->Lu, Mp, sQ, 0.0, e->S, Mp, sQ, Q, 1, 0, 0, 0, HM_GEMM_MIRROR | HM_GEMM_K_LE));
    HM_CUDA(cudaEventRecord(e->ev_S, s2));
    if (!HM_SKIP(4)) HM_CHECK(hm_tri_inverse(s2, e->Lu, e->LuInv, e->T1, Mp, sQ, Q));
    HM_CHECK(hm_dgemm(s2, true, false, Mp, Mp, Mp, 1.0, e->LuInv, Mp, sQ, e->LuInv, Mp, sQ, 0.0, e->Sinv, Mp, sQ, Q, 1, 0, 0, 0, HM_GEMM_MIRROR | HM_GEMM_K_GE));
    HM_CUDA(cudaEventRecord(e->ev_Sinv, s2));
    // K_uu, Cholesky (util.py:197-198); `reuse`: Z, sigma^2, l are those of the resident factorisation (VE phases of VEM,
    // svmogp.py:104-113 with the hyper-parameters fixed: only m_u and L_u move)
    HM_CUDA(cudaMemsetAsync(e->flags_d, 0, sizeof(int) * 2 * HM_MAXQ, s));
    if (!reuse) {
        HM_CHECK(hm_build_kuu(s, e->Zp, e->consts, e->jitter_d, e->Kuu, M, Mp, Xd, Q));
        HM_CUDA(cudaMemcpyAsync(e->Luu, e->Kuu, sizeof(double) * sQ * Q, cudaMemcpyDeviceToDevice, s));
        if (!HM_SKIP(1)) {
            if (graphed_chol) HM_CHECK(cholesky_graphed(e));
            else HM_CHECK(hm_cholesky(s, e->Luu, Mp, sQ, Q, e->flags_d));
        }
    }
    HM_CUDA(cudaStreamWaitEvent(s, e->ev_Sinv, 0));   // join (a captured graph must not leave the side stream dangling)
    return 0;
}

// part B on stream s: K_uu^-1 = Luu^-T Luu^-1 (dpotri, util.py:199), alpha, S K^-1, K^-1 S K^-1, C, KL, operand image
int prepare_partB(hmogp_engine* e, cudaStream_t s, bool reuse = false) {
    const int M = e->M, Mp = e->Mp, Q = e->Q;
    const int64_t sQ = (int64_t)Mp * Mp;
    if (!reuse) {
        if (!HM_SKIP(2)) HM_CHECK(hm_tri_inverse(s, e->Luu, e->LuuInv, e->tmp, Mp, sQ, Q));
        HM_CHECK(hm_dgemm(s, true, false, Mp, Mp, Mp, 1.0, e->LuuInv, Mp, sQ, e->LuuInv, Mp, sQ, 0.0, e->Ki, Mp, sQ, Q, 1, 0, 0, 0, HM_GEMM_MIRROR | HM_GEMM_K_GE));
    }
    {
        dim3 grid((unsigned)hm_cdiv(Mp, 8), (unsigned)Q);
        dgemv_kernel<<<grid, 256, 0, s>>>(e->Ki, e->mp, e->alpha, Mp);
        HM_CUDA(cudaGetLastError());
    }
    if (!HM_SKIP(8)) HM_CHECK(hm_dgemm(s, false, false, Mp, Mp, Mp, 1.0, e->S, Mp, sQ, e->Ki, Mp, sQ, 0.0, e->SK, Mp, sQ, Q));
    if (!HM_SKIP(8)) HM_CHECK(hm_dgemm(s, false, false, Mp, Mp, Mp, 1.0, e->Ki, Mp, sQ, e->SK, Mp, sQ, 0.0, e->KSK, Mp, sQ, Q, 1, 0, 0, 0, HM_GEMM_MIRROR));
    {
        const int64_t n = sQ * Q;
        make_c_kernel<<<(unsigned)hm_cdiv(n, 256), 256, 0, s>>>(e->KSK, e->Ki, e->C, e->Cf, n);
        HM_CUDA(cudaGetLastError());
    }
    {
        const int nblk = 48;   // <= 64 (KLpart)
        dim3 grid((unsigned)nblk, (unsigned)Q);
        if (!HM_SKIP(16)) kl_kernel<<<grid, 256, 0, s>>>(e->Ki, e->S, e->Sinv, e->mp, e->alpha, e->Luu, e->Lu, e->KLpart, e->flags_d + HM_MAXQ, M, Mp);
        HM_CUDA(cudaGetLastError());
        kl_finish_kernel<<<1, Q, 0, s>>>(e->KLpart, nblk, e->KLq, M);
        HM_CUDA(cudaGetLastError());
    }
    if (e->prec == HMOGP_PREC_TC && !HM_SKIP(32)) HM_CHECK(hm_tc_prepare(s, e->C, e->consts, e->tcinfo, e->Cb, M, Mp, e->Mc, Q));
    return 0;
}

// Capture `body` (which issues on the capture stream cs and, forked from it, on e->s2) into an executable graph.
template <typename F> int capture_graph(hmogp_engine* e, cudaGraphExec_t* exec, long long* launches, F body) {
    cudaStream_t cs;
    HM_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    const long long l0 = hm_launch_counter;
    cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (ce == cudaSuccess) {
        rc = body(cs);
        ce = cudaStreamEndCapture(cs, &g);
    }
    cudaStreamDestroy(cs);
    *launches = hm_launch_counter - l0;
    if (rc || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (!rc) hm_set_error("CUDA graph capture of the M-sized chain failed: %s", cudaGetErrorString(ce));
        return rc ? rc : HMOGP_ERR_CUDA;
    }
    const cudaError_t ie = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) { *exec = nullptr; hm_set_error("CUDA graph instantiation failed: %s", cudaGetErrorString(ie)); return HMOGP_ERR_CUDA; }
    return 0;
}

int mm_prepare(hmogp_engine* e, const hmogp_params* p, int mem_kind) {
    cudaStream_t s = e->stream;
    const int M = e->M, Q = e->Q, Xd = e->Xd, J = e->J, T = e->T;
    HM_CHECK(copy_in(e, e->pZ, p->Z, (size_t)M * Q * Xd, mem_kind));
    HM_CHECK(copy_in(e, e->pm, p->m_u, (size_t)M * Q, mem_kind));
    HM_CHECK(copy_in(e, e->pL, p->L_u, (size_t)e->P * Q, mem_kind));
    HM_CHECK(copy_in(e, e->pvar, p->rbf_var, Q, mem_kind));
    HM_CHECK(copy_in(e, e->pls, p->rbf_ls, Q, mem_kind));
    HM_CHECK(copy_in(e, e->pW, p->W, (size_t)J * Q, mem_kind));
    HM_CHECK(copy_in(e, e->pkappa, p->kappa, (size_t)J * Q, mem_kind));
    HM_CHECK(copy_in(e, e->pWc, p->W_chain, (size_t)J * Q, mem_kind));
    HM_CHECK(copy_in(e, e->pkc, p->kappa_chain, (size_t)J * Q, mem_kind));
    HM_CHECK(copy_in(e, e->pbs, p->batch_scale, T, mem_kind));
    e->has_chain = p->W_chain != nullptr || p->kappa_chain != nullptr;
    HM_CHECK(flush_uploads(e));   // row uploads go behind the parameter copies in the copy engine
    prep_consts_kernel<<<1, 256, 0, s>>>(e->consts, e->pvar, e->pls, e->pW, e->pkappa, p->W_chain ? e->pWc : nullptr,
                                         p->kappa_chain ? e->pkc : nullptr, p->batch_scale ? e->pbs : nullptr, Q, J, T);
    HM_CUDA(cudaGetLastError());
    // The first evaluation of an engine issues the chain directly (it also sets the kernels' attributes); from the
    // second on the two halves are graph replays.  HMOGP_NO_GRAPH=1 keeps direct issue.
#ifdef HM_DEBUG_SKIP
    const bool graphs = false;
#else
    const bool graphs = !e->graphs_off && e->prepare_calls > 0;
#endif
    ++e->prepare_calls;
    for (int q = 0; q < Q; ++q) { e->jitter_h[q] = 0.0; e->chol_fail_h[q] = 0; }
    // ---- is the resident factorisation of K_uu the one these parameters need?  Host callers: compared bit by bit with
    // the values it was built from; device callers: only on their explicit word for this call.
    bool reuse = false;
    {
        const size_t nz = (size_t)M * Q * Xd, nk = nz + 2 * (size_t)Q;
        if (mem_kind == HMOGP_MEM_HOST) {
            if (e->kuu_key_h.size() != nk) { e->kuu_key_h.assign(nk, 0.0); e->kuu_key_ok = false; }
            const bool same = e->kuu_key_ok && memcmp(e->kuu_key_h.data(), p->Z, sizeof(double) * nz) == 0 &&
                              memcmp(e->kuu_key_h.data() + nz, p->rbf_var, sizeof(double) * Q) == 0 &&
                              memcmp(e->kuu_key_h.data() + nz + Q, p->rbf_ls, sizeof(double) * Q) == 0;
            reuse = same && e->kuu_valid;
            if (!same) {
                memcpy(e->kuu_key_h.data(), p->Z, sizeof(double) * nz);
                memcpy(e->kuu_key_h.data() + nz, p->rbf_var, sizeof(double) * Q);
                memcpy(e->kuu_key_h.data() + nz + Q, p->rbf_ls, sizeof(double) * Q);
                e->kuu_key_ok = true;
            }
        } else {
            reuse = e->kuu_valid && e->hint_unchanged;
            if (!reuse) e->kuu_key_ok = false;      // the key no longer describes what is resident
        }
        e->hint_unchanged = false;
#ifdef HM_DEBUG_SKIP
        reuse = false;
#endif
        if (e->kuu_cache_off) reuse = false;
    }
    if (reuse) {
        ++e->kuu_reused;
        if (graphs) {
            if (!e->prepR_graph) HM_CHECK(capture_graph(e, &e->prepR_graph, &e->prepR_launches, [&](cudaStream_t cs) {
                const int rc = prepare_partA(e, cs, e->s2, false, true);
                return rc ? rc : prepare_partB(e, cs, true);
            }));
            HM_CUDA(cudaGraphLaunch(e->prepR_graph, s));
            hm_launch_counter += e->prepR_launches - 1;
        } else {
            HM_CHECK(prepare_partA(e, s, e->s2, false, true));
            HM_CHECK(prepare_partB(e, s, true));
        }
        return 0;
    }
    e->kuu_valid = false;
    HM_CUDA(cudaMemsetAsync(e->jitter_d, 0, sizeof(double) * HM_MAXQ, s));
    if (graphs) {
        if (!e->prepA_graph) HM_CHECK(capture_graph(e, &e->prepA_graph, &e->prepA_launches, [&](cudaStream_t cs) { return prepare_partA(e, cs, e->s2, false); }));
        HM_CUDA(cudaGraphLaunch(e->prepA_graph, s));
        hm_launch_counter += e->prepA_launches - 1;
    } else HM_CHECK(prepare_partA(e, s, e->s2, true));
    // jitchol (util.py:198): no jitter unless the plain factorisation fails; then var*1e-6 * 10^k, k<5
    double var_h[HM_MAXQ];
    bool have_var = false;
    for (int attempt = 0;; ++attempt) {
        int fl[HM_MAXQ];
        HM_CUDA(cudaMemcpyAsync(fl, e->flags_d, sizeof(int) * Q, cudaMemcpyDeviceToHost, s));
        if (!HM_SKIP(64)) HM_CUDA(cudaStreamSynchronize(s));
        bool any = false;
        if (HM_SKIP(64)) for (int q = 0; q < Q; ++q) fl[q] = 0;
        for (int q = 0; q < Q; ++q) { any = any || fl[q]; if (fl[q]) ++e->chol_fail_h[q]; }
        if (!any) break;
        if (attempt >= 5) {
            cudaStreamSynchronize(e->s2);
            hm_set_error("not positive definite, even with jitter.");
            return HMOGP_ERR_LINALG;
        }
        if (!have_var) {
            HM_CUDA(cudaMemcpy(var_h, e->pvar, sizeof(double) * Q, cudaMemcpyDeviceToHost));
            have_var = true;
        }
        for (int q = 0; q < Q; ++q)
            if (fl[q]) e->jitter_h[q] = (e->jitter_h[q] == 0.0) ? var_h[q] * 1e-6 : e->jitter_h[q] * 10.0;
        // retry (rare): K_uu with the jitter, copy, cleared flags, Cholesky -- issued directly
        const int64_t sQ = (int64_t)e->Mp * e->Mp;
        HM_CUDA(cudaMemcpyAsync(e->jitter_d, e->jitter_h, sizeof(double) * Q, cudaMemcpyHostToDevice, s));
        HM_CHECK(hm_build_kuu(s, e->Zp, e->consts, e->jitter_d, e->Kuu, M, e->Mp, Xd, Q));
        HM_CUDA(cudaMemcpyAsync(e->Luu, e->Kuu, sizeof(double) * sQ * Q, cudaMemcpyDeviceToDevice, s));
        HM_CUDA(cudaMemsetAsync(e->flags_d, 0, sizeof(int) * 2 * HM_MAXQ, s));
        HM_CHECK(cholesky_graphed(e));
    }
    if (graphs) {
        if (!e->prepB_graph) HM_CHECK(capture_graph(e, &e->prepB_graph, &e->prepB_launches, [&](cudaStream_t cs) { return prepare_partB(e, cs); }));
        HM_CUDA(cudaGraphLaunch(e->prepB_graph, s));
        hm_launch_counter += e->prepB_launches - 1;
    } else HM_CHECK(prepare_partB(e, s));
    // a factorisation that needed jitter is not kept: the retry ladder is part of the call's reported status
    e->kuu_valid = true;
    for (int q = 0; q < Q; ++q) if (e->jitter_h[q] != 0.0) e->kuu_valid = false;
    return 0;
}

int refresh_tasks(hmogp_engine* e) {
    for (int t = 0; t < e->T; ++t) {
        if (!e->Xd_[t]) { hm_set_error("task %d has no data (call hmogp_set_data)", t); return HMOGP_ERR_ARG; }
        e->tk.X[t] = e->Xd_[t];
        e->tk.Y[t] = e->Yd_[t];
    }
    return 0;
}

// queue the deferred uploads on the copy stream, after everything already queued on the compute stream (which may still
// read the old rows)
int flush_uploads(hmogp_engine* e) {
    bool any = false;
    for (int t = 0; t < e->T; ++t) any = any || e->up_X[t];
    if (!any) return 0;
    HM_CUDA(cudaEventRecord(e->ev_cfence, e->stream));
    HM_CUDA(cudaStreamWaitEvent(e->sc, e->ev_cfence, 0));
    // inputs first: the forward pass only needs X, the labels follow while it runs
    for (int t = 0; t < e->T; ++t)
        if (e->up_X[t]) HM_CUDA(cudaMemcpyAsync(e->Xd_[t], e->up_X[t], (size_t)e->N[t] * e->Xd * sizeof(double), cudaMemcpyHostToDevice, e->sc));
    HM_CUDA(cudaEventRecord(e->ev_data, e->sc));
    for (int t = 0; t < e->T; ++t) {
        if (!e->up_X[t]) continue;
        HM_CUDA(cudaMemcpyAsync(e->Yd_[t], e->up_Y[t], (size_t)e->N[t] * sizeof(double), cudaMemcpyHostToDevice, e->sc));
        e->up_X[t] = e->up_Y[t] = nullptr;
    }
    HM_CUDA(cudaEventRecord(e->ev_dataY, e->sc));
    e->dataY_pending = true;
    e->data_pending = true;
    return 0;
}

// order pending row uploads (copy stream) before whatever is queued next on the compute stream: the inputs X, and with
// labels = true also Y
int data_ready(hmogp_engine* e, bool labels = true) {
    HM_CHECK(flush_uploads(e));
    if (e->data_pending) {
        HM_CUDA(cudaStreamWaitEvent(e->stream, e->ev_data, 0));
        e->data_pending = false;
    }
    if (labels && e->dataY_pending) {
        HM_CUDA(cudaStreamWaitEvent(e->stream, e->ev_dataY, 0));
        e->dataY_pending = false;
    }
    return 0;
}

// ---- tensor-core backward: plan, launches, reduction
// The work of all (latent q, output tile) pairs is laid on one tape (position = cost-weighted chunk index) and cut into
// nworkers equal pieces; a piece that crosses a pair boundary becomes several segments, each with a private fp64 partial
// tile ("slot").  Slots of one pair are contiguous, so the reduction order is fixed.
int build_gram_plan(hmogp_engine* e) {
    if (!e->plan_dirty) return 0;
    const int G = e->gram2 ? e->nworkers / 2 : e->nworkers, Q = e->Q, nj = (int)e->jobs_h.size();   // workers: CTAs or CTA pairs
    int64_t NC = 0;
    for (int t = 0; t < e->T; ++t) NC += hm_cdiv(e->tk.count[t], e->gram_chunk);
    std::vector<HmGramSeg> segs;
    std::vector<std::vector<HmGramSeg>> per(G);
    std::vector<int2> jobslots((size_t)Q * nj);
    std::vector<int64_t> cost(nj);
    int64_t per_q = 0;
    for (int j = 0; j < nj; ++j) {   // generated columns per chunk (B + A); pair Gram: one generation serves both on the diagonal
        cost[j] = e->gram2 ? (e->jobs_h[j].j0 == 256 * e->jobs_h[j].I ? e->gram2_cost_diag : 100) : e->jobs_h[j].nw + 128;
        per_q += cost[j];
    }
    const double total = (double)per_q * (double)NC * Q;
    int slot = 0;
    double bs = 0.0;   // tape position of the current block
    for (int q = 0; q < Q; ++q)
        for (int j = 0; j < nj; ++j) {
            const double len = (double)cost[j] * (double)NC;
            jobslots[(size_t)q * nj + j].x = slot;
            if (NC > 0) {
                int g0 = (int)((bs / total) * G), g1 = (int)(((bs + len) / total) * G);
                g0 = g0 > 0 ? g0 - 1 : 0;          // one CTA of slack either side: rounding of the tape positions
                g1 = g1 + 1;
                if (g0 > G - 1) g0 = G - 1;
                if (g1 > G - 1) g1 = G - 1;
                for (int g = g0; g <= g1; ++g) {
                    const double ps = total * g / G, pe = (g == G - 1) ? total * 2.0 : total * (g + 1) / G;
                    int64_t cb = ps <= bs ? 0 : (int64_t)((ps - bs) / (double)cost[j]);
                    int64_t ce = pe >= bs + len ? NC : (int64_t)((pe - bs) / (double)cost[j]);
                    if (cb > NC) cb = NC;
                    if (ce > NC) ce = NC;
                    if (ce <= cb) continue;
                    HmGramSeg sg;
                    sg.q = q; sg.I = e->jobs_h[j].I; sg.j0 = e->jobs_h[j].j0; sg.nw = e->jobs_h[j].nw;
                    sg.chunk_begin = (int)cb; sg.chunk_end = (int)ce; sg.slot = slot++; sg.has_g = (sg.j0 == 0) ? 1 : 0;
                    per[g].push_back(sg);
                }
            }
            jobslots[(size_t)q * nj + j].y = slot;
            bs += len;
        }
    std::vector<int> off(G + 1, 0);
    for (int g = 0; g < G; ++g) {
        off[g] = (int)segs.size();
        segs.insert(segs.end(), per[g].begin(), per[g].end());
    }
    off[G] = (int)segs.size();
    if ((int)segs.size() > e->max_segs || slot > e->nslots_max) {
        hm_set_error("gram plan overflow: %d segments, %d slots", (int)segs.size(), slot);
        return HMOGP_ERR_ARG;
    }
    cudaStream_t s = e->stream;
    HM_CUDA(cudaStreamSynchronize(s));
    if (!segs.empty()) HM_CUDA(cudaMemcpy(e->segs_d, segs.data(), sizeof(HmGramSeg) * segs.size(), cudaMemcpyHostToDevice));
    HM_CUDA(cudaMemcpy(e->segoff_d, off.data(), sizeof(int) * (G + 1), cudaMemcpyHostToDevice));
    HM_CUDA(cudaMemcpy(e->jobslots_d, jobslots.data(), sizeof(int2) * jobslots.size(), cudaMemcpyHostToDevice));
    e->plan_dirty = false;
    return 0;
}

int tc_backward(hmogp_engine* e, int what, double* stats) {
    cudaStream_t s = e->stream;
    const int Q = e->Q, Xd = e->Xd, Mp = e->Mp, M = e->M;
    const int64_t MM = (int64_t)Q * Mp * Mp, gstride = (int64_t)Q * Mp;
    HM_CHECK(build_gram_plan(e));
    HmProjArgs pa = proj_args(e);
    const bool full = what >= HMOGP_WHAT_FULL;
    if (full) {
        // ---- transposed projection: g1 = K^T mu and dz = sum_n GK (x - z) per inducing point (tc_bwd.cu)
        const int nslots = e->nworkers - (e->nworkers % 2);
        HM_CHECK(hm_tc_proj_bwd(s, e->tk, pa, e->Cb, e->tcinfo, e->colpart, nslots, e->tc_npass));
        const int ncol = (1 + Xd) * e->Mc + 1;
        dim3 g1((unsigned)hm_cdiv(ncol, 256), (unsigned)Q);
        reduce_col_kernel<<<g1, 256, 0, s>>>(e->colpart, nslots, ncol, e->Mc, Mp, Xd, stats + e->off_g1, stats + e->off_dz, nullptr);
        HM_CUDA(cudaGetLastError());
    }
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[4], s));
    // ---- H^1 = sum K^T diag(omega) K  (+ g^mu in a VE step; a full step took it from the transposed projection)
    HmGramWeights gw;
    memset(&gw, 0, sizeof(gw));
    gw.nW = 1; gw.wbase[0] = 1; gw.wdim[0] = -1; gw.wdim[1] = -1;
    if (!full) { gw.vbase[0] = 0; gw.vdim[0] = -1; gw.nV = 1; }
    HM_CUDA(cudaMemsetAsync(stats + e->off_H, 0, sizeof(double) * MM, s));
    if (e->gram2) {
        HM_CHECK(hm_tc_gram2(s, e->tk, pa, e->tcinfo, e->segs_d, e->segoff_d, gw.nV, e->slots, e->nworkers / 2, e->tc_f1, e->tc_f2, e->tc_npass));
        HM_CHECK(hm_tc_gram2_reduce(s, e->slots, e->jobs_d, e->jobslots_d, (int)e->jobs_h.size(), Q, gw.nV, stats + e->off_H, e->gvec, M, Mp, e->tc_npass));
    } else {
        HM_CHECK(hm_tc_gram(s, e->tk, pa, e->tcinfo, e->segs_d, e->segoff_d, gw, e->slots, e->nworkers, e->tc_f1, e->tc_f2, e->tc_npass));
        HM_CHECK(hm_tc_gram_reduce(s, e->slots, e->jobs_d, e->jobslots_d, (int)e->jobs_h.size(), Q, gw, stats + e->off_H, e->gvec, gstride, M, Mp));
    }
    if (!full) HM_CUDA(cudaMemcpyAsync(stats + e->off_g1, e->gvec, sizeof(double) * gstride, cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------ C-ABI
extern "C" {

int hmogp_abi_version(void) { return HMOGP_ABI_VERSION; }
const char* hmogp_last_error(void) { return g_err; }
int hmogp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int hmogp_lik_dims(const hmogp_lik_desc* lik, int32_t* dim_y, int32_t* dim_f, int32_t* dim_p) {
    if (!lik) { hm_set_error("null likelihood"); return HMOGP_ERR_ARG; }
    int a, b, c;
    HM_CHECK(lik_dims(*lik, &a, &b, &c));
    if (dim_y) *dim_y = a;
    if (dim_f) *dim_f = b;
    if (dim_p) *dim_p = c;
    return 0;
}

int hmogp_generate_metadata(int32_t T, const hmogp_lik_desc* liks, int64_t* task_index, int64_t* y_index,
                            int64_t* function_index, int64_t* d_index, int64_t* pred_index, int32_t* n_y, int32_t* n_f,
                            int32_t* n_p) {
    // het_likelihood.py:24-44
    int ny = 0, nf = 0, np_ = 0;
    for (int t = 0; t < T; ++t) {
        int dy, df, dp;
        HM_CHECK(lik_dims(liks[t], &dy, &df, &dp));
        if (task_index) task_index[t] = t;
        for (int i = 0; i < dy; ++i) { if (y_index) y_index[ny] = t; ++ny; }
        for (int i = 0; i < df; ++i) { if (function_index) function_index[nf] = t; if (d_index) d_index[nf] = i; ++nf; }
        for (int i = 0; i < dp; ++i) { if (pred_index) pred_index[np_] = t; ++np_; }
    }
    if (n_y) *n_y = ny;
    if (n_f) *n_f = nf;
    if (n_p) *n_p = np_;
    return 0;
}

int hmogp_create(const hmogp_config* cfg, hmogp_engine** out) {
    if (!cfg || !out || !cfg->liks) { hm_set_error("null argument"); return HMOGP_ERR_ARG; }
    if (cfg->M < 1 || cfg->Q < 1 || cfg->Q > HM_MAXQ || cfg->T < 1 || cfg->T > HM_MAXT || cfg->Xdim < 1 || cfg->Xdim > HM_MAXXD) {
        hm_set_error("config out of range: M=%d Q=%d (<=%d) T=%d (<=%d) Xdim=%d (<=%d)", cfg->M, cfg->Q, HM_MAXQ, cfg->T, HM_MAXT, cfg->Xdim, HM_MAXXD);
        return HMOGP_ERR_ARG;
    }
    if (cfg->precision < 0 || cfg->precision > HMOGP_PREC_TC) { hm_set_error("unknown precision %d", cfg->precision); return HMOGP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback");
        return HMOGP_ERR_CUDA;
    }
    HM_CUDA(cudaSetDevice(cfg->device));
    {
        cudaDeviceProp prop;
        HM_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major != 10) {
            hm_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
            return HMOGP_ERR_CUDA;
        }
    }
    hmogp_engine* e = new hmogp_engine();
    memset(&e->tk, 0, sizeof(e->tk));
    e->M = cfg->M; e->Q = cfg->Q; e->Xd = cfg->Xdim; e->T = cfg->T; e->prec = cfg->precision; e->device = cfg->device;
    e->P = cfg->M * (cfg->M + 1) / 2;
    e->Mc = (int)hm_cdiv(cfg->M, 256) * 256;
    e->Mp = 256;
    while (e->Mp < cfg->M) e->Mp *= 2;
    e->stream = nullptr; e->own_stream = false; e->timing = false; e->launches = 0; e->last_what = -1;
    e->has_chain = false;
    int J = 0;
    for (int t = 0; t < cfg->T; ++t) {
        e->liks[t] = cfg->liks[t];
        int dy, df, dp;
        int r = lik_dims(cfg->liks[t], &dy, &df, &dp);
        if (r) { delete e; return r; }
        e->tk.kind[t] = cfg->liks[t].kind; e->tk.K[t] = cfg->liks[t].K; e->tk.sigma[t] = cfg->liks[t].sigma;
        e->tk.dimf[t] = df; e->tk.foff[t] = J;
        J += df;
        e->N[t] = 0; e->Xd_[t] = nullptr; e->Yd_[t] = nullptr; e->cap[t] = 0;
    }
    if (J > HM_MAXJ) { hm_set_error("J=%d output functions > %d", J, HM_MAXJ); delete e; return HMOGP_ERR_ARG; }
    e->J = J;
    e->tk.T = e->T; e->tk.Q = e->Q; e->tk.Xdim = e->Xd; e->tk.J = J;
    e->tk.acs = (e->prec == HMOGP_PREC_TC ? 4 : 2) * e->Q;
    if (e->prec == HMOGP_PREC_TC && !hm_tc_available()) { hm_set_error("tensor-core path not built"); delete e; return HMOGP_ERR_ARG; }
    int rc = 0;
#define A_(ptr, n) if (!rc) rc = dalloc(e, &e->ptr, (size_t)(n))
    const size_t M = e->M, Mp = e->Mp, Q = e->Q, Xd = e->Xd, T = e->T, P = e->P;
    const size_t MM = Mp * Mp * Q;
    A_(pZ, M * Q * Xd); A_(pm, M * Q); A_(pL, P * Q); A_(pvar, Q); A_(pls, Q); A_(pW, J * Q); A_(pkappa, J * Q);
    A_(pWc, J * Q); A_(pkc, J * Q); A_(pbs, T); A_(consts, 1);
    A_(Zp, Q * Mp * Xd); A_(mp, Q * Mp); A_(alpha, Q * Mp); A_(kg, Q * Mp);
    A_(Kuu, MM); A_(Luu, MM); A_(LuuInv, MM); A_(Ki, MM); A_(Lu, MM); A_(LuInv, MM); A_(Sinv, MM); A_(S, MM); A_(SK, MM);
    A_(KSK, MM); A_(C, MM); A_(tmp, MM); A_(T1, MM); A_(E, MM); A_(tmpE, MM); A_(dLdS, MM); A_(dLdLfull, MM); A_(dLdK, MM);
    A_(Cf, MM);
    e->Cb = nullptr; e->tcinfo = nullptr; e->jobs_d = nullptr; e->segs_d = nullptr; e->segoff_d = nullptr; e->jobslots_d = nullptr;
    e->slots = nullptr; e->gvec = nullptr; e->plan_dirty = true;
    e->nworkers = hm_proj_workers(e->prec, e->Mc);
    if (!rc && e->prec == HMOGP_PREC_TC) {
        const char* ev = getenv("HMOGP_TC_NPASS");
        e->tc_npass = ev ? atoi(ev) : 3;
        if (e->tc_npass < 1 || e->tc_npass > 4) e->tc_npass = 3;   // 4: diagnostic (lo x lo product in the pair Gram)
        ev = getenv("HMOGP_TC_FWD_CTAS");            // 2 (default): cta_group::2 CTA pairs; 1: one CTA per SM
        e->tc_ncta = (ev && atoi(ev) == 1) ? 1 : 2;
        if (e->nworkers < 2) e->tc_ncta = 1;
        ev = getenv("HMOGP_TC_FLUSH_ROWS");          // level-1 (tensor-core fp32) accumulation window
        const char* eg = getenv("HMOGP_TC_GRAM_CTAS");   // 2 (default): CTA-pair Gram when the padded M is a multiple of 256
        // M <= 128: one 128 x 128 tile per latent on the one-CTA kernel (every SM its own chunks) instead of a 256 x 256 block
        // per CTA pair that is three quarters padding
        const bool small_m = cfg->M <= 128 && !(eg && atoi(eg) == 2);
        e->gram2 = !(eg && atoi(eg) == 1) && e->Mc % 256 == 0 && e->nworkers >= 2 && !small_m;
        e->gram_chunk = e->gram2 ? HM_GRAM2_CHUNK : HM_GRAM_CHUNK;
        eg = getenv("HMOGP_TC_GRAM_DIAG_COST");
        e->gram2_cost_diag = eg ? atoi(eg) : 75;   // plan weight of a diagonal block against an off-diagonal one (two MMA products + one generated operand tile instead of three + two; measured optimum)
        e->tc_f1 = (ev ? atoi(ev) : (e->gram2 ? 1024 : 512)) / e->gram_chunk;   // pair kernel: one window per TMEM buffer, folded straight into fp64
        if (e->tc_f1 < 1) e->tc_f1 = 1;
        ev = getenv("HMOGP_TC_FLUSH3_ROWS");         // rows between fp64 flushes
        e->tc_f2 = (ev ? atoi(ev) : 16384) / (e->tc_f1 * e->gram_chunk);
        if (e->tc_f2 < 1) e->tc_f2 = 1;
        unsigned short* cb = nullptr;
        rc = dalloc(e, &cb, hm_tc_image_elems(e->Mc, e->Q));
        e->Cb = cb;
        A_(tcinfo, 1);
        // output tiles of the lower block-triangle of an Mc x Mc Gram: rows [128 I, +128) x columns in pieces of <= 256
        if (e->gram2) {   // pair Gram: 256 x 256 blocks of the lower block-triangle
            for (int I = 0; I < e->Mc / 256; ++I)
                for (int j0 = 0; j0 <= I * 256; j0 += 256) {
                    HmGramJob jb; jb.I = I; jb.j0 = j0; jb.nw = 256;
                    e->jobs_h.push_back(jb);
                }
        } else
        for (int I = 0; I < e->Mc / 128; ++I)
            for (int j0 = 0; j0 < (I + 1) * 128; j0 += 256) {
                if (128 * I >= e->M || j0 >= e->M) continue;     // tiles that hold only padding (H is zeroed before the reduce)
                HmGramJob jb; jb.I = I; jb.j0 = j0; jb.nw = ((I + 1) * 128 - j0) < 256 ? ((I + 1) * 128 - j0) : 256;
                e->jobs_h.push_back(jb);
            }
        const size_t nj = e->jobs_h.size();
        e->max_segs = (int)(e->nworkers + Q * nj + 8);
        e->nslots_max = e->max_segs;
        A_(jobs_d, nj); A_(segs_d, e->max_segs); A_(segoff_d, e->nworkers + 1); A_(jobslots_d, Q * nj);
        A_(slots, (size_t)e->nslots_max * (e->gram2 ? 2 : 1) * HM_GRAM_SLOT_DOUBLES);
        A_(gvec, (size_t)HM_GRAM_MAXV * Q * Mp);
        if (!rc && cudaMemcpy(e->jobs_d, e->jobs_h.data(), sizeof(HmGramJob) * nj, cudaMemcpyHostToDevice) != cudaSuccess) rc = HMOGP_ERR_CUDA;
        if (!rc && cudaMemset(e->tcinfo, 0, sizeof(HmTcInfo)) != cudaSuccess) rc = HMOGP_ERR_CUDA;
    }
    A_(KLq, HM_MAXQ); A_(KLpart, HM_MAXQ * 64); A_(jitter_d, HM_MAXQ); A_(rowstat, Q * Mp * 2); A_(trq, 4 * kTraceSlices * HM_MAXQ); A_(dzmm, Q * Xd * Mp); A_(flags_d, 2 * HM_MAXQ);
    // statistics layout
    e->off_nneg = (int)T; e->off_sdv = 2 * (int)T; e->off_sma = e->off_sdv + J; e->off_svc = e->off_sma + J * (int)Q;
    e->off_dls = e->off_svc + J * (int)Q; e->off_g1 = e->off_dls + (int)Q; e->off_dz = e->off_g1 + (int)(Q * Mp);
    e->off_H = e->off_dz + (int)(Q * Xd * Mp);
    e->off_H = (e->off_H + 1) & ~1;
    e->stats_len = (int64_t)e->off_H + (int64_t)MM;
    A_(stats, e->stats_len);
    e->lik_max_blocks = 148 * 8;
    A_(lik_part, (size_t)e->lik_max_blocks * (2 + HM_MAXF * (1 + 2 * HM_MAXQ) + HM_MAXQ));
    A_(colpart, Q * e->nworkers * ((1 + Xd) * e->Mc + 1));
    e->nsplit = hm_gram_splits(e->prec, e->Mc, e->Q);
    A_(Hpart, (size_t)e->nsplit * Q * e->Mc * e->Mc);
    A_(o_lm, 1); A_(o_VE, T); A_(o_KL, 1); A_(o_dmu, M * Q); A_(o_dL, P * Q); A_(o_dKmm, Q * M * M); A_(o_drbf, Q * 2);
    A_(o_dW, J * Q); A_(o_dkappa, J * Q); A_(o_dZ, M * Q * Xd);
#undef A_
    if (!rc && hm_upload_gh_tables()) rc = HMOGP_ERR_CUDA;
    if (!rc) {
        for (int i = 0; i < 7; ++i)
            if (cudaEventCreate(&e->ev[i]) != cudaSuccess) { hm_set_error("cudaEventCreate failed"); rc = HMOGP_ERR_CUDA; break; }
    }
    e->s2 = nullptr; e->ev_fork = e->ev_S = e->ev_Sinv = nullptr;
    e->sc = nullptr; e->ev_cfence = e->ev_data = e->ev_dataY = nullptr; e->data_pending = e->dataY_pending = false;
    e->chol_graph = nullptr;
    e->prepA_graph = e->prepB_graph = e->prepR_graph = nullptr; e->prepA_launches = e->prepB_launches = e->prepR_launches = 0; e->prepare_calls = 0;
    e->kuu_valid = e->kuu_key_ok = e->hint_unchanged = false; e->kuu_reused = 0;
    { const char* v = getenv("HMOGP_NO_KUU_CACHE"); e->kuu_cache_off = v && atoi(v) != 0; }
    { const char* v = getenv("HMOGP_NO_KMM_TRACE"); e->trace_off = v && atoi(v) != 0; }
    { const char* v = getenv("HMOGP_NO_GRAPH"); e->graphs_off = v && atoi(v) != 0; }
    for (int t = 0; t < HM_MAXT; ++t) e->up_X[t] = e->up_Y[t] = nullptr;
    if (!rc && (cudaStreamCreateWithFlags(&e->s2, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_S, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_Sinv, cudaEventDisableTiming) != cudaSuccess ||
                cudaStreamCreateWithFlags(&e->sc, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_cfence, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_data, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_dataY, cudaEventDisableTiming) != cudaSuccess)) {
        hm_set_error("side stream / event creation failed");
        rc = HMOGP_ERR_CUDA;
    }
    if (rc) { hmogp_destroy(e); return rc; }
    *out = e;
    return 0;
}

void hmogp_destroy(hmogp_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->allocs) cudaFree(p);
    for (int t = 0; t < e->T; ++t) {
        if (e->Xd_[t]) cudaFree(e->Xd_[t]);
        if (e->Yd_[t]) cudaFree(e->Yd_[t]);
        if (e->tk.AC[t]) cudaFree(e->tk.AC[t]);
        if (e->tk.MW[t]) cudaFree(e->tk.MW[t]);
    }
    for (int i = 0; i < 7; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_S) cudaEventDestroy(e->ev_S);
    if (e->ev_Sinv) cudaEventDestroy(e->ev_Sinv);
    if (e->s2) cudaStreamDestroy(e->s2);
    if (e->chol_graph) cudaGraphExecDestroy(e->chol_graph);
    if (e->prepA_graph) cudaGraphExecDestroy(e->prepA_graph);
    if (e->prepB_graph) cudaGraphExecDestroy(e->prepB_graph);
    if (e->prepR_graph) cudaGraphExecDestroy(e->prepR_graph);
    for (auto& f : e->fin_graphs) if (f.exec) cudaGraphExecDestroy(f.exec);
    if (e->ev_cfence) cudaEventDestroy(e->ev_cfence);
    if (e->ev_data) cudaEventDestroy(e->ev_data);
    if (e->ev_dataY) cudaEventDestroy(e->ev_dataY);
    if (e->sc) cudaStreamDestroy(e->sc);
    delete e;
}

int hmogp_set_stream(hmogp_engine* e, void* cuda_stream) {
    if (!e) { hm_set_error("null engine"); return HMOGP_ERR_ARG; }
    e->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int hmogp_set_data(hmogp_engine* e, int32_t t, const double* X, const double* Y, int64_t N, int32_t mem_kind) {
    if (!e || t < 0 || t >= e->T || N < 0 || (N > 0 && (!X || !Y))) { hm_set_error("hmogp_set_data: bad argument"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(e->device));
    if (N > e->cap[t] || !e->Xd_[t]) {
        HM_CUDA(cudaStreamSynchronize(e->stream));
        HM_CUDA(cudaStreamSynchronize(e->sc));
        if (e->Xd_[t]) cudaFree(e->Xd_[t]);
        if (e->Yd_[t]) cudaFree(e->Yd_[t]);
        if (e->tk.AC[t]) cudaFree(e->tk.AC[t]);
        if (e->tk.MW[t]) cudaFree(e->tk.MW[t]);
        e->Xd_[t] = e->Yd_[t] = nullptr; e->tk.AC[t] = e->tk.MW[t] = nullptr;
        const size_t n = N > 0 ? (size_t)N : 1;
        HM_CUDA(cudaMalloc((void**)&e->Xd_[t], n * e->Xd * sizeof(double)));
        HM_CUDA(cudaMalloc((void**)&e->Yd_[t], n * sizeof(double)));
        HM_CUDA(cudaMalloc(&e->tk.AC[t], n * e->tk.acs * esize(e->prec)));
        HM_CUDA(cudaMalloc(&e->tk.MW[t], n * 4 * e->Q * esize(e->prec)));
        e->cap[t] = N;
        e->tk.cap[t] = (int64_t)n;
    }
    e->up_X[t] = e->up_Y[t] = nullptr;
    if (N > 0 && mem_kind == HMOGP_MEM_HOST) {
        // Pinned buffers (which the caller must keep valid until the next step has run, as for any asynchronous copy) are
        // uploaded on the copy stream, queued by the next step right after its parameter copies: the M-sized prepare phase
        // then overlaps the upload.  Pageable memory is copied here and now (the driver stages it synchronously anyway).
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, X) == cudaSuccess && at.type == cudaMemoryTypeHost &&
                            cudaPointerGetAttributes(&at, Y) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (pinned) {
            e->up_X[t] = X; e->up_Y[t] = Y;
        } else {
            HM_CUDA(cudaMemcpyAsync(e->Xd_[t], X, (size_t)N * e->Xd * sizeof(double), cudaMemcpyHostToDevice, e->stream));
            HM_CUDA(cudaMemcpyAsync(e->Yd_[t], Y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        }
    } else if (N > 0) {
        HM_CUDA(cudaMemcpyAsync(e->Xd_[t], X, (size_t)N * e->Xd * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
        HM_CUDA(cudaMemcpyAsync(e->Yd_[t], Y, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    }
    e->N[t] = N;
    e->tk.begin[t] = 0;
    e->tk.count[t] = N;
    e->plan_dirty = true;
    return 0;
}

int hmogp_set_rows(hmogp_engine* e, const int64_t* begin, const int64_t* count) {
    if (!e) { hm_set_error("null engine"); return HMOGP_ERR_ARG; }
    for (int t = 0; t < e->T; ++t) {
        const int64_t b = begin ? begin[t] : 0, c = count ? count[t] : e->N[t];
        if (b < 0 || c < 0 || b + c > e->N[t]) { hm_set_error("hmogp_set_rows: slice [%lld,+%lld) outside task %d (N=%lld)", (long long)b, (long long)c, t, (long long)e->N[t]); return HMOGP_ERR_ARG; }
        e->tk.begin[t] = b;
        e->tk.count[t] = c;
    }
    e->plan_dirty = true;
    return 0;
}

int64_t hmogp_stats_len(const hmogp_engine* e) { return e ? e->stats_len : 0; }
double* hmogp_stats_ptr(hmogp_engine* e) { return e ? e->stats : nullptr; }

int hmogp_step_local(hmogp_engine* e, const hmogp_params* p, int32_t mem_kind, int32_t what, double* stats_dev) {
    if (!e || !p || !p->Z || !p->m_u || !p->L_u || !p->rbf_var || !p->rbf_ls || !p->W || !p->kappa) { hm_set_error("hmogp_step_local: null parameter"); return HMOGP_ERR_ARG; }
    if (what < HMOGP_WHAT_ELBO || what > HMOGP_WHAT_FULL) { hm_set_error("bad 'what' %d", what); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(e->device));
    HM_CHECK(refresh_tasks(e));
    cudaStream_t s = e->stream;
    double* stats = stats_dev ? stats_dev : e->stats;
    e->last_what = what;
    e->launch0 = hm_launch_counter;
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[0], s));
    HM_CHECK(mm_prepare(e, p, mem_kind));
    HM_CHECK(data_ready(e, false));   // X only; the labels are awaited before the likelihood kernels
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[1], s));
    HM_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (what >= HMOGP_WHAT_VE ? (int64_t)e->off_H : e->stats_len), s));
    HmProjArgs pa = proj_args(e);
    const bool tc = e->prec == HMOGP_PREC_TC;
    // ---- forward projections
    if (tc) {
        HM_CUDA(cudaMemsetAsync(&e->tcinfo->wmax[0][0], 0, sizeof(unsigned) * 2 * HM_MAXQ, s));
        HM_CHECK(hm_tc_proj_fwd(s, e->tk, pa, e->Cb, e->tcinfo, what >= HMOGP_WHAT_FULL, e->tc_npass, e->tc_ncta));
    } else HM_CHECK(hm_proj_fwd(s, e->prec, e->tk, pa));
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[2], s));
    // ---- likelihoods
    HM_CHECK(data_ready(e, true));
    const int simt_prec = (e->prec == HMOGP_PREC_FP64) ? HMOGP_PREC_FP64 : HMOGP_PREC_FP32;
    for (int t = 0; t < e->T; ++t) {
        int nb = 0;
        HM_CHECK(hm_lik_rows(s, simt_prec, e->tk, e->consts, t, what >= HMOGP_WHAT_VE, e->has_chain, e->lik_part,
                             e->lik_max_blocks, &nb, nullptr, nullptr, nullptr, nullptr, nullptr, tc ? e->tcinfo : nullptr,
                             what >= HMOGP_WHAT_FULL));
        const int nstat = 2 + e->tk.dimf[t] * (1 + 2 * e->Q) + (tc ? e->Q : 0);
        reduce_lik_kernel<<<nstat, 128, 0, s>>>(e->lik_part, nb, nstat, stats, t, e->T, e->J, e->Q, e->tk.foff[t], e->tk.dimf[t],
                                            e->off_sdv, e->off_sma, e->off_svc, e->off_dls);
        HM_CUDA(cudaGetLastError());
    }
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[3], s));
    // ---- backward statistics
    if (what >= HMOGP_WHAT_VE && tc) {
        HM_CHECK(tc_backward(e, what, stats));                    // records ev[4] between projection and Gram
    } else if (what >= HMOGP_WHAT_VE) {
        HM_CHECK(hm_proj_bwd(s, simt_prec, e->tk, pa, what >= HMOGP_WHAT_FULL));
        const int ncol = (1 + e->Xd) * e->Mc + 1;
        dim3 g1((unsigned)hm_cdiv(ncol, 256), (unsigned)e->Q);
        reduce_col_kernel<<<g1, 256, 0, s>>>(e->colpart, e->nworkers, ncol, e->Mc, e->Mp, e->Xd, stats + e->off_g1,
                                             stats + e->off_dz, stats + e->off_dls);
        HM_CUDA(cudaGetLastError());
        if (e->timing) HM_CUDA(cudaEventRecord(e->ev[4], s));
        HM_CHECK(hm_gram(s, simt_prec, e->tk, pa, e->Hpart, e->nsplit));
        dim3 g2((unsigned)hm_cdiv(e->Mp, 128), (unsigned)e->Mp, (unsigned)e->Q);
        reduce_gram_kernel<<<g2, 128, 0, s>>>(e->Hpart, e->nsplit, e->Q, e->Mc, e->Mp, hm_gram_tile(simt_prec), stats + e->off_H);
        HM_CUDA(cudaGetLastError());
    } else if (e->timing) HM_CUDA(cudaEventRecord(e->ev[4], s));
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[5], s));
    return 0;
}

int hmogp_step_finish(hmogp_engine* e, const double* stats_dev, hmogp_grads* g, int32_t mem_kind, int32_t what,
                      hmogp_status* status) {
    if (!e || !g) { hm_set_error("hmogp_step_finish: null argument"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const double* stats = stats_dev ? stats_dev : e->stats;
    const int M = e->M, Mp = e->Mp, Q = e->Q, Xd = e->Xd;
    const int64_t sQ = (int64_t)Mp * Mp;
    AssembleArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.Mp = Mp; a.Q = Q; a.J = e->J; a.T = e->T; a.Xd = Xd; a.what = what;
    a.stats = stats; a.off_sdv = e->off_sdv; a.off_sma = e->off_sma; a.off_svc = e->off_svc; a.off_dls = e->off_dls;
    a.off_g1 = e->off_g1; a.off_dz = e->off_dz;
    a.KLq = e->KLq; a.kg = e->kg; a.alpha = e->alpha; a.dLdLfull = e->dLdLfull; a.dLdK = e->dLdK; a.rowstat = e->rowstat;
    a.dzmm = e->dzmm; a.c = e->consts;
    a.tr = (what >= HMOGP_WHAT_FULL && !e->trace_off) ? e->trq : nullptr; a.jitter = e->jitter_d;
    a.log_marginal = e->o_lm; a.VE = e->o_VE; a.KL = e->o_KL; a.dmu = e->o_dmu; a.dL = e->o_dL;
    a.dKmm = (g->dL_dKmm || what >= HMOGP_WHAT_FULL) ? e->o_dKmm : nullptr;
    a.drbf = e->o_drbf; a.dW = e->o_dW; a.dkappa = e->o_dkappa; a.dZ = e->o_dZ;
    // the finish chain (fixed buffers for a given statistics pointer): direct on first use, then a graph replay
    auto finish_chain = [&](cudaStream_t cs) -> int {
        if (what >= HMOGP_WHAT_VE) {
            const double* g1 = stats + e->off_g1;
            const double* H = stats + e->off_H;
            dim3 gv((unsigned)hm_cdiv(Mp, 8), (unsigned)Q);
            dgemv_kernel<<<gv, 256, 0, cs>>>(e->Ki, g1, e->kg, Mp);  // Ki g1  (svmogp_inf.py:144)
            HM_CUDA(cudaGetLastError());
            HM_CHECK(hm_dgemm(cs, false, false, Mp, Mp, Mp, 1.0, e->Ki, Mp, sQ, H, Mp, sQ, 0.0, e->T1, Mp, sQ, Q));
            HM_CHECK(hm_dgemm(cs, false, false, Mp, Mp, Mp, 1.0, e->T1, Mp, sQ, e->Ki, Mp, sQ, 0.0, e->E, Mp, sQ, Q, 1, 0, 0, 0, HM_GEMM_MIRROR));  // E = Ki H Ki
            const int64_t n = sQ * Q;
            dlds_kernel<<<(unsigned)hm_cdiv(n, 256), 256, 0, cs>>>(e->E, e->Ki, e->Sinv, e->dLdS, n);
            HM_CUDA(cudaGetLastError());
            HM_CHECK(hm_dgemm(cs, false, false, Mp, Mp, Mp, 1.0, e->dLdS, Mp, sQ, e->Lu, Mp, sQ, 0.0, e->dLdLfull, Mp, sQ, Q, 1, 0, 0, 0, HM_GEMM_LOWER | HM_GEMM_KB_GE));   // only the lower triangle is read
            if (what >= HMOGP_WHAT_FULL || a.dKmm) {
                HM_CHECK(hm_dgemm(cs, false, false, Mp, Mp, Mp, 1.0, e->E, Mp, sQ, e->SK, Mp, sQ, 0.0, e->tmpE, Mp, sQ, Q));  // E S Ki
                dim3 gk((unsigned)hm_cdiv(Mp, 128), (unsigned)Mp, (unsigned)Q);
                dldk_kernel<<<gk, 128, 0, cs>>>(e->E, e->tmpE, e->Ki, e->KSK, e->kg, e->alpha, e->dLdK, Mp);
                HM_CUDA(cudaGetLastError());
            }
            if (what >= HMOGP_WHAT_FULL) {
                dim3 gr((unsigned)hm_cdiv(M, 8), (unsigned)Q);
                kmm_grad_kernel<<<gr, 256, 0, cs>>>(e->dLdK, e->Zp, e->consts, e->rowstat, e->dzmm, M, Mp, Xd);
                HM_CUDA(cudaGetLastError());
                if (a.tr) {
                    kmm_trace_kernel<<<dim3(kTraceSlices, (unsigned)Q), 256, 0, cs>>>(stats + e->off_H, e->Ki, e->S, stats + e->off_g1, e->alpha, e->mp, e->trq, M, Mp);
                    HM_CUDA(cudaGetLastError());
                }
            }
        }
        assemble_scalar_kernel<<<1, 256, 0, cs>>>(a);
        HM_CUDA(cudaGetLastError());
        if (what >= HMOGP_WHAT_VE) {
            dim3 gm((unsigned)hm_cdiv(M, 128), (unsigned)M, (unsigned)Q);
            assemble_mat_kernel<<<gm, 128, 0, cs>>>(a);
            HM_CUDA(cudaGetLastError());
        }
        return 0;
    };
#ifdef HM_DEBUG_SKIP
    const bool fin_graphs = false;
#else
    const bool fin_graphs = !e->graphs_off && e->prepare_calls > 1;
#endif
    if (fin_graphs) {
        hmogp_engine::FinGraph* fg = nullptr;
        for (auto& f : e->fin_graphs)
            if (f.stats == stats && f.what == what && f.dkmm == (a.dKmm != nullptr)) fg = &f;
        if (!fg) {
            if (e->fin_graphs.size() >= 8) {   // a caller cycling through statistics buffers: drop the oldest
                if (e->fin_graphs.front().exec) cudaGraphExecDestroy(e->fin_graphs.front().exec);
                e->fin_graphs.erase(e->fin_graphs.begin());
            }
            hmogp_engine::FinGraph f;
            f.exec = nullptr; f.stats = stats; f.what = what; f.dkmm = a.dKmm != nullptr; f.launches = 0;
            HM_CHECK(capture_graph(e, &f.exec, &f.launches, finish_chain));
            e->fin_graphs.push_back(f);
            fg = &e->fin_graphs.back();
        }
        HM_CUDA(cudaGraphLaunch(fg->exec, s));
        hm_launch_counter += fg->launches - 1;
    } else HM_CHECK(finish_chain(s));
    if (e->timing) HM_CUDA(cudaEventRecord(e->ev[6], s));
    // ---- outputs
    const size_t P = e->P;
    HM_CHECK(copy_out(e, g->log_marginal, e->o_lm, 1, mem_kind));
    HM_CHECK(copy_out(e, g->VE, e->o_VE, e->T, mem_kind));
    HM_CHECK(copy_out(e, g->KL, e->o_KL, 1, mem_kind));
    if (what >= HMOGP_WHAT_VE) {
        HM_CHECK(copy_out(e, g->dL_dmu_u, e->o_dmu, (size_t)M * Q, mem_kind));
        HM_CHECK(copy_out(e, g->dL_dL_u, e->o_dL, P * Q, mem_kind));
        if (a.dKmm) HM_CHECK(copy_out(e, g->dL_dKmm, e->o_dKmm, (size_t)Q * M * M, mem_kind));
    }
    if (what >= HMOGP_WHAT_FULL) {
        HM_CHECK(copy_out(e, g->d_rbf, e->o_drbf, (size_t)Q * 2, mem_kind));
        HM_CHECK(copy_out(e, g->dW, e->o_dW, (size_t)e->J * Q, mem_kind));
        HM_CHECK(copy_out(e, g->dkappa, e->o_dkappa, (size_t)e->J * Q, mem_kind));
        HM_CHECK(copy_out(e, g->dZ, e->o_dZ, (size_t)M * Q * Xd, mem_kind));
    }
    int fl[HM_MAXQ];
    double nneg[HM_MAXT];
    HM_CUDA(cudaMemcpyAsync(fl, e->flags_d + HM_MAXQ, sizeof(int) * Q, cudaMemcpyDeviceToHost, s));
    HM_CUDA(cudaMemcpyAsync(nneg, stats + e->off_nneg, sizeof(double) * e->T, cudaMemcpyDeviceToHost, s));
    HM_CUDA(cudaStreamSynchronize(s));
    if (e->timing) {
        for (int i = 0; i < 6; ++i) {
            if (cudaEventElapsedTime(&e->ms[i], e->ev[i], e->ev[i + 1]) != cudaSuccess) e->ms[i] = -1.f;
        }
    }
    e->launches = hm_launch_counter - e->launch0;
    bool unstable = false;
    if (status) {
        memset(status, 0, sizeof(*status));
        for (int q = 0; q < Q; ++q) {
            status->jitter[q] = e->jitter_h[q];
            status->chol_fail[q] = e->chol_fail_h[q];
            status->lu_singular[q] = fl[q];
        }
        for (int t = 0; t < e->T; ++t) status->n_negative_v += (int64_t)nneg[t];
    }
    for (int q = 0; q < Q; ++q) unstable = unstable || fl[q];
    if (unstable && what >= HMOGP_WHAT_VE) {
        hm_set_error("Sqi: Cholesky representation unstable");
        return HMOGP_ERR_UNSTABLE;
    }
    return 0;
}

int hmogp_elbo_and_grads(hmogp_engine* e, const hmogp_params* p, hmogp_grads* g, int32_t mem_kind, int32_t what,
                         hmogp_status* status) {
    HM_CHECK(hmogp_step_local(e, p, mem_kind, what, nullptr));
    return hmogp_step_finish(e, nullptr, g, mem_kind, what, status);
}

int hmogp_inference_host(const hmogp_config* cfg, const double* const* X, const double* const* Y, const int64_t* N,
                         const hmogp_params* p, hmogp_grads* g, int32_t what, hmogp_status* status) {
    hmogp_engine* e = nullptr;
    HM_CHECK(hmogp_create(cfg, &e));
    int rc = 0;
    for (int t = 0; t < cfg->T && !rc; ++t) rc = hmogp_set_data(e, t, X[t], Y[t], N[t], HMOGP_MEM_HOST);
    if (!rc) rc = hmogp_elbo_and_grads(e, p, g, HMOGP_MEM_HOST, what, status);
    hmogp_destroy(e);
    return rc;
}

int hmogp_predict_f(hmogp_engine* e, const hmogp_params* p, int32_t mem_kind, int32_t t, const double* Xnew, int64_t N,
                    double* m_fd, double* v_fd) {
    if (!e || !p || t < 0 || t >= e->T || N < 0 || (N > 0 && (!Xnew || !m_fd || !v_fd))) { hm_set_error("hmogp_predict_f: bad argument"); return HMOGP_ERR_ARG; }
    if (!p->Z || !p->m_u || !p->L_u || !p->rbf_var || !p->rbf_ls || !p->W || !p->kappa) { hm_set_error("hmogp_predict_f: null parameter"); return HMOGP_ERR_ARG; }
    if (N == 0) return 0;
    HM_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    HM_CHECK(mm_prepare(e, p, mem_kind));
    const int F = e->tk.dimf[t];
    const bool host = mem_kind == HMOGP_MEM_HOST;
    double *xd = nullptr, *md = nullptr, *vd = nullptr;
    void* ac = nullptr;
    int rc = 0;
    auto fail = [&](const char* what) { hm_set_error("hmogp_predict_f: %s: %s", what, cudaGetErrorString(cudaGetLastError())); return HMOGP_ERR_CUDA; };
    if (host) {
        if (cudaMalloc((void**)&xd, sizeof(double) * N * e->Xd) != cudaSuccess || cudaMalloc((void**)&md, sizeof(double) * N * F) != cudaSuccess ||
            cudaMalloc((void**)&vd, sizeof(double) * N * F) != cudaSuccess) rc = fail("cudaMalloc");
        if (!rc && cudaMemcpyAsync(xd, Xnew, sizeof(double) * N * e->Xd, cudaMemcpyHostToDevice, s) != cudaSuccess) rc = fail("upload");
    } else { xd = const_cast<double*>(Xnew); md = m_fd; vd = v_fd; }
    if (!rc && cudaMalloc(&ac, (size_t)N * e->tk.acs * esize(e->prec)) != cudaSuccess) rc = fail("cudaMalloc");
    if (!rc) {
        // a one-task view of the task table: only task t has rows, and they are the new inputs
        HmTasks tk = e->tk;
        for (int u = 0; u < e->T; ++u) { tk.count[u] = 0; tk.begin[u] = 0; }
        tk.X[t] = xd; tk.Y[t] = nullptr; tk.count[t] = N; tk.AC[t] = ac; tk.MW[t] = nullptr; tk.cap[t] = N;
        HmProjArgs pa = proj_args(e);
        if (e->prec == HMOGP_PREC_TC) rc = hm_tc_proj_fwd(s, tk, pa, e->Cb, e->tcinfo, false, e->tc_npass, e->tc_ncta);
        else rc = hm_proj_fwd(s, e->prec, tk, pa);
        if (!rc) {
            const unsigned nb = (unsigned)hm_cdiv(N, 256);
            if (e->prec == HMOGP_PREC_FP64) mix_rows_kernel<double><<<nb, 256, 0, s>>>(ac, N, N, e->Q, e->tk.foff[t], F, e->consts, md, vd);
            else mix_rows_kernel<float><<<nb, 256, 0, s>>>(ac, N, N, e->Q, e->tk.foff[t], F, e->consts, md, vd);
            if (cudaGetLastError() != cudaSuccess) rc = fail("mix_rows_kernel");
        }
        if (!rc && host) {
            if (cudaMemcpyAsync(m_fd, md, sizeof(double) * N * F, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaMemcpyAsync(v_fd, vd, sizeof(double) * N * F, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = fail("download");
        }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && !rc) rc = fail("sync");
    if (host) { cudaFree(xd); cudaFree(md); cudaFree(vd); }
    cudaFree(ac);
    return rc;
}

int hmogp_get_rows(hmogp_engine* e, int32_t t, double* m_fd, double* v_fd, double* VE, double* dm, double* dv) {
    if (!e || t < 0 || t >= e->T || e->last_what < 0) { hm_set_error("hmogp_get_rows: no evaluation yet / bad task"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(e->device));
    HM_CHECK(data_ready(e));
    const int64_t n = e->tk.count[t];
    const int F = e->tk.dimf[t];
    if (n == 0) return 0;
    double* buf = nullptr;
    HM_CUDA(cudaMalloc((void**)&buf, sizeof(double) * n * (4 * F + 1)));
    double *rm = buf, *rv = rm + n * F, *rdm = rv + n * F, *rdv = rdm + n * F, *rve = rdv + n * F;
    int nb = 0;
    const int simt_prec = (e->prec == HMOGP_PREC_FP64) ? HMOGP_PREC_FP64 : HMOGP_PREC_FP32;
    double* part = nullptr;
    HM_CUDA(cudaMalloc((void**)&part, sizeof(double) * e->lik_max_blocks * (2 + HM_MAXF * (1 + 2 * HM_MAXQ) + HM_MAXQ)));
    HmTasks tk = e->tk;
    void* mwtmp = nullptr;  // do not disturb the row weights of the last evaluation
    HM_CUDA(cudaMalloc(&mwtmp, (size_t)e->tk.cap[t] * 4 * e->Q * esize(e->prec)));   // same SoA stride as AC
    tk.MW[t] = mwtmp;
    int rc = hm_lik_rows(e->stream, simt_prec, tk, e->consts, t, true, e->has_chain, part, e->lik_max_blocks, &nb, rm, rv, rve, rdm, rdv);
    if (!rc) {
        cudaStream_t s = e->stream;
        if (m_fd) cudaMemcpyAsync(m_fd, rm, sizeof(double) * n * F, cudaMemcpyDeviceToHost, s);
        if (v_fd) cudaMemcpyAsync(v_fd, rv, sizeof(double) * n * F, cudaMemcpyDeviceToHost, s);
        if (dm) cudaMemcpyAsync(dm, rdm, sizeof(double) * n * F, cudaMemcpyDeviceToHost, s);
        if (dv) cudaMemcpyAsync(dv, rdv, sizeof(double) * n * F, cudaMemcpyDeviceToHost, s);
        if (VE) cudaMemcpyAsync(VE, rve, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { hm_set_error("hmogp_get_rows: sync failed"); rc = HMOGP_ERR_CUDA; }
    }
    cudaFree(buf); cudaFree(part); cudaFree(mwtmp);
    return rc;
}

int hmogp_get_dL_dKmn(hmogp_engine* e, int32_t q, int32_t d, double* dL_dKmn, double* dL_dKdiag) {
    if (!e || q < 0 || q >= e->Q || d < 0 || d >= e->J || e->last_what < HMOGP_WHAT_VE) { hm_set_error("hmogp_get_dL_dKmn: bad argument / no gradient evaluation yet"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(e->device));
    HM_CHECK(data_ready(e));
    int t = 0;
    while (t + 1 < e->T && e->tk.foff[t + 1] <= d) ++t;
    const int f = d - e->tk.foff[t], F = e->tk.dimf[t];
    const int64_t n = e->tk.count[t];
    if (n == 0) return 0;
    std::vector<double> dm((size_t)n * F), dv((size_t)n * F);
    HM_CHECK(hmogp_get_rows(e, t, nullptr, nullptr, nullptr, dm.data(), dv.data()));
    if (dL_dKdiag) for (int64_t i = 0; i < n; ++i) dL_dKdiag[i] = dv[(size_t)i * F + f];  // svmogp_inf.py:164
    if (!dL_dKmn) return 0;
    double *ddm = nullptr, *ddv = nullptr, *out = nullptr;
    HM_CUDA(cudaMalloc((void**)&ddm, sizeof(double) * n * F));
    HM_CUDA(cudaMalloc((void**)&ddv, sizeof(double) * n * F));
    HM_CUDA(cudaMalloc((void**)&out, sizeof(double) * n * e->M));
    cudaStream_t s = e->stream;
    cudaMemcpyAsync(ddm, dm.data(), sizeof(double) * n * F, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(ddv, dv.data(), sizeof(double) * n * F, cudaMemcpyHostToDevice, s);
    dim3 grid((unsigned)hm_cdiv(n, 128), (unsigned)e->M);
    dense_dkmn_kernel<<<grid, 128, 0, s>>>(e->Xd_[t] + e->tk.begin[t] * e->Xd, n, e->Zp, e->C + (int64_t)q * e->Mp * e->Mp,
                                          e->alpha + (int64_t)q * e->Mp, e->consts, q, d, ddm, ddv, F, f, out, e->M, e->Mp, e->Xd);
    cudaMemcpyAsync(dL_dKmn, out, sizeof(double) * n * e->M, cudaMemcpyDeviceToHost, s);
    cudaError_t err = cudaStreamSynchronize(s);
    cudaFree(ddm); cudaFree(ddv); cudaFree(out);
    if (err != cudaSuccess) { hm_set_error("hmogp_get_dL_dKmn: %s", cudaGetErrorString(err)); return HMOGP_ERR_CUDA; }
    return 0;
}

int hmogp_get_kuu(hmogp_engine* e, double* Kuu, double* Luu, double* Kuui) {
    if (!e || e->last_what < 0) { hm_set_error("hmogp_get_kuu: no evaluation yet"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(e->device));
    const size_t n = (size_t)e->Q * e->M * e->M;
    double* buf = nullptr;
    HM_CUDA(cudaMalloc((void**)&buf, sizeof(double) * n));
    dim3 grid((unsigned)hm_cdiv(e->M, 128), (unsigned)e->M, (unsigned)e->Q);
    const double* srcs[3] = {e->Kuu, e->Luu, e->Ki};
    double* dsts[3] = {Kuu, Luu, Kuui};
    int rc = 0;
    for (int i = 0; i < 3 && !rc; ++i) {
        if (!dsts[i]) continue;
        extract_mm_kernel<<<grid, 128, 0, e->stream>>>(srcs[i], buf, e->M, e->Mp, i == 1 ? 1 : 0);
        if (cudaMemcpyAsync(dsts[i], buf, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) { hm_set_error("hmogp_get_kuu: copy failed"); rc = HMOGP_ERR_CUDA; }
    }
    cudaFree(buf);
    return rc;
}

// ---- stand-alone likelihood / index entry points
static int staged_call(int mem_kind, cudaStream_t s, std::vector<std::pair<const double*, size_t>> ins,
                       std::vector<std::pair<double*, size_t>> outs, std::vector<double*>& din, std::vector<double*>& dout) {
    // host pointers: stage through device buffers; device pointers: pass through
    din.clear(); dout.clear();
    for (auto& in : ins) {
        if (mem_kind == HMOGP_MEM_DEVICE || !in.first) { din.push_back((double*)in.first); continue; }
        double* d = nullptr;
        HM_CUDA(cudaMalloc((void**)&d, in.second * sizeof(double) + 16));
        HM_CUDA(cudaMemcpyAsync(d, in.first, in.second * sizeof(double), cudaMemcpyHostToDevice, s));
        din.push_back(d);
    }
    for (auto& o : outs) {
        if (mem_kind == HMOGP_MEM_DEVICE || !o.first) { dout.push_back(o.first); continue; }
        double* d = nullptr;
        HM_CUDA(cudaMalloc((void**)&d, o.second * sizeof(double) + 16));
        dout.push_back(d);
    }
    return 0;
}
static int staged_finish(int mem_kind, cudaStream_t s, std::vector<std::pair<const double*, size_t>> ins,
                         std::vector<std::pair<double*, size_t>> outs, std::vector<double*>& din, std::vector<double*>& dout) {
    int rc = 0;
    if (mem_kind == HMOGP_MEM_HOST) {
        for (size_t i = 0; i < outs.size(); ++i)
            if (outs[i].first && cudaMemcpyAsync(outs[i].first, dout[i], outs[i].second * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = HMOGP_ERR_CUDA;
        if (cudaStreamSynchronize(s) != cudaSuccess) rc = HMOGP_ERR_CUDA;
        for (size_t i = 0; i < ins.size(); ++i) if (ins[i].first) cudaFree(din[i]);
        for (size_t i = 0; i < outs.size(); ++i) if (outs[i].first) cudaFree(dout[i]);
        if (rc) hm_set_error("device->host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return rc;
}

int hmogp_lik_var_exp(const hmogp_lik_desc* lik, int64_t N, const double* Y, const double* Mf, const double* Vf, double* VE,
                      double* dm, double* dv, int32_t precision, int32_t mem_kind, void* cuda_stream) {
    if (!lik || N < 0 || !Y || !Mf || !Vf) { hm_set_error("hmogp_lik_var_exp: null argument"); return HMOGP_ERR_ARG; }
    if (hmogp_device_count() == 0) { hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback"); return HMOGP_ERR_CUDA; }
    int dy, F, dp;
    HM_CHECK(lik_dims(*lik, &dy, &F, &dp));
    HM_CHECK(hm_upload_gh_tables());
    cudaStream_t s = (cudaStream_t)cuda_stream;
    std::vector<std::pair<const double*, size_t>> ins = {{Y, (size_t)N}, {Mf, (size_t)N * F}, {Vf, (size_t)N * F}};
    std::vector<std::pair<double*, size_t>> outs = {{VE, (size_t)N}, {dm, (size_t)N * F}, {dv, (size_t)N * F}};
    std::vector<double*> din, dout;
    HM_CHECK(staged_call(mem_kind, s, ins, outs, din, dout));
    int rc = hm_lik_var_exp(s, precision == HMOGP_PREC_FP64 ? HMOGP_PREC_FP64 : HMOGP_PREC_FP32, *lik, N, din[0], din[1], din[2], dout[0], dout[1], dout[2]);
    int rc2 = staged_finish(mem_kind, s, ins, outs, din, dout);
    return rc ? rc : rc2;
}

int hmogp_lik_predictive(const hmogp_lik_desc* lik, int64_t N, const double* Mf, const double* Vf, double* mean_pred,
                         double* var_pred, int32_t gh_tensor, int32_t mem_kind, void* cuda_stream) {
    if (!lik || N < 0 || !Mf || !Vf || !mean_pred || !var_pred) { hm_set_error("hmogp_lik_predictive: null argument"); return HMOGP_ERR_ARG; }
    if (hmogp_device_count() == 0) { hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback"); return HMOGP_ERR_CUDA; }
    int dy, F, dp;
    HM_CHECK(lik_dims(*lik, &dy, &F, &dp));
    HM_CHECK(hm_upload_gh_tables());
    cudaStream_t s = (cudaStream_t)cuda_stream;
    std::vector<std::pair<const double*, size_t>> ins = {{Mf, (size_t)N * F}, {Vf, (size_t)N * F}};
    std::vector<std::pair<double*, size_t>> outs = {{mean_pred, (size_t)N * dp}, {var_pred, (size_t)N * dp}};
    std::vector<double*> din, dout;
    HM_CHECK(staged_call(mem_kind, s, ins, outs, din, dout));
    int rc = hm_lik_predictive(s, *lik, gh_tensor, N, din[0], din[1], dout[0], dout[1]);
    int rc2 = staged_finish(mem_kind, s, ins, outs, din, dout);
    return rc ? rc : rc2;
}

int hmogp_lik_pointwise(const hmogp_lik_desc* lik, int64_t N, const double* F, const double* Y, double* logp, double* dlogp,
                        double* d2logp, int32_t mem_kind, void* cuda_stream) {
    if (!lik || N < 0 || !Y || !F) { hm_set_error("hmogp_lik_pointwise: null argument"); return HMOGP_ERR_ARG; }
    if (hmogp_device_count() == 0) { hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback"); return HMOGP_ERR_CUDA; }
    int dy, nf, dp;
    HM_CHECK(lik_dims(*lik, &dy, &nf, &dp));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    std::vector<std::pair<const double*, size_t>> ins = {{F, (size_t)N * nf}, {Y, (size_t)N}};
    std::vector<std::pair<double*, size_t>> outs = {{logp, (size_t)N}, {dlogp, (size_t)N * nf}, {d2logp, (size_t)N * nf}};
    std::vector<double*> din, dout;
    HM_CHECK(staged_call(mem_kind, s, ins, outs, din, dout));
    int rc = hm_lik_pointwise(s, *lik, N, din[0], din[1], dout[0], dout[1], dout[2]);
    int rc2 = staged_finish(mem_kind, s, ins, outs, din, dout);
    return rc ? rc : rc2;
}

int hmogp_flat_to_triang(const double* flat, double* dense, int32_t M, int32_t D, int32_t mem_kind, void* cuda_stream) {
    if (!flat || !dense || M < 1 || D < 1) { hm_set_error("hmogp_flat_to_triang: bad argument"); return HMOGP_ERR_ARG; }
    if (hmogp_device_count() == 0) { hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback"); return HMOGP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t P = (size_t)M * (M + 1) / 2;
    std::vector<std::pair<const double*, size_t>> ins = {{flat, P * D}};
    std::vector<std::pair<double*, size_t>> outs = {{dense, (size_t)D * M * M}};
    std::vector<double*> din, dout;
    HM_CHECK(staged_call(mem_kind, s, ins, outs, din, dout));
    dim3 grid((unsigned)hm_cdiv(M, 128), (unsigned)M, (unsigned)D);
    flat_to_triang_kernel<<<grid, 128, 0, s>>>(din[0], dout[0], M, D);
    int rc = cudaGetLastError() == cudaSuccess ? 0 : HMOGP_ERR_CUDA;
    int rc2 = staged_finish(mem_kind, s, ins, outs, din, dout);
    return rc ? rc : rc2;
}

int hmogp_triang_to_flat(const double* dense, double* flat, int32_t M, int32_t D, int32_t mem_kind, void* cuda_stream) {
    if (!flat || !dense || M < 1 || D < 1) { hm_set_error("hmogp_triang_to_flat: bad argument"); return HMOGP_ERR_ARG; }
    if (hmogp_device_count() == 0) { hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback"); return HMOGP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t P = (size_t)M * (M + 1) / 2;
    std::vector<std::pair<const double*, size_t>> ins = {{dense, (size_t)D * M * M}};
    std::vector<std::pair<double*, size_t>> outs = {{flat, P * D}};
    std::vector<double*> din, dout;
    HM_CHECK(staged_call(mem_kind, s, ins, outs, din, dout));
    dim3 grid((unsigned)hm_cdiv(M, 128), (unsigned)M, (unsigned)D);
    triang_to_flat_kernel<<<grid, 128, 0, s>>>(din[0], dout[0], M, D);
    int rc = cudaGetLastError() == cudaSuccess ? 0 : HMOGP_ERR_CUDA;
    int rc2 = staged_finish(mem_kind, s, ins, outs, din, dout);
    return rc ? rc : rc2;
}

int hmogp_hint_hyper_unchanged(hmogp_engine* e, int32_t unchanged) {
    if (!e) { hm_set_error("null engine"); return HMOGP_ERR_ARG; }
    e->hint_unchanged = unchanged != 0;
    return 0;
}

int64_t hmogp_kuu_reuse_count(hmogp_engine* e) { return e ? (int64_t)e->kuu_reused : -1; }

int hmogp_enable_timing(hmogp_engine* e, int32_t on) {
    if (!e) { hm_set_error("null engine"); return HMOGP_ERR_ARG; }
    e->timing = on != 0;
    return 0;
}

int hmogp_last_timing(hmogp_engine* e, float* ms, int32_t* launches) {
    if (!e) { hm_set_error("null engine"); return HMOGP_ERR_ARG; }
    if (ms) for (int i = 0; i < 6; ++i) ms[i] = e->ms[i];
    if (launches) *launches = (int32_t)e->launches;
    return 0;
}

int hmogp_tc_built(void) { return hm_tc_available(); }

}  // extern "C"
