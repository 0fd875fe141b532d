// Device functors of the heterogeneous likelihood list: E_q[log p], E_q[dlogp/df], 0.5 E_q[d2logp/df2]
// under q(f) = N(m, diag v), analytic or Gauss-Hermite, templated on the arithmetic type.
//
// Follows /root/reference/likelihoods/*.py (hot-path methods only) including the quirks
// the reference has: Categorical dlogp_df is the constant 1[y=d+1]-1 (C-1), Gamma/Beta
// expectations carry an extra 1/pi (C-2), GH order 20 for 1-D and 10 per axis for tensor grids (C-3).
#pragma once
#include "common.cuh"

// Gauss-Hermite tables: nodes x_i and NORMALISED weights w_i/sqrt(pi) (bernoulli.py:90, categorical.py:166).
// (defined here: this header is included by exactly one translation unit, lik_kernels.cu)
__constant__ double c_gh20_x[20];
__constant__ double c_gh20_w[20];
__constant__ double c_gh10_x[10];
__constant__ double c_gh10_w[10];

template <typename T> struct HmNum;
template <> struct HmNum<double> {
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double lgamma_(double x) { return lgamma(x); }
    static __device__ __forceinline__ double exp_lim() { return 709.782712893384; }  // log(DBL_MAX): GPy safe_exp
    static __device__ __forceinline__ double one_minus_eps() { return 1.0 - 1e-9; }
    static __device__ __forceinline__ double asym_start() { return 10.0; }
    static __device__ __forceinline__ double inf_() { return CUDART_INF; }
};
template <> struct HmNum<float> {
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float lgamma_(float x) { return lgammaf(x); }
    static __device__ __forceinline__ float exp_lim() { return 80.0f; }
    static __device__ __forceinline__ float one_minus_eps() { return 1.0f; }
    static __device__ __forceinline__ float asym_start() { return 6.0f; }
    static __device__ __forceinline__ float inf_() { return CUDART_INF_F; }
};

template <typename T> __device__ __forceinline__ T hm_safe_exp(T f) {
    return HmNum<T>::exp_(f < HmNum<T>::exp_lim() ? f : HmNum<T>::exp_lim());
}
template <typename T> __device__ __forceinline__ T hm_clip(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// digamma (scipy.special.psi) for x > 0: upward recurrence to x >= x0, then the asymptotic series.
template <typename T> __device__ T hm_digamma(T x) {
    T r = T(0);
    while (x < HmNum<T>::asym_start()) { r -= T(1) / x; x += T(1); }
    const T xi = T(1) / x, x2 = xi * xi;
    r += HmNum<T>::log_(x) - T(0.5) * xi -
         x2 * (T(1.0 / 12) - x2 * (T(1.0 / 120) - x2 * (T(1.0 / 252) - x2 * (T(1.0 / 240) - x2 * T(1.0 / 132)))));
    return r;
}
// trigamma = Hurwitz zeta(2, x) (scipy.special.zeta(2, x)) for x > 0.
template <typename T> __device__ T hm_trigamma(T x) {
    T r = T(0);
    while (x < HmNum<T>::asym_start()) { r += T(1) / (x * x); x += T(1); }
    const T xi = T(1) / x, x2 = xi * xi;
    r += xi * (T(1) + T(0.5) * xi +
               x2 * (T(1.0 / 6) - x2 * (T(1.0 / 30) - x2 * (T(1.0 / 42) - x2 * (T(1.0 / 30) - x2 * (T(5.0 / 66) - x2 * T(691.0 / 2730)))))));
    return r;
}

template <typename T> struct HmLikOut {
    T ve;
    T dm[HM_MAXF];
    T dv[HM_MAXF];
};

// ---------------------------------------------------------------- pointwise logpdf / dlogp_df / d2logp_df2
// 1-D likelihoods at a single f.
template <typename T> __device__ __forceinline__ void hm_point_1d(int kind, T f, T y, T lgy1, T& lp, T& d1, T& d2) {
    if (kind == HMOGP_LIK_BERNOULLI) {  // bernoulli.py:31-36,66-80
        const T ef = hm_safe_exp(f);
        const T inv = T(1) / (T(1) + ef);
        const T p = hm_clip(ef * inv, T(1e-9), HmNum<T>::one_minus_eps());
        const T qm = hm_clip(inv, T(1e-9), HmNum<T>::one_minus_eps());  // 1 - p under the same clip
        lp = y * HmNum<T>::log_(p) + (T(1) - y) * HmNum<T>::log_(qm);
        d1 = ((y - p) / qm) * inv;
        d2 = -p * inv;
    } else if (kind == HMOGP_LIK_POISSON) {  // poisson.py:31-34,56-64
        const T ef = hm_safe_exp(f);
        lp = -ef + y * f - lgy1;
        d1 = -ef + y;
        d2 = -ef;
    } else {  // HMOGP_LIK_EXPONENTIAL  exponential.py:28-32,58-68
        const T b = hm_clip(hm_safe_exp(-f), T(1e-9), T(1e9));
        lp = -HmNum<T>::log_(b) - y / b;
        d1 = T(1) - y / b;
        d2 = -y / b;
    }
}

// ---------------------------------------------------------------- variational expectations
template <typename T>
__device__ void hm_lik_eval(int kind, int K, T sigma, T y, const T* m, const T* v, HmLikOut<T>& o) {
    const T half = T(0.5);
    const T log2pi = T(1.8378770664093453);
    if (kind == HMOGP_LIK_GAUSSIAN) {  // gaussian.py:41-62
        const T lv = sigma * sigma;
        o.ve = -half * log2pi - half * HmNum<T>::log_(lv) - half * (y * y + m[0] * m[0] + v[0] - T(2) * m[0] * y) / lv;
        o.dm[0] = -(m[0] - y) / lv;
        o.dv[0] = -half / lv;
        return;
    }
    if (kind == HMOGP_LIK_HETGAUSSIAN) {  // hetgaussian.py:46-73
        T prec = hm_safe_exp(-m[1] + half * v[1]);
        prec = hm_clip(prec, T(-1e9), T(1e9));
        const T sq = hm_clip(y * y + m[0] * m[0] + v[0] - T(2) * m[0] * y, T(-1e9), T(1e9));
        o.ve = -half * log2pi - half * m[1] - half * prec * sq;
        o.dm[0] = prec * (y - m[0]);
        o.dm[1] = half * (prec * sq - T(1));
        o.dv[0] = -half * prec;
        o.dv[1] = -T(0.25) * prec * sq;
        return;
    }
    if (kind == HMOGP_LIK_BERNOULLI || kind == HMOGP_LIK_POISSON || kind == HMOGP_LIK_EXPONENTIAL) {
        // GH-20: f_i = x_i sqrt(2 v) + m, weights w_i/sqrt(pi)  (bernoulli.py:82-111)
        const T s2v = HmNum<T>::sqrt_(T(2) * v[0]);
        const T lgy1 = (kind == HMOGP_LIK_POISSON) ? HmNum<T>::lgamma_(y + T(1)) : T(0);
        T ve = 0, d1 = 0, d2 = 0;
#pragma unroll 4
        for (int i = 0; i < 20; ++i) {
            const T f = T(c_gh20_x[i]) * s2v + m[0];
            const T w = T(c_gh20_w[i]);
            T lp, a, b;
            hm_point_1d<T>(kind, f, y, lgy1, lp, a, b);
            ve += lp * w;
            d1 += a * w;
            d2 += b * w;
        }
        o.ve = ve;
        o.dm[0] = d1;
        o.dv[0] = half * d2;
        return;
    }
    if (kind == HMOGP_LIK_CATEGORICAL) {  // categorical.py:37-46,102-128,130-222
        const int D = K - 1;
        T e[HM_MAXF][10];
        for (int d = 0; d < D; ++d) {
            const T s2v = HmNum<T>::sqrt_(T(2) * v[d]);
            for (int i = 0; i < 10; ++i) e[d][i] = hm_safe_exp(T(c_gh10_x[i]) * s2v + m[d]);
        }
        const int label = (int)y;
        const bool valid = (T(label) == y) && label >= 1 && label <= K;
        int total = 1;
        for (int d = 0; d < D; ++d) total *= 10;
        T ve = 0, wsum = 0;
        T d2[HM_MAXF];
        for (int d = 0; d < HM_MAXF; ++d) d2[d] = 0;
        const T hi = HmNum<T>::one_minus_eps();
        for (int g = 0; g < total; ++g) {
            int idx[HM_MAXF];
            {   // C-order flattening: function 0 is the slowest axis (categorical.py:153-157)
                int r = g;
                for (int d = D - 1; d >= 0; --d) { idx[d] = r % 10; r /= 10; }
            }
            T w = 1, den = 1;
            T ed[HM_MAXF];
            for (int d = 0; d < D; ++d) {
                ed[d] = e[d][idx[d]];
                den += ed[d];
                w *= T(c_gh10_w[idx[d]]);
            }
            const T inv = T(1) / den;
            T psum = hm_clip(inv, T(1e-9), hi);
            T py = psum;  // class K
            for (int d = 0; d < D; ++d) {
                const T pk = hm_clip(ed[d] * inv, T(1e-9), hi);
                psum += pk;
                if (label == d + 1) py = pk;
            }
            ve += w * HmNum<T>::log_(py / psum);
            wsum += w;
            for (int d = 0; d < D; ++d) d2[d] += w * (-(ed[d] * (den - ed[d])) * inv * inv);
        }
        o.ve = valid ? ve : -HmNum<T>::inf_();
        for (int d = 0; d < D; ++d) {
            // quirk C-1: dlogp_df == Y_oneK[:,d] - sum_k Y_oneK[:,k]   (categorical.py:107-111)
            o.dm[d] = valid ? ((label == d + 1 ? T(1) : T(0)) - T(1)) * wsum : T(0);
            o.dv[d] = valid ? half * d2[d] : T(0);
        }
        return;
    }
    // Gamma / Beta: 10x10 grid over (log a, log b), double 1/sqrt(pi) normalisation => extra 1/pi (quirk C-2)
    {
        const T inv_pi = T(0.31830988618379067);
        const T sa = HmNum<T>::sqrt_(T(2) * v[0]), sb = HmNum<T>::sqrt_(T(2) * v[1]);
        T a[10], b[10], lga[10], psa[10], tra[10], lgb[10], psb[10], trb[10], logb[10];
        const bool is_gamma = (kind == HMOGP_LIK_GAMMA);
        for (int i = 0; i < 10; ++i) {
            a[i] = hm_clip(hm_safe_exp(T(c_gh10_x[i]) * sa + m[0]), T(1e-9), T(1e9));
            b[i] = hm_clip(hm_safe_exp(T(c_gh10_x[i]) * sb + m[1]), T(1e-9), T(1e9));
            lga[i] = HmNum<T>::lgamma_(a[i]);
            psa[i] = hm_digamma(a[i]);
            tra[i] = hm_trigamma(a[i]);
            if (is_gamma) {
                logb[i] = HmNum<T>::log_(b[i]);
            } else {
                lgb[i] = HmNum<T>::lgamma_(b[i]);
                psb[i] = hm_digamma(b[i]);
                trb[i] = hm_trigamma(b[i]);
            }
        }
        const T logy = HmNum<T>::log_(y);
        const T log1y = is_gamma ? T(0) : HmNum<T>::log_(T(1) - y);
        T ve = 0, da = 0, db = 0, d2a = 0, d2b = 0;
        for (int i = 0; i < 10; ++i) {
            const T wi = T(c_gh10_w[i]);
            T rve = 0, rda = 0, rdb = 0, r2a = 0, r2b = 0;
            for (int j = 0; j < 10; ++j) {
                const T wj = T(c_gh10_w[j]);
                if (is_gamma) {  // gamma.py:34-41,80-101
                    rve += wj * (-lga[i] + a[i] * logb[j] + (a[i] - T(1)) * logy - b[j] * y);
                    rda += wj * ((-psa[i] + logb[j] + logy) * a[i]);
                    rdb += wj * (a[i] - b[j] * y);
                    r2a += wj * ((-psa[i] - a[i] * tra[i] + logb[j] + logy) * a[i]);
                    r2b += wj * (-y * b[j]);
                } else {  // beta.py:29-36,76-104
                    const T ab = a[i] + b[j];
                    const T lgab = HmNum<T>::lgamma_(ab), psab = hm_digamma(ab), trab = hm_trigamma(ab);
                    rve += wj * ((a[i] - T(1)) * logy + (b[j] - T(1)) * log1y - (lga[i] + lgb[j] - lgab));
                    rda += wj * ((psab - psa[i] + logy) * a[i]);
                    rdb += wj * ((psab - psb[j] + log1y) * b[j]);
                    r2a += wj * ((psab + a[i] * trab - psa[i] - a[i] * tra[i] + logy) * a[i]);
                    r2b += wj * ((psab + b[j] * trab - psb[j] - b[j] * trb[j] + log1y) * b[j]);
                }
            }
            ve += wi * rve; da += wi * rda; db += wi * rdb; d2a += wi * r2a; d2b += wi * r2b;
        }
        o.ve = ve * inv_pi;
        o.dm[0] = da * inv_pi;
        o.dm[1] = db * inv_pi;
        o.dv[0] = half * d2a * inv_pi;
        o.dv[1] = half * d2b * inv_pi;
    }
}

// ================================================================== fp32 fast paths (tensor-core / fp32 engines)
// Same expectations as hm_lik_eval<float>, restructured for throughput; each one states why it is equivalent.
__device__ __forceinline__ float hm_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float hm_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// lgamma, digamma and trigamma of x > 0 in one pass: a predicated 6-step upward recurrence shared by the three
// (product for lgamma, sum 1/x for psi, sum 1/x^2 for zeta(2,.)), then the Stirling / asymptotic series at x >= 6.
__device__ __forceinline__ void hm_lgam_psi_tri(float x, float& lg, float& ps, float& tr) {
    float rec = 0.f, rec2 = 0.f, prod = 1.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (x < 6.f) {
            const float r = hm_rcp(x);
            rec += r;
            rec2 = fmaf(r, r, rec2);
            prod *= x;
            x += 1.f;
        }
    }
    const float xi = hm_rcp(x), x2 = xi * xi;
    const float lx = 0.6931471805599453f * hm_lg2(x);
    lg = fmaf(x - 0.5f, lx, -x) + 0.9189385332046727f +
         xi * (0.08333333333333333f - x2 * (0.002777777777777778f - x2 * (7.936507936507937e-4f - x2 * 5.952380952380953e-4f))) -
         0.6931471805599453f * hm_lg2(prod);
    ps = lx - 0.5f * xi -
         x2 * (0.08333333333333333f - x2 * (0.008333333333333333f - x2 * (0.003968253968253968f - x2 * (0.004166666666666667f - x2 * 0.007575757575757576f)))) -
         rec;
    tr = xi * (1.f + 0.5f * xi +
               x2 * (0.16666666666666666f - x2 * (0.03333333333333333f - x2 * (0.023809523809523808f - x2 * (0.03333333333333333f - x2 * 0.07575757575757576f))))) +
         rec2;
}

// Categorical, D = K - 1 latent functions on a 10^D tensor grid (categorical.py:130-222).
// With den = 1 + sum_d e^{f_d}:  p_d = e^{f_d}/den, p_K = 1/den.  When no probability on the whole grid reaches the
// reference's clip [1e-9, 1-1e-9] (checked per row from the extreme nodes), the clip and the renormalisation
// p / p.sum() are the identity up to round-off, so
//     log p_y = f_y - log den (y <= D) | -log den (y = K),   E[f_y] = m_y exactly (sum w = 1, sum w x = 0),
//     E[log p_y] = m_y [y <= D] - sum_g w_g log den_g,       d2logp_df2_d = -rho_d (1 - rho_d), rho_d = e^{f_d}/den.
// Rows where a clip could be active take the literal path (hm_lik_eval<float>).
template <int D>
__device__ __forceinline__ void hm_categorical_fast(int K, float y, const float* m, const float* v, bool want_grads,
                                                    HmLikOut<float>& o) {
    static_assert(D >= 1 && D <= HM_MAXF, "categorical fast path: 1..4 latent functions");
    float e[D][10], emax[D], emin[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float s2v = sqrtf(2.f * v[d]);
#pragma unroll
        for (int i = 0; i < 10; ++i) e[d][i] = hm_safe_exp((float)c_gh10_x[i] * s2v + m[d]);
        emax[d] = fmaxf(e[d][0], e[d][9]);
        emin[d] = fminf(e[d][0], e[d][9]);
    }
    float den_max = 1.f, pmin = 1.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { den_max += emax[d]; pmin = fminf(pmin, emin[d]); }
    const bool finite_ok = den_max < 1.0e30f && !(v[0] != v[0]);
    bool no_clip = finite_ok && (pmin >= 1.0e-9f * den_max) && (den_max <= 1.0e9f);   // min p >= 1e-9 over the grid
#pragma unroll
    for (int d = 0; d < D; ++d) no_clip = no_clip && (v[d] >= 0.f);
    if (!no_clip) { hm_lik_eval<float>(HMOGP_LIK_CATEGORICAL, K, 0.f, y, m, v, o); return; }
    const int label = (int)y;
    const bool valid = ((float)label == y) && label >= 1 && label <= K;
    float w10[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) w10[i] = (float)c_gh10_w[i];
    // nested loops, last function fastest (C-order of categorical.py:153-157; the sums do not depend on the order)
    float S = 0.f, t[D];
#pragma unroll
    for (int d = 0; d < D; ++d) t[d] = 0.f;
    constexpr int NOUT = (D == 1) ? 1 : (D == 2 ? 10 : (D == 3 ? 100 : 1000));
    for (int g = 0; g < NOUT; ++g) {
        // outer axes 0..D-2 (dynamic), inner axis D-1 unrolled
        float base = 1.f, wo = 1.f, eo[D > 1 ? D - 1 : 1];
        int r = g;
#pragma unroll
        for (int d = D - 2; d >= 0; --d) {
            const int idx = r % 10; r /= 10;
            float ev = e[d][0], wv = w10[0];
#pragma unroll
            for (int i = 1; i < 10; ++i) { ev = (idx == i) ? e[d][i] : ev; wv = (idx == i) ? w10[i] : wv; }
            eo[d] = ev; base += ev; wo *= wv;
        }
        // inner axis D-1.  For the outer axes rho_d = e_d / den with e_d fixed along the inner axis, so
        //   sum_k w_k rho_d (1 - rho_d) = e_d (I1 - e_d I2),   I1 = sum_k w_k / den_k,  I2 = sum_k w_k / den_k^2
        float s_in = 0.f, t_last = 0.f, I1 = 0.f, I2 = 0.f;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const float den = base + e[D - 1][k];
            s_in = fmaf(w10[k], hm_lg2(den), s_in);
            if (want_grads) {
                const float inv = hm_rcp(den), winv = w10[k] * inv;
                I1 += winv;
                I2 = fmaf(winv, inv, I2);
                const float rho = e[D - 1][k] * inv;
                t_last = fmaf(w10[k], fmaf(-rho, rho, rho), t_last);
            }
        }
        float t_in[D];
#pragma unroll
        for (int d = 0; d < D - 1; ++d) t_in[d] = eo[d] * fmaf(-eo[d], I2, I1);
        t_in[D - 1] = t_last;
        S = fmaf(wo, s_in, S);
#pragma unroll
        for (int d = 0; d < D; ++d) t[d] = fmaf(wo, t_in[d], t[d]);
    }
    const float ve = ((label >= 1 && label <= D) ? m[label - 1] : 0.f) - 0.6931471805599453f * S;
    o.ve = valid ? ve : -CUDART_INF_F;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        o.dm[d] = valid ? ((label == d + 1 ? 1.f : 0.f) - 1.f) : 0.f;   // quirk C-1 (weights sum to 1)
        o.dv[d] = valid ? -0.5f * t[d] : 0.f;
    }
}

// Gamma / Beta on the 10 x 10 grid (gamma.py:103-194, beta.py:106-197), special functions through hm_lgam_psi_tri.
template <bool IS_GAMMA>
__device__ __forceinline__ void hm_gamma_beta_fast(float y, const float* m, const float* v, HmLikOut<float>& o) {
    const float inv_pi = 0.31830988618379067f;   // quirk C-2
    const float sa = sqrtf(2.f * v[0]), sb = sqrtf(2.f * v[1]);
    float a[10], b[10], lga[10], psa[10], tra[10], lgb[10], psb[10], trb[10], w10[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        w10[i] = (float)c_gh10_w[i];
        a[i] = hm_clip(hm_safe_exp((float)c_gh10_x[i] * sa + m[0]), 1e-9f, 1e9f);
        b[i] = hm_clip(hm_safe_exp((float)c_gh10_x[i] * sb + m[1]), 1e-9f, 1e9f);
        hm_lgam_psi_tri(a[i], lga[i], psa[i], tra[i]);
        if (IS_GAMMA) { lgb[i] = logf(b[i]); psb[i] = 0.f; trb[i] = 0.f; }
        else hm_lgam_psi_tri(b[i], lgb[i], psb[i], trb[i]);
    }
    const float logy = logf(y), log1y = IS_GAMMA ? 0.f : logf(1.f - y);
    float ve = 0.f, da = 0.f, db = 0.f, d2a = 0.f, d2b = 0.f;
    if (IS_GAMMA) {
        // every term is separable in (i, j): with sum_j w_j = W1, sum_j w_j log b_j = LB, sum_j w_j b_j = B1
        float W1 = 0.f, LB = 0.f, B1 = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) { W1 += w10[j]; LB = fmaf(w10[j], lgb[j], LB); B1 = fmaf(w10[j], b[j], B1); }
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const float wi = w10[i], ai = a[i];
            ve = fmaf(wi, (-lga[i] + (ai - 1.f) * logy) * W1 + ai * LB - y * B1, ve);
            da = fmaf(wi, ((-psa[i] + logy) * W1 + LB) * ai, da);
            db = fmaf(wi, ai * W1 - y * B1, db);
            d2a = fmaf(wi, ((-psa[i] - ai * tra[i] + logy) * W1 + LB) * ai, d2a);
            d2b = fmaf(wi, -y * B1, d2b);
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < 10; ++i) {
            float rve = 0.f, rda = 0.f, rdb = 0.f, r2a = 0.f, r2b = 0.f;
            const float ai = a[i];
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                float lgab, psab, trab;
                hm_lgam_psi_tri(ai + b[j], lgab, psab, trab);
                const float wj = w10[j], bj = b[j];
                rve = fmaf(wj, (ai - 1.f) * logy + (bj - 1.f) * log1y - (lga[i] + lgb[j] - lgab), rve);
                rda = fmaf(wj, (psab - psa[i] + logy) * ai, rda);
                rdb = fmaf(wj, (psab - psb[j] + log1y) * bj, rdb);
                r2a = fmaf(wj, (psab + ai * trab - psa[i] - ai * tra[i] + logy) * ai, r2a);
                r2b = fmaf(wj, (psab + bj * trab - psb[j] - bj * trb[j] + log1y) * bj, r2b);
            }
            ve = fmaf(w10[i], rve, ve); da = fmaf(w10[i], rda, da); db = fmaf(w10[i], rdb, db);
            d2a = fmaf(w10[i], r2a, d2a); d2b = fmaf(w10[i], r2b, d2b);
        }
    }
    o.ve = ve * inv_pi;
    o.dm[0] = da * inv_pi;
    o.dm[1] = db * inv_pi;
    o.dv[0] = 0.5f * d2a * inv_pi;
    o.dv[1] = 0.5f * d2b * inv_pi;
}

// compile-time dispatch used by the fp32 row kernel (KIND = likelihood kind, D = Categorical latent functions)
template <int KIND, int D>
__device__ __forceinline__ void hm_lik_eval_f32(int K, float sigma, float y, const float* m, const float* v, bool want_grads,
                                                HmLikOut<float>& o) {
    if constexpr (KIND == HMOGP_LIK_CATEGORICAL) hm_categorical_fast<D>(K, y, m, v, want_grads, o);
    else if constexpr (KIND == HMOGP_LIK_GAMMA) hm_gamma_beta_fast<true>(y, m, v, o);
    else if constexpr (KIND == HMOGP_LIK_BETA) hm_gamma_beta_fast<false>(y, m, v, o);
    else hm_lik_eval<float>(KIND, K, sigma, y, m, v, o);
}
