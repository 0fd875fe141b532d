/*
 * hetmogp_b200 -- C-ABI of the B200-native HetMOGP ELBO/gradient engine.
 *
 * Drop-in boundary for ONE path of pmorenoz/HetMOGP: the evaluation that
 * SVMOGP.parameters_changed() (hetmogp/svmogp.py:85-166) triggers, i.e.
 * SVMOGPInf.inference (hetmogp/svmogp_inf.py:23-109) plus the hyper-parameter
 * chain rule of svmogp.py:100-166, and the per-likelihood var_exp plug-ins
 * (likelihoods/<name>.py).  The reference is pure Python; a maintainer binds
 * this library with ctypes.  Plain pointers and sizes
 * only -- no torch / numpy types in any signature.
 *
 * All matrices are row-major (C order) float64 exactly as the reference's
 * numpy arrays are laid out.  Pointers are either ALL host or ALL device
 * pointers per call, selected by `mem_kind`.
 *
 * Every function returns 0 on success, non-zero on error (hmogp_last_error()
 * gives the message).  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with HMOGP_ERR_CUDA.
 */
#ifndef HETMOGP_B200_H
#define HETMOGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMOGP_ABI_VERSION 1

/* error codes */
#define HMOGP_OK 0
#define HMOGP_ERR_ARG 1
#define HMOGP_ERR_CUDA 2
#define HMOGP_ERR_LINALG 3   /* jitchol failed after 5 jitter tries (GPy LinAlgError, util.py:198) */
#define HMOGP_ERR_UNSTABLE 4 /* inf in S_q^-1: "Sqi: Cholesky representation unstable" (svmogp_inf.py:126-127) */

/* memory kinds */
#define HMOGP_MEM_HOST 0
#define HMOGP_MEM_DEVICE 1

/* likelihood kinds (likelihoods/<name>.py) */
#define HMOGP_LIK_GAUSSIAN 0    /* gaussian.py:17-62      dim_f=1, analytic        */
#define HMOGP_LIK_HETGAUSSIAN 1 /* hetgaussian.py:46-73   dim_f=2, analytic        */
#define HMOGP_LIK_BERNOULLI 2   /* bernoulli.py:31-111    dim_f=1, GH-20           */
#define HMOGP_LIK_POISSON 3     /* poisson.py:31-95       dim_f=1, GH-20           */
#define HMOGP_LIK_CATEGORICAL 4 /* categorical.py:37-222  dim_f=K-1, GH-10^(K-1)   */
#define HMOGP_LIK_GAMMA 5       /* gamma.py:34-194        dim_f=2, GH-10x10, x1/pi */
#define HMOGP_LIK_BETA 6        /* beta.py:29-197         dim_f=2, GH-10x10, x1/pi */
#define HMOGP_LIK_EXPONENTIAL 7 /* exponential.py:28-99   dim_f=1, GH-20           */
#define HMOGP_MAX_DIMF 4        /* Categorical up to K=5 */
#define HMOGP_MAX_Q 8
#define HMOGP_MAX_TASKS 16
#define HMOGP_MAX_J 32

/* arithmetic of the N-sized kernels (the M x M algebra is always fp64) */
#define HMOGP_PREC_FP64 0 /* SIMT fp64 everywhere: reference-grade parity mode            */
#define HMOGP_PREC_FP32 1 /* SIMT fp32 tiles, fp64 reductions                              */
#define HMOGP_PREC_TC 2   /* tcgen05 tensor-core tiles (split-fp16 operands, fp32 TMEM acc) */

/* what a step computes */
#define HMOGP_WHAT_ELBO 0 /* ELBO only                                                  */
#define HMOGP_WHAT_VE 1   /* + dL_dmu_u, dL_dL_u   (the VE step of svmogp.py:104-113)   */
#define HMOGP_WHAT_FULL 2 /* + kernel, W, kappa, Z gradients (parameters_changed)       */

typedef struct {
    int32_t kind;  /* HMOGP_LIK_* */
    int32_t K;     /* Categorical: number of classes (categorical.py:22) */
    double sigma;  /* Gaussian: noise std (gaussian.py:22, default 0.5) */
} hmogp_lik_desc;

typedef struct {
    int32_t M;         /* inducing points per latent                 */
    int32_t Q;         /* latent GPs                                 */
    int32_t Xdim;      /* input dimension                            */
    int32_t T;         /* tasks/outputs = len(Y) (svmogp_inf.py:26)  */
    int32_t precision; /* HMOGP_PREC_*                               */
    int32_t device;    /* CUDA device ordinal                        */
    const hmogp_lik_desc* liks; /* [T] */
} hmogp_config;

/* Parameters of one evaluation: the arguments of SVMOGPInf.inference
 * (svmogp_inf.py:23-24) with kern_list / B_list flattened to arrays. */
typedef struct {
    const double* Z;           /* [M, Q*Xdim]  column block q = inducing inputs of latent q (util.py:197) */
    const double* m_u;         /* [M, Q]       q_u_means (svmogp.py:66)                                   */
    const double* L_u;         /* [M(M+1)/2,Q] q_u_chols, packed lower, row-major order (svmogp.py:68)    */
    const double* rbf_var;     /* [Q]          kern_list[q].variance                                      */
    const double* rbf_ls;      /* [Q]          kern_list[q].lengthscale                                   */
    const double* W;           /* [J, Q]       B_list[q].W[:,0]                                           */
    const double* kappa;       /* [J, Q]       B_list[q].kappa                                            */
    const double* W_chain;     /* [J, Q] or NULL: multipliers of svmogp.py:141,143,156 (quirk C-5); NULL = W     */
    const double* kappa_chain; /* [J, Q] or NULL                                                          */
    const double* batch_scale; /* [T] or NULL (=1)   svmogp.py:89-90                                      */
} hmogp_params;

/* Outputs; any pointer may be NULL to skip that output. */
typedef struct {
    double* log_marginal; /* [1]            svmogp_inf.py:84-88                            */
    double* VE;           /* [T]            per-task sum of scaled var_exp                 */
    double* KL;           /* [1]            svmogp_inf.py:227-250                          */
    double* dL_dmu_u;     /* [M, Q]         svmogp_inf.py:168                              */
    double* dL_dL_u;      /* [M(M+1)/2, Q]  svmogp_inf.py:175-178                          */
    double* dL_dKmm;      /* [Q, M, M]      svmogp_inf.py:170                              */
    double* d_rbf;        /* [Q, 2]         (variance, lengthscale) svmogp.py:116,139-143  */
    double* dW;           /* [J, Q]         util.py:228-255, svmogp.py:120-129             */
    double* dkappa;       /* [J, Q]                                                        */
    double* dZ;           /* [M, Q*Xdim]    svmogp.py:153-156                              */
} hmogp_grads;

typedef struct {
    int32_t chol_fail[HMOGP_MAX_Q];  /* failed Cholesky attempts of K_uu^q (jitchol retries; 0 = PD as is) */
    double jitter[HMOGP_MAX_Q];      /* jitter finally added to K_uu^q (0 = none; jitchol)     */
    int32_t lu_singular[HMOGP_MAX_Q];/* 1 if S_q^-1 contains inf (svmogp_inf.py:126)           */
    int64_t n_negative_v;            /* rows with v_fd < 0 ('v negative!', svmogp_inf.py:221)  */
} hmogp_status;

typedef struct hmogp_engine hmogp_engine;

int hmogp_abi_version(void);
const char* hmogp_last_error(void);
int hmogp_device_count(void);

/* ---- metadata: HetLikelihood.generate_metadata (het_likelihood.py:24-44), bit-exact integers ---- */
int hmogp_lik_dims(const hmogp_lik_desc* lik, int32_t* dim_y, int32_t* dim_f, int32_t* dim_p);
/* out arrays sized by the caller: y_index[sum dim_y], function_index[J], d_index[J], pred_index[sum dim_p] */
int hmogp_generate_metadata(int32_t T, const hmogp_lik_desc* liks, int64_t* task_index, int64_t* y_index,
                            int64_t* function_index, int64_t* d_index, int64_t* pred_index, int32_t* n_y,
                            int32_t* n_f, int32_t* n_p);

/* ---- engine life cycle ---- */
int hmogp_create(const hmogp_config* cfg, hmogp_engine** out);
void hmogp_destroy(hmogp_engine* e);
int hmogp_set_stream(hmogp_engine* e, void* cuda_stream);
/* Copy one task's rows to the device (resident shard).  X [N,Xdim], Y [N] (labels stored as floats).
 * HMOGP_MEM_HOST with pinned (page-locked) buffers: the upload is queued by the next step call, behind its parameter
 * copies, and overlaps the M-sized prepare phase; the buffers must stay valid until that step has run.  Pageable host
 * memory is copied inside this call. */
int hmogp_set_data(hmogp_engine* e, int32_t t, const double* X, const double* Y, int64_t N, int32_t mem_kind);
/* Restrict the data term to rows [begin[t], begin[t]+count[t]) of each task (minibatch slice,
 * util.py:52-72 / per-rank shard); NULL = all rows. */
int hmogp_set_rows(hmogp_engine* e, const int64_t* begin, const int64_t* count);

/* ---- the hot path: one evaluation equivalent to parameters_changed() ---- */
int hmogp_elbo_and_grads(hmogp_engine* e, const hmogp_params* p, hmogp_grads* g, int32_t mem_kind, int32_t what,
                         hmogp_status* status);
/* Same, split for data-parallel ranks: step_local leaves the per-shard sufficient statistics in a packed
 * fp64 device buffer (hmogp_stats_len doubles); the caller sum-all-reduces it (one NCCL all-reduce) and
 * calls step_finish.  `stats` may be NULL to use the engine's own buffer (see hmogp_stats_ptr). */
int64_t hmogp_stats_len(const hmogp_engine* e);
double* hmogp_stats_ptr(hmogp_engine* e);
int hmogp_step_local(hmogp_engine* e, const hmogp_params* p, int32_t mem_kind, int32_t what, double* stats_dev);
int hmogp_step_finish(hmogp_engine* e, const double* stats_dev, hmogp_grads* g, int32_t mem_kind, int32_t what,
                      hmogp_status* status);

/* Stateless convenience = SVMOGPInf.inference with host arrays (svmogp_inf.py:23): uploads X/Y of every
 * task, runs one evaluation, returns host outputs.  X[t] [N[t],Xdim], Y[t] [N[t]]. */
int hmogp_inference_host(const hmogp_config* cfg, const double* const* X, const double* const* Y, const int64_t* N,
                         const hmogp_params* p, hmogp_grads* g, int32_t what, hmogp_status* status);

/* ---- prediction: q(f_d) at new inputs from q(U) (the O(M^2)-per-point route to what svmogp.py:263-370 obtains through
 *      inference(..., predictive=True), svmogp_inf.py:43-50,216-218): m_fd, v_fd [N, dim_f(t)] for the output functions
 *      of task t at Xnew [N, Xdim].  No labels are needed and no likelihood is evaluated.  Pointers per mem_kind. ---- */
int hmogp_predict_f(hmogp_engine* e, const hmogp_params* p, int32_t mem_kind, int32_t t, const double* Xnew, int64_t N,
                    double* m_fd, double* v_fd);

/* ---- per-row intermediates of the last evaluation (tests; small N) ---- */
/* m_fd, v_fd: [N_t, dim_f];  VE: [N_t];  dm, dv: [N_t, dim_f]  (svmogp_inf.py:54-78). Host pointers. */
int hmogp_get_rows(hmogp_engine* e, int32_t t, double* m_fd, double* v_fd, double* VE, double* dm, double* dv);
/* Dense N-sized gradient blocks of svmogp_inf.py:157-164 for small N (host pointers):
 * dL_dKmn [M, N_t] and dL_dKdiag [N_t] for latent q, output function d. */
int hmogp_get_dL_dKmn(hmogp_engine* e, int32_t q, int32_t d, double* dL_dKmn, double* dL_dKdiag);
/* M x M factors of util.latent_funs_cov (util.py:181-200): each [Q, M, M], host pointers, NULL to skip. */
int hmogp_get_kuu(hmogp_engine* e, double* Kuu, double* Luu, double* Kuui);

/* ---- likelihood plug-in point (likelihoods/<name>.py var_exp / var_exp_derivatives) ---- */
/* Y [N], Mf/Vf [N,dim_f] -> VE [N], dm/dv [N,dim_f] (any output may be NULL). */
int hmogp_lik_var_exp(const hmogp_lik_desc* lik, int64_t N, const double* Y, const double* Mf, const double* Vf,
                      double* VE, double* dm, double* dv, int32_t precision, int32_t mem_kind, void* cuda_stream);
/* predictive(m, v) of likelihoods/<name>.py (e.g. bernoulli.py:113-128, gamma.py:196-238, categorical.py:224-269): mean and
 * variance of p(y*) under q(f*) = N(Mf, diag Vf), Gauss-Hermite.  Mf/Vf [N,dim_f] -> mean_pred/var_pred [N,dim_p].
 * gh_tensor: nodes per axis of the tensor grids (Gamma, Beta, Categorical): 10 = an instance that has run var_exp
 * (a trained model), 20 = a fresh instance (GPy caches the first table it builds). */
int hmogp_lik_predictive(const hmogp_lik_desc* lik, int64_t N, const double* Mf, const double* Vf, double* mean_pred,
                         double* var_pred, int32_t gh_tensor, int32_t mem_kind, void* cuda_stream);
/* logpdf / dlogp_df / d2logp_df2 at given F [N,dim_f]: logp [N], dlogp/d2logp [N,dim_f]. */
int hmogp_lik_pointwise(const hmogp_lik_desc* lik, int64_t N, const double* F, const double* Y, double* logp,
                        double* dlogp, double* d2logp, int32_t mem_kind, void* cuda_stream);

/* ---- packed-lower <-> dense index kernels (GPy choleskies.flat_to_triang / triang_to_flat;
 *      svmogp_inf.py:118,176-178).  flat [M(M+1)/2, D], dense [D, M, M]. ---- */
int hmogp_flat_to_triang(const double* flat, double* dense, int32_t M, int32_t D, int32_t mem_kind, void* cuda_stream);
int hmogp_triang_to_flat(const double* dense, double* flat, int32_t M, int32_t D, int32_t mem_kind, void* cuda_stream);

/* ---- on-device optimiser step of the stochastic loop: climin.Adadelta(model.optimizer_array, model.stochastic_grad,
 *      step_rate, momentum=0.9) of util.py:320-329 over paramz' flat optimizer_array (link order svmogp.py:71-75, Logexp
 *      transform of the positive parameters), with the VE / VM gradient gating of svmogp.py:104-166.  Parameters and
 *      gradients are the device arrays a step reads / writes (hmogp_params / hmogp_grads with HMOGP_MEM_DEVICE), so a
 *      training iteration never leaves the GPU. ---- */
#define HMOGP_OPT_MAX_SEGMENTS 64
typedef struct {
    int64_t offset;      /* first position in the flat vector (segments are contiguous, in link order)      */
    int64_t count;       /* elements                                                                         */
    double* param;       /* device array holding the (constrained) parameter; element k at param[k*stride]   */
    const double* grad;  /* device array holding dELBO/dparam, same indexing                                 */
    int32_t stride;      /* 1, or Q for the columns of W / kappa [J, Q]                                      */
    int32_t positive;    /* 1: Logexp-transformed (variance, lengthscale, kappa)                             */
    int32_t variational; /* 1: m_u / L_u (gradient live in VE steps); 0: Z, kernels, B (live in VM steps)    */
    int32_t reserved;
} hmogp_opt_segment;
typedef struct hmogp_opt hmogp_opt;
int hmogp_opt_create(int32_t device, const hmogp_opt_segment* segs, int32_t nseg, double step_rate, double decay,
                     double momentum, double offset, hmogp_opt** out);
void hmogp_opt_destroy(hmogp_opt* o);
int64_t hmogp_opt_size(const hmogp_opt* o);
/* device pointers of the state vectors: 0 wrt (the flat optimizer_array), 1 gms, 2 sms, 3 step */
double* hmogp_opt_state(hmogp_opt* o, int32_t which);
/* copy a state vector ([hmogp_opt_size] doubles) to host memory */
int hmogp_opt_get_state(hmogp_opt* o, int32_t which, double* host_out, void* cuda_stream);
/* wrt <- unconstrained(parameters): model.optimizer_array */
int hmogp_opt_gather(hmogp_opt* o, void* cuda_stream);
/* wrt -= momentum * step (if apply_momentum), parameters <- constrained(wrt): model.optimizer_array = wrt */
int hmogp_opt_lookahead(hmogp_opt* o, int32_t apply_momentum, void* cuda_stream);
/* g = -transformed gradient (zero where gated off), then the Adadelta update; grad_out [n] receives g (or NULL) */
int hmogp_opt_update(hmogp_opt* o, int32_t ve_active, int32_t vm_active, double* grad_out, void* cuda_stream);

/* K_uu, its Cholesky factor and the two inverses (util.latent_funs_cov, util.py:181-200) stay resident between
 * evaluations and are reused when Z, rbf_var and rbf_ls are bitwise those they were built from: the VE phases of VEM
 * (svmogp.py:104-113 with the hyper-parameters fixed, util.py:284-331) only move m_u and L_u.  With HMOGP_MEM_HOST
 * parameters the engine compares the values itself; with HMOGP_MEM_DEVICE parameters it cannot see them, and reuses the
 * factorisation for the NEXT evaluation only if the caller says so here.  A factorisation that needed jitter is never
 * reused.  hmogp_kuu_reuse_count: evaluations of this engine that reused it (test / diagnostic).  HMOGP_NO_KUU_CACHE=1
 * in the environment turns the reuse off. */
int hmogp_hint_hyper_unchanged(hmogp_engine* e, int32_t unchanged);
int64_t hmogp_kuu_reuse_count(hmogp_engine* e);

/* ---- timing hooks for bench.py: CUDA-event time (ms) and launch count of the N-sized kernels of the
 *      last evaluation, measured on the engine's stream. ---- */
int hmogp_enable_timing(hmogp_engine* e, int32_t on);
int hmogp_last_timing(hmogp_engine* e, float* ms /* [6]: prepare, forward, lik, bwd_proj, bwd_gram, finish */,
                      int32_t* launches);
/* 1 if the tcgen05 tensor-core path (HMOGP_PREC_TC) is compiled into this library */
int hmogp_tc_built(void);

#ifdef __cplusplus
}
#endif
#endif /* HETMOGP_B200_H */
