// On-device optimiser step of the stochastic loop (reference: /root/reference/hetmogp/util.py:320-329,
// climin.Adadelta(model.optimizer_array, model.stochastic_grad, step_rate, momentum=0.9) driving
// SVMOGP.stochastic_grad, /root/reference/hetmogp/svmogp.py:188-199, through paramz' optimizer_array / _grads).
//
// The flat optimiser vector follows paramz' link order (svmogp.py:71-75): Z, m_u, L_u, [variance_q, lengthscale_q] for
// every kernel, [W_q, kappa_q] for every coregionalisation matrix; fixed parameters are left out; positive parameters
// (variance, lengthscale, kappa) are carried through paramz' Logexp transform theta = log(1 + e^x), gradient factor
// 1 - e^-theta.  A segment table maps flat positions onto the engine's parameter / gradient arrays (hmogp_params /
// hmogp_grads layouts, strided for W and kappa whose latent index is the fastest one there).
//
// climin 0.1a1 Adadelta._iterate (third-party, absent from /root/reference; restated in oracle/climin_adadelta.py):
//     step1 = momentum * step;  wrt -= step1                      -> hmogp_opt_lookahead  (+ scatter into the parameters)
//     g = fprime(wrt)                                              -> one engine evaluation, gradients stay on the device
//     gms = d gms + (1 - d) g^2
//     step2 = sqrt(sms + o) / sqrt(gms + o) * g * step_rate;  wrt -= step2
//     step = step1 + step2;  sms = d sms + (1 - d) step^2         -> hmogp_opt_update     (gather + gate + update)
// Compiled with -fmad=false: every product and sum rounds separately, as numpy does, so the update is bit-identical to
// the oracle in fp64 (IEEE sqrt and division).
#include <math.h>
#include <string.h>

#include "common.cuh"

struct hmogp_opt {
    int64_t n;
    int nseg;
    double step_rate, decay, momentum, offset;
    double *wrt, *gms, *sms, *step;
    hmogp_opt_segment* segs_d;
    hmogp_opt_segment segs_h[HMOGP_OPT_MAX_SEGMENTS];
    int device;
};

namespace {

__device__ __forceinline__ int find_segment(const hmogp_opt_segment* segs, int nseg, int64_t i) {
    int s = 0;
    while (s + 1 < nseg && segs[s + 1].offset <= i) ++s;
    return s;
}

// paramz Logexp.f: theta = where(x > 36, x, log1p(exp(clip(x, -log(DBL_MAX), 36))))   (paramz/transformations.py, recalled)
__device__ __forceinline__ double logexp_f(double x) {
    if (x > 36.0) return x;
    return log1p(exp(fmax(x, -709.782712893384)));
}

// wrt -= momentum * step; parameters <- constrained(wrt)
__global__ void opt_lookahead_kernel(hmogp_opt_segment* segs, int nseg, int64_t n, double momentum, double* wrt,
                                     const double* step, int apply_momentum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double w = wrt[i];
    if (apply_momentum) {
        const double step1 = step[i] * momentum;
        w = w - step1;
        wrt[i] = w;
    }
    const hmogp_opt_segment sg = segs[find_segment(segs, nseg, i)];
    sg.param[(i - sg.offset) * sg.stride] = sg.positive ? logexp_f(w) : w;
}

// g = -(dELBO/dtheta) * transform factor, gated; Adadelta update of wrt and its state
__global__ void opt_update_kernel(const hmogp_opt_segment* segs, int nseg, int64_t n, double step_rate, double decay,
                                  double momentum, double offset, double* wrt, double* gms, double* sms, double* step,
                                  int ve_active, int vm_active, double* grad_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const hmogp_opt_segment sg = segs[find_segment(segs, nseg, i)];
    const int64_t k = (i - sg.offset) * sg.stride;
    const bool on = sg.variational ? (ve_active != 0) : (vm_active != 0);
    double g = 0.0;
    if (on) {
        g = sg.grad[k];
        if (sg.positive) {                                     // paramz Logexp.gradfactor: df * where(f > 36, 1, -expm1(-f))
            const double th = sg.param[k];
            g = g * (th > 36.0 ? 1.0 : -expm1(-th));
        }
        g = -g;                                                // objective = -ELBO (paramz Model._grads)
    }
    if (grad_out) grad_out[i] = g;
    const double step1 = step[i] * momentum;                   // the look-ahead already applied to wrt
    const double gm = decay * gms[i] + (1.0 - decay) * (g * g);
    const double step2 = sqrt(sms[i] + offset) / sqrt(gm + offset) * g * step_rate;
    wrt[i] = wrt[i] - step2;
    const double st = step1 + step2;
    gms[i] = gm;
    step[i] = st;
    sms[i] = decay * sms[i] + (1.0 - decay) * (st * st);
}

// wrt <- unconstrained(parameters)   (paramz Logexp.finv: x = log(e^theta - 1), theta for large theta)
__global__ void opt_gather_kernel(const hmogp_opt_segment* segs, int nseg, int64_t n, double* wrt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const hmogp_opt_segment sg = segs[find_segment(segs, nseg, i)];
    const double th = sg.param[(i - sg.offset) * sg.stride];
    wrt[i] = sg.positive ? (th > 36.0 ? th : log(expm1(th))) : th;
}

}  // namespace

extern "C" {

int hmogp_opt_create(int32_t device, const hmogp_opt_segment* segs, int32_t nseg, double step_rate, double decay,
                     double momentum, double offset, hmogp_opt** out) {
    if (!segs || !out || nseg < 1 || nseg > HMOGP_OPT_MAX_SEGMENTS) { hm_set_error("hmogp_opt_create: bad segment table"); return HMOGP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { hm_set_error("no CUDA device: hetmogp_b200 has no CPU fallback"); return HMOGP_ERR_CUDA; }
    HM_CUDA(cudaSetDevice(device));
    int64_t n = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segs[s].offset != n || segs[s].count < 1 || !segs[s].param || segs[s].stride < 1) { hm_set_error("hmogp_opt_create: segment %d is not contiguous / valid", s); return HMOGP_ERR_ARG; }
        n += segs[s].count;
    }
    hmogp_opt* o = new hmogp_opt();
    memset(o, 0, sizeof(*o));
    o->n = n; o->nseg = nseg; o->step_rate = step_rate; o->decay = decay; o->momentum = momentum; o->offset = offset; o->device = device;
    memcpy(o->segs_h, segs, sizeof(hmogp_opt_segment) * nseg);
    cudaError_t e1 = cudaMalloc((void**)&o->wrt, sizeof(double) * 4 * n);
    cudaError_t e2 = cudaMalloc((void**)&o->segs_d, sizeof(hmogp_opt_segment) * nseg);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { hm_set_error("hmogp_opt_create: cudaMalloc failed"); hmogp_opt_destroy(o); return HMOGP_ERR_CUDA; }
    o->gms = o->wrt + n; o->sms = o->gms + n; o->step = o->sms + n;
    HM_CUDA(cudaMemset(o->wrt, 0, sizeof(double) * 4 * n));
    HM_CUDA(cudaMemcpy(o->segs_d, segs, sizeof(hmogp_opt_segment) * nseg, cudaMemcpyHostToDevice));
    *out = o;
    return 0;
}

void hmogp_opt_destroy(hmogp_opt* o) {
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->wrt) cudaFree(o->wrt);
    if (o->segs_d) cudaFree(o->segs_d);
    delete o;
}

int64_t hmogp_opt_size(const hmogp_opt* o) { return o ? o->n : 0; }
double* hmogp_opt_state(hmogp_opt* o, int32_t which) {
    if (!o) return nullptr;
    switch (which) { case 0: return o->wrt; case 1: return o->gms; case 2: return o->sms; case 3: return o->step; }
    return nullptr;
}

int hmogp_opt_get_state(hmogp_opt* o, int32_t which, double* host_out, void* cuda_stream) {
    double* src = hmogp_opt_state(o, which);
    if (!src || !host_out) { hm_set_error("hmogp_opt_get_state: bad argument"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(o->device));
    HM_CUDA(cudaMemcpyAsync(host_out, src, sizeof(double) * o->n, cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream));
    HM_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    return 0;
}

int hmogp_opt_gather(hmogp_opt* o, void* cuda_stream) {
    if (!o) { hm_set_error("null optimiser"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(o->device));
    opt_gather_kernel<<<(unsigned)hm_cdiv(o->n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(o->segs_d, o->nseg, o->n, o->wrt);
    HM_CUDA(cudaGetLastError());
    return 0;
}

int hmogp_opt_lookahead(hmogp_opt* o, int32_t apply_momentum, void* cuda_stream) {
    if (!o) { hm_set_error("null optimiser"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(o->device));
    opt_lookahead_kernel<<<(unsigned)hm_cdiv(o->n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(o->segs_d, o->nseg, o->n, o->momentum, o->wrt,
                                                                                             o->step, apply_momentum);
    HM_CUDA(cudaGetLastError());
    return 0;
}

int hmogp_opt_update(hmogp_opt* o, int32_t ve_active, int32_t vm_active, double* grad_out, void* cuda_stream) {
    if (!o) { hm_set_error("null optimiser"); return HMOGP_ERR_ARG; }
    HM_CUDA(cudaSetDevice(o->device));
    for (int s = 0; s < o->nseg; ++s)
        if (!o->segs_h[s].grad) { hm_set_error("hmogp_opt_update: segment %d has no gradient array", s); return HMOGP_ERR_ARG; }
    opt_update_kernel<<<(unsigned)hm_cdiv(o->n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(o->segs_d, o->nseg, o->n, o->step_rate, o->decay,
                                                                                          o->momentum, o->offset, o->wrt, o->gms, o->sms,
                                                                                          o->step, ve_active, vm_active, grad_out);
    HM_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
