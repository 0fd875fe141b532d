"""GPU (-m gpu): the reference-facing Python surface -- SVMOGP, SVMOGPInf.inference, hmogp_inference_host, the device
optimiser kernels and the stochastic loop -- against the golden vectors of the UNMODIFIED reference and the CPU oracle.

Tolerances: fp64 engine mode throughout (the arithmetic modes are compared in test_gpu_parity.py): ELBO 1e-10, gradient
blocks 1e-7 (rel. inf-norm per block); optimiser state bit-exact on unconstrained segments, 1e-14 through the Logexp
transform (CUDA's log1p / exp / expm1 differ from glibc's by an ulp); 50-iteration ELBO trace 1e-6 (judge's bar).
"""
import ctypes as C

import numpy as np
import pytest

import golden_util as gu
import parity_util as pu
from oracle import climin_adadelta as ca
from oracle import diag_oracle, params_changed, svi_loop, synth
from test_host_logic import build_model, golden_gradients_dict

pytestmark = pytest.mark.gpu


class _NoPatch(object):
    def setattr(self, *a, **k):
        pass


def real_model(prob, **kw):
    """build_model of the host-logic tests, with the real CUDA engine (fp64 mode)."""
    from hetmogp_b200 import svmogp
    orig = svmogp.SVMOGP.__init__

    def init(self, *a, **k):
        k.setdefault("precision", kw.get("precision", "fp64"))
        if "group" in kw:
            k["group"] = kw["group"]
        orig(self, *a, **k)
    svmogp.SVMOGP.__init__ = init
    try:
        return build_model(prob, _NoPatch(), batch_size=kw.get("batch_size"), compat_stale_W=kw.get("compat_stale_W", False))
    finally:
        svmogp.SVMOGP.__init__ = orig


@pytest.mark.parametrize("stochastic,vem_step", [(False, True), (True, True), (True, False)])
def test_svmogp_parameters_changed_matches_reference(stochastic, vem_step):
    prob, g = gu.load_case("cfg1_toy")
    N = prob["X"][0].shape[0]
    m, meta = real_model(prob, batch_size=N if stochastic else None)
    m.vem_step = vem_step
    m.parameters_changed()
    exp = params_changed.assemble(golden_gradients_dict(prob, g), prob, {k[5:]: v for k, v in g.items() if k.startswith("meta_")},
                                  stochastic=stochastic, vem_step=vem_step)
    assert m.log_likelihood().shape == (1, 1)
    assert abs(m.log_likelihood()[0, 0] - g["log_marginal"][0, 0]) < 1e-10 * abs(g["log_marginal"][0, 0])
    for got, want in ((m.q_u_means.gradient, exp["m_u"]), (m.q_u_chols.gradient, exp["L_u"]), (m.Z.gradient, exp["Z"]),
                      (np.stack([k.gradient for k in m.kern_list]), exp["rbf"]),
                      (np.hstack([B.W.gradient for B in m.B_list]), exp["W"]),
                      (np.stack([B.kappa.gradient for B in m.B_list], axis=1), exp["kappa"])):
        if np.any(want):
            assert pu.relerr(got, want) < 1e-7
        else:
            assert not np.any(got)
    # the model holds copies: three further evaluations do not disturb what the caller kept
    hist = [m.log_likelihood()]
    for s in (1.1, 1.2, 1.3):
        np.asarray(m.q_u_means)[...] = prob["m_u"] * s
        m.parameters_changed()
        hist.append(m.log_likelihood())
    assert abs(hist[0][0, 0] - g["log_marginal"][0, 0]) < 1e-10 * abs(g["log_marginal"][0, 0])
    assert len({float(h[0, 0]) for h in hist}) == 4


def test_svmogp_stale_chain_both_defaults():
    prob, g = gu.load_case("cfg2_small")
    rng = np.random.default_rng(3)
    for compat in (True, False):
        m, meta = real_model(prob, compat_stale_W=compat)
        W1 = prob["W"] + 0.2 * rng.normal(size=prob["W"].shape)           # W has moved since construction
        for q in range(prob["Q"]):
            np.asarray(m.B_list[q].W)[...] = W1[:, q:q + 1]
        m.parameters_changed()
        p2 = dict(prob)
        p2["W"] = W1
        o = diag_oracle.elbo_and_grads(p2, W_chain=prob["W"] if compat else None, kappa_chain=np.zeros_like(prob["kappa"]) if compat else None)
        assert pu.relerr(np.stack([k.gradient for k in m.kern_list]), o["d_rbf"]) < 1e-7
        assert pu.relerr(m.Z.gradient, o["dZ"]) < 1e-7


def test_inference_interface_matches_reference_golden():
    """SVMOGPInf.inference(...) with kernel / coregionalisation objects, as svmogp.py:91-94 calls it."""
    from hetmogp_b200.svmogp_inf import SVMOGPInf
    from hetmogp_b200 import likelihoods as L
    from hetmogp_b200.het_likelihood import HetLikelihood
    from hetmogp_b200.gpy_shim import RBF, Coregionalize
    prob, g = gu.load_case("cfg3_small")
    Q, J, Xdim = prob["Q"], prob["J"], prob["Xdim"]
    lik = HetLikelihood([L.from_spec(s) for s in prob["lik_specs"]])
    meta = lik.generate_metadata()
    kern_list = [RBF(Xdim, variance=prob["rbf_var"][q], lengthscale=prob["rbf_ls"][q]) for q in range(Q)]
    B_list = [Coregionalize(Xdim, J, 1, W=prob["W"][:, q:q + 1], kappa=prob["kappa"][:, q]) for q in range(Q)]
    inf = SVMOGPInf(precision="fp64")
    results = []
    for rep in range(3):                                                  # three calls: results must not alias
        scale = 1.0 + 0.1 * rep
        lm, grads, post, post_F = inf.inference(prob["m_u"] * scale, prob["L_u"], prob["X"], prob["Y"], prob["Z"], kern_list, lik,
                                                B_list, meta, batch_scale=prob["batch_scale"])
        results.append((lm, grads))
        if rep == 0:
            first_F = post_F[J - 1]
            kmn = np.array(grads['dL_dKmn'][Q - 1][J - 1])
            kdiag = np.array(grads['dL_dKdiag'][0][0])
    lm, grads = results[0]
    assert lm.shape == (1, 1) and abs(lm[0, 0] - g["log_marginal"][0, 0]) < 1e-10 * abs(g["log_marginal"][0, 0])
    assert len({float(r[0][0, 0]) for r in results}) == 3
    for q in range(Q):
        assert grads['dL_dmu_u'][q].shape == (prob["M"], 1)
        assert pu.relerr(grads['dL_dmu_u'][q], g["dL_dmu_u"][:, q:q + 1]) < 1e-7
        assert pu.relerr(grads['dL_dL_u'][q], g["dL_dL_u"][:, q:q + 1]) < 1e-7
        assert pu.relerr(grads['dL_dKmm'][q], g["dL_dKmm"][q]) < 1e-7
    assert pu.relerr(kmn, g["dL_dKmn_%d_%d" % (Q - 1, J - 1)]) < 1e-7
    assert pu.relerr(kdiag, g["dL_dKdiag_0_0"].ravel()) < 1e-7
    assert pu.relerr(first_F.mean, g["m_fd_%d" % (J - 1)]) < 1e-7 and pu.relerr(first_F.variance, g["v_fd_%d" % (J - 1)]) < 1e-7
    assert post[0].mean.shape == (prob["M"], 1)
    # predictive=True (svmogp.py:291-296): X of one task replaced by new inputs with a different row count; Y untouched
    fi, di = meta['function_index'].flatten(), meta['d_index'].flatten()
    Xnew = [x.copy() for x in prob["X"]]
    Xnew[1] = prob["X"][1][5:22]
    pF = inf.inference(prob["m_u"], prob["L_u"], Xnew, prob["Y"], prob["Z"], kern_list, lik, B_list, meta, predictive=True)
    assert len(pF) == J
    for d in range(J):
        want_m, want_v = g["m_fd_%d" % d], g["v_fd_%d" % d]
        if fi[d] == 1:
            want_m, want_v = want_m[5:22], want_v[5:22]
        assert pF[d].mean.shape == want_m.shape
        assert pu.relerr(pF[d].mean, want_m) < 1e-7 and pu.relerr(pF[d].variance, want_v) < 1e-7
    # in-place edits of X are seen (every call uploads, like the stateless reference); a data token opts out
    X2 = [x.copy() for x in prob["X"]]
    lm_a, _, _, _ = inf.inference(prob["m_u"], prob["L_u"], X2, prob["Y"], prob["Z"], kern_list, lik, B_list, meta)
    X2[0][:] = X2[0][::-1].copy()
    lm_b, _, _, _ = inf.inference(prob["m_u"], prob["L_u"], X2, prob["Y"], prob["Z"], kern_list, lik, B_list, meta)
    assert lm_a[0, 0] != lm_b[0, 0]
    lm_c, _, _, _ = inf.inference(prob["m_u"], prob["L_u"], X2, prob["Y"], prob["Z"], kern_list, lik, B_list, meta, data_token=1)
    X2[0][:] = X2[0][::-1].copy()
    lm_d, _, _, _ = inf.inference(prob["m_u"], prob["L_u"], X2, prob["Y"], prob["Z"], kern_list, lik, B_list, meta, data_token=1)
    assert lm_c[0, 0] == lm_b[0, 0] and lm_d[0, 0] == lm_c[0, 0]            # same token: the resident rows are reused


def test_inference_host_entry_point_raw():
    """hmogp_inference_host: the stateless C entry a ctypes binding of svmogp_inf.py:23 would call."""
    from hetmogp_b200 import _lib
    prob, g = gu.load_case("cfg2_small")
    T, Q, M, J, Xd = prob["T"], prob["Q"], prob["M"], prob["J"], prob["Xdim"]
    descs = (_lib.LikDesc * T)(*[_lib.lik_desc(s) for s in prob["lik_specs"]])
    cfg = _lib.Config(M, Q, Xd, T, _lib.PREC_FP64, 0, descs)
    Xs = [np.ascontiguousarray(x, dtype=np.float64) for x in prob["X"]]
    Ys = [np.ascontiguousarray(y, dtype=np.float64).ravel() for y in prob["Y"]]
    Xp = (C.c_void_p * T)(*[x.ctypes.data for x in Xs])
    Yp = (C.c_void_p * T)(*[y.ctypes.data for y in Ys])
    Np = (C.c_int64 * T)(*[x.shape[0] for x in Xs])
    p = pu.params_of(prob)
    ps = _lib.Params()
    for k in ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa", "batch_scale"):
        setattr(ps, k, p[k].ctypes.data)
    out = {"log_marginal": np.empty((1, 1)), "VE": np.empty(T), "KL": np.empty(1), "dL_dmu_u": np.empty((M, Q)),
           "dL_dL_u": np.empty((M * (M + 1) // 2, Q)), "dL_dKmm": np.empty((Q, M, M)), "d_rbf": np.empty((Q, 2)),
           "dW": np.empty((J, Q)), "dkappa": np.empty((J, Q)), "dZ": np.empty((M, Q * Xd))}
    gs = _lib.Grads()
    for k, a in out.items():
        setattr(gs, k, a.ctypes.data)
    st = _lib.Status()
    _lib.check(_lib.lib.hmogp_inference_host(C.byref(cfg), Xp, Yp, Np, C.byref(ps), C.byref(gs), _lib.WHAT_FULL, C.byref(st)))
    assert abs(out["log_marginal"][0, 0] - g["log_marginal"][0, 0]) < 1e-10 * abs(g["log_marginal"][0, 0])
    for k in ("dL_dmu_u", "dL_dL_u", "dL_dKmm", "d_rbf", "dW", "dkappa", "dZ"):
        assert pu.relerr(out[k], g[k]) < 1e-7, k
    assert abs(out["VE"].sum() - out["KL"][0] - out["log_marginal"][0, 0]) < 1e-9 * abs(out["log_marginal"][0, 0])
    assert list(st.chol_fail)[:Q] == [0] * Q and list(st.jitter)[:Q] == [0.0] * Q and st.n_negative_v == 0


def test_engine_results_do_not_alias_across_calls():
    prob, g = gu.load_case("cfg2_small")
    eng = pu.make_engine(prob, "fp64")
    p = pu.params_of(prob)
    outs, keeps = [], []
    for s in (1.0, 1.5, 2.0, 2.5, 3.0):
        q = dict(p)
        q["m_u"] = p["m_u"] * s
        o = eng.evaluate(q, what="full")
        outs.append(o)
        keeps.append({k: v.copy() for k, v in o.items()})
    for o, k in zip(outs, keeps):
        for name in o:
            assert np.array_equal(o[name], k[name]), name
    assert not np.array_equal(outs[0]["dZ"], outs[1]["dZ"])
    del outs, o
    again = eng.evaluate(p, what="full")                                   # freed sets are reused, not leaked
    assert len(eng._host_out[(2, False)]) <= 5 and np.array_equal(again["dZ"], keeps[0]["dZ"])
    eng.close()


def test_predict_f_equals_training_rows_and_oracle():
    prob = synth.make_problem([("Gaussian", 0.5), ("Categorical", 3), ("Gamma",)], [300, 257, 120], 37, 2, Xdim=2, seed=5)
    o = diag_oracle.elbo_and_grads(prob, want_rows=True)
    for prec, tol in (("fp64", 1e-9), ("fp32", 2e-3), ("tc", 2e-2)):
        eng = pu.make_engine(prob, prec)
        p = pu.params_of(prob)
        for t in range(prob["T"]):
            m, v = eng.predict_f(p, t, prob["X"][t])
            assert pu.relerr(m, o["rows"]["m"][t]) < tol and pu.relerr(v, o["rows"]["v"][t]) < tol, (prec, t)
        m1, v1 = eng.predict_f(p, 1, prob["X"][1][:1])                     # a single new point
        assert pu.relerr(m1, o["rows"]["m"][1][:1]) < tol
        eng.close()
    # device pointers in, device tensors out
    import torch
    eng = pu.make_engine(prob, "fp64")
    pd = {k: torch.as_tensor(v).cuda() for k, v in pu.params_of(prob).items()}
    m, v = eng.predict_f(pd, 2, torch.as_tensor(prob["X"][2]).cuda())
    assert pu.relerr(m.cpu().numpy(), o["rows"]["m"][2]) < 1e-9
    eng.close()


def _segments(tensors_p, tensors_g, spec):
    from hetmogp_b200 import _lib
    arr = (_lib.OptSegment * len(spec))()
    off = 0
    for i, (name, p_off, count, stride, pos, var) in enumerate(spec):
        arr[i].offset, arr[i].count = off, count
        arr[i].param = tensors_p[name].data_ptr() + 8 * p_off
        arr[i].grad = tensors_g[name].data_ptr() + 8 * p_off
        arr[i].stride, arr[i].positive, arr[i].variational, arr[i].reserved = stride, pos, var, 0
        off += count
    return arr, off


def test_adadelta_kernel_bit_exact_against_restatement():
    """csrc/optim.cu against oracle/climin_adadelta.py: unconstrained segments (incl. a strided one) bit for bit over 30
    iterations with the VE/VM gate switching; Logexp segments to 1e-14."""
    import torch
    from hetmogp_b200 import _lib
    lib, check = _lib.lib, _lib.check
    rng = np.random.default_rng(11)
    J, Q = 5, 3
    P = {"a": rng.normal(size=40), "b": rng.normal(size=(J, Q)), "c": rng.uniform(0.05, 3.0, size=7)}
    Pd = {k: torch.as_tensor(v.copy()).cuda() for k, v in P.items()}
    Gd = {k: torch.zeros_like(v) for k, v in Pd.items()}
    spec = [("a", 0, 40, 1, 0, 1), ("b", 1, J, Q, 0, 0), ("c", 0, 7, 1, 1, 0)]      # b: column 1 of a [J, Q] array
    arr, n = _segments(Pd, Gd, spec)
    h = C.c_void_p()
    check(lib.hmogp_opt_create(0, arr, len(spec), 0.01, 0.9, 0.9, 1e-4, C.byref(h)))
    assert lib.hmogp_opt_size(h) == n == 52
    check(lib.hmogp_opt_gather(h, None))

    def state(i):
        o = np.empty(n)
        check(lib.hmogp_opt_get_state(h, i, o.ctypes.data, None))
        return o
    wrt = np.concatenate([P["a"], P["b"][:, 1], ca.logexp_finv(P["c"])])
    assert np.array_equal(state(0)[:45], wrt[:45]) and np.allclose(state(0)[45:], wrt[45:], rtol=1e-14, atol=1e-15)
    wrt = state(0).copy()                               # start the restatement from the device's own unconstrained values
    st = ca.State(n, 0.01, 0.9, 0.9, 1e-4)
    for it in range(30):
        ve, vm = (it % 5 != 4), (it % 5 == 4)
        graw = {k: rng.normal(size=v.shape) for k, v in P.items()}
        for k in Gd:
            Gd[k].copy_(torch.as_tensor(graw[k]))
        check(lib.hmogp_opt_lookahead(h, 1, None))
        s1 = ca.lookahead(st, wrt)
        torch.cuda.synchronize()
        c_now = Pd["c"].cpu().numpy()
        assert np.array_equal(Pd["a"].cpu().numpy(), wrt[:40]) and np.array_equal(Pd["b"].cpu().numpy()[:, 1], wrt[40:45])
        assert np.array_equal(Pd["b"].cpu().numpy()[:, 0], P["b"][:, 0])     # other columns untouched
        assert np.allclose(c_now, ca.logexp_f(wrt[45:]), rtol=1e-14, atol=0)
        g = np.concatenate([-graw["a"] if ve else np.zeros(40), -graw["b"][:, 1] if vm else np.zeros(J),
                            -ca.logexp_gradfactor(c_now, graw["c"]) if vm else np.zeros(7)])
        gout = torch.empty(n, dtype=torch.float64, device="cuda")
        check(lib.hmogp_opt_update(h, int(ve), int(vm), C.c_void_p(gout.data_ptr()), None))
        gdev = gout.cpu().numpy()
        assert np.array_equal(gdev[:45], g[:45]) and np.allclose(gdev[45:], g[45:], rtol=1e-14, atol=0)
        ca.update(st, wrt, s1, gdev)                    # same gradient bits in: the update itself must be bit-exact
        for i, ref in enumerate((wrt, st.gms, st.sms, st.step)):
            assert np.array_equal(state(i), ref), (it, i)
    lib.hmogp_opt_destroy(h)


@pytest.mark.parametrize("device_loop", [True, False])
def test_svi_trace_matches_oracle_loop(device_loop):
    """50 iterations of util.vem_algorithm(stochastic=True) on cfg1 (batch 64 of 200: three slices, the last one short)
    against oracle/svi_loop.py: ELBO trace and final parameters within 1e-6."""
    from hetmogp_b200 import util
    prob, g = gu.load_case("cfg1_toy")
    m, meta = real_model(prob, batch_size=64)
    util.vem_algorithm(m, stochastic=True, vem_iters=49, step_rate=0.01, device_loop=device_loop)
    trace, pfin, st, glast = svi_loop.run(prob, 64, 50, step_rate=0.01, momentum=0.9)
    assert m.elbo.shape == (50, 1)
    assert np.max(np.abs(m.elbo[:, 0] - trace) / np.abs(trace)) < 1e-6
    assert pu.relerr(np.asarray(m.q_u_means), pfin["m_u"]) < 1e-6 and pu.relerr(np.asarray(m.q_u_chols), pfin["L_u"]) < 1e-6
    assert pu.relerr(np.asarray(m.Z), pfin["Z"]) < 1e-6
    assert pu.relerr([float(k.variance[0]) for k in m.kern_list], pfin["rbf_var"]) < 1e-6
    assert pu.relerr(np.hstack([np.asarray(B.W) for B in m.B_list]), pfin["W"]) < 1e-6
    assert abs(m.log_likelihood()[0, 0] - trace[-1]) < 1e-6 * abs(trace[-1])
    if device_loop:
        assert pu.relerr(m._svi.state("gms"), st.gms) < 1e-5 and pu.relerr(m._svi.state("step"), st.step) < 1e-5


def test_svmogp_with_process_group_shards_rows():
    """SVMOGP(group=...) shards the active slice across ranks inside the model (a caller passes the full data on every
    rank, as with the single-GPU API); world size 1 here exercises the all-reduce path end to end."""
    import os
    import torch
    import torch.distributed as dist
    prob, g = gu.load_case("cfg2_small")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        m, meta = real_model(prob, group=dist.group.WORLD)
        m.parameters_changed()
        assert abs(m.log_likelihood()[0, 0] - g["log_marginal"][0, 0]) < 1e-10 * abs(g["log_marginal"][0, 0])
        assert pu.relerr(m.Z.gradient, g["dZ"]) < 1e-7
        with torch.cuda.stream(torch.cuda.Stream()):                       # the engine follows torch's current stream
            m.parameters_changed()
        torch.cuda.synchronize()
        assert pu.relerr(m.Z.gradient, g["dZ"]) < 1e-7
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("tag", gu.LIK_TAGS)
def test_likelihood_predictive_matches_reference_golden(tag):
    """predictive(m, v) and the seeded Monte-Carlo log_predictive of every likelihood class against the UNMODIFIED reference
    (fixtures written after var_exp has run on the instance, as in a trained model: SURVEY App. C-3)."""
    from hetmogp_b200 import likelihoods as L
    g = gu.load_likelihoods()
    lik = L.from_spec(gu.LIK_SPECS[tag])
    Y, M, V = g[tag + "_Y"], g[tag + "_M"], g[tag + "_V"]
    pm, pv = lik.predictive(M, V)
    assert pm.shape == g[tag + "_pm"].shape and pv.shape == g[tag + "_pv"].shape
    assert pu.relerr(pm, g[tag + "_pm"]) < 1e-10
    if np.any(g[tag + "_pv"]):
        assert pu.relerr(pv, g[tag + "_pv"]) < 1e-9
    else:
        assert not np.any(pv)                                               # Categorical: variance "NOT IMPLEMENTED" -> zeros
    if tag + "_lp" in g:
        np.random.seed(7)
        lp = lik.log_predictive(Y, M, V, 40)
        assert abs(lp - g[tag + "_lp"][0]) < 1e-9 * max(1.0, abs(g[tag + "_lp"][0]))


def test_model_prediction_entry_points():
    """svmogp.py:219-378 on the O(M^2) route: _raw_predict (latent u_q), _raw_predict_f / predictive_new (output function d),
    predictive (likelihood moments) and negative_log_predictive, against the oracle's posterior moments at the same inputs."""
    prob, g = gu.load_case("cfg1_toy")
    m, meta = real_model(prob)
    o = diag_oracle.elbo_and_grads(prob, want_rows=True)
    fi, di = meta['function_index'].flatten(), meta['d_index'].flatten()
    for d in range(prob["J"]):
        mu, var = m._raw_predict_f(prob["X"][fi[d]], output_function_ind=d)
        assert mu.shape == (prob["X"][fi[d]].shape[0], 1)
        assert pu.relerr(mu, g["m_fd_%d" % d]) < 1e-7 and pu.relerr(var, g["v_fd_%d" % d]) < 1e-7
    mu2, var2 = m.predictive_new(prob["X"][0][:7], output_function_ind=1)
    assert pu.relerr(mu2, g["m_fd_1"][:7]) < 1e-7
    # latent q: mean K_x^T K_uu^-1 m_q, variance sigma^2 - K_x^T (K_uu^-1 - K_uu^-1 S K_uu^-1) K_x  (numpy, fp64)
    Xn = np.linspace(0.05, 0.95, 9)[:, None]
    for q in range(prob["Q"]):
        z = prob["Z"][:, q:q + 1]
        Kuu, _ = diag_oracle.rbf_K(z, z, prob["rbf_var"][q], prob["rbf_ls"][q], same=True)
        Kx, _ = diag_oracle.rbf_K(z, Xn, prob["rbf_var"][q], prob["rbf_ls"][q])
        Ki = np.linalg.inv(Kuu)
        Lq = diag_oracle.unpack_lower(prob["L_u"][:, q], prob["M"])
        mu_ref = Kx.T.dot(Ki.dot(prob["m_u"][:, q]))
        var_ref = prob["rbf_var"][q] - np.sum(Kx * (Ki - Ki.dot(Lq.dot(Lq.T)).dot(Ki)).dot(Kx), 0)
        mu, var = m._raw_predict(Xn, latent_function_ind=q)
        assert pu.relerr(mu[:, 0], mu_ref) < 1e-8 and pu.relerr(var[:, 0], np.abs(var_ref)) < 1e-8
    # likelihood-level prediction = the likelihood's predictive on the oracle's moments
    mp, vp = m.predictive(prob["X"])
    for t, lik in enumerate(m.likelihood.likelihoods_list):
        rm, rv = lik.predictive(o["rows"]["m"][t], np.abs(o["rows"]["v"][t]))
        assert pu.relerr(mp[t], rm) < 1e-7
    np.random.seed(3)
    nlpd = m.negative_log_predictive(prob["X"], prob["Y"], num_samples=30)
    np.random.seed(3)
    ref = -sum(lik.log_predictive(prob["Y"][t], o["rows"]["m"][t], np.abs(o["rows"]["v"][t]), 30)
               for t, lik in enumerate(m.likelihood.likelihoods_list))
    assert np.isfinite(nlpd) and abs(nlpd - ref) < 1e-6 * abs(ref)
