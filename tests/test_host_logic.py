"""CPU: host-side logic of the drop-in surface, with the CUDA engine replaced by a stand-in that answers from the CPU oracle
(tests may use oracle/; the product never does).  Covers what needs no GPU:

  * the minibatch slice stream and batch scales (util.py:52-72, svmogp.py:38-47,89-90,175-186) -- bit-exact against
    fixtures written from the UNMODIFIED reference (tests/golden/util_streams.npz, oracle/make_golden.py);
  * toy generators / random_W_kappas under fixed seeds (util.py:21-50,92-104,202-206);
  * SVMOGP.parameters_changed's VE / VM gating and .gradient write-back (svmogp.py:100-166) against
    oracle/params_changed.assemble on the reference's dense gradients;
  * paramz-style optimizer_array (link order, Logexp), regular-expression fix / unfix, vem_algorithm's masks;
  * climin's Adadelta update: closed forms, host class == restatement bit for bit, and the full stochastic loop
    (vem_algorithm(stochastic=True, device_loop=False)) against oracle/svi_loop.py.
"""
import os

import numpy as np
import pytest

import golden_util as gu
from oracle import climin_adadelta as ca
from oracle import diag_oracle, params_changed, svi_loop, verbatim

GOLD = dict(np.load(os.path.join(gu.GOLDEN_DIR, "util_streams.npz")))
STREAM_CASES = [(200, 50), (203, 50), (7, 3), (5, 8), (1000, 64), (64, 64)]


# ------------------------------------------------------------------------------------------------ minibatch stream
@pytest.mark.parametrize("n,bs", STREAM_CASES)
def test_minibatch_stream_bit_exact(n, bs):
    from hetmogp_b200 import util
    sl = util.mini_slices(n, bs)
    assert np.array_equal(np.array([[s.start, s.stop] for s in sl], dtype=np.int64), GOLD["mini_%d_%d" % (n, bs)])
    gen = util.draw_mini_slices(n, bs)
    seq = [next(gen) for _ in range(2 * len(sl) + 1)]                       # two epochs and the wrap-around
    assert np.array_equal(np.array([[s.start, s.stop] for s in seq], dtype=np.int64), GOLD["draw_%d_%d" % (n, bs)])
    X = np.zeros((n, 1))
    assert np.array_equal(np.array([X[s].shape[0] for s in seq], dtype=np.int64), GOLD["len_%d_%d" % (n, bs)])
    scales = np.array([util.get_batch_scales([X], [X[s]])[0] for s in seq if X[s].shape[0] > 0])
    assert np.array_equal(scales, GOLD["scale_%d_%d" % (n, bs)])
    # the oracle's own restatement serves the same stream
    o = svi_loop.slice_stream(n, bs)
    assert [(s.start, s.stop) for s in seq] == [(s.start, s.stop) for s in (next(o) for _ in seq)]


@pytest.mark.skipif(not verbatim.available(), reason="reference tree not present")
def test_minibatch_stream_matches_reference_live():
    from hetmogp_b200 import util
    ref = verbatim.load().util
    for n, bs in ((1001, 100), (10, 3), (3, 10)):
        a, b = util.draw_mini_slices(n, bs), ref.draw_mini_slices(n, bs)
        for _ in range(25):
            s, r = next(a), next(b)
            assert (s.start, s.stop, s.step) == (r.start, r.stop, r.step)


def test_seeded_generators_match_reference():
    from hetmogp_b200 import util, likelihoods as L
    from hetmogp_b200.het_likelihood import HetLikelihood
    np.random.seed(101)
    W_list, kappa_list = util.random_W_kappas(3, 5, rank=1)
    assert np.array_equal(np.hstack(W_list), GOLD["rwk_W"]) and np.array_equal(np.stack(kappa_list, axis=1), GOLD["rwk_kappa"])
    np.random.seed(102)
    Xl = [np.linspace(0, 1, 17)[:, None], np.linspace(-1, 2, 9)[:, None]]
    tu = util.true_u_functions(Xl, 3)
    assert np.allclose(tu[0], GOLD["true_u_0"], rtol=0, atol=1e-14) and np.allclose(tu[1], GOLD["true_u_1"], rtol=0, atol=1e-14)
    liks = [L.from_spec(s) for s in (("HetGaussian",), ("Bernoulli",))]
    meta = HetLikelihood(liks).generate_metadata()
    np.random.seed(103)
    W_list, _ = util.random_W_kappas(3, 3, rank=1)
    assert np.array_equal(np.hstack(W_list), GOLD["true_f_W"])
    tf = util.true_f_functions([GOLD["true_u_0"], GOLD["true_u_1"]], W_list, 3, liks, meta)
    assert np.allclose(tf[0], GOLD["true_f_0"], rtol=0, atol=1e-14) and np.allclose(tf[1], GOLD["true_f_1"], rtol=0, atol=1e-14)
    np.random.seed(104)
    assert np.allclose(util.generate_toy_U(np.linspace(0, 1, 11)[:, None], 4), GOLD["toy_U"], rtol=0, atol=1e-14)


# ------------------------------------------------------------------------------------------------ oracle-backed stand-in engine
class OracleEngine(object):
    """Answers Engine's calls from oracle/diag_oracle.py (fp64 numpy): the host logic above it cannot tell."""

    def __init__(self, lik_specs, M, Q, Xdim, precision="fp32", device=0, group=None):
        self.lik_specs, self.M, self.Q, self.Xdim = [tuple(s) for s in lik_specs], M, Q, Xdim
        self.T = len(self.lik_specs)
        self.device, self.calls, self.status = device, [], None

    def set_data(self, X, Y):
        self.X, self.Y = [np.asarray(x) for x in X], [np.asarray(y) for y in Y]
        self.rows = [slice(0, x.shape[0]) for x in self.X]

    def set_rows(self, begin=None, count=None):
        self.rows = [slice(0, x.shape[0]) for x in self.X] if begin is None else [slice(b, b + c) for b, c in zip(begin, count)]

    def set_stream(self, s):
        pass

    def evaluate(self, params, what="full", want_dKmm=False, out=None):
        prob = dict(X=self.X, Y=self.Y, lik_specs=self.lik_specs, Q=self.Q, Xdim=self.Xdim, M=self.M)
        for k in ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa"):
            prob[k] = np.asarray(params[k], dtype=np.float64)
        prob["batch_scale"] = list(np.asarray(params["batch_scale"])) if params.get("batch_scale") is not None else None
        o = diag_oracle.elbo_and_grads(prob, row_slices=self.rows, W_chain=params.get("W_chain"), kappa_chain=params.get("kappa_chain"))
        self.calls.append((what, [(s.start, s.stop) for s in self.rows], prob["batch_scale"]))
        res = {"log_marginal": np.asarray(o["log_marginal"]).reshape(1, 1), "VE": o["VE_sum"], "KL": np.array([o["KL"]])}
        if what in ("ve", "full"):
            res.update(dL_dmu_u=np.hstack(o["dL_dmu_u"]), dL_dL_u=np.hstack(o["dL_dL_u"]))
        if what == "full":
            res.update(dL_dKmm=np.stack(o["dL_dKmm"]), d_rbf=o["d_rbf"], dW=o["dW"], dkappa=o["dkappa"], dZ=o["dZ"])
        self.status = {"jitter": list(o["jitter"]), "chol_fail": [0] * self.Q, "lu_singular": [0] * self.Q, "n_negative_v": 0}
        return res

    def close(self):
        pass


def build_model(prob, monkeypatch, batch_size=None, compat_stale_W=False, W_list=None):
    from hetmogp_b200 import svmogp, util, likelihoods as L
    from hetmogp_b200.het_likelihood import HetLikelihood
    monkeypatch.setattr(svmogp, "Engine", OracleEngine)
    Q, Xdim, J = prob["Q"], prob["Xdim"], prob["J"]
    lik = HetLikelihood([L.from_spec(s) for s in prob["lik_specs"]])
    meta = lik.generate_metadata()
    kern_list = util.latent_functions_prior(Q, lenghtscale=prob["rbf_ls"], variance=prob["rbf_var"], input_dim=Xdim)
    np.random.seed(0)
    m = svmogp.SVMOGP(prob["X"], prob["Y"], prob["Z"][:, :Xdim], kern_list, lik, meta, batch_size=batch_size,
                      W_list=W_list if W_list is not None else [prob["W"][:, q:q + 1].copy() for q in range(Q)],
                      compat_stale_W=compat_stale_W)
    # put the model at the problem's parameter values
    np.asarray(m.Z)[...] = prob["Z"]
    np.asarray(m.q_u_means)[...] = prob["m_u"]
    np.asarray(m.q_u_chols)[...] = prob["L_u"]
    for q in range(Q):
        np.asarray(m.B_list[q].W)[...] = prob["W"][:, q:q + 1]
        np.asarray(m.B_list[q].kappa)[...] = prob["kappa"][:, q]
    return m, meta


def golden_gradients_dict(prob, g):
    Q, J = prob["Q"], prob["J"]
    return {"dL_dmu_u": [g["dL_dmu_u"][:, q:q + 1] for q in range(Q)], "dL_dL_u": [g["dL_dL_u"][:, q:q + 1] for q in range(Q)],
            "dL_dKmm": [g["dL_dKmm"][q] for q in range(Q)],
            "dL_dKmn": [[g["dL_dKmn_%d_%d" % (q, d)] for d in range(J)] for q in range(Q)],
            "dL_dKdiag": [[g["dL_dKdiag_%d_%d" % (q, d)] for d in range(J)] for q in range(Q)]}


@pytest.mark.parametrize("stochastic,vem_step", [(False, True), (True, True), (True, False)])
def test_parameters_changed_gating_and_write_back(monkeypatch, stochastic, vem_step):
    """svmogp.py:100-166: every .gradient field against the line-by-line restatement applied to the UNMODIFIED reference's
    dense gradients dict (golden cfg1), for the non-stochastic model and for VE / VM steps of the stochastic one."""
    prob, g = gu.load_case("cfg1_toy")
    N = prob["X"][0].shape[0]
    m, meta = build_model(prob, monkeypatch, batch_size=N if stochastic else None)   # one slice = every row: batch_scale 1
    m.vem_step = vem_step
    m.parameters_changed()
    exp = params_changed.assemble(golden_gradients_dict(prob, g), prob, {k[5:]: v for k, v in g.items() if k.startswith("meta_")},
                                  stochastic=stochastic, vem_step=vem_step)
    assert m.log_likelihood().shape == (1, 1)
    assert abs(m.log_likelihood()[0, 0] - g["log_marginal"][0, 0]) < 1e-10 * abs(g["log_marginal"][0, 0])
    tol = dict(rtol=1e-9, atol=1e-9)
    assert np.allclose(m.q_u_means.gradient, exp["m_u"], **tol) and np.allclose(m.q_u_chols.gradient, exp["L_u"], **tol)
    assert np.allclose(m.Z.gradient, exp["Z"], **tol)
    for q in range(prob["Q"]):
        assert np.allclose(m.kern_list[q].gradient, exp["rbf"][q], **tol)
        assert np.allclose(m.kern_list[q].variance.gradient, exp["rbf"][q, 0], **tol)
        assert np.allclose(m.B_list[q].W.gradient[:, 0], exp["W"][:, q], **tol)
        assert np.allclose(m.B_list[q].kappa.gradient, exp["kappa"][:, q], **tol)
    what = m._eng.calls[-1][0]
    assert what == ("ve" if (stochastic and vem_step) else "full")          # a VE step never pays for hyper-gradients
    if stochastic and vem_step:
        assert not np.any(m.Z.gradient) and not np.any(m.kern_list[0].gradient)
    if stochastic and not vem_step:
        assert not np.any(m.q_u_means.gradient) and not np.any(m.q_u_chols.gradient)
    # results are the model's own copies
    keep = m.log_likelihood().copy()
    np.asarray(m.q_u_means)[...] *= 1.5
    m.parameters_changed()
    m.parameters_changed()
    assert keep[0, 0] == g["log_marginal"][0, 0] or abs(keep[0, 0] - g["log_marginal"][0, 0]) < 1e-8


def test_stale_W_chain_option(monkeypatch):
    """quirk C-5: with compat_stale_W the kernel / Z chain uses the constructor-time W (svmogp.py:98-99,141,143,156)."""
    prob, g = gu.load_case("cfg2_small")
    rng = np.random.default_rng(3)
    W0 = prob["W"] + 0.2 * rng.normal(size=prob["W"].shape)
    m, meta = build_model(prob, monkeypatch, compat_stale_W=True, W_list=[W0[:, q:q + 1].copy() for q in range(prob["Q"])])
    m.parameters_changed()
    o = diag_oracle.elbo_and_grads(prob, W_chain=W0, kappa_chain=np.zeros_like(prob["kappa"]))
    for q in range(prob["Q"]):
        assert np.allclose(m.kern_list[q].gradient, o["d_rbf"][q], rtol=1e-10, atol=1e-12)
    assert np.allclose(m.Z.gradient, o["dZ"], rtol=1e-10, atol=1e-12)
    m2, _ = build_model(prob, monkeypatch, compat_stale_W=False, W_list=[W0[:, q:q + 1].copy() for q in range(prob["Q"])])
    m2.parameters_changed()
    o2 = diag_oracle.elbo_and_grads(prob)
    assert np.allclose(m2.Z.gradient, o2["dZ"], rtol=1e-10, atol=1e-12)


def test_constructor_consumes_slice_zero_and_batch_scales(monkeypatch):
    """svmogp.py:38-47: the constructor's new_batch() takes slice 0; batches then cycle 1,2,...,0,1 with a short last slice
    and batch_scale = N_all / N_batch (svmogp.py:89-90)."""
    prob, g = gu.load_case("cfg2_small")                     # N = 120 per task
    m, meta = build_model(prob, monkeypatch, batch_size=50)
    assert m._eng.calls[0][1] == [(0, 50)] * 3 and m._eng.calls[0][2] == [120 / 50.0] * 3
    seen = []
    x = m.optimizer_array
    for _ in range(4):
        m.stochastic_grad(x)
        seen.append((m._eng.calls[-1][1][0], m._eng.calls[-1][2][0], m._eng.calls[-1][0]))
    assert [s[0] for s in seen] == [(50, 100), (100, 120), (0, 50), (50, 100)]
    assert [s[1] for s in seen] == [2.4, 6.0, 2.4, 2.4]
    # VE, VE, VE, VE, then VM (svmogp.py:191-198: ve_count runs 0..3)
    assert [s[2] for s in seen] == ["ve", "ve", "ve", "ve"]
    m.stochastic_grad(x)
    assert m._eng.calls[-1][0] == "full"
    m.stochastic_grad(x)
    assert m._eng.calls[-1][0] == "ve"
    assert [xx.shape[0] for xx in m.Xmulti] == [m._slice[t][1] for t in range(3)]


def test_optimizer_array_link_order_and_logexp(monkeypatch):
    prob, g = gu.load_case("cfg2_small")
    prob["kappa"] = prob["kappa"] + 0.3                       # positive, so the Logexp round trip is defined
    m, meta = build_model(prob, monkeypatch)
    Q, J, M, Xd = prob["Q"], prob["J"], prob["M"], prob["Xdim"]
    x = m.optimizer_array
    n_expected = M * Q * Xd + M * Q + (M * (M + 1) // 2) * Q + 2 * Q + 2 * J * Q
    assert x.shape == (n_expected,)
    # link order (svmogp.py:71-75): Z, m_u, L_u, kernels (variance, lengthscale), B's (W, kappa)
    i = 0
    assert np.array_equal(x[i:i + M * Q * Xd], prob["Z"].ravel()); i += M * Q * Xd
    assert np.array_equal(x[i:i + M * Q], prob["m_u"].ravel()); i += M * Q
    assert np.array_equal(x[i:i + prob["L_u"].size], prob["L_u"].ravel()); i += prob["L_u"].size
    for q in range(Q):
        assert np.allclose(ca.logexp_f(x[i:i + 2]), [prob["rbf_var"][q], prob["rbf_ls"][q]], rtol=1e-14); i += 2
    for q in range(Q):
        assert np.array_equal(x[i:i + J], prob["W"][:, q]); i += J
        assert np.allclose(ca.logexp_f(x[i:i + J]), prob["kappa"][:, q], rtol=1e-14); i += J
    # round trip and the transformed gradient (paramz Model._grads)
    m.optimizer_array = x
    assert np.allclose(m.optimizer_array, x, rtol=1e-13, atol=1e-13)
    gneg = m._grads(x)
    o = diag_oracle.elbo_and_grads(prob)
    assert np.allclose(gneg[:M * Q * Xd], -o["dZ"].ravel(), rtol=1e-9, atol=1e-10)
    k0 = M * Q * Xd + M * Q + prob["L_u"].size
    assert np.allclose(gneg[k0], -o["d_rbf"][0, 0] * (1.0 - np.exp(-prob["rbf_var"][0])), rtol=1e-9)
    # fixing by regular expression removes the block (util.py:285-318 usage)
    m['.*.kappa'].fix()
    m['.*.lengthscale'].fix()
    assert m.optimizer_array.shape == (n_expected - J * Q - Q,)
    m.Z.fix()
    assert m.optimizer_array.shape == (n_expected - J * Q - Q - M * Q * Xd,)
    m['.*.lengthscale'].unfix()
    assert m.optimizer_array.shape == (n_expected - J * Q - M * Q * Xd,)
    with pytest.raises(AttributeError):
        m['.*.nothing']
    assert np.allclose(svmogp_logexp_roundtrip(), 0.0, atol=1e-12)


def svmogp_logexp_roundtrip():
    from hetmogp_b200 import svmogp
    t = np.array([1e-8, 1e-3, 0.5, 3.0, 35.0, 36.5, 80.0])
    return svmogp._logexp_f(svmogp._logexp_finv(t)) / t - 1.0


# ------------------------------------------------------------------------------------------------ Adadelta
def test_adadelta_restatement_closed_forms():
    """First two iterations of climin's update by hand (decay d, offset o, momentum m, step rate r)."""
    d, o, mom, r = 0.9, 1e-4, 0.9, 0.01
    g1, g2 = np.array([2.0, -0.5]), np.array([1.0, 0.25])
    wrt = np.array([1.0, 1.0])
    st = ca.State(2, r, d, mom, o)
    s1 = ca.lookahead(st, wrt)
    assert np.array_equal(s1, np.zeros(2)) and np.array_equal(wrt, [1.0, 1.0])
    ca.update(st, wrt, s1, g1)
    gms1 = (1 - d) * g1 ** 2
    step_a = np.sqrt(o) / np.sqrt(gms1 + o) * g1 * r
    assert np.allclose(st.gms, gms1, rtol=1e-15) and np.allclose(st.step, step_a, rtol=1e-15)
    assert np.allclose(wrt, 1.0 - step_a, rtol=1e-15) and np.allclose(st.sms, (1 - d) * step_a ** 2, rtol=1e-15)
    s1 = ca.lookahead(st, wrt)
    assert np.allclose(s1, mom * step_a, rtol=1e-15) and np.allclose(wrt, 1.0 - step_a - mom * step_a, rtol=1e-15)
    ca.update(st, wrt, s1, g2)
    gms2 = d * gms1 + (1 - d) * g2 ** 2
    step_b = mom * step_a + np.sqrt((1 - d) * step_a ** 2 + o) / np.sqrt(gms2 + o) * g2 * r
    assert np.allclose(st.step, step_b, rtol=1e-14) and st.n_iter == 2
    # paramz Logexp: limits and gradient factor
    assert ca.logexp_f(np.array([40.0]))[0] == 40.0 and ca.logexp_finv(np.array([40.0]))[0] == 40.0
    assert abs(ca.logexp_f(np.array([0.0]))[0] - np.log(2.0)) < 1e-16
    assert abs(ca.logexp_gradfactor(np.array([np.log(2.0)]), np.array([3.0]))[0] - 1.5) < 1e-15


def test_host_adadelta_equals_restatement_bitwise():
    from hetmogp_b200.optim import Adadelta
    rng = np.random.default_rng(0)
    A = rng.normal(size=(7, 7))
    A = A.dot(A.T) + np.eye(7)
    fprime = lambda w: A.dot(w) + np.sin(w)
    w0 = rng.normal(size=7)
    wrt = w0.copy()
    opt = Adadelta(wrt, fprime, step_rate=0.01, momentum=0.9)
    seen = []
    opt.minimize_until(lambda info: seen.append(info['n_iter']) or info['n_iter'] > 20)
    assert seen == list(range(1, 22))                                       # max_iter + 1 evaluations (svmogp.py:214-216)
    w = w0.copy()
    st = ca.State(7, 0.01, 0.9, 0.9, 1e-4)
    for _ in range(21):
        s1 = ca.lookahead(st, w)
        ca.update(st, w, s1, fprime(w))
    assert np.array_equal(w, wrt) and np.array_equal(st.gms, opt.gms) and np.array_equal(st.sms, opt.sms)


def test_stochastic_vem_loop_matches_oracle_loop(monkeypatch, capsys):
    """vem_algorithm(stochastic=True) through the host classes == oracle/svi_loop.py: slices, batch scales, gating, VE/VM
    toggle, Logexp chain, link order, Adadelta -- 24 iterations, ELBO trace and final parameters."""
    from hetmogp_b200 import util
    prob, g = gu.load_case("cfg1_toy")
    m, meta = build_model(prob, monkeypatch, batch_size=64)
    util.vem_algorithm(m, stochastic=True, vem_iters=23, step_rate=0.01, device_loop=False, verbose=False)
    trace, pfin, st, glast = svi_loop.run(prob, 64, 24, step_rate=0.01, momentum=0.9)
    assert m.elbo.shape == (24, 1)
    assert np.allclose(m.elbo[:, 0], trace, rtol=1e-12, atol=0)
    assert np.allclose(np.asarray(m.q_u_means), pfin["m_u"], rtol=1e-12, atol=1e-14)
    assert np.allclose(np.asarray(m.Z), pfin["Z"], rtol=1e-12, atol=1e-14)
    assert np.allclose([float(k.variance[0]) for k in m.kern_list], pfin["rbf_var"], rtol=1e-12)
    assert np.array_equal([float(k.lengthscale[0]) for k in m.kern_list], prob["rbf_ls"])       # fixed (util.py:285)
    assert m.kern_list[0].lengthscale.is_fixed and m.B_list[0].kappa.is_fixed


def test_vem_algorithm_masks_non_stochastic(monkeypatch, capsys):
    """util.py:296-318: VE step frees q(U) only; VM step frees variances, lengthscales, W (+Z), fixes q(U); kappa fixed."""
    from hetmogp_b200 import util, svmogp
    prob, g = gu.load_case("cfg2_small")
    m, meta = build_model(prob, monkeypatch)
    seen = []

    def fake_optimize(self, messages=False, max_iters=1000, **kw):
        seen.append({name.split('.', 1)[1]: p.is_fixed for name, p, _, _ in self._named()})
    monkeypatch.setattr(svmogp.SVMOGP, "optimize", fake_optimize)
    util.vem_algorithm(m, stochastic=False, vem_iters=2, optZ=True, non_chained=True)
    assert len(seen) == 4
    ve, vm = seen[0], seen[1]
    assert not ve["m_u"] and not ve["L_u"] and ve["inducing_inputs"] and all(v for k, v in ve.items() if k not in ("m_u", "L_u"))
    assert vm["m_u"] and vm["L_u"] and not vm["inducing_inputs"]
    assert all(v for k, v in vm.items() if k.endswith("kappa")) and not any(v for k, v in vm.items() if k.endswith((".W", "variance", "lengthscale")))
    seen[:] = []
    util.vem_algorithm(m, stochastic=False, vem_iters=1, optZ=False, non_chained=False)
    assert seen[1]["inducing_inputs"] and all(v for k, v in seen[1].items() if k.endswith(".W"))
