"""numpy restatement of climin's Adadelta update and of paramz' Logexp transform.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party arithmetic that is NOT under /root/reference: the reference pins ``climin==0.1a1`` (requirements.txt:6) and
calls ``climin.Adadelta(model.optimizer_array, model.stochastic_grad, step_rate=step_rate, momentum=0.9)`` followed by
``optimizer.minimize_until(c_full)`` (hetmogp/util.py:327-329); GPy==1.9.5 (requirements.txt:5) brings paramz, whose
``Model.optimizer_array`` / ``_grads`` carry positive parameters through ``Logexp``.  Neither package is importable
here (no network): the algorithms below are restated from their published sources (climin/adadelta.py,
paramz/transformations.py; recalled -- "parity unpinned" for these two third-party pieces, pinned only by the
closed-form checks in tests/test_optim_oracle.py).

climin/adadelta.py, Adadelta._iterate (defaults step_rate=1, decay=0.9, momentum=0, offset=1e-4):
    step_m1 = self.step
    step1 = step_m1 * m;  self.wrt -= step1
    gradient = self.fprime(self.wrt)
    self.gms = (d * self.gms) + (1 - d) * gradient ** 2
    step2 = sqrt(self.sms + o) / sqrt(self.gms + o) * gradient * self.step_rate
    self.wrt -= step2
    self.step = step1 + step2
    self.sms = (d * self.sms) + (1 - d) * self.step ** 2
"""
import numpy as np

LIM = 36.0                                   # paramz.transformations._lim_val
LOG_LIM = np.log(np.finfo(np.float64).max)   # paramz.transformations._log_lim_val


def logexp_f(x):
    return np.where(x > LIM, x, np.log1p(np.exp(np.clip(x, -LOG_LIM, LIM))))


def logexp_finv(f):
    return np.where(f > LIM, f, np.log(np.expm1(np.minimum(f, LIM))))


def logexp_gradfactor(f, df):
    return df * np.where(f > LIM, 1.0, -np.expm1(-f))


class State(object):
    def __init__(self, n, step_rate=1.0, decay=0.9, momentum=0.0, offset=1e-4):
        self.gms, self.sms, self.step = np.zeros(n), np.zeros(n), np.zeros(n)
        self.step_rate, self.decay, self.momentum, self.offset = step_rate, decay, momentum, offset
        self.n_iter = 0


def lookahead(st, wrt):
    """First half of an iteration: returns step1 and moves wrt to the point fprime is evaluated at."""
    step1 = st.step * st.momentum
    wrt -= step1
    return step1


def update(st, wrt, step1, gradient):
    d, o = st.decay, st.offset
    st.gms = (d * st.gms) + (1 - d) * gradient ** 2
    step2 = np.sqrt(st.sms + o) / np.sqrt(st.gms + o) * gradient * st.step_rate
    wrt -= step2
    st.step = step1 + step2
    st.sms = (d * st.sms) + (1 - d) * st.step ** 2
    st.n_iter += 1
