"""Host-side driver object with climin's Adadelta interface (climin==0.1a1, requirements.txt:6 of the reference; the
package is not vendored under /root/reference -- semantics recalled).

``Adadelta(wrt, fprime, step_rate=1, decay=0.9, momentum=0, offset=1e-4)`` iterates

    step1 = momentum * step;  wrt -= step1;  g = fprime(wrt)
    gms = decay * gms + (1 - decay) * g**2
    step2 = sqrt(sms + offset) / sqrt(gms + offset) * g * step_rate;  wrt -= step2
    step = step1 + step2;  sms = decay * sms + (1 - decay) * step**2

in place on ``wrt`` and yields ``{'n_iter': ...}`` dicts; ``minimize_until(criterion)`` stops at the first truthy
criterion, as util.vem_algorithm uses it (util.py:327-329).  This class only sequences the iteration for callers that
need a Python callback per step: ``fprime`` is the GPU evaluation (``SVMOGP.stochastic_grad``).  The device-resident loop
(``SVMOGP.svi_device``, csrc/optim.cu) performs the same update without leaving the GPU and is what
``vem_algorithm`` uses by default.
"""
import numpy as np


class Adadelta(object):
    def __init__(self, wrt, fprime, step_rate=1, decay=0.9, momentum=0, offset=1e-4, args=None):
        self.wrt = wrt
        self.fprime = fprime
        self.step_rate, self.decay, self.momentum, self.offset = step_rate, decay, momentum, offset
        self.gms = np.zeros_like(wrt)
        self.sms = np.zeros_like(wrt)
        self.step = np.zeros_like(wrt)
        self.n_iter = 0
        self.args = args

    def __iter__(self):
        d, o, m = self.decay, self.offset, self.momentum
        while True:
            step1 = self.step * m
            self.wrt -= step1
            gradient = self.fprime(self.wrt)
            self.gms = (d * self.gms) + (1 - d) * gradient ** 2
            step2 = np.sqrt(self.sms + o) / np.sqrt(self.gms + o) * gradient * self.step_rate
            self.wrt -= step2
            self.step = step1 + step2
            self.sms = (d * self.sms) + (1 - d) * self.step ** 2
            self.n_iter += 1
            yield {'n_iter': self.n_iter, 'gradient': gradient, 'args': (), 'kwargs': {}}

    def minimize_until(self, criterions):
        if not isinstance(criterions, (list, tuple)):
            criterions = [criterions]
        if not criterions:
            raise ValueError('need to supply at least one criterion')
        for info in self:
            for criterion in criterions:
                if criterion(info):
                    return info
