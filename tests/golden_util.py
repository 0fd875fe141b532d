"""Load the committed golden fixtures (tests/golden/*.npz, written by oracle/make_golden.py from the unmodified
reference) together with the seeded problems they were computed on."""
import os

import numpy as np

from oracle import make_golden

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(make_golden.INFERENCE_CASES)
LIK_TAGS = ["HetGaussian", "Bernoulli", "Categorical3", "Gamma", "Beta", "Poisson", "Gaussian", "Exponential",
            "Categorical4"]
LIK_SPECS = dict(zip(LIK_TAGS, make_golden.ALL))


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, "inference_%s.npz" % name)))
    prob = make_golden.problem_from_case(make_golden.INFERENCE_CASES[name])
    return prob, g


def load_likelihoods():
    return dict(np.load(os.path.join(GOLDEN_DIR, "likelihoods.npz")))
