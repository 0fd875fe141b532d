"""Seeded synthetic problems -- re-exported from hetmogp_b200/synth.py (the generator is shared by
bench.py and the tests; it contains no arithmetic of the hot path)."""
from hetmogp_b200.synth import *  # noqa: F401,F403
from hetmogp_b200.synth import CONFIGS, make_problem, make_config, subsample, inducing_grid, sample_outputs  # noqa: F401
