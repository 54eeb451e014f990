import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "quad-periodic-mpc_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden", "cmpc_golden.npz")
IN_KEYS = ("p", "v", "q", "w", "r", "rpy", "weights", "traj", "alpha", "gait", "x_drag")
CASES = ("trot10", "trot10hard", "mixed16", "stand10", "pronk10")

# parity bars of BASELINE.json's north_star
F_ABS, F_REL, OBJ_REL = 1e-3, 1e-5, 1e-7


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


def golden_case(gold, name):
    inst = {k: gold["%s_in_%s" % (name, k)] for k in IN_KEYS}
    h, dt, mu, fmax = gold[name + "_meta"]
    inst.update(horizon=int(h), dt=float(dt), mu=float(mu), f_max=float(fmax))
    return inst


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as ge
    return ge.build()


def assert_forces_close(got, ref, what=""):
    err = np.abs(got - ref)
    tol = F_ABS + F_REL * np.abs(ref)
    assert (err <= tol).all(), "%s forces off by %.3e N (tolerance 1e-3 abs + 1e-5 rel)" % (what, err.max())
