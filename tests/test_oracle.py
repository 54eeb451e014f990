"""CPU tests: the oracle against the committed golden vectors (which came from the reference's
real qpOASES), the independent numpy restatement against the oracle, the estimator restatement."""
import numpy as np
import pytest

from conftest import CASES, golden_case, assert_forces_close
from oracle import cmpc_numpy as N
from oracle import cmpc_oracle as O

needs_ref = pytest.mark.skipif(not O.available(), reason="oracle/_ref not built (reference sources absent)")


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_golden(golden, case):
    inst = golden_case(golden, case)
    h = inst["horizon"]
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    for i in range(len(inst["p"])):
        r = O.solve(st, O.make_update(inst, i, h))
        assert r["ok"]
        np.testing.assert_allclose(r["x"], golden[case + "_forces"][i], rtol=0, atol=1e-9)
        assert r["nwsr"] == golden[case + "_nwsr"][i]
        assert abs(r["objective"] - golden[case + "_objective"][i]) <= 1e-12 * abs(r["objective"]) + 1e-15


@needs_ref
def test_oracle_float_mode_stays_within_force_bar(golden):
    # the reference condenses in fp32 (common_types.h:14); the fp64 anchor must sit within the
    # 1e-3 N bar of that arithmetic on the nominal trot case
    inst = golden_case(golden, "trot10")
    st = O.make_setup(inst["dt"], 10, inst["mu"], inst["f_max"])
    worst = 0.0
    for i in range(len(inst["p"])):
        u = O.make_update(inst, i, 10)
        r32, r64 = O.solve(st, u, use_float=True), O.solve(st, u)
        worst = max(worst, np.abs(r32["x"] - r64["x"]).max())
        # no constraint flips between the two arithmetics
        keep = N.contact_vars(inst["gait"][i], 10)
        assert (N.active_mask(r32["x"][keep], inst["mu"], inst["f_max"]) ==
                N.active_mask(r64["x"][keep], inst["mu"], inst["f_max"])).all()
    assert worst < 1e-3, worst     # measured 7.6e-4 N over the 24 golden trot instances


@needs_ref
@pytest.mark.parametrize("case", ["trot10", "mixed16"])
def test_numpy_dense_and_closed_form_match_oracle(golden, case):
    inst = golden_case(golden, case)
    h = inst["horizon"]
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    fd = np.array([0.1, -0.2, 0.3, 1.5, -0.7, 0.4])
    for i in range(3):
        r = O.solve(st, O.make_update(inst, i, h), f_dist=fd, want_mats=True)
        Hd, gd, _ = N.condense_dense(inst, i, f_dist=fd)
        Hc, gc = N.condense_closed(inst, i, f_dist=fd)
        sH, sg = np.abs(Hd).max(), np.abs(gd).max()
        assert np.abs(r["H_full"] - Hd).max() <= 1e-13 * sH
        assert np.abs(r["g_full"] - gd).max() <= 1e-13 * sg
        assert np.abs(Hc - Hd).max() <= 1e-13 * sH
        assert np.abs(gc - gd).max() <= 1e-13 * sg


@pytest.mark.parametrize("case", CASES)
def test_closed_form_matches_golden_gradient_and_diagonal(golden, case):
    # no reference library needed: the fixtures carry g and diag(H) from the oracle
    inst = golden_case(golden, case)
    for i in range(2):
        Hc, gc = N.condense_closed(inst, i)
        g_ref, d_ref = golden[case + "_g"][i], golden[case + "_Hdiag"][i]
        assert np.abs(gc - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
        assert np.abs(np.diag(Hc) - d_ref).max() <= 1e-12 * np.abs(d_ref).max()


@pytest.mark.parametrize("case", CASES)
def test_gi_prototype_matches_golden_forces(golden, case):
    """The algorithm the CUDA kernel implements, in numpy, against qpOASES' answers."""
    inst = golden_case(golden, case)
    h = inst["horizon"]
    for i in range(min(4, len(inst["p"]))):
        H, g = N.condense_closed(inst, i)
        keep = N.contact_vars(inst["gait"][i], h)
        x, info = N.gi_solve(H[np.ix_(keep, keep)], g[keep], inst["mu"], inst["f_max"])
        assert info["status"] == 0
        full = np.zeros(12 * h)
        full[keep] = x
        assert_forces_close(full, golden[case + "_forces"][i], case)
        ref_obj = golden[case + "_objective"][i]
        assert abs(info["objective"] - ref_obj) <= 1e-7 * abs(ref_obj)
        ref_mask = N.active_mask(golden[case + "_forces"][i][keep], inst["mu"], inst["f_max"])
        assert (N.active_mask(x, inst["mu"], inst["f_max"]) == ref_mask).all()


def test_primal_mask_agrees_with_qpoases_working_set(golden):
    # wherever qpOASES holds a row in its working set the row is primal-active; the converse can
    # fail only at degenerate vertices (apex of the friction pyramid), which the primal mask covers
    for case in CASES:
        inst = golden_case(golden, case)
        h = inst["horizon"]
        for i in range(len(inst["p"])):
            keep = N.contact_vars(inst["gait"][i], h)
            mask = N.active_mask(golden[case + "_forces"][i][keep], inst["mu"], inst["f_max"]).reshape(-1)
            kept_steps = np.flatnonzero(inst["gait"][i])
            ws = golden[case + "_workingset"][i].reshape(-1, 5)[kept_steps].reshape(-1)
            assert (mask[ws != 0] == ws[ws != 0]).all()


@needs_ref
def test_estimator_oracle_against_golden_and_truth(golden):
    t, d, est = golden["dist_t"], golden["dist_d"], golden["dist_est"]
    for i in range(len(t)):
        got = O.fit_window(t[i], d[i])
        np.testing.assert_allclose(got, est[i], rtol=1e-12, atol=1e-14)
    # the FFT-peak guess resolves frequency to one bin: 1/(400*0.03) Hz
    from cmpc_b200 import synth
    tt, dd, truth = synth.make_disturbance_windows(8, seed=21)
    for i in range(8):
        assert abs(O.fit_window(tt[i], dd[i])[2] - truth["freq"][i, 0]) <= 1.01 / (400 * 0.03)
