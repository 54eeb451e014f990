"""GPU parity at BASELINE.json's sizes, driver-run (`-m gpu`): the CUDA path through the C-ABI against the oracle on
whole batches — qpOASES with the fp64 restatement (the anchor), qpOASES with the reference's own fp32 condensation
(`common_types.h:14`: `typedef float fpt`), and the 65536-instance Adaptive-MPC configuration."""
import os

import numpy as np
import pytest
import torch

from conftest import CASES, F_ABS, F_REL, golden_case
from oracle import cmpc_numpy as N
from oracle import cmpc_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cmpc_b200 import engine, synth

THREADS = len(os.sched_getaffinity(0))
needs_ref = pytest.mark.skipif(not O.available(), reason="oracle/_ref did not travel")


def gpu_solve(inst, resident=False):
    B = len(inst["p"])
    b = engine.Batch(B)
    b.setup(inst["dt"], inst["horizon"], inst["mu"], inst["f_max"])
    if resident:      # upload / solve_range / download: the path bench.py's `value` times (24-row first tier)
        b.upload(inst)
        for k0 in range(0, B, 4096):
            b.solve_range(k0, min(4096, B - k0))
        res = b.download()
    else:             # the end-to-end call
        res = b.solve_host(inst)
    b.close()
    return res


def oracle_batch(inst, use_float):
    h, B = inst["horizon"], len(inst["p"])
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    ups = (O.Update * B)(*[O.make_update(inst, i, h) for i in range(B)])
    forces, ok = O.solve_batch(st, ups, THREADS, use_float=use_float)
    return forces, ok != 0


def mask_from_forces(forces, inst, tol=1e-6):
    """Primal activity of the reference's fmat rows derived from a force vector — the ORACLE's."""
    h = inst["horizon"]
    out = np.zeros((len(forces), 20 * h), dtype=np.int8)
    for i in range(len(forces)):
        steps = np.flatnonzero(inst["gait"][i])
        keep = N.contact_vars(inst["gait"][i], h)
        out[i].reshape(-1, 5)[steps] = N.active_mask(forces[i][keep], inst["mu"], inst["f_max"], tol=tol)
    return out


def over_bar(got, ref):
    return (np.abs(got - ref) > F_ABS + F_REL * np.abs(ref)).any(axis=1)


def lift_nwsr_cap(inst, ref, bad):
    """The reference caps qpOASES at nWSR = 100 (SolverMPC.cpp:854, :955) and then serves the iterate it has reached —
    not the optimum, without any error (qpOASES' status is "homotopy step solved").  For the instances that disagree,
    ask qpOASES again with the cap lifted before calling them failures; returns the corrected reference and how many
    instances that settled."""
    h = inst["horizon"]
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"], nwsr=5000)
    ref = ref.copy()
    lifted = 0
    for i in np.flatnonzero(bad):
        r = O.solve(st, O.make_update(inst, int(i), h))
        if r["ok"] and r["nwsr"] > 100:
            ref[i] = r["x"]
            lifted += 1
    return ref, lifted


@needs_ref
@pytest.mark.parametrize("tag,B,h,gaits,nseg,spread,seed", [
    ("trot h10 (configs[1])", 4096, 10, ("trot",), None, 1.0, 123),
    ("mixed gaits h16 (configs[3])", 2048, 16, ("trot", "bound", "pace", "gallop"), 10, 1.5, 5)])
@pytest.mark.parametrize("resident", [False, True])
def test_batch_parity_against_qpoases(built_lib, tag, B, h, gaits, nseg, spread, seed, resident):
    """Every instance of a BASELINE-sized batch against the reference's qpOASES on the fp64 restatement: forces inside
    the north-star bar, activity masks identical to the ones the ORACLE's forces give."""
    inst = synth.make_batch(B, horizon=h, seed=seed, gaits=gaits, spread=spread, n_segment=nseg)
    res = gpu_solve(inst, resident)
    ref, ok = oracle_batch(inst, use_float=False)
    assert ok.sum() >= 0.98 * B, "qpOASES itself gave up on %d instances" % (~ok).sum()
    assert (res["status"] == engine.ST_SOLVED).all()
    bad = over_bar(res["forces"], ref) & ok
    assert not bad.any(), "%s: %d instances outside the force bar, max |dF| %.3e" % (
        tag, bad.sum(), np.abs(res["forces"] - ref)[ok].max())
    assert np.abs(res["forces"] - ref)[ok].max() <= 1e-6          # measured: < 1e-8 N
    ref_mask = mask_from_forces(ref[ok], {**inst, "gait": inst["gait"][ok]})
    assert (res["active"][ok] == ref_mask).all()


# fp32 condensation: what the reference itself computes.  Its answer is not reproducible bit for bit by anyone (Eigen's
# summation order); measured distance between the fp32 and the fp64 condensation, both through qpOASES
# (DESIGN.md §5): trot h10 median 3e-4 N, 99th percentile 8e-4 N, max 1.5e-3 N; hard / mixed-gait h16 cases up to
# 9e-3 N.  So the bar of 1e-3 N is inside the reference's OWN rounding noise on the hard cases; what can be asserted is
# (i) the bar on (nearly) all nominal trot instances, (ii) a bound of the measured fp32 noise everywhere, (iii) that the
# GPU answer is the fp64 optimum (it coincides with the fp64 oracle), (iv) identical activity masks — no constraint
# flips between the two arithmetics.
FP32_NOISE = {"trot10": 1.0e-3, "trot10hard": 2.0e-3, "mixed16": 4.0e-3, "stand10": 2.0e-3, "pronk10": 1.2e-3}


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_golden_cases_against_fp32_reference_arithmetic(built_lib, golden, case):
    inst = golden_case(golden, case)
    res = gpu_solve(inst)
    ref32, ok = oracle_batch(inst, use_float=True)
    assert ok.all()
    err = np.abs(res["forces"] - ref32)
    assert err.max() <= FP32_NOISE[case], "%s: %.3e N from the fp32 reference arithmetic" % (case, err.max())
    assert (res["active"] == mask_from_forces(ref32, inst)).all()


@needs_ref
def test_trot4096_against_fp32_reference_arithmetic(built_lib):
    inst = synth.make_batch(4096, horizon=10, seed=123)
    res = gpu_solve(inst)
    ref32, ok32 = oracle_batch(inst, use_float=True)
    ref64, ok64 = oracle_batch(inst, use_float=False)
    ok = ok32 & ok64
    assert ok.all()
    per = np.abs(res["forces"] - ref32).max(axis=1)
    bad = over_bar(res["forces"], ref32)
    print("fp32 reference arithmetic, 4096 trot: |dF| median %.2e, p99 %.2e, max %.2e N; %d instances over the bar; "
          % (np.median(per), np.quantile(per, 0.99), per.max(), bad.sum()))
    assert bad.sum() <= 8 and per.max() <= 2.5e-3               # measured: 3 instances, max 1.54e-3 N
    assert np.quantile(per, 0.99) <= 1e-3
    # the GPU answer is the fp64 optimum: its distance to the fp32 result IS the fp32 condensation's own error
    assert np.abs(res["forces"] - ref64).max() <= 1e-6
    masks = mask_from_forces(ref32, inst)
    mism = (res["active"] != masks).any(axis=1).sum()
    assert mism == 0, "%d instances whose activity mask differs from the fp32 reference's" % mism


@needs_ref
def test_adaptive_65536(built_lib):
    """BASELINE.json configs[2] at full size: 65536 Adaptive-MPC instances, disturbance fit + apply fused into the
    launch.  1024 distinct windows (each fitted by the oracle on the CPU), every instance a different robot state:
    the fit of EVERY instance is checked against the oracle's fit of its window (peak bin bit-equal), and a
    512-instance sample of the forces against qpOASES with the same xi in g (SolverMPC.cpp:704-814)."""
    B, h, U = 65536, 10, 1024
    t, d, _ = synth.make_disturbance_windows(U, seed=31)
    est_ref = np.stack([O.fit_window(t[i], d[i]) for i in range(U)])
    idx = np.arange(B) % U
    inst = synth.make_batch(B, horizon=h, seed=802)
    sim_time = t[idx, -1].copy()
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    b.upload_disturbance(t[idx], d[idx], sim_time, 1)
    res = b.solve_host(inst)
    est, fest = b.download_disturbance()
    b.close()
    assert (res["status"] == engine.ST_SOLVED).all()
    assert (est[:, 2] == est_ref[idx, 2]).all()                  # same DFT peak bin -> bit-identical frequency
    np.testing.assert_allclose(est[:, :2], est_ref[idx, :2], rtol=1e-10, atol=1e-12)
    assert (est == est[idx]).all()                               # the same window gives the same bits wherever it runs
    comp = est_ref[idx, 1] + np.sin(2 * np.pi * sim_time.astype(np.float64) * est_ref[idx, 2] + est_ref[idx, 3])
    np.testing.assert_allclose(fest[:, 3], comp.astype(np.float32), rtol=2e-7)
    assert (fest[:, [0, 1, 2, 4, 5]] == 0).all()
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    sample = np.random.default_rng(9).choice(B, 512, replace=False)
    for i in sample:
        r = O.solve(st, O.make_update(inst, i, h), f_dist=fest[i].astype(np.float64))
        assert r["ok"]
        e = np.abs(res["forces"][i] - r["x"])
        assert (e <= F_ABS + F_REL * np.abs(r["x"])).all(), "adaptive instance %d: %.3e N" % (i, e.max())
    f = res["forces"].reshape(B, -1, 3)
    mui = float(np.float32(1.0) / np.float32(inst["mu"]))
    assert (f[..., 2] >= -1e-8).all() and (f[..., 2] <= inst["f_max"] + 1e-8).all()
    assert (np.abs(f[..., 0]) * mui <= f[..., 2] + 1e-8).all() and (np.abs(f[..., 1]) * mui <= f[..., 2] + 1e-8).all()


@needs_ref
@pytest.mark.parametrize("dt,mu,f_max,mass,inertia", [
    (0.026, 0.2, 60.0, 12.0, (0.07, 0.26, 0.242)),        # slippery ground, weak legs: many active cone faces and bounds
    (0.05, 0.8, 240.0, 12.0, (0.07, 0.26, 0.242)),        # long steps, high friction
    (0.03, 0.4, 120.0, 45.0, (0.55, 2.1, 2.3)),           # a heavier robot through cmpc_batch_set_robot
    (0.002, 1.0, 500.0, 5.0, (0.02, 0.05, 0.06))])        # tiny dt: the discretisation polynomials at their small end
def test_setup_parameter_variations_against_qpoases(built_lib, dt, mu, f_max, mass, inertia):
    """problem_setup (dt, mu, f_max; convexMPC_interface.h:15-21) and the robot constants the reference hard-codes
    (RobotState.h:24, RobotState.cpp:49) away from the A1 defaults, against qpOASES on the same setup."""
    h, B = 10, 256
    inst = synth.make_batch(B, horizon=h, dt=dt, seed=321, gaits=("trot", "bound", "pronk"), spread=1.5)
    b = engine.Batch(B)
    b.set_robot(mass, inertia)
    b.setup(dt, h, mu, f_max)
    res = b.solve_host(inst)
    b.close()
    st = O.make_setup(dt, h, mu, f_max, mass=mass, inertia=inertia)
    ups = (O.Update * B)(*[O.make_update(inst, i, h) for i in range(B)])
    ref, ok = O.solve_batch(st, ups, THREADS, use_float=False)
    ok = ok != 0
    assert ok.sum() >= 0.9 * B
    assert (res["status"] == engine.ST_SOLVED).all()
    bad = over_bar(res["forces"], ref) & ok
    assert not bad.any(), "%d instances outside the force bar, max |dF| %.3e" % (bad.sum(), np.abs(res["forces"] - ref)[ok].max())
    ref_mask = mask_from_forces(ref[ok], {**inst, "gait": inst["gait"][ok], "mu": mu, "f_max": f_max})
    assert (res["active"][ok] == ref_mask).all()


@needs_ref
@pytest.mark.parametrize("gaits", [("trot", "pace"), ("trot", "pace", "walk2")],
                         ids=["n60-tensor-core-inversion", "n84-register-tile-sweep"])
def test_weight_and_regularisation_variations_against_qpoases(built_lib, gaits):
    """Per-instance state weights (Q, ConvexMPCLocomotion.cpp:617: some of them zero) and force regularisation alpha
    (:623) away from the A1 defaults — alpha from 1e-6 (H conditioned ~1e5: without the panel refinement of the blocked
    sweeps the 60-variable pipeline returned forces 13 N off, profiles/r2_illcond_accuracy.txt) to 1e-3 — and the x_drag
    coupling from zero to large, against qpOASES.  Two batches: one that stays on the n <= 63 pipeline, one whose
    three-leg stance steps send it through the 96-variable condensation shape."""
    h, B = 10, 512
    inst = synth.make_batch(B, horizon=h, seed=654, gaits=gaits, spread=1.5)
    rng = np.random.default_rng(11)
    w = synth.A1_WEIGHTS[None, :] * rng.uniform(0.2, 5.0, (B, 12)).astype(np.float32)
    w[rng.random((B, 12)) < 0.1] = 0.0
    w[:, 5] = np.maximum(w[:, 5], 1.0)                      # keep the height weighted: an unweighted height is unbounded drift, not a test
    inst["weights"] = w.astype(np.float32)
    inst["alpha"] = rng.choice(np.array([1e-6, 4e-5, 1e-3], np.float32), B)
    inst["x_drag"] = (rng.choice(np.array([0.0, 0.3, 3.0, -2.0], np.float32), B)).astype(np.float32)
    res = gpu_solve(inst)
    ref, ok = oracle_batch(inst, use_float=False)
    assert ok.sum() >= 0.9 * B
    assert (res["status"] == engine.ST_SOLVED).all()
    bad = over_bar(res["forces"], ref) & ok
    ref, lifted = lift_nwsr_cap(inst, ref, bad)      # alpha = 1e-6 makes a few instances need more than 100 working-set changes
    assert lifted <= 4
    bad = over_bar(res["forces"], ref) & ok
    assert not bad.any(), "%d instances outside the force bar, max |dF| %.3e" % (bad.sum(), np.abs(res["forces"] - ref)[ok].max())
    assert np.abs(res["forces"] - ref)[ok].max() < 5e-6      # measured 2e-7 N at alpha = 1e-6 (the bar is 1e-3 N)
    # Rows whose slack at the oracle's forces lies within a decade of the activity tolerance (1e-6) are undecidable at
    # that agreement; every other row must match.
    sub = {**inst, "gait": inst["gait"][ok]}
    m_lo, m_hi = mask_from_forces(ref[ok], sub, tol=1e-7), mask_from_forces(ref[ok], sub, tol=1e-5)
    decided = m_lo == m_hi
    assert decided.mean() > 0.999
    assert (res["active"][ok][decided] == m_lo[decided]).all()
