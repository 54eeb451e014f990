"""GPU tests: the CUDA path, called through the C-ABI, against the golden qpOASES vectors, against the
oracle on seeded batches, and through size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from conftest import CASES, F_ABS, F_REL, OBJ_REL, golden_case, assert_forces_close
from oracle import cmpc_numpy as N
from oracle import cmpc_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cmpc_b200 import engine, synth


def solve(inst, f_dist=None, capacity=None):
    b = engine.Batch(capacity or len(inst["p"]))
    b.setup(inst["dt"], inst["horizon"], inst.get("mu", 0.4), inst.get("f_max", 120.0))
    res = b.solve_host(inst, f_dist=f_dist)
    b.close()
    return res


def full_mask(forces, gait, h, mu, f_max):
    out = np.zeros((len(forces), 20 * h), dtype=np.int8)
    for i in range(len(forces)):
        steps = np.flatnonzero(gait[i])
        keep = N.contact_vars(gait[i], h)
        m = N.active_mask(forces[i][keep], mu, f_max)
        out[i].reshape(-1, 5)[steps] = m
    return out


@pytest.mark.parametrize("case", CASES)
def test_golden_parity(built_lib, golden, case):
    inst = golden_case(golden, case)
    h = inst["horizon"]
    res = solve(inst)
    assert (res["status"] == engine.ST_SOLVED).all()
    ref = golden[case + "_forces"]
    assert_forces_close(res["forces"], ref, case)
    obj = golden[case + "_objective"]
    assert (np.abs(res["objective"] - obj) <= OBJ_REL * np.abs(obj)).all()
    # bit-exact contact / constraint-active mask
    assert (res["active"] == full_mask(ref, inst["gait"], h, inst["mu"], inst["f_max"])).all()
    # swing feet carry exactly zero force (SolverMPC.cpp:973-976)
    swing = np.repeat(inst["gait"] == 0, 3, axis=1)
    assert (res["forces"][swing] == 0.0).all()
    # working-set changes track qpOASES' nWSR (same active-set family; not required to be equal)
    assert np.abs(res["iterations"] - golden[case + "_nwsr"]).max() <= 16


@pytest.mark.skipif(not O.available(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("h,gaits,spread,seed,nseg,B", [
    (10, ("trot",), 1.0, 101, None, 96), (10, ("trot",), 3.0, 102, None, 96),
    (16, ("trot", "bound", "pace", "gallop"), 1.5, 103, 10, 96), (10, ("stand",), 2.0, 104, None, 96),
    (5, ("trot",), 1.0, 105, 10, 96), (1, ("stand",), 1.0, 106, None, 96),
    (12, ("walk2", "trotrun"), 2.0, 107, 10, 96),
    # maximum sizes: every foot down over the longest horizons (n = 192, 228): global-workspace tier
    (16, ("stand",), 1.5, 108, None, 12), (19, ("stand",), 1.5, 109, None, 8), (13, ("stand",), 1.0, 110, None, 12)])
def test_live_oracle_parity(built_lib, h, gaits, spread, seed, nseg, B):
    inst = synth.make_batch(B, horizon=h, seed=seed, gaits=gaits, spread=spread, n_segment=nseg)
    res = solve(inst)
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    checked = []
    ref_forces = np.zeros_like(res["forces"])
    for i in range(B):
        r = O.solve(st, O.make_update(inst, i, h))
        if not r["ok"]:
            continue  # reference itself failed (nWSR > 100): nothing to compare against
        checked.append(i)
        ref_forces[i] = r["x"]
        assert res["status"][i] in (engine.ST_SOLVED, engine.ST_EMPTY)
        assert_forces_close(res["forces"][i], r["x"], "instance %d" % i)
        if r["n_var"]:
            assert abs(res["objective"][i] - r["objective"]) <= OBJ_REL * abs(r["objective"])
    assert len(checked) >= B * 0.9
    # activity mask against the one the ORACLE's forces give (not the GPU's own)
    ref_mask = full_mask(ref_forces[checked], inst["gait"][checked], h, inst["mu"], inst["f_max"])
    assert (res["active"][checked] == ref_mask).all()


@pytest.mark.skipif(not O.available(), reason="oracle/_ref did not travel")
def test_disturbance_hook_parity(built_lib):
    # g gains 2 Bqp' S Qqp xi (SolverMPC.cpp:810)
    h, B = 10, 32
    inst = synth.make_batch(B, horizon=h, seed=201)
    rng = np.random.default_rng(5)
    fd = (rng.normal(0, 1.0, (B, 6)) * np.array([0.3, 0.3, 0.3, 3, 3, 3])).astype(np.float32)
    res = solve(inst, f_dist=fd)
    res0 = solve(inst)
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    for i in range(B):
        r = O.solve(st, O.make_update(inst, i, h), f_dist=fd[i].astype(np.float64))
        assert_forces_close(res["forces"][i], r["x"], "xi instance %d" % i)
    assert np.abs(res["forces"] - res0["forces"]).max() > 1e-2   # the hook does move the forces


def test_edge_cases(built_lib):
    h = 10
    inst = synth.make_batch(8, horizon=h, seed=301)
    inst["gait"][0] = 0                      # nothing in contact
    inst["gait"][1] = 0
    inst["gait"][1][4 * 3 + 2] = 1           # a single foot-step in contact
    inst["gait"][2] = 1                      # everything in contact
    res = solve(inst)
    assert res["status"][0] == engine.ST_EMPTY and (res["forces"][0] == 0).all() and (res["active"][0] == 0).all()
    assert res["status"][1] == engine.ST_SOLVED
    nzcols = np.flatnonzero(res["forces"][1])
    assert set(nzcols) <= {3 * 14, 3 * 14 + 1, 3 * 14 + 2}
    assert (res["status"][2:] == engine.ST_SOLVED).all()
    # empty batch and a capacity larger than the batch
    b = engine.Batch(16)
    b.setup(0.03, h, 0.4, 120.0)
    out = b.solve_host(inst, count=0)
    assert out["forces"].shape == (0, 120)
    out = b.solve_host(inst, count=3)
    assert (out["forces"] == res["forces"][:3]).all()
    with pytest.raises(RuntimeError):
        b.setup(0.03, 20, 0.4, 120.0)        # SolverMPC.cpp:113
    small = engine.Batch(2)
    small.setup(0.03, h, 0.4, 120.0)
    with pytest.raises(RuntimeError):
        small.solve_host(inst, count=8)      # count exceeds capacity
    small.close()
    b.close()


def test_reference_single_instance_interface(built_lib, golden):
    """setup_problem / update_problem_data_floats / get_solution, convexMPC_interface.h:44-52."""
    import ctypes as C
    L = engine.lib()
    inst = golden_case(golden, "trot10")
    h = 10
    for i in range(4):
        L.setup_problem(inst["dt"], h, inst["mu"], inst["f_max"])
        L.update_x_drag(float(inst["x_drag"][i]))
        L.update_solver_settings(100, 1e-7, 1e-8, 1.5, 0.1, 0.0)
        arr = lambda k: np.ascontiguousarray(inst[k][i], dtype=np.float32)
        p, v, q, w, r, wt, tr = [arr(k) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
        gait = np.ascontiguousarray(inst["gait"][i], dtype=np.int32)
        L.update_problem_data_floats(p.ctypes.data, v.ctypes.data, q.ctypes.data, w.ctypes.data, r.ctypes.data,
                                     float(inst["rpy"][i][0]), float(inst["rpy"][i][1]), float(inst["rpy"][i][2]),
                                     wt.ctypes.data, tr.ctypes.data, float(inst["alpha"][i]), gait.ctypes.data)
        got = np.array([L.get_solution(k) for k in range(12 * h)])
        assert_forces_close(got, golden["trot10_forces"][i], "single-instance")
    # the double-precision entry point narrows to float like mfp_to_flt (convexMPC_interface.cpp:69)
    d = lambda k: np.ascontiguousarray(inst[k][0], dtype=np.float64)
    p, v, q, w, r, wt, tr = [d(k) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
    gait = np.ascontiguousarray(inst["gait"][0], dtype=np.int32)
    L.setup_problem(inst["dt"], h, inst["mu"], inst["f_max"])
    L.update_x_drag(float(inst["x_drag"][0]))
    L.update_problem_data(p.ctypes.data, v.ctypes.data, q.ctypes.data, w.ctypes.data, r.ctypes.data, 0.0,
                          wt.ctypes.data, tr.ctypes.data, float(inst["alpha"][0]), gait.ctypes.data)
    got = np.array([L.get_solution(k) for k in range(12 * h)])
    assert_forces_close(got, golden["trot10_forces"][0], "single-instance double")


@pytest.mark.parametrize("B,h,gaits,nseg", [(4096, 10, ("trot",), None), (65536, 10, ("trot",), None),
                                            (8192, 16, ("trot", "bound", "pace", "gallop"), 10)])
def test_full_size_properties(built_lib, B, h, gaits, nseg):
    """BASELINE.json sizes, checked through properties that need no CPU solve."""
    inst = synth.make_batch(B, horizon=h, seed=401, gaits=gaits, spread=1.5, n_segment=nseg)
    mu, fmax = inst["mu"], inst["f_max"]
    res = solve(inst)
    assert (res["status"] == engine.ST_SOLVED).all()
    f = res["forces"].reshape(B, -1, 3)
    mui = float(np.float32(1.0) / np.float32(mu))
    # primal feasibility of every fmat row
    assert (f[..., 2] >= -1e-8).all() and (f[..., 2] <= fmax + 1e-8).all()
    assert (np.abs(f[..., 0]) * mui <= f[..., 2] + 1e-8).all() and (np.abs(f[..., 1]) * mui <= f[..., 2] + 1e-8).all()
    # swing feet exactly zero
    assert (f[inst["gait"] == 0] == 0.0).all()
    # determinism and order independence: a permuted batch gives bit-identical per-instance answers
    perm = np.random.default_rng(0).permutation(B)
    pin = {k: (v[perm] if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
    res_p = solve(pin)
    assert (res_p["forces"] == res["forces"][perm]).all()
    assert (res_p["active"] == res["active"][perm]).all()
    # leg relabelling symmetry: swapping legs 0<->1 and 2<->3 permutes the answer
    sw = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
    sub = slice(0, 512)
    legperm = [1, 0, 3, 2]
    sw["r"] = inst["r"].reshape(B, 3, 4)[:, :, legperm].reshape(B, 12)
    sw["gait"] = inst["gait"].reshape(B, h, 4)[:, :, legperm].reshape(B, 4 * h)
    for k in sw:
        if isinstance(sw[k], np.ndarray):
            sw[k] = np.ascontiguousarray(sw[k][sub])
    res_s = solve(sw)
    f_s = res_s["forces"].reshape(512, h, 4, 3)[:, :, legperm].reshape(512, -1)
    assert np.abs(f_s - res["forces"][sub]).max() <= 1e-7
    # optimality spot check on a few instances: KKT stationarity with numpy-built H, g
    for i in range(0, B, B // 6):
        H, g = N.condense_closed(inst, i)
        keep = N.contact_vars(inst["gait"][i], h)
        x = res["forces"][i][keep]
        grad = H[np.ix_(keep, keep)] @ x + g[keep]
        xg, _ = N.gi_solve(H[np.ix_(keep, keep)], g[keep], mu, fmax)
        assert np.abs(x - xg).max() <= F_ABS
        # stationarity in the free directions: gradient vanishes on foot-steps with no active row
        m = res["active"][i].reshape(-1, 5)[np.flatnonzero(inst["gait"][i])]
        free = np.repeat(~(m != 0).any(1), 3)
        assert np.abs(grad[free]).max(initial=0.0) <= 1e-8


def test_disturbance_estimator_parity(built_lib, golden):
    """Fused estimator stage against the oracle's restatement of fit_sin (SolverMPC.cpp:478-541)."""
    t, d, est_ref = golden["dist_t"], golden["dist_d"], golden["dist_est"]
    B, h = len(t), 10
    inst = synth.make_batch(B, horizon=h, seed=501)
    sim_time = t[:, -1].copy()
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    plain = b.solve_host(inst)
    # mode 1: estimate and apply in the same launch
    b.upload_disturbance(t, d, sim_time, 1)
    res = b.solve_host(inst)
    est, fest = b.download_disturbance()
    assert (est[:, 2] == est_ref[:, 2]).all()                      # same DFT peak bin -> identical frequency
    np.testing.assert_allclose(est[:, :2], est_ref[:, :2], rtol=1e-10, atol=1e-12)
    comp = est_ref[:, 1] + np.sin(2 * np.pi * sim_time.astype(np.float64) * est_ref[:, 2] + est_ref[:, 3])
    np.testing.assert_allclose(fest[:, 3], comp.astype(np.float32), rtol=2e-7)
    assert (fest[:, [0, 1, 2, 4, 5]] == 0).all()
    # the fused launch equals an explicit-xi launch with the same f_est, bit for bit
    b.upload_disturbance(None, None, None, -1)
    explicit = b.solve_host(inst, f_dist=fest)
    assert (explicit["forces"] == res["forces"]).all()
    assert np.abs(res["forces"] - plain["forces"]).max() > 1e-3
    # mode 0: estimate only, g sees no disturbance (history of 400..500 samples)
    b.upload_disturbance(t, d, sim_time, 0)
    res0 = b.solve_host(inst)
    assert (res0["forces"] == plain["forces"]).all()
    # mode 2: no new fit, stored estimate re-evaluated at a later time and applied
    later = sim_time + np.float32(0.03)
    b.upload_disturbance(None, None, later, 2)
    res2 = b.solve_host(inst)
    est2, fest2 = b.download_disturbance()
    assert (est2 == est).all()
    comp2 = est[:, 1] + np.sin(2 * np.pi * later.astype(np.float64) * est[:, 2])
    np.testing.assert_allclose(fest2[:, 3], comp2.astype(np.float32), rtol=2e-7)
    if O.available():
        st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
        for i in range(B):
            r = O.solve(st, O.make_update(inst, i, h), f_dist=fest2[i].astype(np.float64))
            assert_forces_close(res2["forces"][i], r["x"], "adaptive instance %d" % i)
    b.close()


@pytest.mark.skipif(not O.available(), reason="oracle/_ref did not travel")
def test_single_instance_adaptive_state_machine(built_lib, golden):
    """The reference interface with the adaptive hook: 400 / 500-sample rules of SolverMPC.cpp:704-814."""
    import ctypes as C
    L = engine.lib()
    L.cmpc_reset_history()
    inst = golden_case(golden, "trot10")
    h, i = 10, 0
    arr = lambda k: np.ascontiguousarray(inst[k][i], dtype=np.float32)
    p, v, q, w, r, wt, tr = [arr(k) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
    gait = np.ascontiguousarray(inst["gait"][i], dtype=np.int32)
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    upd = O.make_update(inst, i, h)
    ad = O.Adapt()
    fd_ref = np.zeros(6)
    fest = np.zeros(6, dtype=np.float32)
    fext = np.zeros(6, dtype=np.float32)
    L.setup_problem(inst["dt"], h, inst["mu"], inst["f_max"])
    L.update_x_drag(float(inst["x_drag"][i]))
    checked = 0
    for step in range(520):
        tnow = np.float32(0.03 * step)
        fext[3] = np.float32(0.4 + 1.3 * np.sin(2 * np.pi * 0.9 * float(tnow) + 0.3))
        L.cmpc_set_external_force(fext.ctypes.data)
        L.cmpc_set_simulation_time(float(tnow))
        L.update_problem_data_floats(p.ctypes.data, v.ctypes.data, q.ctypes.data, w.ctypes.data, r.ctypes.data,
                                     0.0, 0.0, 0.0, wt.ctypes.data, tr.ctypes.data, float(inst["alpha"][i]),
                                     gait.ctypes.data)
        use = O.lib().cmpc_oracle_adapt_step(C.byref(ad), float(tnow), float(fext[3]),
                                             fd_ref.ctypes.data_as(C.POINTER(C.c_double)))
        if step in (10, 398, 399, 450, 499, 500, 501, 519):
            L.cmpc_get_disturbance_estimate(fest.ctypes.data)
            assert abs(float(fest[3]) - float(ad.f_est[3])) <= 2e-6 * max(1.0, abs(float(ad.f_est[3]))), step
            ref = O.solve(st, upd, f_dist=fd_ref if use else None)
            got = np.array([L.get_solution(k) for k in range(12 * h)])
            assert_forces_close(got, ref["x"], "adaptive step %d" % step)
            checked += 1
    assert checked == 8 and use == 1
    L.cmpc_reset_history()


def _solve_env(inst, env):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return solve(inst)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("h,gaits,nseg,B", [(10, ("trot",), None, 512), (10, ("trot", "bound", "pace", "gallop"), 10, 512),
                                            (16, ("trot", "bound", "pace", "gallop"), 10, 256), (10, ("stand",), None, 128)])
def test_every_kernel_path_agrees(built_lib, h, gaits, nseg, B):
    """Three-kernel pipeline (DMMA inversion / DFMA condensation shapes, fast and any-capacity dual tiers, resumed
    and restarted overflow, hardest-first and natural order) vs the fused single-kernel path: the same optimum to
    rounding, identical activity masks."""
    inst = synth.make_batch(B, horizon=h, seed=601, gaits=gaits, spread=2.0, n_segment=nseg)
    ref = _solve_env(inst, {"CMPC_PATH": "fused"})
    variants = [{}, {"CMPC_QCAP1": "4"}, {"CMPC_DUAL": "generic"}, {"CMPC_CSHAPE": "2"}, {"CMPC_SERIAL": "1"},
                {"CMPC_LPT": "0"},                          # natural instance order instead of hardest first
                {"CMPC_RESUME": "0"},                       # overflowed working sets restart instead of resuming
                {"CMPC_QCAP1": "4", "CMPC_RESUME": "0"}, {"CMPC_QCAP1": "12"}, {"CMPC_NSTREAMS": "2"},
                {"CMPC_DUAL_TEAM": "0"}, {"CMPC_DUAL_TEAM": "0", "CMPC_QCAP1": "4"}]   # one warp per instance beyond the first tier
    for env in variants:
        res = _solve_env(inst, env)
        assert (res["status"] == ref["status"]).all(), env
        assert np.abs(res["forces"] - ref["forces"]).max() <= 1e-7, env
        assert (np.abs(res["objective"] - ref["objective"]) <= 1e-9 * np.abs(ref["objective"]) + 1e-12).all(), env
        assert (res["active"] == ref["active"]).all(), env


def test_small_first_tier_with_a_drop_at_full_capacity(built_lib):
    """A partial step (a row leaves the working set) while the first tier's working set is exactly full: with a
    capacity below 28 rows the cleared slot's row used to be zeroed 32 entries wide — past the end of P when the slot
    was the last row.  Instances 56 / 949 (capacity 8) and 2216 (capacity 16) of this batch take that path."""
    h, B = 16, 2304
    inst = synth.make_batch(2048 * 4, horizon=h, seed=1000, gaits=("trot", "bound", "pace", "gallop"), n_segment=10, spread=1.5)
    inst = {k: (v[:B] if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
    ref = _solve_env(inst, {"CMPC_PATH": "fused"})
    for env in ({"CMPC_QCAP1": "8"}, {"CMPC_QCAP1": "16"}, {"CMPC_QCAP1": "8", "CMPC_DUAL_TEAM": "0"}, {"CMPC_QCAP1": "5"}):
        res = _solve_env(inst, env)
        assert (res["status"] == ref["status"]).all(), env
        assert np.abs(res["forces"] - ref["forces"]).max() <= 1e-7, env
        assert (res["active"] == ref["active"]).all(), env


def test_interleaved_uploads_and_solves_stay_ordered(built_lib):
    """Successive solve_range calls rotate through the engine's streams; uploads, marks and downloads
    must still see them in program order (a stream's first use grows its workspace mid-call)."""
    h, B, ring = 10, 256, 6
    inst = synth.make_batch(B * ring, horizon=h, seed=701, spread=2.0)
    b = engine.Batch(B * ring)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    b.upload(inst)
    for i in range(ring):
        b.solve_range(i * B, B)
    full = b.download()
    # the reference: the same instances solved as ONE range on one stream (same kernels, same capacity tiers)
    ref_b = engine.Batch(B * ring)
    ref_b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    ref_b.upload(inst)
    ref_b.solve_range(0, B * ring)
    whole = ref_b.download()
    ref_b.close()
    assert (full["forces"] == whole["forces"]).all() and (full["status"] == 0).all()
    # and the end-to-end call (first capacity tier of 32 rows instead of 24: a few working sets take the other tier)
    host = solve(inst)
    assert np.abs(full["forces"] - host["forces"]).max() <= 1e-7 and (full["active"] == host["active"]).all()
    # overwrite the records while solves are in flight, solve again: results follow the new records
    perm = np.random.default_rng(1).permutation(B * ring)
    inst2 = {k: (v[perm] if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
    for i in range(ring):
        b.solve_range(i * B, B)
    b.upload(inst2)
    b.mark(0)
    for i in range(ring):
        b.solve_range(i * B, B)
    b.mark(1)
    again = b.download()
    assert b.marked_ms() > 0
    assert (again["forces"] == whole["forces"][perm]).all()
    # per-kernel-class accounting: the pipeline ran, flops were counted per class
    b.reset_counters()
    t = b.profile_range(0, B)
    fl = b.kernel_flops()
    assert t["assemble"] > 0 and t["invert"] > 0 and t["dual"] > 0 and t["fused"] == 0
    assert fl["invert"] >= B * 57 ** 3 and fl["assemble"] > 0 and fl["fused"] == 0
    b.close()


def test_adaptive_batch_at_scale(built_lib, golden):
    """BASELINE.json configs[2] in miniature on the pipeline path: 4096 Adaptive-MPC instances, estimator fused
    into the assembly kernel; every instance reproduces the golden fit of the window it was given."""
    t, d, est_ref = golden["dist_t"], golden["dist_d"], golden["dist_est"]
    nw, B, h = len(t), 4096, 10
    idx = np.arange(B) % nw
    inst = synth.make_batch(B, horizon=h, seed=801)
    sim_time = t[idx, -1].copy()
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    b.upload_disturbance(t[idx], d[idx], sim_time, 1)
    res = b.solve_host(inst)
    est, fest = b.download_disturbance()
    assert (res["status"] == engine.ST_SOLVED).all()
    assert (est[:, 2] == est_ref[idx, 2]).all()
    np.testing.assert_allclose(est[:, :2], est_ref[idx, :2], rtol=1e-10, atol=1e-12)
    b.upload_disturbance(None, None, None, -1)
    explicit = b.solve_host(inst, f_dist=fest)
    assert (explicit["forces"] == res["forces"]).all()
    b.close()


def test_submit_wait_overlaps_two_batches(built_lib):
    """cmpc_batch_submit_bound / cmpc_batch_wait_bound: two batches in flight deliver what the synchronous call does;
    call-order violations are errors, not hangs."""
    h, B = 10, 1024
    insts = [synth.make_batch(B, horizon=h, seed=950 + k, spread=1.5) for k in range(2)]
    want = [solve(i) for i in insts]
    bs = []
    for inst in insts:
        b = engine.Batch(B)
        b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
        b.prepare_host(inst)
        bs.append(b)
    with pytest.raises(RuntimeError):
        bs[0].wait_prepared()                       # nothing submitted
    for rep in range(3):
        bs[0].submit_prepared()
        bs[1].submit_prepared()
        with pytest.raises(RuntimeError):
            bs[0].submit_prepared()                 # still in flight
        for k in (0, 1):
            res = bs[k].wait_prepared()
            for key in ("forces", "objective", "status", "iterations", "active"):
                assert (res[key] == want[k][key]).all(), (rep, k, key)
    for b in bs:
        b.close()


def test_tensor_core_sweep_of_the_large_shapes(built_lib):
    """CMPC_SWEEP=dmma: the 96 / 128-variable condensation shapes with the blocked sweep on the FP64 tensor cores
    (opt-in) against the default DFMA sweep: same optimum within 1e-5 N (its rounding error is larger), same masks."""
    for h, gaits, nseg in ((16, ("trot", "bound", "pace", "gallop"), 10), (10, ("stand",), None)):
        inst = synth.make_batch(192, horizon=h, seed=77, gaits=gaits, n_segment=nseg, spread=1.5)
        ref = _solve_env(inst, {})
        res = _solve_env(inst, {"CMPC_SWEEP": "dmma"})
        assert (res["status"] == ref["status"]).all()
        assert np.abs(res["forces"] - ref["forces"]).max() <= 1e-5
        assert (res["active"] == ref["active"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 7, 4097, 16389])
def test_device_packed_inputs_match_host_packed(built_lib, B):
    """Pinned inputs take the device-side path of the end-to-end call (csrc/cmpc_pack.cu: the copy engine stages the
    reference trajectories, the packing kernel reads the other ten arrays over PCIe with 16-byte loads where a chunk's
    offset allows, scalar loads elsewhere; 16389 instances are four ragged chunks whose 12-byte-per-instance arrays
    start off a 16-byte boundary); pageable inputs are packed by the host.  Same records, so bit-identical results."""
    h = 10
    inst = synth.make_batch(B, horizon=h, seed=4242, spread=1.5)
    want = solve(inst)                                   # pageable arrays: host packing
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    b.prepare_host(inst)                                 # arrays registered: device packing
    for rep in range(2):
        res = b.solve_prepared()
        for key in ("forces", "objective", "status", "iterations", "active"):
            assert (res[key] == want[key]).all(), (rep, key)
    b.close()


@pytest.mark.gpu
def test_inputs_in_cudahostalloc_memory(built_lib):
    """Input arrays that are pinned already (cudaHostAlloc, here through torch) cannot be registered again: the failed
    cmpc_host_register must not poison the next kernel-launch check, and the arrays take the device-packing path."""
    h, B = 10, 2048
    inst = synth.make_batch(B, horizon=h, seed=77, spread=1.5)
    want = solve(inst)
    keep = []
    pinned = dict(inst)
    for k, v in inst.items():
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.uint8 if k == "gait" else np.float32)).pin_memory()
            keep.append(t)
            pinned[k] = t.numpy()
    assert engine.lib().cmpc_host_register(pinned["traj"].ctypes.data, pinned["traj"].nbytes) != 0   # already pinned
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    b.prepare_host(pinned)
    res = b.solve_prepared()
    for key in ("forces", "objective", "status", "iterations", "active"):
        assert (res[key] == want[key]).all(), key
    b.close()
