"""GPU tests of the drop-in boundary: an unmodified C++ caller linked against the library, the failure modes of the
reference interface (non-finite data, error policy), the filtered disturbance estimates, ordering of the pinned
record staging, and the option interface that replaced environment switches."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_case, assert_forces_close
from oracle import cmpc_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cmpc_b200 import engine, synth


def test_cxx_caller_with_reference_declarations_solves_golden_instance(built_lib, golden, tmp_path):
    """g++ translation unit that sees only convexMPC_interface.h's declarations (update_x_drag with C++ linkage),
    linked against libcmpc_b200.so in place of the reference's sources, solves golden instances."""
    from test_host import build_reference_caller
    exe = build_reference_caller(built_lib)
    inst = golden_case(golden, "trot10")
    h = inst["horizon"]
    for i in (0, 5):
        fin, fout = tmp_path / ("in%d.bin" % i), tmp_path / ("out%d.bin" % i)
        with open(fin, "wb") as f:
            f.write(np.int32(h).tobytes())
            f.write(np.array([inst["dt"], inst["mu"], inst["f_max"], inst["x_drag"][i], inst["alpha"][i]], np.float32).tobytes())
            for k in ("p", "v", "q", "w", "r", "weights", "traj"):
                f.write(np.ascontiguousarray(inst[k][i], np.float32).tobytes())
            f.write(np.ascontiguousarray(inst["gait"][i], np.int32).tobytes())
        subprocess.run([exe, str(fin), str(fout)], check=True, timeout=120)
        got = np.fromfile(fout, dtype=np.float64)
        assert_forces_close(got, golden["trot10_forces"][i], "C++ caller, instance %d" % i)


def test_non_finite_inputs_are_reported_not_served(built_lib):
    h, B = 10, 64
    inst = synth.make_batch(B, horizon=h, seed=11)
    clean = engine.Batch(B)
    clean.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    want = clean.solve_host(inst)
    bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
    bad["p"][3, 2] = np.nan          # NaN height
    bad["traj"][17, 5] = np.inf      # Inf in the reference trajectory
    bad["r"][40, 0] = np.nan         # NaN foot position: H itself is NaN
    for opts in ({}, {"path_fused": 1}, {"dual_generic": 1}):
        b = engine.Batch(B, options=opts)
        b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
        res = b.solve_host(bad)
        b.close()
        hit = np.array([3, 17, 40])
        assert (res["status"][hit] == engine.ST_NONFINITE).all(), (opts, res["status"][hit])
        assert (res["forces"][hit] == 0).all() and (res["active"][hit] == 0).all()
        ok = np.setdiff1d(np.arange(B), hit)
        assert (res["status"][ok] == engine.ST_SOLVED).all()
        assert np.abs(res["forces"][ok] - want["forces"][ok]).max() <= 1e-7
    clean.close()


def _single_args(inst, i):
    arr = lambda k: np.array(inst[k][i], dtype=np.float32)      # copies: a test may poison them
    p, v, q, w, r, wt, tr = [arr(k) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
    gait = np.ascontiguousarray(inst["gait"][i], dtype=np.int32)
    return p, v, q, w, r, wt, tr, gait


def _single_solve(L, a, alpha):
    p, v, q, w, r, wt, tr, gait = a
    L.update_problem_data_floats(p.ctypes.data, v.ctypes.data, q.ctypes.data, w.ctypes.data, r.ctypes.data,
                                 0.0, 0.0, 0.0, wt.ctypes.data, tr.ctypes.data, float(alpha), gait.ctypes.data)


def test_reference_interface_holds_last_good_forces_on_request(built_lib, golden):
    """CMPC_ON_ERROR_HOLD: a solve that meets NaN keeps the previous forces for get_solution() and reports the status
    (the default policy aborts the process: a controller must not run on garbage)."""
    L = engine.lib()
    L.cmpc_reset_history()
    inst = golden_case(golden, "trot10")
    h = 10
    L.setup_problem(inst["dt"], h, inst["mu"], inst["f_max"])
    L.update_x_drag(float(inst["x_drag"][0]))
    good = _single_args(inst, 0)
    _single_solve(L, good, inst["alpha"][0])
    first = np.array([L.get_solution(k) for k in range(12 * h)])
    assert_forces_close(first, golden["trot10_forces"][0], "before the fault")
    assert L.cmpc_last_status() == engine.ST_SOLVED
    L.cmpc_set_error_policy(engine.ON_ERROR_HOLD)
    try:
        bad = _single_args(inst, 1)
        bad[0][2] = np.nan
        _single_solve(L, bad, inst["alpha"][1])
        assert L.cmpc_last_status() == engine.ST_NONFINITE
        held = np.array([L.get_solution(k) for k in range(12 * h)])
        assert (held == first).all()
        _single_solve(L, _single_args(inst, 1), inst["alpha"][1])     # recovers with the next good input
        assert L.cmpc_last_status() == engine.ST_SOLVED
        L.update_x_drag(float(inst["x_drag"][1]))
        _single_solve(L, _single_args(inst, 1), inst["alpha"][1])
        assert_forces_close(np.array([L.get_solution(k) for k in range(12 * h)]), golden["trot10_forces"][1], "after")
    finally:
        L.cmpc_set_error_policy(engine.ON_ERROR_ABORT)
        L.cmpc_reset_history()


def test_default_policy_aborts_on_non_finite_forces(built_lib, golden, tmp_path):
    """The default: the process dies with a message instead of returning NaN / stale forces (run in a child)."""
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from cmpc_b200 import engine, synth
L = engine.lib()
inst = synth.make_batch(1, horizon=10, seed=3)
a = lambda k: np.ascontiguousarray(inst[k][0], dtype=np.float32)
p, v, q, w, r, wt, tr = [a(k) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
p[2] = np.nan
gait = np.ascontiguousarray(inst["gait"][0], dtype=np.int32)
L.setup_problem(0.03, 10, 0.4, 120.0)
L.update_problem_data_floats(p.ctypes.data, v.ctypes.data, q.ctypes.data, w.ctypes.data, r.ctypes.data, 0.0, 0.0, 0.0,
                             wt.ctypes.data, tr.ctypes.data, 4e-5, gait.ctypes.data)
print("survived")
""" % (os.path.join(ROOT, "quad-periodic-mpc_b200"), ROOT)
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "survived" not in r.stdout
    assert "status 6" in r.stderr


def test_zero_time_step_in_the_disturbance_history_is_an_error(built_lib, golden):
    """A controller that never sets simulation_time feeds the estimator a window with dt = 0: frequency = inf and
    sin(inf) = NaN in f_est[3] (the reference then hands qpOASES a NaN gradient).  Here the instance reports
    CMPC_ST_NONFINITE as soon as the estimate is applied."""
    t, d = golden["dist_t"][:4].copy(), golden["dist_d"][:4]
    t[1] = t[1, 0]                                  # a frozen clock
    B, h = 4, 10
    inst = synth.make_batch(B, horizon=h, seed=501)
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    b.upload_disturbance(t, d, t[:, -1].copy(), 1)
    res = b.solve_host(inst)
    assert res["status"][1] == engine.ST_NONFINITE and (res["forces"][1] == 0).all()
    assert (res["status"][[0, 2, 3]] == engine.ST_SOLVED).all()
    b.close()


@pytest.mark.skipif(not O.available(), reason="oracle/_ref did not travel")
def test_filtered_disturbance_estimates(built_lib, golden):
    """f_est_smoothed / f_est_static (SolverMPC.h:73-74, SolverMPC.cpp:783, :798) through the reference interface."""
    L = engine.lib()
    L.cmpc_reset_history()
    inst = golden_case(golden, "trot10")
    h = 10
    a = _single_args(inst, 0)
    ad = O.Adapt()
    fd = np.zeros(6)
    fext = np.zeros(6, dtype=np.float32)
    sm, stc, fe = (np.zeros(6, dtype=np.float32) for _ in range(3))
    L.setup_problem(inst["dt"], h, inst["mu"], inst["f_max"])
    for step in range(430):
        tnow = np.float32(0.03 * step)
        fext[3] = np.float32(-0.2 + 0.8 * np.sin(2 * np.pi * 0.6 * float(tnow)))
        L.cmpc_set_external_force(fext.ctypes.data)
        L.cmpc_set_simulation_time(float(tnow))
        _single_solve(L, a, inst["alpha"][0])
        O.lib().cmpc_oracle_adapt_step(C.byref(ad), float(tnow), float(fext[3]), fd.ctypes.data_as(C.POINTER(C.c_double)))
        if step in (0, 50, 399, 400, 429):
            L.cmpc_get_disturbance_estimate_smoothed(sm.ctypes.data)
            L.cmpc_get_disturbance_estimate_static(stc.ctypes.data)
            L.cmpc_get_disturbance_estimate(fe.ctypes.data)
            np.testing.assert_allclose(sm, np.array(ad.f_est_smoothed[:], np.float32), rtol=1e-5, atol=1e-7)
            assert abs(float(stc[3]) - float(ad.f_est_static3)) <= 1e-6 * max(1.0, abs(float(ad.f_est_static3)))
            assert (stc[[0, 1, 2, 4, 5]] == 0).all()
    assert abs(float(sm[3])) > 1e-3            # the estimate did feed the filter after 400 samples
    L.cmpc_reset_history()
    L.cmpc_get_disturbance_estimate_smoothed(sm.ctypes.data)
    assert (sm == 0).all()


def test_back_to_back_uploads_behind_a_deep_queue(built_lib):
    """upload(A); solve(); upload(B) without a sync in between: the copy of A out of the pinned staging is still queued
    behind earlier solves when B is packed — it must still deliver A's records."""
    h, B = 10, 2048
    big = synth.make_batch(B, horizon=h, seed=31, spread=2.0)
    A = synth.make_batch(B, horizon=h, seed=32)
    Bb = synth.make_batch(B, horizon=h, seed=33)
    ref = engine.Batch(B)
    ref.setup(big["dt"], h, big["mu"], big["f_max"])
    ref.upload(A)
    ref.solve()
    want = ref.download()
    ref.close()
    b = engine.Batch(B)
    b.setup(big["dt"], h, big["mu"], big["f_max"])
    for rep in range(3):
        b.upload(big)
        for _ in range(48):
            b.solve_range(0, B)          # ~10 ms of queued work on the engine's streams
        b.upload(A)
        b.solve()
        b.upload(Bb)                     # packs into the same staging buffer
        got = b.download()
        assert (got["forces"] == want["forces"]).all(), rep
        b.solve()                        # and B's records did arrive afterwards
        after = b.download()
        assert np.abs(after["forces"] - want["forces"]).max() > 1e-3
    b.close()


def test_options_replace_environment_switches(built_lib):
    h, B = 10, 256
    inst = synth.make_batch(B, horizon=h, seed=41, spread=2.0)
    b = engine.Batch(B)
    b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    base = b.solve_host(inst)
    with pytest.raises(RuntimeError):
        b.set_option("no_such_switch", 1)
    for key, val in (("path_fused", 1), ("path_fused", 0), ("qcap1", 6), ("serial", 1), ("d2h_copy", 1), ("host_pack", 1)):
        b.set_option(key, val)
        res = b.solve_host(inst)
        assert np.abs(res["forces"] - base["forces"]).max() <= 1e-7, key
        assert (res["active"] == base["active"]).all(), key
    b.close()


def test_inversion_kernel_switches_agree(built_lib):
    """The pivot blocks of the n <= 63 inversion kernel by the fp32 chain + FP64 Newton steps (default) and by the FP64
    chain (inv_f32 = 0), the panel refinement never / always / by the default threshold: the same forces to rounding on a
    well conditioned batch, and on a badly conditioned one (alpha = 1e-6) every variant WITH the refinement agrees —
    without it the FP64 chain is 13 N off there (profiles/r2_illcond_accuracy.txt), which is why it exists."""
    h, B = 10, 512
    inst = synth.make_batch(B, horizon=h, seed=77, gaits=("trot", "pace"), spread=1.5)
    hard = dict(inst)
    hard["alpha"] = np.full(B, 1e-6, np.float32)
    w = np.array(inst["weights"], copy=True)
    w[:, [0, 7]] = 0.0                                   # unweighted roll and y velocity: H conditioned ~1e5
    hard["weights"] = w

    def run(data, **opts):
        b = engine.Batch(B, options=opts)
        b.setup(data["dt"], h, data["mu"], data["f_max"])
        res = b.solve_host(data)
        b.close()
        assert (res["status"] == engine.ST_SOLVED).all(), opts
        return res

    base = run(inst)
    for opts in ({"inv_f32": 0}, {"inv_refine": -1}, {"inv_refine": 0}, {"inv_f32": 0, "inv_refine": 0}):
        res = run(inst, **opts)
        assert np.abs(res["forces"] - base["forces"]).max() < 1e-7, opts
        assert (res["active"] == base["active"]).all(), opts
    hbase = run(hard)
    for opts in ({"inv_f32": 0}, {"inv_refine": 0}, {"inv_f32": 0, "inv_refine": 0}, {"path_fused": 1}):
        res = run(hard, **opts)
        assert np.abs(res["forces"] - hbase["forces"]).max() < 2e-6, opts


def test_commands_reject_a_changed_robot_count(built_lib):
    h, B = 10, 64
    b = engine.Batch(B)
    b.setup(0.03, h, 0.4, 120.0)
    cmds = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=h, seed=7)
    b.solve_commands(cmds)
    with pytest.raises(RuntimeError):
        b.solve_commands(cmds[:32])      # the histories were built for 64 robots
    b.reset_history()
    b.solve_commands(cmds[:32])
    b.close()


def test_sharded_engine_calls_cover_the_batch_once(built_lib):
    """What `bench.py --gpus N` / `--mode sweep1m` do, on one device: the batch is cut with bench.shard_bounds, every shard is
    solved by its own engine handle (as a rank would), and the concatenated results equal the unsharded call bit for bit
    (instances are independent: no collective, nothing shared between the shards)."""
    import bench
    h, B, world = 10, 5000, 3                      # a ragged split: 1667 + 1667 + 1666
    inst = synth.make_batch(B, horizon=h, seed=61, spread=1.5)
    whole = engine.Batch(B)
    whole.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    want = whole.solve_host(inst)
    whole.close()
    covered = np.zeros(B, dtype=np.int32)
    parts = []
    for rank in range(world):
        lo, hi = bench.shard_bounds(B, rank, world)
        covered[lo:hi] += 1
        shard = {k: (v[lo:hi] if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
        b = engine.Batch(hi - lo)
        b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
        parts.append(b.solve_host(shard))
        b.close()
    assert (covered == 1).all()
    for key in ("forces", "objective", "status", "iterations", "active"):
        got = np.concatenate([p[key] for p in parts])
        assert (got == want[key]).all(), key
