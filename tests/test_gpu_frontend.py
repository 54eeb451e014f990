"""GPU tests of the caller of the path on the device (cmpc_batch_solve_commands: updateMPCIfNeeded / solveDenseMPC /
getMpcTable as CUDA kernels, csrc/cmpc_frontend.cu) against the oracle's restatement (oracle/cmpc_frontend.py)."""
import numpy as np
import pytest
import torch

from conftest import assert_forces_close
from oracle import cmpc_frontend as F
from oracle import cmpc_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cmpc_b200 import engine, synth

H, DT, MU, FMAX = 10, 0.03, 0.4, 120.0


def _batch(cap):
    b = engine.Batch(cap)
    b.setup(DT, H, MU, FMAX)
    return b


@pytest.mark.parametrize("with_log", [True, False])
def test_records_are_bit_identical_to_the_oracle(built_lib, with_log):
    B = 512
    c = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=H, gaits=("trot", "bound", "pace", "gallop"), seed=21,
                            mixed_fraction=0.2, stand_fraction=0.1, with_log=with_log, sim_time=1.25)
    b = _batch(B)
    res, forces = b.solve_commands(c, want_forces=True)
    inst, ex = F.solver_inputs(c, H, DT, np.zeros((B, 6), np.float32))
    want = F.pack_records(inst, sim_time=c["sim_time"])
    got = b.copy_records(0, B)
    assert (got == want).all(), "records differ in %d bytes" % int((got != want).sum())
    # the command state solveDenseMPC leaves behind
    assert (res["world_position_desired"] == ex["world_position_desired"]).all()
    assert (res["x_comp_integral"] == ex["x_comp_integral"]).all()
    assert (res["f_ext"] == ex["f_ext"]).all()
    if not with_log:
        assert (res["f_ext"] == 0).all()
    # Fr_des / f_ff from the first horizon step
    f, f_ff = F.leg_outputs(c, forces)
    assert (res["fr_des"] == f).all() and (res["f_ff"] == f_ff).all()
    assert (res["status"] == engine.ST_SOLVED).all()
    # the same instances through the update_problem_data-level call: identical records, identical solve
    b2 = _batch(B)
    inst.update(mu=MU, f_max=FMAX)
    ref = b2.solve_host(inst)
    assert (ref["forces"] == forces).all() and (ref["iterations"] == res["iterations"]).all()
    if O.available():
        st = O.make_setup(DT, H, MU, FMAX)
        for i in range(0, B, 37):
            r = O.solve(st, O.make_update(inst, i, H))
            assert_forces_close(forces[i], r["x"], "command instance %d" % i)
    b.close()
    b2.close()


def test_f_ext_persists_without_log_data(built_lib):
    """received_log_data_ == false leaves the global f_ext as it was (ConvexMPCLocomotion.cpp:773-776)."""
    B = 64
    b = _batch(B)
    c = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=H, seed=5, with_log=True)
    first = b.solve_commands(c)
    c2 = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=H, seed=6, with_log=False)
    second = b.solve_commands(c2)
    assert (second["f_ext"] == first["f_ext"]).all() and (first["f_ext"] != 0).any()
    b.reset_history()
    third = b.solve_commands(c2)
    assert (third["f_ext"] == 0).all() and b.history_length() == 1
    b.close()


def test_history_and_estimator_state_machine(built_lib):
    """solve_mpc's bookkeeping on the device (SolverMPC.cpp:688-813): (simulation_time, f_ext[3]) pushed every call;
    the sinusoid is fitted while the history holds 400..500 samples and not applied; beyond 500 the stored fit is
    refreshed at simulation_time and applied in g.  Checked against the explicit-window path of the batched engine
    fed with the history the test keeps itself."""
    B = 8
    rng = np.random.default_rng(77)
    amp, freq, phase = rng.uniform(0.5, 2.0, B), rng.uniform(0.3, 1.2, B), rng.uniform(-3, 3, B)
    b = _batch(B)
    ref = _batch(B)
    hist_t, hist_d = [], []
    checked = 0
    for step in range(506):
        t = np.float32(DT * step)
        c = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=H, seed=900, sim_time=float(t),
                                disturbance=(amp, freq, phase))
        res, forces = b.solve_commands(c, want_forces=True)
        hist_t.append(np.full(B, t, np.float32))
        hist_d.append(res["f_ext"][:, 3].copy())
        n = len(hist_t)
        assert b.history_length() == n
        if n in (1, 399, 400, 401, 450, 500, 501, 506):
            inst, _ = F.solver_inputs(c, H, DT, np.zeros((B, 6), np.float32))
            inst.update(mu=MU, f_max=FMAX)
            if n < 400:
                ref.upload_disturbance(None, None, None, -1)
                want = ref.solve_host(inst)
            else:
                wt = np.stack(hist_t[-400:], axis=1)
                wd = np.stack(hist_d[-400:], axis=1)
                if n <= 500:
                    ref.upload_disturbance(wt, wd, hist_t[-1], 0)      # fit, not applied
                    want = ref.solve_host(inst)
                    est_ref, fest_ref = ref.download_disturbance()
                else:
                    ref.upload_disturbance(None, None, hist_t[-1], 2)  # the fit of sample 500, applied
                    want = ref.solve_host(inst)
                    est_ref2, fest_ref = ref.download_disturbance()
                    assert (est_ref2 == est_ref).all()
                est, fest = b.download_disturbance()
                assert (est == est_ref).all() and (fest == fest_ref).all(), n
                if n == 450:   # the fit recovers the injected frequency to the DFT's resolution (1 / (400 dt) Hz)
                    assert np.abs(est[:, 2] - freq).max() <= 1.0 / (400 * DT) + 1e-9
            assert (want["forces"] == forces).all(), n
            checked += 1
    assert checked == 8
    # beyond 500 samples the estimate moves the forces
    ref.upload_disturbance(None, None, None, -1)
    plain = ref.solve_host(inst)
    assert np.abs(plain["forces"] - forces).max() > 1e-6
    b.close()
    ref.close()


def test_command_edge_cases(built_lib):
    b = _batch(16)
    c = synth.make_commands(16, engine.COMMAND_DTYPE, horizon=H, seed=1)
    assert len(b.solve_commands(c[:0])) == 0
    # a robot with every leg in swing over the horizon: zero forces, status EMPTY
    c["gait_durations"][3] = 0
    c["stand"][3] = 0
    c["gait_kind"][3] = 0
    res, forces = b.solve_commands(c, want_forces=True)
    assert res["status"][3] == engine.ST_EMPTY and (forces[3] == 0).all() and (res["fr_des"][3] == 0).all()
    assert (np.delete(res["status"], 3) == engine.ST_SOLVED).all()
    with pytest.raises(RuntimeError):
        b.solve_commands(np.zeros(17, dtype=engine.COMMAND_DTYPE))
    b.close()
