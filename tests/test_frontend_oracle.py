"""CPU tests of the oracle's restatement of the CALLER of the path (updateMPCIfNeeded / solveDenseMPC / getMpcTable,
oracle/cmpc_frontend.py) and of the cmpc_command layout shared by the header, the binding and the oracle."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from oracle import cmpc_frontend as F
from cmpc_b200 import engine, synth


def _commands(B=96, **kw):
    kw.setdefault("gaits", ("trot", "bound", "pace", "gallop"))
    kw.setdefault("stand_fraction", 0.1)
    return synth.make_commands(B, engine.COMMAND_DTYPE, seed=11, mixed_fraction=0.25, **kw)


def test_command_layout_matches_the_header(tmp_path):
    """sizeof / offsetof of cmpc_command and cmpc_command_result as gcc sees include/cmpc_b200.h == the numpy dtypes."""
    src = tmp_path / "layout.c"
    fields_c = [n for n in engine.COMMAND_DTYPE.names]
    fields_r = [n for n in engine.RESULT_DTYPE.names]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cmpc_b200.h"', 'int main(void) {',
             'printf("%zu %zu\\n", sizeof(cmpc_command), sizeof(cmpc_command_result));']
    lines += ['printf("%%zu\\n", offsetof(cmpc_command, %s));' % f for f in fields_c]
    lines += ['printf("%%zu\\n", offsetof(cmpc_command_result, %s));' % f for f in fields_r]
    lines += ['return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(out[0]), int(out[1])] == [engine.COMMAND_DTYPE.itemsize, engine.RESULT_DTYPE.itemsize]
    offs = [int(x) for x in out[2:]]
    want = [engine.COMMAND_DTYPE.fields[f][1] for f in fields_c] + [engine.RESULT_DTYPE.fields[f][1] for f in fields_r]
    assert offs == want


def test_mpc_table_against_the_gait_mirror_and_a_scalar_restatement():
    h = 10
    c = _commands(stand_fraction=0.0)
    tab = F.mpc_table(c, h).reshape(len(c), h, 4)
    for b in range(len(c)):
        if c["gait_kind"][b] == F.GAIT_OFFSET_DURATION:
            ref = synth.mpc_table(h, c["gait_offsets"][b], c["gait_durations"][b], int(c["gait_iteration"][b]), h)
        else:  # MixedFrequncyGait::getMpcTable, Gait.cpp:188-215, one robot at a time
            ref = np.zeros((h, 4), dtype=np.uint8)
            for i in range(h):
                for j in range(4):
                    period = int(c["gait_offsets"][b, j])
                    progress = (i + int(c["gait_iteration"][b]) + 1) % period
                    ref[i, j] = np.float32(progress) < np.float32(period) * c["gait_duty"][b]
        assert (tab[b] == ref).all(), b
    # an offset-duration table over one full period holds every leg down for exactly its duration
    od = c["gait_kind"] == F.GAIT_OFFSET_DURATION
    assert (tab[od].sum(1) == np.minimum(c["gait_durations"][od], h)).all()


def test_reference_trajectory_rules():
    h, dt = 10, 0.03
    c = _commands()
    traj, wpd = F.reference_trajectory(c, h, dt)
    traj = traj.reshape(len(c), h, 12)
    mv = c["stand"] == 0
    # clamp: the desired position never leaves the +-0.1 m box around the robot (ConvexMPCLocomotion.cpp:534-548)
    assert (np.abs(wpd[mv] - c["position"][mv, :2]) <= np.float32(0.1) + 1e-6).all()
    inside = np.abs(c["world_position_desired"] - c["position"][:, :2]).max(1) <= 0.0999
    assert (wpd[inside] == c["world_position_desired"][inside]).all()
    # "start at current position": only the yaw starts at the measured value (:575)
    assert (traj[mv, 0, 2] == c["rpy"][mv, 2]).all() and (traj[mv, 0, 3] == wpd[mv, 0]).all()
    # the walk is a float accumulation of dt * v (:579-581), constant columns stay constant
    vw0 = traj[mv, 0, 9]
    step = (np.float32(dt) * vw0).astype(np.float32)
    for i in range(1, h):
        assert (traj[mv, i, 3] == (traj[mv, i - 1, 3] + step).astype(np.float32)).all()
    for col in (0, 1, 5, 6, 7, 8, 9, 10, 11):
        assert (traj[:, :, col] == traj[:, :1, col]).all()
    omni = mv & (c["omni_mode"] != 0)
    assert (traj[omni, 0, 9] == c["x_vel_des"][omni]).all() and (traj[omni, 0, 10] == c["y_vel_des"][omni]).all()
    # stand: the same state at every step, world_position_desired untouched (:524-531)
    st = c["stand"] != 0
    assert st.any()
    assert (traj[st, :, 3] == c["stand_traj"][st, 0][:, None]).all() and (traj[st, :, 2] == c["stand_traj"][st, 2][:, None]).all()
    assert (traj[st, :, 6:] == 0).all() and (wpd[st] == c["world_position_desired"][st]).all()


def test_external_force_against_the_dense_matrices():
    """The row-wise fp32 restatement == x_k - A_prev x_prev - B_prev u_prev with the 13x13 / 13x12 matrices of
    ConvexMPCLocomotion.cpp:650-771 built densely in float64."""
    c = _commands(B=40)
    prev = np.random.default_rng(5).normal(0, 1, (len(c), 6)).astype(np.float32)
    c["have_log"][::5] = 0
    got = F.external_force(c, prev)
    for b in range(len(c)):
        if not c["have_log"][b]:
            assert (got[b] == prev[b]).all()
            continue
        A = np.zeros((13, 13))
        A[3, 9] = A[4, 10] = A[5, 11] = A[11, 12] = 1.0
        A[11, 9] = c["log_x_drag"][b]
        R = c["log_R"][b].astype(np.float64).reshape(3, 3)
        A[0:3, 6:9] = R.T
        Iw = R @ np.diag([0.07, 0.26, 0.242]) @ R.T
        Ii = np.linalg.inv(Iw)
        Bm = np.zeros((13, 12))
        rf = c["log_r_feet"][b].astype(np.float64).reshape(3, 4)
        for leg in range(4):
            r = rf[:, leg]
            cm = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
            Bm[6:9, 3 * leg:3 * leg + 3] = Ii @ cm
            Bm[9:12, 3 * leg:3 * leg + 3] = np.eye(3) / 12.0
        xk = np.concatenate([c["rpy"][b], [c["position"][b, 0], c["position"][b, 1], c["ground_z"][b]],
                             c["omega_world"][b], c["v_world"][b], [-9.81]]).astype(np.float64)
        xp = np.concatenate([c["log_x_prev"][b], [-9.81]]).astype(np.float64)
        u = -c["log_foot_force"][b].astype(np.float64)
        fx = (xk - A @ xp - Bm @ u)[6:12]
        ref = np.array([-fx[0], -fx[1], fx[2], fx[3], fx[4], fx[5]])
        np.testing.assert_allclose(got[b], ref, rtol=2e-4, atol=2e-4)


def test_solver_inputs_and_leg_outputs():
    h, dt = 10, 0.03
    c = _commands(B=32)
    inst, ex = F.solver_inputs(c, h, dt, np.zeros((len(c), 6), np.float32))
    pf = c["p_foot"].reshape(-1, 4, 3)
    for i in range(12):   # r[i] = pFoot[i % 4][i / 4] - position[i / 4]  (:779)
        assert (inst["r"][:, i] == pf[:, i % 4, i // 4] - c["position"][:, i // 4]).all()
    assert (inst["p"][:, 2] == c["ground_z"]).all() and (inst["x_drag"] == c["x_comp_integral"]).all()
    fast = np.abs(c["v_world"][:, 0]) > 0.3
    assert (ex["x_comp_integral"][~fast] == c["x_comp_integral"][~fast]).all() and fast.any()
    assert (ex["x_comp_integral"][fast] != c["x_comp_integral"][fast]).any()
    rec = F.pack_records(inst)
    assert rec.shape == (len(c), 720)
    forces = np.random.default_rng(2).normal(0, 30, (len(c), 12 * h))
    f, f_ff = F.leg_outputs(c, forces)
    R = c["r_body"].astype(np.float64).reshape(-1, 3, 3)
    ref = -np.einsum("bij,blj->bli", R, forces[:, :12].reshape(-1, 4, 3)).reshape(-1, 12)
    np.testing.assert_allclose(f_ff, ref, rtol=1e-5, atol=1e-4)
    assert (f == forces[:, :12].astype(np.float32)).all()
