"""ORACLE (test infrastructure, not product code): CPU restatement of the CALLER of the solve path —
ConvexMPCLocomotion::updateMPCIfNeeded + solveDenseMPC (ConvexMPCLocomotion.cpp:511-870) and the gaits'
getMpcTable (Gait.cpp:158-215) — in numpy float32, vectorised over robots.

Parity status: UNPINNED by a reference run (the reference needs Eigen and ROS, absent from this image, and holds no
tests or vectors for these functions).  The restatement follows the reference statement by statement; every fp32
operation is a separately rounded numpy float32 operation in the order written here, which is the order
csrc/cmpc_frontend.cu uses, so the GPU records are compared bit for bit.  Eigen's own summation order for the 3x3
products is not knowable without building the reference; the difference is at fp32 rounding level.

Input: a numpy structured array with the fields of `cmpc_command` (include/cmpc_b200.h).
"""
import numpy as np

f32 = np.float32
Q_WEIGHTS = np.array([0.25, 0.25, 10, 10, 2, 50, 0, 0, 0.3, 0.2, 0.2, 0.1], dtype=f32)  # ConvexMPCLocomotion.cpp:627
ALPHA = f32(4e-5)                                                                        # :634
GAIT_OFFSET_DURATION, GAIT_MIXED_FREQUENCY = 0, 1


def _dot3(a0, b0, a1, b1, a2, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def mpc_table(cmd, h):
    """OffsetDurationGait::getMpcTable (Gait.cpp:158-186) / MixedFrequncyGait::getMpcTable (:188-215), nIterations = h."""
    B = len(cmd)
    tab = np.zeros((B, h, 4), dtype=np.uint8)
    it = cmd["gait_iteration"].astype(np.int64)
    off = cmd["gait_offsets"].astype(np.int64)
    dur = cmd["gait_durations"].astype(np.int64)
    mixed = cmd["gait_kind"] == GAIT_MIXED_FREQUENCY
    for i in range(h):
        for j in range(4):
            it_ = (i + it + 1) % h                       # :162
            progress = it_ - off[:, j]                   # :163
            progress = np.where(progress < 0, progress + h, progress)
            on = progress < dur[:, j]                    # :171
            period = np.where(off[:, j] > 0, off[:, j], 1)
            pm = (i + it + 1) % period                   # :196
            onm = pm.astype(f32) < period.astype(f32) * cmd["gait_duty"].astype(f32)   # :197
            tab[:, i, j] = np.where(mixed, onm, on)
    return tab.reshape(B, 4 * h)


def reference_trajectory(cmd, h, dt_mpc):
    """updateMPCIfNeeded (:511-594): trajAll and the clamped world_position_desired."""
    B = len(cmd)
    dt = f32(dt_mpc)
    traj = np.zeros((B, h, 12), dtype=f32)
    R = cmd["r_body"].astype(f32)
    xv, yv = cmd["x_vel_des"].astype(f32), cmd["y_vel_des"].astype(f32)
    zero = np.zeros(B, dtype=f32)
    vw0 = _dot3(R[:, 0], xv, R[:, 3], yv, R[:, 6], zero)      # rBody^T v_des_robot (:518)
    vw1 = _dot3(R[:, 1], xv, R[:, 4], yv, R[:, 7], zero)
    omni = cmd["omni_mode"] != 0
    vw0 = np.where(omni, xv, vw0).astype(f32)
    vw1 = np.where(omni, yv, vw1).astype(f32)
    mpe = f32(0.1)                                             # max_pos_error (:534)
    p0, p1 = cmd["position"][:, 0].astype(f32), cmd["position"][:, 1].astype(f32)
    xs, ys = cmd["world_position_desired"][:, 0].astype(f32), cmd["world_position_desired"][:, 1].astype(f32)
    xs = np.where(xs - p0 > mpe, p0 + mpe, xs).astype(f32)     # :538-541
    xs = np.where(p0 - xs > mpe, p0 - mpe, xs).astype(f32)
    ys = np.where(ys - p1 > mpe, p1 + mpe, ys).astype(f32)     # :543-546
    ys = np.where(p1 - ys > mpe, p1 - mpe, ys).astype(f32)
    t0 = np.stack([cmd["rpy_comp"][:, 0], cmd["rpy_comp"][:, 1], cmd["yaw_des"], xs, ys, cmd["body_height"],
                   zero, zero, cmd["yaw_turn_rate"], vw0, vw1, zero], axis=1).astype(f32)     # :551-563
    dx, dy, dyaw = dt * vw0, dt * vw1, dt * cmd["yaw_turn_rate"].astype(f32)
    x, y, yaw = xs.copy(), ys.copy(), cmd["rpy"][:, 2].astype(f32).copy()                   # :575 trajAll[2] = rpy[2]
    for i in range(h):
        if i > 0:                                                                               # :579-581
            x = (x + dx).astype(f32)
            y = (y + dy).astype(f32)
            yaw = (yaw + dyaw).astype(f32)
        traj[:, i, :] = t0
        traj[:, i, 2], traj[:, i, 3], traj[:, i, 4] = yaw, x, y
    # stand gait (:524-531)
    st = cmd["stand"] != 0
    s0 = np.stack([cmd["roll_des"], cmd["pitch_des"], cmd["stand_traj"][:, 2], cmd["stand_traj"][:, 0],
                   cmd["stand_traj"][:, 1], cmd["body_height"], zero, zero, zero, zero, zero, zero], axis=1).astype(f32)
    traj[st] = s0[st][:, None, :]
    wpd = np.stack([xs, ys], axis=1)
    wpd[st] = cmd["world_position_desired"][st]
    return traj.reshape(B, 12 * h), wpd.astype(f32)


def external_force(cmd, f_ext_prev):
    """solveDenseMPC (:647-776): f_external = x_k - A_prev x_prev - B_prev u_prev, rows 6..11, signs of :771."""
    B = len(cmd)
    R = cmd["log_R"].astype(f32)
    Ib = np.array([0.07, 0.26, 0.242], dtype=f32)
    Iw = np.zeros((B, 9), dtype=f32)
    for a in range(3):
        for b in range(3):
            Iw[:, 3 * a + b] = _dot3(R[:, 3 * a] * Ib[0], R[:, 3 * b], R[:, 3 * a + 1] * Ib[1], R[:, 3 * b + 1],
                                     R[:, 3 * a + 2] * Ib[2], R[:, 3 * b + 2])
    c00 = Iw[:, 4] * Iw[:, 8] - Iw[:, 5] * Iw[:, 7]
    c01 = Iw[:, 5] * Iw[:, 6] - Iw[:, 3] * Iw[:, 8]
    c02 = Iw[:, 3] * Iw[:, 7] - Iw[:, 4] * Iw[:, 6]
    det = (Iw[:, 0] * c00 + Iw[:, 1] * c01) + Iw[:, 2] * c02
    with np.errstate(divide="ignore", invalid="ignore"):
        idet = (f32(1.0) / det).astype(f32)
    Ii = np.stack([c00 * idet,
                   (Iw[:, 2] * Iw[:, 7] - Iw[:, 1] * Iw[:, 8]) * idet,
                   (Iw[:, 1] * Iw[:, 5] - Iw[:, 2] * Iw[:, 4]) * idet,
                   c01 * idet,
                   (Iw[:, 0] * Iw[:, 8] - Iw[:, 2] * Iw[:, 6]) * idet,
                   (Iw[:, 2] * Iw[:, 3] - Iw[:, 0] * Iw[:, 5]) * idet,
                   c02 * idet,
                   (Iw[:, 1] * Iw[:, 6] - Iw[:, 0] * Iw[:, 7]) * idet,
                   (Iw[:, 0] * Iw[:, 4] - Iw[:, 1] * Iw[:, 3]) * idet], axis=1).astype(f32)
    bu_w = np.zeros((B, 3), dtype=f32)
    bu_v = np.zeros((B, 3), dtype=f32)
    minv = f32(1.0) / f32(12.0)
    ff, rf = cmd["log_foot_force"].astype(f32), cmd["log_r_feet"].astype(f32)
    for leg in range(4):
        u0, u1, u2 = -ff[:, 3 * leg], -ff[:, 3 * leg + 1], -ff[:, 3 * leg + 2]       # u_prev = -foot_force (:748-759)
        rx, ry, rz = rf[:, leg], rf[:, 4 + leg], rf[:, 8 + leg]
        t0 = ry * u2 - rz * u1
        t1 = rz * u0 - rx * u2
        t2 = rx * u1 - ry * u0
        for a in range(3):
            bu_w[:, a] = bu_w[:, a] + _dot3(Ii[:, 3 * a], t0, Ii[:, 3 * a + 1], t1, Ii[:, 3 * a + 2], t2)
        bu_v[:, 0] = bu_v[:, 0] + minv * u0
        bu_v[:, 1] = bu_v[:, 1] + minv * u1
        bu_v[:, 2] = bu_v[:, 2] + minv * u2
    a11 = cmd["log_x_drag"].astype(f32) * cmd["log_x_prev"][:, 9].astype(f32) + f32(-9.81)
    w, v = cmd["omega_world"].astype(f32), cmd["v_world"].astype(f32)
    f6 = np.stack([w[:, 0] - bu_w[:, 0], w[:, 1] - bu_w[:, 1], w[:, 2] - bu_w[:, 2],
                   v[:, 0] - bu_v[:, 0], v[:, 1] - bu_v[:, 1], (v[:, 2] - a11) - bu_v[:, 2]], axis=1).astype(f32)
    fe = np.stack([-f6[:, 0], -f6[:, 1], f6[:, 2], f6[:, 3], f6[:, 4], f6[:, 5]], axis=1).astype(f32)   # :771
    have = cmd["have_log"] != 0
    out = np.array(f_ext_prev, dtype=f32, copy=True)
    out[have] = fe[have]
    return out


def solver_inputs(cmd, h, dt_mpc, f_ext_prev, weights=Q_WEIGHTS, alpha=ALPHA):
    """What solveDenseMPC hands update_problem_data_floats (:779-828), plus the command state it leaves behind.
    Returns (inputs dict in the layout of cmpc_inputs, extras dict)."""
    B = len(cmd)
    traj, wpd = reference_trajectory(cmd, h, dt_mpc)
    gait = mpc_table(cmd, h)
    f_ext = external_force(cmd, f_ext_prev)
    pos = cmd["position"].astype(f32)
    p = np.stack([pos[:, 0], pos[:, 1], cmd["ground_z"].astype(f32)], axis=1)                # p_v (:640)
    pf = cmd["p_foot"].astype(f32).reshape(B, 4, 3)
    r = np.zeros((B, 12), dtype=f32)
    for i in range(12):
        r[:, i] = pf[:, i % 4, i // 4] - pos[:, i // 4]                                       # :779
    a = f32(alpha)
    if a > f32(1e-4):                                                                          # :785-789
        a = f32(1e-5)
    xci = cmd["x_comp_integral"].astype(f32)
    vx = cmd["v_world"][:, 0].astype(f32)
    pz_err = cmd["ground_z"].astype(f32) - cmd["body_height"].astype(f32)                      # :794
    with np.errstate(divide="ignore", invalid="ignore"):
        upd = (xci + ((cmd["cmpc_x_drag"].astype(f32) * pz_err) * f32(dt_mpc)) / vx).astype(f32)
    xci_new = np.where((vx > f32(0.3)) | (vx < f32(-0.3)), upd, xci).astype(f32)              # :811-816
    inst = {
        "p": p, "v": cmd["v_world"].astype(f32), "q": cmd["orientation"].astype(f32), "w": cmd["omega_world"].astype(f32),
        "r": r, "rpy": cmd["rpy"].astype(f32), "weights": np.tile(np.asarray(weights, dtype=f32), (B, 1)), "traj": traj,
        "alpha": np.full(B, a, dtype=f32), "gait": gait, "x_drag": xci,                       # update_x_drag(old integral) :809
        "horizon": h, "dt": dt_mpc,
    }
    extras = {"world_position_desired": wpd, "x_comp_integral": xci_new, "f_ext": f_ext}
    return inst, extras


def leg_outputs(cmd, forces):
    """:833-845: f = (float) get_solution(leg*3+axis); f_ff[leg] = -rBody * f; Fr_des[leg] = f."""
    B = len(cmd)
    R = cmd["r_body"].astype(f32)
    f = np.asarray(forces)[:, :12].astype(f32)
    f_ff = np.zeros((B, 12), dtype=f32)
    for leg in range(4):
        f0, f1, f2 = f[:, 3 * leg], f[:, 3 * leg + 1], f[:, 3 * leg + 2]
        for a in range(3):
            f_ff[:, 3 * leg + a] = -_dot3(R[:, 3 * a], f0, R[:, 3 * a + 1], f1, R[:, 3 * a + 2], f2)
    return f, f_ff


def pack_records(inst, f_dist=None, sim_time=None):
    """The instance records as csrc/cmpc_device.h lays them out (what the GPU front end writes), for bit comparisons."""
    h = inst["horizon"]
    B = len(inst["p"])
    stride = (4 * (48 + 12 * h) + 4 * h + 15) & ~15
    rec = np.zeros((B, stride), dtype=np.uint8)
    fl = rec[:, :4 * (48 + 12 * h)].view(f32)
    fl[:, 0:3], fl[:, 3:6], fl[:, 6:10], fl[:, 10:13] = inst["p"], inst["v"], inst["q"], inst["w"]
    fl[:, 13:25], fl[:, 25:37] = inst["r"], inst["weights"]
    fl[:, 37], fl[:, 38] = inst["alpha"], inst["x_drag"]
    if f_dist is not None:
        fl[:, 39:45] = f_dist
    if sim_time is not None:
        fl[:, 45] = sim_time
    fl[:, 48:48 + 12 * h] = inst["traj"]
    rec[:, 4 * (48 + 12 * h):4 * (48 + 12 * h) + 4 * h] = inst["gait"]
    return rec
