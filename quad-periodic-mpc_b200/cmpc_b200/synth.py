"""Synthetic randomised A1 MPC instances (states, gaits, reference trajectories,
periodic disturbances) in the shapes the reference feeds solve_mpc().

Field meaning follows update_data_t (convexMPC_interface.h:23-42):
  p,v,w  world position / velocity / angular velocity           float32 [B,3]
  q      orientation quaternion (w,x,y,z)                        float32 [B,4]
  r      foot positions relative to the COM, r[axis*4+leg]       float32 [B,12]
  rpy    roll,pitch,yaw (carried, unused by the dense solver)    float32 [B,3]
  weights 12 state weights, traj 12*h reference states           float32
  gait   contact table, gait[step*4+leg] in {0,1}                uint8 [B,4h]
"""
import numpy as np

# ConvexMPCLocomotion.cpp:617, :623 and :62
A1_WEIGHTS = np.array([0.25, 0.25, 10, 10, 2, 50, 0, 0, 0.3, 0.2, 0.2, 0.1], dtype=np.float32)
A1_ALPHA = 4e-5
A1_MU = 0.4
A1_FMAX = 120.0
A1_MASS = 12.0
A1_INERTIA = (0.07, 0.26, 0.242)

# (offsets, durations) in tenths of the gait period; ConvexMPCLocomotion.cpp:41-50
GAITS = {
    "trot": ((0, 5, 5, 0), (5, 5, 5, 5)),
    "bound": ((5, 5, 0, 0), (4, 4, 4, 4)),
    "pronk": ((0, 0, 0, 0), (8, 8, 8, 8)),
    "pace": ((5, 0, 5, 0), (5, 5, 5, 5)),
    "gallop": ((0, 2, 7, 9), (4, 4, 4, 4)),
    "trotrun": ((0, 5, 5, 0), (4, 4, 4, 4)),
    "walk2": ((0, 5, 5, 0), (7, 7, 7, 7)),
    "stand": ((0, 0, 0, 0), (10, 10, 10, 10)),
}


def mpc_table(n_segment, offsets, durations, iteration, horizon):
    """OffsetDurationGait::getMpcTable, Gait.cpp:158-187, first `horizon` rows."""
    tab = np.zeros((horizon, 4), dtype=np.uint8)
    for i in range(horizon):
        it = (i + iteration + 1) % n_segment
        for j in range(4):
            prog = it - offsets[j]
            if prog < 0:
                prog += n_segment
            tab[i, j] = 1 if prog < durations[j] else 0
    return tab


def scaled_gait(name, n_segment):
    off, dur = GAITS[name]
    s = n_segment / 10.0
    return tuple(int(o * s) for o in off), tuple(max(1, int(d * s)) for d in dur)


def quat_from_rpy(roll, pitch, yaw):
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], axis=-1)


def make_batch(batch, horizon=10, dt=0.03, gaits=("trot",), seed=0, n_segment=None, spread=1.0,
               body_height=0.24):
    """Randomised A1 instances.  `spread` scales the state perturbations
    (1.0 = brisk locomotion with pushes; larger activates more friction-cone faces)."""
    rng = np.random.default_rng(seed)
    B, h = batch, horizon
    nseg = n_segment or h
    f32 = np.float32
    roll = rng.normal(0, 0.08 * spread, B)
    pitch = rng.normal(0, 0.08 * spread, B)
    yaw = rng.uniform(-np.pi, np.pi, B)
    q = quat_from_rpy(roll, pitch, yaw)
    p = np.stack([rng.uniform(-2, 2, B), rng.uniform(-2, 2, B), body_height + rng.normal(0, 0.02 * spread, B)], -1)
    v = np.stack([rng.normal(0, 0.4 * spread, B), rng.normal(0, 0.25 * spread, B), rng.normal(0, 0.1 * spread, B)], -1)
    w = rng.normal(0, 0.4 * spread, (B, 3))
    # A1 hips (+/-0.1805, +/-0.047) + abad 0.0838; leg order FR, FL, RR, RL
    hips = np.array([[0.1805, -0.1308], [0.1805, 0.1308], [-0.1805, -0.1308], [-0.1805, 0.1308]])
    cy, sy = np.cos(yaw), np.sin(yaw)
    r = np.zeros((B, 12))
    for leg in range(4):
        fx = hips[leg, 0] + rng.normal(0, 0.04 * spread, B)
        fy = hips[leg, 1] + rng.normal(0, 0.03 * spread, B)
        r[:, 0 * 4 + leg] = cy * fx - sy * fy
        r[:, 1 * 4 + leg] = sy * fx + cy * fy
        r[:, 2 * 4 + leg] = -p[:, 2] + rng.normal(0, 0.005 * spread, B)
    # reference trajectory, ConvexMPCLocomotion.cpp:554-585
    vx_des = rng.uniform(-0.7, 0.7, B)
    vy_des = rng.uniform(-0.4, 0.4, B)
    yaw_rate = rng.uniform(-1.0, 1.0, B)
    vdw = np.stack([cy * vx_des - sy * vy_des, sy * vx_des + cy * vy_des], -1)
    traj = np.zeros((B, 12 * h))
    x_start = p[:, 0] + rng.uniform(-0.1, 0.1, B)
    y_start = p[:, 1] + rng.uniform(-0.1, 0.1, B)
    init = np.zeros((B, 12))
    init[:, 0] = rng.normal(0, 0.02, B)
    init[:, 1] = rng.normal(0, 0.02, B)
    init[:, 2] = yaw
    init[:, 3], init[:, 4], init[:, 5] = x_start, y_start, body_height
    init[:, 8] = yaw_rate
    init[:, 9], init[:, 10] = vdw[:, 0], vdw[:, 1]
    for i in range(h):
        traj[:, 12 * i:12 * i + 12] = init
        if i > 0:
            traj[:, 12 * i + 3] = traj[:, 12 * (i - 1) + 3] + dt * vdw[:, 0]
            traj[:, 12 * i + 4] = traj[:, 12 * (i - 1) + 4] + dt * vdw[:, 1]
            traj[:, 12 * i + 2] = traj[:, 12 * (i - 1) + 2] + dt * yaw_rate
    gait = np.zeros((B, 4 * h), dtype=np.uint8)
    gait_id = rng.integers(0, len(gaits), B)
    iters = rng.integers(0, nseg, B)
    cache = {}
    for b in range(B):
        key = (int(gait_id[b]), int(iters[b]))
        if key not in cache:
            off, dur = scaled_gait(gaits[key[0]], nseg)
            cache[key] = mpc_table(nseg, off, dur, key[1], h).reshape(-1)
        gait[b] = cache[key]
    return {
        "p": p.astype(f32), "v": v.astype(f32), "q": q.astype(f32), "w": w.astype(f32), "r": r.astype(f32),
        "rpy": np.stack([roll, pitch, yaw], -1).astype(f32),
        "weights": np.tile(A1_WEIGHTS, (B, 1)), "traj": traj.astype(f32),
        "alpha": np.full(B, A1_ALPHA, f32), "gait": gait,
        "x_drag": rng.normal(0, 0.05, B).astype(f32),
        "horizon": h, "dt": dt, "mu": A1_MU, "f_max": A1_FMAX,
    }


def make_disturbance_windows(batch, n=400, dt=0.03, seed=0):
    """Randomised periodic disturbance histories (amplitude/frequency/offset + noise):
    the f_ext[3] samples the reference accumulates in diff_history (SolverMPC.cpp:692)."""
    rng = np.random.default_rng(seed + 7919)
    t0 = rng.uniform(0, 5, (batch, 1))
    t = t0 + dt * np.arange(n)[None, :]
    amp = rng.uniform(0.2, 3.0, (batch, 1))
    freq = rng.uniform(0.15, 1.5, (batch, 1))
    phase = rng.uniform(-np.pi, np.pi, (batch, 1))
    off = rng.normal(0, 0.5, (batch, 1))
    d = off + amp * np.sin(2 * np.pi * freq * t + phase) + rng.normal(0, 0.05, (batch, n))
    return t.astype(np.float32), d.astype(np.float32), dict(amp=amp, freq=freq, phase=phase, off=off)


def rot_from_quat(q):
    """Rotation matrix of the estimator (rBody: body <- world), row major, from (w,x,y,z)."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y),
                  2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x),
                  2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], axis=-1)
    return R


def make_commands(batch, dtype, horizon=10, gaits=("trot",), seed=0, spread=1.0, body_height=0.24, mixed_fraction=0.0,
                  stand_fraction=0.0, with_log=True, sim_time=0.0, disturbance=None):
    """Randomised controller-level inputs (`cmpc_command`, include/cmpc_b200.h): what ConvexMPCLocomotion's
    updateMPCIfNeeded / solveDenseMPC read per robot.  `dtype` is engine.COMMAND_DTYPE.
    disturbance = (amp, freq, phase) arrays adds a periodic force residual to the logged model so that f_ext[3]
    oscillates (the signal the estimator fits)."""
    rng = np.random.default_rng(seed + 4241)
    B, h = batch, horizon
    base = make_batch(B, horizon=h, gaits=gaits, seed=seed, spread=spread, body_height=body_height)
    c = np.zeros(B, dtype=dtype)
    p, q = base["p"], base["q"]
    c["position"] = p
    c["ground_z"] = p[:, 2] + rng.normal(0, 0.003, B)
    c["v_world"], c["omega_world"], c["orientation"], c["rpy"] = base["v"], base["w"], q, base["rpy"]
    c["r_body"] = rot_from_quat(q.astype(np.float64))
    r = base["r"].reshape(B, 3, 4)                                    # r[axis][leg] = pFoot[leg][axis] - position[axis]
    c["p_foot"] = (r.transpose(0, 2, 1) + p[:, None, :]).reshape(B, 12)
    c["x_vel_des"], c["y_vel_des"] = rng.uniform(-0.7, 0.7, B), rng.uniform(-0.4, 0.4, B)
    c["yaw_turn_rate"] = rng.uniform(-1.0, 1.0, B)
    c["yaw_des"] = base["rpy"][:, 2] + rng.normal(0, 0.05, B)
    c["body_height"] = body_height
    c["rpy_comp"] = rng.normal(0, 0.02, (B, 2))
    c["world_position_desired"] = p[:, :2] + rng.uniform(-0.25, 0.25, (B, 2))   # some beyond max_pos_error
    c["roll_des"], c["pitch_des"] = rng.normal(0, 0.02, B), rng.normal(0, 0.02, B)
    c["stand_traj"] = np.stack([p[:, 0], p[:, 1], base["rpy"][:, 2]], -1)
    c["x_comp_integral"] = rng.normal(0, 0.05, B)
    c["cmpc_x_drag"] = 3.0
    names = list(gaits)
    gid = rng.integers(0, len(names), B)
    for b in range(B):
        off, dur = scaled_gait(names[gid[b]], h)
        c["gait_offsets"][b], c["gait_durations"][b] = off, dur
    c["gait_iteration"] = rng.integers(0, h, B)
    mixed = rng.random(B) < mixed_fraction
    c["gait_kind"] = mixed.astype(np.int32)
    periods = rng.integers(min(3, h), h + 1, (B, 4))
    c["gait_offsets"][mixed] = periods[mixed]
    c["gait_duty"] = rng.uniform(0.35, 0.65, B)
    c["omni_mode"] = rng.random(B) < 0.25
    c["stand"] = rng.random(B) < stand_fraction
    c["gait_durations"][c["stand"] != 0] = h
    c["gait_offsets"][(c["stand"] != 0) & ~mixed] = 0
    c["gait_kind"][c["stand"] != 0] = 0
    c["have_log"] = 1 if with_log else 0
    c["log_x_prev"] = np.concatenate([base["rpy"], p, base["w"], base["v"]], axis=1) + rng.normal(0, 0.01, (B, 12))
    c["log_R"] = rot_from_quat(q.astype(np.float64)).reshape(B, 3, 3).transpose(0, 2, 1).reshape(B, 9)
    c["log_r_feet"] = base["r"] + rng.normal(0, 0.005, (B, 12))
    ff = rng.normal(0, 8.0, (B, 12))
    ff[:, 2::3] = rng.uniform(20, 70, (B, 4))
    if disturbance is not None:
        amp, freq, phase = disturbance
        ff[:, 0] += 12.0 * (amp * np.sin(2 * np.pi * freq * sim_time + phase))   # shows up in f_ext[3] = v_x - sum u_x / 12
    c["log_foot_force"] = ff
    c["log_x_drag"] = c["x_comp_integral"]
    c["sim_time"] = sim_time
    return c
