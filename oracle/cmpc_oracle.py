"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/_ref/libcmpc_ref.so (the restated condensation of
SolverMPC.cpp + the reference's real qpOASES 3.2.0, see cmpc_oracle.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libcmpc_ref.so")
MAX_SEG = 36
HIST_MAX = 4096


class Setup(C.Structure):
    _fields_ = [("dt", C.c_float), ("mu", C.c_float), ("f_max", C.c_float), ("horizon", C.c_int),
                ("mass", C.c_float), ("inertia", C.c_float * 3), ("nwsr", C.c_int)]


class Update(C.Structure):
    _fields_ = [("p", C.c_float * 3), ("v", C.c_float * 3), ("q", C.c_float * 4), ("w", C.c_float * 3),
                ("r", C.c_float * 12), ("roll", C.c_float), ("pitch", C.c_float), ("yaw", C.c_float),
                ("weights", C.c_float * 12), ("traj", C.c_float * (12 * MAX_SEG)), ("alpha", C.c_float),
                ("gait", C.c_ubyte * (4 * MAX_SEG)), ("x_drag", C.c_float)]


class Result(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_double)), ("y_con", C.POINTER(C.c_double)),
                ("con_status", C.POINTER(C.c_int8)), ("var_elim", C.POINTER(C.c_int8)),
                ("H_full", C.POINTER(C.c_double)), ("g_full", C.POINTER(C.c_double)),
                ("H_red", C.POINTER(C.c_double)), ("g_red", C.POINTER(C.c_double)),
                ("AdtBdtQdt", C.POINTER(C.c_double)),
                ("n_var", C.c_int), ("n_con", C.c_int), ("nwsr", C.c_int), ("qp_return", C.c_int),
                ("qp_status_ok", C.c_int), ("objective", C.c_double)]


class Adapt(C.Structure):
    _fields_ = [("len", C.c_int), ("t_hist", C.c_float * HIST_MAX), ("d_hist", C.c_float * HIST_MAX),
                ("est", C.c_double * 4), ("f_est", C.c_float * 6), ("f_est_smoothed", C.c_float * 6),
                ("f_est_static3", C.c_float)]


def build(quiet=True):
    """Build oracle/_ref when the reference sources are present (this container)."""
    if os.path.isdir("/root/reference/be2r_cmpc_unitree/src/third_party/qpOASES/src"):
        subprocess.run(["make", "-j8", "-C", _HERE], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)
    return os.path.exists(_LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.cmpc_oracle_solve.argtypes = [C.POINTER(Setup), C.POINTER(Update), C.POINTER(C.c_double), C.c_int,
                                           C.POINTER(Result)]
        _lib.cmpc_oracle_fit_window.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                                C.POINTER(C.c_double)]
        _lib.cmpc_oracle_adapt_step.argtypes = [C.POINTER(Adapt), C.c_double, C.c_double, C.POINTER(C.c_double)]
        _lib.cmpc_oracle_solve_batch.argtypes = [C.POINTER(Setup), C.POINTER(Update), C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(C.c_double), C.POINTER(C.c_int)]
    return _lib


def available():
    return os.path.exists(_LIB_PATH) or build()


def make_setup(dt, horizon, mu, f_max, mass=12.0, inertia=(0.07, 0.26, 0.242), nwsr=100):
    s = Setup()
    s.dt, s.mu, s.f_max, s.horizon, s.mass, s.nwsr = dt, mu, f_max, horizon, mass, nwsr
    s.inertia[:] = inertia
    return s


def make_update(inst, i, horizon):
    """inst: dict of batch arrays as produced by cmpc_b200.synth (float32 fields)."""
    u = Update()
    u.p[:] = inst["p"][i]
    u.v[:] = inst["v"][i]
    u.q[:] = inst["q"][i]
    u.w[:] = inst["w"][i]
    u.r[:] = inst["r"][i]
    u.roll, u.pitch, u.yaw = [float(x) for x in inst["rpy"][i]]
    u.weights[:] = inst["weights"][i]
    u.traj[:12 * horizon] = inst["traj"][i][:12 * horizon]
    u.alpha = float(inst["alpha"][i])
    u.gait[:4 * horizon] = [int(x) for x in inst["gait"][i][:4 * horizon]]
    u.x_drag = float(inst["x_drag"][i])
    return u


def solve(setup, update, f_dist=None, use_float=False, want_mats=False):
    h = setup.horizon
    n, m = 12 * h, 20 * h
    x = np.zeros(n)
    y = np.zeros(m)
    cs = np.zeros(m, dtype=np.int8)
    ve = np.zeros(n, dtype=np.int8)
    r = Result()
    r.x = x.ctypes.data_as(C.POINTER(C.c_double))
    r.y_con = y.ctypes.data_as(C.POINTER(C.c_double))
    r.con_status = cs.ctypes.data_as(C.POINTER(C.c_int8))
    r.var_elim = ve.ctypes.data_as(C.POINTER(C.c_int8))
    out = {}
    if want_mats:
        Hf, gf, Hr, gr, E = np.zeros((n, n)), np.zeros(n), np.zeros(n * n), np.zeros(n), np.zeros((13, 31))
        r.H_full = Hf.ctypes.data_as(C.POINTER(C.c_double))
        r.g_full = gf.ctypes.data_as(C.POINTER(C.c_double))
        r.H_red = Hr.ctypes.data_as(C.POINTER(C.c_double))
        r.g_red = gr.ctypes.data_as(C.POINTER(C.c_double))
        r.AdtBdtQdt = E.ctypes.data_as(C.POINTER(C.c_double))
    fd = None
    if f_dist is not None:
        fdarr = np.ascontiguousarray(f_dist, dtype=np.float64)
        fd = fdarr.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib().cmpc_oracle_solve(C.byref(setup), C.byref(update), fd, int(use_float), C.byref(r))
    if rc != 0:
        raise RuntimeError("cmpc_oracle_solve failed rc=%d" % rc)
    out.update(x=x, y_con=y, con_status=cs, var_elim=ve, n_var=r.n_var, n_con=r.n_con, nwsr=r.nwsr,
               qp_return=r.qp_return, ok=bool(r.qp_status_ok), objective=r.objective)
    if want_mats:
        nv = r.n_var
        out.update(H_full=Hf, g_full=gf, H_red=Hr[:nv * nv].reshape(nv, nv).copy(), g_red=gr[:nv].copy(),
                   Adt=E[:, :13].copy(), Bdt=E[:, 13:25].copy(), Qdt=E[:, 25:31].copy())
    return out


def solve_batch(setup, updates, threads, use_float=False):
    """updates: ctypes array of Update. Returns (forces[count,12h], ok[count])."""
    count = len(updates)
    h = setup.horizon
    forces = np.zeros((count, 12 * h))
    ok = np.zeros(count, dtype=np.int32)
    lib().cmpc_oracle_solve_batch(C.byref(setup), updates, count, int(use_float), threads,
                                  forces.ctypes.data_as(C.POINTER(C.c_double)),
                                  ok.ctypes.data_as(C.POINTER(C.c_int)))
    return forces, ok


def fit_window(t, d):
    t = np.ascontiguousarray(t, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    out = np.zeros(4)
    rc = lib().cmpc_oracle_fit_window(t.ctypes.data_as(C.POINTER(C.c_double)),
                                      d.ctypes.data_as(C.POINTER(C.c_double)), len(t),
                                      out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise RuntimeError("fit_window failed")
    return out
