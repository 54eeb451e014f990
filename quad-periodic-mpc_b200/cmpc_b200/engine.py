"""ctypes binding of the C-ABI in include/cmpc_b200.h (libcmpc_b200.so).

There is no CPU path: every call below ends in the CUDA library, and loading
or using it without a usable B200 raises.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("CMPC_LIB") or os.path.join(_PKG, "libcmpc_b200.so")  # CMPC_LIB: experiment builds

ST_SOLVED, ST_EMPTY, ST_MAXITER, ST_INFEASIBLE, ST_WSOVERFLOW, ST_CAPACITY, ST_NONFINITE = range(7)
ON_ERROR_ABORT, ON_ERROR_HOLD = 0, 1

# The library reads no environment variable on a solve path; its diagnostic / test switches exist only through
# cmpc_batch_set_option.  For the scripts and tests that select kernel paths from the shell, THIS test-side binding
# maps the historical variable names onto options when a Batch is created.
_ENV_OPTIONS = {
    "CMPC_NSTREAMS": ("nstreams", int), "CMPC_SPLIT": ("split", int), "CMPC_SERIAL": ("serial", int),
    "CMPC_LPT": ("lpt", int), "CMPC_RESUME": ("resume", int), "CMPC_INV_STAGGER": ("inv_stagger", int),
    "CMPC_INV_REFINE": ("inv_refine", int), "CMPC_INV_F32": ("inv_f32", int), "CMPC_INV_CTAS": ("inv_ctas", int),
    "CMPC_TRAJ_COPY": ("traj_copy", int), "CMPC_CSHAPE": ("cshape", int), "CMPC_WS_MB": ("ws_mb", int),
    "CMPC_QCAP1": ("qcap1", int), "CMPC_WPC": ("wpc", int), "CMPC_NO_MID_TIER": ("no_mid_tier", lambda v: 1),
    "CMPC_SHAPE": ("shape", int), "CMPC_HOST_PACK": ("host_pack", int), "CMPC_D2H_COPY": ("d2h_copy", int),
    "CMPC_CHUNKS": ("chunks", int), "CMPC_SUBMIT_COPY": ("submit_copy", int), "CMPC_HOST_THREADS": ("host_threads", int),
    "CMPC_PATH": ("path_fused", lambda v: int(v == "fused")), "CMPC_DUAL": ("dual_generic", lambda v: int(v == "generic")),
    "CMPC_SWEEP": ("sweep_dmma", lambda v: int(v == "dmma")), "CMPC_EXP_SKIP_PACK": ("exp_skip_pack", lambda v: 1),
    "CMPC_DUAL_TEAM": ("dual_team", int),
}


class Inputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("p", "v", "q", "w", "r", "weights", "traj", "alpha", "gait", "x_drag", "f_dist")]


class Outputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("forces", "objective", "status", "iterations", "active")]


# cmpc_command / cmpc_command_result (include/cmpc_b200.h), as numpy structured dtypes
COMMAND_DTYPE = np.dtype([
    ("position", "<f4", 3), ("ground_z", "<f4"), ("v_world", "<f4", 3), ("omega_world", "<f4", 3),
    ("orientation", "<f4", 4), ("rpy", "<f4", 3), ("r_body", "<f4", 9), ("p_foot", "<f4", 12),
    ("x_vel_des", "<f4"), ("y_vel_des", "<f4"), ("yaw_turn_rate", "<f4"), ("yaw_des", "<f4"), ("body_height", "<f4"),
    ("rpy_comp", "<f4", 2), ("world_position_desired", "<f4", 2), ("roll_des", "<f4"), ("pitch_des", "<f4"),
    ("stand_traj", "<f4", 3), ("x_comp_integral", "<f4"), ("cmpc_x_drag", "<f4"),
    ("gait_kind", "<i4"), ("gait_iteration", "<i4"), ("gait_offsets", "<i4", 4), ("gait_durations", "<i4", 4),
    ("gait_duty", "<f4"), ("omni_mode", "<i4"), ("stand", "<i4"), ("have_log", "<i4"),
    ("log_x_prev", "<f4", 12), ("log_R", "<f4", 9), ("log_r_feet", "<f4", 12), ("log_foot_force", "<f4", 12),
    ("log_x_drag", "<f4"), ("sim_time", "<f4"), ("pad_", "<f4"),
])
RESULT_DTYPE = np.dtype([
    ("fr_des", "<f4", 12), ("f_ff", "<f4", 12), ("world_position_desired", "<f4", 2), ("x_comp_integral", "<f4"),
    ("f_ext", "<f4", 6), ("status", "<i4"), ("iterations", "<i4"), ("pad_", "<f4", 1),
])
assert COMMAND_DTYPE.itemsize == 464 and RESULT_DTYPE.itemsize == 144

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libcmpc_b200.so is not built: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.cmpc_last_error.restype = C.c_char_p
        L.cmpc_batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
        L.cmpc_batch_destroy.argtypes = [C.c_void_p]
        L.cmpc_batch_setup.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double]
        L.cmpc_batch_set_robot.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.cmpc_batch_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.cmpc_batch_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(Inputs)]
        L.cmpc_batch_solve.argtypes = [C.c_void_p]
        L.cmpc_batch_sync.argtypes = [C.c_void_p]
        L.cmpc_batch_download.argtypes = [C.c_void_p, C.POINTER(Outputs)]
        L.cmpc_batch_solve_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(Inputs), C.POINTER(Outputs)]
        L.cmpc_batch_bind_host.argtypes = [C.c_void_p, C.POINTER(Inputs), C.POINTER(Outputs)]
        L.cmpc_batch_solve_bound.argtypes = [C.c_void_p, C.c_int]
        L.cmpc_batch_submit_bound.argtypes = [C.c_void_p, C.c_int]
        L.cmpc_batch_wait_bound.argtypes = [C.c_void_p]
        L.cmpc_batch_upload_disturbance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.cmpc_batch_download_disturbance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.cmpc_batch_set_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cmpc_batch_last_solve_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.cmpc_batch_kernel_launches.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.cmpc_batch_last_flops.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.cmpc_batch_solve_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cmpc_batch_mark.argtypes = [C.c_void_p, C.c_int]
        L.cmpc_batch_marked_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.cmpc_batch_reset_counters.argtypes = [C.c_void_p]
        L.cmpc_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.cmpc_host_unregister.argtypes = [C.c_void_p]
        L.cmpc_batch_kernel_flops.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.cmpc_batch_profile_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.cmpc_batch_enable_phase_clocks.argtypes = [C.c_void_p, C.c_int]
        L.cmpc_batch_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
        L.cmpc_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.cmpc_measure_dmma_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.cmpc_batch_set_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.cmpc_batch_solve_commands.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cmpc_batch_reset_history.argtypes = [C.c_void_p]
        L.cmpc_batch_history_length.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.cmpc_batch_copy_records.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cmpc_batch_device_records.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.cmpc_batch_device_forces.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.setup_problem.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double]
        L.update_problem_data_floats.argtypes = [C.c_void_p] * 5 + [C.c_float] * 3 + [C.c_void_p] * 2 + [C.c_float,
                                                                                                         C.c_void_p]
        L.update_problem_data.argtypes = [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.get_solution.argtypes = [C.c_int]
        L.get_solution.restype = C.c_double
        L.update_x_drag.argtypes = [C.c_float]
        L.update_solver_settings.argtypes = [C.c_int] + [C.c_double] * 5
        L.cmpc_set_external_force.argtypes = [C.c_void_p]
        L.cmpc_set_simulation_time.argtypes = [C.c_float]
        L.cmpc_get_disturbance_estimate.argtypes = [C.c_void_p]
        L.cmpc_get_disturbance_estimate_smoothed.argtypes = [C.c_void_p]
        L.cmpc_get_disturbance_estimate_static.argtypes = [C.c_void_p]
        L.cmpc_set_error_policy.argtypes = [C.c_int]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().cmpc_last_error().decode()))


def measure_fp64_peak(device=0):
    t = C.c_double()
    _check(lib().cmpc_measure_fp64_peak(device, C.byref(t)), "cmpc_measure_fp64_peak")
    return t.value


def measure_dmma_peak(device=0):
    t = C.c_double()
    _check(lib().cmpc_measure_dmma_peak(device, C.byref(t)), "cmpc_measure_dmma_peak")
    return t.value


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Batch:
    """Batched engine handle; mirrors setup_problem / update_problem_data / get_solution for many instances."""

    def __init__(self, capacity, device=0, options=None):
        self._h = C.c_void_p()
        self.capacity = capacity
        _check(lib().cmpc_batch_create(C.byref(self._h), device, capacity), "cmpc_batch_create")
        self.horizon = 0
        self.count = 0
        self._keep = None
        for env, (key, conv) in _ENV_OPTIONS.items():
            if env in os.environ:
                self.set_option(key, conv(os.environ[env]))
        for key, value in (options or {}).items():
            self.set_option(key, value)

    def set_option(self, key, value):
        _check(lib().cmpc_batch_set_option(self._h, key.encode(), int(value)), "cmpc_batch_set_option(%s)" % key)

    def close(self):
        if self._h:
            self.release_prepared()
            lib().cmpc_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self, dt, horizon, mu, f_max):
        _check(lib().cmpc_batch_setup(self._h, dt, horizon, mu, f_max), "cmpc_batch_setup")
        self.horizon = horizon

    def set_robot(self, mass, inertia):
        arr = (C.c_double * 3)(*inertia)
        _check(lib().cmpc_batch_set_robot(self._h, mass, arr), "cmpc_batch_set_robot")

    def _inputs(self, inst, count, f_dist=None):
        f32 = lambda k: np.ascontiguousarray(inst[k][:count], dtype=np.float32)
        arrs = {k: f32(k) for k in ("p", "v", "q", "w", "r", "weights", "traj", "alpha", "x_drag")}
        arrs["gait"] = np.ascontiguousarray(inst["gait"][:count], dtype=np.uint8)
        assert arrs["traj"].shape[1] == 12 * self.horizon and arrs["gait"].shape[1] == 4 * self.horizon
        if f_dist is not None:
            arrs["f_dist"] = np.ascontiguousarray(f_dist[:count], dtype=np.float32)
        s = Inputs()
        for k, a in arrs.items():
            setattr(s, k, a.ctypes.data)
        if f_dist is None:
            s.f_dist = None
        self._keep = arrs
        return s

    def upload(self, inst, count=None, f_dist=None):
        count = len(inst["p"]) if count is None else count
        s = self._inputs(inst, count, f_dist)
        _check(lib().cmpc_batch_upload(self._h, count, C.byref(s)), "cmpc_batch_upload")
        self.count = count

    def solve(self):
        _check(lib().cmpc_batch_solve(self._h), "cmpc_batch_solve")

    def solve_range(self, first, count):
        _check(lib().cmpc_batch_solve_range(self._h, first, count), "cmpc_batch_solve_range")

    def mark(self, which):
        _check(lib().cmpc_batch_mark(self._h, which), "cmpc_batch_mark")

    def marked_ms(self):
        ms = C.c_float()
        _check(lib().cmpc_batch_marked_ms(self._h, C.byref(ms)), "cmpc_batch_marked_ms")
        return ms.value

    def reset_counters(self):
        _check(lib().cmpc_batch_reset_counters(self._h), "cmpc_batch_reset_counters")

    def sync(self):
        _check(lib().cmpc_batch_sync(self._h), "cmpc_batch_sync")

    def _outputs(self, count, want_active=True):
        h = self.horizon
        res = {"forces": np.zeros((count, 12 * h)), "objective": np.zeros(count),
               "status": np.zeros(count, dtype=np.int32), "iterations": np.zeros(count, dtype=np.int32)}
        if want_active:
            res["active"] = np.zeros((count, 20 * h), dtype=np.int8)
        o = Outputs()
        for k, a in res.items():
            setattr(o, k, a.ctypes.data)
        return o, res

    def download(self, want_active=True):
        o, res = self._outputs(self.count, want_active)
        _check(lib().cmpc_batch_download(self._h, C.byref(o)), "cmpc_batch_download")
        return res

    def solve_host(self, inst, count=None, f_dist=None, want_active=True):
        count = len(inst["p"]) if count is None else count
        s = self._inputs(inst, count, f_dist)
        o, res = self._outputs(count, want_active)
        _check(lib().cmpc_batch_solve_host(self._h, count, C.byref(s), C.byref(o)), "cmpc_batch_solve_host")
        self.count = count
        return res

    def prepare_host(self, inst, count=None, f_dist=None, want_active=True, pin_outputs=True, pin_inputs=True):
        """Bind host input arrays and preallocated output arrays once; solve_prepared() then is one C call.
        pin_outputs registers the output arrays with CUDA so results are copied straight into them;
        pin_inputs registers the input arrays so the device reads them itself and packs the records (cmpc_pack.cu)."""
        count = len(inst["p"]) if count is None else count
        s = self._inputs(inst, count, f_dist)
        o, res = self._outputs(count, want_active)
        self.release_prepared()
        pinned = []
        if pin_inputs:
            for a in self._keep.values():
                if a.nbytes and lib().cmpc_host_register(a.ctypes.data, a.nbytes) == 0:
                    pinned.append(a)
        if pin_outputs:
            for a in res.values():
                if a.nbytes and lib().cmpc_host_register(a.ctypes.data, a.nbytes) == 0:
                    pinned.append(a)
        self._prepared = (count, s, o, res)
        self._prepared_inputs = self._keep   # later uploads rebind self._keep; these arrays stay alive with the binding
        _check(lib().cmpc_batch_bind_host(self._h, C.byref(s), C.byref(o)), "cmpc_batch_bind_host")
        self._pinned = pinned
        return res

    def release_prepared(self):
        if getattr(self, "_prepared", None) is not None and self._h:
            lib().cmpc_batch_bind_host(self._h, None, None)
            self._prepared = None
        for a in getattr(self, "_pinned", []):
            lib().cmpc_host_unregister(a.ctypes.data)
        self._pinned = []

    def solve_prepared(self):
        count, s, o, res = self._prepared
        _check(lib().cmpc_batch_solve_bound(self._h, count), "cmpc_batch_solve_bound")
        self.count = count
        return res

    def submit_prepared(self):
        """Enqueue the bound solve and return; wait_prepared() delivers the results."""
        _check(lib().cmpc_batch_submit_bound(self._h, self._prepared[0]), "cmpc_batch_submit_bound")

    def wait_prepared(self):
        count, s, o, res = self._prepared
        _check(lib().cmpc_batch_wait_bound(self._h), "cmpc_batch_wait_bound")
        self.count = count
        return res

    def last_solve_ms(self):
        ms = C.c_float()
        _check(lib().cmpc_batch_last_solve_ms(self._h, C.byref(ms)), "cmpc_batch_last_solve_ms")
        return ms.value

    def launches(self):
        n = C.c_longlong()
        _check(lib().cmpc_batch_kernel_launches(self._h, C.byref(n)), "cmpc_batch_kernel_launches")
        return n.value

    def last_flops(self):
        f = C.c_double()
        _check(lib().cmpc_batch_last_flops(self._h, C.byref(f)), "cmpc_batch_last_flops")
        return f.value

    KERNELS = ("assemble", "invert", "dual", "fused")

    def kernel_flops(self):
        arr = (C.c_double * 4)()
        _check(lib().cmpc_batch_kernel_flops(self._h, arr), "cmpc_batch_kernel_flops")
        return dict(zip(self.KERNELS, [float(x) for x in arr]))

    def profile_range(self, first, count):
        arr = (C.c_float * 4)()
        _check(lib().cmpc_batch_profile_range(self._h, first, count, arr), "cmpc_batch_profile_range")
        return dict(zip(self.KERNELS, [float(x) for x in arr]))

    PHASES = ("wait", "adapt", "prep", "hess", "load", "sweep", "store", "qp", "out", "publish", "dvwait", "x0_tilewait", "x1_tileload", "x2_kstore", "x3_looptop")

    def enable_phase_clocks(self, on=True):
        _check(lib().cmpc_batch_enable_phase_clocks(self._h, int(on)), "cmpc_batch_enable_phase_clocks")

    def phase_cycles(self):
        arr = (C.c_ulonglong * len(self.PHASES))()
        _check(lib().cmpc_batch_phase_cycles(self._h, arr, len(self.PHASES)), "cmpc_batch_phase_cycles")
        return dict(zip(self.PHASES, [int(x) for x in arr]))

    # ---- the caller of the path on the device (updateMPCIfNeeded / solveDenseMPC / getMpcTable) ----
    def set_weights(self, weights, alpha):
        w = np.ascontiguousarray(weights, dtype=np.float32)
        _check(lib().cmpc_batch_set_weights(self._h, _ptr(w), float(alpha)), "cmpc_batch_set_weights")

    def solve_commands(self, commands, want_forces=False, results=None):
        """One MPC update of len(commands) robots from cmpc_command structs (a COMMAND_DTYPE array)."""
        cmds = np.ascontiguousarray(commands, dtype=COMMAND_DTYPE)
        n = len(cmds)
        res = np.zeros(n, dtype=RESULT_DTYPE) if results is None else results
        forces = np.zeros((n, 12 * self.horizon)) if want_forces else None
        _check(lib().cmpc_batch_solve_commands(self._h, n, _ptr(cmds), _ptr(res), None if forces is None else _ptr(forces)),
               "cmpc_batch_solve_commands")
        self.count = n
        return (res, forces) if want_forces else res

    def reset_history(self):
        _check(lib().cmpc_batch_reset_history(self._h), "cmpc_batch_reset_history")

    def history_length(self):
        n = C.c_int()
        _check(lib().cmpc_batch_history_length(self._h, C.byref(n)), "cmpc_batch_history_length")
        return n.value

    def copy_records(self, first, count):
        h = self.horizon
        stride = (4 * (48 + 12 * h) + 4 * h + 15) & ~15
        out = np.zeros((count, stride), dtype=np.uint8)
        _check(lib().cmpc_batch_copy_records(self._h, first, count, _ptr(out)), "cmpc_batch_copy_records")
        return out

    def upload_disturbance(self, win_t, win_d, sim_time, mode):
        """mode 0 estimate, 1 estimate+apply, 2 apply the stored estimate (windows may be None), <0 off."""
        if mode < 0:
            _check(lib().cmpc_batch_upload_disturbance(self._h, 0, None, None, None, -1), "upload_disturbance")
            return
        stt = np.ascontiguousarray(sim_time, dtype=np.float32)
        if win_t is None:
            _check(lib().cmpc_batch_upload_disturbance(self._h, len(stt), None, None, _ptr(stt), mode),
                   "cmpc_batch_upload_disturbance")
            return
        wt = np.ascontiguousarray(win_t, dtype=np.float32)
        wd = np.ascontiguousarray(win_d, dtype=np.float32)
        _check(lib().cmpc_batch_upload_disturbance(self._h, len(wt), _ptr(wt), _ptr(wd), _ptr(stt), mode),
               "cmpc_batch_upload_disturbance")

    def download_disturbance(self):
        est = np.zeros((self.count, 4))
        fest = np.zeros((self.count, 6), dtype=np.float32)
        _check(lib().cmpc_batch_download_disturbance(self._h, _ptr(est), _ptr(fest)), "download_disturbance")
        return est, fest
