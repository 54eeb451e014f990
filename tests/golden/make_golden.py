"""Generate tests/golden/cmpc_golden.npz.

Run in the build container (where /root/reference exists): the outputs come
from the REAL reference solver — qpOASES 3.2.0 compiled from the reference's
vendored sources by oracle/Makefile, driven through the call sequence of
SolverMPC.cpp:955-964 — on H, g condensed by the fp64 restatement in
oracle/cmpc_oracle.cpp.  Inputs are seeded synthetic A1 instances
(cmpc_b200/synth.py).  Re-running this script reproduces the file bit for bit.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200"))
sys.path.insert(0, ROOT)
from cmpc_b200 import synth  # noqa: E402
from oracle import cmpc_oracle as O  # noqa: E402

IN_KEYS = ("p", "v", "q", "w", "r", "rpy", "weights", "traj", "alpha", "gait", "x_drag")

CASES = {
    # name: (horizon, gaits, spread, count, seed, n_segment)
    "trot10": (10, ("trot",), 1.0, 24, 11, None),       # BASELINE configs[0]/[1] shape
    "trot10hard": (10, ("trot",), 3.0, 16, 12, None),   # many active cone faces
    "mixed16": (16, ("trot", "bound", "pace", "gallop"), 1.5, 12, 13, 10),  # configs[3] shape
    "stand10": (10, ("stand",), 2.0, 6, 14, None),      # all feet down, n = 120
    "pronk10": (10, ("pronk",), 1.0, 6, 15, None),      # flight phases: some steps have no contact
}


def main():
    out = {}
    for name, (h, gaits, spread, count, seed, nseg) in CASES.items():
        inst = synth.make_batch(count, horizon=h, seed=seed, gaits=gaits, spread=spread, n_segment=nseg)
        st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
        forces = np.zeros((count, 12 * h))
        obj = np.zeros(count)
        nwsr = np.zeros(count, dtype=np.int32)
        ws = np.zeros((count, 20 * h), dtype=np.int8)
        ok = np.zeros(count, dtype=np.int8)
        g_full = np.zeros((count, 12 * h))
        h_diag = np.zeros((count, 12 * h))
        for i in range(count):
            r = O.solve(st, O.make_update(inst, i, h), want_mats=True)
            forces[i], obj[i], nwsr[i], ws[i], ok[i] = r["x"], r["objective"], r["nwsr"], r["con_status"], r["ok"]
            g_full[i] = r["g_full"]
            h_diag[i] = np.diag(r["H_full"])
        assert ok.all(), name
        for k in IN_KEYS:
            out["%s_in_%s" % (name, k)] = inst[k]
        out[name + "_forces"] = forces
        out[name + "_objective"] = obj
        out[name + "_nwsr"] = nwsr
        out[name + "_workingset"] = ws
        out[name + "_g"] = g_full
        out[name + "_Hdiag"] = h_diag
        out[name + "_meta"] = np.array([h, inst["dt"], inst["mu"], inst["f_max"]], dtype=np.float64)
        print(name, "h=%d count=%d nwsr max=%d" % (h, count, nwsr.max()))
    # disturbance estimator windows (SolverMPC.cpp:704-753)
    t, d, _ = synth.make_disturbance_windows(8, seed=21)
    est = np.stack([O.fit_window(t[i], d[i]) for i in range(len(t))])
    out["dist_t"], out["dist_d"], out["dist_est"] = t, d, est
    path = os.path.join(ROOT, "tests", "golden", "cmpc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
