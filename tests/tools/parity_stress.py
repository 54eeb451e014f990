"""Large-sample parity of the CUDA path — the end-to-end call and the device-resident solve_range path — against the
reference's qpOASES (oracle/_ref, fp64 condensation, all host threads): forces, and the count of instances on which
qpOASES itself gave up (nWSR cap of 100)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
from oracle import cmpc_oracle as O

assert O.available(), "oracle/_ref did not travel"
threads = len(os.sched_getaffinity(0))
CASES = [("trot h=10", 10, ("trot",), None, 1.0, 8192), ("trot h=10, hard (spread 2.5)", 10, ("trot",), None, 2.5, 4096),
         ("all gaits h=10", 10, ("trot", "bound", "pronk", "pace", "gallop", "trotrun", "walk2"), None, 1.5, 4096),
         ("stand h=10 (n = 120)", 10, ("stand",), None, 1.0, 1024),
         ("mixed gaits h=16", 16, ("trot", "bound", "pace", "gallop"), 10, 1.5, 2048),
         ("trot h=6", 6, ("trot",), None, 1.0, 2048), ("trot h=19 (n_segment 10)", 19, ("trot",), 10, 1.0, 1024)]
for tag, h, gaits, nseg, spread, B in CASES:
    inst = synth.make_batch(B, horizon=h, seed=4242, gaits=gaits, n_segment=nseg, spread=spread)
    b = engine.Batch(B); b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    res = b.solve_host(inst)
    b.upload(inst)                         # the device-resident path (solve_range: first capacity tier of 24 rows)
    for k0 in range(0, B, 1024):
        b.solve_range(k0, min(1024, B - k0))
    res_r = b.download(); b.close()
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    ups = (O.Update * B)(*[O.make_update(inst, i, h) for i in range(B)])
    t0 = time.perf_counter()
    ref, ok = O.solve_batch(st, ups, threads, use_float=False)
    dt = time.perf_counter() - t0
    good = ok != 0
    err = np.abs(res["forces"] - ref)
    tol = 1e-3 + 1e-5 * np.abs(ref)
    within = (err <= tol).all(1) & (np.abs(res_r["forces"] - ref) <= tol).all(1)
    err = np.maximum(err, np.abs(res_r["forces"] - ref))
    print("%-30s %5d instances: qpOASES solved %5d; of those within 1e-3 N + 1e-5 rel: %5d, max |dF| %.2e N; GPU status solved %d; "
          "iterations mean %.1f max %d; qpOASES %.0f solves/s on %d threads"
          % (tag, B, good.sum(), (within & good).sum(), err[good].max() if good.any() else 0.0, (res["status"] == 0).sum(),
             res["iterations"].mean(), res["iterations"].max(), B / dt, threads), flush=True)
