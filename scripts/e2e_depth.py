"""End-to-end throughput of the bound call against the number of batches in flight
(cmpc_batch_submit_bound / cmpc_batch_wait_bound on D engine handles; every step reads its inputs from and writes
its results to pinned host arrays).  Usage: python scripts/e2e_depth.py [steps] [depths...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
from cmpc_b200 import engine, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
depths = [int(a) for a in sys.argv[2:]] or [1, 2, 3, 4, 6, 8]
BATCH, h, DT = 4096, 10, 0.03
MODE = os.environ.get("MODE", "e2e")  # "resident": the same loop over handles with device-resident inputs and outputs (no PCIe)
inst = synth.make_batch(BATCH * max(depths), horizon=h, seed=77)
for D in depths:
    pipe = []
    for k in range(D):
        sk = {kk: (v[k * BATCH:(k + 1) * BATCH] if isinstance(v, np.ndarray) else v) for kk, v in inst.items()}
        bk = engine.Batch(BATCH)
        bk.setup(DT, h, inst["mu"], inst["f_max"])
        if MODE == "resident":
            bk.upload(sk)
            bk.submit_prepared = bk.solve
            bk.wait_prepared = bk.sync
            bk.solve_prepared = lambda bk=bk: (bk.solve(), bk.sync())
        else:
            if os.environ.get("PINNED") == "hostalloc":  # inputs in cudaHostAlloc memory instead of cudaHostRegister-ed numpy arrays
                import torch
                keep = []
                for kk, v in list(sk.items()):
                    if isinstance(v, np.ndarray):
                        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.uint8 if kk == "gait" else np.float32)).pin_memory()
                        keep.append(t)
                        sk[kk] = t.numpy()
                bk._hostalloc_keep = keep
            bk.prepare_host(sk, want_active=False)
        pipe.append(bk)
    for k in range(3 * D):
        pipe[k % D].solve_prepared()
    t0 = time.perf_counter()
    t_sub = t_wait = 0.0
    for k in range(steps):
        ta = time.perf_counter()
        if k >= D:
            rp = pipe[k % D].wait_prepared()
        tb = time.perf_counter()
        pipe[k % D].submit_prepared()
        t_sub += time.perf_counter() - tb
        t_wait += tb - ta
    for k in range(steps, steps + D):
        rp = pipe[k % D].wait_prepared()
    wall = time.perf_counter() - t0
    if MODE != "resident":
        assert (rp["status"] == 0).all()
    print("depth %d: %.3f ms/step  %.2f M solves/s   host: submit %.1f us, wait %.1f us per step"
          % (D, 1e3 * wall / steps, steps * BATCH / wall / 1e6, 1e6 * t_sub / steps, 1e6 * t_wait / steps), flush=True)
    for bk in pipe:
        bk.close()
