"""A/B of engine options on the synchronous end-to-end call (cmpc_batch_solve_bound, 4096 trot, pinned host arrays) and on
two batches in flight.  usage: e2e_ab.py [steps] "k=v,k=v" "k=v" ...   (an empty string = defaults)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
variants = sys.argv[2:] or [""]
B, h = 4096, 10
inst = synth.make_batch(2 * B, horizon=h, seed=77)
ref = None
for v in variants:
    opts = dict((k, int(x)) for k, x in (kv.split("=") for kv in v.split(",") if kv))
    pipe = []
    for k in range(2):
        sk = {kk: (a[k * B:(k + 1) * B] if isinstance(a, np.ndarray) else a) for kk, a in inst.items()}
        bk = engine.Batch(B, options=opts)
        bk.setup(0.03, h, inst["mu"], inst["f_max"])
        bk.prepare_host(sk, want_active=True)
        pipe.append(bk)
    b = pipe[0]
    for _ in range(5):
        res = b.solve_prepared()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(steps):
            res = b.solve_prepared()
        best = min(best, (time.perf_counter() - t0) / steps)
    assert (res["status"] == 0).all()
    if ref is None:
        ref = {k: a.copy() for k, a in res.items()}
    same = all((res[k] == ref[k]).all() for k in ref)
    for k in range(4):
        pipe[k % 2].solve_prepared()
    t0 = time.perf_counter()
    for k in range(steps):
        if k >= 2:
            pipe[k % 2].wait_prepared()
        pipe[k % 2].submit_prepared()
    for k in range(steps, steps + 2):
        pipe[k % 2].wait_prepared()
    two = (time.perf_counter() - t0) / steps
    print("%-32s sync %.4f ms/step (%.2f M/s)   two in flight %.4f ms/step (%.2f M/s)   same bits as first variant: %s"
          % (v or "(defaults)", 1e3 * best, B / best / 1e6, 1e3 * two, B / two / 1e6, same), flush=True)
    for bk in pipe:
        bk.close()
