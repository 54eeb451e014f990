"""End-to-end bound call: chunk count against batch size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
for B, chunks in ((4096, (1, 2)), (16384, (1, 2, 4)), (65536, (1, 2, 4, 8))):
    inst = synth.make_batch(B, horizon=10, seed=1000)
    for ch in chunks:
        os.environ["CMPC_CHUNKS"] = str(ch)
        b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0)
        b.prepare_host(inst, want_active=False)
        for _ in range(3):
            b.solve_prepared()
        n = max(5, 200000 // B)
        t0 = time.perf_counter()
        for _ in range(n):
            b.solve_prepared()
        dt = (time.perf_counter() - t0) / n
        print("B=%6d chunks %d: %.3f ms/call  %.2f M solves/s" % (B, ch, dt * 1e3, B / dt / 1e6), flush=True)
        b.close()
