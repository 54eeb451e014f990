"""Host enqueue time vs wait time of cmpc_batch_solve_host (CMPC_TRACE=1), per chunk count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
os.environ["CMPC_TRACE"] = "1"
import numpy as np
from cmpc_b200 import synth, engine
B = int(os.environ.get("B", 4096))
inst = synth.make_batch(B, horizon=10, seed=1000)
for ch in (1, 2, 3):
    os.environ["CMPC_CHUNKS"] = str(ch)
    b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0)
    b.prepare_host(inst, want_active=False)
    t0 = time.perf_counter()
    for _ in range(100):
        b.solve_prepared()
    dt = (time.perf_counter() - t0) / 100
    print("chunks %d: %.3f ms/call, device time of the last call %.3f ms" % (ch, dt * 1e3, b.last_solve_ms()), flush=True)
    b.close()
