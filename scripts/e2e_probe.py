"""End-to-end host-buffer call (cmpc_batch_solve_host) at batch 4096: chunk count and host threads."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
B = 4096
inst = synth.make_batch(B, horizon=10, seed=1000)
for env in ({"CMPC_CHUNKS": "1"}, {"CMPC_CHUNKS": "2"}, {"CMPC_CHUNKS": "3"}, {"CMPC_CHUNKS": "4"}, {"CMPC_CHUNKS": "2", "CMPC_HOST_THREADS": "1"},
            {"CMPC_CHUNKS": "2", "PIN": "0"}):
    for k in ("CMPC_CHUNKS", "CMPC_HOST_THREADS"):
        os.environ.pop(k, None)
    os.environ.update({k: v for k, v in env.items() if k != "PIN"})
    b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0)
    b.prepare_host(inst, want_active=False, pin_outputs=env.get("PIN", "1") == "1")
    for _ in range(5):
        b.solve_prepared()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        res = b.solve_prepared()
    dt = (time.perf_counter() - t0) / n
    print("%-50s %.3f ms/call  %.2f M solves/s  kernels %.3f ms" % (env, dt * 1e3, B / dt / 1e6, b.last_solve_ms()), flush=True)
    b.close()
