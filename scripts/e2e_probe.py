"""End-to-end host-buffer call (cmpc_batch_solve_host) at batch 4096: device-side vs host-side record packing,
chunk count; checks that both packing paths return identical results."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
B = int(os.environ.get("B", 4096))
inst = synth.make_batch(B, horizon=10, seed=1000)
ref = None
KNOBS = ("CMPC_CHUNKS", "CMPC_HOST_PACK", "CMPC_INV_REGS", "CMPC_D2H_COPY")
for env in ({"CMPC_HOST_PACK": "1", "CMPC_D2H_COPY": "1", "CMPC_CHUNKS": "2"}, {"CMPC_D2H_COPY": "1", "CMPC_CHUNKS": "2"},
            {"CMPC_CHUNKS": "1"}, {"CMPC_CHUNKS": "2"}, {"CMPC_CHUNKS": "3"}, {"CMPC_CHUNKS": "4"}, {"CMPC_CHUNKS": "2", "PIN_OUT": "0"},
            {"CMPC_CHUNKS": "1", "ACTIVE": "0"}, {"CMPC_CHUNKS": "2", "ACTIVE": "0"}, {"CMPC_CHUNKS": "3", "ACTIVE": "0"}):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: v for k, v in env.items() if k in KNOBS})
    b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0)
    b.prepare_host(inst, want_active=env.get("ACTIVE", "1") == "1", pin_outputs=env.get("PIN_OUT", "1") == "1")
    for _ in range(5):
        b.solve_prepared()
    t0 = time.perf_counter()
    n = 100
    for _ in range(n):
        res = b.solve_prepared()
    dt = (time.perf_counter() - t0) / n
    if ref is None:
        ref = {k: v.copy() for k, v in res.items()}
    same = all(np.array_equal(ref[k], res[k]) for k in res)
    print("%-50s %.3f ms/call  %.2f M solves/s  identical to host packing: %s" % (env, dt * 1e3, B / dt / 1e6, same), flush=True)
    b.close()
