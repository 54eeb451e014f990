"""Steady-state throughput of back-to-back solve_range calls over a ring of resident batches (what bench.py times)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine

def run(B, ring, steps, env):
    for k in ("CMPC_PATH", "CMPC_SERIAL", "CMPC_QCAP1"):
        os.environ.pop(k, None)
    os.environ.update(env)
    inst = synth.make_batch(B * ring, horizon=10, seed=1000)
    b = engine.Batch(B * ring); b.setup(0.03, 10, 0.4, 120.0); b.upload(inst); b.sync()
    for i in range(5):
        b.solve_range((i % ring) * B, B)
    b.sync()
    best = 1e9
    for rep in range(3):
        b.mark(0)
        for i in range(steps):
            b.solve_range((i % ring) * B, B)
        b.mark(1); b.sync()
        best = min(best, b.marked_ms())
    res = b.download()
    print("%-28s B=%d: %.3f ms/step  %.2f M solves/s  status ok=%s" % (env, B, best / steps, B * steps / best / 1e3, bool((res["status"] == 0).all())), flush=True)
    b.close()

if __name__ == "__main__":
    for env in ({"CMPC_PATH": "fused", "CMPC_SERIAL": "1"}, {"CMPC_SERIAL": "1"}, {}):
        run(4096, 8, 48, env)
    run(16384, 2, 16, {"CMPC_SERIAL": "1"})
    run(16384, 2, 16, {})
