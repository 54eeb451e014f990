"""Steady-state throughput (bench.py's timed loop) under environment knobs given as KEY=VALUE arguments.
   python scripts/knob_sweep.py [B=4096] [RING=8] [STEPS=48] CMPC_INV_REGS=168 CMPC_QCAP1=24 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
kv = dict(a.split("=", 1) for a in sys.argv[1:])
B, ring, steps = int(kv.pop("B", 4096)), int(kv.pop("RING", 8)), int(kv.pop("STEPS", 48))
h, gaits = int(kv.pop("H", 10)), tuple(kv.pop("GAITS", "trot").split(","))
os.environ.update(kv)
import numpy as np
from cmpc_b200 import synth, engine

inst = synth.make_batch(B * ring, horizon=h, seed=1000, gaits=gaits)
b = engine.Batch(B * ring); b.setup(0.03, h, 0.4, 120.0); b.upload(inst); b.sync()
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
t = {k: 0.0 for k in engine.Batch.KERNELS}
for i in range(4):
    tt = b.profile_range((i % ring) * B, B)
    for k in t: t[k] += tt[k] / 4
print("%-40s B=%d h=%d: %.4f ms/step %.2f M/s ok=%s | serial us: %s" % (
    " ".join("%s=%s" % x for x in kv.items()) or "(default)", B, h, best / steps, B * steps / best / 1e3,
    bool((res["status"] == 0).all()), " ".join("%s=%.1f" % (k, 1e3 * v) for k, v in t.items() if v > 0)), flush=True)
b.close()
