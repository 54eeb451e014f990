"""A few launches of the adaptive solve (estimator fit + apply fused) on 4096 instances, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
from cmpc_b200 import synth, engine
B = 4096
inst = synth.make_batch(B, horizon=10, seed=1000)
b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0); b.upload(inst)
t, d, _ = synth.make_disturbance_windows(B, seed=3)
b.upload_disturbance(t, d, t[:, -1].copy(), 1)
for _ in range(3):
    b.solve()
b.sync()
print("kernel ms", b.last_solve_ms())
