"""A few solves of the mixed-gait h=16 workload (BASELINE configs[3]) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
from cmpc_b200 import synth, engine
B = 4096
inst = synth.make_batch(B, horizon=16, seed=1000, gaits=("trot", "bound", "pace", "gallop"), n_segment=10, spread=1.5)
b = engine.Batch(B); b.setup(0.03, 16, 0.4, 120.0); b.upload(inst)
for _ in range(3):
    b.solve(); b.sync()
print("kernel ms", b.last_solve_ms())
