"""One warm + a few measured launches of the solve kernel on the bench workload (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
from cmpc_b200 import synth, engine
B = int(os.environ.get("PROF_BATCH", "4096"))
inst = synth.make_batch(B, horizon=10, seed=1000)
b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0); b.upload(inst)
for _ in range(int(os.environ.get("PROF_LAUNCHES", "4"))):
    b.solve()
b.sync()
print("kernel ms", b.last_solve_ms())
