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
# per-phase cycles of the adaptive assembly kernel (thread-0 clocks summed over CTAs)
b.enable_phase_clocks(True)
for _ in range(4):
    b.solve(); b.sync()
cyc = b.phase_cycles()
tot = sum(cyc.values())
for k, v in cyc.items():
    if v:
        print("   %-12s %5.1f%%  %8.0f cyc/instance" % (k, 100.0 * v / tot, v / (B * 4)))
print("class times (ms):", b.profile_range(0, B))
