"""Device-resident throughput of every BASELINE.json config on one GPU (the bench line is configs[1] only):
  [1] 4096 A1 trot h=10   [2] 65536 Adaptive-MPC instances, estimator fused (fit + apply in the launch)
  [3] mixed gaits (trot/bound/pace/gallop), h=16   [4] 1M instances, walked as resident 65536-instance batches
plus the device-side caller (commands) on [2]'s history rules.  Prints one line per config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine


def resident(tag, B, h, gaits, nseg=None, spread=1.0, steps=8, ring=2, adaptive=False, total=None):
    inst = synth.make_batch(B * ring, horizon=h, seed=1000, gaits=gaits, n_segment=nseg, spread=spread)
    b = engine.Batch(B * ring); b.setup(0.03, h, 0.4, 120.0); b.upload(inst)
    if adaptive:
        t, d, _ = synth.make_disturbance_windows(B * ring, seed=3)
        b.upload_disturbance(t, d, t[:, -1].copy(), 1)
    for i in range(3):
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
    ok = bool((res["status"] == 0).all())
    n = total or B * steps
    ms = best * (n / (B * steps))
    print("%-58s %9d instances  %8.3f ms  %6.2f M solves/s  iters mean %.1f max %d  all solved: %s"
          % (tag, n, ms, B * steps / best / 1e3, res["iterations"].mean(), res["iterations"].max(), ok), flush=True)
    b.close()


if __name__ == "__main__":
    resident("[1] A1 trot h=10, batch 4096", 4096, 10, ("trot",), steps=48, ring=8)
    resident("[2] Adaptive-MPC h=10, batch 65536, fit + apply fused", 65536, 10, ("trot",), steps=4, ring=1, adaptive=True)
    resident("[3] mixed gaits h=16 (n_segment 10), batch 4096", 4096, 16, ("trot", "bound", "pace", "gallop"), nseg=10, spread=1.5, steps=8, ring=2)
    resident("[4] 1M A1 trot h=10 as 16 resident batches of 65536", 65536, 10, ("trot",), steps=16, ring=2, total=1 << 20)
