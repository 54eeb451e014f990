"""Latency of the end-to-end call for SMALL batches (pageable inputs, the way the reference's single-instance interface calls it):
the three-kernel pipeline against the fused single-kernel path.  usage: latency_small.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
for B in (1, 4, 16, 64, 148, 296, 592, 1184):
    inst = synth.make_batch(B, horizon=10, seed=5)
    row = []
    for opts in ({}, {"path_fused": 1}):
        b = engine.Batch(B, options=opts); b.setup(0.03, 10, 0.4, 120.0)
        for _ in range(20): res = b.solve_host(inst)
        lat = []
        for _ in range(300):
            t0 = time.perf_counter(); res = b.solve_host(inst); lat.append(1e6 * (time.perf_counter() - t0))
        lat.sort(); row.append((lat[len(lat) // 2], b.last_solve_ms() * 1e3))
        assert (res["status"] == 0).all()
        b.close()
    print("B=%5d  pipeline: median %.1f us (device %.1f us)   fused kernel: median %.1f us (device %.1f us)" % (B, row[0][0], row[0][1], row[1][0], row[1][1]), flush=True)
