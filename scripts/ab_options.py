"""A/B of engine options on resident batches: serial solve time, pipelined throughput, bit-agreement of forces.
usage: ab_options.py workload(trot|mixed) B steps "k=v,k=v" ...   (an empty string = defaults)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine

wl, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
variants = sys.argv[4:] or [""]
ring = 4
if wl == "mixed":
    h = 16
    inst = synth.make_batch(B * ring, horizon=h, seed=1000, gaits=("trot", "bound", "pace", "gallop"), n_segment=10, spread=1.5)
else:
    h = 10
    inst = synth.make_batch(B * ring, horizon=h, seed=1000)
ref = None
for v in variants:
    opts = dict((k, int(x)) for k, x in (kv.split("=") for kv in v.split(",") if kv))
    b = engine.Batch(B * ring, options=opts)
    b.setup(0.03, h, 0.4, 120.0)
    b.upload(inst)
    for i in range(3):
        b.solve_range((i % ring) * B, B)
    b.sync()
    ser = []
    for i in range(4):
        b.solve_range((i % ring) * B, B); b.sync(); ser.append(b.last_solve_ms())
    best = 1e9
    for rep in range(3):
        b.mark(0)
        for i in range(steps):
            b.solve_range((i % ring) * B, B)
        b.mark(1); b.sync()
        best = min(best, b.marked_ms() / steps)
    t = b.profile_range(0, B)
    res = b.download()
    if ref is None:
        ref = res
    dev = np.abs(res["forces"] - ref["forces"]).max()
    print("%-28s serial %.3f ms  pipelined %.3f ms/step (%.2f M/s)  classes asm %.3f inv %.3f dual %.3f  solved %s  max|dF| vs first %.1e  iters max %d"
          % (v or "(defaults)", min(ser), best, B / best / 1e3, t["assemble"], t["invert"], t["dual"], bool((res["status"] == 0).all()), dev, res["iterations"].max()), flush=True)
    b.close()
