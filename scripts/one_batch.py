"""A few serial solves of one resident batch (what the ncu captures are taken on).
usage: one_batch.py [B] [h] [reps] [workload: trot|mixed] [key=value options ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
h = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = sys.argv[4] if len(sys.argv) > 4 else "trot"
opts = dict((k, int(v)) for k, v in (a.split("=") for a in sys.argv[5:]))
if wl == "mixed":
    inst = synth.make_batch(B, horizon=h, seed=1000, gaits=("trot", "bound", "pace", "gallop"), n_segment=10, spread=1.5)
else:
    inst = synth.make_batch(B, horizon=h, seed=1000)
b = engine.Batch(B, options=opts)
b.setup(0.03, h, 0.4, 120.0)
b.upload(inst)
ms = []
for _ in range(reps):
    b.solve_range(0, B)     # throughput configuration (24-row first tier) ...
    b.sync()
    ms.append(b.last_solve_ms())
res = b.download()
print("B=%d h=%d %s %s: %s ms per solve; all solved %s; iterations mean %.1f max %d"
      % (B, h, wl, opts, ["%.3f" % m for m in ms], bool((res["status"] == 0).all()), res["iterations"].mean(), res["iterations"].max()))
b.close()
