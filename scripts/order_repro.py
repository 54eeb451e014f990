import os, sys
sys.path.insert(0, "quad-periodic-mpc_b200"); sys.path.insert(0, ".")
kv = dict(a.split("=", 1) for a in sys.argv[1:]); os.environ.update(kv)
import numpy as np
from cmpc_b200 import synth, engine
h, B, ring = 10, 256, 6
inst = synth.make_batch(B * ring, horizon=h, seed=701, spread=2.0)
b = engine.Batch(B * ring); b.setup(inst["dt"], h, inst["mu"], inst["f_max"]); b.upload(inst)
for rep in range(3):
    for i in range(ring):
        b.solve_range(i * B, B)
    full = b.download()
    b2 = engine.Batch(B * ring); b2.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    whole = b2.solve_host(inst); b2.close()
    bad = np.flatnonzero((full["forces"] != whole["forces"]).any(1))
    print(kv, "rep", rep, "mismatching instances:", len(bad), bad[:12], "status", np.unique(full["status"]), "iters of bad", full["iterations"][bad][:12], whole["iterations"][bad][:12],
          "maxdiff", np.abs(full["forces"] - whole["forces"]).max())
