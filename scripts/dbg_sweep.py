import os, sys
sys.path.insert(0, "quad-periodic-mpc_b200"); sys.path.insert(0, ".")
import numpy as np
from cmpc_b200 import synth, engine
def run(env, inst):
    for k in ("CMPC_SWEEP",): os.environ.pop(k, None)
    os.environ.update(env)
    b = engine.Batch(len(inst["p"])); b.setup(inst["dt"], inst["horizon"], inst["mu"], inst["f_max"])
    r = b.solve_host(inst); b.close(); return r
for h, gaits, nseg in ((16, ("trot","bound","pace","gallop"), 10), (10, ("stand",), None), (12, ("trot",), None), (13, ("trot",), None)):
    inst = synth.make_batch(64, horizon=h, seed=9, gaits=gaits, n_segment=nseg)
    a = run({}, inst); b = run({"CMPC_SWEEP": "dmma"}, inst)
    n = 3 * inst["gait"].astype(bool).sum(1)
    d = np.abs(a["forces"] - b["forces"]).max(1)
    bad = d > 1e-6
    print("h", h, gaits[0], "n values", sorted(set(n.tolist())), "bad", bad.sum(), "of", len(n), "bad n:", sorted(set(n[bad].tolist())), "good n:", sorted(set(n[~bad].tolist())), "status", np.unique(b["status"]), "maxdiff", d.max())
