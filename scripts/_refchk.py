import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/quad-periodic-mpc_b200")
from cmpc_b200 import engine, synth
h, B = 10, 512
inst = synth.make_batch(B, horizon=h, seed=654, gaits=("trot", "pace", "walk2"), spread=1.5)
inst["alpha"] = np.full(B, 1e-6, np.float32)
out = {}
for v in (-1, 0, 512):
    b = engine.Batch(B, options={"inv_refine": v}); b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    out[v] = b.solve_host(inst)["forces"].copy(); b.close()
print("bitwise equal -1 vs 0:", np.array_equal(out[-1], out[0]), "max diff", np.abs(out[-1] - out[0]).max())
print("bitwise equal 0 vs 512:", np.array_equal(out[512], out[0]), "max diff", np.abs(out[512] - out[0]).max())
