"""The fuzz configurations of tests/tools/fuzz.py (seed 7771, 400 rounds) once more: every instance whose forces are more than
1e-6 N from the reference's qpOASES call (nWSR capped at 100) is solved again with the cap lifted.  Test infrastructure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
from oracle import cmpc_oracle as O
threads = len(os.sched_getaffinity(0))
rng = np.random.default_rng(7771)
names = list(synth.GAITS)
found = 0
for it in range(400):
    h = int(rng.integers(1, 20)); k = int(rng.integers(1, 4))
    gaits = tuple(rng.choice(names, size=k, replace=False))
    nseg = int(rng.choice([h, 10, 16])) if rng.random() < 0.5 else None
    spread = float(rng.choice([0.5, 1.0, 1.5, 2.5]))
    B = int(rng.choice([1, 2, 3, 17, 64, 257, 1024]))
    inst = synth.make_batch(B, horizon=h, seed=int(rng.integers(1 << 30)), gaits=gaits, n_segment=nseg, spread=spread)
    b = engine.Batch(B); b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    res = b.solve_host(inst); b.close()
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"])
    ups = (O.Update * B)(*[O.make_update(inst, i, h) for i in range(B)])
    ref, ok = O.solve_batch(st, ups, threads, use_float=False)
    err = np.abs(res["forces"] - ref).max(1)
    err[ok == 0] = 0
    for i in np.flatnonzero(err > 1e-6):
        st2 = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"], nwsr=5000)
        r2 = O.solve(st2, O.make_update(inst, int(i), h))
        e2 = np.abs(res["forces"][i] - r2["x"]).max()
        print("round %d h=%d gaits=%s B=%d inst %d: |dF| vs capped qpOASES %.2e, vs qpOASES with the cap lifted %.2e (nWSR used %s), GPU iterations %d"
              % (it, h, gaits, B, i, err[i], e2, r2.get("nwsr"), res["iterations"][i]), flush=True)
        found += 1
print("instances above 1e-6 N against the capped reference:", found)
