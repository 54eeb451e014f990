"""How accurate is each path of the engine on badly conditioned instances (alpha = 1e-6, some state weights zero)?
Prints, per path, the largest force difference against qpOASES (working-set cap lifted) and the largest slack the
engine's own forces leave on rows that are active at qpOASES' solution (exact arithmetic leaves none)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200"))
from cmpc_b200 import engine, synth
from oracle import cmpc_numpy as N, cmpc_oracle as O

h, B = 10, 512
alphas = [float(a) for a in os.environ.get("ALPHAS", "1e-6,1e-5,4e-5,1e-3").split(",")]


def slacks(f, mu_inv):             # (contacts, 5) slacks of the reference's rows for forces f (contacts, 3)
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    return np.stack([fx * mu_inv + fz, -fx * mu_inv + fz, fy * mu_inv + fz, -fy * mu_inv + fz, fz], 1)


def study(title, gaits, variants):
    inst = synth.make_batch(B, horizon=h, seed=654, gaits=gaits, spread=1.5)
    rng = np.random.default_rng(11)
    w = synth.A1_WEIGHTS[None, :] * rng.uniform(0.2, 5.0, (B, 12)).astype(np.float32)
    w[rng.random((B, 12)) < 0.1] = 0.0
    w[:, 5] = np.maximum(w[:, 5], 1.0)
    inst["weights"] = w.astype(np.float32)
    inst["alpha"] = rng.choice(np.array(alphas, np.float32), B)
    inst["x_drag"] = (rng.choice(np.array([0.0, 0.3, 3.0, -2.0], np.float32), B)).astype(np.float32)
    st = O.make_setup(inst["dt"], h, inst["mu"], inst["f_max"], nwsr=5000)
    ref = np.zeros((B, 12 * h)); ok = np.zeros(B, bool)
    for i in range(B):
        r = O.solve(st, O.make_update(inst, i, h)); ref[i] = r["x"]; ok[i] = r["ok"]
    mu_inv = float(np.float32(1.0) / np.float32(inst["mu"]))
    nvar = [len(N.contact_vars(inst["gait"][i], h)) for i in range(B)]
    print("#### %s: gaits %s, reduced variables %d..%d" % (title, gaits, min(nvar), max(nvar)))
    for name, opts in variants:
        b = engine.Batch(B, options=opts); b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
        res = b.solve_host(inst); b.close()
        print("== %s: statuses %s" % (name, dict(zip(*[x.tolist() for x in np.unique(res["status"], return_counts=True)]))))
        for a in alphas:
            sel = ok & (np.abs(inst["alpha"] - np.float32(a)) < 1e-12) & (res["status"] == engine.ST_SOLVED)
            dF, sl = 0.0, 0.0
            for i in np.flatnonzero(sel):
                keep = N.contact_vars(inst["gait"][i], h)
                sr = slacks(ref[i][keep].reshape(-1, 3), mu_inv); sg = slacks(res["forces"][i][keep].reshape(-1, 3), mu_inv)
                act = sr < 1e-9
                if act.any(): sl = max(sl, np.abs(sg[act]).max())
                dF = max(dF, np.abs(res["forces"][i] - ref[i]).max())
            print("   alpha %.0e: %3d instances, max |dF| %.2e N, max slack on rows active at the optimum %.2e" % (a, sel.sum(), dF, sl))


common = [("default (panel refinement above 1024)", {}), ("no panel refinement (inv_refine = -1)", {"inv_refine": -1}),
          ("always refine (inv_refine = 0)", {"inv_refine": 0}), ("inv_refine = 512", {"inv_refine": 512}),
          ("inv_refine = 4096", {"inv_refine": 4096}), ("fused kernel (scalar sweep)", {"path_fused": 1})]
which = os.environ.get("STUDY", "12")
if "1" in which:
    study("n <= 63 pipeline (tensor-core inversion kernel)", ("trot", "pace"), common)
if "2" in which:
  study("64..128-variable shapes (register-tile sweep in the condensation kernel)", ("trot", "pace", "walk2"),
      common + [("sweep_dmma", {"sweep_dmma": 1}), ("sweep_dmma, no refinement", {"sweep_dmma": 1, "inv_refine": -1}),
              ("sweep_dmma, always refine", {"sweep_dmma": 1, "inv_refine": 0})])
