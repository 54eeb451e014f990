"""Small runs of every kernel family (pipeline shapes, fused path, estimator, commands): a quick end-to-end regression."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
for tag, h, gaits, nseg, B, spread in (("trot10", 10, ("trot",), None, 300, 2.0), ("mixed16", 16, ("trot", "bound", "pace", "gallop"), 10, 48, 1.5),
                                       ("stand10", 10, ("stand",), None, 24, 1.0), ("trot19", 19, ("walk2",), None, 8, 1.0)):
    inst = synth.make_batch(B, horizon=h, seed=5, gaits=gaits, n_segment=nseg, spread=spread)
    b = engine.Batch(B); b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
    res = b.solve_host(inst)
    b.prepare_host(inst); res2 = b.solve_prepared()
    assert (res["forces"] == res2["forces"]).all()
    if h == 10 and gaits == ("trot",):
        t, d, _ = synth.make_disturbance_windows(B, seed=3)
        b.upload_disturbance(t, d, t[:, -1].copy(), 1)
        b.solve_host(inst)
        b.upload_disturbance(None, None, None, -1)
        c = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=h, seed=2, mixed_fraction=0.2, stand_fraction=0.1)
        for _ in range(2):
            b.solve_commands(c)
    print(tag, "ok, status", np.unique(res["status"]), "iters max", res["iterations"].max(), flush=True)
    b.close()
    # the capacity tiers behind a small first tier: CTA-per-instance tier, one-warp tiers, hand-over with and without P
    for opts in ({"qcap1": 8}, {"qcap1": 8, "dual_team": 0}, {"qcap1": 5, "resume_p": 0}, {"qcap1": 6, "resume": 0}, {"path_fused": 1}):
        b = engine.Batch(B, options=opts); b.setup(inst["dt"], h, inst["mu"], inst["f_max"])
        r = b.solve_host(inst)
        assert np.abs(r["forces"] - res["forces"]).max() <= 1e-7, (tag, opts)
        b.close()
    print(tag, "tiers ok", flush=True)
