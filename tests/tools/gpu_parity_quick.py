"""Quick GPU parity run: CUDA path vs the oracle (real qpOASES) on seeded batches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
from oracle import cmpc_oracle as O

def run(h, gaits, spread, B, seed, nseg=None, ncheck=64):
    inst = synth.make_batch(B, horizon=h, seed=seed, gaits=gaits, spread=spread, n_segment=nseg)
    b = engine.Batch(B)
    b.setup(0.03, h, 0.4, 120.0)
    t0 = time.time()
    res = b.solve_host(inst)
    t1 = time.time()
    res = b.solve_host(inst)
    t2 = time.time()
    b.upload(inst); b.solve(); b.solve(); b.sync()
    ms = b.last_solve_ms()
    st = O.make_setup(0.03, h, 0.4, 120.0)
    worst = 0; objrel = 0; itd = 0
    for i in range(min(ncheck, B)):
        r = O.solve(st, O.make_update(inst, i, h))
        if not r["ok"]: continue
        worst = max(worst, np.abs(res["forces"][i] - r["x"]).max())
        if r["n_var"]: objrel = max(objrel, abs(res["objective"][i] - r["objective"]) / abs(r["objective"]))
        itd = max(itd, abs(int(res["iterations"][i]) - r["nwsr"]))
    print(f"h={h} gaits={gaits} spread={spread} B={B}: status={np.bincount(res['status'])} "
          f"iters mean/max={res['iterations'].mean():.1f}/{res['iterations'].max()} worst_abs={worst:.2e} "
          f"objrel={objrel:.2e} iter-nwsr max diff={itd} kernel_ms={ms:.3f} ({B/ms*1e3:.0f} solves/s) "
          f"host_call_ms first={1e3*(t1-t0):.1f} second={1e3*(t2-t1):.1f} flops={b.last_flops():.3e}", flush=True)
    b.close()

if __name__ == "__main__":
    run(10, ("trot",), 1.0, 64, 1)
    run(10, ("trot",), 1.0, 4096, 1)
    run(10, ("trot",), 3.0, 4096, 2)
    run(16, ("trot", "bound", "pace", "gallop"), 1.5, 2048, 3, nseg=10)
    run(10, ("stand",), 2.0, 1024, 4)
    for shape in (0, 1, 3):
        for qcap in (32, 64):
            os.environ["CMPC_SHAPE"] = str(shape); os.environ["CMPC_QCAP1"] = str(qcap)
            print("SHAPE", shape, "QCAP1", qcap)
            run(10, ("trot",), 1.0, 16384, 5, ncheck=8)
