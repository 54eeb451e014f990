"""A/B of the two-kernel pipeline vs the fused kernel on the bench workload (kernel ms, parity of the two)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine

def run(B, h=10, gaits=("trot",), spread=1.0, nseg=None, tag="", envs=({},)):
    inst = synth.make_batch(B, horizon=h, seed=1000, gaits=gaits, spread=spread, n_segment=nseg)
    ref = None
    for env in envs:
        for k in ("CMPC_PATH", "CMPC_QCAP1", "CMPC_WPC", "CMPC_WS_MB", "CMPC_CSHAPE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        b = engine.Batch(B); b.setup(0.03, h, 0.4, 120.0); b.upload(inst)
        for _ in range(3):
            b.solve()
        b.sync()
        ms = []
        for _ in range(6):
            b.mark(0); b.solve(); b.mark(1); b.sync(); ms.append(b.marked_ms())
        res = b.download()
        if ref is None:
            ref = res
        df = np.abs(res["forces"] - ref["forces"]).max()
        print("%-8s %-40s B=%d h=%d: %.3f ms (%.2f M/s) status=%s iters mean %.1f max %d  max|dF| vs first=%.2e"
              % (tag, env, B, h, min(ms), B / min(ms) / 1e3, np.bincount(res["status"]).tolist(),
                 res["iterations"].mean(), res["iterations"].max(), df), flush=True)
        b.close()

if __name__ == "__main__":
    envs = [{"CMPC_PATH": "fused"}, {}, {"CMPC_QCAP1": "8"}, {"CMPC_QCAP1": "12"}, {"CMPC_QCAP1": "24"}, {"CMPC_QCAP1": "32"},
            {"CMPC_WPC": "2"}, {"CMPC_WPC": "8"}]
    run(4096, tag="trot", envs=envs)
    run(16384, tag="trot", envs=envs[:2])
    run(2048, h=16, gaits=("trot", "bound", "pace", "gallop"), spread=1.5, nseg=10, tag="mixed16", envs=envs[:2] + [{"CMPC_QCAP1": "32"}])
    run(1024, h=10, gaits=("stand",), spread=2.0, tag="stand", envs=envs[:2])
