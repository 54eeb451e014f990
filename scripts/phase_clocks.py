"""Per-phase SM cycles of the solve kernel on the bench workload (thread-0 clocks summed over CTAs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine

def run(B, h=10, gaits=("trot",), spread=1.0, nseg=None, launches=4, tag=""):
    inst = synth.make_batch(B, horizon=h, seed=1000, gaits=gaits, spread=spread, n_segment=nseg)
    b = engine.Batch(B); b.setup(0.03, h, 0.4, 120.0); b.upload(inst)
    for _ in range(3):
        b.solve()
    b.sync()
    plain = []
    for _ in range(launches):
        b.solve(); b.sync(); plain.append(b.last_solve_ms())
    b.enable_phase_clocks(True)
    ms = []
    for _ in range(launches):
        b.solve(); b.sync(); ms.append(b.last_solve_ms())
    cyc = b.phase_cycles()
    res = b.download()
    print("   iterations histogram:", np.bincount(np.minimum(res["iterations"], 40)).tolist())
    tot = sum(cyc.values())
    print("%s B=%d h=%d: kernel %.3f ms plain, %.3f ms with clocks; iters mean %.1f max %d; cycles/instance %.0f"
          % (tag, B, h, min(plain), min(ms), res["iterations"].mean(), res["iterations"].max(), tot / (B * launches)))
    for k, v in cyc.items():
        print("   %-6s %5.1f%%  %8.0f cyc/instance" % (k, 100.0 * v / tot, v / (B * launches)))
    b.close()

if __name__ == "__main__":
    run(4096, tag="trot")
    run(16384, tag="trot")
    if len(sys.argv) > 1:
        run(2048, h=16, gaits=("trot", "bound", "pace", "gallop"), spread=1.5, nseg=10, tag="mixed16")
