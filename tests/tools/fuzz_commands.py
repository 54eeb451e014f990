"""Randomised configurations of the device-side caller (cmpc_batch_solve_commands) against the oracle's restatement
(oracle/cmpc_frontend.py): records bit for bit, command state, Fr_des / f_ff, and the forces against the
update_problem_data-level call on the oracle-built inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
from oracle import cmpc_frontend as F

def run(rounds=60, seed=42, verbose=True):
    """Returns (robots checked, configurations with a mismatch)."""
    rng = np.random.default_rng(seed)
    names = [g for g in synth.GAITS if g != "stand"]
    bad = total = 0
    for it in range(rounds):
        h = int(rng.integers(1, 20))
        gaits = tuple(rng.choice(names, size=int(rng.integers(1, 4)), replace=False))
        B = int(rng.choice([1, 2, 5, 33, 200, 1000]))
        dt = float(rng.choice([0.03, 0.026, 0.05]))
        c = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=h, gaits=gaits, seed=int(rng.integers(1 << 30)),
                                spread=float(rng.choice([0.5, 1.0, 2.0])), mixed_fraction=float(rng.choice([0.0, 0.3, 1.0])),
                                stand_fraction=float(rng.choice([0.0, 0.2])), with_log=bool(rng.random() < 0.7),
                                sim_time=float(rng.uniform(0, 10)))
        b = engine.Batch(B); b.setup(dt, h, 0.4, 120.0)
        res, forces = b.solve_commands(c, want_forces=True)
        inst, ex = F.solver_inputs(c, h, dt, np.zeros((B, 6), np.float32))
        rec_ok = bool((b.copy_records(0, B) == F.pack_records(inst, sim_time=c["sim_time"])).all())
        st_ok = bool((res["world_position_desired"] == ex["world_position_desired"]).all() and (res["x_comp_integral"] == ex["x_comp_integral"]).all()
                     and (res["f_ext"] == ex["f_ext"]).all())
        f, f_ff = F.leg_outputs(c, forces)
        out_ok = bool((res["fr_des"] == f).all() and (res["f_ff"] == f_ff).all())
        b2 = engine.Batch(B); b2.setup(dt, h, 0.4, 120.0)
        inst.update(mu=0.4, f_max=120.0)
        ref = b2.solve_host(inst)
        f_ok = bool((ref["forces"] == forces).all() and (ref["status"] == res["status"]).all())
        b.close(); b2.close()
        total += B
        ok = rec_ok and st_ok and out_ok and f_ok
        bad += 0 if ok else 1
        if verbose: print("h=%2d dt=%.3f gaits=%-26s B=%4d: records %s, command state %s, leg outputs %s, forces %s"
              % (h, dt, ",".join(gaits), B, rec_ok, st_ok, out_ok, f_ok), flush=True)
    if verbose:
        print("FUZZ commands: %d robots in %d configurations, %d configurations with a mismatch" % (total, rounds, bad))
    return total, bad


if __name__ == "__main__":
    run(int(os.environ.get("FUZZ_ROUNDS", "60")), int(os.environ.get("FUZZ_SEED", "42")))
