import os, sys, time
sys.path.insert(0, "quad-periodic-mpc_b200"); sys.path.insert(0, ".")
import numpy as np
from cmpc_b200 import synth, engine
from oracle import cmpc_frontend as F
B=4096
cmds = synth.make_commands(B, engine.COMMAND_DTYPE, horizon=10, gaits=("trot",), seed=2000)
b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0)
res = b.solve_commands(cmds)
it = res["iterations"]
print("commands: iters mean %.2f max %d  >=32: %d  hist %s" % (it.mean(), it.max(), (it>=32).sum(), np.bincount(np.minimum(it,60))[:61].tolist()))
inst = synth.make_batch(B, horizon=10, seed=1000)
r2 = b.solve_host(inst)
it = r2["iterations"]
print("make_batch: iters mean %.2f max %d  >=32: %d" % (it.mean(), it.max(), (it>=32).sum()))
ins, ex = F.solver_inputs(cmds, 10, 0.03, np.zeros((B,6),np.float32))
for k in ("p","v","w","r","traj"):
    print(k, "cmd std", np.std(ins[k],axis=0)[:6], "batch std", np.std(inst[k],axis=0)[:6])
print("traj err z cmd", np.abs(ins["traj"][:,5]-ins["p"][:,2]).mean(), "batch", np.abs(inst["traj"][:,5]-inst["p"][:,2]).mean())
print("traj err xy cmd", np.abs(ins["traj"][:,3]-ins["p"][:,0]).mean(), "batch", np.abs(inst["traj"][:,3]-inst["p"][:,0]).mean())
print("traj vel cmd", np.abs(ins["traj"][:,9]).mean(), "batch", np.abs(inst["traj"][:,9]).mean())
print("traj yaw err cmd", np.abs(ins["traj"][:,2]-cmds["rpy"][:,2]).mean())
