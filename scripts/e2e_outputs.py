"""The synchronous end-to-end call (cmpc_batch_solve_bound, 4096 trot, pinned arrays) against WHICH result arrays are read back:\nhow much of the call the result bytes cost.  usage: e2e_outputs.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np, ctypes as C
from cmpc_b200 import synth, engine
B, h, steps = 4096, 10, 300
inst = synth.make_batch(B, horizon=h, seed=77)
L = engine.lib()
for tag, keys in (("all outputs", ("forces","objective","status","iterations","active")), ("no active", ("forces","objective","status","iterations")),
                  ("status only", ("status",)), ("forces only", ("forces",))):
    b = engine.Batch(B); b.setup(0.03, h, 0.4, 120.0)
    s = b._inputs(inst, B)
    pinned = []
    for a in b._keep.values():
        if L.cmpc_host_register(a.ctypes.data, a.nbytes) == 0: pinned.append(a)
    o, res = b._outputs(B, True)
    for k in ("forces","objective","status","iterations","active"):
        if k not in keys: setattr(o, k, None)
        elif L.cmpc_host_register(res[k].ctypes.data, res[k].nbytes) == 0: pinned.append(res[k])
    engine._check(L.cmpc_batch_bind_host(b._h, C.byref(s), C.byref(o)), "bind")
    for _ in range(5): engine._check(L.cmpc_batch_solve_bound(b._h, B), "solve")
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(steps): L.cmpc_batch_solve_bound(b._h, B)
        best = min(best, (time.perf_counter() - t0) / steps)
    print("%-14s sync %.4f ms/step  (%.2f M/s)  out bytes %d" % (tag, 1e3*best, B/best/1e6, sum(res[k].nbytes for k in keys)), flush=True)
    L.cmpc_batch_bind_host(b._h, None, None)
    for a in pinned: L.cmpc_host_unregister(a.ctypes.data)
    b.close()
