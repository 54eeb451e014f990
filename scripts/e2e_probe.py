import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import numpy as np
from cmpc_b200 import synth, engine
B = 4096
inst = synth.make_batch(B, horizon=10, seed=1000)
b = engine.Batch(B); b.setup(0.03, 10, 0.4, 120.0)
for _ in range(3): b.solve_host(inst, want_active=False)
L = engine.lib()
s = b._inputs(inst, B); o, res = b._outputs(B, False)
def t(f, n=20):
    t0 = time.perf_counter()
    for _ in range(n): f()
    return 1e3 * (time.perf_counter() - t0) / n
print("python solve_host     ms", t(lambda: b.solve_host(inst, want_active=False)))
print("C solve_host (reuse)  ms", t(lambda: L.cmpc_batch_solve_host(b._h, B, C.byref(s), C.byref(o))))
print("C upload+sync         ms", t(lambda: (L.cmpc_batch_upload(b._h, B, C.byref(s)), L.cmpc_batch_sync(b._h))))
print("C solve+sync          ms", t(lambda: (L.cmpc_batch_solve(b._h), L.cmpc_batch_sync(b._h))))
print("C download            ms", t(lambda: L.cmpc_batch_download(b._h, C.byref(o))))
print("_inputs python        ms", t(lambda: b._inputs(inst, B)))
print("_outputs python       ms", t(lambda: b._outputs(B, False)))
