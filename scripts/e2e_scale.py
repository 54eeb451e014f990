"""What limits the end-to-end call when N ranks share one host: the synchronous bound call on every rank at once, with
variants that take pieces of the PCIe traffic away.  Launch with torchrun (one rank per GPU); rank 0 prints, per variant,
the aggregate rate (units of all ranks / slowest rank's wall time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quad-periodic-mpc_b200")); sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch, torch.distributed as dist
import bench
from cmpc_b200 import synth, engine

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pin = os.environ.get("PIN", "1") == "1"
cores = bench.pin_rank_to_cores(lr, world) if pin else None
B, h, steps = 4096, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 300
inst = synth.make_batch(B, horizon=h, seed=77 + rank)
L = engine.lib()
ALL = ("forces", "objective", "status", "iterations", "active")
variants = [("all outputs, zero-copy stores", ALL, {}, 0), ("all outputs, copy engine", ALL, {"d2h_copy": 1}, 0),
            ("status only", ("status",), {}, 0), ("all outputs, ranks de-phased by rank x 40 us", ALL, {}, 40),
            ("forces only", ("forces",), {}, 0)]
for tag, keys, opts, dephase in variants:
    b = engine.Batch(B, device=lr, options=opts); b.setup(0.03, h, 0.4, 120.0)
    s = b._inputs(inst, B)
    pinned = [a for a in b._keep.values() if L.cmpc_host_register(a.ctypes.data, a.nbytes) == 0]
    o, res = b._outputs(B, True)
    for k in ALL:
        if k not in keys: setattr(o, k, None)
        elif L.cmpc_host_register(res[k].ctypes.data, res[k].nbytes) == 0: pinned.append(res[k])
    engine._check(L.cmpc_batch_bind_host(b._h, C.byref(s), C.byref(o)), "bind")
    for _ in range(5): engine._check(L.cmpc_batch_solve_bound(b._h, B), "solve")
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    if dephase: time.sleep(rank * dephase * 1e-6)
    t0 = time.perf_counter()
    for _ in range(steps): L.cmpc_batch_solve_bound(b._h, B)
    wall = time.perf_counter() - t0
    units, secs = bench.allreduce_sum_max(float(steps * B), wall)
    if rank == 0:
        print("N=%d %-46s %.2f M solves/s aggregate, %.2f M per GPU, %.4f ms per call (cores per rank: %s)"
              % (world, tag, units / secs / 1e6, units / secs / 1e6 / world, 1e3 * secs / steps, len(cores) if cores else "unpinned"), flush=True)
    L.cmpc_batch_bind_host(b._h, None, None)
    for a in pinned: L.cmpc_host_unregister(a.ctypes.data)
    b.close()
if world > 1:
    dist.destroy_process_group()
