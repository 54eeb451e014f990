#!/usr/bin/env python
"""bench.py — cMPC QP solves/sec (horizon 10, batched A1 trot) on N B200s, next to the
reference's qpOASES path on the host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode headline|sweep1m]

  --mode headline (default): weak scaling, 4096 instances per GPU per step (BASELINE.json configs[1]).
  --mode sweep1m: BASELINE.json configs[4] — ONE fixed batch of 2^20 instances sharded over the N ranks
             (shard_bounds), a step is one pass over the whole million; strong scaling.

One "step" = one pass of the condensation + inversion + QP kernels over one batch of 4096 synthetic
randomised A1 trot instances (BASELINE.json configs[1]) per GPU.  Instances are independent, so the
path shards with no collective (weak scaling: every rank solves its own 4096-instance batches).

  value      solves/s with the instance records resident in HBM.  A ring of RING distinct batches
             (> L2 in total) is uploaded once; timed steps walk the ring so every step reads cold
             records.  CUDA events on the engine's stream, max over ranks.
  e2e        the same metric through the host-buffer calls (cmpc_batch_bind_host, then cmpc_batch_submit_bound /
             cmpc_batch_wait_bound with eight batches in flight — throughput, as `value` and the reference arm are):
             every step the device reads the pinned input arrays over PCIe and packs the records, and
             forces/objective/status/iterations/activity mask land in pinned host arrays, all inside the timed
             region.  e2e.synchronous is one cmpc_batch_solve_bound call at a time (latency-oriented use),
             e2e.two_in_flight two batches, e2e.commands the controller-level call one level up (row (f)).
  roofline   the FP64 tensor-core inversion kernel (K = H^-1 by blocked sweeps of DMMA m8n8k4, ~90 % of the
             step's flops) against the measured FP64 peak of the device; every kernel class of the step
             (assembly, inversion, dual active set) is listed beside it with its own time, flops and
             fraction, and so is the whole step.  The path is not HBM bound; the HBM fraction is reported.
  cpu_baseline / --impl reference
             the reference's CPU path: restated fp32 condensation + the reference's real
             qpOASES 3.2.0 (oracle/_ref), one solve per thread on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "quad-periodic-mpc_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

HORIZON, DT, BATCH = 10, 0.03, 4096
SWEEP_TOTAL = 1 << 20
METRIC = "cmpc_qp_solves_per_sec_h10_batched"
L2_BYTES = 126 * 1024 * 1024
# dram__bytes_read.sum + dram__bytes_write.sum of one launch (bytes), from the committed ncu --set full capture
# (profiles/r2_ncu_full_summary.txt; cold cache: ncu flushes L2 between kernels)
NCU_TRAFFIC = {"assemble": 25.9e6, "invert": 97.5e6, "dual": 29.3e6, "fused": None}


WORKLOADS = {
    "headline": "batch 4096 A1 trot instances, horizon 10, dt 0.03, randomised states (BASELINE.json configs[1]) per GPU per step",
    "sweep1m": "one fixed batch of 1048576 A1 trot instances, horizon 10, dt 0.03, randomised states, sharded over the "
               "ranks (BASELINE.json configs[4])",
}
# the CPU arm is a port (restated fp32 condensation) around the reference's REAL qpOASES 3.2.0 (compiled from its sources)
CPU_KIND = "port+reference-qpOASES"


def shard_bounds(total, rank, world):
    """Contiguous split of `total` independent instances over `world` ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_max(units, seconds):
    """(sum of units, max of seconds) over ranks; identity without torch.distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return units, seconds
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    a = torch.tensor([units], dtype=torch.float64, device=dev)
    b = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    dist.all_reduce(b, op=dist.ReduceOp.MAX)
    return a.item(), b.item()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def single_solve_latency(device, reps=1000):
    """BASELINE.json configs[0]: one update_problem_data_floats -> get_solution round trip through the reference's own
    single-instance interface (convexMPC_interface.h:44-52): pageable inputs, host packing, one instance, the whole
    kernel chain, the twelve first-step forces read back.  Microseconds, wall clock."""
    from cmpc_b200 import engine, synth
    L = engine.lib()
    os.environ["CMPC_DEVICE"] = str(device)
    inst = synth.make_batch(64, horizon=HORIZON, dt=DT, seed=77)
    L.cmpc_reset_history()
    L.setup_problem(DT, HORIZON, inst["mu"], inst["f_max"])
    args = []
    for i in range(64):
        a = [np.array(inst[k][i], dtype=np.float32) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
        a.append(np.ascontiguousarray(inst["gait"][i], dtype=np.int32))
        args.append(a)
    lat = []
    fext = np.zeros(6, dtype=np.float32)
    for k in range(reps + 20):
        p, v, q, w, r, wt, tr, gait = args[k % 64]
        # what the controller does before every solve: the two globals of the adaptive hook (f_ext, simulation_time)
        fext[3] = 0.3 + 0.8 * np.sin(2 * np.pi * 0.7 * DT * k)
        t0 = time.perf_counter()
        L.cmpc_set_external_force(fext.ctypes.data)
        L.cmpc_set_simulation_time(DT * k)
        L.update_problem_data_floats(p.ctypes.data, v.ctypes.data, q.ctypes.data, w.ctypes.data, r.ctypes.data, 0.0, 0.0,
                                     0.0, wt.ctypes.data, tr.ctypes.data, float(inst["alpha"][k % 64]), gait.ctypes.data)
        f = [L.get_solution(j) for j in range(12)]
        t1 = time.perf_counter()
        if k >= 20:
            lat.append(1e6 * (t1 - t0))
    assert all(np.isfinite(f))
    L.cmpc_reset_history()
    lat.sort()
    return {"median": lat[len(lat) // 2], "p99": lat[int(0.99 * len(lat))], "min": lat[0], "reps": reps,
            "call": "setup once, then update_problem_data_floats + 12 x get_solution per solve (one A1 trot instance, "
                    "horizon 10) through the reference's single-instance interface; includes the ctypes call overhead"}


def cpu_reference_run(steps, warmup, sample, threads=None):
    """The reference's CPU implementation of the path on the host cores (oracle/_ref)."""
    from cmpc_b200 import synth
    from oracle import cmpc_oracle as O
    if not O.available():
        raise RuntimeError("oracle/_ref/libcmpc_ref.so is missing: build it where /root/reference exists")
    threads = threads or len(os.sched_getaffinity(0))
    inst = synth.make_batch(sample, horizon=HORIZON, dt=DT, seed=1234)
    st = O.make_setup(DT, HORIZON, inst["mu"], inst["f_max"])
    ups = (O.Update * sample)(*[O.make_update(inst, i, HORIZON) for i in range(sample)])
    for _ in range(warmup):
        O.solve_batch(st, ups, threads, use_float=True)
    t0 = time.perf_counter()
    ok_all = 0
    for _ in range(steps):
        _, ok = O.solve_batch(st, ups, threads, use_float=True)
        ok_all += int(ok.sum())
    dt = time.perf_counter() - t0
    return {"value": steps * sample / dt, "seconds": dt, "threads": threads, "sample": sample, "solved": ok_all}


def run_reference(args, rank, world):
    if rank != 0:
        return
    # the same per-step batch as the GPU arm's headline workload; in sweep1m mode a step of the GPU arm is 2^20 solves,
    # here a bounded 4096-instance sample of the same distribution (a million qpOASES solves take ~25 s per step)
    sample = BATCH
    r = cpu_reference_run(args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 condensation + f64 qpOASES",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.mode], "batch_per_gpu": sample, "horizon": HORIZON,
                   "sample": "every step solves %d instances of the workload on the host cores" % sample},
        "cpu_baseline": {"value": r["value"], "unit": "solves/s", "cores": r["threads"], "kind": CPU_KIND,
                         "sample": "%d instances of the bench workload per step, qpOASES 3.2.0 compiled from the "
                                   "reference's sources + restated fp32 condensation, one solve per thread" % sample},
        "e2e": {"value": r["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pin_rank_to_cores(local_rank, world):
    """N ranks on one host: give every rank its own slice of the cores this process may use, so the eight enqueue
    threads (and each engine's host workers) do not migrate across each other."""
    # measured on the 8-GPU box (32 cores, one NUMA node; profiles/r2_e2e_scale_n8.txt): four cores per rank are too few
    # for a rank's enqueue thread plus the CUDA runtime's own threads — 7.84 M solves/s per GPU pinned against 8.50 M
    # unpinned — so pinning is opt-in
    if world <= 1 or os.environ.get("CMPC_BENCH_PIN", "0") != "1":
        return None
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // world)
    mine = cores[local_rank * per:(local_rank + 1) * per] or cores
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from cmpc_b200 import engine, synth
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    cores = pin_rank_to_cores(local_rank, world)
    sweep = args.mode == "sweep1m"
    h = HORIZON
    rec_bytes = 4 * (48 + 12 * h) + 4 * h
    rec_bytes = (rec_bytes + 15) & ~15
    out_bytes = 12 * h * 8 + 20 * h + 8 + 4 + 4
    if sweep:
        # ONE fixed batch of 2^20 instances, cut over the ranks; a step is one pass over the whole batch (this rank:
        # its shard, in launches of 65536 instances that rotate through the engine's streams)
        lo, hi = shard_bounds(SWEEP_TOTAL, rank, world)
        total = hi - lo
        launch = 65536
        ring = -(-total // launch)
        per_step_units = float(total)
    else:
        ring = max(2, -(-int(1.25 * L2_BYTES) // (BATCH * (rec_bytes + out_bytes))))
        total = ring * BATCH
        launch = BATCH
        per_step_units = float(BATCH)
    inst = synth.make_batch(total, horizon=h, dt=DT, seed=1000 + rank)
    b = engine.Batch(total, device=local_rank)
    b.setup(DT, h, inst["mu"], inst["f_max"])
    b.upload(inst)
    b.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        b.sync()

    def resident_step(i):
        if sweep:
            for k0 in range(0, total, launch):
                b.solve_range(k0, min(launch, total - k0))
        else:
            b.solve_range((i % ring) * BATCH, BATCH)

    # ---- device-resident throughput ----
    for i in range(args.warmup):
        resident_step(i)
    b.sync()
    b.reset_counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    b.mark(0)
    for i in range(args.steps):
        resident_step(i + args.warmup)
    b.mark(1)
    barrier()
    region_ms = b.marked_ms()          # CUDA events on the launching stream around exactly K steps
    launches = b.launches()
    flops_total = b.last_flops()
    units, seconds = allreduce_sum_max(args.steps * per_step_units, region_ms / 1e3)
    value = units / seconds
    # every instance the timed region touched reached its optimum (the ring covers them all)
    chk = b.download(want_active=False)
    touched = total if (sweep or args.steps + args.warmup >= ring) else (args.steps + args.warmup) * BATCH
    assert (chk["status"][:touched] == engine.ST_SOLVED).all(), "resident loop: %d instances not solved" % (
        chk["status"][:touched] != engine.ST_SOLVED).sum()
    f = chk["forces"][:touched].reshape(touched, -1, 3)
    assert np.isfinite(f).all() and (f[..., 2] >= -1e-8).all() and (f[..., 2] <= inst["f_max"] + 1e-8).all()
    del chk, f

    # ---- end to end through the host-buffer call: forces, objective, status, iterations AND the activity mask ----
    if sweep:
        be, sub, e2e_n = b, inst, total     # the shard itself, through the same handle (a second one would double its HBM)
    else:
        sub = {k: (v[:BATCH] if isinstance(v, np.ndarray) else v) for k, v in inst.items()}
        be = engine.Batch(BATCH, device=local_rank)
        be.setup(DT, h, inst["mu"], inst["f_max"])
        e2e_n = BATCH
    be.prepare_host(sub, want_active=True)   # host arrays bound once; every step below is one C-ABI call
    for _ in range(max(1, args.warmup if not sweep else 1)):
        res = be.solve_prepared()
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = be.solve_prepared()                # inputs read over PCIe -> records -> kernels -> results in host memory
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_units, e2e_seconds = allreduce_sum_max(float(e2e_steps * e2e_n), e2e_wall)
    assert (res["status"] == 0).all()
    h2d = int(sum(a.nbytes for a in be._prepared_inputs.values()))   # the eleven input arrays the device reads over PCIe
    d2h = int(sum(a.nbytes for a in res.values()))                    # forces, objective, status, iterations, active

    # ---- the same call with D batches in flight (submit / wait): step k is submitted before step k-D+1 is waited for ----
    def in_flight(depth):
        pipe = []
        # from four batches in flight on, the copy engine carries the results better than the kernels' own PCIe stores
        # (profiles/r2_submit_copy.txt): the documented option for deep pipelines
        opts = {"submit_copy": 1} if depth >= 4 else {}
        for k in range(depth):
            sk = {kk: (v[(k + 1) * BATCH:(k + 2) * BATCH] if isinstance(v, np.ndarray) else v) for kk, v in inst.items()}
            bk = engine.Batch(BATCH, device=local_rank, options=opts)
            bk.setup(DT, h, inst["mu"], inst["f_max"])
            bk.prepare_host(sk, want_active=True)
            pipe.append(bk)
        for k in range(depth * max(1, args.warmup)):
            pipe[k % depth].solve_prepared()
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            if k >= depth:
                rp = pipe[k % depth].wait_prepared()
            pipe[k % depth].submit_prepared()
        for k in range(max(args.steps, depth), max(args.steps, depth) + min(depth, args.steps)):
            rp = pipe[k % depth].wait_prepared()
        barrier()
        wall = time.perf_counter() - t0
        units, secs = allreduce_sum_max(float(args.steps * BATCH), wall)
        assert (rp["status"] == 0).all()
        for bk in pipe:
            bk.close()
        return units, secs

    extra = {}
    if not sweep:
        pipe_units, pipe_seconds = in_flight(2)
        deep = min(8, ring - 1)
        deep_units, deep_seconds = in_flight(deep)
        extra["two_in_flight"] = {"value": pipe_units / pipe_seconds, "unit": "solves/s",
                                  "ms_per_step": 1e3 * pipe_seconds / args.steps,
                                  "call": "cmpc_batch_submit_bound / cmpc_batch_wait_bound on two batches: every step "
                                          "still reads its inputs from and writes its results to pinned host arrays"}
        extra["deep_in_flight"] = {"value": deep_units / deep_seconds, "unit": "solves/s", "batches_in_flight": deep,
                                   "ms_per_step": 1e3 * deep_seconds / args.steps,
                                   "call": "cmpc_batch_submit_bound / cmpc_batch_wait_bound on %d engine handles (step k is submitted before step k-%d is waited for), option submit_copy = 1: results by the copy engine; every step reads its inputs from and writes its results to pinned host arrays (profiles/r2_submit_copy.txt)" % (deep, deep - 1)}
        # ---- end to end one level up: the controller-level call (updateMPCIfNeeded / solveDenseMPC on the device) ----
        cmds = synth.make_commands(BATCH, engine.COMMAND_DTYPE, horizon=h, gaits=("trot",), seed=2000 + rank)
        cres = np.zeros(BATCH, dtype=engine.RESULT_DTYPE)
        bc = engine.Batch(BATCH, device=local_rank)
        bc.setup(DT, h, inst["mu"], inst["f_max"])
        for a in (cmds, cres):
            engine.lib().cmpc_host_register(a.ctypes.data, a.nbytes)
        for _ in range(max(1, args.warmup)):
            bc.solve_commands(cmds, results=cres)
        csteps = max(1, min(args.steps, 200))
        barrier()
        t0 = time.perf_counter()
        for _ in range(csteps):
            bc.solve_commands(cmds, results=cres)
        barrier()
        cmd_wall = time.perf_counter() - t0
        cmd_units, cmd_seconds = allreduce_sum_max(float(csteps * BATCH), cmd_wall)
        assert (cres["status"] == 0).all()
        for a in (cmds, cres):
            engine.lib().cmpc_host_unregister(a.ctypes.data)
        bc.close()
        extra["commands"] = {"value": cmd_units / cmd_seconds, "unit": "solves/s", "steps": csteps,
                             "h2d_bytes_per_step": int(cmds.nbytes), "d2h_bytes_per_step": int(cres.nbytes),
                             "ms_per_step": 1e3 * cmd_seconds / csteps,
                             "call": "cmpc_batch_solve_commands: the controller-level call (ConvexMPCLocomotion::"
                                     "updateMPCIfNeeded + solveDenseMPC + getMpcTable on the device), one cmpc_command "
                                     "in and one cmpc_command_result out per robot"}
    clocks = sampler.stop()            # sampled through the timed regions (resident, e2e, e2e commands)
    # `value` above is THROUGHPUT (eight batches in flight on the device), and so is the reference arm (all host threads over
    # the sample): the end-to-end figure that compares with both is the host-buffer call with batches in flight, every
    # step's inputs read from and results written to pinned host memory inside the timed region.  The one-call-at-a-time
    # figure (latency-oriented use) stays next to it as e2e.synchronous.
    sync_block = {"value": e2e_units / e2e_seconds, "unit": "solves/s", "ms_per_step": 1e3 * e2e_seconds / e2e_steps,
                  "call": "cmpc_batch_solve_bound: one synchronous call per step, pinned host arrays in the layout of "
                          "update_problem_data (update_data_t); the device reads the inputs over PCIe and packs the "
                          "records, the kernels write the results into host memory"}
    if "deep_in_flight" in extra:
        d = extra.pop("deep_in_flight")
        e2e_block = dict({"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "ms_per_step": d["ms_per_step"], "batches_in_flight": d["batches_in_flight"],
                          "outputs": "forces, objective, status, iterations, active mask", "call": d["call"],
                          "synchronous": sync_block}, **extra)
    else:
        e2e_block = dict(dict(sync_block, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                              outputs="forces, objective, status, iterations, active mask"), **extra)
    if sweep:
        be.release_prepared()

    # ---- per-kernel-class device time: serial solves (CUDA events between the classes) on cold batches ----
    kflops = b.kernel_flops()                     # algorithmic flops of the timed region, per kernel class
    prof_n = min(8, ring)
    kms = {k: 0.0 for k in engine.Batch.KERNELS}
    launches_per_step = ring if sweep else 1
    for i in range(prof_n):
        k0 = ((i + 3) % ring) * launch
        t = b.profile_range(k0, min(launch, total - k0))
        for k in kms:
            kms[k] += t[k] / prof_n

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        fp64 = engine.measure_fp64_peak(local_rank)
        dmma = engine.measure_dmma_peak(local_rank)
        step_ms = region_ms / args.steps
        names = {"assemble": "cmpc_assemble_mma_kernel", "invert": "cmpc_invert_ws_kernel",
                 "dual": "cmpc_lpt_order_kernel + cmpc_dual_fast_kernel + cmpc_dual_team_kernel (working sets beyond the first tier, CTA per instance)", "fused": "cmpc_solve_kernel"}
        bounds = {"assemble": "latency / issue (FP64 FMA + shared memory)", "invert": "tensor (FP64 DMMA)",
                  "dual": "latency (dependent chain per active-set iteration)", "fused": "latency"}
        peak_of = {"assemble": fp64, "invert": dmma, "dual": fp64, "fused": fp64}
        kernels = []
        for k in engine.Batch.KERNELS:
            if kms[k] <= 0.0:
                continue
            fl = kflops[k] / args.steps / launches_per_step
            ach = fl / (kms[k] * 1e-3) / 1e12
            kernels.append({"kernel": names[k], "ms": kms[k], "flops_per_launch": fl, "achieved_tflops": ach,
                            "frac_of_fp64_peak": ach / peak_of[k] if peak_of[k] else None, "bound": bounds[k],
                            "share_of_serial_time": kms[k] / sum(kms.values())})
        fl_step = sum(kflops.values()) / args.steps
        dom = "invert" if kms["invert"] > 0 else max(kms, key=kms.get)   # the flop-dominant kernel (80 % of the step's flops)
        tdom = max(kms, key=kms.get)                                       # the time-dominant kernel class
        fl_dom = kflops[dom] / args.steps / launches_per_step
        achieved = fl_dom / (kms[dom] * 1e-3) / 1e12
        fl_t = kflops[tdom] / args.steps / launches_per_step
        hbm_bytes = (total if sweep else BATCH) * (rec_bytes + out_bytes)
        hbm_ach = hbm_bytes / (step_ms * 1e-3) / 1e9
        cpu = None
        try:
            c = cpu_reference_run(10, 1, 4096)
            cpu = {"value": c["value"], "unit": "solves/s", "cores": c["threads"], "kind": CPU_KIND,
                   "sample": "10 x 4096 instances of the bench workload (about 20 core-seconds), qpOASES 3.2.0 built "
                             "from the reference's sources + restated fp32 condensation, one solve per thread on all "
                             "host cores"}
        except Exception as e:  # the checker library did not travel
            cpu = {"value": None, "unit": "solves/s", "cores": 0, "kind": CPU_KIND, "sample": "unavailable: %s" % e}
        latency = single_solve_latency(local_rank) if not sweep else None
        cfg = {"workload": WORKLOADS[args.mode], "batch_per_gpu": total if sweep else BATCH, "horizon": h,
               "l2": "inputs rotate through a ring of %d distinct resident batches (%.0f MB > L2)"
                     % (ring, total * (rec_bytes + out_bytes) / 1e6) if not sweep else
                     "every step reads this rank's whole shard (%.0f MB of records and outputs, > L2 up to 8 ranks)"
                     % (total * (rec_bytes + out_bytes) / 1e6),
               "sharding": "independent instances, no data-path collective" +
                           (" (bench.shard_bounds over %d ranks)" % world if sweep else ""),
               "pipelining": "successive launches rotate through the engine's eight CUDA streams (a launch is three "
                             "kernels: assembly, FP64-tensor inversion, dual active set, plus the hardest-first "
                             "ordering pass and the working-set overflow launch)",
               "host_cores_per_rank": len(cores) if cores else None}
        if sweep:
            cfg["launch_instances"] = launch
        else:
            cfg["ring_batches"] = ring
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True,
            "scaling": "strong" if sweep else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": dmma, "unit": "TFLOP/s",
                         "frac": achieved / dmma if dmma else None,
                         "traffic": NCU_TRAFFIC.get(dom),
                         "kernel": names[dom], "kernel_ms": kms[dom],
                         "peak_source": "FP64 DMMA m8n8k4 microbenchmark run in this process (cmpc_measure_dmma_peak: "
                                        "independent accumulators on every SM sub-partition); the FP64 FMA peak measured "
                                        "the same way is fp64_fma_peak; MEASURED_PEAKS.json holds no FP64 figure",
                         "fp64_fma_peak": fp64, "fp64_dmma_peak": dmma,
                         "flops_per_launch": fl_dom,
                         "flops_note": "algorithmic: n^3 + 2 n^2 per instance (symmetric inverse + x0 = -K g), n = 60; "
                                       "the kernel executes 1.6x that (padding to 64, full diagonal tiles, D^-1 C)",
                         "timing": "CUDA events around each kernel class of %d serial launches of cold batches" % prof_n,
                         "traffic_note": "dram bytes of one ncu --set full launch (profiles/); ncu flushes L2 between "
                                         "kernels, in the pipeline the tiles written by the assembly kernel are L2 hits",
                         "time_dominant": {"kernel": names[tdom], "ms": kms[tdom],
                                           "share_of_serial_time": kms[tdom] / sum(kms.values()),
                                           "flops_per_launch": fl_t, "achieved": fl_t / (kms[tdom] * 1e-3) / 1e12,
                                           "frac": fl_t / (kms[tdom] * 1e-3) / 1e12 / peak_of[tdom], "bound": bounds[tdom]},
                         "step": {"flops_per_step": fl_step, "achieved": fl_step / (step_ms * 1e-3) / 1e12,
                                  "frac": (fl_step / (step_ms * 1e-3) / 1e12) / fp64 if fp64 else None,
                                  "ms": step_ms, "serial_ms": sum(kms.values()) * launches_per_step},
                         "kernels": kernels,
                         "hbm": {"achieved": hbm_ach, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": hbm_ach / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                                 "bytes_per_launch": hbm_bytes, "peak_source": peak_kind}},
            "cpu_baseline": cpu,
            "e2e": e2e_block,
            "gpu_launches": launches, "clocks": clocks,
        }
        if latency:
            line["latency_single_us"] = latency
        print(json.dumps(line), flush=True)
    b.close()
    if be is not b:
        be.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--mode", default="headline", choices=["headline", "sweep1m"])
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.steps is None:
        args.steps = 500 if args.mode == "headline" else 10
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
