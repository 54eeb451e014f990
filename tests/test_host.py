"""CPU tests of the host side: the library builds, loads and exports every symbol of
include/cmpc_b200.h, refuses to run without a GPU, and the host-side mirrors behave."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cmpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:void|int|double|char\s*\*|const char\s*\*)\s*\*?\s*(\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(built_lib):
    syms = declared_symbols()
    assert "setup_problem" in syms and "cmpc_batch_solve_host" in syms and len(syms) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    lib = C.CDLL(built_lib)
    for s in syms:
        getattr(lib, s)


def build_reference_caller(built_lib):
    """g++ translation unit that includes only the reference's declarations, linked against libcmpc_b200.so."""
    src = os.path.join(ROOT, "tests", "cxx", "reference_caller.cpp")
    out = os.path.join(ROOT, "build", "reference_caller")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libdir = os.path.dirname(built_lib)
    subprocess.run(["g++", "-O1", "-std=c++14", src, "-o", out, "-L" + libdir, "-lcmpc_b200", "-Wl,-rpath," + libdir],
                   check=True)
    return out


def test_unmodified_cxx_caller_links(built_lib):
    """convexMPC_interface.h:52 declares update_x_drag with C++ linkage: both names must be exported, and a g++
    caller that sees only the reference's declarations must link (it is RUN by the gpu tests)."""
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    assert "update_x_drag" in exported and "_Z13update_x_dragf" in exported
    exe = build_reference_caller(built_lib)
    und = subprocess.run(["nm", "-u", exe], check=True, capture_output=True, text=True).stdout
    assert "_Z13update_x_dragf" in und and "setup_problem" in und


def test_library_reads_no_environment_on_solve_paths():
    """Diagnostic switches exist only through cmpc_batch_set_option; the one getenv left in the default build is
    CMPC_DEVICE of the single-instance reference interface (read once, at the first setup_problem)."""
    csrc = os.path.join(ROOT, "quad-periodic-mpc_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        text = open(os.path.join(csrc, f)).read()
        # drop the -DCMPC_EXPERIMENTS-only regions
        text = re.sub(r"#ifdef CMPC_EXPERIMENTS.*?#e(?:lse|ndif)", "", text, flags=re.S)
        for m in re.finditer(r'getenv\("(\w+)"\)', text):
            assert m.group(1) == "CMPC_DEVICE", (f, m.group(1))


def test_kernel_image_is_sm100a_with_bulk_copy(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], check=True, capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass      # cp.async.bulk record prefetch
    assert "DFMA" in sass        # FP64 pipe
    assert "DMMA.8x8x4" in sass  # FP64 tensor-core sweep of the inversion kernel
    assert "REDUX" in sass       # warp argmin of the dual active-set kernel


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_cpu_fallback(built_lib):
    from cmpc_b200 import engine
    with pytest.raises(RuntimeError) as e:
        engine.Batch(4)
    assert "no CUDA device" in str(e.value) or "failed" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "quad-periodic-mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/cmpc_numpy", ""), os.path.join(dirpath, f)


def test_mpc_table_mirror_of_offset_duration_gait():
    from cmpc_b200 import synth
    # trot, 10 segments, iteration 0: legs 0,3 start in stance at phase 1..; Gait.cpp:158-187
    tab = synth.mpc_table(10, (0, 5, 5, 0), (5, 5, 5, 5), 0, 10)
    assert tab.shape == (10, 4)
    assert (tab.sum(1) == 2).all()                # trot: exactly two feet down at every step
    assert (tab[:, 0] == tab[:, 3]).all() and (tab[:, 1] == tab[:, 2]).all()
    assert (tab[:, 0] + tab[:, 1] == 1).all()
    assert list(tab[:, 0]) == [1, 1, 1, 1, 0, 0, 0, 0, 0, 1]
    stand = synth.mpc_table(10, (0, 0, 0, 0), (10, 10, 10, 10), 3, 10)
    assert stand.all()
    # periodicity in the iteration counter
    a = synth.mpc_table(10, (0, 2, 7, 9), (4, 4, 4, 4), 3, 16)
    b = synth.mpc_table(10, (0, 2, 7, 9), (4, 4, 4, 4), 13, 16)
    assert (a == b).all()


def test_synth_shapes_and_determinism():
    from cmpc_b200 import synth
    a = synth.make_batch(16, horizon=10, seed=5)
    b = synth.make_batch(16, horizon=10, seed=5)
    for k in ("p", "q", "traj", "gait"):
        assert (a[k] == b[k]).all()
    assert a["traj"].shape == (16, 120) and a["gait"].shape == (16, 40) and a["gait"].dtype == np.uint8
    np.testing.assert_allclose(np.linalg.norm(a["q"], axis=1), 1.0, atol=1e-6)


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    lo, hi = bench.shard_bounds(1000, rank, world)
    t = torch.tensor([float(hi - lo), 10.0 + rank])
    total = bench.allreduce_sum_max(t[0].item(), t[1].item())
    q.put((rank, lo, hi, total))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_covers_the_batch_once():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    (r0, lo0, hi0, tot0), (r1, lo1, hi1, tot1) = got
    assert lo0 == 0 and hi0 == lo1 and hi1 == 1000
    assert tot0 == tot1 == (1000.0, 11.0)   # units summed over ranks, time max over ranks
