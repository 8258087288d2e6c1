"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): z-slab decomposition with NCCL ghost-plane
exchange and allreduce must reproduce the oracle (and hence the single-GPU run) to the same tolerances."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_worker(nproc, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    cmd += [str(a) for a in args]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    for line in p.stdout.splitlines():
        if line.startswith("MGPU_RESULT "):
            return json.loads(line[len("MGPU_RESULT "):])
    raise AssertionError("worker failed:\n" + p.stdout[-3000:] + "\n" + p.stderr[-3000:])


def ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
@pytest.mark.parametrize("args", [
    ("--refine", 5, "--rtol", 1e-10),                                   # 65^3: levels <= 65^3 replicated
    ("--refine", 6, "--rtol", 1e-10, "--levels", 5),                    # 129^3 distributed, coarse 9^3
    ("--refine", 6, "--rtol", 1e-10, "--levels", 5, "--march-min-plane", 1),   # plane-marching kernel on slabs
    ("--refine", 5, "--rtol", 1e-8, "--cycle", "w"),
    ("--refine", 6, "--rtol", 1e-10, "--levels", 5, "--comm-peer", 0),          # NCCL send/recv/allreduce path
    ("--refine", 5, "--rtol", 1e-8, "--cycle", "w", "--rep-points", 1, "--comm-peer", 0),
    ("--refine", 5, "--rtol", 1e-10, "--rep-points", 1, "--march-min-plane", 1),   # every level that can be is distributed
])
def test_slab_solve_matches_oracle(nproc, args):
    if ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = run_worker(nproc, *args)
    assert r["world"] == nproc
    assert r["its"] == r["oracle_its"]
    assert r["hist_rel"] is not None and r["hist_rel"] < 1e-10
    assert r["sol_rel"] < 1e-12
    assert r["bnorm_rel"] < 1e-13


@pytest.mark.parametrize("nproc", [2, 4, 8])
@pytest.mark.parametrize("args", [
    ("--refine", 6, "--rtol", 1e-10, "--levels", 5, "--march-min-plane", 1, "--repeat", 3),
    ("--refine", 5, "--rtol", 1e-10, "--rep-points", 1, "--repeat", 2),      # small distributed levels: generic kernels
    ("--refine", 5, "--rtol", 1e-8, "--cycle", "w", "--rep-points", 1),
])
def test_fused_exchange_is_bit_identical_to_push_kernels(nproc, args):
    """Ghost planes pushed by the producing kernel and awaited by the consuming kernel (comm.h HaloPort) must give the
    same bits as one exchange kernel per halo (both in one process, fresh hierarchies), and the oracle's answer."""
    if ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = run_worker(nproc, *args, "--fused-halo", 1, "--compare-fused")
    assert r["fused_equal"] is True
    assert r["its"] == r["oracle_its"]
    assert r["hist_rel"] < 1e-10 and r["sol_rel"] < 1e-12


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_slab_solve_matches_c_oracle_at_257(nproc):
    """BASELINE config 2's grid (257^3, -pc_mg_levels 6) on slabs, against oracle/fish_cpu.c directly."""
    if ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = run_worker(nproc, "--refine", 7, "--levels", 6, "--rtol", 1e-10, "--c-oracle")
    assert r["its"] == r["oracle_its"]
    assert r["hist_rel"] is not None and r["hist_rel"] < 1e-10 and r["sol_rel"] < 1e-12
    assert r["bnorm_rel"] < 1e-10          # (the C oracle sums 17 M squares per-thread serially: its own rounding)


def test_peer_transport_is_bit_identical_to_nccl_on_two_ranks():
    """A two-term sum has one rounding whatever the allreduce algorithm, so at 2 ranks NCCL send/recv/allreduce and the
    peer-memory transport must agree to the bit."""
    if ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    args = ("--refine", 6, "--rtol", 1e-10, "--levels", 5, "--march-min-plane", 1, "--no-oracle")
    peer = run_worker(2, *args)
    nccl = run_worker(2, *args, "--comm-peer", 0)
    assert peer["history"] == nccl["history"]
    assert peer["sol_sha1"] == nccl["sol_sha1"]


# ---- pattern.c on y-slabs (BASELINE config 5 names 8 GPUs; c/ch5/pattern.c:79-84 periodic DMDA, ring neighbours) ----
def run_pattern_worker(nproc, argv, *extra, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "mgpu_pattern_worker.py"),
           "--argv", argv, *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    for line in p.stdout.splitlines():
        if line.startswith("MGPU_RESULT "):
            return json.loads(line[len("MGPU_RESULT "):])
    raise AssertionError("worker failed:\n" + p.stdout[-3000:] + "\n" + p.stderr[-3000:])


@pytest.mark.parametrize("nproc", [2, 4, 8])
@pytest.mark.parametrize("argv", [
    "-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_monitor -ts_max_time 40 -pc_type mg",                   # ARKIMEX, adaptive (64^2)
    "-da_grid_x 4 -da_grid_y 4 -da_refine 5 -ts_type beuler -ts_dt 5 -ts_max_time 15 -ts_monitor -pc_type mg "
    "-snes_converged_reason -ksp_converged_reason",                                                       # Newton + GMRES + MG (128^2)
    "-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_type cn -ts_dt 4 -ts_max_time 12 -ts_monitor -pc_type none",
    "-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_type bdf -ts_max_time 20 -ts_monitor -pc_type mg",
    "-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_monitor -ts_max_time 20 -pc_type mg -ptn_noisy_init 0.15",
])
def test_pattern_on_slabs_equals_the_single_gpu_run(nproc, argv):
    """Same printed lines (adaptive step sequence, Newton and Krylov counts), final state within 1e-12 relative of the
    one-GPU run: per node the slab kernels do the single-GPU arithmetic; only the all-reduced dot products round
    differently."""
    if ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = run_pattern_worker(nproc, argv)
    assert r["world"] == nproc and r["same_steps"], r["lines"]
    # printed lines: identical, except that a GMRES count may move by one where the residual sits on the tolerance (the
    # all-reduced dot products round differently): the north star's "same iteration count +-1"
    import re
    krylov_moved = False
    assert len(r["lines"]) == len(r["one_gpu_lines"])
    for a, b in zip(r["lines"], r["one_gpu_lines"]):
        if a == b:
            continue
        ma, mb = re.fullmatch(r"(.*iterations )(\d+)", a), re.fullmatch(r"(.*iterations )(\d+)", b)
        assert ma and mb and ma.group(1) == mb.group(1) and "Linear solve" in a, (a, b)
        assert abs(int(ma.group(2)) - int(mb.group(2))) <= 1, (a, b)
        krylov_moved = True
    # same Krylov counts: the states agree to rounding; a moved count leaves the difference of two inexact solves
    assert r["rel_diff"] < (1e-7 if krylov_moved else 1e-12), r["rel_diff"]


# ---- the unchanged fish.c on N GPUs from ONE process: -p4b_gpus N (one host thread per GPU inside the shim) ----
@pytest.mark.parametrize("ngpu", [2, 4, 8])
@pytest.mark.parametrize("argv", [
    "-fsh_dim 3 -da_refine 5 -pc_type mg -mg_levels_pc_type jacobi -ksp_rtol 1e-10 -ksp_converged_reason -ksp_monitor",
    "-fsh_dim 3 -da_refine 6 -pc_mg_levels 5 -pc_type mg -mg_levels_pc_type jacobi -ksp_rtol 1e-10 -ksp_converged_reason "
    "-snes_monitor_short",
    "-fsh_dim 2 -da_refine 8 -pc_type mg -mg_levels_pc_type jacobi -ksp_rtol 1e-8 -ksp_converged_reason",
])
def test_unchanged_fish_c_on_n_gpus_from_one_process(ngpu, argv):
    """c/testit.sh:22 runs the reference's binary under `mpiexec -n P`; here the same binary takes -p4b_gpus P: the
    callbacks see one logical rank, the KSP solve runs on P slabs (peer memory between host threads of one process).
    Same report as the one-GPU run: iteration counts equal, monitored norms to rounding, error norms to the digits
    printed."""
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    exe = os.path.join(ROOT, "p4pdes_b200", "bin", "fish")
    if not os.path.exists(exe):
        pytest.skip("unchanged driver not built")

    def run(extra):
        p = subprocess.run([exe] + (argv + extra).split(), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr
        return [l for l in p.stdout.splitlines() if not l.startswith("NCCL version")]

    one, many = run(""), run(" -p4b_gpus %d" % ngpu)
    assert len(one) == len(many)
    for a, b in zip(one, many):
        if "KSP Residual norm" in a:
            fa, fb = float(a.split()[-1]), float(b.split()[-1])
            assert a.split()[:4] == b.split()[:4] and abs(fa - fb) <= 1e-10 * max(fa, 1e-300) + 1e-16 * float(one[0].split()[-1] if "KSP" in one[0] else 1.0)
        else:
            assert a == b, (a, b)


@pytest.mark.parametrize("ngpu", [2, 8])
@pytest.mark.parametrize("argv", [
    "-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_monitor -ts_max_time 40 -pc_type mg -mg_levels_pc_type jacobi",
    "-da_grid_x 4 -da_grid_y 4 -da_refine 5 -ts_type beuler -ts_dt 5 -ts_max_time 15 -ts_monitor -pc_type mg "
    "-mg_levels_pc_type jacobi -snes_converged_reason -ptn_noisy_init 0.1",
])
def test_unchanged_pattern_c_on_n_gpus_from_one_process(ngpu, argv):
    """BASELINE config 5 ("pattern.c ... 8 B200") through the reference's own driver: ./pattern ... -p4b_gpus N runs the
    time stepping on y-slabs, one host thread per GPU; stdout equals the one-GPU run's."""
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    exe = os.path.join(ROOT, "p4pdes_b200", "bin", "pattern")
    if not os.path.exists(exe):
        pytest.skip("unchanged driver not built")

    def run(extra):
        p = subprocess.run([exe] + (argv + extra).split(), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr
        return [l for l in p.stdout.splitlines() if not l.startswith("NCCL version")]

    assert run("") == run(" -p4b_gpus %d" % ngpu)
