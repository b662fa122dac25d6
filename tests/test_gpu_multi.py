"""GPU (needs >= 2 devices, else skipped): domain decomposition over NCCL vs the one-rank reference fixtures
(tests/mgpu_check.py, launched with torchrun: one rank per GPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_decomposition_matches_fixtures():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if " grid " in l]
    print("\n".join(lines))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert len(lines) >= 8 and not any("FAIL" in l for l in lines)


def test_two_rank_halo_overlap_matches_fixtures():
    """the same check with B200_OVERLAP=1: interior tiles of the single-phase decks run while the NCCL halo is in flight"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, B200_OVERLAP="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "tests", "mgpu_check.py"), "dam3d", "dam2d", "heat3d", "heat2d_rhosum"],
                       capture_output=True, text=True, timeout=900, env=env)
    lines = [l for l in p.stdout.splitlines() if " grid " in l]
    print("\n".join(lines))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert len(lines) >= 4 and not any("FAIL" in l for l in lines)


def test_two_rank_run_with_an_empty_brick():
    """one rank owns no atoms: it must keep taking part in the halos and collectives instead of hanging its neighbour"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29535", os.path.join(ROOT, "tests", "mgpu_check.py"), "--empty-rank"], capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if " grid " in l]
    print("\n".join(lines))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert len(lines) == 1 and "OK" in lines[0]


def test_two_rank_balanced_bricks_match_fixtures():
    """non-uniform bricks (parallel.balance_shift = the `balance ... shift` command) on the inhomogeneous dam-break decks"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29536", os.path.join(ROOT, "tests", "mgpu_check.py"), "--balance", "x", "--grid", "2x1x1", "dam3d", "dam2d"],
                       capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if " grid " in l]
    print("\n".join(lines))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert len(lines) == 2 and not any("FAIL" in l for l in lines)


def test_two_rank_local_order_is_the_reference_order():
    """after migration, Atom::sort and fix phase_change insertions every rank holds its atoms in LAMMPS' own local order: exchange() packs
    the leavers while filling each hole with the rank's last atom (comm_brick.cpp:628-650), arrivals are appended, sort() re-numbers
    (atom.cpp:1555), tag_extend numbers new atoms rank after rank -- checked against the P-rank oracle, order and fields"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29537", os.path.join(ROOT, "tests", "mgpu_check.py"), "--vs-world", "dam3d", "dam2d_1000", "droplet2d", "bubble2d_1000", "shock3d"],
                       capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if " grid " in l]
    print("\n".join(lines))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert len(lines) == 5 and not any("FAIL" in l or "LOCAL ORDER DIFFERS" in l for l in lines)
