"""N>1 device path: one process per GPU (torch.distributed.run), NCCL ghost import inside
b200s_evaluate, owner-computes rows against the oracle.  Skipped on boxes with a single GPU (the
host-side lists are covered on CPU by tests/test_partition_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(world, args):
    """torchrun on a free rendezvous port; a port that got taken between the probe and the bind is retried."""
    for attempt in range(3):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
               os.path.join(ROOT, "tests", "helpers", "mgpu_worker.py")] + args
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            break
    return r


@pytest.mark.parametrize("world,n", [(2, (12, 7, 5)), (4, (10, 9, 6)), (8, (10, 8, 8))])
def test_nccl_halo_rows_match_oracle(world, n):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _launch(world, [",".join(map(str, n))])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MGPU_OK") == world, r.stdout[-3000:]


@pytest.mark.parametrize("world,n", [(2, (12, 7, 5)), (4, (10, 9, 6))])
def test_nccl_actions_around_the_path(world, n):
    """multiply (x through the halo), energy, mass action with imported acceleration, Dirichlet + CG with
    all-reduced dot products on a partitioned mesh."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _launch(world, [",".join(map(str, n)), "post"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MGPU_OK") == world and r.stdout.count("MGPU_POST") == world, r.stdout[-3000:]


@pytest.mark.parametrize("world,n", [(2, (12, 7, 5)), (4, (10, 9, 6))])
def test_nccl_row_export_matches_owner_computes(world, n):
    """Row-element mode (FE_MATRIX): row elements only, export of the off-rank rows over NCCL inside b200s_evaluate."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _launch(world, [",".join(map(str, n)), "export"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MGPU_EXPORT_OK") == world, r.stdout[-3000:]
