"""Slab-partitioned multi-GPU path: needs >= 2 CUDA devices (skipped on the single-GPU test box).
Runs tests/mgpu_check.py under torchrun: assembly, SpMV and a Newton solve over 2 slabs must reproduce
the single-domain CPU restatement (F, J, J x to 1e-12; converged fields to 1e-8).  Two gather thresholds of the
multi-rank multigrid: the default (this small grid is gathered whole, so the hierarchy is the single-domain one
and the Krylov counts match the CPU run) and 300 cells (slab-local levels above an all-gathered coarse level); each with the exchanges through the
peer-memory mailboxes (csrc/tpb_comm.cu) and through NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# TPB_P2P: 15 = every exchange through the peer-memory mailboxes with the halo fused into the SpMV kernel,
# 7 = mailboxes with separate push/pull kernels, 0 = NCCL only
@pytest.mark.parametrize("gather,p2p", [(None, "15"), ("300", "15"), (None, "7"), ("300", "0")])
def test_two_slabs_reproduce_single_domain(gather, p2p):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.pop("TPB_MG_GATHER", None)
    if gather:
        env["TPB_MG_GATHER"] = gather
    env["TPB_P2P"] = p2p
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "mgpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("| OK") == 4      # two checks (kernels + Newton, model.solve()) on two ranks


def test_two_slabs_on_the_bench_workload_reproduce_single_domain():
    """tests/mgpu_bench_check.py: the 2x stacked 60x220x85 bench workload on two slabs against the same grid as ONE domain
    (fields 1e-8, no runaway Newton step)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_bench_check.py"), "3"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("| OK") == 2
