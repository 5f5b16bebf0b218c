"""The N>1 path: one process per GPU launched by torch.distributed.run.  world_size-2 (and 4) gloo runs on
CPU cover the host logic; the NCCL run needs two GPUs (`gpurun --gpus 2`) and is skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(mode, world, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mp_worker.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        assert "WORKER_OK %d" % rank in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("world", [2, 4])
def test_host_logic_under_gloo(world):
    _launch("gloo", world)


@pytest.mark.gpu
def test_sharding_pooling_and_diagnostics_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _launch("nccl", 2)
