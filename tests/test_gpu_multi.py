"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): the multi-rank paths on NCCL -- snapshot broadcast,
Monte-Carlo fan-out with per-start-state all-reduce (BASELINE.json configs[4]) -- give bit-identical results to
the same job run in one process."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_monte_carlo_fan_out_over_nccl_equals_one_process():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n, "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_nccl_monte_carlo_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    for r in range(n):
        assert "nccl monte carlo worker %d/%d ok" % (r, n) in res.stdout
