"""CPU: the multi-rank host logic (sharding, statistics all-reduce, snapshot fan-out) on gloo,
world_size 2, rendezvous on 127.0.0.1."""
import os
import socket
import subprocess
import sys

from bc_gym_planning_env_b200 import parallel

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_range_partitions():
    for n, ws in ((65536, 8), (10, 3), (5, 8)):
        ranges = [parallel.shard_range(n, r, ws) for r in range(ws)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_single_process_is_a_no_op_world():
    import torch
    assert parallel.world() == (0, 1)
    t = torch.ones(3, dtype=torch.float64)
    assert parallel.allreduce_sum_(t) is t and float(t.sum()) == 3.0


def test_two_rank_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "gloo worker 0 ok" in res.stdout and "gloo worker 1 ok" in res.stdout
