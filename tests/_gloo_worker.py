"""Worker for tests/test_parallel_gloo.py: exercises bc_gym_planning_env_b200.parallel on the gloo
backend (world_size 2, CPU tensors) -- the same code path the NCCL ranks run on the GPU box."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_gym_planning_env_b200 import parallel  # noqa: E402


class FakeEnv(object):
    """Just enough of VecPlanEnv for allreduce_episode_stats."""

    def __init__(self, rank):
        self._stats = torch.arange(8, dtype=torch.float64) * (rank + 1)

    def episode_stats(self, reset=False):
        out = self._stats.clone()
        if reset:
            self._stats.zero_()
        return out


def main():
    dist.init_process_group("gloo")
    rank, ws = parallel.world()
    assert ws == 2

    # shards tile the env index space with no overlap
    lo, hi = parallel.shard_range(65537)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    assert sum(int(s) for s in sizes) == 65537 and abs(int(sizes[0]) - int(sizes[1])) <= 1
    assert (lo == 0) == (rank == 0)

    # statistics all-reduce: sum over ranks, optional reset of the local accumulators
    env = FakeEnv(rank)
    total = parallel.allreduce_episode_stats(env, reset=True)
    assert torch.equal(total, torch.arange(8, dtype=torch.float64) * 3)
    assert float(env._stats.abs().sum()) == 0.0

    # snapshot fan-out: rank 0's columns reach every rank bit-exactly
    f = torch.full((34, 5), float(rank + 1), dtype=torch.float64)
    i = torch.full((6, 5), rank + 7, dtype=torch.int32)
    if rank == 0:
        f = torch.arange(34 * 5, dtype=torch.float64).reshape(34, 5) * 0.1
        i = torch.arange(6 * 5, dtype=torch.int32).reshape(6, 5)
    f, i = parallel.broadcast_snapshot(f, i, src=0)
    assert torch.equal(f, torch.arange(34 * 5, dtype=torch.float64).reshape(34, 5) * 0.1)
    assert torch.equal(i, torch.arange(6 * 5, dtype=torch.int32).reshape(6, 5))

    # fan-out assignment: global env g starts from state g % k on every rank
    cols = parallel.fan_out_columns(3, 4)
    assert cols.tolist() == [(rank * 4 + j) % 3 for j in range(4)]

    dist.barrier()
    dist.destroy_process_group()
    print("gloo worker %d ok" % rank)


if __name__ == "__main__":
    main()
