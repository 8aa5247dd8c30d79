"""Worker for tests/test_gpu_multi.py (torchrun, one rank per GPU, NCCL): BASELINE.json configs[4] -- K start
states taken mid-episode on rank 0, broadcast, fanned out to R noisy rollouts each over all ranks, per-start-state
statistics all-reduced.  Because Philox streams are keyed by the GLOBAL env id, the all-reduced result must equal,
bit for bit, what one process computes with all the envs -- every rank checks that locally."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bc_gym_planning_env_b200 import parallel  # noqa: E402
from bc_gym_planning_env_b200.envs.base.params import EnvParams  # noqa: E402
from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool  # noqa: E402
from bc_gym_planning_env_b200.vec_env import VecPlanEnv, VecState  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    rank, ws = parallel.world()
    K, R, H = 16, 64 * ws, 48                        # R rollouts per start state over the whole job
    params = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
    costmaps, paths = random_aisle_pool(K, 4242, params)          # same pool on every rank (same seeds)

    # start states: rank 0 drives K envs for a while and snapshots them; the other ranks receive the columns
    src = VecPlanEnv(costmaps, paths, params, noise_parameters=None, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(7)
    low, high = src.action_bounds()
    lo, hi = torch.from_numpy(low).to(device), torch.from_numpy(high).to(device)
    plan = (lo + (hi - lo) * torch.rand((H, K, 2), generator=gen, device=device)).contiguous()   # same on all ranks
    if rank == 0:
        for t in range(25):
            src.step(plan[t % H])
        snap = src.get_state()
        f, i = snap.f, snap.i
    else:
        f = torch.zeros((src.layout.n_frows, K), dtype=torch.float64, device=device)
        i = torch.zeros((src.layout.n_irows, K), dtype=torch.int32, device=device)
    f, i = parallel.broadcast_snapshot(f, i, src=0)
    start = VecState(f, i)

    # this rank's share of the K * R rollouts: global env g = rank * n_local + e starts from state g % K
    n_total = K * R
    n_local = n_total // ws
    ids = ((np.arange(n_local) + rank * n_local) % K)
    fan = VecPlanEnv(costmaps, paths, params, n_envs=n_local, map_ids=ids, path_ids=ids, seed=99, device=device,
                     env_id_base=rank * n_local)
    out = parallel.monte_carlo_rollouts(fan, start, plan)          # all-reduced over ranks
    assert out.shape == (K, 4) and bool((out[:, 3] == R).all())

    # the same job in one process: all n_total envs here, no collective
    ids_all = np.arange(n_total) % K
    one = VecPlanEnv(costmaps, paths, params, n_envs=n_total, map_ids=ids_all, path_ids=ids_all, seed=99, device=device)
    cols = (torch.arange(n_total, device=device) % K)
    one.set_state(VecState(start.f[:, cols].contiguous(), start.i[:, cols].contiguous()))
    ret = torch.zeros(n_total, dtype=torch.float64, device=device)
    for h in range(H):
        _, r, _, _ = one.step(plan[h][cols].contiguous())
        ret += r
    want = torch.zeros((K, 4), dtype=torch.float64, device=device)
    want[:, 0].index_add_(0, cols, ret)
    want[:, 1].index_add_(0, cols, one.state_i[2].to(torch.float64))
    n_path = torch.as_tensor([len(one.full_path(e)) for e in range(n_total)], device=device)
    want[:, 2].index_add_(0, cols, (one.state_i[1].to(torch.int64) > n_path - 1).to(torch.float64))
    want[:, 3].index_add_(0, cols, torch.ones_like(ret))
    assert torch.equal(out, want), (out - want).abs().max()
    assert float(out[:, 0].sum()) > 0 and float(out[:, 1].sum()) > 0      # progress was made, some rollouts crashed
    # the episode-statistics all-reduce sums over ranks
    stats = parallel.allreduce_episode_stats(fan)
    assert stats.shape[0] == 8
    dist.barrier()
    dist.destroy_process_group()
    print("nccl monte carlo worker %d/%d ok" % (rank, ws))


if __name__ == "__main__":
    main()
