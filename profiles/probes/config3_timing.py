"""Step time of BASELINE.json configs[3] (corridor stand-in with filled obstacle regions, 16 384 envs): the windows hold
thousands of occupied cells, so the sparse egocentric kernel hands (almost) every env to the dense cell-tile kernel.
    python profiles/probes/config3_timing.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bc_gym_planning_env_b200.envs.base.params import EnvParams  # noqa: E402
from bc_gym_planning_env_b200.envs.rw_corridors.tdwa_test_environments import \
    get_random_maps_squeeze_between_obstacle_in_corridor_on_path  # noqa: E402
from bc_gym_planning_env_b200.vec_env import VecPlanEnv  # noqa: E402

n = 16384
original, path, variants = get_random_maps_squeeze_between_obstacle_in_corridor_on_path(n_variants=256, seed=1)
ep = EnvParams(iteration_timeout=1200, pose_delay=1, control_delay=0, state_delay=1)
for sparse in (True, False):
    env = VecPlanEnv(list(variants), [path], ep, n_envs=n, map_ids=np.arange(n) % len(variants), path_ids=np.zeros(n, dtype=np.int64),
                     auto_reset=True, with_ego=True, private_map_copies=True, ego_sparse=sparse)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    low, high = env.action_bounds()
    lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
    acts = [(lo + (hi - lo) * torch.rand((n, 2), generator=gen, device="cuda")).contiguous() for _ in range(8)]
    for t in range(20):
        env.step(acts[t % 8])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(200):
        env.step(acts[t % 8])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 200
    handed = int(env._ego_list[n]) if env._ego_list is not None else n
    print("ego_sparse=%s: %.4f ms/step, %.3g env-steps/s, %d of %d envs rendered by the dense kernel in the last step"
          % (sparse, ms, n / ms * 1e3, handed, n))
    env.check_status()
    del env
    torch.cuda.empty_cache()
