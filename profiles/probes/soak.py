"""Soak run (GPU box): two batches fed the same actions and noise for thousands of steps, one rendering the egocentric
observation with the sparse scatter kernel, the other with the dense gather kernel; states must stay bit-identical and
every compared image equal.  Also steps a device-generated batch with reset storms and watches the status counters.
    python profiles/probes/soak.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bc_gym_planning_env_b200.envs.base.params import EnvParams  # noqa: E402
from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool  # noqa: E402
from bc_gym_planning_env_b200.vec_aisle_env import VecRandomAisleTurnEnv, VecRandomMiniEnv  # noqa: E402
from bc_gym_planning_env_b200.vec_env import VecPlanEnv  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n = 32768
params = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
costmaps, paths = random_aisle_pool(512, 9000, params)
a = VecPlanEnv(costmaps, paths, params, n_envs=n, seed=5, auto_reset=True, with_ego=True)
b = VecPlanEnv(costmaps, paths, params, n_envs=n, seed=5, auto_reset=True, with_ego=True, ego_sparse=False)
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
low, high = a.action_bounds()
lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
t0 = time.time()
compared = handed = 0
for t in range(steps):
    act = (lo + (hi - lo) * torch.rand((n, 2), generator=gen, device="cuda")).contiguous()
    a.step(act)
    b.step(act)
    if t % 20 == 19:
        assert torch.equal(a.state_f, b.state_f) and torch.equal(a.state_i, b.state_i), t
        assert torch.equal(a.ego_image, b.ego_image), t
        assert torch.equal(a.goal_n_state, b.goal_n_state), t
        handed += int(a._ego_list[n])
        compared += 1
a.check_status()
b.check_status()
print("sparse == dense over %d steps x %d envs (%d image comparisons, %d envs handed to the dense kernel, episodes %d) in %.1f s"
      % (steps, n, compared, handed, int(a.episode_stats()[0]), time.time() - t0))

g = VecRandomAisleTurnEnv(8192, params, seed=3, auto_reset=True, with_ego=True)
m = VecRandomMiniEnv(4096, seed=4, auto_reset=True, with_ego=True)
for env in (g, m):
    k = env.n_envs
    resets = 0
    lo2, hi2 = env.action_bounds()
    lo2, hi2 = torch.from_numpy(lo2).cuda(), torch.from_numpy(hi2).cuda()
    for t in range(steps // 3):
        act = (lo2 + (hi2 - lo2) * torch.rand((k, 2), generator=gen, device="cuda")).contiguous()
        _, _, done, _ = env.step(act)
        if t % 10 == 9 and bool(done.any()):
            resets += int(done.sum())
            env.reset(done.clone())
    env.check_status()
    print("%s: %d steps, %d worlds regenerated on the device, status clean" % (type(env).__name__, steps // 3, resets))
