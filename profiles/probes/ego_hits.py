"""Mean number of non-zero pixels per egocentric crop on the bench workload (run on the GPU box)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bc_gym_planning_env_b200.envs.base.params import EnvParams  # noqa: E402
from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool  # noqa: E402
from bc_gym_planning_env_b200.vec_env import VecPlanEnv  # noqa: E402
n = 16384
params = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
costmaps, paths = random_aisle_pool(256, 9000, params)
env = VecPlanEnv(costmaps, paths, params, n_envs=n, seed=5, auto_reset=True, with_ego=True)
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
low, high = env.action_bounds()
lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
for t in range(300):
    env.step((lo + (hi - lo) * torch.rand((n, 2), generator=gen, device="cuda")).contiguous())
    if t % 100 == 99:
        nz = (env.ego_image != 0).flatten(1).sum(1).double()
        print("step %d: non-zero pixels per crop mean %.1f max %d; crops with none %.3f" % (t, nz.mean().item(), int(nz.max().item()), (nz == 0).double().mean().item()))
