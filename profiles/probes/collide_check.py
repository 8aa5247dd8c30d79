"""The three collision kernels against one another on random poses near the walls (run on the GPU box):
    python profiles/probes/collide_check.py
thread-per-env kernel (collide_warp_balanced, the code move_kernel inlines), warp-per-env tile kernel and the uint8
kernel must agree for full warps and for batches that end in a partly filled warp (37 and 5 envs)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bc_gym_planning_env_b200.envs.base.params import EnvParams
from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
from bc_gym_planning_env_b200.vec_env import VecPlanEnv
params = EnvParams()
costmaps, paths = random_aisle_pool(32, 5, params)
for n in (4096, 37, 5):
    env = VecPlanEnv(costmaps, paths, params, n_envs=n, seed=1)
    rng = np.random.RandomState(3)
    poses = np.zeros((n, 3))
    for e in range(n):
        p = env.full_path(e)
        k = rng.randint(len(p))
        poses[e] = p[k] + np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.8, 0.8), rng.uniform(-3, 3)])
    pt = torch.from_numpy(poses).cuda()
    a = env.pose_collides(pt).cpu().numpy()
    b = env.pose_collides(pt, use_u8=True).cpu().numpy()
    c = env.pose_collides(pt, count_pixels=True)[0].cpu().numpy()
    bad = np.nonzero(a != b)[0]
    print(n, "hits u8", b.sum(), "thread", a.sum(), "warp", c.sum(), "mismatch", len(bad), "false neg", int((b & ~a).sum()), "false pos", int((a & ~b).sum()))
    print(" bad lanes", np.bincount(bad % 32, minlength=32).tolist())
    print(" first bad", bad[:20].tolist())
