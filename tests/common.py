"""Shared helpers for the parity tests: fixture loading, oracle and VecPlanEnv construction."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# The contract of BASELINE.json's north_star: poses/rewards within 1e-5 relative, flags bit-exact.
# The CUDA path computes in fp64, so the tests also assert the much tighter bound it actually meets.
CONTRACT_RTOL = 1e-5
TIGHT_ATOL = 1e-9


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "params" in d:
        d["params"] = json.loads(str(d["params"]))
    return d


def fixture_envs(d):
    n = int(d["n_envs"])
    return [(d["costmap_%d" % i], d["origin_%d" % i], d["path_%d" % i]) for i in range(n)]


def make_oracles(d, alphas=None, normal_sources=None):
    from oracle import plan_env_oracle as O
    p = d["params"]
    out = []
    for i, (cm, origin, path) in enumerate(fixture_envs(d)):
        out.append(O.OraclePlanEnv(cm, origin, float(d["resolution"]), path, robot=p["robot"], dt=p["dt"], sp=p["sp"],
                                   ap=p["ap"], multiplier=p["multiplier"], timeout=p["timeout"], delays=p["delays"],
                                   alphas=alphas, normal_source=None if normal_sources is None else normal_sources[i],
                                   refine=False, reward_provider=p.get("reward_provider", "continuous_reward")))
    return out


def env_params(p, **over):
    from bc_gym_planning_env_b200.envs.base.params import EnvParams, RewardParams
    kw = dict(dt=p["dt"], goal_spat_dist=p["sp"], goal_ang_dist=p["ap"], iteration_timeout=p["timeout"],
              control_delay=p["delays"][0], pose_delay=p["delays"][1], state_delay=p["delays"][2], robot_name=p["robot"],
              refine_path=False, reward_provider_name=p.get("reward_provider", "continuous_reward"),
              reward_provider_params=RewardParams(spatial_precision=p["sp"], angular_precision=p["ap"],
                                                  spatial_progress_multiplier=p["multiplier"]))
    kw.update(over)
    return EnvParams(**kw)


def make_vec_env(d, **kw):
    from bc_gym_planning_env_b200.utilities.costmap_2d import CostMap2D
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    envs = fixture_envs(d)
    res = float(d["resolution"])
    costmaps = [CostMap2D(cm, res, np.array(origin, dtype=np.float64)) for cm, origin, _ in envs]
    paths = [path for _, _, path in envs]
    params = kw.pop("params", None)
    if params is None:
        params = env_params(d["params"]) if "params" in d else None
    kw.setdefault("noise_parameters", None)
    return VecPlanEnv(costmaps, paths, params, **kw)
