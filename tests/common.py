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
                                   refine=False, reward_provider=p.get("reward_provider", "continuous_reward"),
                                   footprint_scale=p.get("footprint_scale", 1.0)))
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
    if "params" in d:
        kw.setdefault("footprint_scale", d["params"].get("footprint_scale", 1.0))
    return VecPlanEnv(costmaps, paths, params, **kw)


def tiny_worlds():
    """Edge-case worlds (costmap uint8, origin, coarse path): maps far smaller than an aisle, which the robot leaves;
    one is smaller than the tricycle's footprint, one is all lethal and driven into from outside.  Shared by
    oracle/gen_golden.py (fixture edge_worlds: the reference stepped on them) and the GPU edge-case tests."""
    worlds = []
    m = np.zeros((50, 40), dtype=np.uint8)             # 1.5 m x 1.2 m: a lethal post in a corner the footprint misses,
    m[0:3, 0:3] = 254                                  # non-lethal costs under the robot
    m[20:24, 15:19] = 253
    m[30, :] = 255
    worlds.append((m, np.array([-0.3, -0.6]), np.array([[0., 0., 0.], [4., 0.5, 0.2]])))
    m = np.zeros((20, 20), dtype=np.uint8)             # 0.6 m square, smaller than the footprint that passes over it
    m[0, :] = 253
    m[5:9, 5:9] = 100
    m[:, 19] = 255
    worlds.append((m, np.array([0.4, -0.3]), np.array([[-1., 0., 0.], [3., 0., 0.]])))
    m = np.full((33, 47), 254, dtype=np.uint8)         # all lethal, the robot starts outside and drives in
    worlds.append((m, np.array([1.5, -0.5]), np.array([[0., 0., 0.], [4., 0., 0.]])))
    m = np.zeros((64, 64), dtype=np.uint8)             # empty map, path leaves through a corner
    worlds.append((m, np.array([-1., -1.]), np.array([[0., 0., np.pi / 4], [3., 3., np.pi / 4]])))
    return worlds
