"""A/B of two builds or modes of the step (BCG_B200_LIB=<variant>; in round 2 also BCG_STEP_KERNELS=split for round 1's three
state kernels and BCG_EGO_KERNEL=cta / warp -- modes that no longer exist), and of plain launches against the captured
CUDA graph (VecPlanEnv.step_graph).

    python profiles/probes/fused_vs_split.py run OUT.npz [--envs N] [--steps K]
    python profiles/probes/fused_vs_split.py compare A.npz B.npz                    # bit equality of the two runs
    python profiles/probes/fused_vs_split.py times [--sizes 256,8192,65536]         # per-kernel and per-step times

`run` steps the bench workload (aisle pool, delays 2/1/1, Philox noise, auto-reset, egocentric observation) and saves the
final state rows, the summed rewards, the episode statistics and a digest of every egocentric image.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def make_env(n, pool=512, seed=1234, noise=True, auto_reset=True):
    import torch
    from bc_gym_planning_env_b200.envs.base.params import EnvParams
    from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
    from bc_gym_planning_env_b200.vec_env import DEFAULT_NOISE, VecPlanEnv
    params = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
    costmaps, paths = random_aisle_pool(min(pool, n), 777, params)
    env = VecPlanEnv(costmaps, paths, params, n_envs=n, seed=seed, auto_reset=auto_reset, private_map_copies=True,
                     with_ego=True, noise_parameters=DEFAULT_NOISE if noise else None)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99)
    low, high = env.action_bounds()
    lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
    actions = [(lo + (hi - lo) * torch.rand((n, 2), generator=gen, device="cuda")).contiguous() for _ in range(16)]
    return env, actions


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def run():
    import torch
    out = sys.argv[2]
    n, steps = arg("--envs", 16384), arg("--steps", 400)
    env, actions = make_env(n)
    rsum = torch.zeros(n, dtype=torch.float64, device="cuda")
    digest = torch.zeros(n, dtype=torch.int64, device="cuda")
    weights = (torch.arange(env.ego_image[0].numel(), device="cuda", dtype=torch.int64) % 8191) + 1
    for k in range(steps):
        _, r, d, _ = env.step(actions[k % len(actions)])
        rsum += r
        digest = digest * 31 + (env.ego_image.view(n, -1).to(torch.int64) * weights).sum(dim=1) + d.to(torch.int64)
    env.check_status()
    np.savez(out, state_f=env.state_f.cpu().numpy(), state_i=env.state_i.cpu().numpy(), rsum=rsum.cpu().numpy(),
             digest=digest.cpu().numpy(), stats=env.episode_stats().cpu().numpy(), goal=env.goal_n_state.cpu().numpy(),
             obs=env.obs_vec.cpu().numpy())
    print("saved", out, "mode", os.environ.get("BCG_STEP_KERNELS", "fused"), "episodes", float(env.episode_stats()[0]))


def compare():
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
    print("compared", a.files, "-> differing:", bad)
    if bad:
        for k in bad:
            print(k, "max abs diff", np.nanmax(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))))
        sys.exit(1)


def times():
    import torch
    sizes = [int(x) for x in arg("--sizes", "256,8192,65536").split(",")]
    res = {}
    for n in sizes:
        env, actions = make_env(n)
        for k in range(10):
            env.step(actions[k % 16])
        K = 100
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
        for evs in ev:
            for e in evs:
                e.record()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(K):
            env.step_timed(actions[k % 16], ev[k])
        t1.record()
        torch.cuda.synchronize()
        r = {"step_ms_timed": t0.elapsed_time(t1) / K,
             "kin": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
             "cr": float(np.mean([e[1].elapsed_time(e[2]) for e in ev])),
             "state_or_commit": float(np.mean([e[2].elapsed_time(e[3]) for e in ev])),
             "ego": float(np.mean([e[3].elapsed_time(e[4]) for e in ev]))}
        t0.record()
        for k in range(K):
            env.step(actions[k % 16])
        t1.record()
        torch.cuda.synchronize()
        r["step_ms_plain"] = t0.elapsed_time(t1) / K
        if True:
            env.step_graph(actions[0])
            torch.cuda.synchronize()
            t0.record()
            for k in range(K):
                env.actions.copy_(actions[k % 16])
                env.step_graph()
            t1.record()
            torch.cuda.synchronize()
            r["step_ms_graph"] = t0.elapsed_time(t1) / K
            t0.record()
            for k in range(K):
                env.step_graph()
            t1.record()
            torch.cuda.synchronize()
            r["step_ms_graph_same_actions"] = t0.elapsed_time(t1) / K
        env.check_status()
        res[n] = r
        print(n, json.dumps(r), flush=True)
        del env, actions
        torch.cuda.empty_cache()
    print(json.dumps({"mode": os.environ.get("BCG_STEP_KERNELS", "fused"), "times": res}))


def profile():
    """a short run for ncu: python profiles/probes/fused_vs_split.py profile [--envs N] [--steps K]"""
    import torch
    n, steps = arg("--envs", 65536), arg("--steps", 12)
    env, actions = make_env(n)
    for k in range(steps):
        env.step(actions[k % 16])
    torch.cuda.synchronize()
    env.check_status()
    print("profiled run done")


if __name__ == "__main__":
    {"run": run, "compare": compare, "times": times, "profile": profile}[sys.argv[1]]()
