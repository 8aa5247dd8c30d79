"""Overlap probe (GPU box): how much of the latency-bound kernels (kin / collide+reward / commit) hides under the
HBM-bound egocentric kernel when the batch is split into independent sub-batches stepped on separate streams.
Envs are independent, so C sub-batches of N / C envs compute exactly what one batch of N computes.
    python profiles/probes/stream_overlap.py [envs] [steps]
Prints ms per step of all N envs for C = 1, 2, 4, 8 in two modes: `free` (streams never join: upper bound, what an
open-loop action source gets) and `join` (every step ends with all streams joined, like one VecPlanEnv.step call)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bc_gym_planning_env_b200.envs.base.params import EnvParams  # noqa: E402
from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool  # noqa: E402
from bc_gym_planning_env_b200.vec_env import VecPlanEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
params = EnvParams(control_delay=2, pose_delay=1, state_delay=1)
costmaps, paths = random_aisle_pool(256, 9000, params)
gen = torch.Generator(device="cuda")
gen.manual_seed(1)

for C in (1, 2, 4, 8):
    k = n // C
    envs = [VecPlanEnv(costmaps, paths, params, n_envs=k, seed=5, auto_reset=True, with_ego=True, private_map_copies=True,
                       env_id_base=c * k) for c in range(C)]
    low, high = envs[0].action_bounds()
    lo, hi = torch.from_numpy(low).cuda(), torch.from_numpy(high).cuda()
    acts = [[(lo + (hi - lo) * torch.rand((k, 2), generator=gen, device="cuda")).contiguous() for _ in range(4)] for _ in range(C)]
    streams = [torch.cuda.Stream(priority=-1 if c % 2 == 0 else 0) for c in range(C)]
    main = torch.cuda.current_stream()
    for mode in ("free", "join"):
        for rep in range(2):                               # rep 0 = warm-up
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(main)
            for s in streams:
                s.wait_stream(main)
            for t in range(steps):
                for c in range(C):
                    with torch.cuda.stream(streams[c]):
                        envs[c].step(acts[c][t % 4])
                if mode == "join":
                    for s in streams:
                        main.wait_stream(s)
                    for s in streams:
                        s.wait_stream(main)
            for s in streams:
                main.wait_stream(s)
            t1.record(main)
            torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
        print("C=%d %-4s  %.4f ms/step  %.3e env-steps/s" % (C, mode, ms, n / (ms * 1e-3)), flush=True)
    for e in envs:
        e.check_status()
    del envs, acts
    torch.cuda.empty_cache()
