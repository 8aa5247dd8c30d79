#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PlanEnv.step path (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # CPU reference arm (oracle port)

Workload (`config.workload`): AisleTurnEnv, 65 536 envs per GPU, tricycle robot,
EnvParams(control_delay=2, pose_delay=1, state_delay=1), PlanEnv's odometry noise on (Philox4x32-10),
auto-reset on done, egocentric observation assembled every step.  A "step" is one `VecPlanEnv.step`
over all envs of the rank.  Envs shard across ranks with no data-path collective ("weak" scaling:
65 536 envs per GPU); the only collective is one all-reduce of the episode statistics at the end of
the timed region.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
DELAYS = (2, 1, 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--pool", type=int, default=2048, help="distinct random aisle maps generated on the host")
    ap.add_argument("--no-ego", action="store_true", help="skip the egocentric observation kernel (not the headline)")
    ap.add_argument("--ego-staging", default="tiles", choices=["tiles", "tma", "spans"],
                    help="how the egocentric kernel stages its source window (see VecPlanEnv)")
    ap.add_argument("--worlds", default="pool", choices=["pool", "device"],
                    help="pool: --pool host-built maps replicated on device (default); device: every env's own world drawn "
                         "and rasterised on the GPU (bcg_generate_aisles; about 1.2 MB of slots per env)")
    ap.add_argument("--gen-envs", type=int, default=8192, help="envs of the device-generation (reset storm) measurement; 0 = skip")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--e2e-image-steps", type=int, default=5, help="steps of the images-to-host e2e variant; 0 = skip")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-envs", type=int, default=16, help="reference arm: envs advanced per worker per step")
    return ap.parse_args()


def workload_config(args, n_envs):
    return {
        "workload": "AisleTurnEnv x %d envs/GPU, tricycle, control/pose/state delay %d/%d/%d, Philox odometry noise, "
                    "auto-reset, egocentric obs %s" % (n_envs, DELAYS[0], DELAYS[1], DELAYS[2], "off" if args.no_ego else "on"),
        "envs_per_gpu": n_envs,
        "map_pool": args.pool,
        "maps": ("random aisle turns (RandomAisleTurnEnv distribution); pool of %d distinct maps generated on the host, "
                 "replicated on device so every env owns a private costmap copy in HBM" % args.pool) if args.worlds == "pool"
        else "random aisle turns (RandomAisleTurnEnv distribution), one distinct world per env drawn and rasterised on the "
             "GPU (bcg_generate_aisles, Philox)",
        "ego_staging": args.ego_staging,
        "l2": "inputs larger than L2 (per-env costmaps + egocentric output are GBs per step); no explicit flush",
    }


def aisle_params():
    from bc_gym_planning_env_b200.envs.base.params import EnvParams
    return EnvParams(control_delay=DELAYS[0], pose_delay=DELAYS[1], state_delay=DELAYS[2])


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region: the sampler starts before the warm-up
    (nvidia-smi needs a few hundred ms to come up), every sample carries a timestamp, and only the samples between
    `mark_start()` and `mark_end()` are kept (all samples under load if the region was shorter than one poll)."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc, self.path = None, None
        self.t_load = self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.t_load = time.time()

    def wait_first_sample(self, timeout=15.0):
        """Block until nvidia-smi has written its first line (bounded), so that short runs are sampled too."""
        t = time.time()
        ok = False
        while self.proc is not None and time.time() - t < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    ok = True
                    break
            except OSError:
                pass
            time.sleep(0.02)
        self.t_load = time.time()            # the load (warm-up, then the timed region) starts now
        return ok

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()
        self.extended = False

    def extend_end(self):
        """The same load kept running after the timed region (see run_b200): samples up to now count."""
        self.t1 = time.time()
        self.extended = True

    def samples_since_start(self):
        """Lines nvidia-smi has written since mark_start (cheap estimate: the file is only read, never parsed twice)."""
        if self.proc is None or self.t0 is None:
            return 1 << 30
        n = 0
        try:
            for line in open(self.path):
                f = line.split(",")
                try:
                    if len(f) >= 9 and self._stamp(f[0]) >= self.t0 - 0.005:
                        n += 1
                except ValueError:
                    continue
        except OSError:
            return 1 << 30
        return n

    @staticmethod
    def _stamp(text):
        import datetime
        return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    rows.append((self._stamp(f[0]), float(f[2]), float(f[3]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        t0, t1 = self.t0 or self.t_load, self.t1 or time.time()
        inside = [r for r in rows if t0 - 0.005 <= r[0] <= t1 + 0.005]
        scope = "timed region + the same load kept running right after it" if getattr(self, "extended", False) else "timed region"
        if not inside:                       # region shorter than a poll: everything sampled since the warm-up began
            inside, scope = [r for r in rows if self.t_load <= r[0] <= t1 + 0.005], "warm-up + timed region"
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(np.max([r[2] for r in inside])),
                       reasons=sorted(reasons), samples=len(inside), scope=scope)
        return out


# ------------------------------------------------------------------------------------------------
# CPU legs (the ONLY places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def _oracle_envs(n, seed):
    from oracle import plan_env_oracle as O
    from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
    costmaps, paths = random_aisle_pool(n, seed, aisle_params())
    envs = []
    for i, (cm, path) in enumerate(zip(costmaps, paths)):
        src = (lambda env_id: (lambda step: O.philox_normal_source(0, env_id, step)))(seed + i)
        envs.append(O.OraclePlanEnv(cm.get_data(), cm.get_origin(), cm.get_resolution(), path, delays=DELAYS,
                                    alphas=O.DEFAULT_NOISE, normal_source=src))
    return envs


def _oracle_advance(envs, rng, with_ego, low, high):
    """One env.step (+ egocentric observation) for each oracle env, auto-reset on done."""
    from oracle import plan_env_oracle as O
    n = 0
    for env in envs:
        obs, _, done, _ = env.step(rng.uniform(low, high).astype(np.float32))
        if with_ego:
            O.ego_costmap(env.costmap, obs["pose"], env.origin, env.resolution)
            O.goal_n_state(obs["path"], obs["pose"], obs["robot_state"], env.resolution)
        if done:
            env.reset()
        n += 1
    return n


def _action_bounds():
    s = 60. * np.pi / 180.
    return np.array([s / 10, -np.pi / 2]), np.array([s / 2, np.pi / 2])


def cpu_baseline_sample(seconds, with_ego):
    """Single-process oracle port on one host core for ~`seconds` s."""
    _single_threaded_math()
    envs = _oracle_envs(4, 900)
    rng = np.random.RandomState(0)
    low, high = _action_bounds()
    _oracle_advance(envs, rng, with_ego, low, high)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        n += _oracle_advance(envs, rng, with_ego, low, high)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d env-steps of the NumPy oracle port (4 aisle envs, same delays/noise/ego settings) in %.1f s on 1 core; "
                      "host has %d cores" % (n, dt, os.cpu_count())}


def cpu_generation_sample(seconds):
    """Worlds per second of the host path the reference takes at every RandomAisleTurnEnv.reset (draw, geometry,
    cv2-equivalent walls, refine_path, initial reward state), restated by the oracle, on one core."""
    from oracle import aisle_oracle as A
    from oracle import plan_env_oracle as O
    rng = np.random.RandomState(5)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        coarse, costmap, origin = A.aisle_world(A.draw_turn_params(rng), 0.03)
        path = O.refine_path(coarse, 0.05)
        O.initial_reward_state(path, 1.0, np.pi / 2)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "worlds/s", "cores": 1, "kind": "port",
            "sample": "%d aisle worlds of the NumPy oracle port in %.1f s on 1 core" % (n, dt)}


def _single_threaded_math():
    """One worker per core: keep BLAS / OpenCV from spawning their own thread pools in every worker."""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    try:
        import cv2
        cv2.setNumThreads(0)
    except Exception:
        pass


_REF_BARRIER = None


def _ref_init(barrier):
    global _REF_BARRIER
    _REF_BARRIER = barrier


def _ref_worker(args):
    wid, n_envs, n_steps, warmup, with_ego = args
    _single_threaded_math()
    envs = _oracle_envs(n_envs, 1000 + 100 * wid)
    rng = np.random.RandomState(wid)
    low, high = _action_bounds()
    for _ in range(warmup):
        _oracle_advance(envs, rng, with_ego, low, high)
    _REF_BARRIER.wait()                      # all workers start the timed steps together
    t0 = time.perf_counter()
    n = 0
    for _ in range(n_steps):
        n += _oracle_advance(envs, rng, with_ego, low, high)
    return n, t0, time.perf_counter()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (its oracle port; the reference
    itself is pure Python and cannot travel to the GPU box) on all host cores.  A step = every worker
    advances `--ref-envs` envs once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    workers = max(1, os.cpu_count() or 1)
    with_ego = not args.no_ego
    from oracle import plan_env_oracle  # noqa: F401  (import once here, not in every forked worker)
    from bc_gym_planning_env_b200.envs import synth_turn_env  # noqa: F401
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(workers)
    with ctx.Pool(workers, initializer=_ref_init, initargs=(barrier,)) as pool:
        res = pool.map(_ref_worker, [(w, args.ref_envs, args.steps, args.warmup, with_ego) for w in range(workers)], chunksize=1)
    n = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    value = n / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.envs),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": "%d workers x %d envs x %d steps of the NumPy oracle port" % (workers, args.ref_envs, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def build_env(args, rank, device):
    from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    params = aisle_params()
    if args.worlds == "device":
        from bc_gym_planning_env_b200.vec_aisle_env import VecRandomAisleTurnEnv
        return VecRandomAisleTurnEnv(args.envs, params, seed=1234, auto_reset=True, device=device,
                                     env_id_base=rank * args.envs, with_ego=not args.no_ego)
    costmaps, paths = random_aisle_pool(args.pool, 10000 + rank * args.pool, params)
    env = VecPlanEnv(costmaps, paths, params, n_envs=args.envs, seed=1234, auto_reset=True, device=device,
                     env_id_base=rank * args.envs, private_map_copies=True, with_ego=not args.no_ego,
                     ego_staging=args.ego_staging)
    return env


def run_b200(args):
    import torch
    import torch.distributed as dist
    from bc_gym_planning_env_b200 import _native as nat
    from bc_gym_planning_env_b200.parallel import allreduce_episode_stats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nat.require_cuda()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    sampler = ClockSampler(local_rank)
    sampler.start()                           # nvidia-smi needs a few hundred ms to come up: start it before the set-up
    env = build_env(args, rank, device)
    n = env.n_envs
    low, high = env.action_bounds()
    gen = torch.Generator(device=device)
    gen.manual_seed(4321 + rank)
    n_sets = 8
    lo_t, hi_t = torch.from_numpy(low).to(device), torch.from_numpy(high).to(device)
    actions = [(lo_t + (hi_t - lo_t) * torch.rand((n, 2), generator=gen, device=device)).contiguous() for _ in range(n_sets)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler.wait_first_sample()
    for w in range(args.warmup):
        env.step(actions[w % n_sets])
    env.episode_stats(reset=True)
    barrier()

    # ---- timed region: K steps, inputs resident in HBM ------------------------------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    for evs in ev:
        for e in evs:
            e.record()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_start()
    t_start.record()
    for k in range(args.steps):
        env.step_timed(actions[k % n_sets], ev[k])
    stats = allreduce_episode_stats(env)      # the path's only collective
    t_end.record()
    barrier()
    sampler.mark_end()
    elapsed_ms = t_start.elapsed_time(t_end)
    # a timed region shorter than a few polls of nvidia-smi: keep the identical load running (untimed) until the sampler
    # has seen it, so that the clocks line always describes this workload under load
    t_roll = time.time()
    while sampler.samples_since_start() < 5 and time.time() - t_roll < 3.0:
        for k in range(20):
            env.step(actions[k % n_sets])
        torch.cuda.synchronize()
        sampler.extend_end()
    clocks = sampler.stop()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    env.check_status()
    value = n * world * args.steps / (elapsed_ms * 1e-3)
    split = os.environ.get("BCG_STEP_KERNELS") == "split"      # round 1's three state kernels (A/B)
    kin_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    cr_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    commit_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))      # fused build: the state kernel
    ego_ms = float(np.mean([e[3].elapsed_time(e[4]) for e in ev]))

    # ---- same loop without the egocentric kernel (reported beside the headline, not instead of it) -----
    no_ego_value = None
    if not args.no_ego:
        saved = (env._out.ego_image, env._out.goal_n_state)
        env._out.ego_image, env._out.goal_n_state = None, None
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for k in range(args.steps):
            env.step(actions[k % n_sets])
        a1.record()
        barrier()
        tt = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        no_ego_value = n * world * args.steps / (float(tt.item()) * 1e-3)
        env._out.ego_image, env._out.goal_n_state = saved

    # ---- algorithmic bytes -----------------------------------------------------------
    cand_pose = env._cand[:3].t().contiguous()
    _, pixels = env.pose_collides(cand_pose, count_pixels=True)
    coll_bytes = float(pixels.sum().item())                       # 1 B per in-map footprint pixel (uint8 costmap)
    n_path = torch.tensor([len(env.full_path(e)) for e in range(min(n, 4096))], dtype=torch.float64)
    remaining = float(n_path.mean().item()) - float(env.state_i[nat.I_TARGET].double().mean().item())
    scan_bytes = 24.0 * max(remaining, 0.0) * n                  # (N - target_idx) x 24 B per env
    cp = env._c_params
    ego_bytes = 2.0 * cp.ego_w * cp.ego_h * n                     # gather read + image write per env
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # DRAM bytes per launch of each kernel from the committed `ncu --set full` capture of this workload
    # (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum at 65 536 envs); null for other sizes
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("envs_per_gpu", -1)) == n:
            traffic = tj.get("dram_bytes_per_launch", {})

    def roof(kernel, alg_bytes, ms, note):
        ach = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        tr = traffic.get(kernel.split(" ")[0])
        phys = tr / (ms * 1e-3) / 1e9 if (tr and ms > 0) else None      # DRAM bytes ncu counted / launch time measured here
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": tr, "traffic_gbs": phys, "traffic_frac": None if phys is None else phys / peak,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms, "peak_source": peak_src, "note": note}

    roof_commit = roof("collide_reward_kernel", coll_bytes + scan_bytes, cr_ms,
                       "algorithmic bytes (SURVEY 8d) = in-map footprint pixels x 1 B (uint8 definition; the kernel reads "
                       "the derived 1-bit lethal tile plane) + remaining path points x 24 B (the kernel skips path chunks "
                       "farther than the reach radius); collision-only share: %.1f MB" % (coll_bytes / 1e6))
    sparse = getattr(env, "_ego_list", None) is not None and os.environ.get("BCG_EGO_KERNEL") != "dense"
    ego_name = "ego_sparse_kernel" if sparse else ("ego_tiles_kernel" if env.ego_staging == "tiles" else "ego_kernel")
    dense_envs = int(env._ego_list[n].item()) if sparse else n
    # Algorithmic bytes of the egocentric observation.  SURVEY 8d counts a gather: source read + image write = 2 x W x H
    # per env.  The sparse kernel is a scatter and never reads the source bytes: what it must move is the image
    # (W x H written) and one occupancy bit per source cell of the crop (W x H / 8 read); rooflining it against bytes it
    # does not touch would give fractions above 1, so `frac` uses the scatter's own bytes and the 8d figure is kept
    # beside it.
    if sparse:
        scatter_bytes = float(cp.ego_w * cp.ego_h + (cp.ego_w * cp.ego_h + 7) // 8) * n
        roof_ego = roof(ego_name, scatter_bytes, ego_ms,
                        "algorithmic bytes = W x H image write + W x H / 8 occupancy bits read per env (scatter formulation; "
                        "the kernel never reads the source bytes).  `survey_8d` rooflines the same launch against the gather "
                        "definition of SURVEY 8d (2 x W x H per env) and can exceed 1.  ms_per_launch covers ego_sparse_kernel "
                        "plus the dense ego_tiles_kernel launch for the %d envs it handed over in the last step." % dense_envs)
        roof_ego["survey_8d"] = {"algorithmic_bytes_per_launch": ego_bytes, "achieved": ego_bytes / (ego_ms * 1e-3) / 1e9,
                                 "frac": ego_bytes / (ego_ms * 1e-3) / 1e9 / peak}
    else:
        roof_ego = roof(ego_name, ego_bytes, ego_ms, "algorithmic bytes = 2 x ego_w x ego_h per env (SURVEY 8d: source read + "
                        "image write); staging: " + env.ego_staging)
    dominant = roof_ego if (not args.no_ego and ego_ms >= cr_ms) else roof_commit

    # ---- stand-alone collision kernels, cold L2 (flush between launches) ----------------------------
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=device)
    rows = cand_pose.t().contiguous()

    def time_kernel(fn, reps=5):
        ms = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    import ctypes as C
    flags = torch.empty(n, dtype=torch.uint8, device=device)
    s = env._stream()
    # work records of the poses the last step proposed (one prep launch), then only the collision kernel is timed
    nat.check(nat.lib().bcg_collision(C.byref(env._c_params), C.byref(env._batch), None, nat.ptr(flags), None, s))
    tiles_ms = time_kernel(lambda: nat.check(nat.lib().bcg_collision_recheck(C.byref(env._c_params), C.byref(env._batch), nat.ptr(flags), 0, s)))
    u8_ms = time_kernel(lambda: nat.check(nat.lib().bcg_collision_recheck(C.byref(env._c_params), C.byref(env._batch), nat.ptr(flags), 1, s)))
    del flush

    # ---- reset storm: every env of a batch gets a new world on the device -----------
    generation = None
    if args.gen_envs > 0 and rank == 0:
        from bc_gym_planning_env_b200.vec_aisle_env import VecRandomAisleTurnEnv
        genv = env if args.worlds == "device" else VecRandomAisleTurnEnv(args.gen_envs, aisle_params(), seed=99, device=device)
        for _ in range(2):
            genv.generate()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        g0.record()
        for _ in range(reps):
            genv.generate()
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / reps
        genv.check_status()
        generation = {"worlds_per_sec": genv.n_envs / (gms * 1e-3), "ms_per_launch": gms, "envs": genv.n_envs,
                      "kernel": "generate_aisles_kernel",
                      "what": "RandomAisleTurnEnv.reset with draw_new_turn_on_reset for every env: turn draw, five walls "
                              "(old ones erased), both tile planes, refined path, initial state"}
        if genv is not env:
            del genv
            torch.cuda.empty_cache()
        # RandomMiniEnv.reset for a configs[1]-sized batch: sample (with its collision / too-close rejection), rasterise,
        # initial state -- the reference does ~100 of these per second on a core
        from bc_gym_planning_env_b200.vec_aisle_env import VecRandomMiniEnv
        menv = VecRandomMiniEnv(4096, seed=98, device=device)
        menv.generate()
        torch.cuda.synchronize()
        g0.record()
        for _ in range(reps):
            menv.generate()
        g1.record()
        torch.cuda.synchronize()
        menv.check_status()
        generation["mini"] = {"worlds_per_sec": menv.n_envs / (g0.elapsed_time(g1) / reps * 1e-3), "envs": menv.n_envs,
                              "ms_per_launch": g0.elapsed_time(g1) / reps, "kernel": "generate_minis_kernel"}
        del menv
        torch.cuda.empty_cache()

    # ---- e2e: host actions in, host results out, every step -----------------------------------------
    h_actions = [a.cpu().pin_memory() for a in actions]
    h2d = h_actions[0].numel() * 4
    d2h = n * 8 + n + n * 12 * 4
    env.step_host(h_actions[0])                               # allocates the pinned buffers and the side stream
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.e2e_steps):
        env.step_host(h_actions[k % n_sets])                  # synchronises: the caller acts on the result
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n * world * args.e2e_steps / (float(t.item()) * 1e-3)

    # same, with the egocentric crops and goal vectors copied to pinned host memory every step as well (a host-side
    # consumer of the images): bound by the host link, reported beside the headline e2e
    e2e_images = None
    if not args.no_ego and args.e2e_image_steps > 0:
        env.step_host(h_actions[0], images=True)              # allocates the pinned image buffer
        barrier()
        e0.record()
        for k in range(args.e2e_image_steps):
            env.step_host(h_actions[k % n_sets], images=True)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        img_bytes = env.ego_image.numel() + env.goal_n_state.numel() * 4
        e2e_images = {"value": n * world * args.e2e_image_steps / (float(t.item()) * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + img_bytes, "steps": args.e2e_image_steps,
                      "d2h_gbs": (d2h + img_bytes) * args.e2e_image_steps / (float(t.item()) * 1e-3) / 1e9}

    if rank == 0:
        cpu = cpu_baseline_sample(args.cpu_seconds, not args.no_ego) if args.gpus == 1 else None
        if generation is not None and args.gpus == 1:
            generation["cpu_baseline"] = cpu_generation_sample(min(args.cpu_seconds, 3.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, n),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "VecPlanEnv.step_host: pinned-host actions in; reward f64, done u8 and the 12-float compact "
                            "observation out to pinned host memory every step (copies overlap the egocentric kernel); "
                            "egocentric images stay in HBM for a GPU-resident policy; `with_images_to_host` = the same "
                            "call with every crop and goal vector copied to pinned host memory too (host-link bound)",
                    "with_images_to_host": e2e_images},
            "gpu_launches": args.steps * ((3 if split else 1) + (0 if args.no_ego else (1 if (sparse and env._batch.flags & 1) else (2 if sparse else 1)))) * world,
            "ego_dense_fallback_envs_last_step": None if args.no_ego else dense_envs,
            "roofline": dominant,
            "roofline_collision": roof_commit,
            "roofline_ego": None if args.no_ego else roof_ego,
            "kernels_ms": dict({"kin_kernel": kin_ms, "collide_reward_kernel": cr_ms, "commit_kernel": commit_ms} if split
                               else {"state_kernel": commit_ms},
                               **{ego_name: ego_ms, "collision_tiles_cold_l2": tiles_ms, "collision_u8_cold_l2": u8_ms}),
            "collision_standalone": {
                "tiles": roof("collision_kernel (lethal tile plane), cold L2", coll_bytes, tiles_ms, "uint8-definition bytes"),
                "u8": roof("collision_kernel (uint8 rows), cold L2", coll_bytes, u8_ms, "uint8-definition bytes"),
            },
            "value_without_ego_obs": no_ego_value,
            "generation": generation,
            "episode_stats": {k: float(v) for k, v in zip(nat.STAT_NAMES, stats.tolist())},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
