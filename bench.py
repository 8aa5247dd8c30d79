#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PlanEnv.step path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path, configs[2] (the headline)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm: the UNMODIFIED reference when it
                                                                     # is importable (oracle/ref_loader.py), else the port
    python bench.py --workload mini4096 | corridor16384 | montecarlo # the other configs, same JSON contract

Headline workload (`config.workload`): AisleTurnEnv, tricycle robot, EnvParams(control_delay=2, pose_delay=1,
state_delay=1), PlanEnv's odometry noise on (Philox4x32-10), auto-reset on done, egocentric observation assembled every
step.  A "step" is one `VecPlanEnv.step` over all envs of the rank.  Envs shard across ranks with no data-path
collective; the only collective is one all-reduce of the episode statistics at the end of the timed region.
  * `value` (scaling "weak"): 65 536 envs PER GPU.  The timed region is K calls of the user's `VecPlanEnv.step`; every
    4th of them records CUDA events between its kernels (the per-kernel times of the roofline objects);
  * `e2e`: the same through `VecPlanEnv.step_host`: pinned-host actions up, reward / done / observation vector down to
    pinned host memory every step, the caller waiting for them before it submits the next actions;
  * `strong` (N > 1): BASELINE's literal "65 536 envs at 1/2/4/8 B200": 65 536 envs in TOTAL, 65 536 / N per GPU,
    stepped the faster of two ways (`step`: dependent launches; `step_graph`: one captured graph per step).
One JSON line is printed by rank 0.
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
TOTAL_ENVS = 65536


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="aisle", choices=["aisle", "mini4096", "corridor16384", "montecarlo"],
                    help="aisle: BASELINE configs[2] (default, the headline); mini4096: configs[1]; corridor16384: configs[3] "
                         "(synthetic stand-in); montecarlo: configs[4]")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's own size)")
    ap.add_argument("--pool", type=int, default=2048, help="distinct random aisle maps generated on the host")
    ap.add_argument("--no-ego", action="store_true", help="skip the egocentric observation kernel (not the headline)")
    ap.add_argument("--ego-staging", default="tiles", choices=["tiles", "tma", "spans"],
                    help="how the egocentric kernel stages its source window (see VecPlanEnv)")
    ap.add_argument("--worlds", default="pool", choices=["pool", "device"],
                    help="pool: --pool host-built maps replicated on device (default); device: every env's own world drawn "
                         "and rasterised on the GPU (bcg_generate_aisles; about 1.2 MB of slots per env)")
    ap.add_argument("--gen-envs", type=int, default=8192, help="envs of the device-generation (reset storm) measurement; 0 = skip")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--e2e-image-steps", type=int, default=5, help="steps of the images-to-host e2e variants; 0 = skip")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-envs", type=int, default=16, help="reference arm: envs advanced per worker per step")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling measurement (N > 1 only)")
    ap.add_argument("--mc-states", type=int, default=256)
    ap.add_argument("--mc-rollouts", type=int, default=1024)
    ap.add_argument("--mc-horizon", type=int, default=128)
    return ap.parse_args()


def workload_config(args, n_envs, what=None):
    return {
        "workload": what or ("AisleTurnEnv x %d envs/GPU, tricycle, control/pose/state delay %d/%d/%d, Philox odometry noise, "
                             "auto-reset, egocentric obs %s" % (n_envs, DELAYS[0], DELAYS[1], DELAYS[2], "off" if args.no_ego else "on")),
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
        self.extended = False

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
        """The same load kept running after the timed region (see keep_load_until_sampled): samples up to now count."""
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
        scope = "timed region + the same load kept running right after it" if self.extended else "timed region"
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


def keep_load_until_sampled(sampler, env, actions, chunk=20):
    """A timed region shorter than a few polls of nvidia-smi: keep the identical load running (untimed) until the sampler
    has seen it, so that the clocks line always describes this workload under load."""
    import torch
    t_roll = time.time()
    while sampler.samples_since_start() < 5 and time.time() - t_roll < 3.0:
        for k in range(chunk):
            env.step(actions[k % len(actions)])
        torch.cuda.synchronize()
        sampler.extend_end()


# ------------------------------------------------------------------------------------------------
# CPU legs (the ONLY places bench.py executes oracle/): the unmodified reference where it is importable, the port else
# ------------------------------------------------------------------------------------------------
def _action_bounds():
    s = 60. * np.pi / 180.
    return np.array([s / 10, -np.pi / 2]), np.array([s / 2, np.pi / 2])


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


def reference_kind():
    """'reference' when the unmodified reference package can be imported here (the build container's /root/reference, the
    driver's baseline/_ref, or the install oracle/make_ref.py leaves in oracle/_ref/, which travels to the GPU box), else
    'port'."""
    from oracle.ref_loader import reference_available
    return "reference" if reference_available() else "port"


class _PortEnvs(object):
    """n aisle envs of the NumPy oracle port, stepped with egocentric observation and auto-reset"""

    def __init__(self, n, seed, with_ego):
        from oracle import plan_env_oracle as O
        from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
        self.O, self.with_ego = O, with_ego
        costmaps, paths = random_aisle_pool(n, seed, aisle_params())
        self.envs = []
        for i, (cm, path) in enumerate(zip(costmaps, paths)):
            src = (lambda env_id: (lambda step: O.philox_normal_source(0, env_id, step)))(seed + i)
            self.envs.append(O.OraclePlanEnv(cm.get_data(), cm.get_origin(), cm.get_resolution(), path, delays=DELAYS,
                                             alphas=O.DEFAULT_NOISE, normal_source=src))
        self.rng = np.random.RandomState(seed)
        self.low, self.high = _action_bounds()

    def advance(self):
        O = self.O
        for env in self.envs:
            obs, _, done, _ = env.step(self.rng.uniform(self.low, self.high).astype(np.float32))
            if self.with_ego:
                O.ego_costmap(env.costmap, obs["pose"], env.origin, env.resolution)
                O.goal_n_state(obs["path"], obs["pose"], obs["robot_state"], env.resolution)
            if done:
                env.reset()
        return len(self.envs)


class _ReferenceEnvs(object):
    """n envs of the UNMODIFIED reference: EgocentricCostmap(RandomAisleTurnEnv(EnvParams(delays 2/1/1))) with PlanEnv's own
    odometry noise, random actions within its action space, reset() (a new random turn, synth_turn_env.py:278-291) on done"""

    def __init__(self, n, seed, with_ego):
        from oracle.ref_loader import load_reference
        load_reference()
        from bc_gym_planning_env.envs.base.action import Action
        from bc_gym_planning_env.envs.base.params import EnvParams
        from bc_gym_planning_env.envs.egocentric import EgocentricCostmap
        from bc_gym_planning_env.envs.synth_turn_env import RandomAisleTurnEnv
        self.Action = Action
        ep = EnvParams(control_delay=DELAYS[0], pose_delay=DELAYS[1], state_delay=DELAYS[2])
        self.envs = []
        for i in range(n):
            env = RandomAisleTurnEnv(params=ep, seed=seed + i)
            self.envs.append(EgocentricCostmap(env) if with_ego else env)
        self.rng = np.random.RandomState(seed)
        self.low, self.high = _action_bounds()

    def advance(self):
        for env in self.envs:
            a = self.rng.uniform(self.low, self.high).astype(np.float32)
            _, _, done, _ = env.step(self.Action(command=a))
            if done:
                env.reset()
        return len(self.envs)


def make_cpu_envs(kind, n, seed, with_ego):
    return (_ReferenceEnvs if kind == "reference" else _PortEnvs)(n, seed, with_ego)


def _time_cpu(envs, seconds):
    envs.advance()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        n += envs.advance()
    return n, time.perf_counter() - t0


def cpu_baseline_sample(seconds, with_ego):
    """Single process, one host core, ~`seconds` s: the unmodified reference when importable (kind "reference"), and the
    NumPy port beside it (a third of the budget) so that both are on record."""
    _single_threaded_math()
    kind = reference_kind()
    out = {"unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count()}
    if kind == "reference":
        n, dt = _time_cpu(make_cpu_envs("reference", 4, 900, with_ego), seconds * 2. / 3.)
        out["value"] = n / dt
        out["sample"] = ("%d env-steps of the unmodified reference (4 x EgocentricCostmap(RandomAisleTurnEnv), delays 2/1/1, its own "
                         "noise, reset on done) in %.1f s on 1 core" % (n, dt))
        n, dt = _time_cpu(make_cpu_envs("port", 4, 900, with_ego), seconds / 3.)
        out["port_value"] = n / dt
    else:
        n, dt = _time_cpu(make_cpu_envs("port", 4, 900, with_ego), seconds)
        out["value"] = out["port_value"] = n / dt
        out["sample"] = ("%d env-steps of the NumPy oracle port (4 aisle envs, same delays/noise/ego settings) in %.1f s on 1 core "
                         "(the reference package is not importable here)" % (n, dt))
    return out


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


_REF_BARRIER = None


def _ref_init(barrier):
    global _REF_BARRIER
    _REF_BARRIER = barrier


def _ref_worker(args):
    wid, kind, n_envs, n_steps, warmup, with_ego = args
    _single_threaded_math()
    envs = make_cpu_envs(kind, n_envs, 1000 + 100 * wid, with_ego)
    for _ in range(warmup):
        envs.advance()
    _REF_BARRIER.wait()                      # all workers start the timed steps together
    t0 = time.perf_counter()
    n = 0
    for _ in range(n_steps):
        n += envs.advance()
    return n, t0, time.perf_counter()


def _run_cpu_pool(kind, args, with_ego):
    import multiprocessing as mp
    workers = max(1, os.cpu_count() or 1)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(workers)
    with ctx.Pool(workers, initializer=_ref_init, initargs=(barrier,)) as pool:
        res = pool.map(_ref_worker, [(w, kind, args.ref_envs, args.steps, args.warmup, with_ego) for w in range(workers)], chunksize=1)
    n = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    return n, wall, workers


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores -- the UNMODIFIED reference
    package when it is importable (oracle/_ref on the GPU box), else its NumPy port.  A step = every worker advances
    `--ref-envs` envs once.  The port's throughput under the same load is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with_ego = not args.no_ego
    kind = reference_kind()
    from oracle import plan_env_oracle  # noqa: F401  (import once here, not in every forked worker)
    from bc_gym_planning_env_b200.envs import synth_turn_env  # noqa: F401
    n, wall, workers = _run_cpu_pool(kind, args, with_ego)
    value = n / wall
    port_value = value
    if kind == "reference":
        pn, pwall, _ = _run_cpu_pool("port", args, with_ego)
        port_value = pn / pwall
    what = ("the unmodified reference (EgocentricCostmap(RandomAisleTurnEnv), delays 2/1/1, its own noise, reset on done)"
            if kind == "reference" else "the NumPy oracle port (the reference package is not importable here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.envs or TOTAL_ENVS),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "port_value": port_value,
                         "sample": "%d workers x %d envs x %d steps of %s" % (workers, args.ref_envs, args.steps, what)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the CUDA arm: shared pieces
# ------------------------------------------------------------------------------------------------
class Dist(object):
    """rank / world / device of this process, barrier and max-over-ranks"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from bc_gym_planning_env_b200 import _native as nat
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        nat.require_cuda()
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.dist = dist
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        import torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def build_aisle_env(args, rank, device, n_envs, env_id_base, compact=True):
    from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    params = aisle_params()
    if args.worlds == "device":
        from bc_gym_planning_env_b200.vec_aisle_env import VecRandomAisleTurnEnv
        return VecRandomAisleTurnEnv(n_envs, params, seed=1234, auto_reset=True, device=device,
                                     env_id_base=env_id_base, with_ego=not args.no_ego)
    costmaps, paths = random_aisle_pool(min(args.pool, n_envs), 10000 + rank * args.pool, params)
    return VecPlanEnv(costmaps, paths, params, n_envs=n_envs, seed=1234, auto_reset=True, device=device,
                      env_id_base=env_id_base, private_map_copies=True, with_ego=not args.no_ego,
                      ego_staging=args.ego_staging,
                      compact_ego=compact and not args.no_ego and args.ego_staging == "tiles" and args.e2e_image_steps > 0)


def random_actions(env, device, seed, n_sets=8):
    import torch
    low, high = env.action_bounds()
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    lo_t, hi_t = torch.from_numpy(low).to(device), torch.from_numpy(high).to(device)
    return [(lo_t + (hi_t - lo_t) * torch.rand((env.n_envs, 2), generator=gen, device=device)).contiguous() for _ in range(n_sets)]


EVENT_EVERY = 4          # timed region: per-kernel CUDA events on every 4th step


def timed_steps(env, actions, steps, D, sampler=None):
    """The timed region: `steps` env steps bracketed by a barrier + synchronize on both sides, then one all-reduce of the
    episode statistics.  Every EVENT_EVERY-th step is VecPlanEnv.step_timed (CUDA events between the kernels, on the stream
    the kernels run on: the per-kernel times of the roofline objects come from inside the timed region); the others are the
    call a user makes, VecPlanEnv.step, whose three launches are programmatic dependents of one another -- an event record
    between two kernels is a stream operation of its own and switches that overlap off.
    Returns (ms max over ranks, per-kernel mean ms, stats)."""
    import torch
    from bc_gym_planning_env_b200.parallel import allreduce_episode_stats
    n_sets = len(actions)
    timed = [k for k in range(steps) if k % EVENT_EVERY == 0]
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(5)] for k in timed}
    ev = [ev[k] for k in timed]
    for evs in ev:
        for e in evs:
            e.record()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    if sampler is not None:
        sampler.mark_start()
    t_start.record()
    for k in range(steps):
        if k % EVENT_EVERY == 0:
            env.step_timed(actions[k % n_sets], ev[k // EVENT_EVERY])
        else:
            env.step(actions[k % n_sets])
    stats = allreduce_episode_stats(env)      # the path's only collective
    t_end.record()
    D.barrier()
    if sampler is not None:
        sampler.mark_end()
    ms = D.max_ms(t_start.elapsed_time(t_end))
    kern = {"move_kernel": float(np.mean([e[1].elapsed_time(e[2]) for e in ev])),
            "reward_kernel": float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))}
    kern["ego"] = float(np.mean([e[3].elapsed_time(e[4]) for e in ev]))
    return ms, kern, stats


def plain_steps(env, actions, steps, D, graph=False):
    """`steps` steps without per-kernel events (graph: through the captured CUDA graph); ms max over ranks"""
    import torch
    n_sets = len(actions)
    if graph:
        env.step_graph(actions[0])           # captures on the first call
    D.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for k in range(steps):
        if graph:
            env.step_graph(actions[k % n_sets])
        else:
            env.step(actions[k % n_sets])
    a1.record()
    D.barrier()
    return D.max_ms(a0.elapsed_time(a1))


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(n_envs):
    """DRAM bytes per launch of each kernel from the committed `ncu --set full` capture of the headline workload
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum at 65 536 envs); {} for other sizes"""
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("envs_per_gpu", -1)) == n_envs:
            return tj.get("dram_bytes_per_launch", {})
    return {}


def roof(kernel, alg_bytes, ms, note, traffic=None):
    """roofline object: achieved = algorithmic bytes / launch time; `traffic` = DRAM bytes ncu counted per launch"""
    peak, peak_src = hbm_peak()
    ach = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    phys = traffic / (ms * 1e-3) / 1e9 if (traffic and ms > 0) else None
    return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "traffic_gbs": phys, "traffic_frac": None if phys is None else phys / peak,
            "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms, "peak_source": peak_src, "note": note}


def ego_roofline(env, ego_ms, n, sparse, traffic):
    cp = env._c_params
    gather_bytes = 2.0 * cp.ego_w * cp.ego_h * n                  # source read + image write per env
    if not sparse:
        name = "ego_tiles_kernel" if env.ego_staging == "tiles" else "ego_kernel"
        return name, roof(name, gather_bytes, ego_ms, "algorithmic bytes = 2 x ego_w x ego_h per env (SURVEY 8d: source read + "
                          "image write); staging: " + env.ego_staging, traffic.get(name))
    # The sparse kernel is a scatter: it never reads the source bytes, so SURVEY 8d's gather bytes (2 x W x H per env)
    # are not what it moves and give "fractions" above 1 (kept under `survey_8d`).  Its algorithmic bytes are the crop it
    # must produce, W x H per env; `traffic` is what ncu counted in DRAM for one launch (profiles/ncu_traffic.json) -- less
    # than the crops' bytes, because VecPlanEnv keeps them in memory with compute-data compression.
    name = "ego_sparse_kernel"
    image_bytes = float(cp.ego_w * cp.ego_h) * n
    compressed = bool(getattr(env, "image_memory_compressed", False))
    tr = traffic.get(name if compressed else name + "_plain_memory")
    r = roof(name, image_bytes, ego_ms,
             "scatter kernel: algorithmic bytes = the crops it writes (ego_w x ego_h per env); `traffic` = DRAM bytes per "
             "launch from the committed ncu capture (crops in %s memory); `survey_8d` = the gather definition of SURVEY 8d "
             "(2 x W x H per env), which a scatter does not move and which can exceed 1; `image_fill` = plain fills of the "
             "same bytes (the write-only floor)" % ("compressible" if compressed else "plain"), tr)
    peak = r["peak"]
    r["image_memory_compressed"] = compressed
    r["survey_8d"] = {"algorithmic_bytes_per_launch": gather_bytes, "frac": gather_bytes / (ego_ms * 1e-3) / 1e9 / peak}
    r["image_fill"] = image_fill_floor(env, ego_ms)
    return name, r


def image_fill_floor(env, ego_ms):
    """Context for the scatter kernel's time: plain fills (library calls, not the product) of a buffer of the images' size,
    timed like the kernels (CUDA events, mean of 20 back-to-back launches, buffer far larger than L2): torch's `zero_`
    (the driver's memset) and `fill_` of the same bytes viewed as int64 and as uint8."""
    import torch
    nbytes = env.ego_image.numel()
    scratch = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=env.ego_image.device)
    u8 = scratch.view(torch.uint8)

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    ms = {"memset": timed(scratch.zero_), "fill_int64": timed(lambda: scratch.fill_(1)), "fill_uint8": timed(lambda: u8.fill_(1))}
    best = min(ms.values())
    del scratch, u8
    return {"ms": ms, "best_ms": best, "best_gbs": nbytes / (best * 1e-3) / 1e9,
            "kernel_ms_over_best_fill_ms": ego_ms / best if best > 0 else None,
            "what": "write-only floor of the observation on this GPU: plain fills of a buffer of the crops' size"}


def measure_e2e(env, actions, args, D, n):
    """VecPlanEnv.step_host: pinned-host actions in, reward / done / compact observation out to pinned host memory every
    step.  `resident`: the egocentric images stay in HBM for a GPU-resident policy -- the headline e2e; `compact` /
    `dense`: the images go to the host as well (as per-env lists of non-zero pixels packed on the device / as they are)."""
    import torch
    world = D.world
    n_sets = len(actions)
    h_actions = [a.cpu().pin_memory() for a in actions]
    h2d = h_actions[0].numel() * 4
    d2h = n * 8 + n + n * 12 * 4
    out = {"launches_per_step": env.launches_per_step()}

    def loop(steps, **kw):
        env.step_host(h_actions[0], **kw)                        # allocates the pinned buffers and the side stream
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        extra = 0
        for k in range(steps):
            res = env.step_host(h_actions[k % n_sets], **kw)     # synchronises: the caller acts on the result
            if kw.get("images") == "compact":
                extra += int(res[3]["bytes"])
        e1.record()
        D.barrier()
        return D.max_ms(e0.elapsed_time(e1)), extra

    ms, _ = loop(args.e2e_steps)
    out["resident"] = {"value": n * world * args.e2e_steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h,
                       "note": "VecPlanEnv.step_host: pinned-host actions in (upload stream); reward f64, done u8 and the "
                               "12-float compact observation out to pinned host memory every step (copies on a side stream "
                               "right after reward_kernel); the call returns when those are on the host, so the host "
                               "prepares and uploads the next actions while the egocentric kernel of this step still runs -- "
                               "that is why this figure is close to `value`; egocentric images stay in HBM for a GPU-resident "
                               "consumer (e2e_images_to_host has the variants that wait for them)"}
    out["compact"] = out["dense"] = None
    if getattr(env, "_ego_hits", None) is not None and args.e2e_image_steps > 0:
        steps = args.e2e_image_steps * 4
        ms, extra = loop(steps, images="compact")
        per_step = d2h + extra / steps + n * 9 * 4
        out["compact"] = {"value": n * world * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                          "d2h_bytes_per_step": per_step, "steps": steps, "d2h_gbs": per_step * steps / (ms * 1e-3) / 1e9}
    if env.with_ego and args.e2e_image_steps > 0:
        ms, _ = loop(args.e2e_image_steps, images=True)
        img_bytes = env.ego_image.numel() + env.goal_n_state.numel() * 4
        out["dense"] = {"value": n * world * args.e2e_image_steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h + img_bytes, "steps": args.e2e_image_steps,
                        "d2h_gbs": (d2h + img_bytes) * args.e2e_image_steps / (ms * 1e-3) / 1e9}
    return out


IMAGES_NOTE = ("the same call with the egocentric observation delivered to pinned host memory every step as well: `compact` = "
               "per env the list of its non-zero crop pixels (offset | value << 16), packed on the device (bcg_pack_ego_hits) "
               "and expanded by VecPlanEnv.expand_compact; `dense` = every ego_h x ego_w crop as it is (host-link bound)")


# ------------------------------------------------------------------------------------------------
# the headline: BASELINE configs[2]
# ------------------------------------------------------------------------------------------------
def run_aisle(args, D):
    import ctypes as C
    import torch
    from bc_gym_planning_env_b200 import _native as nat

    world, rank, device = D.world, D.rank, D.device
    n = args.envs or TOTAL_ENVS
    sampler = ClockSampler(D.local_rank)
    sampler.start()                           # nvidia-smi needs a few hundred ms to come up: start it before the set-up
    env = build_aisle_env(args, rank, device, n, rank * n)
    actions = random_actions(env, device, 4321 + rank)
    n_sets = len(actions)
    sampler.wait_first_sample()
    for w in range(args.warmup):
        env.step(actions[w % n_sets])
    env.episode_stats(reset=True)
    D.barrier()

    # ---- timed region: K steps, inputs resident in HBM ------------------------------------------
    elapsed_ms, kern, stats = timed_steps(env, actions, args.steps, D, sampler)
    keep_load_until_sampled(sampler, env, actions)
    clocks = sampler.stop()
    env.check_status()
    value = n * world * args.steps / (elapsed_ms * 1e-3)
    # the driver's --steps can make the timed region a few ms: the same loop over >= 200 steps is printed beside it
    long_steps = max(200, args.steps)
    long_ms = plain_steps(env, actions, long_steps, D)
    value_long = {"steps": long_steps, "ms_per_step": long_ms / long_steps, "value": n * world * long_steps / (long_ms * 1e-3),
                  "note": "the same loop without the per-kernel events, over %d steps" % long_steps}

    # ---- same loop without the egocentric kernel (reported beside the headline, not instead of it) -----
    no_ego_value = None
    if not args.no_ego:
        saved = (env._out.ego_image, env._out.goal_n_state)
        env._out.ego_image = env._out.goal_n_state = None
        ms = plain_steps(env, actions, args.steps, D)
        no_ego_value = n * world * args.steps / (ms * 1e-3)
        env._out.ego_image, env._out.goal_n_state = saved

    # ---- algorithmic bytes -----------------------------------------------------------
    cand_pose = env._cand[:3].t().contiguous()
    _, pixels = env.pose_collides(cand_pose, count_pixels=True)
    coll_bytes = float(pixels.sum().item())                       # 1 B per in-map footprint pixel (uint8 costmap)
    remaining = float((env.path_lengths() - env.state_i[nat.I_TARGET].to(torch.int64)).clamp(min=0).double().mean().item())
    scan_bytes = 24.0 * remaining * n                             # (N - target_idx) x 24 B per env
    traffic = ncu_traffic(n)
    sparse = getattr(env, "_ego_list", None) is not None and os.environ.get("BCG_EGO_KERNEL") != "dense"
    ego_ms = kern.pop("ego")
    ego_name, roof_ego = ego_roofline(env, ego_ms, n, sparse, traffic)
    kern[ego_name] = ego_ms
    dense_envs = int(env._ego_list[n].item()) if sparse else n
    roof_reward = None
    if "reward_kernel" in kern:
        roof_reward = roof("reward_kernel", scan_bytes, kern["reward_kernel"],
                           "algorithmic bytes (SURVEY 8d) = remaining path points x 24 B (the kernel skips path chunks farther than "
                           "the reach radius, so it moves far fewer)", traffic.get("reward_kernel"))

    # ---- stand-alone collision kernels, cold L2 (flush between launches) ----------------------------
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=device)

    def time_kernel(fn, reps=7):
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

    flags = torch.empty(n, dtype=torch.uint8, device=device)
    s = env._stream()
    # work records of the poses the last step proposed (one prep launch), then only the collision kernel is timed
    nat.check(nat.lib().bcg_collision(C.byref(env._c_params), C.byref(env._batch), None, nat.ptr(flags), None, s))
    tiles_ms = time_kernel(lambda: nat.check(nat.lib().bcg_collision_recheck(C.byref(env._c_params), C.byref(env._batch), nat.ptr(flags), 0, s)))
    u8_ms = time_kernel(lambda: nat.check(nat.lib().bcg_collision_recheck(C.byref(env._c_params), C.byref(env._batch), nat.ptr(flags), 1, s)))
    del flush
    roof_collision = roof("collision_thread_kernel (lethal tile plane + tile summary), cold L2", coll_bytes, tiles_ms,
                          "the north_star's collision roofline: algorithmic bytes (SURVEY 8d) = in-map footprint pixels x 1 B on the "
                          "reference's uint8 costmap (%.1f MB); the kernel reads the derived 1-bit lethal tile plane, and of it only "
                          "the non-empty tiles under the footprint; stand-alone launch of the very code the step's move_kernel "
                          "inlines, L2 flushed before every launch" % (coll_bytes / 1e6), traffic.get("collision_thread_kernel"))

    # ---- reset storm: every env of a batch gets a new world on the device -----------
    generation = None
    if args.gen_envs > 0 and rank == 0:
        from bc_gym_planning_env_b200.vec_aisle_env import VecRandomAisleTurnEnv, VecRandomMiniEnv
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
    e2e = measure_e2e(env, actions, args, D, n)

    # ---- the same step as one CUDA-graph launch -----------------------------------------------------------
    graph_ms = plain_steps(env, actions, long_steps, D, graph=True)
    value_graph = {"value": n * world * long_steps / (graph_ms * 1e-3), "ms_per_step": graph_ms / long_steps, "steps": long_steps,
                   "note": "VecPlanEnv.step_graph: the step's kernels replayed as one captured CUDA graph (device-side step counter)"}

    # ---- strong scaling: BASELINE's "65 536 envs at 1/2/4/8 B200" = 65 536 envs in total ----------------
    strong = None
    if world > 1 and not args.no_strong and args.worlds == "pool":
        del env
        torch.cuda.empty_cache()
        strong = measure_strong(args, D)
    elif world == 1 and n == TOTAL_ENVS:
        strong = {"value": value_graph["value"], "ms_per_step": value_graph["ms_per_step"], "envs_total": TOTAL_ENVS,
                  "envs_per_gpu": n, "steps": long_steps, "scaling": "strong",
                  "note": "one GPU: the strong and the weak workload are the same batch (the CUDA-graph figure)"}

    if rank == 0:
        cpu = cpu_baseline_sample(args.cpu_seconds, not args.no_ego) if args.gpus == 1 else None
        if generation is not None and args.gpus == 1:
            generation["cpu_baseline"] = cpu_generation_sample(min(args.cpu_seconds, 3.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, n),
            "clocks": clocks,
            "e2e": e2e["resident"],
            "e2e_images_to_host": {"compact": e2e["compact"], "dense": e2e["dense"], "note": IMAGES_NOTE},
            "gpu_launches": args.steps * e2e["launches_per_step"] * world,
            "ego_dense_fallback_envs_last_step": None if args.no_ego else dense_envs,
            "roofline": roof_ego if not args.no_ego else roof_collision,
            "roofline_collision": roof_collision,
            "roofline_reward": roof_reward,
            "kernels_ms": dict(kern, collision_tiles_cold_l2=tiles_ms, collision_u8_cold_l2=u8_ms),
            "collision_standalone": {
                "tiles": roof_collision,
                "u8": roof("collision_kernel (uint8 rows), cold L2", coll_bytes, u8_ms, "uint8-definition bytes"),
            },
            "value_200_steps": value_long,
            "value_graph": value_graph,
            "strong": strong,
            "value_without_ego_obs": no_ego_value,
            "generation": generation,
            "episode_stats": {k: float(v) for k, v in zip(nat.STAT_NAMES, stats.tolist())},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))


def measure_strong(args, D):
    """65 536 envs in TOTAL over the ranks (BASELINE's metric, SURVEY 8e "N / G per GPU"), every rank stepping its share
    through the captured CUDA graph; same workload, same barrier + max-over-ranks timing."""
    from bc_gym_planning_env_b200.parallel import shard_range
    lo, hi = shard_range(TOTAL_ENVS, D.rank, D.world)
    env = build_aisle_env(args, D.rank, D.device, hi - lo, lo, compact=False)
    actions = random_actions(env, D.device, 977 + D.rank)
    for w in range(max(args.warmup, 3)):
        env.step(actions[w % len(actions)])
    steps = max(200, args.steps)
    ms_timed, kern, _ = timed_steps(env, actions, steps, D)
    ms_graph = plain_steps(env, actions, steps, D, graph=True)
    ms_plain = plain_steps(env, actions, steps, D, graph=False)
    env.check_status()
    best, how = min((ms_plain, "VecPlanEnv.step: three launches per step, each a programmatic dependent of the one before"),
                    (ms_graph, "VecPlanEnv.step_graph: one CUDA-graph launch per step"))
    return {"value": TOTAL_ENVS * steps / (best * 1e-3), "ms_per_step": best / steps, "envs_total": TOTAL_ENVS,
            "envs_per_gpu": hi - lo, "steps": steps, "scaling": "strong", "stepped_with": how,
            "value_graph": TOTAL_ENVS * steps / (ms_graph * 1e-3), "value_step": TOTAL_ENVS * steps / (ms_plain * 1e-3),
            "value_with_events_between_kernels": TOTAL_ENVS * steps / (ms_timed * 1e-3), "kernels_ms": kern,
            "note": "65 536 envs in total, sharded by contiguous env range; the faster of the two ways of stepping"}


# ------------------------------------------------------------------------------------------------
# BASELINE configs[1] and configs[3]
# ------------------------------------------------------------------------------------------------
def run_small_workload(args, D):
    """configs[1] (mini4096) and configs[3] (corridor16384, synthetic stand-in): same contract, one batch per GPU."""
    from bc_gym_planning_env_b200 import _native as nat
    from bc_gym_planning_env_b200.envs.base.params import EnvParams
    world, rank, device = D.world, D.rank, D.device
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    if args.workload == "mini4096":
        from bc_gym_planning_env_b200.vec_aisle_env import VecRandomMiniEnv
        n = args.envs or 4096
        env = VecRandomMiniEnv(n, seed=2024, noise_parameters=None, auto_reset=True, device=device, env_id_base=rank * n,
                               with_ego=not args.no_ego)
        what = ("RandomMiniEnv x %d envs/GPU (BASELINE configs[1]): per-env world sampled and rasterised on the device "
                "(bcg_generate_minis), tricycle, noise off, goal tolerances 0.2 m / pi/8, auto-reset, egocentric obs %s"
                % (n, "off" if args.no_ego else "on"))
    else:
        from bc_gym_planning_env_b200.envs.rw_corridors.tdwa_test_environments import \
            get_random_maps_squeeze_between_obstacle_in_corridor_on_path
        from bc_gym_planning_env_b200.vec_env import VecPlanEnv
        n = args.envs or 16384
        _, path, variants = get_random_maps_squeeze_between_obstacle_in_corridor_on_path(n_variants=1000, seed=1 + rank)
        ep = EnvParams(iteration_timeout=1200, pose_delay=1, control_delay=0, state_delay=1, goal_spat_dist=1.0,
                       goal_ang_dist=np.pi / 2, dt=0.05)          # the reference runner's (rw_randomized_corridor_3_boxes.py:20-29)
        env = VecPlanEnv(list(variants), [path], ep, n_envs=n, map_ids=np.arange(n) % len(variants),
                         path_ids=np.zeros(n, dtype=np.int64), auto_reset=True, device=device, env_id_base=rank * n,
                         with_ego=not args.no_ego)
        what = ("rw_randomized_corridor_3_boxes x %d envs/GPU (BASELINE configs[3]) on a SYNTHETIC STAND-IN (the S3 corridor data "
                "cannot be fetched): 667 x 267 corridor with 255 / 253 / 254 / 0 cells, 3 jittered boxes, 1000 variants, tricycle, "
                "pose/state delay 1/1, Philox odometry noise, auto-reset, egocentric obs %s" % (n, "off" if args.no_ego else "on"))
    actions = random_actions(env, device, 555 + rank)
    sampler.wait_first_sample()
    for w in range(args.warmup):
        env.step(actions[w % len(actions)])
    env.episode_stats(reset=True)
    elapsed_ms, kern, stats = timed_steps(env, actions, args.steps, D, sampler)
    keep_load_until_sampled(sampler, env, actions, chunk=50)
    clocks = sampler.stop()
    env.check_status()
    value = n * world * args.steps / (elapsed_ms * 1e-3)
    long_steps = max(200, args.steps)
    graph_ms = plain_steps(env, actions, long_steps, D, graph=True)
    sparse = getattr(env, "_ego_list", None) is not None and os.environ.get("BCG_EGO_KERNEL") != "dense"
    ego_ms = kern.pop("ego")
    roof_ego = None
    if not args.no_ego:
        ego_name, roof_ego = ego_roofline(env, ego_ms, n, sparse, {})
        kern[ego_name] = ego_ms
    e2e = measure_e2e(env, actions, args, D, n)
    if rank == 0:
        cfg = workload_config(args, n, what)
        cfg.pop("map_pool", None)
        cfg.pop("maps", None)
        cfg["l2"] = ("nothing is flushed between timed iterations: the batch's own working set (maps, rows, crops: %.2f GB of crops "
                     "alone) is what a user of this config steps" % (n * 15561 / 1e9))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e["resident"], "e2e_images_to_host": {"compact": e2e["compact"], "dense": e2e["dense"], "note": IMAGES_NOTE},
            "gpu_launches": args.steps * e2e["launches_per_step"] * world,
            "roofline": roof_ego, "kernels_ms": kern,
            "value_graph": {"value": n * world * long_steps / (graph_ms * 1e-3), "ms_per_step": graph_ms / long_steps, "steps": long_steps},
            "episode_stats": {k: float(v) for k, v in zip(nat.STAT_NAMES, stats.tolist())},
            "cpu_baseline": None,
        }
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]
# ------------------------------------------------------------------------------------------------
def run_montecarlo(args, D):
    """BASELINE configs[4]: K start states taken mid-episode from a configs[2]-style batch on rank 0 (bcg_gather_state),
    NCCL broadcast of the snapshot columns, R noisy rollouts of a fixed H-step action sequence per start state fanned out
    over all ranks (global rollout g belongs to start state g // R), per-start-state statistics all-reduced -- all of it
    inside the timed region, repeated `--steps` times.  With several ranks the all-reduced statistics of two start states
    are checked, bit for bit, against a one-process replay of their rollouts."""
    import torch
    from bc_gym_planning_env_b200 import parallel
    from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
    from bc_gym_planning_env_b200.vec_env import DEFAULT_NOISE, VecPlanEnv, VecState
    world, rank, device = D.world, D.rank, D.device
    K, R, H = args.mc_states, args.mc_rollouts, args.mc_horizon
    params = aisle_params()
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    costmaps, paths = random_aisle_pool(K, 31337, params)            # the same pool on every rank (same seed)
    # the batch the start states are taken from: K envs driven for a while (every rank builds it, rank 0's is used)
    src = VecPlanEnv(costmaps, paths, params, seed=7, device=device, noise_parameters=DEFAULT_NOISE)
    gen = torch.Generator(device=device)
    gen.manual_seed(99)
    low, high = src.action_bounds()
    lo_t, hi_t = torch.from_numpy(low).to(device), torch.from_numpy(high).to(device)
    mid = torch.tensor([0.35, 0.0], device=device)                    # gentle driving: most start states are still alive
    plan = (mid + 0.35 * (lo_t + (hi_t - lo_t) * torch.rand((H, K, 2), generator=gen, device=device) - mid)).contiguous()
    for t in range(20):
        src.step(plan[t % H])
    lo_env, hi_env = parallel.shard_range(K * R, rank, world)
    n_local = hi_env - lo_env
    state_of = (np.arange(n_local) + lo_env) // R                    # a start state's R rollouts are contiguous global ids
    fan = VecPlanEnv(costmaps, paths, params, n_envs=n_local, map_ids=state_of, path_ids=state_of, seed=4242, device=device,
                     env_id_base=lo_env, noise_parameters=DEFAULT_NOISE)
    cols = torch.from_numpy(state_of).to(device)
    idx = torch.arange(K, dtype=torch.int64, device=device)
    f = torch.zeros((src.layout.n_frows, K), dtype=torch.float64, device=device)
    i = torch.zeros((src.layout.n_irows, K), dtype=torch.int32, device=device)

    def one_evaluation(events=None):
        if rank == 0:
            snap = src.get_state(idx)                                 # bcg_gather_state
            f.copy_(snap.f)
            i.copy_(snap.i)
        if events:
            events[0].record()
        parallel.broadcast_snapshot(f, i, src=0)                      # NCCL broadcast of the snapshot columns
        if events:
            events[1].record()
        out = parallel.monte_carlo_rollouts(fan, VecState(f, i), plan, reduce=False, cols=cols)
        if events:
            events[2].record()
        parallel.allreduce_sum_(out)                                  # ncclAllReduce(sum) of fp64 [K, 4]
        if events:
            events[3].record()
        return out

    sampler.wait_first_sample()
    first = one_evaluation().clone()
    # equivalence: the statistics of two start states against a one-process replay of their R rollouts (Philox streams
    # are keyed by the global env id, so how the job is split over ranks must not change a bit)
    check = {"states": [0, K // 2], "bit_equal": True, "ranks": world}
    for st in check["states"]:
        one = VecPlanEnv([costmaps[st]], [paths[st]], params, n_envs=R, seed=4242, device=device, env_id_base=st * R,
                         noise_parameters=DEFAULT_NOISE)
        got = parallel.monte_carlo_rollouts(one, VecState(f[:, st:st + 1].contiguous(), i[:, st:st + 1].contiguous()),
                                            plan[:, st:st + 1].contiguous(), reduce=False,
                                            cols=torch.zeros(R, dtype=torch.int64, device=device))
        check["bit_equal"] = check["bit_equal"] and bool(torch.equal(got[0], first[st]))
        del one
    if not check["bit_equal"]:
        raise RuntimeError("Monte-Carlo fan-out over %d ranks differs from the one-process replay" % world)
    D.barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for evs in ev:
        for e in evs:
            e.record()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    sampler.mark_start()
    t0.record()
    for k in range(args.steps):
        out = one_evaluation(ev[k])
    t1.record()
    D.barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = D.max_ms(t0.elapsed_time(t1))
    bcast_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    roll_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    red_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    fan.check_status()
    if rank == 0:
        env_steps = float(K) * R * H * args.steps
        o = out.cpu().numpy()
        line = {
            "metric": METRIC, "value": env_steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": 1, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Monte-Carlo rollouts (BASELINE configs[4]): %d start states x %d rollouts x %d steps = %d envs in "
                                   "total (%d on this GPU), AisleTurnEnv, delays 2/1/1, Philox noise, no egocentric observation; a "
                                   "step = one whole evaluation: bcg_gather_state on rank 0, NCCL broadcast, set_state fan-out, H env "
                                   "steps, all-reduce of fp64 [%d, 4]" % (K, R, H, K * R, n_local, K),
                       "start_states": K, "rollouts_per_state": R, "horizon": H, "envs_per_gpu": n_local,
                       "l2": "state rows and scratch of the batch exceed L2 at the default size; no explicit flush"},
            "clocks": clocks,
            "rollouts_per_sec": float(K) * R * args.steps / (ms * 1e-3),
            "collectives": {"broadcast_ms": bcast_ms, "allreduce_ms": red_ms, "rollouts_ms": roll_ms,
                            "share_of_step": (bcast_ms + red_ms) / (ms / args.steps),
                            "broadcast_bytes": int(f.numel() * 8 + i.numel() * 4), "allreduce_bytes": int(K * 4 * 8)},
            "equivalence_check": check,
            "result_summary": {"mean_return": float(o[:, 0].sum() / o[:, 3].sum()), "collision_rate": float(o[:, 1].sum() / o[:, 3].sum()),
                               "goal_rate": float(o[:, 2].sum() / o[:, 3].sum()), "rollouts": float(o[:, 3].sum())},
            "e2e": {"value": env_steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(K * 4 * 8),
                    "note": "the evaluation is GPU-resident end to end: start states, plan and statistics live in HBM; the [K, 4] "
                            "result is what a planner reads back"},
            "gpu_launches": args.steps * (H * fan.launches_per_step() + 3) * world,
            "roofline": None, "cpu_baseline": None,
        }
        print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    D = Dist()
    try:
        if args.workload == "aisle":
            run_aisle(args, D)
        elif args.workload == "montecarlo":
            run_montecarlo(args, D)
        else:
            run_small_workload(args, D)
    finally:
        D.close()


if __name__ == "__main__":
    main()
