"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous index range.

Envs are independent (reference envs/base/env.py:225-249: each PlanEnv owns its robot, reward state
and queues; the costmap is read-only during step), so stepping needs NO collective.  torch.distributed
(NCCL over NVLink on the box, gloo in CPU tests) is used for exactly two things:
  * all-reduce(sum) of the small episode-statistics vector, once per report interval;
  * fan-out of `get_state` snapshots for Monte-Carlo rollouts (README.md:45-61 of the reference):
    broadcast the snapshot columns from one rank, every rank scatters them into its own envs.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total, rank=None, world_size=None):
    """Contiguous [lo, hi) range of global env ids owned by `rank`; sizes differ by at most one."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t):
    """In-place sum over ranks (no-op for a single process).  Returns t."""
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_episode_stats(env, reset=False):
    """Whole-job episode statistics: fp64 [STATS_WORDS] summed over ranks (see _native.STAT_NAMES)."""
    return allreduce_sum_(env.episode_stats(reset=reset))


def broadcast_snapshot(f, i, src=0):
    """Broadcast snapshot columns (fp64 [F, k], int32 [I, k]) from rank `src`; other ranks pass
    tensors of the right shape to be overwritten.  Returns (f, i)."""
    if world()[1] > 1:
        dist.broadcast(f, src=src)
        dist.broadcast(i, src=src)
    return f, i


def fan_out_columns(k_states, n_local, first_global_env=None):
    """Which snapshot column each local env starts from when k_states start states are replicated over all envs of
    all ranks: global env g takes state g % k_states.  `first_global_env` is the global id of this rank's env 0
    (VecPlanEnv's env_id_base, or shard_range(...)[0]); default: every rank holds n_local envs."""
    if first_global_env is None:
        first_global_env = world()[0] * int(n_local)
    return (torch.arange(n_local, dtype=torch.int64) + int(first_global_env)) % int(k_states)


def monte_carlo_rollouts(env, start, actions, reduce=True, cols=None):
    """Monte-Carlo evaluation of `start` states (README.md:45-61 of the reference, batched).

    env: a VecPlanEnv (auto_reset off) whose env e has the map and path of the start state it is assigned;
    start: VecState with k columns (already identical on every rank, e.g. via broadcast_snapshot);
    actions: [H, k, 2] fixed action sequence per start state.  Local env e is loaded with column cols[e] (default:
    fan_out_columns from the env's env_id_base, i.e. global env g takes state g % k) and stepped H times with its own
    Philox stream.
    Returns fp64 [k, 4]: (sum of returns, rollouts that collided, rollouts that finished their path, rollouts) per
    start state, summed over ranks when reduce."""
    if env.auto_reset:
        raise ValueError("Monte-Carlo rollouts need auto_reset off: a reset rollout would restart from the initial state")
    k = start.f.shape[1]
    if cols is None:
        cols = fan_out_columns(k, env.n_envs, int(env._c_params.env_id_base))
    cols = cols.to(env.device)
    if tuple(cols.shape) != (env.n_envs,) or int(cols.min()) < 0 or int(cols.max()) >= k:
        raise ValueError("cols must assign one of the %d start states to each of the %d envs" % (k, env.n_envs))
    if tuple(actions.shape[1:]) != (k, 2):
        raise ValueError("actions must have shape [H, %d, 2]" % k)
    env.set_state(type(start)(start.f.to(env.device)[:, cols].contiguous(), start.i.to(env.device)[:, cols].contiguous()))
    from bc_gym_planning_env_b200 import _native as nat
    acts = actions.to(env.device)[:, cols].contiguous()          # [H, n, 2]: one gather for the whole horizon
    # the rollout's return accumulates on the device in the episode-return row, started from zero (the same additions in
    # the same order as summing the step rewards); the H steps are launched from one library call
    env.state_f[nat.F_EP_RETURN].zero_()
    env.rollout(acts)
    ret = env.state_f[nat.F_EP_RETURN]
    out = torch.zeros((k, 4), dtype=torch.float64, device=env.device)
    out[:, 0].index_add_(0, cols, ret)
    out[:, 1].index_add_(0, cols, env.state_i[nat.I_COLLIDED].to(torch.float64))
    out[:, 2].index_add_(0, cols, env.goal_reached().to(torch.float64))
    out[:, 3].index_add_(0, cols, torch.ones_like(ret))
    return allreduce_sum_(out) if reduce else out
