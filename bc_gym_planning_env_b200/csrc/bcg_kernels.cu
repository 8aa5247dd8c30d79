// Kernels + C-ABI of the batched PlanEnv.step path for B200 (sm_100a).  See include/bcg_b200.h.
//
// Launch shapes (DESIGN.md has the rooflines):
//   kin_kernel      1 thread / env   -- delay ring + robot model + Philox noise, coalesced SoA fp64
//   commit_kernel   1 warp   / env   -- footprint-vs-lethal-tile collision, rollback, delay rings,
//                                       chunk-culled reached-index scan, reward, done, auto-reset
//   ego_kernel      1 CTA    / env   -- cv2.warpAffine(INTER_NEAREST)-exact egocentric gather
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "bcg_b200.h"
#include "bcg_device.cuh"

using namespace bcg;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define BCG_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t err__ = (expr);                                                               \
    if (err__ != cudaSuccess)                                                                 \
      return fail(BCG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));       \
  } while (0)

#define BCG_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return fail(BCG_ERR_INVALID, msg); \
  } while (0)

BcgStateLayout make_layout(const BcgParams& p) {
  BcgStateLayout L;
  L.ring_control = BCG_F_FIXED;
  L.ring_pose = L.ring_control + 2 * p.delay_control;
  L.ring_state = L.ring_pose + 3 * p.delay_pose;
  L.n_frows = L.ring_state + 7 * p.delay_state;
  L.n_irows = BCG_I_FIXED;
  return L;
}

int check_batch(const BcgParams* p, const BcgBatch* b) {
  BCG_REQUIRE(p && b, "null params/batch");
  BCG_REQUIRE(b->n_envs > 0, "n_envs must be positive");
  BCG_REQUIRE(p->delay_control >= 0 && p->delay_pose >= 0 && p->delay_state >= 0, "negative delay");
  BCG_REQUIRE(p->delay_control < 4096 && p->delay_pose < 4096 && p->delay_state < 4096, "delay too large");
  const BcgStateLayout L = make_layout(*p);
  BCG_REQUIRE(b->n_frows == L.n_frows && b->n_irows == L.n_irows, "state row count does not match bcg_state_layout");
  BCG_REQUIRE(b->state_f && b->state_i && b->init_f && b->init_i && b->cand, "null state pointers");
  BCG_REQUIRE(b->map_id && b->path_id && b->maps && b->paths && b->map_arena && b->tile_arena && b->path_arena,
              "null arena pointers");
  BCG_REQUIRE(b->lut.edges && b->lut.verts && b->lut.header && b->lut.rows && b->lut.fp_pix, "null footprint table");
  BCG_REQUIRE(b->lut.n_verts > 0 && b->lut.n_verts <= 32, "footprint must have 1..32 vertices");
  BCG_REQUIRE(b->lut.n_bins > 0 && b->lut.wpr > 0 && b->lut.max_rows > 0, "empty footprint table");
  BCG_REQUIRE(b->status && b->stats, "null status/stats");
  return BCG_OK;
}

// ---- kernels -------------------------------------------------------------------------------------

// robot.step for every env: envs/base/env.py:371-373 (control delay) + robot model into b.cand.
__global__ void __launch_bounds__(256) kin_kernel(const BcgParams p, const BcgBatch b, const BcgStateLayout L,
                                                  const void* __restrict__ actions, const int action_is_f64,
                                                  const uint64_t step_index) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.n_envs) return;
  const int64_t N = b.n_envs;
  double u[2];
  if (action_is_f64) {
    const double2 a = reinterpret_cast<const double2*>(actions)[e];
    u[0] = a.x;
    u[1] = a.y;
  } else {
    const float2 a = reinterpret_cast<const float2*>(actions)[e];
    u[0] = (double)a.x;
    u[1] = (double)a.y;
  }
  if (p.delay_control > 0) {
    int q = b.state_i[BCG_I_QC * N + e];
    delay_line<2>(b.state_f + (int64_t)L.ring_control * N + e, N, q, p.delay_control, u);
    b.state_i[BCG_I_QC * N + e] = q;
  }
  double s[7];
#pragma unroll
  for (int r = 0; r < 7; ++r) s[r] = b.state_f[(BCG_F_ROBOT + r) * N + e];
  robot_step(s, u[0], u[1], p, p.env_id_base + (uint64_t)e, step_index);
#pragma unroll
  for (int r = 0; r < 7; ++r) b.cand[r * N + e] = s[r];
}

// Collision + the rest of _resolve_state_transition (env.py:363-398) + reward (reward.py:214-259) +
// done (env.py:407-419); one warp per env, scalars are computed warp-uniformly and stored by lane 0.
__global__ void __launch_bounds__(256) commit_kernel(const BcgParams p, const BcgBatch b, const BcgStateLayout L,
                                                     const BcgStepOut out) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (e >= b.n_envs) return;
  const int64_t N = b.n_envs;
  double* sf = b.state_f + e;
  int32_t* si = b.state_i + e;

  double c[7];
#pragma unroll
  for (int r = 0; r < 7; ++r) c[r] = b.cand[r * N + e];
  const BcgMapDesc m = b.maps[b.map_id[e]];
  const BcgPathDesc pd = b.paths[b.path_id[e]];

  const bool hit = collide_tiles<false>(p, b, m, c[0], c[1], c[2], lane, nullptr);
  if (hit) {  // env.py:458-459 + tricycle_model.py:471-476: pose restored, v = w = 0, wheel/steer kept
    c[0] = sf[(BCG_F_ROBOT + 0) * N];
    c[1] = sf[(BCG_F_ROBOT + 1) * N];
    c[2] = sf[(BCG_F_ROBOT + 2) * N];
    c[3] = 0.0;
    c[4] = 0.0;
  }
  int iter = si[BCG_I_ITER * N];
  int target = si[BCG_I_TARGET * N];
  int collided = si[BCG_I_COLLIDED * N];
  double min_dist = sf[BCG_F_MIN_DIST * N];
  const bool done_before = (target > pd.n - 1) || (iter >= p.iteration_timeout) || (collided != 0);

  // delay lines (env.py:377-389).  All lanes compute the same values; only lane 0 writes.
  double dpose[3] = {c[0], c[1], c[2]};
  double dstate[7];
#pragma unroll
  for (int r = 0; r < 7; ++r) dstate[r] = c[r];
  int qp = si[BCG_I_QP * N], qs = si[BCG_I_QS * N];
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 7; ++r) sf[(BCG_F_ROBOT + r) * N] = c[r];
    delay_line<3>(sf + (int64_t)L.ring_pose * N, N, qp, p.delay_pose, dpose);
    delay_line<7>(sf + (int64_t)L.ring_state * N, N, qs, p.delay_state, dstate);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) dpose[r] = __shfl_sync(BCG_FULL, dpose[r], 0);
#pragma unroll
  for (int r = 0; r < 7; ++r) dstate[r] = __shfl_sync(BCG_FULL, dstate[r], 0);

  const double time = sf[BCG_F_TIME * N] + p.dt;
  iter += 1;
  collided |= hit ? 1 : 0;

  // reward on the *delayed* pose (reward.py:227-232 reads state.pose)
  double reward = 0.0;
  const double* P = b.path_arena + pd.off;
  if (!(target > pd.n - 1)) {
    const int last = last_reached_from(p, b, pd, target, dpose[0], dpose[1], dpose[2], lane);
    if (last >= target) {
      target = last + 1;
      if (target > pd.n - 1) {
        min_dist = 0.0;
      } else {
        min_dist = hypot(__ldg(P + target) - dpose[0], __ldg(P + pd.pitch + target) - dpose[1]);
      }
      reward = 1.0;
    } else {
      const double d = hypot(__ldg(P + target) - dpose[0], __ldg(P + pd.pitch + target) - dpose[1]);
      if (d < min_dist) {
        reward = (min_dist - d) * p.progress_multiplier;
        min_dist = d;
      }
    }
  }
  const bool goal = target > pd.n - 1;
  const bool timed_out = iter >= p.iteration_timeout;
  const bool done = goal || timed_out || (collided != 0);
  const double ep_return = sf[BCG_F_EP_RETURN * N] + reward;
  __syncwarp();  // every lane has read the old state; lane 0 (or the reset loop) may now overwrite it

  if (lane == 0) {
    if (out.reward) out.reward[e] = reward;
    if (out.done) out.done[e] = done ? 1 : 0;
    if (out.hit) out.hit[e] = hit ? 1 : 0;
    if (done && !done_before) {  // episode statistics, once per episode
      atomicAdd(b.stats + BCG_STAT_EPISODES, 1.0);
      atomicAdd(b.stats + BCG_STAT_RETURN, ep_return);
      atomicAdd(b.stats + BCG_STAT_LENGTH, (double)iter);
      if (collided) atomicAdd(b.stats + BCG_STAT_COLLIDED, 1.0);
      if (goal) atomicAdd(b.stats + BCG_STAT_GOAL, 1.0);
      if (timed_out) atomicAdd(b.stats + BCG_STAT_TIMEOUT, 1.0);
    }
  }
  if (done && p.auto_reset) {
    // PlanEnv.reset (env.py:293-303): every row back to the stored initial state
    for (int r = lane; r < L.n_frows; r += 32) sf[(int64_t)r * N] = b.init_f[(int64_t)r * N + e];
    for (int r = lane; r < L.n_irows; r += 32) si[(int64_t)r * N] = b.init_i[(int64_t)r * N + e];
  } else if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 7; ++r) sf[(BCG_F_DROBOT + r) * N] = dstate[r];
#pragma unroll
    for (int r = 0; r < 3; ++r) sf[(BCG_F_DPOSE + r) * N] = dpose[r];
    sf[BCG_F_TIME * N] = time;
    sf[BCG_F_MIN_DIST * N] = min_dist;
    sf[BCG_F_EP_RETURN * N] = ep_return;
    si[BCG_I_ITER * N] = iter;
    si[BCG_I_TARGET * N] = target;
    si[BCG_I_COLLIDED * N] = collided;
    si[BCG_I_QP * N] = qp;
    si[BCG_I_QS * N] = qs;
  }
  if (out.obs_vec) {
    __syncwarp();
    if (lane < 12) {
      float v;
      if (lane < 3) v = (float)sf[(BCG_F_DPOSE + lane) * N];
      else if (lane < 10) v = (float)sf[(BCG_F_DROBOT + lane - 3) * N];
      else if (lane == 10) v = (float)sf[BCG_F_TIME * N];
      else v = (float)si[BCG_I_TARGET * N];
      out.obs_vec[(int64_t)e * 12 + lane] = v;
    }
  }
}

// make_initial_state (env.py:179-214) + generate_initial_state (reward.py:261-288); one warp per env.
__global__ void __launch_bounds__(256) init_kernel(const BcgParams p, const BcgBatch b, const BcgStateLayout L) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (e >= b.n_envs) return;
  const int64_t N = b.n_envs;
  const BcgPathDesc pd = b.paths[b.path_id[e]];
  const double* P = b.path_arena + pd.off;
  const double x0 = P[0], y0 = P[pd.pitch], t0 = P[2 * pd.pitch];
  const int last = last_reached_from(p, b, pd, 0, x0, y0, t0, lane);
  int target = last + 1;
  double min_dist = 0.0;
  if (target > pd.n - 1) {
    if (lane == 0) atomicAdd(b.status + BCG_STATUS_PATH_EXHAUSTED, 1u);
  } else {
    min_dist = hypot(P[target] - x0, P[pd.pitch + target] - y0);
  }
  for (int r = lane; r < L.n_frows; r += 32) {
    double v = 0.0;
    if (r == BCG_F_ROBOT + 0 || r == BCG_F_DROBOT + 0 || r == BCG_F_DPOSE + 0) v = x0;
    if (r == BCG_F_ROBOT + 1 || r == BCG_F_DROBOT + 1 || r == BCG_F_DPOSE + 1) v = y0;
    if (r == BCG_F_ROBOT + 2 || r == BCG_F_DROBOT + 2 || r == BCG_F_DPOSE + 2) v = t0;
    if (r == BCG_F_MIN_DIST) v = min_dist;
    b.init_f[(int64_t)r * N + e] = v;
    b.state_f[(int64_t)r * N + e] = v;
  }
  for (int r = lane; r < L.n_irows; r += 32) {
    const int v = (r == BCG_I_TARGET) ? target : 0;
    b.init_i[(int64_t)r * N + e] = v;
    b.state_i[(int64_t)r * N + e] = v;
  }
}

__global__ void __launch_bounds__(256) reset_kernel(const BcgBatch b, const uint8_t* __restrict__ mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.n_envs) return;
  if (mask && !mask[e]) return;
  const int64_t N = b.n_envs;
  for (int r = 0; r < b.n_frows; ++r) b.state_f[(int64_t)r * N + e] = b.init_f[(int64_t)r * N + e];
  for (int r = 0; r < b.n_irows; ++r) b.state_i[(int64_t)r * N + e] = b.init_i[(int64_t)r * N + e];
}

// lethal bit-plane: bit x&31 of word ((ty*tiles_x + tx)*16 + (y&15)) <-> costmap[y][x] == 254
__global__ void __launch_bounds__(256) tiles_kernel(const BcgBatch b, const int first) {
  const BcgMapDesc m = b.maps[first + blockIdx.y];
  const int words = m.tiles_x * m.tiles_y * 16;
  uint32_t* dst = const_cast<uint32_t*>(b.tile_arena) + m.tile_off;
  const uint8_t* src = b.map_arena + m.data_off;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
    const int tile = w >> 4, r = w & 15;
    const int ty = tile / m.tiles_x, tx = tile - ty * m.tiles_x;
    const int y = (ty << 4) + r;
    uint32_t bits = 0;
    if (y < m.height) {
      const uint4* row = reinterpret_cast<const uint4*>(src + (int64_t)y * m.pitch + (tx << 5));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 v = row[h];
        const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t eq = __vcmpeq4(ws[k], 0xFEFEFEFEu) & 0x01010101u;  // bit 0 of each byte
          const uint32_t nib = (eq * 0x10204080u) >> 28;                    // gather to 4 bits
          bits |= nib << (h * 16 + k * 4);
        }
      }
      const int over = (tx << 5) + 32 - m.width;
      if (over > 0) bits &= (over >= 32) ? 0u : (0xffffffffu >> over);
    }
    dst[w] = bits;
  }
}

__global__ void __launch_bounds__(256) collision_kernel(const BcgParams p, const BcgBatch b,
                                                        const double* __restrict__ poses, uint8_t* __restrict__ flags,
                                                        int32_t* __restrict__ pixels, const int use_u8) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (e >= b.n_envs) return;
  const int64_t N = b.n_envs;
  const BcgMapDesc m = b.maps[b.map_id[e]];
  const double x = poses[e], y = poses[N + e], th = poses[2 * N + e];
  bool hit;
  int cnt = 0;
  if (use_u8) {
    hit = collide_u8(p, b, m, x, y, th, lane);
  } else if (pixels) {
    hit = collide_tiles<true>(p, b, m, x, y, th, lane, &cnt);
  } else {
    hit = collide_tiles<false>(p, b, m, x, y, th, lane, nullptr);
  }
  if (lane == 0) {
    flags[e] = hit ? 1 : 0;
    if (pixels) pixels[e] = cnt;
  }
}

// cv2::saturate_cast<int>(double) == cvRound with saturation
__device__ __forceinline__ int cv_round_sat(double v) {
  const double r = rint(v);
  if (r >= 2147483647.0) return 2147483647;
  if (r <= -2147483648.0) return (-2147483647 - 1);
  return (int)r;
}

#define BCG_EGO_MAX 256

// EgocentricCostmap.observation (envs/egocentric.py:125-160): extract_egocentric_costmap
// (utilities/costmap_utils.py:25-75) = cv2.getRotationMatrix2D composed in float32 with the crop
// shift, then cv2.warpAffine(INTER_NEAREST, borderValue=0): fp64 inverse and 10-bit fixed-point source
// coordinates; plus the 9-vector goal_n_state (:152-159).
__global__ void __launch_bounds__(256) ego_kernel(const BcgParams p, const BcgBatch b, uint8_t* __restrict__ image,
                                                  float* __restrict__ goal_n_state) {
  __shared__ int adx[BCG_EGO_MAX], ady[BCG_EGO_MAX], bdx[BCG_EGO_MAX], bdy[BCG_EGO_MAX];
  const int e = blockIdx.x;
  const int64_t N = b.n_envs;
  const double* sf = b.state_f + e;
  const BcgMapDesc m = b.maps[b.map_id[e]];
  const double px = sf[(BCG_F_DPOSE + 0) * N], py = sf[(BCG_F_DPOSE + 1) * N], pth = sf[(BCG_F_DPOSE + 2) * N];

  if (image) {
    // costmap_utils.py:42-65
    const double cx = (double)world_to_pixel_1d(px, m.origin_x, p.inv_resolution);
    const double cy = (double)world_to_pixel_1d(py, m.origin_y, p.inv_resolution);
    const double deg = 180 * pth / BCG_PI;
    const double rad = deg * (BCG_PI / 180.);
    double bsn, acs;
    sincos(rad, &bsn, &acs);
    const float r00 = (float)acs, r01 = (float)bsn, r02 = (float)((1 - acs) * cx - bsn * cy);
    const float r10 = (float)(-bsn), r11 = (float)acs, r12 = (float)(bsn * cx + (1 - acs) * cy);
    const double dsx = rint((p.ego_x0 - (m.origin_x - px)) * p.inv_resolution);
    const double dsy = rint((p.ego_y0 - (m.origin_y - py)) * p.inv_resolution);
    const double M0 = (double)r00, M1 = (double)r01, M2 = (double)(r02 - (float)dsx);
    const double M3 = (double)r10, M4 = (double)r11, M5 = (double)(r12 - (float)dsy);
    // cv::warpAffine inverts the forward map in double
    double D = M0 * M4 - M1 * M3;
    D = (D != 0.0) ? 1. / D : 0.0;
    const double A11 = M4 * D, A22 = M0 * D, A12 = M1 * (-D), A21 = M3 * (-D);
    const double b1 = -A11 * M2 - A12 * M5;
    const double b2 = -A21 * M2 - A22 * M5;
    for (int t = threadIdx.x; t < p.ego_w; t += blockDim.x) {
      adx[t] = cv_round_sat(A11 * t * 1024);
      ady[t] = cv_round_sat(A21 * t * 1024);
    }
    for (int t = threadIdx.x; t < p.ego_h; t += blockDim.x) {
      bdx[t] = cv_round_sat((A12 * t + b1) * 1024) + 512;
      bdy[t] = cv_round_sat((A22 * t + b2) * 1024) + 512;
    }
    __syncthreads();
    const uint8_t* src = b.map_arena + m.data_off;
    const int npx = p.ego_w * p.ego_h;
    uint8_t* dst = image + (int64_t)e * npx;
    for (int i = threadIdx.x; i < npx; i += blockDim.x) {
      const int v = i / p.ego_w, u = i - v * p.ego_w;
      const int X = (adx[u] + bdx[v]) >> 10, Y = (ady[u] + bdy[v]) >> 10;
      uint8_t val = 0;
      if ((unsigned)X < (unsigned)m.width && (unsigned)Y < (unsigned)m.height) val = __ldg(src + (int64_t)Y * m.pitch + X);
      dst[i] = val;
    }
  }
  if (goal_n_state && threadIdx.x == 0) {
    float* g = goal_n_state + (int64_t)e * 9;
    const BcgPathDesc pd = b.paths[b.path_id[e]];
    const int target = b.state_i[BCG_I_TARGET * N + e];
    if (target > pd.n - 1) {
#pragma unroll
      for (int k = 0; k < 9; ++k) g[k] = 0.f;
    } else {
      // from_global_to_egocentric (coordinate_transformations.py:341-362 -> :57-84 -> :310-328)
      const double* P = b.path_arena + pd.off;
      const double gx = P[target], gy = P[pd.pitch + target], gt = P[2 * pd.pitch + target];
      double sn, cs;
      sincos(pth, &sn, &cs);
      const double tx = -px * cs - py * sn;
      const double ty = px * sn - py * cs;
      const double tt = wrap_angle(-pth);
      double st, ct;
      sincos(tt, &st, &ct);
      const double ex = ct * gx - st * gy + tx;
      const double ey = st * gx + ct * gy + ty;
      const double ea = wrap_angle(gt + tt);
      g[0] = (float)clampd(ex / p.ego_world_w, -1.0, 1.0);
      g[1] = (float)clampd(ey / p.ego_world_h, -1.0, 1.0);
      g[2] = (float)ea;
      g[3] = (float)sf[(BCG_F_DROBOT + 0) * N];
      g[4] = (float)sf[(BCG_F_DROBOT + 1) * N];
      g[5] = (float)sf[(BCG_F_DROBOT + 2) * N];
      g[6] = (float)sf[(BCG_F_DROBOT + 3) * N];
      g[7] = (float)sf[(BCG_F_DROBOT + 4) * N];
      g[8] = (float)sf[(BCG_F_DROBOT + 6) * N];
    }
  }
}

__global__ void __launch_bounds__(256) gather_kernel(const BcgBatch b, const int64_t* __restrict__ idx, const int k,
                                                     double* __restrict__ out_f, int32_t* __restrict__ out_i) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const int64_t e = idx[j], N = b.n_envs;
  if (e < 0 || e >= N) return;
  for (int r = 0; r < b.n_frows; ++r) out_f[(int64_t)r * k + j] = b.state_f[(int64_t)r * N + e];
  for (int r = 0; r < b.n_irows; ++r) out_i[(int64_t)r * k + j] = b.state_i[(int64_t)r * N + e];
}

__global__ void __launch_bounds__(256) scatter_kernel(const BcgBatch b, const int64_t* __restrict__ idx, const int k,
                                                      const double* __restrict__ in_f, const int32_t* __restrict__ in_i,
                                                      const int load_delayed_robot) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const int64_t e = idx[j], N = b.n_envs;
  if (e < 0 || e >= N) return;
  for (int r = 0; r < b.n_frows; ++r) {
    int src = r;
    if (load_delayed_robot && r >= BCG_F_ROBOT && r < BCG_F_ROBOT + 7) src = r - BCG_F_ROBOT + BCG_F_DROBOT;  // env.py:284
    b.state_f[(int64_t)r * N + e] = in_f[(int64_t)src * k + j];
  }
  for (int r = 0; r < b.n_irows; ++r) b.state_i[(int64_t)r * N + e] = in_i[(int64_t)r * k + j];
}

__global__ void w2p_kernel(const double* __restrict__ xy, const int64_t n, const double ox, const double oy,
                           const double inv_res, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[2 * i] = world_to_pixel_1d(xy[2 * i], ox, inv_res);
  out[2 * i + 1] = world_to_pixel_1d(xy[2 * i + 1], oy, inv_res);
}

__global__ void wrap_kernel(const double* __restrict__ in, const int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = wrap_angle(in[i]);
}

inline int blocks_for(int64_t work, int per_block) { return (int)((work + per_block - 1) / per_block); }

}  // namespace

// ---- C-ABI -------------------------------------------------------------------------------------------
extern "C" {

int bcg_abi_version(void) { return BCG_ABI_VERSION; }

size_t bcg_last_error(char* buf, size_t cap) {
  if (buf && cap) {
    const size_t n = g_last_error.size() < cap - 1 ? g_last_error.size() : cap - 1;
    memcpy(buf, g_last_error.data(), n);
    buf[n] = 0;
  }
  return g_last_error.size();
}

int64_t bcg_sizeof(int32_t which) {
  switch (which) {
    case 0: return sizeof(BcgParams);
    case 1: return sizeof(BcgMapDesc);
    case 2: return sizeof(BcgPathDesc);
    case 3: return sizeof(BcgFootprintLut);
    case 4: return sizeof(BcgBatch);
    case 5: return sizeof(BcgStateLayout);
    case 6: return sizeof(BcgStepOut);
    default: return -1;
  }
}

int bcg_device_count(void) {
  int n = 0;
  const cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(BCG_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
  }
  return n;
}

int bcg_state_layout(const BcgParams* p, BcgStateLayout* out) {
  BCG_REQUIRE(p && out, "null argument");
  BCG_REQUIRE(p->delay_control >= 0 && p->delay_pose >= 0 && p->delay_state >= 0, "negative delay");
  *out = make_layout(*p);
  return BCG_OK;
}

int bcg_build_lethal_tiles(const BcgBatch* b, int32_t first, int32_t count, void* stream) {
  BCG_REQUIRE(b && b->maps && b->map_arena && b->tile_arena, "null map arena");
  BCG_REQUIRE(first >= 0 && count >= 0 && first + count <= b->n_maps, "map range out of bounds");
  cudaStream_t s = (cudaStream_t)stream;
  for (int32_t done = 0; done < count; done += 32768) {
    const int32_t chunk = (count - done) < 32768 ? (count - done) : 32768;
    tiles_kernel<<<dim3(8, chunk), 256, 0, s>>>(*b, first + done);
    BCG_CHECK_CUDA(cudaGetLastError());
  }
  return BCG_OK;
}

int bcg_init_state(const BcgParams* p, const BcgBatch* b, void* stream) {
  if (int rc = check_batch(p, b)) return rc;
  init_kernel<<<blocks_for((int64_t)b->n_envs * 32, 256), 256, 0, (cudaStream_t)stream>>>(*p, *b, make_layout(*p));
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_reset_where(const BcgBatch* b, const uint8_t* mask, void* stream) {
  BCG_REQUIRE(b && b->state_f && b->init_f && b->state_i && b->init_i && b->n_envs > 0, "bad batch");
  reset_kernel<<<blocks_for(b->n_envs, 256), 256, 0, (cudaStream_t)stream>>>(*b, mask);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_kinematic_step(const BcgParams* p, const BcgBatch* b, const void* actions, int32_t action_is_f64,
                       uint64_t step_index, void* stream) {
  if (int rc = check_batch(p, b)) return rc;
  BCG_REQUIRE(actions, "null actions");
  kin_kernel<<<blocks_for(b->n_envs, 256), 256, 0, (cudaStream_t)stream>>>(*p, *b, make_layout(*p), actions,
                                                                          action_is_f64, step_index);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_observe_ego(const BcgParams* p, const BcgBatch* b, uint8_t* ego_image, float* goal_n_state, void* stream) {
  if (int rc = check_batch(p, b)) return rc;
  BCG_REQUIRE(ego_image || goal_n_state, "nothing to compute");
  BCG_REQUIRE(p->ego_w > 0 && p->ego_h > 0 && p->ego_w <= BCG_EGO_MAX && p->ego_h <= BCG_EGO_MAX,
              "egocentric crop must be 1..256 pixels per side");
  ego_kernel<<<b->n_envs, ego_image ? 256 : 32, 0, (cudaStream_t)stream>>>(*p, *b, ego_image, goal_n_state);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_step(const BcgParams* p, const BcgBatch* b, const void* actions, int32_t action_is_f64, uint64_t step_index,
             const BcgStepOut* out, void* stream) {
  if (int rc = check_batch(p, b)) return rc;
  BCG_REQUIRE(actions && out, "null actions/out");
  cudaStream_t s = (cudaStream_t)stream;
  const BcgStateLayout L = make_layout(*p);
  kin_kernel<<<blocks_for(b->n_envs, 256), 256, 0, s>>>(*p, *b, L, actions, action_is_f64, step_index);
  BCG_CHECK_CUDA(cudaGetLastError());
  commit_kernel<<<blocks_for((int64_t)b->n_envs * 32, 256), 256, 0, s>>>(*p, *b, L, *out);
  BCG_CHECK_CUDA(cudaGetLastError());
  if (out->ego_image || out->goal_n_state) return bcg_observe_ego(p, b, out->ego_image, out->goal_n_state, stream);
  return BCG_OK;
}

int bcg_collision(const BcgParams* p, const BcgBatch* b, const double* poses, uint8_t* flags_out, int32_t* pixels_out,
                  void* stream) {
  if (int rc = check_batch(p, b)) return rc;
  BCG_REQUIRE(poses && flags_out, "null poses/flags");
  collision_kernel<<<blocks_for((int64_t)b->n_envs * 32, 256), 256, 0, (cudaStream_t)stream>>>(*p, *b, poses, flags_out,
                                                                                              pixels_out, 0);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_collision_u8(const BcgParams* p, const BcgBatch* b, const double* poses, uint8_t* flags_out, void* stream) {
  if (int rc = check_batch(p, b)) return rc;
  BCG_REQUIRE(poses && flags_out, "null poses/flags");
  collision_kernel<<<blocks_for((int64_t)b->n_envs * 32, 256), 256, 0, (cudaStream_t)stream>>>(*p, *b, poses, flags_out,
                                                                                              nullptr, 1);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_gather_state(const BcgBatch* b, const int64_t* idx, int32_t k, double* out_f, int32_t* out_i, void* stream) {
  BCG_REQUIRE(b && idx && out_f && out_i && k >= 0, "bad gather arguments");
  if (k == 0) return BCG_OK;
  gather_kernel<<<blocks_for(k, 256), 256, 0, (cudaStream_t)stream>>>(*b, idx, k, out_f, out_i);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_scatter_state(const BcgBatch* b, const int64_t* idx, int32_t k, const double* in_f, const int32_t* in_i,
                      int32_t load_delayed_robot, void* stream) {
  BCG_REQUIRE(b && idx && in_f && in_i && k >= 0, "bad scatter arguments");
  if (k == 0) return BCG_OK;
  scatter_kernel<<<blocks_for(k, 256), 256, 0, (cudaStream_t)stream>>>(*b, idx, k, in_f, in_i, load_delayed_robot);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_world_to_pixel(const double* xy, int64_t n, double origin_x, double origin_y, double resolution, int32_t* out,
                       void* stream) {
  BCG_REQUIRE(xy && out && n >= 0 && resolution > 0, "bad world_to_pixel arguments");
  if (n == 0) return BCG_OK;
  w2p_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(xy, n, origin_x, origin_y, 1. / resolution, out);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

int bcg_normalize_angle(const double* in, int64_t n, double* out, void* stream) {
  BCG_REQUIRE(in && out && n >= 0, "bad normalize_angle arguments");
  if (n == 0) return BCG_OK;
  wrap_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, out);
  BCG_CHECK_CUDA(cudaGetLastError());
  return BCG_OK;
}

}  // extern "C"
