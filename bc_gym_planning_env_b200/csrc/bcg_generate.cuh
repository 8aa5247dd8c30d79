// Device-side AisleTurnEnv generation: the reference rebuilds a random aisle turn on
// the host at every reset (envs/synth_turn_env.py:278-291 -> :317-332 -> :110-192 -> cv2.line), ~1 k maps/s
// per core.  Here one warp regenerates one env inside its fixed-size slot of the map / tile / path arenas:
// geometry, 5 walls rasterised like cv2.line, refined path, chunk bounds.  Included by bcg_kernels.cu.
#pragma once
#include "bcg_device.cuh"

namespace bcg {

// uniform in [0, 1) from 53 bits of one Philox block half
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t env, uint64_t draw, uint32_t k) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)draw, 0x41495345u + (k >> 1), (uint32_t)(draw >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint64_t bits = (k & 1) ? (((uint64_t)r.z << 21) | (r.w >> 11)) : (((uint64_t)r.x << 21) | (r.y >> 11));
  return (double)bits * 0x1.0p-53;
}

// RandomAisleTurnEnv._draw_random_turn_params (envs/synth_turn_env.py:317-332), Philox instead of MT19937
__device__ __forceinline__ BcgTurnParams draw_turn_params(uint64_t seed, uint64_t env, uint64_t draw) {
  BcgTurnParams t;
  t.main_corridor_length = 10.0 + 6.0 * philox_uniform(seed, env, draw, 0);
  t.turn_corridor_length = 4.0 + 8.0 * philox_uniform(seed, env, draw, 1);
  const double lim = 3. / 8. * BCG_PI;
  t.turn_corridor_angle = -lim + (2 * lim) * philox_uniform(seed, env, draw, 2);
  t.main_corridor_width = 0.5 + philox_uniform(seed, env, draw, 3);
  t.turn_corridor_width = 0.5 + philox_uniform(seed, env, draw, 4);
  t.margin = 1.0;
  t.flip_arnd_oy = philox_uniform(seed, env, draw, 5) < 0.5;
  t.flip_arnd_ox = philox_uniform(seed, env, draw, 6) < 0.5;
  t.rot_theta = BCG_TWO_PI * philox_uniform(seed, env, draw, 7);
  return t;
}

struct AisleGeometry {
  double wall[5][4];    // world end points (x0, y0, x1, y1) of the walls a-i, c-d, d-e, j-g, g-h
  double way[4][3];     // oriented way points
  double origin_x, origin_y;
  int width, height;    // costmap cells
};

// path_and_costmap_from_config (envs/synth_turn_env.py:41-192): corridor corners, way points, flips, rotation,
// world bounds.  Same operation order as the host restatement in envs/synth_turn_env.py.
__device__ __forceinline__ AisleGeometry aisle_geometry(const BcgTurnParams& tp, double inv_res) {
  const double h = tp.main_corridor_length / 2, far = tp.turn_corridor_length / 2;
  const double alpha = tp.turn_corridor_angle, d = tp.main_corridor_width, z = tp.turn_corridor_width;
  const double ta = tan(alpha), ca = cos(alpha);
  const double lower = -z / ca, upper = z / ca;
  // a b c d e f g h i j
  const double cx[10] = {-d, 0, d, d, far, far, d, far, -d, d};
  const double cy[10] = {-h, -h, -h, d * ta + lower, far * ta + lower, far * ta, d * ta + upper, far * ta + upper, h, h};
  double sn, cs;
  sincos(tp.rot_theta, &sn, &cs);
  const double fx = tp.flip_arnd_oy ? -1. : 1., fy = tp.flip_arnd_ox ? -1. : 1.;
  const double t00 = cs * fx, t01 = -sn * fy, t10 = sn * fx, t11 = cs * fy;
  double mx[10], my[10];
  double min_x = 1e300, max_x = -1e300, min_y = 1e300, max_y = -1e300;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    mx[k] = t00 * cx[k] + t01 * cy[k];
    my[k] = t10 * cx[k] + t11 * cy[k];
    min_x = fmin(min_x, mx[k]); max_x = fmax(max_x, mx[k]);
    min_y = fmin(min_y, my[k]); max_y = fmax(max_y, my[k]);
  }
  AisleGeometry g;
  const int wa[5] = {0, 2, 3, 9, 6}, wb[5] = {8, 3, 4, 6, 7};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    g.wall[k][0] = mx[wa[k]]; g.wall[k][1] = my[wa[k]];
    g.wall[k][2] = mx[wb[k]]; g.wall[k][3] = my[wb[k]];
  }
  const double wx[4] = {0, 0, d, far * ca};
  const double wy[4] = {-h, d * ta + lower, d * ta, far * ca * ta};
  const double wt[4] = {BCG_PI / 2, BCG_PI / 2, alpha, alpha};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g.way[k][0] = t00 * wx[k] + t01 * wy[k];
    g.way[k][1] = t10 * wx[k] + t11 * wy[k];
    double t = wt[k];
    if (tp.flip_arnd_ox) t = -t;
    if (tp.flip_arnd_oy) t = BCG_PI - t;
    g.way[k][2] = py_mod(t + tp.rot_theta, BCG_TWO_PI);
  }
  const double size_x = fabs(max_x - min_x) + 2 * tp.margin, size_y = fabs(max_y - min_y) + 2 * tp.margin;
  g.origin_x = min_x - tp.margin;
  g.origin_y = min_y - tp.margin;
  g.width = (int)rint(size_x * inv_res);      // CostMap2D.create_empty -> world_to_pixel(size, (0, 0), res)
  g.height = (int)rint(size_y * inv_res);
  return g;
}

// cv2.line(thickness 1, 8-connected) = LineIterator(left to right): the major axis advances every step, the
// minor axis after step j has moved ceil((2 d j - D) / (2 D)) times (clamped at 0), D/d = major/minor extent.
// Verified pixel for pixel against cv2.line on the host (tests/test_host_logic.py).  Lanes take pixels
// j = lane, lane + 32, ...; `value` 254 draws, 0 erases; the lethal tile plane is kept in step.
__device__ __forceinline__ void draw_wall(uint8_t* data, uint32_t* tiles, int pitch, int tiles_x, int rows, int x0, int y0,
                                          int x1, int y1, uint8_t value, unsigned lane) {
  if (x1 < x0) {
    int t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
  }
  const int dx = x1 - x0, dy = y1 - y0;
  const int sy = dy >= 0 ? 1 : -1;
  const int adx = dx, ady = dy >= 0 ? dy : -dy;
  const bool steep = ady > adx;
  const int D = steep ? ady : adx, dm = steep ? adx : ady;
  for (int j = lane; j <= D; j += 32) {
    const long long t = 2ll * dm * j - D;
    const int m = t > 0 ? (int)((t + 2ll * D - 1) / (2ll * D)) : 0;
    const int x = steep ? x0 + m : x0 + j;
    const int y = steep ? y0 + j * sy : y0 + m * sy;
    if (x < 0 || y < 0 || x >= pitch || y >= rows) continue;      // never for aisle walls (1 m margin); be safe
    data[(int64_t)y * pitch + x] = value;
    uint32_t* word = tiles + (((int64_t)(y >> 4) * tiles_x + (x >> 5)) << 4) + (y & 15);
    if (value == 254) atomicOr(word, 1u << (x & 31));
    else atomicAnd(word, ~(1u << (x & 31)));
  }
}

}  // namespace bcg
