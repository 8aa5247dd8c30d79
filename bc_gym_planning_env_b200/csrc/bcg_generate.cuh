// Device-side AisleTurnEnv generation.  The reference rebuilds a random aisle turn on the
// host at every reset (envs/synth_turn_env.py:278-291 -> :317-332 -> :110-192 -> cv2.line), about 1 k maps/s per
// core.  Here one CTA regenerates one env inside its fixed-size slots of the map / tile / path arenas: geometry,
// five walls rasterised like cv2.line (previous walls erased pixel by pixel), both tile planes kept in step,
// refined path, chunk bounds and the initial state.  Included by bcg_kernels.cu.
#pragma once
#include "bcg_device.cuh"

namespace bcg {

// uniform in [0, 1) from 53 bits of one Philox block half; draws 2 k and 2 k + 1 share a block
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t env, uint64_t draw, uint32_t k) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)draw, 0x41495345u + (k >> 1), (uint32_t)(draw >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint64_t bits = (k & 1) ? (((uint64_t)r.z << 21) | (r.w >> 11)) : (((uint64_t)r.x << 21) | (r.y >> 11));
  return (double)bits * 0x1.0p-53;
}

// RandomAisleTurnEnv._draw_random_turn_params (envs/synth_turn_env.py:317-332): same distributions, same draw order,
// Philox instead of MT19937
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
// world bounds.  np.dot of the 2 x 2 transform with a point accumulates fused (BLAS), hence the fma.
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
    mx[k] = fma(t01, cy[k], t00 * cx[k]);
    my[k] = fma(t11, cy[k], t10 * cx[k]);
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
    g.way[k][0] = fma(t01, wy[k], t00 * wx[k]);
    g.way[k][1] = fma(t11, wy[k], t10 * wx[k]);
    double t = wt[k];
    if (tp.flip_arnd_ox) t = -t;
    if (tp.flip_arnd_oy) t = BCG_PI - t;
    g.way[k][2] = py_mod(t + tp.rot_theta, BCG_TWO_PI);
  }
  const double size_x = fabs(max_x - min_x) + 2 * tp.margin, size_y = fabs(max_y - min_y) + 2 * tp.margin;
  g.origin_x = min_x - tp.margin;
  g.origin_y = min_y - tp.margin;
  g.width = world_to_pixel_1d(size_x, 0.0, inv_res);      // CostMap2D.create_empty -> world_to_pixel(size, (0, 0), res)
  g.height = world_to_pixel_1d(size_y, 0.0, inv_res);
  return g;
}

// layout of one env's map as the drawing code needs it
struct MapLayout {
  uint8_t* data;       // uint8 rows
  uint32_t* tiles;     // lethal tile plane
  uint32_t* occ;       // occupancy plane (same layout; may be null)
  uint8_t* ctiles;     // cell tiles
  int pitch, rows, tiles_x, ctiles_x;
};

// cv2.line(thickness 1, 8-connected) = LineIterator(left to right): the major axis advances every step, the
// minor axis after step j has moved ceil((2 d j - D) / (2 D)) times (clamped at 0), D/d = major/minor extent --
// Bresenham with the initial error D - 2 d in closed form, pixel for pixel equal to cv2.line
// (tests/test_host_logic.py::test_line_pixels_match_cv2).  Threads take pixels j = first, first + stride, ...;
// `value` 254 draws, 0 erases; the lethal and occupancy tile planes and the cell tiles are kept in step.
__device__ __forceinline__ void draw_wall(const MapLayout& m, int x0, int y0, int x1, int y1, uint8_t value, int first,
                                          int stride) {
  if (x1 < x0) {
    int t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
  }
  const int dx = x1 - x0, dy = y1 - y0;
  const int sy = dy >= 0 ? 1 : -1;
  const int ady = dy >= 0 ? dy : -dy;
  const bool steep = ady > dx;
  const int D = steep ? ady : dx, dm = steep ? dx : ady;
  for (int j = first; j <= D; j += stride) {
    const long long t = 2ll * dm * j - D;
    const int mv = t > 0 ? (int)((t + 2ll * D - 1) / (2ll * D)) : 0;
    const int x = steep ? x0 + mv : x0 + j;
    const int y = steep ? y0 + j * sy : y0 + mv * sy;
    if (x < 0 || y < 0 || x >= m.pitch || y >= m.rows) continue;      // never for aisle walls (1 m margin); be safe
    m.data[(int64_t)y * m.pitch + x] = value;
    m.ctiles[(((int64_t)(y >> 3) * m.ctiles_x + (x >> 4)) << 7) + ((y & 7) << 4) + (x & 15)] = value;
    const int64_t widx = (((int64_t)(y >> 4) * m.tiles_x + (x >> 5)) << 4) + (y & 15);
    if (value == 254) {
      atomicOr(m.tiles + widx, 1u << (x & 31));
      if (m.occ) atomicOr(m.occ + widx, 1u << (x & 31));
    } else {
      atomicAnd(m.tiles + widx, ~(1u << (x & 31)));
      if (m.occ) atomicAnd(m.occ + widx, ~(1u << (x & 31)));
    }
  }
}

// what a slot currently holds (BcgAisleSlots.gen_state, 128 bytes per env)
struct __align__(16) AisleGenState {
  int32_t wall[5][4];   // pixel end points of the walls drawn
  int32_t pitch, rows, tiles_x, ctiles_x;
  int32_t valid;
  int32_t pad[7];
};
static_assert(sizeof(AisleGenState) == 128, "AisleGenState is 128 bytes");

// refine_path (utilities/path_tools.py:178-240, angle_delta None) of the 4 way points: a segment longer than delta
// becomes int(d / delta) + 2 np.linspace points (k * step + start, the end point dropped) with the angle of its
// start; the last way point closes the path.
struct RefinedShape {
  int count[3];   // points segment i contributes
  int n;          // total, including the closing point
};

__device__ __forceinline__ RefinedShape refined_shape(const double way[4][3], double delta) {
  RefinedShape s;
  s.n = 1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double dx = way[i + 1][0] - way[i][0], dy = way[i + 1][1] - way[i][1];
    const double d = sqrt(dx * dx + dy * dy);                 // np.linalg.norm(axis=1)
    s.count[i] = d > delta ? (int)(d / delta) + 1 : 1;       // (int(d / delta) + 2 points)[:-1]
    s.n += s.count[i];
  }
  return s;
}

__device__ __forceinline__ void refined_point(const double way[4][3], const RefinedShape& s, int idx, double& x, double& y,
                                              double& th) {
  int i = 0, k = idx;
  while (i < 3 && k >= s.count[i]) {
    k -= s.count[i];
    ++i;
  }
  if (i == 3) {
    x = way[3][0]; y = way[3][1]; th = way[3][2];
    return;
  }
  th = way[i][2];
  if (s.count[i] == 1) {
    x = way[i][0]; y = way[i][1];
    return;
  }
  const double div = (double)s.count[i];                      // num - 1
  const double sx = (way[i + 1][0] - way[i][0]) / div, sy = (way[i + 1][1] - way[i][1]) / div;
  x = (double)k * sx + way[i][0];
  y = (double)k * sy + way[i][1];
}

}  // namespace bcg
