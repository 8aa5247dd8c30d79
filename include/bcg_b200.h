/*
 * bcg_b200.h -- C-ABI of the B200-native batched PlanEnv.step path.
 *
 * The upstream reference (braincorp/bc-gym-planning-env, pure Python) has no FFI of its own; the
 * seams this library plugs into are (paths relative to bc_gym_planning_env/ in the reference):
 *   (i)  the env API            envs/base/env.py:278-361 (get_state/set_state/reset/step)
 *   (ii) the native-hook seam   `brain.shining_utils.*` imports at utilities/path_tools.py:101-103,
 *        utilities/coordinate_transformations.py:17-20,39-41,169-171, utilities/costmap_utils.py:106-107
 * Each entry point below names the reference function(s) it replaces for a *batch* of N environments.
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is a CUDA device pointer owned by the caller
 *     (the Python host allocates them as torch tensors); the library never allocates or frees them.
 *   - every function returns 0 on success, <0 on error; bcg_last_error() gives the message of the
 *     last failure on the calling thread.  No C++ exception crosses the boundary.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - per-env dynamic state is structure-of-arrays: state_f[row * n_envs + env] (fp64 rows) and
 *     state_i[row * n_envs + env] (int32 rows); row numbers come from bcg_state_layout().
 *   - kernels never trap: out-of-map pixels are skipped exactly like envs/base/env.py:483-484 and
 *     anomalies are counted in the device `status` words (BCG_STATUS_*), which the host polls.
 */
#ifndef BCG_B200_H_
#define BCG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCG_ABI_VERSION 11

/* error codes */
#define BCG_OK 0
#define BCG_ERR_INVALID (-1) /* bad argument                     */
#define BCG_ERR_CUDA (-2)    /* a CUDA runtime call failed       */
#define BCG_ERR_NO_DEVICE (-3)

/* robot kinds (robot_models/robot_drive_types.py) */
/* reward providers (envs/base/reward_provider_examples.py) */
#define BCG_REWARD_CONTINUOUS 0
#define BCG_REWARD_PURE_PURSUIT 1

#define BCG_ROBOT_TRICYCLE 0
#define BCG_ROBOT_DIFFDRIVE 1

/* fixed fp64 state rows; ring rows follow (see BcgStateLayout) */
#define BCG_F_ROBOT 0    /* 7 rows: true robot x,y,th,v,w,steer_cmd,wheel   (tricycle_model.py:234-244) */
#define BCG_F_DROBOT 7   /* 7 rows: delayed robot state = State.robot_state (env.py:386-396)            */
#define BCG_F_DPOSE 14   /* 3 rows: delayed pose       = State.pose         (env.py:377-394)            */
#define BCG_F_TIME 17    /* State.current_time                                                         */
#define BCG_F_MIN_DIST 18 /* ContinuousRewardProviderState.min_spat_dist_so_far (reward.py:15)          */
#define BCG_F_EP_RETURN 19 /* sum of rewards since the last reset (statistics only)                     */
#define BCG_F_FIXED 20
/* int32 state rows */
#define BCG_I_ITER 0     /* State.current_iter                                 */
#define BCG_I_TARGET 1   /* ContinuousRewardProviderState.target_idx            */
#define BCG_I_COLLIDED 2 /* State.robot_collided (sticky, env.py:393)           */
#define BCG_I_QC 3       /* control queue: head | len << 16                     */
#define BCG_I_QP 4       /* pose queue                                          */
#define BCG_I_QS 5       /* robot-state queue                                   */
#define BCG_I_FIXED 6

/* device status words (uint32 counters, length BCG_STATUS_WORDS) */
#define BCG_STATUS_LUT_MISS 0     /* footprint tuple not found in the angle-bin table */
#define BCG_STATUS_PATH_EXHAUSTED 1 /* init: "Goal pose too close to initial pose" (reward.py:275-277) */
#define BCG_STATUS_SLOT_OVERFLOW 2  /* bcg_generate_aisles: a drawn map / path did not fit its slot; env left as it was */
#define BCG_STATUS_SAMPLER_EMPTY 3  /* bcg_generate_minis: "the sampling space looks empty" (mini_env.py:138,361)      */
#define BCG_STATUS_WORDS 8

/* episode statistics accumulated on device at episode end (fp64, length BCG_STATS_WORDS) */
#define BCG_STAT_EPISODES 0
#define BCG_STAT_RETURN 1
#define BCG_STAT_LENGTH 2
#define BCG_STAT_COLLIDED 3
#define BCG_STAT_GOAL 4
#define BCG_STAT_TIMEOUT 5
#define BCG_STATS_WORDS 8

/* Environment constants shared by the batch: EnvParams (envs/base/params.py:15-43), RewardParams
 * (envs/base/reward.py:162-171), robot dimension table (robot_models/robot_dimensions_examples.py). */
typedef struct BcgParams {
  double dt;
  double resolution;
  double inv_resolution;      /* 1./resolution, as coordinate_transformations.py:204 computes it */
  double spatial_precision;   /* sp */
  double angular_precision;   /* ap */
  double progress_multiplier; /* RewardParams.spatial_progress_multiplier */
  double wheel_base;          /* front_wheel_from_axis */
  double max_wheel_angle;
  double max_wheel_delta;     /* max_front_wheel_speed * dt (tricycle_model.py:143) */
  double p_gain;
  double max_lin_acc;
  double max_ang_acc;
  double alpha[6];            /* odometry noise; used when noise_on */
  double ego_x0, ego_y0;      /* egocentric crop origin in the robot frame (egocentric.py:113-114) */
  double ego_world_w, ego_world_h; /* CostMap2D.world_size() of the crop (costmap_2d.py:115-121) */
  uint64_t seed;              /* Philox key */
  uint64_t env_id_base;       /* global id of env 0 of this shard (distinct noise streams per GPU) */
  int32_t robot_kind;
  int32_t noise_on;
  int32_t delay_control, delay_pose, delay_state;
  int32_t iteration_timeout;
  int32_t ego_w, ego_h;       /* crop size in pixels */
  int32_t reward_kind;        /* 0 ContinuousRewardProvider (reward.py:173-288), 1 ContinuousRewardPurePursuitProvider
                                 (reward.py:291-371) */
  int32_t reserved;
  int32_t auto_reset;         /* restore the initial state of envs that report done */
  int32_t ego_variant;        /* 0: EgocentricCostmap wrapper (egocentric.py:125-160): crop about the delayed pose,
                                 9-vector goal_n_state.  1: ColoredEgoCostmapRandomAisleTurnEnv
                                 (synth_turn_env.py:396-420): crop about the TRUE robot pose, goal = unit vector
                                 towards the last path point, 5 values [gx, gy, v, w, wheel] (slots 5..8 zero) */
} BcgParams;

/* one costmap of the arena: CostMap2D (utilities/costmap_2d.py:13-37) + its derived planes: the lethal
 * bit-plane (collision) and the cell tiles (the same uint8 cells re-laid out as 128-byte tiles of 16 px x 8
 * rows, tile (ty, tx) at cell_tile_off + (ty * ctiles_x + tx) * 128, row r of a tile at + r * 16, zero
 * beyond the map; the egocentric kernel fetches whole HBM lines of a rotated window from it) */
#define BCG_MAP_ONLY_LETHAL 1
typedef struct BcgMapDesc {
  int64_t data_off;  /* byte offset of uint8 [H][pitch] in the map arena                      */
  int64_t tile_off;  /* uint32 offset of the lethal tile plane in the tile arena               */
  double origin_x, origin_y;
  int32_t height, width, pitch;
  int32_t tiles_x, tiles_y; /* 32 px x 16 rows per 64-byte tile                                */
  int32_t flags;            /* BCG_MAP_ONLY_LETHAL: every non-zero cell is 254 (set by the host before
                               bcg_build_lethal_tiles, cleared there when another value is met)       */
  int64_t cell_tile_off;    /* byte offset of the map's cell tiles in the cell-tile arena (multiple of 128) */
  int32_t ctiles_x, ctiles_y; /* 16 px x 8 rows per 128-byte cell tile: pitch / 16, ceil(height / 8)     */
  int32_t occupied;         /* cells != 0, counted by bcg_build_lethal_tiles: maps with more than 1 cell in 20
                               occupied skip the sparse egocentric kernel (their windows overflow its list) */
  int32_t sum_off;          /* uint32 offset of the map's tile summary in BcgBatch.occ_sum_arena: [tiles_y][(tiles_x + 31) / 32]
                               words, bit tx & 31 of word [ty][tx >> 5] set when tile (tx, ty) of the occupancy plane
                               holds a cell (a clear bit promises an empty tile; set by the host, kept by the
                               generators)                                                                          */
} BcgMapDesc;

/* one refined path: fp64 SoA rows x,y,th,cos(th),sin(th), each `pitch` long, then chunk bounds */
typedef struct BcgPathDesc {
  int64_t off;       /* fp64 offset of the 5 point rows in the path arena                       */
  int64_t chunk_off; /* fp64 offset of the 3 chunk rows (cx, cy, radius), each `chunk_pitch`     */
  int32_t n;         /* number of points                                                       */
  int32_t pitch;
  int32_t n_chunks;  /* ceil(n / 32)                                                           */
  int32_t chunk_pitch;
} BcgPathDesc;

/* angle-binned footprint table replacing get_pixel_footprint (utilities/path_tools.py:122-162) */
typedef struct BcgFootprintLut {
  const double* edges;    /* device [n_bins + 1] ascending, edges[0] = -pi, edges[n_bins] = +pi    */
  const int16_t* verts;   /* device [n_bins][2 * n_verts]  rounded vertices (x0,y0,x1,y1,...)       */
  const int16_t* header;  /* device [n_bins][4] = xmin, ymin, n_rows, width (mask bounding box)     */
  const uint64_t* rows;   /* device [n_bins][max_rows][wpr] bit i of word j <-> column xmin+64j+i   */
  const double* fp_pix;   /* device [2 * n_verts] footprint / resolution (fx0,fy0,...)             */
  const int32_t* bucket_first; /* device [n_buckets]: bin containing the left end of uniform bucket k  */
  double bucket_scale;    /* n_buckets / (2 pi)                                                     */
  int32_t n_bins, n_verts, max_rows, wpr;
  int32_t n_buckets;
  int32_t bin_stride;     /* int16 elements per row of `bins` (multiple of 8), 0 when bins is NULL         */
  const int16_t* bins;    /* optional, device [n_bins][bin_stride]: a bin's vertex tuple followed by its header
                             (xmin, ymin, n_rows, width), 16-byte aligned rows -- lets the lookup fetch two
                             candidate bins, tuple and header, in one memory round trip                     */
} BcgFootprintLut;

/* BcgBatch.flags */
#define BCG_BATCH_SPARSE_EGO_ONLY 1 /* no map of the batch is dense (BcgMapDesc.occupied below 1 cell in 20, read back by the
                                       host after bcg_build_lethal_tiles; always true for device-generated worlds): the
                                       sparse egocentric kernel renders the rare window that overflows its cell list
                                       itself and the dense kernel is not launched behind it */

/* everything a step touches; all pointers are device pointers */
typedef struct BcgBatch {
  int32_t n_envs;
  int32_t n_frows, n_irows; /* must equal bcg_state_layout() for the params in use */
  int32_t n_maps, n_paths;
  int32_t flags;            /* BCG_BATCH_* */
  double* state_f;
  int32_t* state_i;
  double* init_f;  /* the state reset() restores (env.py:247,302) */
  int32_t* init_i;
  double* cand;    /* scratch [9][n_envs]: rows 0..2 = the pose the last step proposed (before the collision
                      verdict); bcg_kinematic_step writes rows 0..6 (the proposed robot state)              */
  int32_t* cand_i; /* scratch [2][n_envs] (reserved; must be allocated)                                      */
  void* ego_work;  /* scratch [n_envs][128 bytes]: per-env affine map + source window of the egocentric crop,
                      written by move_kernel / reward_kernel (or bcg_observe_ego) for the egocentric kernel; may
                      be NULL when no egocentric image is ever requested                                    */
  void* work;      /* scratch [n_envs][192 bytes]: the per-env record move_kernel leaves for reward_kernel, and the
                      work records (map / path / footprint references) of the stand-alone collision entry points */
  const int32_t* map_id;  /* [n_envs] index into maps  */
  const int32_t* path_id; /* [n_envs] index into paths */
  const BcgMapDesc* maps;
  const BcgPathDesc* paths;
  const uint8_t* map_arena;
  const uint32_t* tile_arena;
  const double* path_arena;
  BcgFootprintLut lut;
  const void* map_tmaps; /* optional: [n_maps][tmap_n_widths] 128-byte TMA tensor maps from
                            bcg_encode_map_tensor_maps; NULL = the egocentric kernel stages its source
                            window with plain coalesced loads                                             */
  int32_t tmap_n_widths;  /* 1..4 box-width classes, ascending                                           */
  int32_t tmap_box_h;     /* rows per box                                                                */
  int32_t tmap_box_w[4];  /* box widths (multiples of 16) the tensor maps were encoded with              */
  const uint8_t* cell_tile_arena; /* optional: cell tiles of every map (bcg_build_cell_tiles); when set the
                            egocentric kernel stages only the 128-byte tiles its rotated source window
                            touches (16-byte cp.async pieces, zero fill outside the map) and map_tmaps is
                            not used                                                                      */
  const uint32_t* occ_tile_arena; /* optional: the occupancy plane (1 bit per cell, value != 0), laid out exactly
                            like the lethal tile plane (same tile_off); filled by bcg_build_lethal_tiles.
                            With it, cell_tile_arena and ego_list the egocentric observation takes the sparse
                            path: zero the crop, then scatter only the occupied source cells of its window   */
  int32_t* ego_list;      /* optional scratch [n_envs + 4]: envs whose window holds too many occupied cells
                            for the sparse path, handed to the dense cell-tile kernel (count at [n_envs]; [n_envs + 1] is
                            the sparse kernel's env counter)                                                 */
  const uint32_t* occ_sum_arena; /* optional: one bit per tile of the occupancy plane (BcgMapDesc.sum_off), filled by
                            bcg_build_lethal_tiles and kept by the generators; must start out zero.  With it the sparse
                            egocentric kernel loads only the non-empty tiles of a window                              */
  uint32_t* status; /* [BCG_STATUS_WORDS] */
  double* stats;    /* [BCG_STATS_WORDS]  */
  uint64_t* step_counter; /* optional device [2], zero at start: when set, bcg_step ignores its step_index argument,
                            takes the step index from word 0 and increments it on the device when the step is done
                            (word 1 is its scratch) -- so that a captured CUDA graph of a step can be replayed
                            without changing a kernel argument                                                    */
} BcgBatch;

/* TurnParams (envs/synth_turn_env.py:18-31): geometry of one aisle turn */
typedef struct BcgTurnParams {
  double main_corridor_length, turn_corridor_length, turn_corridor_angle;
  double main_corridor_width, turn_corridor_width, margin, rot_theta;
  int32_t flip_arnd_oy, flip_arnd_ox;
} BcgTurnParams;

/* Fixed-size per-env slots for environments that are (re)generated on the device (bcg_generate_aisles).
 * Env e owns maps[e] and paths[e] (n_maps == n_paths == n_envs, map_id[e] == path_id[e] == e); the offsets in
 * its descriptors are set once by the host (e * slot size) and never change, the generator rewrites the
 * geometry fields (height, width, pitch, origin, tile counts, path length).  The map, cell-tile and lethal-tile
 * slots must be zero when an env is generated for the first time. */
typedef struct BcgAisleSlots {
  int64_t map_slot_bytes;    /* capacity of one env's uint8 rows, and of its cell tiles           */
  int64_t tile_slot_words;   /* capacity of one env's lethal tile plane (uint32 words)            */
  int32_t path_pitch;        /* row pitch of one env's 5 path rows (points); multiple of 4        */
  int32_t chunk_pitch;       /* row pitch of its 3 chunk rows; >= ceil(path_pitch / 32)           */
  void* gen_state;           /* device scratch [n_envs][128 bytes], zero before the first call:
                                the walls currently drawn in each slot (to erase them cheaply)     */
  BcgTurnParams* params_out; /* optional device [n_envs]: the turn each regenerated env now has   */
} BcgAisleSlots;

/* RandomMiniEnvParams (envs/mini_env.py:30-46) + the two goal tolerances of its EnvParams the sampler's final check
 * uses (:349-353) */
typedef struct BcgMiniGenParams {
  double inner_h, inner_w, mid_margin, out_margin;
  double min_obstacle_angle, max_obstacle_angle;
  double lim_euc_dist, lim_ang_dist, angular_pose_noise_scale;
  double goal_spat_dist, goal_ang_dist;
} BcgMiniGenParams;

/* MiniEnvParams (envs/mini_env.py:80-92): world size, start / end poses, the corner obstacle (two walls o-a, o-b) */
typedef struct BcgMiniParams {
  double h, w;
  double start[3], end[3];
  double a[2], o[2], b[2];
} BcgMiniParams;

typedef struct BcgStateLayout {
  int32_t n_frows, n_irows;
  int32_t ring_control; /* first fp64 row of the control ring: delay_control slots x 2 rows */
  int32_t ring_pose;    /* delay_pose slots x 3 rows                                         */
  int32_t ring_state;   /* delay_state slots x 7 rows                                        */
} BcgStateLayout;

/* outputs of one step; any pointer may be NULL to skip that output */
typedef struct BcgStepOut {
  double* reward;      /* [n] reward of this step (reward.py:214-259)                          */
  uint8_t* done;       /* [n] env.py:407-419                                                   */
  uint8_t* hit;        /* [n] collision verdict of this step's pose_collides (env.py:455)       */
  uint8_t* ego_image;  /* [n][ego_h][ego_w] egocentric crop (egocentric.py:125-160 'env')        */
  float* goal_n_state; /* [n][9] egocentric.py:152-159                                          */
  float* obs_vec;      /* [n][12] fp32 copy of delayed pose(3), delayed robot state(7), time, target_idx */
  /* compact form of ego_image for consumers behind a narrow link (optional; needs ego_image and the sparse kernel): per
   * env the list of its non-zero crop pixels, entry = pixel offset (v * ego_w + u) | cost value << 16, in no particular
   * order.  ego_hit_count[e] = number of entries (only the first ego_hit_cap are stored: a larger count means "read
   * ego_image[e]"), or -1 when the env was rendered by the dense kernel.  A crop is ~1 % walls, so this is ~0.8 KB per
   * env instead of 15.6 KB. */
  uint32_t* ego_hits;      /* [n][ego_hit_cap] */
  int32_t* ego_hit_count;  /* [n]              */
  int32_t ego_hit_cap;
  int32_t reserved;
} BcgStepOut;

/* -- library ------------------------------------------------------------------------------------ */
int bcg_abi_version(void);
/* copies the calling thread's last error message (NUL terminated) and returns its length */
size_t bcg_last_error(char* buf, size_t cap);
/* sizeof() of the ABI structs, in declaration order: 0 BcgParams, 1 BcgMapDesc, 2 BcgPathDesc,
 * 3 BcgFootprintLut, 4 BcgBatch, 5 BcgStateLayout, 6 BcgStepOut, 7 BcgTurnParams, 8 BcgAisleSlots, 9 BcgMiniGenParams,
 * 10 BcgMiniParams; -1 for anything else.  Lets a
 * binding (ctypes, cffi, ...) verify its struct mirrors before the first call. */
int64_t bcg_sizeof(int32_t which);
/* number of CUDA devices visible, <0 on error; makes a missing GPU a loud failure for callers */
int bcg_device_count(void);
/* state row map for the delays in `p` (mirrors the three queues of env.py:64-66) */
int bcg_state_layout(const BcgParams* p, BcgStateLayout* out);

/* -- setup (at reset time, not per step) ------------------------------------------------------------ */
/* derive the lethal bit-plane (cell == 254, costmap_2d.py:21) of maps [first, first+count) */
int bcg_build_lethal_tiles(const BcgBatch* b, int32_t first, int32_t count, void* stream);
/* derive the cell tiles (see BcgMapDesc) of maps [first, first+count) from the uint8 rows (CostMap2D.get_data,
 * costmap_2d.py:91-96: the bytes extract_egocentric_costmap samples, costmap_utils.py:25-75) */
int bcg_build_cell_tiles(const BcgBatch* b, int32_t first, int32_t count, void* stream);
/* Host-side: encode n_widths CUtensorMaps (128 B each, cuTensorMapEncodeTiled) per costmap -- uint8
 * [H][pitch] tensor, boxes (box_w[j], box_h), zero fill outside the map (= cv2.warpAffine's
 * borderValue 0, utilities/costmap_utils.py:72) -- into out_host [n_maps][n_widths][128 bytes].  The
 * caller uploads the buffer and points BcgBatch.map_tmaps at it.  Widths must be ascending multiples
 * of 16, and widths and box_h <= 256. */
int bcg_encode_map_tensor_maps(const BcgMapDesc* maps_host, int32_t n_maps, const void* map_arena_dev,
                               const int32_t* box_w, int32_t n_widths, int32_t box_h, void* out_host);
/* make_initial_state (env.py:179-214) + ContinuousRewardProvider.generate_initial_state
 * (reward.py:261-288) for every env: writes init_f/init_i and copies them into state_f/state_i */
int bcg_init_state(const BcgParams* p, const BcgBatch* b, void* stream);
/* PlanEnv.reset (env.py:293-303) for envs with mask[e] != 0 (mask NULL = all) */
int bcg_reset_where(const BcgBatch* b, const uint8_t* mask, void* stream);

/* RandomAisleTurnEnv.reset with draw_new_turn_on_reset (envs/synth_turn_env.py:278-291) on the device, for
 * envs with mask[e] != 0 (mask NULL = all): draw a turn (:317-332; Philox keyed (p->seed; env id, draw_index)
 * unless `turn_params`, device [n_envs], supplies them), build its walls like cv2.line and its way points
 * (path_and_costmap_from_config :110-192), refine the path (utilities/path_tools.py:178-240, spacing
 * `path_delta`), derive both tile planes, and make the initial state (env.py:179-214, reward.py:261-288).
 * The previous walls of the slot are erased pixel by pixel, so a reset costs O(wall pixels), not O(map).
 * Not available together with TMA tensor maps (map_tmaps), which cannot be re-encoded on the device. */
int bcg_generate_aisles(const BcgParams* p, const BcgBatch* b, const BcgAisleSlots* slots, const uint8_t* mask,
                        const BcgTurnParams* turn_params, uint64_t draw_index, double path_delta, void* stream);

/* RandomMiniEnv's env construction (envs/mini_env.py:323-389, what its reset repeats) on the device, for envs with
 * mask[e] != 0 (mask NULL = all), in the same per-env slots as bcg_generate_aisles: sample MiniEnvParams (:269-320,
 * Philox keyed (p->seed; env id, draw_index) in the reference's draw order) until neither end pose collides and the
 * two are not within the goal tolerances of each other (:336-358), rasterise the two walls like cv2.line incl. its
 * clipping of far end points (cv::clipLine), refine the path, make the initial state.  With `mini_params` (device
 * [n_envs]) the worlds are built from those parameters as they are, like MiniEnv(config) does.  `params_out`
 * (optional, device [n_envs]) receives the accepted parameters.  BCG_STATUS_SLOT_OVERFLOW counts worlds that do not
 * fit their slots, BCG_STATUS_SAMPLER_EMPTY envs whose sampling ran out of tries (both left as they were). */
int bcg_generate_minis(const BcgParams* p, const BcgBatch* b, const BcgAisleSlots* slots, const uint8_t* mask,
                       const BcgMiniGenParams* gen, const BcgMiniParams* mini_params, BcgMiniParams* params_out,
                       uint64_t draw_index, double path_delta, void* stream);

/* -- the hot path -------------------------------------------------------------------------------- */
/* PlanEnv.step (env.py:334-361) for all envs.  actions: device [n][2] (wheel_v, wheel_angle) or
 * (v, w) for diff-drive; action_is_f64 selects fp64 vs fp32 elements.  Launches move_kernel (one thread per env:
 * control delay env.py:371-373, robot model tricycle_model.py:478-538 / differential_drive.py:236-265, pose_collides
 * env.py:464-489, rollback :452-461, delay lines :377-389, compact observation, egocentric record), reward_kernel (a
 * group of lanes per env: reward.py:214-259, done env.py:400-419, statistics, goal vector egocentric.py:152-159,
 * auto-reset env.py:293-303) and -- only if out->ego_image is set -- the egocentric observation (egocentric.py:125-160):
 * with the occupancy plane and ego_list the sparse scatter kernel, followed by the dense cell-tile kernel for the
 * envs it hands over unless BCG_BATCH_SPARSE_EGO_ONLY, else one dense kernel.  Each kernel is launched as a programmatic
 * dependent of the one before it on `stream`.  step_index is the caller's global step counter (ignored when
 * BcgBatch.step_counter is set): odometry noise is Philox4x32-10 keyed by p->seed at counter (env id, step_index, draw),
 * replacing the reference's global np.random (differential_drive.py:50). */
int bcg_step(const BcgParams* p, const BcgBatch* b, const void* actions, int32_t action_is_f64,
             uint64_t step_index, const BcgStepOut* out, void* stream);

/* bcg_step with per-kernel hooks: events[0..4] are caller-created cudaEvent_t handles recorded on `stream`: [0], [1]
 * before move_kernel, [2] between move_kernel and reward_kernel, [3] after reward_kernel (reward / done / compact
 * observation are final: a host consumer's copies can start here while the egocentric kernel runs), [4] after the
 * egocentric kernels.  Any entry may be NULL (not recorded).  A kernel with an event recorded right before it is not
 * launched as a programmatic dependent of its predecessor.  events == NULL behaves exactly like bcg_step. */
int bcg_step_events(const BcgParams* p, const BcgBatch* b, const void* actions, int32_t action_is_f64,
                    uint64_t step_index, const BcgStepOut* out, void* const* events, void* stream);

/* the pieces, individually addressable (used by tests, ncu and the hook seam) */
/* robot.step (tricycle_model.py:478-538 / differential_drive.py:236-265) into b->cand */
int bcg_kinematic_step(const BcgParams* p, const BcgBatch* b, const void* actions,
                       int32_t action_is_f64, uint64_t step_index, void* stream);
/* pose_collides (env.py:464-489) of poses [3][n] (rows x,y,th) against each env's map, using the
 * lethal tile plane; flags_out [n].  pixels_out (optional, [n]) = in-map footprint pixel count (warp-per-env
 * kernel; without it the thread-per-env kernel the step itself uses).  poses == NULL checks the poses the last
 * bcg_step / bcg_kinematic_step proposed (rows 0..2 of b->cand).  Two launches: per-env work records (footprint
 * angle bin, mask box, map references) into b->work, then the collision kernel. */
int bcg_collision(const BcgParams* p, const BcgBatch* b, const double* poses, uint8_t* flags_out,
                  int32_t* pixels_out, void* stream);
/* same verdicts read straight from the uint8 costmap rows (no derived plane) */
int bcg_collision_u8(const BcgParams* p, const BcgBatch* b, const double* poses, uint8_t* flags_out,
                     void* stream);
/* only the collision kernel, on the work records the last bcg_collision / bcg_collision_u8 call left in b->work
 * (use_u8: the uint8-row kernel) -- the form the collision roofline is measured on */
int bcg_collision_recheck(const BcgParams* p, const BcgBatch* b, uint8_t* flags_out, int32_t use_u8, void* stream);
/* EgocentricCostmap.observation (egocentric.py:125-160) from the current state */
int bcg_observe_ego(const BcgParams* p, const BcgBatch* b, uint8_t* ego_image, float* goal_n_state,
                    void* stream);

/* -- snapshots: get_state / set_state (env.py:278-291) ------------------------------------------ */
/* gather columns idx[0..k) of state_f/state_i into out_f [n_frows][k], out_i [n_irows][k] */
int bcg_gather_state(const BcgBatch* b, const int64_t* idx, int32_t k, double* out_f,
                     int32_t* out_i, void* stream);
/* scatter columns back; load_delayed_robot != 0 applies env.py:284 (robot := delayed robot state) */
int bcg_scatter_state(const BcgBatch* b, const int64_t* idx, int32_t k, const double* in_f,
                      const int32_t* in_i, int32_t load_delayed_robot, void* stream);

/* -- compact observation ------------------------------------------------------------------------------ */
/* Packs the per-env hit lists a step left in BcgStepOut.ego_hits (fixed stride `cap`) into one contiguous array:
 * env e's min(count, cap) entries go to packed[offsets[e] ..], offsets = exclusive prefix sum of the counts with
 * dense / overflowed envs (count < 0 or > cap) counted as 0 -- what a host consumer copies over the link instead
 * of n crops of ego_w x ego_h bytes.  offsets_only == 0: packed is uint32 (pixel offset | value << 16);
 * offsets_only != 0: packed is uint16, the pixel offsets alone (for batches whose maps hold no cost value but 254,
 * BCG_MAP_ONLY_LETHAL: the value is implied). */
int bcg_pack_ego_hits(const uint32_t* hits, const int32_t* counts, int32_t cap, int32_t n, const int64_t* offsets,
                      void* packed, int32_t offsets_only, void* stream);

/* -- hook-seam helpers (batched forms of the brain.shining_utils.* scalars) --------------------- */
/* world_to_pixel (coordinate_transformations.py:185-205): xy [n][2] fp64 -> int32 [n][2] */
int bcg_world_to_pixel(const double* xy, int64_t n, double origin_x, double origin_y,
                       double resolution, int32_t* out, void* stream);
/* normalize_angle (coordinate_transformations.py:28-36) */
int bcg_normalize_angle(const double* in, int64_t n, double* out, void* stream);
/* is_footprint_colliding_impl (costmap_utils.py:106-136; contract test_costmap_utils.py:251-325): does any cell of the
 * blitted footprint equal `value`?  values, mask: device [n] (the image slice and the blit mask, flattened);
 * flag_out: device int32, 1 when any(values[mask != 0] == value) */
int bcg_masked_any_equal(const uint8_t* values, const uint8_t* mask, int64_t n, int32_t value, int32_t* flag_out, void* stream);
/* inverse_transform_2d_impl (coordinate_transformations.py:39-84): device [n][3] (x, y, angle) -> [n][3] */
int bcg_inverse_transform(const double* transforms, int64_t n, double* out, void* stream);
/* native_project_poses (coordinate_transformations.py:289-328): poses device [n][3] moved by the HOST transform
 * (x, y, angle): rotate, translate, wrap the angle */
int bcg_project_poses(const double* transform_host, const double* poses, int64_t n, double* out, void* stream);
/* Observation.path of every env (env.py:421-433: path[target_idx:]; pure pursuit: path[:target_idx + 1]) as a device
 * tensor in the robot frame (from_global_to_egocentric, coordinate_transformations.py:341-362, about the observed pose):
 * out device [n][max_points][3] zero padded, len_out (optional) device [n] = way points left (may exceed max_points) */
int bcg_observe_ego_path(const BcgParams* p, const BcgBatch* b, int32_t max_points, double* out, int32_t* len_out, void* stream);

/* H steps of a fixed action plan (Monte-Carlo fan-outs: the reference steps a copied env through a candidate plan,
 * README.md:45-61): bcg_step for plan[h], h = 0 .. horizon - 1, launched back to back from one call -- every kernel a
 * programmatic dependent of the one before, no host work in between.  plan: device float32 [horizon][n_envs][2];
 * step h uses the Philox step index step_index + h (or the device-side counter when BcgBatch.step_counter is set). */
int bcg_rollout(const BcgParams* p, const BcgBatch* b, const float* plan, int32_t horizon, uint64_t step_index,
                const BcgStepOut* out, void* stream);

/* -- image memory --------------------------------------------------------------------------------
 * Egocentric crops are mostly zeros.  A device allocation with compute-data compression (CUDA virtual memory
 * management, CU_MEM_ALLOCATION_COMP_GENERIC) makes the hardware write and read such lines at a fraction of their DRAM
 * cost; contents and addressing are unchanged for kernels and copies.  The one place where this library allocates: the
 * caller owns the block and returns it with bcg_free_image_memory.  (Nothing in the reference corresponds: its images
 * are NumPy arrays, egocentric.py:125-160.)
 * bytes: wanted size; *dptr: device address; *mapped_bytes: size actually mapped (granularity rounded), to be passed
 * back on free; *compressed: 1 when the allocation got compression, 0 when the device / driver gave a plain one. */
int bcg_alloc_image_memory(int64_t bytes, int32_t want_compression, void** dptr, int64_t* mapped_bytes, int32_t* compressed);
int bcg_free_image_memory(void* dptr, int64_t mapped_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BCG_B200_H_ */
