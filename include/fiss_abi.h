/* fiss_abi.h -- C ABI of libfissgpu.so: the B200 (sm_100a) Frenet lattice engine.
 *
 * The reference (SS47816/fiss_plus_planner, pure Python) has no FFI layer; its boundary for this
 * path is the planner class API.  Each entry point below names the reference
 * code it replaces (paths relative to the reference root).  Plain pointers and sizes only; no
 * C++/torch types; every call returns an int32 status (FISS_OK or a negative FISS_ERR_*) and
 * never lets a C++ exception escape.  A handle is bound to one device, is not thread-safe, and
 * owns: the uploaded spline and obstacle tables, pinned host staging and device scratch for the
 * *_host entry points.  The *_dev entry points work on caller-owned device buffers (e.g. torch
 * tensors' data_ptr()) and are asynchronous on `stream` (a cudaStream_t passed as void*).
 *
 * Candidate numbering: candidate (b, c) = problem b in [0, B), end state c in [0, C); flat id
 * b*C + c.  For FrenetOptimalPlanner the host enumerates end states in the reference's loop order
 * d (outer), T, v (inner) (frenet_optimal_planner.py:75,78,89), so "last minimal cost wins"
 * (:263-268) is "largest flat id among the minima".
 */
#ifndef FISS_ABI_H_
#define FISS_ABI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fiss_handle fiss_handle;

enum {
  FISS_OK = 0,
  FISS_ERR_INVALID = -1,  /* bad argument */
  FISS_ERR_CUDA = -2,     /* a CUDA runtime call failed; see fiss_last_error */
  FISS_ERR_CAPACITY = -3, /* request exceeds what one launch supports */
  FISS_ERR_STATE = -4,    /* spline not set, lane busy, no communicator, ... */
  FISS_ERR_NCCL = -5      /* libnccl could not be loaded or an NCCL call failed; see fiss_last_error */
};

/* per-candidate flags word written by the device */
#define FISS_FLAG_SPEED 1u      /* any(s_d > max_speed)        frenet_optimal_planner.py:152 */
#define FISS_FLAG_ACCEL 2u      /* any(|s_dd| > max_accel)     frenet_optimal_planner.py:155 */
#define FISS_FLAG_CURVATURE 4u  /* any(|c| > max_curvature); only if check_curvature (off in the reference, :145-150) */
#define FISS_FLAG_COLLISION 8u  /* has_collision()             frenet_optimal_planner.py:168-195 */
#define FISS_FLAG_INFEASIBLE_MASK 15u
#define FISS_FLAG_NCART_SHIFT 8 /* bits 8..23: n' = number of Cartesian points (truncation, :112-113) */
#define FISS_FLAG_NCART_MASK 0xFFFFu

/* rows of a "full record" [FISS_REC_ROWS][n_stride] (FrenetTrajectory fields, frenet.py:131-148) */
enum {
  FISS_REC_T = 0, FISS_REC_S, FISS_REC_S_D, FISS_REC_S_DD, FISS_REC_S_DDD,
  FISS_REC_D, FISS_REC_D_D, FISS_REC_D_DD, FISS_REC_D_DDD,
  FISS_REC_X, FISS_REC_Y, FISS_REC_YAW, FISS_REC_DS, FISS_REC_C, FISS_REC_C_D, FISS_REC_C_DD,
  FISS_REC_ROWS
};

/* rows of the dense materialisation [FISS_MAT_ROWS][B*C][n_stride]: the output contract
 * (x, y, yaw, v = s_d, kappa = c) of SURVEY 8(a) row a7 */
enum { FISS_MAT_X = 0, FISS_MAT_Y, FISS_MAT_YAW, FISS_MAT_V, FISS_MAT_KAPPA, FISS_MAT_ROWS };

/* One POD per launch: the planner settings, vehicle limits and cost weights the path reads. */
typedef struct fiss_params {
  double tick_t;           /* settings.tick_t                          frenet_optimal_planner.py:41 */
  double target_speed;     /* settings.highest_speed (= max_target_speed, :250) -> cost target */
  double max_speed;        /* vehicle.max_speed                        vehicle.py:39 */
  double max_accel;        /* vehicle.max_accel                        vehicle.py:40 */
  double max_curvature;    /* vehicle.max_curvature; used iff check_curvature */
  double ego_length;       /* vehicle.l                                vehicle.py:17 */
  double ego_width;        /* vehicle.w                                vehicle.py:18 */
  double cost_time_offset; /* the literal 10.0 of cost_function.py:42 */
  double w_speed;          /* w_V = 1      cost_function.py:9  */
  double w_accel;          /* w_A = 0.1    cost_function.py:10 */
  double w_jerk;           /* w_J = 0.1    cost_function.py:11 */
  double w_offset;         /* w_LC = 10    cost_function.py:12 */
  int32_t time_step_now;   /* plan(..., time_step_now)                 frenet_optimal_planner.py:247 */
  int32_t check_res;       /* has_collision check_res; the planners pass 2 (:202) */
  int32_t check_curvature; /* 0 = reference behaviour */
  int32_t collide_all;     /* 0: collision only for constraint survivors (plan(), :257-259); 1: for every candidate */
} fiss_params;

/* The product lattice the FOP / FOP+ / FISS planners sample: every (d_end[i], v_end[j], T[k]).
 * Candidate c = i*stride_d + j*stride_v + k*stride_t, a dense numbering of [0, nd*nv*nt):
 *   FrenetOptimalPlanner order (d outer, T, v inner; frenet_optimal_planner.py:75,78,89):
 *       stride_d = nt*nv, stride_t = nv, stride_v = 1
 *   FissPlanner grid [i_d][j_v][k_t] (fiss_planner.py:48,60,70):
 *       stride_d = nv*nt, stride_v = nt, stride_t = 1
 * The three axis arrays are HOST pointers (<= FISS_GRID_AXIS_MAX entries each); the step count of
 * horizon T[k] is fiss_arange_len(T[k], tick_t). */
#define FISS_GRID_AXIS_MAX 64
typedef struct fiss_grid {
  const double* d_end;
  const double* v_end;
  const double* T;
  int32_t nd, nv, nt;
  int32_t stride_d, stride_v, stride_t;
} fiss_grid;

/* ---- lifetime / errors ------------------------------------------------------------------- */
int32_t fiss_create(int32_t device, fiss_handle** out);
int32_t fiss_destroy(fiss_handle* h);
/* message of the last failing call on `h` (or of the last failing fiss_create when h == NULL) */
const char* fiss_last_error(const fiss_handle* h);
int32_t fiss_abi_version(void);

/* len(np.arange(0.0, T, tick)): the step count n of a candidate with horizon T
 * (frenet_optimal_planner.py:82, fiss_planner.py:113, fiss_plus_planner.py:181). */
int32_t fiss_arange_len(double T, double tick);

/* ---- scene tables (host pointers in, copied to handle-owned device memory) ----------------- */
/* table: [9][K] float64 rows = knots s_k, then a,b,c,d of x(s), then a,b,c,d of y(s)
 * (CubicSpline2D.sx / .sy, cubic_spline.py:162-168; b and d have K-1 live entries).
 * Replaces the per-step CubicSpline2D.calc_position / calc_yaw calls (:170-190,214-232). */
int32_t fiss_set_spline(fiss_handle* h, void* stream, const double* table, int32_t K);

/* Fit the natural cubic splines of L centre lines ON THE DEVICE: xy [L][K][2] way points ->
 * arc-length knots (np.cumsum of segment lengths) and the a, b, c, d rows of x(s), y(s) -- CubicSpline2D.__init__ /
 * CubicSpline1D.__init__ (cubic_spline.py:19-43,118-142,157-168), with the dense K x K np.linalg.solve replaced by
 * the tridiagonal (Thomas) recurrence, one CTA per lane.  tables [L][9][K] (host, may be NULL) receives the
 * coefficient tables; install_lane >= 0 makes that lane the handle's reference line (as fiss_set_spline would),
 * -1 installs nothing.  Agrees with the host fit to rounding (~1e-15 relative), not bit for bit. */
int32_t fiss_fit_splines_host(fiss_handle* h, void* stream, const double* xy, int32_t L, int32_t K, double* tables,
                              int32_t install_lane);

/* The 0.1 m polyline of generate_frenet_frame (frenet_optimal_planner.py:274-278): ref [m][4] = (x, y, yaw,
 * curvature) of the handle's reference line at s_i = i * step, i < m (CubicSpline2D.calc_position / calc_yaw /
 * calc_curvature, cubic_spline.py:170-232); rows with s_i beyond the last knot are NaN. */
int32_t fiss_frame_samples_host(fiss_handle* h, void* stream, double step, int32_t m, double* ref);

/* Dense obstacle predictions: xyth [M][T_obs][3] = (x, y, orientation) at absolute time step t,
 * lw [M][2] = rectangle (length, width), valid [M][T_obs] != 0 where obstacle.state_at_time(t)
 * is not None, final_time_step = obstacles[0].prediction.final_time_step
 * (frenet_optimal_planner.py:173,185-189).  M == 0 clears the table (no obstacles: :170-171). */
int32_t fiss_set_obstacles(fiss_handle* h, void* stream, const double* xyth, const double* lw,
                           const uint8_t* valid, int32_t M, int32_t T_obs, int32_t final_time_step);

/* Same table from the Waymo wire format, with the semantics of convert_waymo_obstacle_to_cr (waymo_interface.py:24-76,146):
 * trajs [N][T][11] float32 = (x, y, z, l, w, h, heading, vx, vy, valid, type), mask [N][T] (host pointers).
 * Agent i becomes an obstacle iff mask[i][1] is set (its trajectory, steps 1.. up to the first masked step, :45-54, is
 * non-empty, :58); a kept agent has states at step 0 (the initial state, never masked, :36-40) through the last unmasked
 * step of that run and none behind; a dropped agent has no state at all and the agents behind it move up.  Length /
 * width from step 0 (:33-34).  final_time_step < 0 derives obstacles[0].prediction.final_time_step
 * (frenet_optimal_planner.py:173) from the FIRST KEPT agent, as the reference's list would; >= 0 overrides it.
 * n_kept (may be NULL) receives the number of obstacles kept; 0 clears the table (no obstacles: :170-171). */
int32_t fiss_set_obstacles_waymo(fiss_handle* h, void* stream, const float* trajs, const uint8_t* mask,
                                 int32_t N, int32_t T, int32_t final_time_step, int32_t* n_kept);

/* ---- device-pointer API ------------------------------------------------------------------- */
/* The hot kernel: every candidate (b, c) through quintic/quartic solve, n-step evaluation, cost
 * (calc_frenet_paths :69-104 + cost_total), Frenet->Cartesian (calc_global_paths :106-138),
 * constraint masks (:140-160) and collision mask (:168-208).
 *   d_ego  [B][6]  (s, s_d, s_dd, d, d_d, d_dd)        FrenetState fields read at :81,92
 *   d_end  [C][4]  (d_end, v_end, T, n) ; n = fiss_arange_len(T, tick_t) stored as a double
 *   d_cost [B*C], d_flags [B*C]
 *   d_mat  optional (may be NULL): [FISS_MAT_ROWS][B*C][n_stride] float64, NaN beyond each row's length
 */
int32_t fiss_eval_candidates_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B,
                                 const double* d_end, int32_t C, const fiss_params* p,
                                 double* d_cost, uint32_t* d_flags, double* d_mat, int32_t n_stride);

/* The same outputs for a product lattice, by the lattice kernel (csrc/fiss_grid_kernel.cuh): the
 * work shared between candidates -- the longitudinal polynomial, spline frame, masks and cost terms
 * per (v_end, T); the lateral polynomial and cost terms per (d_end, T); obstacle proximity per
 * (v_end, T, step) -- is computed once per ego state in shared memory, then one warp per candidate
 * does the Frenet->Cartesian conversion, heading/curvature, collision predicate and the stores.
 * This is what FrenetOptimalPlanner.plan() (:247-270) launches.  Same d_cost / d_flags / d_mat
 * layout and numbering as fiss_eval_candidates_dev with the expanded [C][4] table. */
int32_t fiss_eval_grid_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B, const fiss_grid* g,
                           const fiss_params* p, double* d_cost, uint32_t* d_flags, double* d_mat,
                           int32_t n_stride);

/* argmin with the reference's tie rule (:263-268: `min_cost >= cost` => last minimum wins) over the
 * feasible candidates of each problem, then the winner's full record.
 *   d_best_idx [B] (-1 when nothing survives), d_best_cost [B],
 *   d_records optional [B][FISS_REC_ROWS][n_stride], d_best_meta optional [B][2] = (n, n') */
int32_t fiss_pick_winners_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B,
                              const double* d_end, int32_t C, const fiss_params* p,
                              const double* d_cost, const uint32_t* d_flags,
                              int32_t* d_best_idx, double* d_best_cost, double* d_records,
                              int32_t* d_best_meta, int32_t n_stride);

/* Full records (all FrenetTrajectory arrays) of selected candidates of ONE problem:
 * generate_trajectory / generate_trajectory_by_end_state + calc_global_paths for a list
 * (fiss_planner.py:101-138, fiss_plus_planner.py:172-205).  d_sel [N] indexes d_end; NULL = 0..N-1.
 *   d_records [N][FISS_REC_ROWS][n_stride], d_cost [N], d_flags [N] (masks included) */
int32_t fiss_full_records_dev(fiss_handle* h, void* stream, const double* d_ego6, const double* d_end,
                              const int32_t* d_sel, int32_t N, const fiss_params* p,
                              double* d_records, double* d_cost, uint32_t* d_flags, int32_t n_stride);

/* One plan step on caller-owned DEVICE buffers: fiss_eval_grid_dev + fiss_pick_winners_dev (the pick fused into the
 * record launch when d_records != NULL) as ONE call.  A small batch (launch-latency bound) runs, from the second call with
 * the same buffers on, as one cudaGraphLaunch of an instantiated graph whose kernel nodes are patched when the parameters
 * change (time_step_now moves every cycle, planning.py:124-128).  A batch that fills the GPU (B * nt >= 2 * SMs) is issued
 * launch by launch and CHAINED to the previous kernel of `stream` (programmatic dependent launch): called back to back, the
 * next step's CTAs start on the SMs this step's last work items leave idle and wait for the earlier launches only before
 * they write (so consecutive calls may share every buffer); an event record, a copy or a foreign kernel between two calls
 * simply ends a chain.  Not meant to be stream-captured into a caller's own graph with chaining on (FISS_CHAIN=0 then).
 * Asynchronous on `stream`; use a handle from one stream at a time.  d_mat / d_records / d_best_meta may be NULL. */
int32_t fiss_plan_grid_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B, const fiss_grid* g,
                           const fiss_params* p, double* d_cost, uint32_t* d_flags, double* d_mat, int32_t* d_best_idx,
                           double* d_best_cost, int32_t* d_best_meta, double* d_records, int32_t n_stride);

/* ---- host-pointer API (what the Python planners and a C caller use) ------------------------ */
/* plan() for B ego states over one shared lattice: H2D of ego/end states, the hot kernel, the pick
 * kernel, D2H of winners -- all inside, synchronous on return.
 *   ego [B][6], end [C][4]; out: best_idx [B], best_cost [B], best_meta [B][2] (n, n'),
 *   records optional [B][FISS_REC_ROWS][n_stride]; cost / flags optional [B*C] (the whole volume). */
int32_t fiss_plan_lattice_host(fiss_handle* h, void* stream, const double* ego, int32_t B,
                               const double* end, int32_t C, const fiss_params* p,
                               int32_t* best_idx, double* best_cost, int32_t* best_meta,
                               double* records, int32_t n_stride, double* cost, uint32_t* flags);

/* plan() over a product lattice: fiss_eval_grid_dev + the pick + the winners' records, host buffers. */
int32_t fiss_plan_grid_host(fiss_handle* h, void* stream, const double* ego, int32_t B, const fiss_grid* g,
                            const fiss_params* p, int32_t* best_idx, double* best_cost, int32_t* best_meta,
                            double* records, int32_t n_stride, double* cost, uint32_t* flags);

/* Streaming twin of fiss_plan_grid_host for a sequence of batches (planning.py:120-128 calls plan() once per step; a
 * batched caller submits one batch of ego states per step): `submit` enqueues H2D + kernels + D2H of a batch on one of
 * FISS_LANES lanes and returns; `wait` blocks until that lane's winners / records are in the caller's buffers.  The
 * device->host copy of one lane runs on a second stream under the kernels of the others: two lanes in rotation hide the
 * copy-back (a third measured no faster on a B200: the step is bound by what the kernels' stream carries).  Buffers as in
 * fiss_plan_grid_host (no cost / flags volume); they must stay valid until the lane has been waited for.  A lane with a call
 * in flight refuses another submit (FISS_ERR_STATE). */
#define FISS_LANES 4
int32_t fiss_plan_grid_submit(fiss_handle* h, void* stream, int32_t lane, const double* ego, int32_t B, const fiss_grid* g,
                              const fiss_params* p, int32_t* best_idx, double* best_cost, int32_t* best_meta,
                              double* records, int32_t n_stride);
int32_t fiss_plan_grid_wait(fiss_handle* h, int32_t lane);

/* "evaluate a list of end states" for one ego state: cost + masks for every entry,
 * and full records when `records` != NULL.  */
int32_t fiss_eval_end_states_host(fiss_handle* h, void* stream, const double* ego6, const double* end,
                                  int32_t N, const fiss_params* p, double* cost, uint32_t* flags,
                                  double* records, int32_t n_stride);

/* ---- multi-GPU: the cross-GPU best-cost pick (BASELINE config 5 with few problems) ------------------ */
/* One lattice can be split across the GPUs of a box: rank r evaluates a slab of it -- some lateral rows, or some
 * horizons -- and picks the slab's winner (fiss_plan_grid_dev on the slab's own fiss_grid).  Local candidate id c maps
 * to the id of FrenetOptimalPlanner's full numbering as (c / id_inner) * id_outer + c % id_inner + id_offset: a slab of
 * lateral rows [i_lo, i_hi) is a contiguous range (id_inner >= C_local, id_offset = i_lo * stride_d); a slab of
 * horizons [k_lo, k_hi) is (k_hi - k_lo) * nv consecutive ids out of every nt * nv (id_inner = (k_hi - k_lo) * nv,
 * id_outer = nt * nv, id_offset = k_lo * nv).  Either way the local order is the global order restricted.  fiss_allreduce_pick then makes every rank hold the GLOBAL winner under the reference's rule --
 * minimum cost, the LAST minimum in enumeration order on exact ties (frenet_optimal_planner.py:263-268) -- with
 *   * ONE ncclAllReduce(MIN, uint64) for the pick: a [B][nranks][2] slot table of (order-preserving cost key, global id),
 *     each rank filling its own slot (a float64 cost and an id do not fit one 64-bit key without dropping cost bits);
 *   * one ncclAllReduce(SUM, uint64) that moves the winners' records (+ (n, n')) from their owners -- the other ranks
 *     add zero words, so the bit patterns arrive exact; no root rank, hence no host round trip.
 * Everything is enqueued on `stream`; nothing synchronises.  In place: d_best_idx [B] local ids (-1 = none) -> global
 * ids, d_best_cost [B] -> the global minima (+inf = none), d_best_meta [B][2] and d_records [B][FISS_REC_ROWS][n_stride]
 * (either may be NULL) -> the global winner's (all-NaN record where nothing is feasible).
 * `comm` is an ncclComm_t of the caller (with its nranks / rank), or NULL for the communicator the handle owns
 * (fiss_comm_init).  libnccl is resolved at run time (the copy already in the process, e.g. torch's, else libnccl.so.2;
 * nccl_path may name one explicitly or be NULL); a single-GPU user never loads it. */
int32_t fiss_comm_unique_id(void* out128, const char* nccl_path);                       /* ncclGetUniqueId: 128 bytes */
int32_t fiss_comm_init(fiss_handle* h, const void* id128, int32_t nranks, int32_t rank, const char* nccl_path);
int32_t fiss_comm_destroy(fiss_handle* h);
int32_t fiss_allreduce_pick(fiss_handle* h, void* comm, int32_t nranks, int32_t rank, void* stream, int32_t B,
                            int64_t id_inner, int64_t id_outer, int64_t id_offset, int32_t* d_best_idx, double* d_best_cost, int32_t* d_best_meta,
                            double* d_records, int32_t n_stride);

/* number of kernel launches issued through this handle so far (bench.py's gpu_launches) */
int64_t fiss_launch_count(const fiss_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* FISS_ABI_H_ */
