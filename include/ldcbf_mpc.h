/*
 * ldcbf_mpc.h — C ABI of the B200-native LDCBF-MPC hot path (libldcbf_b200.so).
 *
 * The reference (salvatore373/Humanoid-Navigation-using-MPC-LDCBF) has no FFI: the seam is Python method
 * calls inside `HumanoidNavigation/MPC/HumanoidMpc.py`.  Each entry point below names the reference call
 * site(s) it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every array pointer is caller-owned, contiguous, row-major DEVICE memory unless the name ends in
 *    `_host`; fp64 / int32 / int8 as declared; the library allocates nothing persistent;
 *  - calls are asynchronous on `cuda_stream` (a cudaStream_t passed as void*; NULL = default stream),
 *    re-entrant, no global state;
 *  - return value: 0 on success, negative LDCBF_E_* on argument / launch errors (nothing is thrown);
 *    the numerical outcome of each scenario is reported in status[b] (LDCBF_STATUS_*).
 *  - state layout (p_x, v_x, p_y, v_y) and input layout (f_x, f_y) as in `HumanoidMpc.py:34-48`;
 *    obstacles are convex polygons given by their hull vertices in counter-clockwise order
 *    (`polygon.points[polygon.vertices]`, `Utils/ObstaclesUtils.py:54`), zero padded to max_verts.
 */
#ifndef LDCBF_MPC_H
#define LDCBF_MPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDCBF_ABI_VERSION 2   /* 2: ldcbf_rollout_f64 gained total_iters / end_code and streams obstacles beyond 8;
                                    ldcbf_rollout_unknown_f64 */

/* return codes */
#define LDCBF_OK 0
#define LDCBF_E_ARG (-1)      /* null pointer / non-positive size */
#define LDCBF_E_SHAPE (-2)    /* horizon or obstacle count not supported by the compiled kernels */
#define LDCBF_E_LAUNCH (-3)   /* CUDA launch error (see ldcbf_last_cuda_error) */

/* per-scenario status */
#define LDCBF_STATUS_SOLVED 0
#define LDCBF_STATUS_MAX_ITER 1
#define LDCBF_STATUS_INFEASIBLE 2   /* reference: IPOPT raises, loop breaks (HumanoidMpc.py:419-429) */
#define LDCBF_STATUS_DEGENERATE 3   /* CoM exactly on an obstacle edge: ||x-c|| = 0 (ObstaclesUtils.py:104) */
#define LDCBF_STATUS_DONE 4         /* rollout only: scenario already stopped (objective < stop_objective) */

/* why a closed loop ended (ldcbf_rollout_f64 end_code[b]; with sub-goals: the ending of the last run) */
#define LDCBF_END_STOP_RULE 0         /* previous objective < stop_objective (HumanoidMpc.py:392) */
#define LDCBF_END_BUDGET 1            /* num_inputs steps of the run / T steps of the buffers used up */
#define LDCBF_END_INFEASIBLE_NOW 2    /* a constant k = 0 LDCBF row is violated by more than eps_const_row */
#define LDCBF_END_INFEASIBLE_AHEAD 3  /* the QP over the future stages is infeasible (LDCBF rows against kinematic rows) */
#define LDCBF_END_DEGENERATE 4        /* CoM exactly on an obstacle edge */
#define LDCBF_END_MAX_ITER 5

/* supported shapes: horizon 1..4 by the register-resident solver (one thread per scenario) and 5..48 by the
 * long-horizon solver (one thread block per scenario, iteration cap max_iter * ceil(N / 4)); any number of
 * obstacles per scenario in ldcbf_mpc_qp_f64 / ldcbf_mpc_step_f64 (as far as shared memory holds N * max_obs row
 * flags for the long horizons: 64 obstacles at N = 48), at most 8 obstacles and horizon 4 in ldcbf_rollout_f64 */
#define LDCBF_MAX_HORIZON 4
#define LDCBF_MAX_HORIZON_LONG 48
#define LDCBF_MAX_OBSTACLES 8

/* Every key of HumanoidNavigation/config.yml:2-17 that the hot path reads, the constants the reference
 * hard-codes, and the solver controls.  Fill with ldcbf_params_default() and override. */
typedef struct ldcbf_params {
    double delta_t;         /* DELTA_T        config.yml:2  */
    double gravity;         /* GRAVITY_CONST  config.yml:3  */
    double com_height;      /* COM_HEIGHT     config.yml:4  */
    double alpha;           /* ALPHA          config.yml:5  */
    double l_max_x, l_max_y, l_min_x, l_min_y;  /* config.yml:6-9 */
    double v_min[2];        /* V_MIN          config.yml:10 */
    double v_max[2];        /* V_MAX          config.yml:11 */
    double omega_max;       /* 0.156*pi       HumanoidMpc.py:21 */
    double omega_min;       /* -omega_max     HumanoidMpc.py:22 */
    double foot_offset;     /* 0.05           HumanoidMpc.py:200 */
    double stop_objective;  /* 0.05           HumanoidMpc.py:392 (rollout) */
    double sampling_time;   /* ctor argument  HumanoidMpc.py:50,159 */
    double eps_active;      /* a constraint counts as violated below -eps_active (normalised rows) */
    double eps_const_row;   /* tolerance on the constant k = 0 LDCBF rows (1e-6, BASELINE.json) */
    int32_t max_iter;       /* active-set iteration cap per solve */
    int32_t flags;          /* LDCBF_FLAG_* */
    double eps_infeasible;  /* a QP is reported infeasible only when the row that cannot be satisfied is violated by more
                               than this (1e-9, natural units): below it the feasible set is a point to rounding and the
                               solve is repeated with rows counted as violated only below -eps_infeasible (ABI 2) */
} ldcbf_params;

/* flags.  Default 0: the half-plane builder restates the reference's per-edge arithmetic operation by operation (it walks
 * the hull ring in vertex order where the reference iterates ConvexHull.simplices: (c, eta) equal the reference's to
 * ~1e-13, and are bit-equal to the ring-order oracle).
 * FAST_GEOMETRY trades that for ~1 ulp agreement (one reciprocal instead of two square roots and a division per
 * edge); use it only when bit parity of (c, eta) with the reference is not needed. */
#define LDCBF_FLAG_FAST_GEOMETRY 1
/* ldcbf_rollout_f64 warm-starts every solve from the previous step's active set shifted by one stage (the result is
 * exact either way; the reference warm-starts IPOPT the same way, HumanoidMpc.py:450-455).  COLD_START disables it. */
#define LDCBF_FLAG_COLD_START 2
/* The open-loop entry points (ldcbf_mpc_qp_f64, ldcbf_mpc_step_f64, ldcbf_mpc_step_packed_f64) start every solve
 * from a geometric active-set guess (all velocity rows on the side the goal lies on) instead of the empty set: same
 * optimum, about a third of the iterations.  COLD_START disables that as well.
 * COOP_LANES: batches of at most 1024 scenarios with N <= 3 and at most 8 obstacles are solved by one warp per
 * scenario (QR-updated active set, rows scanned in parallel) instead of one thread per scenario. */
#define LDCBF_FLAG_COOP_LANES 4

/* Optional per-scenario overrides of the limits `bounds_tuning.py:22-26` mutates:
 * limits[b] = (ALPHA, V_MAX[0], V_MAX[1], OMEGA_MAX, OMEGA_MIN, reserved).  NaN entries fall back to ldcbf_params. */
#define LDCBF_LIMITS_STRIDE 6

int ldcbf_abi_version(void);
void ldcbf_params_default(ldcbf_params* prm);
const char* ldcbf_last_cuda_error(void);

/* Caller-provided workspace: none (returns 0; kept so callers written against SURVEY.md §8b's proposal link).
 * One internal buffer exists: for batches of 151 552 scenarios and more the solve runs as two kernels (prepare:
 * heading schedule + warm start, one thread per scenario; resume: active-set iterations with lane refill) that hand the
 * solver state over through a stream-ordered allocation (cudaMallocFromPoolAsync / cudaFreeAsync on the caller's
 * stream; 848 B per scenario at N = 3, at most 2^20 scenarios at a time) from a memory pool the library owns per
 * device.  The pool keeps that memory between calls; ldcbf_trim_workspace() hands it back to the driver (call it
 * when no step is in flight on the current device). */
size_t ldcbf_workspace_bytes(int B, int N, int max_obs, int max_verts);
int ldcbf_trim_workspace(void);

/* K1 — LDCBF half-plane builder.
 * Replaces ObstaclesUtils.get_closest_point_and_normal_vector_from_obs (Utils/ObstaclesUtils.py:60-109),
 * is_point_inside_polygon (:50-57) and the loop of HumanoidMPC._get_list_c_and_eta (MPC/HumanoidMpc.py:296-319).
 *   pos    [B,2]                      current CoM (p_x, p_y)
 *   verts  [B,max_obs,max_verts,2]    hull vertices, CCW, zero padded
 *   nverts [B,max_obs] int32, nobs [B] int32
 *   c_eta  [B,max_obs,4] out          (c_x, c_y, eta_x, eta_y); rows o >= nobs[b] are zero */
int ldcbf_halfplanes_f64(int B, int max_obs, int max_verts, const double* pos, const double* verts,
                         const int32_t* nverts, const int32_t* nobs, double* c_eta, void* cuda_stream);

/* K2+K3 — heading schedule, QP assembly, exact QP solve, one LIP integration, given the half-planes.
 * Replaces _precompute_theta_omega_naive (HumanoidMpc.py:137-160), the constraint/cost builders (:162-249,
 * :252-294, :321-333), optim_prob.solve() (:417) and _integrate (:335-343, :441-447).
 *   x0 [B,4], theta0 [B], goal [B,2], foot [B,N+1] int8 (+1 right / -1 left, the s_v window of :403)
 *   c_eta [B,max_obs,4], nobs [B]; delta [B] or NULL (HumanoidMPCCustomLCBF.py:30-31);
 *   limits [B,6] or NULL
 * out: U [B,N,2] footsteps, X [B,N+1,4] predicted states (X[:,1] is the next state), theta [B,N+1],
 *      omega [B,N], obj [B] (value of the reference's cost incl. the constant k=0 term), status [B], iters [B].
 * For status != 0 the U, X, obj entries are NaN. */
int ldcbf_mpc_qp_f64(const ldcbf_params* prm, int B, int N, int max_obs, const double* x0, const double* theta0,
                     const double* goal, const int8_t* foot, const double* c_eta, const int32_t* nobs,
                     const double* delta, const double* limits, double* U, double* X, double* theta,
                     double* omega, double* obj, int32_t* status, int32_t* iters, void* cuda_stream);

/* One full MPC step = K1 then K2+K3 on the same stream (HumanoidMpc.py:387-447 for one k).
 * `warm` [B,2N] or NULL is accepted for API parity with the reference's set_initial (:450-455); the exact
 * active-set solver does not need it. */
int ldcbf_mpc_step_f64(const ldcbf_params* prm, int B, int N, int max_obs, int max_verts, const double* x0,
                       const double* theta0, const double* goal, const int8_t* foot, const double* verts,
                       const int32_t* nverts, const int32_t* nobs, const double* delta, const double* warm,
                       const double* limits, double* U, double* X, double* theta, double* omega, double* c_eta,
                       double* obj, int32_t* status, int32_t* iters, void* cuda_stream);

/* Loop-shaped variant of ldcbf_mpc_step_f64 for closed loops driven from the host (the reference's
 * run_simulation keeps its state in X_pred[:, k], HumanoidMpc.py:396-447): ONE row in and ONE row out per scenario, so
 * a step costs one upload and one download.
 *   state [B,6] (p_x, v_x, p_y, v_y, theta, first stance foot +1/-1); the foot window alternates from the first
 *               foot exactly like the reference's s_v list (HumanoidMpc.py:104-108, :403)
 *   next  [B,10] out = (x_next[4], theta_1, u0_x, u0_y, omega_0, objective, status as a double)
 *   c_eta [B,max_obs,4] scratch / out;  iters [B] out or NULL */
int ldcbf_mpc_step_packed_f64(const ldcbf_params* prm, int B, int N, int max_obs, int max_verts, const double* state,
                              const double* goal, const double* verts, const int32_t* nverts, const int32_t* nobs,
                              const double* delta, const double* limits, double* next, double* c_eta,
                              int32_t* iters, void* cuda_stream);

/* K4 — LiDAR ray casting.  Replaces compute_lidar_readings / get_closest_point
 * (RangeFinder/range_finder_wth_polygons_dbscan.py:13-63) and line_polygon_intersection (Utils/obstacles.py:95-139).
 *   ray_dirs [R,2]  lidar_range*(cos, sin)(i*2*pi/R) computed on the host with libm exactly as :28-37
 *   pos [B,2]; verts/nverts/nobs as above but in the order the reference casts against
 *   (ConvexHull.points rows, HumanoidMPCUnknownEnvironment.py:46)
 * out: hit_obs [B,R] int32 (-1 = no hit), hit_edge [B,R] int32, hit_xy [B,R,2] (NaN = no hit). */
int ldcbf_lidar_cast_f64(int B, int R, const double* ray_dirs, double lidar_range, const double* pos, int max_obs,
                         int max_verts, const double* verts, const int32_t* nverts, const int32_t* nobs,
                         int32_t* hit_obs, int32_t* hit_edge, double* hit_xy, void* cuda_stream);

/* f1 — LiDAR post-processing: clustering and convex hulls of the clusters, one scan per scenario.
 * Replaces retrieve_clusters (sklearn DBSCAN, eps = 0.3, min_samples = 3), create_convex_hull and build_local_obstacles
 * (RangeFinder/range_finder_wth_polygons_dbscan.py:65-126) and the Gaussian-noise step (:161-172; the noise is an
 * input tensor because the reference draws it from the unseeded global numpy RNG).
 *   hit_xy [B,R,2] readings (NaN = none), R <= 512;  noise [B,R,2] or NULL
 * out: labels [B,R] int32   cluster number per ray in sklearn's label order, -1 = noise / no reading
 *      hull_verts [B,max_hulls,max_hull_verts,2]  hull vertices, counter-clockwise, no closing vertex, zero padded
 *      hull_nverts [B,max_hulls], n_hulls [B]     clusters the reference discards (< 3 distinct points, collinear)
 *                                                 produce no hull; hulls keep the cluster order
 *      overflow [B] or NULL                       1 when a scan had more hulls / vertices than the buffers hold
 * The outputs have the layout of the verts / nverts / nobs inputs of ldcbf_halfplanes_f64 and ldcbf_mpc_step_f64. */
int ldcbf_lidar_clusters_f64(int B, int R, const double* hit_xy, const double* noise, double eps, int min_samples,
                             int max_hulls, int max_hull_verts, int32_t* labels, double* hull_verts,
                             int32_t* hull_nverts, int32_t* n_hulls, int32_t* overflow, void* cuda_stream);

/* f3 — map front-end of the sub-goal planner: occupancy grid, exact Euclidean distance transform and clearance
 * cost, batched.  Replaces HumanoidMPCWithRRT.py:21-88 (_build_occupancy_grid: frame = bounding box of the hull
 * vertices, the origin and the goal padded by 3; height = ceil(width * dy / dx); np.round to cells; cells of
 * [xmin,xmax) x [ymin,ymax) inside the triangulated rounded hull) and :103-108 (scipy distance_transform_edt(1 - og),
 * exp(-d)).  The RRT* search that consumes the cost map lives in the un-vendored rrtplanner package and stays on
 * the host (mirror: HumanoidNavigation/MPC/HumanoidMPCVariants/rrt_star.py).
 *   goal [B,2]; verts/nverts/nobs as in K1 (hull rings)
 *   meta [B,6] out = (min_x, min_y, max_x, max_y, height_grid_size, occupied cells or -1 when height > h_cap)
 *   og   [B,width+1,h_cap+1] uint8 out (cell (i,j), j fastest; columns above height_grid_size are zero)
 *   dist, cost [B,width+1,h_cap+1] out (cost may be NULL); entries with j > height_grid_size are not written
 *   work [B,width+1,h_cap+1] int32 scratch
 * og and dist are bit-exact against the reference; cost within 1 ulp of np.exp. */
int ldcbf_clearance_grid_f64(int B, int width, int h_cap, int max_obs, int max_verts, const double* goal,
                             const double* verts, const int32_t* nverts, const int32_t* nobs, double* meta,
                             uint8_t* og, double* dist, double* cost, int32_t* work, void* cuda_stream);

/* Closed loop (HumanoidMpc.py:380-459, incl. the mpc_step = int(DELTA_T/sampling_time) sub-stepping of :74-78,
 * :384,:443-446) with optional sub-goal sequencing (HumanoidMPCVariants/HumanoidMPCWithRRT.py:153-181: a fresh
 * run per sub-goal — objective memory and foot parity restart, the state carries over).  One kernel launch.
 *   state  [B,5] in/out   (p_x, v_x, p_y, v_y, theta)
 *   goals  [B,n_goals,2]; right_first [B] int8 (1 = start_with_right_foot)
 *   T = total loop-iteration budget per scenario (size of the trajectory buffers);
 *   max_steps_per_goal = num_inputs of one run (mpc_step * N_mpc_timesteps)
 *   traj_X [B,T+1,5] or NULL (row 0 = initial state), traj_U [B,T,3] or NULL (f_x, f_y, omega)
 *   steps [B] out: loop iterations executed; goal_steps [B,n_goals] out: iterations spent on each sub-goal;
 *   status [B] out: status of the last solve; total_solves, total_iters: optional device counters (+= QP solves,
 *   += active-set iterations of those solves); end_code [B] out or NULL: LDCBF_END_* of the (last) run. */
int ldcbf_rollout_f64(const ldcbf_params* prm, int B, int N, int T, int n_goals, int max_steps_per_goal, int max_obs,
                      int max_verts, double* state, const double* goals, const int8_t* right_first,
                      const double* verts, const int32_t* nverts, const int32_t* nobs, const double* delta,
                      const double* limits, double* traj_X, double* traj_U, int32_t* steps, int32_t* goal_steps,
                      int32_t* status, int64_t* total_solves, int64_t* total_iters, int32_t* end_code,
                      void* cuda_stream);

/* Closed loop of the unknown-environment variant (HumanoidMPCVariants/HumanoidMPCUnknownEnvironment.py:30-68 inside the
 * step loop HumanoidMpc.py:380-459): every step scans the TRUE map (verts/nverts/nobs in ConvexHull.points order, as
 * ldcbf_lidar_cast_f64) from the current CoM, clusters the readings and builds the hulls (as ldcbf_lidar_clusters_f64:
 * noise [B,R,2] or NULL is added to the readings of every step, eps / min_samples / max_hulls / max_hull_verts as
 * there) and solves the MPC step against the INFERRED obstacles; stop rule, break on a failed solve, integration and
 * the alternating foot-parity window as in the reference's loop.  5 launches per step on the caller's stream, no host
 * round trip; sampling_time must equal delta_t (no sub-stepping, else LDCBF_E_SHAPE).  Work buffers (about
 * 28*R + 16*max_hulls*max_hull_verts bytes per scenario) come from the library's stream-ordered pool.
 *   state [B,5] in/out; goal [B,2]; right_first [B] int8
 *   traj_X [B,T+1,5] or NULL, traj_U [B,T,3] or NULL; steps, status, end_code [B] out (LDCBF_END_*);
 *   overflow [B] out or NULL: 1 when a scan produced more hulls / hull vertices than max_hulls / max_hull_verts hold
 *   (the surplus was dropped); total_solves: optional device counter. */
int ldcbf_rollout_unknown_f64(const ldcbf_params* prm, int B, int N, int T, int R, const double* ray_dirs,
                              double lidar_range, int max_obs, int max_verts, double* state, const double* goal,
                              const int8_t* right_first, const double* verts, const int32_t* nverts,
                              const int32_t* nobs, const double* noise, double eps, int min_samples, int max_hulls,
                              int max_hull_verts, const double* delta, const double* limits, double* traj_X,
                              double* traj_U, int32_t* steps, int32_t* status, int32_t* end_code, int32_t* overflow,
                              int64_t* total_solves, void* cuda_stream);

/* FP64 FMA-chain probe used by bench.py to measure the FP64 pipe peak on the box (roofline denominator).
 * Launches `blocks` x `threads` threads, each running `iters` x 8 independent FMAs; out[blocks*threads]. */
int ldcbf_probe_fp64_fma(int blocks, int threads, int iters, double* out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LDCBF_MPC_H */
