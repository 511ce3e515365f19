// Closed loop of the unknown-environment variant, entirely on the device.
//
// Reference control flow restated (HumanoidNavigation/):
//   MPC/HumanoidMpc.py:380-459                            the step loop (stop rule :392, break on a failed solve :419-429,
//                                                         record u_0 / omega_0 :432-433, integrate :441-447)
//   MPC/HumanoidMPCVariants/HumanoidMPCUnknownEnvironment.py:30-68   _get_list_c_and_eta of the variant: every step scans
//                                                         the TRUE map from the current CoM (range_finder), clusters the
//                                                         readings, builds the hulls and takes the half-planes of the
//                                                         INFERRED obstacles
// Per step, for all B scenarios at once, on one stream and without a host round trip:
//   K4 lidar_kernel -> f1 lidar_clusters_kernel -> K1 half-planes of the hulls -> K2+K3 solve -> unknown_advance_kernel
// (stop rule, status, integration, trajectory rows, foot parity; a scenario that has ended keeps its state and is
// masked out).  All scenarios are in lockstep, so the foot-parity window of step t is the same function of t for all
// of them and the loop-shaped step entry (state rows [B,6] in, result rows [B,10] out) carries everything.
// Sub-stepping (DELTA_T / sampling_time > 1) is not supported here (LDCBF_E_SHAPE); the reference's unknown-environment
// runs use sampling_time = DELTA_T (simulation_1.py:222-231).
#include "ldcbf_common.cuh"

namespace ldcbf {

struct UnknownLoop {
    double* state;            // [B,5] in/out
    double* state6;           // [B,6] work: (p_x, v_x, p_y, v_y, theta, first stance foot)
    double2* pos;             // [B] work: LiDAR position of the next scan
    const double* next10;     // [B,10] result rows of the step
    const int32_t* iters;     // [B]
    const int32_t* ovf_step;  // [B] overflow flag of this step's clustering
    double* last_obj;         // [B]
    int32_t* alive;           // [B]
    double* traj_X; double* traj_U;
    int32_t* steps; int32_t* status; int32_t* end_code; int32_t* overflow;
    unsigned long long* total_solves;
};

__global__ void unknown_init_kernel(int B, int T, const int8_t* __restrict__ right_first, UnknownLoop io) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* s = io.state + 5 * (size_t)b;
    double* s6 = io.state6 + 6 * (size_t)b;
    for (int i = 0; i < 5; ++i) s6[i] = s[i];
    s6[5] = right_first[b] ? 1.0 : -1.0;                       // s_v[0], HumanoidMpc.py:104-108
    io.pos[b] = make_double2(s[0], s[2]);
    io.last_obj[b] = INFINITY;
    io.alive[b] = 1;
    io.steps[b] = 0;
    io.status[b] = LDCBF_STATUS_SOLVED;
    io.end_code[b] = LDCBF_END_BUDGET;
    if (io.overflow) io.overflow[b] = 0;
    if (io.traj_X) for (int i = 0; i < 5; ++i) io.traj_X[(size_t)b * (T + 1) * 5 + i] = s[i];
}

__global__ void unknown_advance_kernel(int B, int T, int t, double stop_objective, UnknownLoop io) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int solved = 0;
    if (b < B && io.alive[b]) {
        if (io.last_obj[b] < stop_objective) {                 // :392 — checked before the solve of this step counts
            io.alive[b] = 0;
            io.end_code[b] = LDCBF_END_STOP_RULE;
        } else {
            solved = 1;
            const double* r = io.next10 + 10 * (size_t)b;      // (x_next[4], theta_1, u0_x, u0_y, omega_0, objective, status)
            const int st = (int)r[9];
            if (io.overflow && io.ovf_step[b]) io.overflow[b] = 1;
            if (st != LDCBF_STATUS_SOLVED) {                   // :419-429 break
                io.alive[b] = 0;
                io.status[b] = st;
                io.end_code[b] = st == LDCBF_STATUS_DEGENERATE ? LDCBF_END_DEGENERATE
                               : st == LDCBF_STATUS_MAX_ITER ? LDCBF_END_MAX_ITER
                               : io.iters[b] == 0 ? LDCBF_END_INFEASIBLE_NOW : LDCBF_END_INFEASIBLE_AHEAD;
            } else {
                double* s6 = io.state6 + 6 * (size_t)b;
                for (int i = 0; i < 5; ++i) s6[i] = r[i];      // :441-447
                s6[5] = -s6[5];                                // next window starts with the other foot (:401-403)
                io.pos[b] = make_double2(r[0], r[2]);
                io.last_obj[b] = r[8];
                io.steps[b] = t + 1;
                if (io.traj_U) {
                    double* u = io.traj_U + ((size_t)b * T + t) * 3;
                    u[0] = r[5]; u[1] = r[6]; u[2] = r[7];     // :432-433
                }
                if (io.traj_X) {
                    double* x = io.traj_X + ((size_t)b * (T + 1) + t + 1) * 5;
                    for (int i = 0; i < 5; ++i) x[i] = r[i];
                }
            }
        }
    }
    if (io.total_solves) {
        const unsigned s = __reduce_add_sync(0xffffffffu, (unsigned)solved);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(io.total_solves, (unsigned long long)s);
    }
}

__global__ void unknown_finish_kernel(int B, UnknownLoop io) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int i = 0; i < 5; ++i) io.state[5 * (size_t)b + i] = io.state6[6 * (size_t)b + i];
}

}  // namespace ldcbf

extern "C" int ldcbf_rollout_unknown_f64(const ldcbf_params* prm, int B, int N, int T, int R, const double* ray_dirs,
                                         double lidar_range, int max_obs, int max_verts, double* state,
                                         const double* goal, const int8_t* right_first, const double* verts,
                                         const int32_t* nverts, const int32_t* nobs, const double* noise, double eps,
                                         int min_samples, int max_hulls, int max_hull_verts, const double* delta,
                                         const double* limits, double* traj_X, double* traj_U, int32_t* steps,
                                         int32_t* status, int32_t* end_code, int32_t* overflow, int64_t* total_solves,
                                         void* cuda_stream) {
    using namespace ldcbf;
    if (!prm || B < 0 || T <= 0 || R <= 0 || max_obs <= 0 || max_verts <= 0 || max_hulls <= 0 || max_hull_verts < 3)
        return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!ray_dirs || !state || !goal || !right_first || !verts || !nverts || !nobs || !steps || !status || !end_code)
        return LDCBF_E_ARG;
    if ((int)(prm->delta_t / prm->sampling_time) > 1) return LDCBF_E_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaMemPool_t pool = workspace_pool();
    if (!pool) { set_last_error(cudaErrorMemoryAllocation); return LDCBF_E_LAUNCH; }
    // one stream-ordered allocation, carved up (every piece a multiple of 16 B)
    auto up = [](size_t n) { return (n + 15) / 16 * 16; };
    const size_t Bs = (size_t)B;
    const size_t sz[] = {up(Bs * 6 * 8), up(Bs * 16), up(Bs * 10 * 8), up(Bs * max_hulls * 32), up(Bs * R * 4), up(Bs * R * 4),
                         up(Bs * R * 16), up(Bs * R * 4), up(Bs * max_hulls * max_hull_verts * 16), up(Bs * max_hulls * 4),
                         up(Bs * 4), up(Bs * 4), up(Bs * 8), up(Bs * 4), up(Bs * 4)};
    size_t total = 0;
    for (size_t s : sz) total += s;
    char* base = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&base, total, pool, st);
    if (e != cudaSuccess) { set_last_error(e); cudaGetLastError(); return LDCBF_E_LAUNCH; }
    char* cur = base;
    size_t k = 0;
    auto take = [&]() { char* p = cur; cur += sz[k++]; return p; };
    double* state6 = (double*)take(); double2* pos = (double2*)take(); double* next10 = (double*)take();
    double* c_eta = (double*)take(); int32_t* hit_obs = (int32_t*)take(); int32_t* hit_edge = (int32_t*)take();
    double* hit_xy = (double*)take(); int32_t* labels = (int32_t*)take(); double* hull_verts = (double*)take();
    int32_t* hull_nverts = (int32_t*)take(); int32_t* n_hulls = (int32_t*)take(); int32_t* ovf = (int32_t*)take();
    double* last_obj = (double*)take(); int32_t* alive = (int32_t*)take(); int32_t* iters = (int32_t*)take();

    const UnknownLoop io{state, state6, pos, next10, iters, ovf, last_obj, alive, traj_X, traj_U, steps, status, end_code,
                         overflow, reinterpret_cast<unsigned long long*>(total_solves)};
    const unsigned grid = (unsigned)((B + 127) / 128);
    unknown_init_kernel<<<grid, 128, 0, st>>>(B, T, right_first, io);
    int rc = check_launch();
    for (int t = 0; t < T && rc == LDCBF_OK; ++t) {
        rc = ldcbf_lidar_cast_f64(B, R, ray_dirs, lidar_range, reinterpret_cast<const double*>(pos), max_obs, max_verts,
                                  verts, nverts, nobs, hit_obs, hit_edge, hit_xy, cuda_stream);
        if (rc == LDCBF_OK)
            rc = ldcbf_lidar_clusters_f64(B, R, hit_xy, noise, eps, min_samples, max_hulls, max_hull_verts, labels,
                                          hull_verts, hull_nverts, n_hulls, ovf, cuda_stream);
        if (rc == LDCBF_OK)
            rc = ldcbf_mpc_step_packed_f64(prm, B, N, max_hulls, max_hull_verts, state6, goal, hull_verts, hull_nverts,
                                           n_hulls, delta, limits, next10, c_eta, iters, cuda_stream);
        if (rc == LDCBF_OK) {
            unknown_advance_kernel<<<grid, 128, 0, st>>>(B, T, t, prm->stop_objective, io);
            rc = check_launch();
        }
    }
    if (rc == LDCBF_OK) {
        unknown_finish_kernel<<<grid, 128, 0, st>>>(B, io);
        rc = check_launch();
    }
    cudaFreeAsync(base, st);
    return rc;
}
