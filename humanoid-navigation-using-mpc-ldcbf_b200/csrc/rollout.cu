// Closed-loop rollout: the whole step loop of HumanoidMPC.run_simulation in ONE kernel launch.
//
// Reference control flow restated (HumanoidNavigation/MPC/HumanoidMpc.py):
//   :380-459  for k in range(num_inputs): half-planes at the current CoM (:387) -> stop when the previous
//             objective < 0.05 (:392) -> foot-parity window (:401-403) -> heading schedule (:406-411) ->
//             solve on MPC timesteps (:415-429, break on failure) -> record u_0, omega_0 (:432-433) ->
//             integrate / hold position on sub-steps (:439-447)
//   :74-78    mpc_step = int(DELTA_T / sampling_time) (at least 1), num_inputs = mpc_step * N_simul
// and the sub-goal sequencing of MPC/HumanoidMPCVariants/HumanoidMPCWithRRT.py:153-181 (a fresh run per
// sub-goal: objective memory and foot parity restart, the state carries over).
//
// Mapping: one thread (or a group of 2 / 4 lanes that split the ring walk) per scenario, persistent over all its
// steps: no per-step launch latency; the scenario's map (vertex rings, edge constants, half-planes of the step) is
// staged in shared memory for the whole run when it fits, state and active data live in registers.  Scenarios are
// independent, so a block never synchronises.  Every solve is warm-started from the previous step's active set, shifted
// by one stage and with the last stage repeated (mpc_qp.cuh: shift_codes, qp_warm_start).
#include <cstdio>
#include <cstdlib>

#include "halfplane_dev.cuh"
#include "mpc_qp.cuh"

namespace ldcbf {

// development aid (make profile-lib): cycles per phase of a loop step, accumulated by thread 0 of block 0 and printed
#if defined(LDCBF_ROLLOUT_PROFILE) && defined(__CUDA_ARCH__)
#define RO_T(k) do { const long long now_ = clock64(); prof_[k] += now_ - prof_t_; prof_t_ = now_; } while (0)
#else
#define RO_T(k) do { } while (0)
#endif

struct RolloutIO {
    double* state; const double* goals; const int8_t* right_first; const double2* verts; const int32_t* nverts;
    const int32_t* nobs; const double* delta; const double* limits; double* traj_X; double* traj_U;
    int32_t* steps; int32_t* goal_steps; int32_t* status; unsigned long long* total_solves; bool fast_geometry;
    bool warm_start;
    // ABI 2: half-planes of the obstacles beyond the MO register-resident ones ([B,max_obs] double4, library-owned
    // scratch), iteration counter, per-scenario reason the loop ended (LDCBF_END_*)
    double4* ce_scratch; unsigned long long* total_iters; int32_t* end_code;
    bool prune;       // LDCBF_ROLLOUT_NOPRUNE=1 switches the pruned ring walk off (A/B runs)
    int start_mode;   // LDCBF_ROLLOUT_START: 2 shifted active set + last stage repeated (default), 0 shifted only, 1 geometric guess
    double* den_scratch;   // [B,max_obs,max_verts] edge constants of the EXACT ring walk, tabulated once per run
};

// G lanes per scenario (small batches): the kernel time is the slowest scenario's chain of sequential steps, and the
// largest fixed part of a step is the walk over the scenario's rings (52 edges, each a division and a square root).  The
// G lanes split every ring (halfplane_group, bit-equal to the serial walk) and then all run the same solve in lockstep
// (same data, same control flow: no divergence, no communication); lane 0 of the group writes.  G is chosen by batch
// size in launch_rollout (4 / 2 / 1: more lanes shorten the walk, more warps cost instruction fetch).
template <int N, int MO, bool EXACT, int BLOCK, int G>
__global__ void __launch_bounds__(BLOCK) rollout_kernel(StepConst C, int B, int T, int n_goals, int max_steps_per_goal,
                                                      int substeps, int max_obs, int max_verts, int map_doubles,
                                                      RolloutIO io) {
    extern __shared__ __align__(32) double qp_ws[];
    const int b = (blockIdx.x * BLOCK + threadIdx.x) / G;
    if (b >= B) return;                                       // whole groups leave together (G divides BLOCK)
    const int glane = threadIdx.x % G;
    const bool writer = glane == 0;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) / G * G));
    double px = io.state[5 * (size_t)b], vx = io.state[5 * (size_t)b + 1], py = io.state[5 * (size_t)b + 2],
           vy = io.state[5 * (size_t)b + 3], th = io.state[5 * (size_t)b + 4];
    const bool right_first = io.right_first[b] != 0;
    const int nt = min(io.nobs[b], max_obs);
    const int nb = min(nt, MO);
    const int n_stream = nt - nb;
    // The scenario's map for the whole run — vertex rings, the edge constants of the bit-exact ring walk (|AB|^2 as the
    // reference forms it, a square root and a product: tabulated once, off the per-step critical path, same bits), the
    // vertex counts and the half-planes of the current step — lives in SHARED memory when it fits (map_doubles > 0: the
    // launcher reserved map_doubles doubles per scenario behind the solver workspaces): no global load is left in the
    // step loop (clock64 profile: the ring walk was the largest phase of a step, 12 k of ~32 k cycles, most of it
    // waiting for L2).  Otherwise the same pointers refer to global memory (vertices in place, scratch from the pool).
    double4* ces;
    const double2* rings;
    double* dens;
    const int32_t* nvs;
    int32_t* prevs = nullptr;
    if (map_doubles > 0) {
        double* m = qp_ws + (size_t)QpWorkspace<N>::DOUBLES * BLOCK + (size_t)(threadIdx.x / G) * map_doubles;
        ces = reinterpret_cast<double4*>(m);
        double2* sv = reinterpret_cast<double2*>(m + 4 * max_obs);
        dens = m + 4 * max_obs + 2 * max_obs * max_verts;
        int32_t* snv = reinterpret_cast<int32_t*>(dens + max_obs * max_verts);
        prevs = io.prune ? snv + max_obs : nullptr;              // closest edge of every obstacle at the previous step
        const double2* gv = io.verts + (size_t)b * max_obs * max_verts;
        for (int i = glane; i < nt * max_verts; i += G) sv[i] = gv[i];
        for (int o = glane; o < max_obs; o += G) {
            snv[o] = o < nt ? min(io.nverts[(size_t)b * max_obs + o], max_verts) : 0;
            if (prevs) prevs[o] = 0;
        }
        __syncwarp(gmask);
        rings = sv; nvs = snv;
    } else {
        ces = io.ce_scratch + (size_t)b * max_obs;
        rings = io.verts + (size_t)b * max_obs * max_verts;
        dens = (EXACT && io.den_scratch) ? io.den_scratch + (size_t)b * max_obs * max_verts : nullptr;
        nvs = io.nverts + (size_t)b * max_obs;
    }
    if (!EXACT) dens = nullptr;
    const double dl = io.delta ? io.delta[b] : 0.0;
    const Limits lim = load_limits(C, io.limits, (size_t)b);
    double* tX = (io.traj_X && writer) ? io.traj_X + (size_t)b * (T + 1) * 5 : nullptr;
    double* tU = (io.traj_U && writer) ? io.traj_U + (size_t)b * T * 3 : nullptr;
    if (tX) { tX[0] = px; tX[1] = vx; tX[2] = py; tX[3] = vy; tX[4] = th; }

#if defined(LDCBF_ROLLOUT_PROFILE) && defined(__CUDA_ARCH__)
    long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_t_ = clock64();
#endif
    // lane l of the group tabulates the constants of the edges it will walk, so every lane later reads its own stores
    if (dens) {
        for (int o = 0; o < nt; ++o) {
            const int V = min(nvs[o], max_verts);
            const double2* ring = rings + (size_t)o * max_verts;
            for (int e = glane; e < V; e += G) dens[(size_t)o * max_verts + e] = edge_den_exact(ring[e], ring[(e + 1 == V) ? 0 : e + 1]);
        }
    }
    int gi = 0, kstep = 0, total = 0, solves = 0, iters_sum = 0, last_status = LDCBF_STATUS_SOLVED;
    int end_code = LDCBF_END_BUDGET;
    int warm[2 * N];                       // active set of the previous step, shifted by one stage (-1: none)
#pragma unroll
    for (int j = 0; j < 2 * N; ++j) warm[j] = -1;
    double last_obj = INFINITY, ux = 0.0, uy = 0.0;
    if (writer) for (int i = 0; i < n_goals; ++i) io.goal_steps[(size_t)b * n_goals + i] = 0;

    while (gi < n_goals && total < T) {
        if (last_obj < C.stop_objective || kstep >= max_steps_per_goal) {     // :392 / loop exhausted
            end_code = last_obj < C.stop_objective ? LDCBF_END_STOP_RULE : LDCBF_END_BUDGET;
            if (!(last_obj < C.stop_objective) && gi + 1 < n_goals) {
                // a run that used up its num_inputs steps returns X_pred[:, :num_inputs] (HumanoidMpc.py:458) and the
                // next sub-goal run starts from its LAST column (HumanoidMPCWithRRT.py:178): the state before the last
                // integration, parked in io.state below
                px = io.state[5 * (size_t)b]; vx = io.state[5 * (size_t)b + 1]; py = io.state[5 * (size_t)b + 2];
                vy = io.state[5 * (size_t)b + 3]; th = io.state[5 * (size_t)b + 4];
            }
            if (writer) io.goal_steps[(size_t)b * n_goals + gi] = kstep;
            ++gi; kstep = 0; last_obj = INFINITY;
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) warm[j] = -1;                      // a fresh run per sub-goal
            continue;
        }
        const double gx = io.goals[((size_t)b * n_goals + gi) * 2], gy = io.goals[((size_t)b * n_goals + gi) * 2 + 1];
        double th1, om0;
        RO_T(0);
        if (n_goals > 1 && kstep + 1 >= max_steps_per_goal) {      // last step of a run: park the state before it (lanes
            io.state[5 * (size_t)b] = px; io.state[5 * (size_t)b + 1] = vx; io.state[5 * (size_t)b + 2] = py;   // of a group
            io.state[5 * (size_t)b + 3] = vy; io.state[5 * (size_t)b + 4] = th;                  // store equal values)
        }
        if (kstep % substeps == 0) {
            // half-planes at the current CoM (:387).  ONE copy of the ring walk in the instruction stream (the loop over
            // obstacles is not unrolled: this kernel is bound by instruction fetch — ncu: 1.9 of 5.3 cycles per issued
            // instruction wait for the next instruction, the loop body of a step is ~170 KB against a 32 KB L1.5
            // instruction cache); the results go through the scratch row of the scenario, every lane of the group
            // stores the same value and reads back its own store
#pragma unroll 1
            for (int o = 0; o < nt; ++o) {
                const int V = min(nvs[o], max_verts);
                if (V <= 0) { ces[o] = make_double4(0.0, 0.0, 0.0, 0.0); continue; }
                // staged map: walk with pruning around the previous step's closest edge (every lane of the group computes
                // the same new index and stores it: equal values); otherwise the plain walk
                int prev = prevs ? prevs[o] : -1;
                ces[o] = prevs ? halfplane_group_pruned<EXACT, G>(px, py, rings + (size_t)o * max_verts, V, glane, gmask,
                                                                  dens ? dens + (size_t)o * max_verts : nullptr, prev)
                               : halfplane_group<EXACT, G, false>(px, py, rings + (size_t)o * max_verts, V, glane, gmask,
                                                                  dens ? dens + (size_t)o * max_verts : nullptr);
                if (prevs) prevs[o] = prev;
            }
            double4 ce[MO];
#pragma unroll
            for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? ces[o] : make_double4(0.0, 0.0, 0.0, 0.0);
            RO_T(1);
            int ft[N + 1];
            const int step_number = kstep / substeps;                                 // :401
#pragma unroll
            for (int k = 0; k <= N; ++k) ft[k] = (((step_number + k) & 1) == (right_first ? 0 : 1)) ? 1 : -1;
            QpSolution<N> S;
            {
                QpState<N, MO> qs;
                double* ws = qp_ws + threadIdx.x;
                qp_setup<N, MO, BLOCK>(C, px, vx, py, vy, th, gx, gy, ft, ce, nb, ces + MO, n_stream, dl, lim, ws, qs);
                RO_T(2);
                if (io.warm_start) {
                    bool any = false;                       // nothing carried over (first step of a run): geometric guess
#pragma unroll
                    for (int j = 0; j < 2 * N; ++j) any |= warm[j] >= 0;
                    if (!any || io.start_mode == 1) guess_codes<N, MO>(qs, warm);
                    qp_warm_start<N, MO, BLOCK, true>(C, warm, ws, qs);
                }
                RO_T(3);
                while (!qs.done) qp_trip<N, MO, BLOCK>(C, ws, qs);
                RO_T(4);
                qp_finish<N, MO>(C, qs, S);
                shift_codes<N, MO, BLOCK>(qs, ws, warm, io.start_mode == 2);
                RO_T(5);
            }
            solves += writer ? 1 : 0;
            iters_sum += writer ? S.iters : 0;
            last_status = S.status;
            if (S.status != LDCBF_STATUS_SOLVED) {                                    // :419-429 break
                end_code = S.status == LDCBF_STATUS_DEGENERATE ? LDCBF_END_DEGENERATE
                         : S.status == LDCBF_STATUS_MAX_ITER ? LDCBF_END_MAX_ITER
                         : S.iters == 0 ? LDCBF_END_INFEASIBLE_NOW : LDCBF_END_INFEASIBLE_AHEAD;
                if (writer) io.goal_steps[(size_t)b * n_goals + gi] = kstep;
                ++gi; kstep = 0; last_obj = INFINITY;
#pragma unroll
                for (int j = 0; j < 2 * N; ++j) warm[j] = -1;
                continue;
            }
            last_obj = S.obj;
            ux = S.ux[0]; uy = S.uy[0];
            // x_{k+1} = A x_k + B u_0 (:441-442), evaluated as the reference does
            const double npx = C.ch * px + C.sh_over_beta * vx + (1.0 - C.ch) * ux;
            const double nvx = C.beta_sh * px + C.ch * vx - C.beta_sh * ux;
            const double npy = C.ch * py + C.sh_over_beta * vy + (1.0 - C.ch) * uy;
            const double nvy = C.beta_sh * py + C.ch * vy - C.beta_sh * uy;
            px = npx; vx = nvx; py = npy; vy = nvy;
            th1 = S.th[1]; om0 = S.om[0];
        } else {
            // sub-step: only the heading advances (:443-447)
            const double phi = atan2(gy - py, gx - px);
            om0 = fmin(fmax(phi - th, lim.omega_min), lim.omega_max);
            th1 = __dadd_rn(th, __dmul_rn(om0, C.sampling_time));
        }
        th = th1;
        if (tU) { tU[3 * total] = ux; tU[3 * total + 1] = uy; tU[3 * total + 2] = om0; }
        ++total; ++kstep;
        if (tX) { double* x = tX + 5 * total; x[0] = px; x[1] = vx; x[2] = py; x[3] = vy; x[4] = th; }
    }
#if defined(LDCBF_ROLLOUT_PROFILE) && defined(__CUDA_ARCH__)
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("rollout profile (cycles, thread 0): steps %d iters %d | loop %lld K1 %lld setup %lld warm %lld trips %lld finish %lld\n",
               total, iters_sum, prof_[0], prof_[1], prof_[2], prof_[3], prof_[4], prof_[5]);
#endif
    if (writer) {
        if (gi < n_goals) io.goal_steps[(size_t)b * n_goals + gi] = kstep;
        io.state[5 * (size_t)b] = px; io.state[5 * (size_t)b + 1] = vx; io.state[5 * (size_t)b + 2] = py;
        io.state[5 * (size_t)b + 3] = vy; io.state[5 * (size_t)b + 4] = th;
        io.steps[b] = total;
        io.status[b] = last_status;
        if (io.end_code) io.end_code[b] = (gi < n_goals && total >= T) ? LDCBF_END_BUDGET : end_code;
    }
    if (io.total_solves) {
        // one atomic per warp, whatever subset of its lanes is still here (blocks of 8 threads, ragged last block)
        const unsigned m = __activemask();
        const unsigned s = __reduce_add_sync(m, (unsigned)solves);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(io.total_solves, (unsigned long long)s);
        if (io.total_iters) {
            const unsigned it = __reduce_add_sync(m, (unsigned)iters_sum);
            if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(io.total_iters, (unsigned long long)it);
        }
    }
}

template <int N, int MO, bool EXACT, int BLOCK, int G = 1>
static int launch_rollout_block(const StepConst& C, int B, int T, int n_goals, int msg, int sub, int max_obs,
                                int max_verts, const RolloutIO& io, cudaStream_t st) {
    size_t smem = (size_t)QpWorkspace<N>::DOUBLES * sizeof(double) * BLOCK;
    // per scenario: half-planes (4 per obstacle), vertices (2 per vertex), edge constants (1 per vertex), vertex counts;
    // rounded to 32 B; staged when the block's share stays below 96 KB (several blocks per SM), else read from global
    int map_doubles = 4 * max_obs + 3 * max_obs * max_verts + max_obs;       // ... and two ints per obstacle
    map_doubles = (map_doubles + 3) / 4 * 4;
    const size_t map_bytes = (size_t)map_doubles * sizeof(double) * (BLOCK / G);
    if (smem + map_bytes <= 96 * 1024) smem += map_bytes; else map_doubles = 0;
    auto kern = rollout_kernel<N, MO, EXACT, BLOCK, G>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    }
    kern<<<(unsigned)(((size_t)B * G + BLOCK - 1) / BLOCK), BLOCK, smem, st>>>(C, B, T, n_goals, msg, sub, max_obs, max_verts,
                                                                              map_doubles, io);
    return check_launch();
}

template <int N, int MO>
static int launch_rollout(const StepConst& C, int B, int T, int n_goals, int msg, int sub, int max_obs, int max_verts,
                          const RolloutIO& io, cudaStream_t st) {
    // Lanes per scenario for the ring walk (the solve is then run redundantly by the lanes of a group).  The kernel is a
    // chain of dependent steps per scenario, so more lanes shorten the ring walk — until the warps get in each other's
    // way: every warp streams ~100 KB of code per step, and beyond ~2 warps per SM instruction fetch dominates (DESIGN.md
    // §6).  Measured after the warm-start changes of round 2 (ms per config-2 pass, lanes per scenario 1 / 2 / 4):
    // B = 2048: - / 4.73 / 4.04;  3072: - / 4.79 / 4.89;  4096: 6.38 / 5.00 / 6.07;  6144: 6.39 / 5.80 / -;
    // 8192: 6.46 / 7.26 / -.  Blocks of 16 or 8 threads (more, smaller warps) always lost.
    if (B < 7168 && max_verts > 8) {
        static const int rg_env = getenv("LDCBF_ROLLOUT_G") ? atoi(getenv("LDCBF_ROLLOUT_G")) : 0;
        static const int rb = getenv("LDCBF_ROLLOUT_BLOCK") ? atoi(getenv("LDCBF_ROLLOUT_BLOCK")) : 32;
        const int rg = rg_env ? rg_env : (B < 3072 ? 4 : 2);
        if (io.fast_geometry) {
            if (rg == 2) return launch_rollout_block<N, MO, false, 32, 2>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
            return launch_rollout_block<N, MO, false, 32, 4>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
        }
        if (rb == 8) return launch_rollout_block<N, MO, true, 8, 4>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
        if (rb == 16) return launch_rollout_block<N, MO, true, 16, 4>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
        if (rg == 1) return launch_rollout_block<N, MO, true, 32, 1>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
        if (rg == 2) return launch_rollout_block<N, MO, true, 32, 2>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
        if (rg == 8) return launch_rollout_block<N, MO, true, 32, 8>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
        return launch_rollout_block<N, MO, true, 32, 4>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
    }
    const bool big = B >= 148 * 4 * 128;
    if (io.fast_geometry)
        return big ? launch_rollout_block<N, MO, false, 128>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st)
                   : launch_rollout_block<N, MO, false, 32>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
    return big ? launch_rollout_block<N, MO, true, 128>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st)
               : launch_rollout_block<N, MO, true, 32>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
}

template <int N>
static int dispatch_rollout(const StepConst& C, int B, int T, int n_goals, int msg, int sub, int max_obs,
                            int max_verts, const RolloutIO& io, cudaStream_t st) {
    if (max_obs <= 4) return launch_rollout<N, 4>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
    // up to 8 obstacles are register-resident; further ones are streamed from the scratch the caller of this function
    // allocated (their half-planes are rebuilt there every step)
    return launch_rollout<N, 8>(C, B, T, n_goals, msg, sub, max_obs, max_verts, io, st);
}

}  // namespace ldcbf

extern "C" int ldcbf_rollout_f64(const ldcbf_params* prm, int B, int N, int T, int n_goals, int max_steps_per_goal,
                                 int max_obs, int max_verts, double* state, const double* goals,
                                 const int8_t* right_first, const double* verts, const int32_t* nverts,
                                 const int32_t* nobs, const double* delta, const double* limits, double* traj_X,
                                 double* traj_U, int32_t* steps, int32_t* goal_steps, int32_t* status,
                                 int64_t* total_solves, int64_t* total_iters, int32_t* end_code, void* cuda_stream) {
    using namespace ldcbf;
    if (!prm || B < 0 || T <= 0 || n_goals <= 0 || max_steps_per_goal <= 0 || max_obs <= 0 || max_verts <= 0)
        return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!state || !goals || !right_first || !verts || !nverts || !nobs || !steps || !goal_steps || !status)
        return LDCBF_E_ARG;
    const StepConst C = make_const(*prm);
    int sub = (int)(prm->delta_t / prm->sampling_time);                      // HumanoidMpc.py:74-75
    if (sub <= 0) sub = 1;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // the half-planes of a step go through a stream-ordered scratch from the library's pool, 32 B per (scenario,
    // obstacle): up to 8 are then held in registers, further ones (the reference's CROWDED maps have 20,
    // simulation_1.py:201-231) are streamed from it during the scans
    double4* scratch = nullptr;
    const bool exact = (prm->flags & LDCBF_FLAG_FAST_GEOMETRY) == 0;
    const size_t ce_bytes = sizeof(double4) * (size_t)B * max_obs;
    const size_t den_bytes = exact ? sizeof(double) * (size_t)B * max_obs * max_verts : 0;   // edge constants, exact mode
    {
        cudaMemPool_t pool = workspace_pool();
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&scratch, ce_bytes + den_bytes, pool, st) : cudaErrorMemoryAllocation;
        if (e != cudaSuccess) { set_last_error(e); cudaGetLastError(); return LDCBF_E_LAUNCH; }
    }
    double* den_table = exact ? reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + ce_bytes) : nullptr;
    const RolloutIO io{state, goals, right_first, reinterpret_cast<const double2*>(verts), nverts, nobs, delta, limits,
                       traj_X, traj_U, steps, goal_steps, status,
                       reinterpret_cast<unsigned long long*>(total_solves),
                       (prm->flags & LDCBF_FLAG_FAST_GEOMETRY) != 0, (prm->flags & LDCBF_FLAG_COLD_START) == 0,
                       scratch, reinterpret_cast<unsigned long long*>(total_iters), end_code, getenv("LDCBF_ROLLOUT_NOPRUNE") == nullptr,
                       getenv("LDCBF_ROLLOUT_START") ? atoi(getenv("LDCBF_ROLLOUT_START")) : 2, den_table};
    int rc;
    switch (N) {
        case 1: rc = dispatch_rollout<1>(C, B, T, n_goals, max_steps_per_goal, sub, max_obs, max_verts, io, st); break;
        case 2: rc = dispatch_rollout<2>(C, B, T, n_goals, max_steps_per_goal, sub, max_obs, max_verts, io, st); break;
        case 3: rc = dispatch_rollout<3>(C, B, T, n_goals, max_steps_per_goal, sub, max_obs, max_verts, io, st); break;
        case 4: rc = dispatch_rollout<4>(C, B, T, n_goals, max_steps_per_goal, sub, max_obs, max_verts, io, st); break;
        default: rc = LDCBF_E_SHAPE;
    }
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
}
