// K2+K3 kernel (one thread per scenario) and the C-ABI entry points for one batched MPC step.
// See mpc_qp.cuh for the formulation and the reference lines each piece replaces.
#include <mutex>

#include <cstdio>
#include <cstdlib>

#include "mpc_qp.cuh"
#include "mpc_qp_coop.cuh"
#include "step_io.cuh"

#ifndef LDCBF_QP_REFILL_TRIPS
#define LDCBF_QP_REFILL_TRIPS 1
#endif
// lanes that must be free before a warp stops to retire / refill them (cold-start kernel / resume kernel).  With the
// geometric guess the setup + warm start inside the divergent refill region cost more than the trips (B = 2^20:
// 8 -> 2.31 ms, 16 -> 1.80, 24 -> 1.63, 32 -> 1.99 against 1.90 ms cold), hence the prepare / resume split below.
#ifndef LDCBF_QP_REFILL_MIN_IDLE
#define LDCBF_QP_REFILL_MIN_IDLE 8
#endif
#ifndef LDCBF_QP_RESUME_MIN_IDLE
#define LDCBF_QP_RESUME_MIN_IDLE 16
#endif

namespace ldcbf {

static thread_local cudaError_t g_last_error = cudaSuccess;
void set_last_error(cudaError_t e) { g_last_error = e; }

template <int N>
__device__ __forceinline__ void store_solution(const QpSolution<N>& S, int b, const StepIO& io) {
    if (io.next10) {    // (x_next[4], theta_1, u0_x, u0_y, omega_0, objective, status)
        double2* o = reinterpret_cast<double2*>(io.next10) + 5 * (size_t)b;
        o[0] = make_double2(S.px[1], S.vx[1]);
        o[1] = make_double2(S.py[1], S.vy[1]);
        o[2] = make_double2(S.th[1], S.ux[0]);
        o[3] = make_double2(S.uy[0], S.om[0]);
        o[4] = make_double2(S.obj, (double)S.status);
    }
    if (io.U) {
        double2* U2 = reinterpret_cast<double2*>(io.U) + (size_t)b * N;
        double4* X4 = reinterpret_cast<double4*>(io.X) + (size_t)b * (N + 1);
#pragma unroll
        for (int k = 0; k < N; ++k) U2[k] = make_double2(S.ux[k], S.uy[k]);
#pragma unroll
        for (int k = 0; k <= N; ++k) X4[k] = make_double4(S.px[k], S.vx[k], S.py[k], S.vy[k]);
#pragma unroll
        for (int k = 0; k <= N; ++k) io.theta[(size_t)b * (N + 1) + k] = S.th[k];
#pragma unroll
        for (int k = 0; k < N; ++k) io.omega[(size_t)b * N + k] = S.om[k];
    }
    if (io.obj) io.obj[b] = S.obj;
    if (io.status) io.status[b] = S.status;
    if (io.iters) io.iters[b] = S.iters;
}

// Foot-parity window: read from the foot array, or derived from the first stance foot of the packed state row
// (the reference's s_v alternates strictly, HumanoidMpc.py:104-108, so foot[j] = foot[0] * (-1)^j).
template <int N>
__device__ __forceinline__ void load_foot(const StepIO& io, int b, int (&ft)[N + 1]) {
    if (io.state6) {
        const int f0 = io.state6[6 * (size_t)b + 5] < 0.0 ? -1 : 1;
#pragma unroll
        for (int k = 0; k <= N; ++k) ft[k] = (k & 1) ? -f0 : f0;
    } else {
#pragma unroll
        for (int k = 0; k <= N; ++k) ft[k] = io.foot[(size_t)b * (N + 1) + k];
    }
}

// One thread per scenario; BLOCK threads per block (32 for small batches so the warps spread over all SMs, 128
// otherwise).  Dynamic shared memory: QpWorkspace<N>::DOUBLES doubles per thread, element-major.
template <int N, int MO, int BLOCK>
__global__ void __launch_bounds__(BLOCK) mpc_qp_kernel(StepConst C, int B, int max_obs, StepIO io) {
    extern __shared__ double qp_ws[];
    const int b = blockIdx.x * BLOCK + threadIdx.x;
    if (b >= B) return;
    double4 x;
    double th0;
    load_state(io, b, x, th0);
    const double2 g = reinterpret_cast<const double2*>(io.goal)[b];
    int ft[N + 1];
    load_foot<N>(io, b, ft);
    const Limits lim = load_limits(C, io.limits, (size_t)b);
    const int nt = min(io.nobs[b], max_obs);
    const int nb = min(nt, MO);
    double4 ce[MO];
    const double4* gce = reinterpret_cast<const double4*>(io.c_eta) + (size_t)b * max_obs;
#pragma unroll
    for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? gce[o] : make_double4(0.0, 0.0, 0.0, 0.0);
    QpSolution<N> S;
    solve_scenario<N, MO, BLOCK>(C, x.x, x.y, x.z, x.w, th0, g.x, g.y, ft, ce, nb, gce + MO, nt - nb,
                                 io.delta ? io.delta[b] : 0.0, lim, qp_ws + threadIdx.x, S);
    store_solution<N>(S, b, io);
}

// Large-batch variant: persistent warps with lane refill.  A warp of the plain kernel runs until its slowest lane
// has converged (mean 14.8 trips, maximum over 32 lanes ~30: 14 of 32 lanes active on average, ncu).  Here every
// warp owns a contiguous chunk of scenarios; a lane that has converged stores its result and takes the next
// scenario of the chunk while the other lanes keep iterating, TRIPS trips between two refill points.  No global
// counter, no atomics: the chunk cursor is warp-uniform and lanes rank themselves with a ballot.
// RESUME = true: the scenarios of the chunk [b0, b0 + Bc) were set up and warm-started by mpc_qp_prepare_kernel
// (below); a refill is then a copy of the scenario's record (106 doubles at N = 3, coalesced over the refilling
// lanes) instead of ~2500 instructions of setup and warm start inside this divergent region.
template <int N, int MO, int BLOCK, int TRIPS, bool RESUME, int MIN_IDLE>
__global__ void __launch_bounds__(BLOCK) mpc_qp_refill_kernel(StepConst C, int b0, int Bc, int max_obs, int per_warp,
                                                            StepIO io, const double* __restrict__ rec,
                                                            const int* __restrict__ ids, const int* __restrict__ count) {
    extern __shared__ double qp_ws[];
    double* ws = qp_ws + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    // RESUME: the work list is the compacted set of scenarios the prepare kernel left unfinished — record k belongs to
    // scenario ids[k], k < *count; the chunk cursor then runs over k.  Otherwise it runs over the scenarios themselves.
    if (RESUME) per_warp = (*count + (int)(gridDim.x * (BLOCK / 32)) - 1) / (int)(gridDim.x * (BLOCK / 32));
    const int B = RESUME ? *count : b0 + Bc;
    long long first = (RESUME ? 0LL : (long long)b0) + (long long)warp * per_warp;
    int next = first < B ? (int)first : B;                       // warp-uniform chunk cursor
    const int end = min(B, next + per_warp);
    QpState<N, MO> s;
    s.done = true;
    int b = -1;                                                  // scenario this lane is working on
    for (;;) {
        const unsigned done_m = __ballot_sync(0xffffffffu, b >= 0 && s.done);    // converged, not yet retired
        const unsigned idle_m = __ballot_sync(0xffffffffu, b < 0);
        const bool none_busy = (done_m | idle_m) == 0xffffffffu;
        const bool more = next < end;
        // service point: retire converged lanes and refill free lanes in one (divergent) region, entered only when
        // enough lanes are free to amortise it, or when no lane has anything left to iterate on
        if ((more && __popc(done_m) + __popc(idle_m) >= MIN_IDLE) || none_busy) {
            if (b >= 0 && s.done) {
                QpSolution<N> S;
                qp_finish<N, MO>(C, s, S);
                store_solution<N>(S, b, io);
                b = -1;
            }
            if (more) {
                const unsigned free_m = __ballot_sync(0xffffffffu, b < 0);
                const int cand = next + __popc(free_m & ((1u << lane) - 1u));
                if (b < 0 && cand < end) {
                    b = RESUME ? ids[cand] : cand;
                    double4 x;
                    double th0;
                    load_state(io, b, x, th0);
                    const double2 g = reinterpret_cast<const double2*>(io.goal)[b];
                    int ft[N + 1];
                    load_foot<N>(io, b, ft);
                    const Limits lim = load_limits(C, io.limits, (size_t)b);
                    const int nt = min(io.nobs[b], max_obs);
                    const int nb = min(nt, MO);
                    double4 ce[MO];
                    const double4* gce = reinterpret_cast<const double4*>(io.c_eta) + (size_t)b * max_obs;
#pragma unroll
                    for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? gce[o] : make_double4(0.0, 0.0, 0.0, 0.0);
                    qp_setup<N, MO, BLOCK, RESUME>(C, x.x, x.y, x.z, x.w, th0, g.x, g.y, ft, ce, nb, gce + MO, nt - nb,
                                                   io.delta ? io.delta[b] : 0.0, lim, ws, s,
                                                   RESUME ? rec + cand : nullptr, (size_t)Bc);
                }
                next = min(end, next + __popc(free_m));
            }
        }
        if (__ballot_sync(0xffffffffu, b >= 0) == 0u) break;
#pragma unroll 1
        for (int t = 0; t < TRIPS; ++t) {
            if (b >= 0 && !s.done) qp_trip<N, MO, BLOCK>(C, ws, s);
        }
    }
}

// Large batches, first half: one thread per scenario runs the fixed part of a solve — heading schedule (atan2, N+1
// sincos), bounds, geometric guess, warm start (Gram matrix, Cholesky, sign repair) — and the first K0 trips, every
// warp starting with all lanes busy.  Half of the scenarios are finished by then (p50 of the trip count is 3) and
// store their result here; the others append their solver state as a record to a COMPACT list (warp-aggregated
// atomic; element-major over the list positions, so both sides stay coalesced) which the resume kernel works through
// with lane refill.  Which position a scenario gets depends on the order the warps arrive in, its result does not.
template <int N, int MO, int BLOCK, int K0>
__global__ void __launch_bounds__(BLOCK) mpc_qp_prepare_kernel(StepConst C, int b0, int Bc, int max_obs, StepIO io,
                                                             double* __restrict__ rec, int* __restrict__ ids,
                                                             int* __restrict__ count) {
    extern __shared__ double qp_ws[];
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const bool live = i < Bc;
    const int b = b0 + (live ? i : 0);
    double* ws = qp_ws + threadIdx.x;
    QpState<N, MO> s;
    s.done = true;
    if (live) {
        double4 x;
        double th0;
        load_state(io, b, x, th0);
        const double2 g = reinterpret_cast<const double2*>(io.goal)[b];
        int ft[N + 1];
        load_foot<N>(io, b, ft);
        const Limits lim = load_limits(C, io.limits, (size_t)b);
        const int nt = min(io.nobs[b], max_obs);
        const int nb = min(nt, MO);
        double4 ce[MO];
        const double4* gce = reinterpret_cast<const double4*>(io.c_eta) + (size_t)b * max_obs;
#pragma unroll
        for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? gce[o] : make_double4(0.0, 0.0, 0.0, 0.0);
        qp_setup<N, MO, BLOCK>(C, x.x, x.y, x.z, x.w, th0, g.x, g.y, ft, ce, nb, gce + MO, nt - nb,
                               io.delta ? io.delta[b] : 0.0, lim, ws, s);
        int codes[2 * N];
        guess_codes<N, MO>(s, codes);
        qp_warm_start<N, MO, BLOCK>(C, codes, ws, s);
#pragma unroll 1
        for (int t = 0; t < K0; ++t) {
            if (!s.done) qp_trip<N, MO, BLOCK>(C, ws, s);
        }
        if (s.done) {
            QpSolution<N> S;
            qp_finish<N, MO>(C, s, S);
            store_solution<N>(S, b, io);
        }
    }
    const bool open = live && !s.done;
    const unsigned m = __ballot_sync(0xffffffffu, open);
    if (m) {
        const unsigned lane = threadIdx.x & 31u;
        int base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (open) {
            const int k = base + __popc(m & ((1u << lane) - 1u));
            ids[k] = b;
            qp_dump_state<N, MO, BLOCK>(C, s, ws, rec + k, (size_t)Bc);
        }
    }
}

// Batches that do not fill the GPU: G lanes per scenario (mpc_qp_coop.cuh).  BLOCK / G scenarios per block, each with
// its own slice of shared memory; a group whose scenario index is past the batch leaves at once.
template <int N, int MO, int G, int BLOCK>
__global__ void __launch_bounds__(BLOCK) mpc_qp_coop_kernel(StepConst C, int B, int max_obs, StepIO io) {
    extern __shared__ double qp_ws[];
    const int gi = threadIdx.x / G;
    const int b = blockIdx.x * (BLOCK / G) + gi;
    if (b >= B) return;
    LaneGroup<G> grp;
    grp.lane = threadIdx.x % G;
    grp.mask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) / G * G));
    double4 x;
    double th0;
    load_state(io, b, x, th0);
    const double2 g = reinterpret_cast<const double2*>(io.goal)[b];
    int ft[N + 1];
    load_foot<N>(io, b, ft);
    const Limits lim = load_limits(C, io.limits, (size_t)b);
    const int nobs = min(min(io.nobs[b], max_obs), MO);
    const double4* gce = reinterpret_cast<const double4*>(io.c_eta) + (size_t)b * max_obs;
    QpSolution<N> S;
    coop_solve_scenario<N, MO, G>(C, grp, x.x, x.y, x.z, x.w, th0, g.x, g.y, ft, gce, nobs,
                                  io.delta ? io.delta[b] : 0.0, lim, qp_ws + (size_t)gi * CoopShape<N, MO>::DOUBLES, S);
    if (grp.lane == 0) store_solution<N>(S, b, io);
}

template <int N, int MO, int G, int BLOCK>
static int launch_qp_coop(const StepConst& C, int B, int max_obs, const StepIO& io, cudaStream_t st) {
    constexpr int PER_BLOCK = BLOCK / G;
    const size_t smem = (size_t)CoopShape<N, MO>::DOUBLES * sizeof(double) * PER_BLOCK;
    auto kern = mpc_qp_coop_kernel<N, MO, G, BLOCK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    }
    kern<<<(unsigned)((B + PER_BLOCK - 1) / PER_BLOCK), BLOCK, smem, st>>>(C, B, max_obs, io);
    return check_launch();
}

// Library-owned memory pool (one per device) for the record buffer of the prepare / resume split: stream-ordered
// allocations that are kept between calls (release threshold = max) so that a step never pays for a driver allocation
// after the first one.
static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64] = {};
cudaMemPool_t workspace_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        g_pools[dev] = pool;
    }
    return g_pools[dev];
}

// development aid: LDCBF_COOP_MAX_B overrides the largest batch routed to the cooperative kernel, LDCBF_COOP_G the
// lanes per scenario (8, 16 or 32)
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Small batches, racing variant: every scenario is solved by TWO lanes at once with different (equally exact) paths
// through the active-set method — lane 0 starts from the geometric guess and pivots on velocity rows first, lane 1
// starts cold and pivots on leg rows first — and the first lane to reach the optimum writes the result.  At the
// benchmark batch the kernel time is the trip count of the slowest of 4096 scenarios (the GPU is mostly idle, one
// 8-lane warp per SM sub-partition); the two paths are slow on different scenarios: max 33 / 38 trips alone, 23
// for the faster of the two (p99 25 / 32 -> 18; measured on three seeds of config 2).  A lane that ends without a
// solution (infeasible, iteration cap) waits for its partner, so a status other than "solved" is only reported
// when both paths agree.
template <int N, int MO, int BLOCK>
__global__ void __launch_bounds__(BLOCK) mpc_qp_race_kernel(StepConst C, int B, int max_obs, StepIO io) {
    extern __shared__ double qp_ws[];
    const int b = (blockIdx.x * BLOCK + threadIdx.x) >> 1;
    if (b >= B) return;                                   // both lanes of a pair leave together
    const int racer = threadIdx.x & 1;
    const unsigned pair_mask = 3u << (threadIdx.x & 30);
    double4 x;
    double th0;
    load_state(io, b, x, th0);
    const double2 g = reinterpret_cast<const double2*>(io.goal)[b];
    int ft[N + 1];
    load_foot<N>(io, b, ft);
    const Limits lim = load_limits(C, io.limits, (size_t)b);
    const int nt = min(io.nobs[b], max_obs);
    const int nb = min(nt, MO);
    double4 ce[MO];
    const double4* gce = reinterpret_cast<const double4*>(io.c_eta) + (size_t)b * max_obs;
#pragma unroll
    for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? gce[o] : make_double4(0.0, 0.0, 0.0, 0.0);
    double* ws = qp_ws + threadIdx.x;
    QpState<N, MO> s;
    qp_setup<N, MO, BLOCK>(C, x.x, x.y, x.z, x.w, th0, g.x, g.y, ft, ce, nb, gce + MO, nt - nb,
                           io.delta ? io.delta[b] : 0.0, lim, ws, s);
    s.pref = racer ? 0 : 1;
    if (racer == 0) {
        int codes[2 * N];
        guess_codes<N, MO>(s, codes);
        qp_warm_start<N, MO, BLOCK>(C, codes, ws, s);
    }
    int winner;
    for (;;) {
        if (!s.done) qp_trip<N, MO, BLOCK>(C, ws, s);
        const int mine = s.done ? (s.status == LDCBF_STATUS_SOLVED ? 2 : 1) : 0;
        const int other = __shfl_xor_sync(pair_mask, mine, 1);
        if (mine == 2 || other == 2) { winner = (mine == 2 && (other != 2 || racer == 0)) ? racer : (racer ^ 1); break; }
        if (mine == 1 && other == 1) { winner = 0; break; }
    }
    if (racer == winner) {
        QpSolution<N> S;
        qp_finish<N, MO>(C, s, S);
        store_solution<N>(S, b, io);
    }
}

template <int N, int MO, int BLOCK>
static int launch_qp_race(const StepConst& C, int B, int max_obs, const StepIO& io, cudaStream_t st) {
    const size_t smem = (size_t)QpWorkspace<N>::DOUBLES * sizeof(double) * BLOCK;
    auto kern = mpc_qp_race_kernel<N, MO, BLOCK>;
    kern<<<(unsigned)((2 * (size_t)B + BLOCK - 1) / BLOCK), BLOCK, smem, st>>>(C, B, max_obs, io);
    return check_launch();
}

template <int N, int MO, int BLOCK>
static int launch_qp_block(const StepConst& C, int B, int max_obs, const StepIO& io, cudaStream_t st) {
    const size_t smem = (size_t)QpWorkspace<N>::DOUBLES * sizeof(double) * BLOCK;
    auto kern = mpc_qp_kernel<N, MO, BLOCK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    }
    kern<<<(unsigned)((B + BLOCK - 1) / BLOCK), BLOCK, smem, st>>>(C, B, max_obs, io);
    return check_launch();
}

template <int N, int MO>
static int launch_qp(const StepConst& C, int B, int max_obs, const StepIO& io, cudaStream_t st) {
    // Large batches: 128-thread blocks.  Small batches leave SM sub-partitions idle, so the scenarios are spread
    // over more, narrower warps (8 lanes used per warp): a warp runs until its slowest lane has converged, and the
    // expected maximum iteration count over 8 scenarios is well below that over 32.
    // Opt-in (LDCBF_FLAG_COOP_LANES): a whole warp per scenario with the QR-updated solver of mpc_qp_coop.cuh.
    // Measured (cold start on both sides): 25-35 % faster than one thread per scenario for 64 <= B <= 1024 (B = 512:
    // 61 -> 41 us), equal at B = 4096 with 8 lanes per scenario, slower beyond; no named configuration falls in its
    // range and it has no initial guess yet, so it is not the default.
    if constexpr (N <= 3) {
        static const int coop_max_b = env_int("LDCBF_COOP_MAX_B", 1024);
        static const int coop_g = env_int("LDCBF_COOP_G", 0);
        if (C.coop_lanes && max_obs <= MO && B <= coop_max_b) {
            const int g = coop_g ? coop_g : 32;
            if (g == 32) return launch_qp_coop<N, MO, 32, 32>(C, B, max_obs, io, st);
            if (g == 16) return launch_qp_coop<N, MO, 16, 32>(C, B, max_obs, io, st);
            return launch_qp_coop<N, MO, 8, 32>(C, B, max_obs, io, st);
        }
    }
    if (B >= 148 * 2 * 128 * 4) {
        // persistent grid: 2 blocks of 128 threads per SM (register-limited occupancy), >= 128 scenarios per warp
        constexpr int BLOCK = 128, TRIPS = LDCBF_QP_REFILL_TRIPS;
        const int warps = 148 * 2 * (BLOCK / 32);
        const size_t smem = (size_t)QpWorkspace<N>::DOUBLES * sizeof(double) * BLOCK;
        if (C.cold_start) {
            auto kern = mpc_qp_refill_kernel<N, MO, BLOCK, TRIPS, false, LDCBF_QP_REFILL_MIN_IDLE>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
            kern<<<148 * 2, BLOCK, smem, st>>>(C, 0, B, max_obs, (B + warps - 1) / warps, io, nullptr, nullptr, nullptr);
            return check_launch();
        }
        // prepare (convergent: fixed part + first trips) + resume (lane refill from the records of the unfinished), in
        // chunks of at most 2^20 scenarios so that the stream-ordered record buffer stays below 1 GB
        constexpr int CHUNK = 1 << 20;
        const int Bmax = B < CHUNK ? B : CHUNK;
        char* buf = nullptr;
        cudaMemPool_t pool = workspace_pool();
        const size_t rec_bytes = (size_t)QpRecord<N>::DOUBLES * sizeof(double) * Bmax;
        const size_t ids_bytes = ((size_t)Bmax * sizeof(int) + 15) / 16 * 16;
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&buf, rec_bytes + ids_bytes + 16, pool, st) : cudaErrorMemoryAllocation;
        if (e != cudaSuccess) { set_last_error(e); cudaGetLastError(); return LDCBF_E_LAUNCH; }
        double* rec = reinterpret_cast<double*>(buf);
        int* ids = reinterpret_cast<int*>(buf + rec_bytes);
        int* count = reinterpret_cast<int*>(buf + rec_bytes + ids_bytes);
        static const int k0 = env_int("LDCBF_PREP_TRIPS", 3);     // measured at B = 2^20: 0 -> 1.34 ms, 2 -> 1.40, 3 -> 1.21, 4 -> 1.28
        auto prep = k0 <= 0 ? mpc_qp_prepare_kernel<N, MO, BLOCK, 0>
                  : k0 == 1 ? mpc_qp_prepare_kernel<N, MO, BLOCK, 1>
                  : k0 == 2 ? mpc_qp_prepare_kernel<N, MO, BLOCK, 2>
                  : k0 == 3 ? mpc_qp_prepare_kernel<N, MO, BLOCK, 3> : mpc_qp_prepare_kernel<N, MO, BLOCK, 4>;
        auto kern = mpc_qp_refill_kernel<N, MO, BLOCK, TRIPS, true, LDCBF_QP_RESUME_MIN_IDLE>;
        e = cudaFuncSetAttribute(prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int rc = LDCBF_OK;
        if (e != cudaSuccess) { set_last_error(e); rc = LDCBF_E_LAUNCH; }
        for (int b0 = 0; b0 < B && rc == LDCBF_OK; b0 += CHUNK) {
            const int Bc = (B - b0) < CHUNK ? (B - b0) : CHUNK;
            cudaMemsetAsync(count, 0, sizeof(int), st);
            prep<<<(unsigned)((Bc + BLOCK - 1) / BLOCK), BLOCK, smem, st>>>(C, b0, Bc, max_obs, io, rec, ids, count);
            kern<<<148 * 2, BLOCK, smem, st>>>(C, b0, Bc, max_obs, 0, io, rec, ids, count);
            rc = check_launch();
        }
        cudaFreeAsync(buf, st);
        return rc;
    }
    if (B >= 148 * 4 * 128) return launch_qp_block<N, MO, 128>(C, B, max_obs, io, st);
    if (B >= 148 * 4 * 16) return launch_qp_block<N, MO, 32>(C, B, max_obs, io, st);
    if (!C.cold_start) {
        static const int race_block = env_int("LDCBF_RACE_BLOCK", 0);
        // lanes per block: the fewer scenarios share a warp, the less the warp waits on its slowest lane, as long as
        // every warp still finds an SM sub-partition of its own (measured at 512 / 2048 / 4096 / 8192 scenarios)
        const int rb = race_block ? race_block : (B <= 1024 ? 8 : (B <= 4096 ? 16 : 32));
        if (rb == 32) return launch_qp_race<N, MO, 32>(C, B, max_obs, io, st);
        if (rb == 16) return launch_qp_race<N, MO, 16>(C, B, max_obs, io, st);
        return launch_qp_race<N, MO, 8>(C, B, max_obs, io, st);
    }
    return launch_qp_block<N, MO, 8>(C, B, max_obs, io, st);
}

template <int N>
static int dispatch_obs(const StepConst& C, int B, int max_obs, const StepIO& io, cudaStream_t st) {
    if (max_obs <= 4) return launch_qp<N, 4>(C, B, max_obs, io, st);
    // up to 8 obstacles per scenario live in registers; any further ones are streamed from c_eta during the scan
    return launch_qp<N, 8>(C, B, max_obs, io, st);
}

static int dispatch_horizon(const ldcbf_params& prm, int B, int N, int max_obs, const StepIO& io, void* cuda_stream) {
    const StepConst C = make_const(prm);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    switch (N) {
        case 1: return dispatch_obs<1>(C, B, max_obs, io, st);
        case 2: return dispatch_obs<2>(C, B, max_obs, io, st);
        case 3: return dispatch_obs<3>(C, B, max_obs, io, st);
        case 4: return dispatch_obs<4>(C, B, max_obs, io, st);
        default: return N > LDCBF_MAX_HORIZON ? launch_long_horizon(C, B, N, max_obs, io, st) : LDCBF_E_SHAPE;
    }
}

}  // namespace ldcbf

using namespace ldcbf;

extern "C" int ldcbf_abi_version(void) { return LDCBF_ABI_VERSION; }

extern "C" const char* ldcbf_last_cuda_error(void) { return cudaGetErrorString(g_last_error); }

extern "C" void ldcbf_params_default(ldcbf_params* p) {
    if (!p) return;
    p->delta_t = 0.4; p->gravity = 9.81; p->com_height = 1.0; p->alpha = 3.6;      // config.yml:2-5
    p->l_max_x = 0.10; p->l_max_y = 0.10; p->l_min_x = -0.1; p->l_min_y = -0.1;    // config.yml:6-9
    p->v_min[0] = -0.1; p->v_min[1] = 0.1; p->v_max[0] = 0.8; p->v_max[1] = 0.4;   // config.yml:10-11
    p->omega_max = 0.156 * 3.141592653589793; p->omega_min = -p->omega_max;        // HumanoidMpc.py:21-22
    p->foot_offset = 0.05; p->stop_objective = 0.05;                               // HumanoidMpc.py:200,392
    p->sampling_time = 1e-3;                                                       // HumanoidMpc.py:50
    p->eps_active = 1e-12; p->eps_const_row = 1e-6; p->eps_infeasible = 1e-9;
    p->max_iter = 200; p->flags = 0;
}

extern "C" size_t ldcbf_workspace_bytes(int, int, int, int) { return 0; }

extern "C" int ldcbf_trim_workspace(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return LDCBF_E_ARG;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pools[dev] && cudaMemPoolTrimTo(g_pools[dev], 0) != cudaSuccess) { cudaGetLastError(); return LDCBF_E_LAUNCH; }
    return LDCBF_OK;
}

extern "C" int ldcbf_mpc_qp_f64(const ldcbf_params* prm, int B, int N, int max_obs, const double* x0,
                                const double* theta0, const double* goal, const int8_t* foot, const double* c_eta,
                                const int32_t* nobs, const double* delta, const double* limits, double* U, double* X,
                                double* theta, double* omega, double* obj, int32_t* status, int32_t* iters,
                                void* cuda_stream) {
    if (!prm || B < 0 || max_obs <= 0) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!x0 || !theta0 || !goal || !foot || !c_eta || !nobs || !U || !X || !theta || !omega || !obj || !status || !iters)
        return LDCBF_E_ARG;
    const StepIO io{x0, theta0, goal, foot, c_eta, nobs, delta, limits, U, X, theta, omega, obj, status, iters,
                    nullptr, nullptr};
    return dispatch_horizon(*prm, B, N, max_obs, io, cuda_stream);
}

extern "C" int ldcbf_mpc_step_packed_f64(const ldcbf_params* prm, int B, int N, int max_obs, int max_verts,
                                         const double* state, const double* goal, const double* verts,
                                         const int32_t* nverts, const int32_t* nobs, const double* delta,
                                         const double* limits, double* next, double* c_eta, int32_t* iters,
                                         void* cuda_stream) {
    if (!prm || B < 0 || max_obs <= 0 || max_verts <= 0) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!state || !goal || !verts || !nverts || !nobs || !next || !c_eta) return LDCBF_E_ARG;
    int rc = launch_halfplanes(B, max_obs, max_verts, state, 6, 2, verts, nverts, nobs, c_eta,
                               (prm->flags & LDCBF_FLAG_FAST_GEOMETRY) != 0, cuda_stream);
    if (rc != LDCBF_OK) return rc;
    const StepIO io{nullptr, nullptr, goal, nullptr, c_eta, nobs, delta, limits, nullptr, nullptr, nullptr, nullptr,
                    nullptr, nullptr, iters, state, next};
    return dispatch_horizon(*prm, B, N, max_obs, io, cuda_stream);
}

extern "C" int ldcbf_mpc_step_f64(const ldcbf_params* prm, int B, int N, int max_obs, int max_verts, const double* x0,
                                  const double* theta0, const double* goal, const int8_t* foot, const double* verts,
                                  const int32_t* nverts, const int32_t* nobs, const double* delta, const double* warm,
                                  const double* limits, double* U, double* X, double* theta, double* omega,
                                  double* c_eta, double* obj, int32_t* status, int32_t* iters, void* cuda_stream) {
    (void)warm;
    if (!prm || B < 0 || max_obs <= 0 || max_verts <= 0) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!x0 || !c_eta) return LDCBF_E_ARG;
    // K1 reads the CoM position straight out of the state rows (p_x, v_x, p_y, v_y): stride 4, y at +2.
    int rc = launch_halfplanes(B, max_obs, max_verts, x0, 4, 2, verts, nverts, nobs, c_eta,
                               (prm->flags & LDCBF_FLAG_FAST_GEOMETRY) != 0, cuda_stream);
    if (rc != LDCBF_OK) return rc;
    return ldcbf_mpc_qp_f64(prm, B, N, max_obs, x0, theta0, goal, foot, c_eta, nobs, delta, limits, U, X, theta,
                            omega, obj, status, iters, cuda_stream);
}
