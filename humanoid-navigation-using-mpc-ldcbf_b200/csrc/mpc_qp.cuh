// K2+K3 device code: heading schedule, QP assembly in CoM-position space and the exact dual active-set
// solve for ONE scenario, executed by ONE thread with all problem data in registers.
//
// Reference semantics restated (paths relative to the reference root, HumanoidNavigation/):
//   MPC/HumanoidMpc.py:137-160  heading schedule (current CoM for every k, no angle wrap, theta += omega*Ts)
//   MPC/HumanoidMpc.py:183-202,233-236  leg reachability      k = 0..N-1
//   MPC/HumanoidMpc.py:204-219,238-243  maneuverability       k = 0..N-1 on x_{k+1}, theta_{k+1}, omega_k
//   MPC/HumanoidMpc.py:162-181,245-249  walking velocities    k = 1..N (parity only on the cos term of row 2)
//   MPC/HumanoidMpc.py:252-294 + HumanoidMPCVariants/HumanoidMPCCustomLCBF.py:30-31  LDCBF rows k = 0..N
//   MPC/HumanoidMpc.py:321-333  cost sum_{k=0..N} ||p_k - goal||^2
//   MPC/HumanoidMpc.py:417      optim_prob.solve()  (IPOPT on what is a strictly convex QP)
//   MPC/HumanoidMpc.py:335-343,441-447  one LIP integration step
//
// Formulation.  The reference's unknowns are the footsteps u_k; condensing onto them is
// ill-conditioned because the LIP is unstable (cond(P) = 2.7e4 at N = 3, singular in fp64 for N >= 12,
// SURVEY.md §0).  The same QP written in the future CoM positions w = (p_1..p_N) is benign:
//     v_{k+1} = -v_k + gtil (p_{k+1} - p_k),      gtil = beta*sinh(beta T)/(cosh(beta T) - 1)
//     u_k     = (p_{k+1} - cosh(beta T) p_k - sinh(beta T)/beta v_k) / (1 - cosh(beta T))
// are exact consequences of x_{k+1} = A x_k + B u_k, the cost becomes ||w - (g,..,g)||^2 (Hessian 2I) and
// every constraint row is (scalar pattern over k) x (unit 2-vector): the QP is the Euclidean projection of
// the stacked goal onto a polytope.  It is solved exactly with the Goldfarb-Idnani dual active-set
// method.  With Hessian I the step directions need only the Gram matrix of the active normals, kept in
// 2N fixed slots (free slot = zero normal, unit diagonal) so that the 2N x 2N Cholesky, both triangular
// solves and every update are fully unrolled with static register indices: no local memory, and the
// scenarios of a warp diverge only in their iteration count.
// The maneuverability row k and the longitudinal walking-velocity row k+1 are the same linear form
// (cos theta_{k+1}, sin theta_{k+1}) . v_{k+1}; they are merged into one row with the tighter upper bound.
#pragma once
#include "ldcbf_common.cuh"

namespace ldcbf {

// The solver source also compiles for the host (tests/cpu_harness builds it into a test-only library so the
// algorithm is exercised by the CPU test suite); the product library only ever launches it on the device.
#define LDCBF_HD __host__ __device__ __forceinline__
#if defined(LDCBF_HOST_COUNTERS)
static long long ldcbf_host_dependent_guesses = 0;      // analysis aid of the host build (tests/cpu_harness)
#endif
LDCBF_HD double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
LDCBF_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
// 1/sqrt(x) for the Cholesky pivots (x > 0, normal range): hardware seed (MUFU.RSQ64H, ~20 bits) and two
// Newton steps, ~10 instructions instead of the ~20 of the library rsqrt with its special-case handling.
LDCBF_HD double rsqrt_f64(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * (1.5 - hx * y * y);
    y = y * (1.5 - hx * y * y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
LDCBF_HD int first_free_slot(unsigned amask) {
#ifdef __CUDA_ARCH__
    return __ffs(~amask) - 1;
#else
    return __builtin_ffs((int)~amask) - 1;
#endif
}
LDCBF_HD unsigned nonneg_bit(double m) {      // 1 when the sign bit of m is clear
#ifdef __CUDA_ARCH__
    return ((unsigned)__double2hiint(m) >> 31) ^ 1u;
#else
    return m < 0.0 ? 0u : 1u;
#endif
}
LDCBF_HD double quiet_nan() {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}

// Output of one solve, kept in registers by the caller.
template <int N>
struct QpSolution {
    double px[N + 1], py[N + 1];   // CoM positions p_0..p_N
    double vx[N + 1], vy[N + 1];   // CoM velocities
    double ux[N], uy[N];           // footsteps
    double th[N + 1], om[N];       // heading schedule
    double obj;
    int status, iters;
};

// Per-scenario limits (defaults from ldcbf_params, optionally overridden by the `limits` array of the ABI:
// limits[b] = (ALPHA, V_MAX[0], V_MAX[1], OMEGA_MAX, OMEGA_MIN, reserved), NaN = keep the default).
struct Limits {
    double alpha_over_pi, vmax0, vmax1, omega_max, omega_min;
};
LDCBF_HD Limits load_limits(const StepConst& C, const double* limits, size_t b) {
    Limits L{C.alpha_over_pi, C.v_max0, C.v_max1, C.omega_max, C.omega_min};
    if (limits) {
        const double* q = limits + 6 * b;
        if (q[0] == q[0]) L.alpha_over_pi = q[0] / 3.141592653589793;
        if (q[1] == q[1]) L.vmax0 = q[1];
        if (q[2] == q[2]) L.vmax1 = q[2];
        if (q[3] == q[3]) L.omega_max = q[3];
        if (q[4] == q[4]) L.omega_min = q[4];
    }
    return L;
}

// Row identifiers: leg(k,sub) = 2k+sub (k<N); vel(k,sub) = 2N + 2(k-1) + sub (k=1..N);
// cbf(k,o) = 4N + (k-1)*MO + o (k=1..N).  A candidate is (id, sg): sg=+1 means a.w >= lo, -1 means -a.w >= -hi.
//
// Dense signed normal in p-space, a[2i], a[2i+1] = coefficient on p_{i+1}:
//   leg(k):  +r on p_{k+1}, -r on p_k                      r = (c_k, s_k) or (-s_k, c_k)
//   vel(k):  gtil r on p_k, -+2 gtil r on p_{k-1}, ..      r = (c_k, s_k) or (-s_k, foot_k c_k)
//   cbf(k):  eta_o on p_k
template <int N, int MO>
LDCBF_HD void row_normal(int id, double sg, double gtil, const double (&rc)[N + 1],
                                           const double (&rs)[N + 1], const int (&ft)[N + 1],
                                           const double (&ex)[MO], const double (&ey)[MO], const double4* ces, int ns,
                                           double (&a)[2 * N]) {
    // decode without dynamic register indexing: select chains over the (static) k and o
    int typ, k, sub;   // typ 0 leg, 1 vel, 2 cbf (registers), 3 cbf (streamed); k = state index the row "ends" at
    if (id < 2 * N) { typ = 0; k = (id >> 1) + 1; sub = id & 1; }
    else if (id < 4 * N) { typ = 1; k = ((id - 2 * N) >> 1) + 1; sub = id & 1; }
    else if (id < 4 * N + N * MO) { typ = 2; k = (id - 4 * N) / MO + 1; sub = (id - 4 * N) - (k - 1) * MO; }
    else { typ = 3; const int j = id - (4 * N + N * MO); k = j / ns + 1; sub = j - (k - 1) * ns; }
    const int kth = (typ == 0) ? k - 1 : k;     // heading index used by the row
    double c = 0.0, s = 0.0, f = 1.0;
#pragma unroll
    for (int j = 0; j <= N; ++j) if (j == kth) { c = rc[j]; s = rs[j]; f = (double)ft[j]; }
    double rx, ry;
    if (typ == 3) {
        const double4 ce = ces[sub];
        rx = ce.z; ry = ce.w;
    } else if (typ == 2) {
        rx = 0.0; ry = 0.0;
#pragma unroll
        for (int o = 0; o < MO; ++o) if (o == sub) { rx = ex[o]; ry = ey[o]; }
    } else if (sub == 0) { rx = c; ry = s; }
    else { rx = -s; ry = (typ == 1) ? f * c : c; }
    rx *= sg; ry *= sg;
#pragma unroll
    for (int i = 0; i < N; ++i) {               // coefficient pattern on p_{i+1}
        const int d = k - 1 - i;                // 0 on the row's own state
        double kap = 0.0;
        if (d == 0) kap = (typ == 1) ? gtil : 1.0;
        else if (d > 0) {
            if (typ == 0) kap = (d == 1) ? -1.0 : 0.0;
            else if (typ == 1) kap = (d & 1) ? -2.0 * gtil : 2.0 * gtil;
        }   // typ 2, 3: only the row's own state
        a[2 * i] = kap * rx; a[2 * i + 1] = kap * ry;
    }
}

// Per-thread workspace of the solver: the signed normals of the 2N slots and their Gram matrix, 2*(2N)^2 doubles.
// On the device it lives in shared memory, element e of this thread at ws[e * WS] with WS = threads per block
// (consecutive threads -> consecutive 8-byte words: conflict-free); a slot is then written with a dynamic
// address instead of a chain of predicated register moves, and ~110 registers are freed.
template <int N>
struct QpWorkspace { static constexpr int DOUBLES = 2 * (2 * N) * (2 * N) + 2 * N; };   // normals, Gram, row code per slot

// Solver state of one scenario, kept in registers across trips of the active-set loop.
template <int N, int MO>
struct QpState {
    double rc[N + 1], rs[N + 1];          // cos / sin of the heading schedule
    double th[N + 1], om[N];
    int ft[N + 1];                        // foot parity window
    double ex[MO], ey[MO], hb[MO];        // half-planes of the first MO obstacles, in registers: eta . p >= hb
    int nb;
    const double4* ces;                   // obstacles beyond MO are streamed from global memory (c, eta) per scan
    int ns;                               // how many of them
    double delta;
    double vmid[N + 1], vhalf[N + 1];     // merged longitudinal velocity row at state k: mid +- half
    double vlat_mid, vlat_half;           // lateral velocity row: [V_MIN1, V_MAX1]
    double p0x, p0y, v0x, v0y, gx, gy;
    double px[N + 1], py[N + 1];          // current iterate w = (p_1..p_N), p_0 fixed
    double u[2 * N];                      // multipliers per slot
    double np[2 * N], nn, s_p, u_p;       // row being added: signed normal, |n|^2, slack, multiplier
    int p_code;                           // ... and its row code 2*id + (upper side)
    unsigned amask;                       // occupied slots
    int pref;                             // pivot preference of the scan: -1 none, 0 leg rows first, 1 velocity rows first
    double tol;                           // a row counts as violated below -tol: eps_active, eps_infeasible after a relaxed restart
    int status, iters;
    bool need_scan, done;
};

#define AN(j, i) ws[((j) * NV + (i)) * WS]
#define GM(i, j) ws[(NV * NV + (i) * NV + (j)) * WS]
#define RC(j) ws[(2 * NV * NV + (j)) * WS]      // row code of slot j: 2*id + (1 if the upper side), as a double

// Heading schedule, half-plane offsets, row bounds, unconstrained optimum, empty active set.
// ce[o] = (c_x, c_y, eta_x, eta_y) for o < nb.
//
// RESUME = true (large batches, mpc_step.cu): the heading schedule, the iterate, the multipliers and the slot
// workspace are not computed but read back from the record `rec` (element e of this scenario at rec[e * stride]) that
// the prepare kernel wrote with qp_dump_state after its own qp_setup + warm start; everything else (inputs, bounds,
// half-plane offsets, status checks) is recomputed from the inputs, which costs a few dozen instructions.
template <int N>
struct QpRecord {       // layout of a record, in doubles
    static constexpr int TH = 0, OM = TH + N + 1, RC_ = OM + N, RS_ = RC_ + N + 1, W = RS_ + N + 1, U = W + 2 * N,
                         NP = U + 2 * N, NN = NP + 2 * N, SP = NN + 1, UP = SP + 1, PACKED = UP + 1, WS0 = PACKED + 1,
                         DOUBLES = WS0 + QpWorkspace<N>::DOUBLES;
};

template <int N, int MO, int WS, bool RESUME = false>
LDCBF_HD void qp_setup(const StepConst& C, double p0x, double v0x, double p0y, double v0y, double th0, double gx,
                       double gy, const int (&ft)[N + 1], const double4 (&ce)[MO], int nb, const double4* ce_stream,
                       int n_stream, double delta, const Limits& lim, double* ws, QpState<N, MO>& s,
                       const double* rec = nullptr, size_t stride = 0) {
    const double alpha_over_pi = lim.alpha_over_pi, vmax0 = lim.vmax0, omega_max = lim.omega_max, omega_min = lim.omega_min;
    s.vlat_mid = 0.5 * (lim.vmax1 + C.v_min1); s.vlat_half = 0.5 * (lim.vmax1 - C.v_min1);
    constexpr int NV = 2 * N;
    s.p0x = p0x; s.p0y = p0y; s.v0x = v0x; s.v0y = v0y; s.gx = gx; s.gy = gy; s.nb = nb;
    s.ces = ce_stream; s.ns = n_stream; s.delta = delta;
#pragma unroll
    for (int k = 0; k <= N; ++k) s.ft[k] = ft[k];
    // ---- heading schedule (HumanoidMpc.py:137-160)
    if (RESUME) {
        using R = QpRecord<N>;
#pragma unroll
        for (int k = 0; k <= N; ++k) {
            s.th[k] = rec[(R::TH + k) * stride]; s.rc[k] = rec[(R::RC_ + k) * stride]; s.rs[k] = rec[(R::RS_ + k) * stride];
        }
#pragma unroll
        for (int k = 0; k < N; ++k) s.om[k] = rec[(R::OM + k) * stride];
    } else {
        const double phi = atan2(gy - p0y, gx - p0x);
        double thk = th0;
        s.th[0] = thk;
        sincos(thk, &s.rs[0], &s.rc[0]);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double w = fmin(fmax(phi - thk, omega_min), omega_max);
            s.om[k] = w;
            thk = add_rn(thk, mul_rn(w, C.sampling_time));
            s.th[k + 1] = thk;
            sincos(thk, &s.rs[k + 1], &s.rc[k + 1]);
        }
    }
    // ---- half-planes: eta_o . p >= eta_o . c_o + delta
    int status = LDCBF_STATUS_SOLVED;
#pragma unroll
    for (int o = 0; o < MO; ++o) {
        s.ex[o] = 0.0; s.ey[o] = 0.0; s.hb[o] = -INFINITY;    // absent obstacle: slack 0*p - (-inf) = +inf, never chosen
        if (o < nb) {
            s.ex[o] = ce[o].z; s.ey[o] = ce[o].w;
            s.hb[o] = ce[o].z * ce[o].x + ce[o].w * ce[o].y + delta;
            if (!(s.ex[o] == s.ex[o]) || !(s.ey[o] == s.ey[o])) status = LDCBF_STATUS_DEGENERATE;
            // constant k = 0 row (HumanoidMpc.py:284-292 with k = 0)
            else if (s.ex[o] * p0x + s.ey[o] * p0y - s.hb[o] < -C.eps_const_row) status = LDCBF_STATUS_INFEASIBLE;
        }
    }
    for (int o = 0; o < n_stream; ++o) {      // streamed obstacles: same two checks
        const double4 c4 = ce_stream[o];
        if (!(c4.z == c4.z) || !(c4.w == c4.w)) status = LDCBF_STATUS_DEGENERATE;
        else if (c4.z * (p0x - c4.x) + c4.w * (p0y - c4.y) - delta < -C.eps_const_row && status == LDCBF_STATUS_SOLVED)
            status = LDCBF_STATUS_INFEASIBLE;
    }
    // merged longitudinal velocity row at state k: [V_MIN0, min(V_MAX0, V_MAX0 - alpha/pi |omega_{k-1}|)]
    s.vmid[0] = 0.0; s.vhalf[0] = 0.0;
#pragma unroll
    for (int k = 1; k <= N; ++k) {
        const double vhi = fmin(vmax0, vmax0 - alpha_over_pi * fabs(s.om[k - 1]));
        s.vmid[k] = 0.5 * (vhi + C.v_min0);
        s.vhalf[k] = 0.5 * (vhi - C.v_min0);
    }
    // ---- Goldfarb-Idnani dual active set on  min 1/2 ||w - g||^2  s.t. rows: start at the unconstrained optimum
    s.px[0] = p0x; s.py[0] = p0y;
#pragma unroll
    for (int k = 1; k <= N; ++k) { s.px[k] = gx; s.py[k] = gy; }
    s.amask = 0;
    s.pref = -1;
    s.tol = C.eps_active;
    // AN(j, .): signed normal of slot j (zero when free);  GM(i, j): Gram matrix of the slots (identity on free slots)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        s.u[j] = 0.0; s.np[j] = 0.0;
        if (!RESUME) {
#pragma unroll
            for (int i = 0; i < NV; ++i) { AN(j, i) = 0.0; GM(j, i) = (i == j) ? 1.0 : 0.0; }
            RC(j) = -1.0;
        }
    }
    s.nn = 1.0; s.s_p = 0.0; s.u_p = 0.0; s.p_code = 0;
    s.iters = 0;
    s.need_scan = true;
    s.status = status;
    s.done = status != LDCBF_STATUS_SOLVED;
    if (RESUME) {
        using R = QpRecord<N>;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            s.px[k] = rec[(R::W + 2 * (k - 1)) * stride]; s.py[k] = rec[(R::W + 2 * (k - 1) + 1) * stride];
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) s.u[j] = rec[(R::U + j) * stride];
        // amask | need_scan << 8 | relaxed << 9 | p_code << 10 | iters << 24 (exact in a double)
        const long long packed = (long long)rec[R::PACKED * stride];
        s.amask = (unsigned)(packed & 0xff);
        s.need_scan = ((packed >> 8) & 1) != 0;
        if ((packed >> 9) & 1) s.tol = C.eps_infeasible;
        s.p_code = (int)((packed >> 10) & 0x3fff);
        s.iters = (int)(packed >> 24);
#pragma unroll
        for (int j = 0; j < NV; ++j) s.np[j] = rec[(R::NP + j) * stride];
        s.nn = rec[R::NN * stride]; s.s_p = rec[R::SP * stride]; s.u_p = rec[R::UP * stride];
#pragma unroll
        for (int e = 0; e < QpWorkspace<N>::DOUBLES; ++e) ws[e * WS] = rec[(R::WS0 + e) * stride];
    }
}

// The counterpart of qp_setup<RESUME>: what a scenario's solver state consists of after setup (+ warm start).
template <int N, int MO, int WS>
LDCBF_HD void qp_dump_state(const StepConst& C, const QpState<N, MO>& s, const double* ws, double* rec, size_t stride) {
    using R = QpRecord<N>;
    constexpr int NV = 2 * N;
#pragma unroll
    for (int k = 0; k <= N; ++k) {
        rec[(R::TH + k) * stride] = s.th[k]; rec[(R::RC_ + k) * stride] = s.rc[k]; rec[(R::RS_ + k) * stride] = s.rs[k];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) rec[(R::OM + k) * stride] = s.om[k];
#pragma unroll
    for (int k = 1; k <= N; ++k) {
        rec[(R::W + 2 * (k - 1)) * stride] = s.px[k]; rec[(R::W + 2 * (k - 1) + 1) * stride] = s.py[k];
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) rec[(R::U + j) * stride] = s.u[j];
#pragma unroll
    for (int j = 0; j < NV; ++j) rec[(R::NP + j) * stride] = s.np[j];
    rec[R::NN * stride] = s.nn; rec[R::SP * stride] = s.s_p; rec[R::UP * stride] = s.u_p;
    rec[R::PACKED * stride] = (double)((long long)s.amask | ((long long)(s.need_scan ? 1 : 0) << 8) |
                                       ((long long)(s.tol > C.eps_active ? 1 : 0) << 9) | ((long long)(s.p_code & 0x3fff) << 10) |
                                       ((long long)s.iters << 24));
#pragma unroll
    for (int e = 0; e < QpWorkspace<N>::DOUBLES; ++e) rec[(R::WS0 + e) * stride] = ws[e * WS];
}

// (value, index) tournament over sl[LO .. LO+LEN): the minimum ends in sl[LO], its index in ti[LO]
template <int LO, int LEN, int NR>
LDCBF_HD void argmin_range(double (&sl)[NR], int (&ti)[NR]) {
#pragma unroll
    for (int n = LEN; n > 1; n = (n + 1) / 2) {
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const int a = LO + i, o = LO + n - 1 - i;
            const bool take = sl[o] < sl[a];
            sl[a] = take ? sl[o] : sl[a];
            ti[a] = take ? ti[o] : ti[a];
        }
    }
}

// One trip of the active-set loop: (scan for the most violated row if the previous trip ended with a full step,)
// one primal/dual step, one slot update.  Every lane of a warp runs the same instruction stream; lanes differ
// only in how many trips they need.  Sets s.done on convergence or failure.
//
// Each trip factorises the slots' Gram matrix from scratch (fully unrolled NV x NV Cholesky): r = G^-1 (N^T n+),
// z = n+ - N r.  Updating the inverse instead (one rank-1 update per add / drop) was measured 30 % faster but
// loses the active set on ill-conditioned vertices (two nearly anti-parallel velocity rows, multipliers ~3e4):
// a false "infeasible" in 1 of 4096 scenarios.
template <int N, int MO, int WS>
LDCBF_HD void qp_trip(const StepConst& C, double* ws, QpState<N, MO>& s) {
    constexpr int NV = 2 * N;
    if (s.need_scan) {
        // -- most violated row at the current point (natural units: m, m/s).
        // A two-sided row lo <= v <= hi has slack  half - |v - mid|  (one value for both sides; the side is the
        // sign of v - mid, collected as a bit).  The NR slacks are reduced by a tournament over (value, index)
        // pairs (fmin on doubles costs ~7 instructions on sm_100: NaN handling, no DMNMX).
        constexpr int NR = 4 * N + N * MO;
        double sl[NR];
        unsigned upper = 0;                 // bit i: row i is on the v > mid side
        {
            double Vx = s.v0x, Vy = s.v0y;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double dx = s.px[k + 1] - s.px[k], dy = s.py[k + 1] - s.py[k];
                const double off = (double)s.ft[k] * C.foot_offset;
                double m;
                m = (s.rc[k] * dx + s.rs[k] * dy) - C.legx_mid;
                sl[2 * k] = C.legx_half - fabs(m);
                upper |= nonneg_bit(m) << (2 * k);
                m = (s.rc[k] * dy - s.rs[k] * dx) - (C.legy_mid - off);
                sl[2 * k + 1] = C.legy_half - fabs(m);
                upper |= nonneg_bit(m) << (2 * k + 1);
                Vx = C.gtil * dx - Vx; Vy = C.gtil * dy - Vy;       // v_{k+1}
                const int kk = k + 1;
                m = (s.rc[kk] * Vx + s.rs[kk] * Vy) - s.vmid[kk];
                sl[2 * N + 2 * k] = s.vhalf[kk] - fabs(m);
                upper |= nonneg_bit(m) << (2 * N + 2 * k);
                m = ((double)s.ft[kk] * s.rc[kk] * Vy - s.rs[kk] * Vx) - s.vlat_mid;
                sl[2 * N + 2 * k + 1] = s.vlat_half - fabs(m);
                upper |= nonneg_bit(m) << (2 * N + 2 * k + 1);
#pragma unroll
                for (int o = 0; o < MO; ++o)
                    sl[4 * N + k * MO + o] = s.ex[o] * s.px[kk] + s.ey[o] * s.py[kk] - s.hb[o];
            }
        }
        // tournament argmin: (value, index) pairs, one compare and three selects per node — one tournament per row
        // family (leg [0,2N), velocity [2N,4N), LDCBF [4N,NR)), then the three winners
        int ti[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) ti[i] = i;
        argmin_range<0, 2 * N>(sl, ti);
        argmin_range<2 * N, 2 * N>(sl, ti);
        argmin_range<4 * N, N * MO>(sl, ti);
        // Pivot preference (the racing kernel runs the same scenario with two different preferences, mpc_step.cu):
        // pref = 0 takes the most violated LEG row while any leg row is violated, pref = 1 the most violated
        // VELOCITY row, pref < 0 (default) the most violated row overall.  Any violated row is a valid pivot of the
        // dual method; the choice only changes the path, not the optimum.
        const double b_leg = sl[0], b_vel = sl[2 * N], b_cbf = sl[4 * N];
        const int i_leg = ti[0], i_vel = ti[2 * N], i_cbf = ti[4 * N];
        double best = b_leg;
        int bid = i_leg;
        if (b_vel < best) { best = b_vel; bid = i_vel; }
        if (b_cbf < best) { best = b_cbf; bid = i_cbf; }
        const bool force_leg = s.pref == 0 && b_leg < -s.tol, force_vel = s.pref == 1 && b_vel < -s.tol;
        // obstacles beyond the register-resident MO: one streaming pass over (c, eta) in global memory (L1/L2)
        for (int o = 0; o < s.ns; ++o) {
            const double4 c4 = s.ces[o];
            const double hbo = c4.z * c4.x + c4.w * c4.y + s.delta;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const double v = c4.z * s.px[k] + c4.w * s.py[k] - hbo;
                if (v < best) { best = v; bid = 4 * N + N * MO + (k - 1) * s.ns + o; }
            }
        }
        if (!(best < -s.tol)) { s.done = true; return; }   // primal feasible: optimal
        if (force_leg) { best = b_leg; bid = i_leg; }
        if (force_vel) { best = b_vel; bid = i_vel; }
        const double bsg = (bid < 4 * N && ((upper >> bid) & 1u)) ? -1.0 : 1.0;
        row_normal<N, MO>(bid, bsg, C.gtil, s.rc, s.rs, s.ft, s.ex, s.ey, s.ces, s.ns, s.np);
        double nn = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) nn += s.np[i] * s.np[i];
        s.nn = nn; s.s_p = best; s.u_p = 0.0;
        s.p_code = 2 * bid + (bsg < 0.0 ? 1 : 0);
        s.need_scan = false;
    }
    if (++s.iters > C.max_iter) { s.status = LDCBF_STATUS_MAX_ITER; s.done = true; return; }

    // d = N^T n+ ;  r = G^-1 d ;  z = n+ - N r
    double d[NV], r[NV], z[NV], zz = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) acc += AN(j, i) * s.np[i];
        d[j] = acc;
    }
    {
        double L[NV][NV];      // Cholesky factor, inverse diagonal kept in L[j][j]
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double dj = GM(j, j);
#pragma unroll
            for (int l = 0; l < j; ++l) dj -= L[j][l] * L[j][l];
            const double inv = rsqrt_f64(fmax(dj, 1e-300));
            L[j][j] = inv;
#pragma unroll
            for (int i = j + 1; i < NV; ++i) {
                double v = GM(i, j);
#pragma unroll
                for (int l = 0; l < j; ++l) v -= L[i][l] * L[j][l];
                L[i][j] = v * inv;
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double v = d[j];
#pragma unroll
            for (int l = 0; l < j; ++l) v -= L[j][l] * r[l];
            r[j] = v * L[j][j];
        }
#pragma unroll
        for (int j = NV - 1; j >= 0; --j) {
            double v = r[j];
#pragma unroll
            for (int l = j + 1; l < NV; ++l) v -= L[l][j] * r[l];
            r[j] = v * L[j][j];
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double v = s.np[i];
#pragma unroll
        for (int j = 0; j < NV; ++j) v -= r[j] * AN(j, i);
        z[i] = v;
        zz += v * v;
    }
    const bool dependent = !(zz > 1e-13 * s.nn) || s.amask == (1u << NV) - 1u;
    // dual step length t1 = min u_j / r_j over r_j > 0, compared by cross-multiplication
    // as a tournament over (numerator, denominator, slot) triples — depth log2(NV) instead of a chain of NV
    // dependent compare-and-selects; (1, 0) stands for "no candidate" and loses against every real one
    double tn[NV], td[NV];
    int tj[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const bool cand = ((s.amask >> j) & 1u) && r[j] > 1e-13;
        tn[j] = cand ? s.u[j] : 1.0; td[j] = cand ? r[j] : 0.0; tj[j] = cand ? j : -1;
    }
#pragma unroll
    for (int n = NV; n > 1; n = (n + 1) / 2) {
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const int o = n - 1 - i;
            const bool take = tn[o] * td[i] < tn[i] * td[o];
            tn[i] = take ? tn[o] : tn[i]; td[i] = take ? td[o] : td[i]; tj[i] = take ? tj[o] : tj[i];
        }
    }
    const double t1n = tn[0], t1d = td[0];          // t1 = t1n / t1d, "infinite" while t1d == 0
    const int ldrop = tj[0];
    // full step t2 = -s_p / zz
    const bool full = !dependent && (ldrop < 0 || (-s.s_p) * t1d <= t1n * zz);
    if (!full && ldrop < 0) {
        // Row p cannot be satisfied together with the active rows.  When it is only violated by rounding-level amounts
        // (<= eps_infeasible, 1e-9 m or m/s: kinematic rows that meet as equalities leave a feasible set that is a
        // single point to rounding — IPOPT with constr_viol_tol 1e-5 and the NNLS oracle both return that point) the
        // solve is restarted ONCE from the empty active set with rows counted as violated only below -eps_infeasible;
        // otherwise the QP is infeasible (the reference's IPOPT raises, HumanoidMpc.py:419-429).
        if (s.tol < C.eps_infeasible && -s.s_p <= C.eps_infeasible) {
            s.tol = C.eps_infeasible;
#pragma unroll
            for (int k = 1; k <= N; ++k) { s.px[k] = s.gx; s.py[k] = s.gy; }
            s.amask = 0;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                s.u[j] = 0.0; s.np[j] = 0.0;
#pragma unroll
                for (int i = 0; i < NV; ++i) { AN(j, i) = 0.0; GM(j, i) = (i == j) ? 1.0 : 0.0; }
                RC(j) = -1.0;
            }
            s.nn = 1.0; s.s_p = 0.0; s.u_p = 0.0; s.p_code = 0;
            s.need_scan = true;
            return;
        }
        s.status = LDCBF_STATUS_INFEASIBLE; s.done = true; return;
    }
    // one division, not two (a 123-cycle chain on B200).  A reciprocal seed + two Newton steps (60 cycles, 1-2 ulp)
    // was tried: 2 % faster, but a full step then no longer lands its row on the bound to the last bit, and the
    // closed loop of config 1 with delta = 0 ended on an obstacle edge (status 3) at step 31 instead of at the goal.
    const double t = (full ? -s.s_p : t1n) / (full ? zz : t1d);
#pragma unroll
    for (int j = 0; j < NV; ++j) s.u[j] -= t * r[j];
    s.u_p += t;
    if (!dependent) {
#pragma unroll
        for (int k = 1; k <= N; ++k) { s.px[k] += t * z[2 * (k - 1)]; s.py[k] += t * z[2 * (k - 1) + 1]; }
        s.s_p += t * zz;
    }
    // -- slot bookkeeping: row p enters the first free slot (full step) or slot ldrop leaves (partial step)
    const int slot = full ? first_free_slot(s.amask) : ldrop;
    if (full) { s.amask |= 1u << slot; s.need_scan = true; }
    else s.amask &= ~(1u << slot);
#pragma unroll
    for (int j = 0; j < NV; ++j) if (j == slot) s.u[j] = full ? s.u_p : 0.0;
    {
        double* an = &AN(slot, 0);             // dynamic slot address
        double* grow = &GM(slot, 0);
        double* gcol = &GM(0, slot);
#pragma unroll
        for (int i = 0; i < NV; ++i) an[i * WS] = full ? s.np[i] : 0.0;
#pragma unroll
        for (int l = 0; l < NV; ++l) {
            const double val = full ? d[l] : 0.0;   // d[slot] = 0 for a free slot; the diagonal is set below
            grow[l * WS] = val;
            gcol[l * NV * WS] = val;
        }
        grow[slot * WS] = full ? s.nn : 1.0;
        (&RC(0))[slot * WS] = full ? (double)s.p_code : -1.0;
    }
}

// Initial active-set guess for a solve with no history (open-loop batches, first step of a loop): all 2N velocity
// rows, each on the side the goal lies on (r . (goal - p_0) > 0 -> upper bound).  Measured on config 2 (4096
// scenarios): the optimum is a vertex with ~6 active rows, three quarters of them velocity rows, and the cold start from
// the unconstrained optimum wanders (adds late-stage rows first, drops them again); this guess, repaired by
// qp_warm_start, replaces 14.8 trips per solve by 5.1 with identical results.  Any guess is safe: qp_warm_start
// keeps the method exact.  LDCBF_FLAG_COLD_START disables it.
template <int N, int MO>
LDCBF_HD void guess_codes(const QpState<N, MO>& s, int (&codes)[2 * N]) {
    const double dgx = s.gx - s.p0x, dgy = s.gy - s.p0y;
#pragma unroll
    for (int k = 1; k <= N; ++k) {
        const double lon = s.rc[k] * dgx + s.rs[k] * dgy;
        const double lat = (double)s.ft[k] * s.rc[k] * dgy - s.rs[k] * dgx;
        codes[2 * (k - 1)] = 2 * (2 * N + 2 * (k - 1)) + (lon > 0.0 ? 1 : 0);
        codes[2 * (k - 1) + 1] = 2 * (2 * N + 2 * (k - 1) + 1) + (lat > 0.0 ? 1 : 0);
    }
}

// Warm start of the dual active-set method from a guessed active set (closed loop: the final active set of the
// previous MPC step shifted by one stage, `shift_codes`).  Goldfarb-Idnani may start from any point that is the
// optimum of the equality-constrained problem on a set of independent rows with non-negative multipliers, so:
// load the guessed rows into the slots, u = G^-1 (b - N^T g) (one Cholesky), drop every row with u_j < 0 and solve
// again (at most NV rounds), w = g + N u.  A guess that turns out dependent falls back to the cold start.  The
// result is exact whatever the guess; a good guess replaces ~15 trips by 2-4.
// codes[j] = 2*id + (upper side) or -1.  Must be called right after qp_setup (w = g, empty active set).
// ROLLED (the closed-loop kernel): the loop that decodes the guessed rows is not unrolled — one copy in the instruction
// stream, that kernel is bound by instruction fetch; the open-loop kernels keep the unrolled form (the race kernel is
// bound by the latency of one warp's dependent chain: 60 -> 79 us at B = 4096 with the rolled loop) — and a guessed row
// that depends on the rows before it is taken out individually (see the rounds below); the open-loop guess, velocity
// rows only, cannot be dependent and keeps the plain test.
template <int N, int MO, int WS, bool ROLLED = false>
LDCBF_HD void qp_warm_start(const StepConst& C, const int (&codes)[2 * N], double* ws, QpState<N, MO>& s) {
    constexpr int NV = 2 * N;
    if (s.done) return;
    // signed deviations m = v - mid of all two-sided rows and slacks of the LDCBF rows at w = g (same formulas as
    // the scan), to turn a row code into its right-hand side: b - n.g = -(slack of that side at g)
    double mdev[4 * N], half[4 * N], cbf[N * MO];
    {
        double Vx = s.v0x, Vy = s.v0y;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double dx = s.px[k + 1] - s.px[k], dy = s.py[k + 1] - s.py[k];
            const double off = (double)s.ft[k] * C.foot_offset;
            mdev[2 * k] = (s.rc[k] * dx + s.rs[k] * dy) - C.legx_mid;          half[2 * k] = C.legx_half;
            mdev[2 * k + 1] = (s.rc[k] * dy - s.rs[k] * dx) - (C.legy_mid - off); half[2 * k + 1] = C.legy_half;
            Vx = C.gtil * dx - Vx; Vy = C.gtil * dy - Vy;
            const int kk = k + 1;
            mdev[2 * N + 2 * k] = (s.rc[kk] * Vx + s.rs[kk] * Vy) - s.vmid[kk]; half[2 * N + 2 * k] = s.vhalf[kk];
            mdev[2 * N + 2 * k + 1] = ((double)s.ft[kk] * s.rc[kk] * Vy - s.rs[kk] * Vx) - s.vlat_mid;
            half[2 * N + 2 * k + 1] = s.vlat_half;
#pragma unroll
            for (int o = 0; o < MO; ++o) cbf[k * MO + o] = s.ex[o] * s.px[kk] + s.ey[o] * s.py[kk] - s.hb[o];
        }
    }
    // Slot j is addressed dynamically in the workspace, its right-hand side is parked in the first row of the Gram
    // block, which is only built in the rounds below.
    unsigned mask = 0;
    auto load_row = [&](int j) {
        int code = -1;
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) if (jj == j) code = codes[jj];
        const int id = code >> 1;
        bool ok = code >= 0 && id < 4 * N + N * MO;
        double sl = 0.0;
        const double sg = (code & 1) ? -1.0 : 1.0;
        if (ok) {
            if (id < 4 * N) {
                double m = 0.0, h = 0.0;
#pragma unroll
                for (int i = 0; i < 4 * N; ++i) if (i == id) { m = mdev[i]; h = half[i]; }
                sl = h + sg * m;                      // lower side: v - lo = half + m ; upper side: hi - v = half - m
            } else {
                const int q = id - 4 * N;
                ok = (q % MO) < s.nb && !(code & 1);
#pragma unroll
                for (int i = 0; i < N * MO; ++i) if (i == q) sl = cbf[i];
            }
        }
        GM(0, j) = 0.0;
        if (ok) {
            double a[NV];
            row_normal<N, MO>(id, sg, C.gtil, s.rc, s.rs, s.ft, s.ex, s.ey, s.ces, s.ns, a);
#pragma unroll
            for (int i = 0; i < NV; ++i) AN(j, i) = a[i];
            RC(j) = (double)code;
            GM(0, j) = -sl;
            mask |= 1u << j;
        }
    };
    if (ROLLED) {
#pragma unroll 1
        for (int j = 0; j < NV; ++j) load_row(j);
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) load_row(j);
    }
    double rhs[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) rhs[j] = GM(0, j);
    if (mask == 0u) return;
    bool fail = false;
    double uu[NV];
    // Gram matrix of the loaded slots (identity on the empty ones), once: a round that drops rows only has to put the
    // identity back on their rows and columns, the entries between the rows that stay do not change.  Only shared
    // memory and the mask are involved, so the closed-loop kernel (ROLLED) runs it as a real loop: 40 instructions of
    // code instead of ~500 in a kernel that waits for its own code.
    if (ROLLED) {
#pragma unroll 1
        for (int j = 0; j < NV; ++j) {
            const bool aj = (mask >> j) & 1u;
#pragma unroll 1
            for (int l = 0; l <= j; ++l) {
                double g2 = 0.0;
#pragma unroll
                for (int i = 0; i < NV; ++i) g2 += AN(j, i) * AN(l, i);
                const bool al = (mask >> l) & 1u;
                g2 = (aj && al) ? g2 : (j == l ? 1.0 : 0.0);
                GM(j, l) = g2; GM(l, j) = g2;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const bool aj = (mask >> j) & 1u;
#pragma unroll
            for (int l = 0; l <= j; ++l) {
                double g2 = 0.0;
#pragma unroll
                for (int i = 0; i < NV; ++i) g2 += AN(j, i) * AN(l, i);
                const bool al = (mask >> l) & 1u;
                g2 = (aj && al) ? g2 : (j == l ? 1.0 : 0.0);
                GM(j, l) = g2; GM(l, j) = g2;
            }
        }
    }
#pragma unroll 1
    for (int round = 0; round <= NV; ++round) {
        ++s.iters;      // a round costs about one trip (Cholesky, two solves) and is counted as one
        // Cholesky, u = G^-1 rhs
        double L[NV][NV];
        // A guessed row whose pivot falls below 1e-10 of its diagonal entry (compared without the division) depends on
        // the rows before it: it is taken out on the spot — its row and column of the factor become the identity, which
        // is exactly the factor of the Gram matrix without that row — instead of giving up the whole guess (5 % of the
        // closed-loop steps had such a row, each then cost a cold start of ~25 trips: the p99 of the trips per step).
        unsigned dep = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double dj = GM(j, j);
            const double diag = dj;
#pragma unroll
            for (int l = 0; l < j; ++l) dj -= L[j][l] * L[j][l];
            const bool bad = ((mask >> j) & 1u) && !(dj > 1e-10 * diag);
            dep |= bad ? (1u << j) : 0u;
            if (!ROLLED && bad) fail = true;        // open-loop guesses (velocity rows only) are independent by construction
            const double inv = (ROLLED && bad) ? 1.0 : rsqrt_f64(fmax(dj, 1e-300));
            L[j][j] = inv;
#pragma unroll
            for (int i = j + 1; i < NV; ++i) {
                double v = GM(i, j);
#pragma unroll
                for (int l = 0; l < j; ++l) v -= L[i][l] * L[j][l];
                L[i][j] = (ROLLED && bad) ? 0.0 : v * inv;
            }
            if (ROLLED) {
#pragma unroll
                for (int l = 0; l < j; ++l) L[j][l] = bad ? 0.0 : L[j][l];
            }
        }
        if (!ROLLED && fail) break;                 // dependent guess: back to the cold start
        if (!ROLLED) dep = 0;
        mask &= ~dep;
#if !defined(__CUDA_ARCH__) && defined(LDCBF_HOST_COUNTERS)
        if (dep) ++ldcbf_host_dependent_guesses;
#endif
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double v = ((mask >> j) & 1u) ? rhs[j] : 0.0;
#pragma unroll
            for (int l = 0; l < j; ++l) v -= L[j][l] * uu[l];
            uu[j] = v * L[j][j];
        }
#pragma unroll
        for (int j = NV - 1; j >= 0; --j) {
            double v = uu[j];
#pragma unroll
            for (int l = j + 1; l < NV; ++l) v -= L[l][j] * uu[l];
            uu[j] = v * L[j][j];
        }
        unsigned neg = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) if (((mask >> j) & 1u) && !(uu[j] >= 0.0)) neg |= 1u << j;
        if (round == NV && neg != 0u) { fail = true; break; }
        mask &= ~neg;                                           // drop the rows with negative multipliers
        const unsigned gone = neg | dep;                        // ... and clear them and the dependent ones from the workspace
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if ((gone >> j) & 1u) {
#pragma unroll
                for (int i = 0; i < NV; ++i) { AN(j, i) = 0.0; GM(j, i) = 0.0; GM(i, j) = 0.0; }
                GM(j, j) = 1.0;
                RC(j) = -1.0;
            }
        }
        if (neg == 0u || mask == 0u) break;
    }
    if (fail || mask == 0u) {                                   // back to the cold start
#pragma unroll
        for (int j = 0; j < NV; ++j) {
#pragma unroll
            for (int i = 0; i < NV; ++i) { AN(j, i) = 0.0; GM(j, i) = (i == j) ? 1.0 : 0.0; }
            RC(j) = -1.0;
        }
        return;
    }
    // accept: multipliers, active mask, w = g + N u  (GM already holds the Gram matrix of the accepted slots)
    s.amask = mask;
#pragma unroll
    for (int j = 0; j < NV; ++j) s.u[j] = ((mask >> j) & 1u) ? uu[j] : 0.0;
#pragma unroll
    for (int k = 1; k <= N; ++k) {
        double ax = s.gx, ay = s.gy;
#pragma unroll
        for (int j = 0; j < NV; ++j) { ax += s.u[j] * AN(j, 2 * (k - 1)); ay += s.u[j] * AN(j, 2 * (k - 1) + 1); }
        s.px[k] = ax; s.py[k] = ay;
    }
}

// Row codes of the final active set, shifted by one stage for the next MPC step (stage k of this step is stage k-1 of
// the next one; rows of the first stage disappear).  out[j] = -1 where nothing carries over.
// dup_last: the rows that were active at the LAST stage are also guessed active at the new last stage (which otherwise
// starts without rows), as far as free entries remain: a walking gait is close to periodic over one stage.
template <int N, int MO, int WS>
LDCBF_HD void shift_codes(const QpState<N, MO>& s, const double* ws, int (&out)[2 * N], bool dup_last = false) {
    constexpr int NV = 2 * N;
    int extra[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        int code = ((s.amask >> j) & 1u) ? (int)RC(j) : -1;
        extra[j] = -1;
        if (dup_last && code >= 0) {
            const int id = code >> 1;
            const bool last = id < 2 * N ? id >= 2 * (N - 1)
                            : id < 4 * N ? id - 2 * N >= 2 * (N - 1)
                            : id < 4 * N + N * MO ? id - 4 * N >= (N - 1) * MO : false;
            if (last) extra[j] = code;
        }
        if (code >= 0) {
            const int id = code >> 1, side = code & 1;
            int nid = -1;
            if (id < 2 * N) nid = (id >= 2) ? id - 2 : -1;                           // leg(k) -> leg(k-1)
            else if (id < 4 * N) nid = (id - 2 * N >= 2) ? id - 2 : -1;              // vel(k) -> vel(k-1)
            else if (id < 4 * N + N * MO) nid = (id - 4 * N >= MO) ? id - MO : -1;   // cbf(k,o) -> cbf(k-1,o)
            code = nid >= 0 ? 2 * nid + side : -1;
        }
        out[j] = code;
    }
    if (dup_last) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            bool pending = extra[j] >= 0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (pending && out[i] < 0) { out[i] = extra[j]; pending = false; }
            }
        }
    }
}

#undef AN
#undef GM
#undef RC

// Outputs: states, footsteps, objective (NaN when not solved).
template <int N, int MO>
LDCBF_HD void qp_finish(const StepConst& C, const QpState<N, MO>& s, QpSolution<N>& S) {
    S.status = s.status;
    S.iters = s.iters;
    const double nan = quiet_nan();
    const double gx = s.gx, gy = s.gy;
#pragma unroll
    for (int k = 0; k <= N; ++k) S.th[k] = s.th[k];
#pragma unroll
    for (int k = 0; k < N; ++k) S.om[k] = s.om[k];
    S.px[0] = s.p0x; S.py[0] = s.p0y; S.vx[0] = s.v0x; S.vy[0] = s.v0y;
    double obj = (s.p0x - gx) * (s.p0x - gx) + (s.p0y - gy) * (s.p0y - gy);
    const bool ok = s.status == LDCBF_STATUS_SOLVED;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double dx = s.px[k + 1] - s.px[k], dy = s.py[k + 1] - s.py[k];
        S.vx[k + 1] = C.gtil * dx - S.vx[k];
        S.vy[k + 1] = C.gtil * dy - S.vy[k];
        S.ux[k] = (s.px[k + 1] - C.ch * s.px[k] - C.sh_over_beta * S.vx[k]) * C.inv_one_m_ch;
        S.uy[k] = (s.py[k + 1] - C.ch * s.py[k] - C.sh_over_beta * S.vy[k]) * C.inv_one_m_ch;
        S.px[k + 1] = s.px[k + 1]; S.py[k + 1] = s.py[k + 1];
        obj += (s.px[k + 1] - gx) * (s.px[k + 1] - gx) + (s.py[k + 1] - gy) * (s.py[k + 1] - gy);
    }
    S.obj = ok ? obj : nan;
    if (!ok) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            S.ux[k] = nan; S.uy[k] = nan;
            S.px[k + 1] = nan; S.py[k + 1] = nan; S.vx[k + 1] = nan; S.vy[k + 1] = nan;
        }
    }
}

// One scenario from start to end (one thread).
template <int N, int MO, int WS>
LDCBF_HD void solve_scenario(const StepConst& C, double p0x, double v0x, double p0y, double v0y, double th0, double gx,
                             double gy, const int (&ft)[N + 1], const double4 (&ce)[MO], int nb, const double4* ce_stream,
                             int n_stream, double delta, const Limits& lim, double* ws, QpSolution<N>& S) {
    QpState<N, MO> s;
    qp_setup<N, MO, WS>(C, p0x, v0x, p0y, v0y, th0, gx, gy, ft, ce, nb, ce_stream, n_stream, delta, lim, ws, s);
    if (!C.cold_start) {
        int codes[2 * N];
        guess_codes<N, MO>(s, codes);
        qp_warm_start<N, MO, WS>(C, codes, ws, s);
    }
    while (!s.done) qp_trip<N, MO, WS>(C, ws, s);
    qp_finish<N, MO>(C, s, S);
}

}  // namespace ldcbf
