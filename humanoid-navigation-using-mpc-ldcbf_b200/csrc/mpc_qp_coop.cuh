// K2+K3, cooperative variant for batches that do not fill the GPU: G lanes of a warp solve ONE scenario.
//
// Same reference semantics and the same QP as mpc_qp.cuh (heading schedule HumanoidMpc.py:137-160, rows :162-249 and
// :252-294 with HumanoidMPCCustomLCBF.py:30-31, cost :321-333, solve :417, integration :335-343,441-447), the same
// CoM-position formulation and the same dual active-set method (Goldfarb-Idnani, Hessian I).  What
// changes is the mapping and the factorisation:
//   * the thread-per-scenario kernel is bound, at the benchmark batch (4096 scenarios on 592 SM sub-partitions), by the
//     latency of ONE thread's instruction stream: ~800 dependent-ish instructions per trip of the active-set loop,
//     a third of them the scan for the most violated row, another third a 2N x 2N Cholesky from scratch (a chain of
//     2N reciprocal square roots);
//   * here the rows of the QP are built once, as dense normals, into shared memory by the G lanes (row i by lane
//     i mod G, the four sincos of the heading schedule by four different lanes); the scan is one row or a few rows per
//     lane and a (value|side|row) 64-bit min over the group by xor shuffles; the chosen row is read back by every lane
//     as a shared-memory broadcast;
//   * the active set is kept as an orthogonal factorisation N_A = J [R; 0] (J 2N x 2N, R upper triangular), replicated
//     in the registers of every lane of the group (so no further communication is needed): a row enters with ONE
//     Householder reflection of the free columns of J (one rsqrt), leaves with a Givens chain on R; the step
//     directions are z = J2 d2 and r = R^-1 d1 with d = J^T n+ — no Gram matrix, no squared condition number, one
//     triangular solve per trip.  Positions j >= q of R hold an identity padding so every loop is fully unrolled with
//     static register indices and group-uniform predicates on (q, l).
// The source compiles for the host with G = 1 (tests/cpu_harness), where the group collectives are the identity.
#pragma once
#include "mpc_qp.cuh"

namespace ldcbf {

template <int N, int MO>
struct CoopShape {
    static constexpr int NV = 2 * N;              // unknowns w = (p_1..p_N)
    static constexpr int NR = 4 * N + N * MO;     // folded rows: leg 2N, velocity 2N (both two-sided), LDCBF N*MO
    static constexpr int RS = NV + 2;             // per row: NV coefficients, c0, half
    static constexpr int RAW = RS * NR + 3 * (N + 1);
    // doubles of shared memory per scenario, = 8 mod 16: two groups of one half-warp then sit in different banks
    static constexpr int DOUBLES = ((RAW + 7) / 16) * 16 + 8;
    static_assert(NR <= 64, "row index must fit the 6 low bits of the scan key");
};

// The G lanes that share a scenario.  Collectives are group-uniform: every lane of `mask` calls them together.
template <int G>
struct LaneGroup {
    unsigned mask;
    int lane;
    LDCBF_HD long long min_ll(long long v) const {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const long long t = __shfl_xor_sync(mask, v, o);
            v = t < v ? t : v;
        }
#endif
        return v;
    }
    LDCBF_HD void sync() const {
#ifdef __CUDA_ARCH__
        if (G > 1) __syncwarp(mask);
#endif
    }
};

// order-preserving map double -> signed 64-bit integer (a < b  <=>  key(a) < key(b), no NaNs)
LDCBF_HD long long order_key(double x) {
#ifdef __CUDA_ARCH__
    const long long b = __double_as_longlong(x);
#else
    long long b;
    memcpy(&b, &x, 8);
#endif
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

template <int N>
struct CoopState {
    static constexpr int NV = 2 * N;
    double J[NV][NV];        // orthogonal; columns < q span the active normals
    double R[NV][NV];        // upper triangular q x q block (entries [i][j], i <= j + 1 used), identity padding beyond q
    double Rinv[NV];         // 1 / R[j][j]
    double u[NV];            // multipliers in factor order
    int code[NV];            // row code 2*id + (upper side) in factor order, -1 beyond q
    double w[NV];            // current iterate (p_1..p_N), interleaved x, y
    double np[NV], nn, s_p, u_p;   // row being added: signed normal, |n|^2, slack, multiplier
    int p_code;
    int q;
    int status, iters;
    bool need_scan, done;
    double tol;              // a row counts as violated below -tol (eps_active; eps_infeasible after a relaxed restart)
    double p0x, p0y, v0x, v0y, gx, gy;
    double th[N + 1], om[N];
#ifdef LDCBF_COOP_PROFILE
    long long prof[8];
#endif
};

#define CA(c, i) sm[(c) * NR + (i)]

// development aid (make profile-lib): cycles per phase of the trip, accumulated by the first lane of the first block
#if defined(LDCBF_COOP_PROFILE) && defined(__CUDA_ARCH__)
#define COOP_T(k) do { const long long now_ = clock64(); s.prof[k] += now_ - prof_t_; prof_t_ = now_; } while (0)
#define COOP_T0() long long prof_t_ = clock64()
#else
#define COOP_T(k) do { } while (0)
#define COOP_T0() do { } while (0)
#endif

// Heading schedule, rows into shared memory, empty active set at the unconstrained optimum.  `ce` = (c, eta) of the
// scenario's obstacles in global memory (nobs <= MO of them).
template <int N, int MO, int G>
LDCBF_HD void coop_setup(const StepConst& C, const LaneGroup<G>& grp, double p0x, double v0x, double p0y, double v0y,
                         double th0, double gx, double gy, const int (&ft)[N + 1], const double4* ce, int nobs,
                         double delta, const Limits& lim, double* sm, CoopState<N>& s) {
    using SH = CoopShape<N, MO>;
    constexpr int NV = SH::NV, NR = SH::NR;
    double* sm_rc = sm + SH::RS * NR;
    double* sm_rs = sm_rc + (N + 1);
    double* sm_ft = sm_rs + (N + 1);
    s.p0x = p0x; s.p0y = p0y; s.v0x = v0x; s.v0y = v0y; s.gx = gx; s.gy = gy;
    // ---- heading schedule (HumanoidMpc.py:137-160): the cheap recurrence on every lane, one sincos per lane
    {
        const double phi = atan2(gy - p0y, gx - p0x);
        double thk = th0;
        s.th[0] = thk;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double om = fmin(fmax(phi - thk, lim.omega_min), lim.omega_max);
            s.om[k] = om;
            thk = add_rn(thk, mul_rn(om, C.sampling_time));
            s.th[k + 1] = thk;
        }
        for (int k = grp.lane; k <= N; k += G) {
            double t = 0.0, f = 1.0;
#pragma unroll
            for (int j = 0; j <= N; ++j) if (j == k) { t = s.th[j]; f = (double)ft[j]; }
            double sn, cs;
            sincos(t, &sn, &cs);
            sm_rc[k] = cs; sm_rs[k] = sn; sm_ft[k] = f;
        }
    }
    // ---- constant k = 0 LDCBF row and degenerate half-planes (HumanoidMpc.py:284-292 with k = 0)
    int status = LDCBF_STATUS_SOLVED;
    for (int o = 0; o < nobs; ++o) {
        const double4 c4 = ce[o];
        if (!(c4.z == c4.z) || !(c4.w == c4.w)) status = LDCBF_STATUS_DEGENERATE;
        else if (c4.z * p0x + c4.w * p0y - (c4.z * c4.x + c4.w * c4.y + delta) < -C.eps_const_row &&
                 status == LDCBF_STATUS_SOLVED)
            status = LDCBF_STATUS_INFEASIBLE;
    }
    grp.sync();
    // ---- rows: lane builds rows lane, lane + G, ...   slack_i(w) = half_i - |a_i.w + c0_i| (two-sided), a_i.w + c0_i (LDCBF)
    const double vlat_mid = 0.5 * (lim.vmax1 + C.v_min1), vlat_half = 0.5 * (lim.vmax1 - C.v_min1);
    for (int i = grp.lane; i < NR; i += G) {
        int typ, k, sub;    // k = index of the last state the row involves
        if (i < 2 * N) { typ = 0; k = (i >> 1) + 1; sub = i & 1; }
        else if (i < 4 * N) { typ = 1; k = ((i - 2 * N) >> 1) + 1; sub = i & 1; }
        else { typ = 2; k = (i - 4 * N) / MO + 1; sub = (i - 4 * N) - (k - 1) * MO; }
        const int kth = (typ == 0) ? k - 1 : k;
        const double c = sm_rc[kth], sn = sm_rs[kth], f = sm_ft[kth];
        double rx, ry, c0, half = 0.0;
        if (typ == 0) {
            // leg reachability k' = k-1 (HumanoidMpc.py:183-202): R(theta_k')^T (p_k - p_k') in the box, lateral box
            // shifted by -foot * 0.05
            rx = sub ? -sn : c; ry = sub ? c : sn;
            const double mid = sub ? C.legy_mid - f * C.foot_offset : C.legx_mid;
            half = sub ? C.legy_half : C.legx_half;
            c0 = -mid - (k == 1 ? rx * p0x + ry * p0y : 0.0);
        } else if (typ == 1) {
            // velocity rows at state k (HumanoidMpc.py:162-181 merged with :204-219):
            // v_k = (-1)^k v_0 + gtil (p_k - 2 p_{k-1} + 2 p_{k-2} - ... +- p_0)
            rx = sub ? -sn : c; ry = sub ? f * c : sn;
            double om_prev = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) if (j == k - 1) om_prev = s.om[j];
            const double vhi = fmin(lim.vmax0, lim.vmax0 - lim.alpha_over_pi * fabs(om_prev));
            const double mid = sub ? vlat_mid : 0.5 * (vhi + C.v_min0);
            half = sub ? vlat_half : 0.5 * (vhi - C.v_min0);
            const double sk = (k & 1) ? -1.0 : 1.0;
            c0 = sk * (rx * v0x + ry * v0y) + sk * C.gtil * (rx * p0x + ry * p0y) - mid;
        } else {
            // LDCBF row of obstacle `sub` at state k: eta.(p_k - c) - delta >= 0; absent obstacle: slack +inf
            rx = 0.0; ry = 0.0; c0 = INFINITY;
            if (sub < nobs) {
                const double4 c4 = ce[sub];
                rx = c4.z; ry = c4.w;
                c0 = -(c4.z * c4.x + c4.w * c4.y + delta);
            }
        }
#pragma unroll
        for (int j = 1; j <= N; ++j) {
            double kap = 0.0;
            if (j == k) kap = (typ == 1) ? C.gtil : 1.0;
            else if (j < k) {
                if (typ == 0) kap = (j == k - 1) ? -1.0 : 0.0;
                else if (typ == 1) kap = ((k - j) & 1) ? -2.0 * C.gtil : 2.0 * C.gtil;
            }
            CA(2 * (j - 1), i) = kap * rx;
            CA(2 * (j - 1) + 1, i) = kap * ry;
        }
        CA(NV, i) = c0;
        CA(NV + 1, i) = half;
    }
    // ---- empty active set at the unconstrained optimum w = (g, .., g)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
        for (int i = 0; i < NV; ++i) { s.J[i][j] = (i == j) ? 1.0 : 0.0; s.R[i][j] = (i == j) ? 1.0 : 0.0; }
        s.Rinv[j] = 1.0; s.u[j] = 0.0; s.code[j] = -1; s.np[j] = 0.0;
        s.w[j] = (j & 1) ? gy : gx;
    }
    s.nn = 1.0; s.s_p = 0.0; s.u_p = 0.0; s.p_code = 0; s.q = 0;
    s.iters = 0; s.need_scan = true;
    s.tol = C.eps_active;
    s.status = status;
    s.done = status != LDCBF_STATUS_SOLVED;
    grp.sync();
}

// One trip of the active-set loop (group-uniform control flow).
template <int N, int MO, int G>
LDCBF_HD void coop_trip(const StepConst& C, const LaneGroup<G>& grp, const double* sm, CoopState<N>& s) {
    using SH = CoopShape<N, MO>;
    constexpr int NV = SH::NV, NR = SH::NR;
    COOP_T0();
    if (s.need_scan) {
        // -- most violated row: every lane evaluates its rows, the group takes the min of (slack | side | row) keys
        long long best = 0x7fffffffffffffffLL;
        constexpr int RPL = (NR + G - 1) / G;
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) {
            const int i = grp.lane + ii * G;
            if (i < NR) {
                double m = CA(NV, i);
#pragma unroll
                for (int c = 0; c < NV; ++c) m += CA(c, i) * s.w[c];
                const bool two = i < 4 * N;
                const double sl = two ? CA(NV + 1, i) - fabs(m) : m;
                const long long side = (two && m >= 0.0) ? 64 : 0;
                const long long key = (order_key(sl) & ~127LL) | side | (long long)i;
                best = key < best ? key : best;
            }
        }
        best = grp.min_ll(best);
        const int id = (int)(best & 63);
        const bool upper = (best & 64) != 0;
        // the chosen row, read back by every lane (broadcast); its slack is recomputed exactly
        double a[NV], m = CA(NV, id);
#pragma unroll
        for (int c = 0; c < NV; ++c) { a[c] = CA(c, id); m += a[c] * s.w[c]; }
        const bool two = id < 4 * N;
        const double sl = two ? CA(NV + 1, id) - fabs(m) : m;
        if (!(sl < -s.tol)) { s.done = true; return; }     // primal feasible: optimal
        const double sg = upper ? -1.0 : 1.0;
        double nn = 0.0;
#pragma unroll
        for (int c = 0; c < NV; ++c) { s.np[c] = sg * a[c]; nn += a[c] * a[c]; }
        s.nn = nn; s.s_p = sl; s.u_p = 0.0;
        s.p_code = 2 * id + (upper ? 1 : 0);
        s.need_scan = false;
        COOP_T(0);
    }
    if (++s.iters > C.max_iter) { s.status = LDCBF_STATUS_MAX_ITER; s.done = true; return; }
    const int q = s.q;
    // one-hot masks of q and q-1: `(qbit >> c) & 1` instead of `c == q` keeps nvcc from turning the predicated
    // static-index moves below into dynamically indexed (local-memory) accesses
    const unsigned qbit = 1u << q, qm1bit = qbit >> 1;

    // d = J^T n+ ;  zz = |d2|^2 ;  r = R^-1 d1
    double d[NV], r[NV], zz = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) acc += s.J[i][j] * s.np[i];
        d[j] = acc;
        zz += (j >= q) ? acc * acc : 0.0;
    }
#pragma unroll
    for (int j = NV - 1; j >= 0; --j) {
        double v = (j < q) ? d[j] : 0.0;
#pragma unroll
        for (int l = j + 1; l < NV; ++l) v -= s.R[j][l] * r[l];
        r[j] = v * s.Rinv[j];
    }
    COOP_T(1);
    const bool dependent = !(zz > 1e-13 * s.nn) || q == NV;
    // dual step length t1 = min u_j / r_j over r_j > 0: tournament on (numerator, denominator) pairs compared by
    // cross-multiplication; (1, 0) stands for "none"
    double tn[NV], td[NV];
    int tj[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const bool ok = j < q && r[j] > 1e-13;
        tn[j] = ok ? s.u[j] : 1.0; td[j] = ok ? r[j] : 0.0; tj[j] = ok ? j : -1;
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
    const double t1n = tn[0], t1d = td[0];
    const int ldrop = tj[0];
    const bool full = !dependent && (ldrop < 0 || (-s.s_p) * t1d <= t1n * zz);
    if (!full && ldrop < 0) {
        // same rule as mpc_qp.cuh:qp_trip — a row that cannot be satisfied but is violated by no more than eps_infeasible
        // (a feasible set that is a point to rounding) restarts the solve once with that tolerance
        if (s.tol < C.eps_infeasible && -s.s_p <= C.eps_infeasible) {
            s.tol = C.eps_infeasible;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
#pragma unroll
                for (int i = 0; i < NV; ++i) { s.J[i][j] = (i == j) ? 1.0 : 0.0; s.R[i][j] = (i == j) ? 1.0 : 0.0; }
                s.Rinv[j] = 1.0; s.u[j] = 0.0; s.code[j] = -1; s.np[j] = 0.0;
                s.w[j] = (j & 1) ? s.gy : s.gx;
            }
            s.nn = 1.0; s.s_p = 0.0; s.u_p = 0.0; s.p_code = 0; s.q = 0;
            s.need_scan = true;
            return;
        }
        s.status = LDCBF_STATUS_INFEASIBLE; s.done = true; return;
    }
    COOP_T(2);
    const double t = full ? (-s.s_p) / zz : t1n / t1d;
#pragma unroll
    for (int j = 0; j < NV; ++j) s.u[j] -= t * r[j];      // r_j = 0 beyond q
    s.u_p += t;
    if (!dependent) {
        // primal step along z = J2 d2
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double zi = 0.0;
#pragma unroll
            for (int j = 0; j < NV; ++j) zi += (j >= q) ? s.J[i][j] * d[j] : 0.0;
            s.w[i] += t * zi;
        }
        s.s_p += t * zz;
    }
    COOP_T(3);
    if (full) {
        // -- row p enters at position q: Householder reflection H of the free columns so that H d2 = rho e_1
        const double rsz = rsqrt_f64(zz);
        const double nrm = zz * rsz;
        double alpha = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) alpha = ((qbit >> j) & 1u) ? d[j] : alpha;
        const double sgn = alpha < 0.0 ? -1.0 : 1.0;
        double v[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = (j > q) ? d[j] : (((qbit >> j) & 1u) ? alpha + sgn * nrm : 0.0);
        const double beta = rsz / (nrm + fabs(alpha));        // 2 / v.v = 1 / (nrm (nrm + |alpha|))
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < NV; ++j) acc += s.J[i][j] * v[j];
            acc *= beta;
#pragma unroll
            for (int j = 0; j < NV; ++j) s.J[i][j] -= acc * v[j];
        }
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const bool at = (qbit >> c) & 1u;
#pragma unroll
            for (int i = 0; i < c; ++i) s.R[i][c] = at ? d[i] : s.R[i][c];
            s.R[c][c] = at ? -sgn * nrm : s.R[c][c];
            s.Rinv[c] = at ? -sgn * rsz : s.Rinv[c];
            s.u[c] = at ? s.u_p : s.u[c];
            s.code[c] = at ? s.p_code : s.code[c];
        }
        s.q = q + 1;
        s.need_scan = true;
        COOP_T(4);
    } else {
        // -- position ldrop leaves: shift the columns to its right one place left (R becomes upper Hessenberg there)
        // and restore the triangle with Givens rotations of rows (c, c+1), applied to columns (c, c+1) of J as well
#pragma unroll
        for (int c = 0; c < NV - 1; ++c) {
            if (c >= ldrop && c + 1 < q) {
#pragma unroll
                for (int i = 0; i <= c + 1; ++i) s.R[i][c] = s.R[i][c + 1];
                s.u[c] = s.u[c + 1];
                s.code[c] = s.code[c + 1];
            }
        }
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const bool at = (qm1bit >> c) & 1u;                 // the vacated last column: zeros during the rotations
#pragma unroll
            for (int i = 0; i <= c; ++i) s.R[i][c] = at ? 0.0 : s.R[i][c];
            s.u[c] = at ? 0.0 : s.u[c];
            s.code[c] = at ? -1 : s.code[c];
        }
#pragma unroll
        for (int c = 0; c < NV - 1; ++c) {
            if (c >= ldrop && c + 1 < q) {
                const double a = s.R[c][c], b = s.R[c + 1][c];
                const double h2 = a * a + b * b;
                const double ih = rsqrt_f64(h2);
                const double cs = a * ih, sn = b * ih;
                s.R[c][c] = h2 * ih;
                s.Rinv[c] = ih;
                s.R[c + 1][c] = 0.0;
#pragma unroll
                for (int j = c + 1; j < NV; ++j) {
                    const double x = s.R[c][j], y = s.R[c + 1][j];
                    s.R[c][j] = cs * x + sn * y;
                    s.R[c + 1][j] = cs * y - sn * x;
                }
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const double x = s.J[i][c], y = s.J[i][c + 1];
                    s.J[i][c] = cs * x + sn * y;
                    s.J[i][c + 1] = cs * y - sn * x;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const bool at = (qm1bit >> c) & 1u;
            s.R[c][c] = at ? 1.0 : s.R[c][c];
            s.Rinv[c] = at ? 1.0 : s.Rinv[c];
        }
        s.q = q - 1;
        COOP_T(5);
    }
}

#undef CA

template <int N>
LDCBF_HD void coop_finish(const StepConst& C, const CoopState<N>& s, QpSolution<N>& S) {
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
        const double pxn = s.w[2 * k], pyn = s.w[2 * k + 1];
        const double dx = pxn - S.px[k], dy = pyn - S.py[k];
        S.vx[k + 1] = C.gtil * dx - S.vx[k];
        S.vy[k + 1] = C.gtil * dy - S.vy[k];
        S.ux[k] = (pxn - C.ch * S.px[k] - C.sh_over_beta * S.vx[k]) * C.inv_one_m_ch;
        S.uy[k] = (pyn - C.ch * S.py[k] - C.sh_over_beta * S.vy[k]) * C.inv_one_m_ch;
        S.px[k + 1] = pxn; S.py[k + 1] = pyn;
        obj += (pxn - gx) * (pxn - gx) + (pyn - gy) * (pyn - gy);
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

// One scenario from start to end (the G lanes of `grp`).
template <int N, int MO, int G>
LDCBF_HD void coop_solve_scenario(const StepConst& C, const LaneGroup<G>& grp, double p0x, double v0x, double p0y,
                                  double v0y, double th0, double gx, double gy, const int (&ft)[N + 1],
                                  const double4* ce, int nobs, double delta, const Limits& lim, double* sm,
                                  QpSolution<N>& S) {
    CoopState<N> s;
#if defined(LDCBF_COOP_PROFILE) && defined(__CUDA_ARCH__)
    for (int k = 0; k < 8; ++k) s.prof[k] = 0;
    const long long t_begin = clock64();
#endif
    coop_setup<N, MO, G>(C, grp, p0x, v0x, p0y, v0y, th0, gx, gy, ft, ce, nobs, delta, lim, sm, s);
#if defined(LDCBF_COOP_PROFILE) && defined(__CUDA_ARCH__)
    const long long t_setup = clock64();
#endif
    while (!s.done) coop_trip<N, MO, G>(C, grp, sm, s);
    coop_finish<N>(C, s, S);
#if defined(LDCBF_COOP_PROFILE) && defined(__CUDA_ARCH__)
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("coop profile (scenario 0): trips %d | setup %lld  scan %lld  d+solve %lld  ratio %lld  step %lld  add %lld  drop %lld | total %lld cycles\n",
               s.iters, t_setup - t_begin, s.prof[0], s.prof[1], s.prof[2], s.prof[3], s.prof[4], s.prof[5], clock64() - t_begin);
#endif
}

}  // namespace ldcbf
