// K2+K3 device code: heading schedule, QP assembly in CoM-position space and the exact dual active-set
// solve for ONE scenario, executed by ONE thread with all problem data in registers / thread-local memory.
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
// method; with Hessian I the projector onto the active normals is a <= 2N Gram/Cholesky solve rebuilt
// from scratch every iteration (<= 6x6 here), so there is no factor-update bookkeeping.
// The maneuverability row k and the longitudinal walking-velocity row k+1 are the same linear form
// (cos theta_{k+1}, sin theta_{k+1}) . v_{k+1}; they are merged into one row with the tighter upper bound.
#pragma once
#include "ldcbf_common.cuh"

namespace ldcbf {

template <int N>
struct RowScale {
    // 1/sqrt(4k-3): norm of the velocity-row pattern (1, -2, 2, ..., +-2) over p_k, p_{k-1}, .., p_1
    __device__ __forceinline__ static double vel(int k) { return rsqrt((double)(4 * k - 3)); }
    __device__ __forceinline__ static double leg(int k) { return k == 0 ? 1.0 : 0.70710678118654752440; }
};

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

// Dense signed normal (length 2N) of candidate row `id`, with sign sg (+1: a.w >= lo, -1: -a.w >= -hi).
template <int N, int MO>
__device__ __forceinline__ void row_normal(int id, double sg, const double* rc, const double* rs, const int* ft,
                                           const double* ex, const double* ey, double* a) {
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) a[i] = 0.0;
    if (id < 2 * N) {                       // leg reachability (k, sub): r.(p_{k+1} - p_k)
        const int k = id >> 1, sub = id & 1;
        const double s = sg * RowScale<N>::leg(k);
        const double rx = (sub ? -rs[k] : rc[k]) * s, ry = (sub ? rc[k] : rs[k]) * s;
        a[2 * k] = rx; a[2 * k + 1] = ry;
        if (k > 0) { a[2 * k - 2] = -rx; a[2 * k - 1] = -ry; }
    } else if (id < 4 * N) {                // velocity rows at state k = 1..N: r.v_k / (gtil sqrt(4k-3))
        const int j = id - 2 * N, k = (j >> 1) + 1, sub = j & 1;
        const double s = sg * RowScale<N>::vel(k);
        const double rx = (sub ? -rs[k] : rc[k]) * s, ry = (sub ? (double)ft[k] * rc[k] : rs[k]) * s;
        a[2 * (k - 1)] = rx; a[2 * (k - 1) + 1] = ry;
        double alt = -2.0;
        for (int i = k - 1; i >= 1; --i) { a[2 * (i - 1)] = alt * rx; a[2 * (i - 1) + 1] = alt * ry; alt = -alt; }
    } else {                                // LDCBF (k, o): eta_o . p_k
        const int j = id - 4 * N, k = j / MO + 1, o = j - (k - 1) * MO;
        a[2 * (k - 1)] = sg * ex[o]; a[2 * (k - 1) + 1] = sg * ey[o];
    }
}

// One scenario.  c_eta points at this scenario's [max_obs][4] block (c_x, c_y, eta_x, eta_y).
template <int N, int MO>
__device__ void solve_scenario(const StepConst& C, double p0x, double v0x, double p0y, double v0y, double th0,
                               double gx, double gy, const int* ft /*[N+1]*/, const double4* c_eta, int nb,
                               double delta, double alpha_over_pi, double vmax0, double omega_max,
                               double omega_min, QpSolution<N>& S) {
    constexpr int NV = 2 * N;
    double rc[N + 1], rs[N + 1];
    // ---- heading schedule (HumanoidMpc.py:137-160)
    {
        const double phi = atan2(gy - p0y, gx - p0x);
        double thk = th0;
        S.th[0] = thk;
        sincos(thk, &rs[0], &rc[0]);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double w = fmin(fmax(phi - thk, omega_min), omega_max);
            S.om[k] = w;
            thk = __dadd_rn(thk, __dmul_rn(w, C.sampling_time));
            S.th[k + 1] = thk;
            sincos(thk, &rs[k + 1], &rc[k + 1]);
        }
    }
    // ---- half-planes: eta_o . p >= eta_o . c_o + delta
    double ex[MO], ey[MO], hb[MO];
    int status = LDCBF_STATUS_SOLVED;
#pragma unroll
    for (int o = 0; o < MO; ++o) {
        ex[o] = 0.0; ey[o] = 0.0; hb[o] = -1.0;
        if (o < nb) {
            const double4 ce = c_eta[o];
            ex[o] = ce.z; ey[o] = ce.w;
            hb[o] = ce.z * ce.x + ce.w * ce.y + delta;
            if (!(ce.z == ce.z) || !(ce.w == ce.w)) status = LDCBF_STATUS_DEGENERATE;
            // constant k = 0 row (HumanoidMpc.py:284-292 with k = 0)
            else if (ce.z * p0x + ce.w * p0y - hb[o] < -C.eps_const_row) status = LDCBF_STATUS_INFEASIBLE;
        }
    }
    // ---- row bounds (normalised)
    double vhi[N + 1];
#pragma unroll
    for (int k = 1; k <= N; ++k) vhi[k] = fmin(vmax0, vmax0 - alpha_over_pi * fabs(S.om[k - 1]));

    // ---- Goldfarb-Idnani dual active set on  min ||w - g||^2  s.t. rows
    double px[N + 1], py[N + 1];
    px[0] = p0x; py[0] = p0y;
#pragma unroll
    for (int k = 1; k <= N; ++k) { px[k] = gx; py[k] = gy; }   // unconstrained optimum

    int na = 0;
    int aid[NV];
    double asg[NV], u[NV];
    double An[NV][NV];      // signed normals of the active rows
    int iters = 0;
    const double inv_gtil = 1.0 / C.gtil;

    while (status == LDCBF_STATUS_SOLVED) {
        // -- most violated row at the current point
        double best = -C.eps_active, bsg = 0.0;
        int bid = -1;
        {
            double Vx = v0x, Vy = v0y;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double dx = px[k + 1] - px[k], dy = py[k + 1] - py[k];
                const double il = RowScale<N>::leg(k);
                const double lg = (rc[k] * dx + rs[k] * dy) * il;
                const double lt = (rc[k] * dy - rs[k] * dx) * il;
                const double off = (double)ft[k] * C.foot_offset;
                double s;
                s = lg - C.l_min_x * il;          if (s < best) { best = s; bid = 2 * k; bsg = 1.0; }
                s = C.l_max_x * il - lg;          if (s < best) { best = s; bid = 2 * k; bsg = -1.0; }
                s = lt - (C.l_min_y - off) * il;  if (s < best) { best = s; bid = 2 * k + 1; bsg = 1.0; }
                s = (C.l_max_y - off) * il - lt;  if (s < best) { best = s; bid = 2 * k + 1; bsg = -1.0; }
                Vx = C.gtil * dx - Vx; Vy = C.gtil * dy - Vy;       // v_{k+1}
                const int kk = k + 1;
                const double iv = RowScale<N>::vel(kk) * inv_gtil;
                const double vl = (rc[kk] * Vx + rs[kk] * Vy) * iv;
                const double vt = ((double)ft[kk] * rc[kk] * Vy - rs[kk] * Vx) * iv;
                s = vl - C.v_min0 * iv;           if (s < best) { best = s; bid = 2 * N + 2 * k; bsg = 1.0; }
                s = vhi[kk] * iv - vl;            if (s < best) { best = s; bid = 2 * N + 2 * k; bsg = -1.0; }
                s = vt - C.v_min1 * iv;           if (s < best) { best = s; bid = 2 * N + 2 * k + 1; bsg = 1.0; }
                s = C.v_max1 * iv - vt;           if (s < best) { best = s; bid = 2 * N + 2 * k + 1; bsg = -1.0; }
#pragma unroll
                for (int o = 0; o < MO; ++o) {
                    if (o < nb) {
                        s = ex[o] * px[kk] + ey[o] * py[kk] - hb[o];
                        if (s < best) { best = s; bid = 4 * N + k * MO + o; bsg = 1.0; }
                    }
                }
            }
        }
        if (bid < 0) break;   // primal feasible: optimal

        double np[NV];
        row_normal<N, MO>(bid, bsg, rc, rs, ft, ex, ey, np);
        double s_p = best, u_p = 0.0;
        // -- add row p: partial steps until it can enter the active set
        for (;;) {
            if (++iters > C.max_iter) { status = LDCBF_STATUS_MAX_ITER; break; }
            // r = (N^T N)^-1 N^T n+ ,  z = n+ - N r
            double r[NV], z[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) z[i] = np[i];
            if (na > 0) {
                double G[NV][NV], d[NV];
                for (int j = 0; j < na; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NV; ++i) acc += An[j][i] * np[i];
                    d[j] = acc;
                    for (int l = 0; l <= j; ++l) {
                        double g = 0.0;
#pragma unroll
                        for (int i = 0; i < NV; ++i) g += An[j][i] * An[l][i];
                        G[j][l] = g;
                    }
                }
                // Cholesky G = L L^T in place (lower), then two triangular solves
                for (int j = 0; j < na; ++j) {
                    double dj = G[j][j];
                    for (int l = 0; l < j; ++l) dj -= G[j][l] * G[j][l];
                    dj = sqrt(fmax(dj, 1e-300));
                    G[j][j] = dj;
                    const double inv = 1.0 / dj;
                    for (int i = j + 1; i < na; ++i) {
                        double v = G[i][j];
                        for (int l = 0; l < j; ++l) v -= G[i][l] * G[j][l];
                        G[i][j] = v * inv;
                    }
                }
                for (int j = 0; j < na; ++j) {
                    double v = d[j];
                    for (int l = 0; l < j; ++l) v -= G[j][l] * r[l];
                    r[j] = v / G[j][j];
                }
                for (int j = na - 1; j >= 0; --j) {
                    double v = r[j];
                    for (int l = j + 1; l < na; ++l) v -= G[l][j] * r[l];
                    r[j] = v / G[j][j];
                }
                for (int j = 0; j < na; ++j) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) z[i] -= r[j] * An[j][i];
                }
            }
            double zz = 0.0;
#pragma unroll
            for (int i = 0; i < NV; ++i) zz += z[i] * z[i];
            const bool dependent = !(zz > 1e-14) || na >= NV;
            // dual step length: largest t keeping the active multipliers non-negative
            double t1 = INFINITY;
            int ldrop = -1;
            for (int j = 0; j < na; ++j) {
                if (r[j] > 1e-14) {
                    const double tj = u[j] / r[j];
                    if (tj < t1) { t1 = tj; ldrop = j; }
                }
            }
            const double t2 = dependent ? INFINITY : -s_p / zz;
            const double t = fmin(t1, t2);
            if (!(t < INFINITY)) { status = LDCBF_STATUS_INFEASIBLE; break; }
            for (int j = 0; j < na; ++j) u[j] -= t * r[j];
            u_p += t;
            if (!dependent) {
#pragma unroll
                for (int k = 1; k <= N; ++k) { px[k] += t * z[2 * (k - 1)]; py[k] += t * z[2 * (k - 1) + 1]; }
                s_p += t * zz;
            }
            if (t2 <= t1) {          // full step: row p becomes active
                aid[na] = bid; asg[na] = bsg; u[na] = u_p;
#pragma unroll
                for (int i = 0; i < NV; ++i) An[na][i] = np[i];
                ++na;
                break;
            }
            // partial step: drop row ldrop and try again
            for (int j = ldrop; j < na - 1; ++j) {
                aid[j] = aid[j + 1]; asg[j] = asg[j + 1]; u[j] = u[j + 1];
#pragma unroll
                for (int i = 0; i < NV; ++i) An[j][i] = An[j + 1][i];
            }
            --na;
        }
    }
    (void)aid; (void)asg;

    // ---- outputs: states, footsteps, objective
    S.status = status;
    S.iters = iters;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    S.px[0] = p0x; S.py[0] = p0y; S.vx[0] = v0x; S.vy[0] = v0y;
    double obj = (p0x - gx) * (p0x - gx) + (p0y - gy) * (p0y - gy);
    const bool ok = status == LDCBF_STATUS_SOLVED;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double dx = px[k + 1] - px[k], dy = py[k + 1] - py[k];
        S.vx[k + 1] = C.gtil * dx - S.vx[k];
        S.vy[k + 1] = C.gtil * dy - S.vy[k];
        S.ux[k] = (px[k + 1] - C.ch * px[k] - C.sh_over_beta * S.vx[k]) * C.inv_one_m_ch;
        S.uy[k] = (py[k + 1] - C.ch * py[k] - C.sh_over_beta * S.vy[k]) * C.inv_one_m_ch;
        S.px[k + 1] = px[k + 1]; S.py[k + 1] = py[k + 1];
        obj += (px[k + 1] - gx) * (px[k + 1] - gx) + (py[k + 1] - gy) * (py[k + 1] - gy);
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

}  // namespace ldcbf
