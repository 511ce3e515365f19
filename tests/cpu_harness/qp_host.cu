// TEST INFRASTRUCTURE: compiles the solver source of the CUDA kernels (csrc/mpc_qp.cuh) for the HOST so that the
// CPU test suite can run the exact same algorithm code without a GPU.  Never linked into libldcbf_b200.so.
#define LDCBF_HOST_COUNTERS
#include "mpc_qp.cuh"

namespace ldcbf { void set_last_error(cudaError_t) {} }

// One scenario with a warm start from `codes_in[6]` (-1 = none); returns the shifted final active set in codes_out.
extern "C" int qp_host_solve_n3_warm(const ldcbf_params* prm, int max_obs, const double* x0, double theta0,
                                     const double* goal, const int8_t* foot, const double* c_eta, int nobs_b,
                                     double delta, const int* codes_in, int* codes_out, double* U, double* X,
                                     double* obj, int32_t* status, int32_t* iters) {
    using namespace ldcbf;
    constexpr int N = 3, MO = 4;
    const StepConst C = make_const(*prm);
    int ft[N + 1];
    for (int k = 0; k <= N; ++k) ft[k] = foot[k];
    double4 ce[MO];
    const int nb = nobs_b < MO ? nobs_b : MO;
    for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? make_double4(c_eta[4 * o], c_eta[4 * o + 1], c_eta[4 * o + 2], c_eta[4 * o + 3]) : make_double4(0, 0, 0, 0);
    double ws[QpWorkspace<N>::DOUBLES];
    QpState<N, MO> s;
    qp_setup<N, MO, 1>(C, x0[0], x0[1], x0[2], x0[3], theta0, goal[0], goal[1], ft, ce, nb, nullptr, 0, delta,
                       load_limits(C, nullptr, 0), ws, s);
    int codes[2 * N];
    for (int j = 0; j < 2 * N; ++j) codes[j] = codes_in[j];
    if (codes_in[0] == -2) guess_codes<N, MO>(s, codes);      // -2: the kernel's own initial guess
    qp_warm_start<N, MO, 1, true>(C, codes, ws, s);      // the closed-loop kernel's instantiation
    while (!s.done) qp_trip<N, MO, 1>(C, ws, s);
    QpSolution<N> S;
    qp_finish<N, MO>(C, s, S);
    shift_codes<N, MO, 1>(s, ws, codes, (prm->flags & 0x100) != 0);      // test-only bit: repeat the last stage's rows
    for (int j = 0; j < 2 * N; ++j) codes_out[j] = codes[j];
    for (int k = 0; k < N; ++k) { U[2 * k] = S.ux[k]; U[2 * k + 1] = S.uy[k]; }
    for (int k = 0; k <= N; ++k) { X[4 * k] = S.px[k]; X[4 * k + 1] = S.vx[k]; X[4 * k + 2] = S.py[k]; X[4 * k + 3] = S.vy[k]; }
    *obj = S.obj; *status = S.status; *iters = S.iters;
    (void)max_obs;
    return 0;
}

extern "C" int qp_host_solve_n3(const ldcbf_params* prm, int B, int max_obs, const double* x0, const double* theta0,
                                const double* goal, const int8_t* foot, const double* c_eta, const int32_t* nobs,
                                const double* delta, double* U, double* X, double* theta, double* omega, double* obj,
                                int32_t* status, int32_t* iters) {
    using namespace ldcbf;
    constexpr int N = 3, MO = 4;
    const StepConst C = make_const(*prm);
    for (int b = 0; b < B; ++b) {
        int ft[N + 1];
        for (int k = 0; k <= N; ++k) ft[k] = foot[b * (N + 1) + k];
        double4 ce[MO];
        const int nb = nobs[b] < MO ? nobs[b] : MO;
        for (int o = 0; o < MO; ++o) {
            const double* p = c_eta + ((size_t)b * max_obs + o) * 4;
            ce[o] = (o < nb) ? make_double4(p[0], p[1], p[2], p[3]) : make_double4(0, 0, 0, 0);
        }
        QpSolution<N> S;
        double ws[QpWorkspace<N>::DOUBLES];
        solve_scenario<N, MO, 1>(C, x0[4 * b], x0[4 * b + 1], x0[4 * b + 2], x0[4 * b + 3], theta0[b], goal[2 * b],
                              goal[2 * b + 1], ft, ce, nb,
                                 reinterpret_cast<const double4*>(c_eta + ((size_t)b * max_obs + MO) * 4),
                                 (nobs[b] < max_obs ? nobs[b] : max_obs) - nb, delta ? delta[b] : 0.0,
                                 load_limits(C, nullptr, 0), ws, S);
        for (int k = 0; k < N; ++k) { U[(b * N + k) * 2] = S.ux[k]; U[(b * N + k) * 2 + 1] = S.uy[k]; omega[b * N + k] = S.om[k]; }
        for (int k = 0; k <= N; ++k) {
            double* x = X + ((size_t)b * (N + 1) + k) * 4;
            x[0] = S.px[k]; x[1] = S.vx[k]; x[2] = S.py[k]; x[3] = S.vy[k];
            theta[b * (N + 1) + k] = S.th[k];
        }
        obj[b] = S.obj; status[b] = S.status; iters[b] = S.iters;
    }
    return 0;
}

// Cooperative variant (csrc/mpc_qp_coop.cuh) with a group of one lane: same QR-updated active-set code as the GPU
// kernel, the group collectives reduce to the identity.
#include "mpc_qp_coop.cuh"

template <int N, int MO>
static void coop_host_batch(const ldcbf_params* prm, int B, int max_obs, const double* x0, const double* theta0,
                            const double* goal, const int8_t* foot, const double* c_eta, const int32_t* nobs,
                            const double* delta, double* U, double* X, double* theta, double* omega, double* obj,
                            int32_t* status, int32_t* iters) {
    using namespace ldcbf;
    const StepConst C = make_const(*prm);
    LaneGroup<1> grp{1u, 0};
    for (int b = 0; b < B; ++b) {
        int ft[N + 1];
        for (int k = 0; k <= N; ++k) ft[k] = foot[b * (N + 1) + k];
        const int nb = nobs[b] < MO ? nobs[b] : MO;
        QpSolution<N> S;
        double sm[CoopShape<N, MO>::DOUBLES];
        coop_solve_scenario<N, MO, 1>(C, grp, x0[4 * b], x0[4 * b + 1], x0[4 * b + 2], x0[4 * b + 3], theta0[b],
                                      goal[2 * b], goal[2 * b + 1], ft,
                                      reinterpret_cast<const double4*>(c_eta + (size_t)b * max_obs * 4), nb,
                                      delta ? delta[b] : 0.0, load_limits(C, nullptr, 0), sm, S);
        for (int k = 0; k < N; ++k) { U[(b * N + k) * 2] = S.ux[k]; U[(b * N + k) * 2 + 1] = S.uy[k]; omega[b * N + k] = S.om[k]; }
        for (int k = 0; k <= N; ++k) {
            double* x = X + ((size_t)b * (N + 1) + k) * 4;
            x[0] = S.px[k]; x[1] = S.vx[k]; x[2] = S.py[k]; x[3] = S.vy[k];
            theta[b * (N + 1) + k] = S.th[k];
        }
        obj[b] = S.obj; status[b] = S.status; iters[b] = S.iters;
    }
}

extern "C" int qp_host_coop_solve(const ldcbf_params* prm, int B, int N, int max_obs, const double* x0,
                                  const double* theta0, const double* goal, const int8_t* foot, const double* c_eta,
                                  const int32_t* nobs, const double* delta, double* U, double* X, double* theta,
                                  double* omega, double* obj, int32_t* status, int32_t* iters) {
#define COOP_CASE(n, mo) coop_host_batch<n, mo>(prm, B, max_obs, x0, theta0, goal, foot, c_eta, nobs, delta, U, X, theta, omega, obj, status, iters)
    if (max_obs > 8) return -1;
    if (max_obs <= 4) {
        if (N == 1) COOP_CASE(1, 4); else if (N == 2) COOP_CASE(2, 4); else if (N == 3) COOP_CASE(3, 4); else return -1;
    } else {
        if (N == 1) COOP_CASE(1, 8); else if (N == 2) COOP_CASE(2, 8); else if (N == 3) COOP_CASE(3, 8); else return -1;
    }
#undef COOP_CASE
    return 0;
}

// Analysis aid: final active set (row codes, -1 padded) and the order in which rows entered, for one N = 3 scenario.
extern "C" int qp_host_coop_active_set(const ldcbf_params* prm, int max_obs, const double* x0, double theta0,
                                       const double* goal, const int8_t* foot, const double* c_eta, int nobs_b,
                                       double delta, int* codes_out, int* trips_out) {
    using namespace ldcbf;
    constexpr int N = 3, MO = 4;
    if (max_obs > MO) return -1;
    const StepConst C = make_const(*prm);
    LaneGroup<1> grp{1u, 0};
    int ft[N + 1];
    for (int k = 0; k <= N; ++k) ft[k] = foot[k];
    double sm[CoopShape<N, MO>::DOUBLES];
    CoopState<N> s;
    coop_setup<N, MO, 1>(C, grp, x0[0], x0[1], x0[2], x0[3], theta0, goal[0], goal[1], ft,
                         reinterpret_cast<const double4*>(c_eta), nobs_b < MO ? nobs_b : MO, delta,
                         load_limits(C, nullptr, 0), sm, s);
    while (!s.done) coop_trip<N, MO, 1>(C, grp, sm, s);
    for (int j = 0; j < 2 * N; ++j) codes_out[j] = j < s.q ? s.code[j] : -1;
    *trips_out = s.iters;
    return s.status;
}

// Analysis aid: per-trip trace (row code being added, active-set size after the trip, most negative slack) of one scenario.
extern "C" int qp_host_coop_trace(const ldcbf_params* prm, int max_obs, const double* x0, double theta0,
                                  const double* goal, const int8_t* foot, const double* c_eta, int nobs_b,
                                  double delta, int* trace /*[3*cap]*/, double* slack /*[cap]*/, int cap) {
    using namespace ldcbf;
    constexpr int N = 3, MO = 4;
    if (max_obs > MO) return -1;
    const StepConst C = make_const(*prm);
    LaneGroup<1> grp{1u, 0};
    int ft[N + 1];
    for (int k = 0; k <= N; ++k) ft[k] = foot[k];
    double sm[CoopShape<N, MO>::DOUBLES];
    CoopState<N> s;
    coop_setup<N, MO, 1>(C, grp, x0[0], x0[1], x0[2], x0[3], theta0, goal[0], goal[1], ft,
                         reinterpret_cast<const double4*>(c_eta), nobs_b < MO ? nobs_b : MO, delta,
                         load_limits(C, nullptr, 0), sm, s);
    int n = 0;
    while (!s.done) {
        const int q0 = s.q;
        int codes0[2 * N];
        for (int j = 0; j < 2 * N; ++j) codes0[j] = s.code[j];
        coop_trip<N, MO, 1>(C, grp, sm, s);
        if (n < cap && !s.done) {
            trace[3 * n] = s.p_code; trace[3 * n + 1] = s.q;
            int dropped = -1;
            if (s.q < q0) { for (int j = 0; j < q0; ++j) { bool found = false; for (int l = 0; l < s.q; ++l) found |= s.code[l] == codes0[j]; if (!found) dropped = codes0[j]; } }
            trace[3 * n + 2] = dropped;
            slack[n] = s.s_p;
            ++n;
        }
    }
    return n;
}

// The two paths the racing kernel (mpc_qp_race_kernel) runs per scenario: pref = 1 from the geometric guess,
// pref = 0 from the cold start.  N = 3, at most 4 obstacles.
extern "C" int qp_host_solve_n3_pref(const ldcbf_params* prm, int B, int max_obs, int pref, int use_guess,
                                     const double* x0, const double* theta0, const double* goal, const int8_t* foot,
                                     const double* c_eta, const int32_t* nobs, const double* delta, double* U, double* X,
                                     double* obj, int32_t* status, int32_t* iters) {
    using namespace ldcbf;
    constexpr int N = 3, MO = 4;
    if (max_obs > MO) return -1;
    const StepConst C = make_const(*prm);
    for (int b = 0; b < B; ++b) {
        int ft[N + 1];
        for (int k = 0; k <= N; ++k) ft[k] = foot[b * (N + 1) + k];
        double4 ce[MO];
        const int nb = nobs[b] < MO ? nobs[b] : MO;
        for (int o = 0; o < MO; ++o) {
            const double* p = c_eta + ((size_t)b * max_obs + o) * 4;
            ce[o] = (o < nb) ? make_double4(p[0], p[1], p[2], p[3]) : make_double4(0, 0, 0, 0);
        }
        double ws[QpWorkspace<N>::DOUBLES];
        QpState<N, MO> s;
        qp_setup<N, MO, 1>(C, x0[4 * b], x0[4 * b + 1], x0[4 * b + 2], x0[4 * b + 3], theta0[b], goal[2 * b],
                           goal[2 * b + 1], ft, ce, nb, nullptr, 0, delta ? delta[b] : 0.0, load_limits(C, nullptr, 0), ws, s);
        s.pref = pref;
        if (use_guess) {
            int codes[2 * N];
            guess_codes<N, MO>(s, codes);
            qp_warm_start<N, MO, 1>(C, codes, ws, s);
        }
        while (!s.done) qp_trip<N, MO, 1>(C, ws, s);
        QpSolution<N> S;
        qp_finish<N, MO>(C, s, S);
        for (int k = 0; k < N; ++k) { U[(b * N + k) * 2] = S.ux[k]; U[(b * N + k) * 2 + 1] = S.uy[k]; }
        for (int k = 0; k <= N; ++k) {
            double* xx = X + ((size_t)b * (N + 1) + k) * 4;
            xx[0] = S.px[k]; xx[1] = S.vx[k]; xx[2] = S.py[k]; xx[3] = S.vy[k];
        }
        obj[b] = S.obj; status[b] = S.status; iters[b] = S.iters;
    }
    return 0;
}

// Analysis aid: the kernel's geometric guess and the FINAL active set (row codes 2*id + side, -1 = empty slot) of one
// open-loop solve, to study how far the guess is from the optimum's active set.
extern "C" int qp_host_active_sets_n3(const ldcbf_params* prm, const double* x0, double theta0, const double* goal,
                                      const int8_t* foot, const double* c_eta, int nobs_b, double delta,
                                      const int* codes_in, int* guess_out, int* final_out, int32_t* status, int32_t* iters) {
    using namespace ldcbf;
    constexpr int N = 3, MO = 4;
    const StepConst C = make_const(*prm);
    int ft[N + 1];
    for (int k = 0; k <= N; ++k) ft[k] = foot[k];
    double4 ce[MO];
    const int nb = nobs_b < MO ? nobs_b : MO;
    for (int o = 0; o < MO; ++o) ce[o] = (o < nb) ? make_double4(c_eta[4 * o], c_eta[4 * o + 1], c_eta[4 * o + 2], c_eta[4 * o + 3]) : make_double4(0, 0, 0, 0);
    double ws[QpWorkspace<N>::DOUBLES];
    QpState<N, MO> s;
    qp_setup<N, MO, 1>(C, x0[0], x0[1], x0[2], x0[3], theta0, goal[0], goal[1], ft, ce, nb, nullptr, 0, delta,
                       load_limits(C, nullptr, 0), ws, s);
    int codes[2 * N];
    for (int j = 0; j < 2 * N; ++j) codes[j] = codes_in[j];
    if (codes_in[0] == -2) guess_codes<N, MO>(s, codes);
    for (int j = 0; j < 2 * N; ++j) guess_out[j] = codes[j];
    qp_warm_start<N, MO, 1>(C, codes, ws, s);
    while (!s.done) qp_trip<N, MO, 1>(C, ws, s);
    for (int j = 0; j < 2 * N; ++j) final_out[j] = ((s.amask >> j) & 1u) ? (int)ws[2 * (2 * N) * (2 * N) + j] : -1;
    *status = s.status; *iters = s.iters;
    return 0;
}

extern "C" long long qp_host_dependent_guesses(void) { return ldcbf::ldcbf_host_dependent_guesses; }
