// TEST INFRASTRUCTURE: compiles the solver source of the CUDA kernels (csrc/mpc_qp.cuh) for the HOST so that the
// CPU test suite can run the exact same algorithm code without a GPU.  Never linked into libldcbf_b200.so.
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
    qp_warm_start<N, MO, 1>(C, codes, ws, s);
    while (!s.done) qp_trip<N, MO, 1>(C, ws, s);
    QpSolution<N> S;
    qp_finish<N, MO>(C, s, S);
    shift_codes<N, MO, 1>(s, ws, codes);
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
