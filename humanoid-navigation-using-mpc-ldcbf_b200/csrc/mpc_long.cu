// K2+K3 for long horizons (LDCBF_MAX_HORIZON < N <= LDCBF_MAX_HORIZON_LONG): one thread block per scenario.
//
// Same problem and same method as mpc_qp.cuh (heading schedule, rows and cost of HumanoidMpc.py:137-333 written in
// the CoM positions w = (p_1..p_N), Goldfarb-Idnani dual active set), but with up to 2N = 96 unknowns, 4N + N*n_obs
// rows and active sets of up to 2N rows the per-thread register formulation no longer fits, and the Gram/Cholesky
// factor of the active normals loses the degenerate vertices long horizons produce (leg row k, velocity rows k and
// k+1 are dependent to within the heading increment).  Here the block keeps an orthogonal factorisation of the
// active normals in shared memory, N_A = Q [R; 0] with Q = J (n x n) and R upper triangular:
//     d = J^T n+,   r = R^-1 d[0:q)  (dual step),   z = J[:, q:) d[q:)  (primal step),   |z|^2 = z.n+ = |d[q:)|^2
//     add a row:    one Householder reflection on the free columns of J, new column (d[0:q), -+|d[q:)|) of R
//     drop a row:   delete the column of R, restore the triangle with Givens rotations, same rotations on J
// which is backward stable: the 150-instance sweeps of tools/proto_long_horizon.py (N = 20, 40) match the
// Lawson-Hanson oracle (oracle/qp_pspace.py) to 1e-7.
//
// Work split inside a block (T threads): matrix-vector products by thread-per-column / thread-per-row over the padded
// (odd leading dimension: conflict-free both ways) shared arrays, the violation scan by thread-per-row with a block
// arg-min, the triangular solve by warp 0 (column sweep, one shuffle per column) while the other warps form z.
#include <limits.h>

#include "mpc_qp.cuh"
#include "step_io.cuh"

namespace ldcbf {

// R is stored packed: row i keeps the columns c >= i-1 (the triangle plus the one sub-diagonal that exists while a
// column is being deleted); element (i, c) lives at R[long_rbase(i, n) + c].
__host__ __device__ __forceinline__ int long_rbase(int i, int n) {
    return i == 0 ? 0 : n + (i - 1) * (n + 1) - ((i - 1) * i) / 2 - (i - 1);
}
__host__ __device__ __forceinline__ int long_rsize(int n) { return n + (n - 1) * (n + 1) - ((n - 1) * n) / 2; }

struct LongShared {
    double *J, *R, *d, *r, *z, *u, *cs, *sn, *rdi;      // factorisation and step vectors
    double *P, *V;                                       // iterate p_0..p_N (x, y interleaved) and velocities
    double *th, *rc, *rs, *om, *vmid, *vhalf, *vinrm;    // heading schedule, merged longitudinal bounds, 1/|row|
    double *ex, *ey, *hb, *eni;                          // half-planes: eta . p >= hb, 1/|eta|
    double *redv;
    int *ft, *acode, *redi;
    unsigned char* act;                                  // row id -> currently in the active set
};

__host__ __device__ inline size_t long_carve(int N, int max_obs, char* base, LongShared* S) {
    const int n = 2 * N, ld = n | 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base + off; off += (bytes + 15) & ~size_t(15); return p; };
    double* J = (double*)take(sizeof(double) * n * ld);
    double* R = (double*)take(sizeof(double) * long_rsize(n));
    double* vec = (double*)take(sizeof(double) * 7 * n);
    double* P = (double*)take(sizeof(double) * 4 * (N + 1));
    double* hs = (double*)take(sizeof(double) * 7 * (N + 1));
    double* ob = (double*)take(sizeof(double) * 4 * max_obs);
    double* redv = (double*)take(sizeof(double) * 32);
    int* ints = (int*)take(sizeof(int) * (N + 1 + n + 32));
    unsigned char* act = (unsigned char*)take((size_t)4 * N + (size_t)N * max_obs);
    if (S) {
        S->J = J; S->R = R;
        S->d = vec; S->r = vec + n; S->z = vec + 2 * n; S->u = vec + 3 * n; S->cs = vec + 4 * n; S->sn = vec + 5 * n;
        S->rdi = vec + 6 * n;
        S->P = P; S->V = P + 2 * (N + 1);
        S->th = hs; S->rc = hs + (N + 1); S->rs = hs + 2 * (N + 1); S->om = hs + 3 * (N + 1);
        S->vmid = hs + 4 * (N + 1); S->vhalf = hs + 5 * (N + 1); S->vinrm = hs + 6 * (N + 1);
        S->ex = ob; S->ey = ob + max_obs; S->hb = ob + 2 * max_obs; S->eni = ob + 3 * max_obs;
        S->redv = redv;
        S->ft = ints; S->acode = ints + (N + 1); S->redi = ints + (N + 1 + n);
        S->act = act;
    }
    return off;
}

struct LongScalars {
    double p0x, p0y, v0x, v0y, gx, gy, delta, vlat_mid, vlat_half;
    double rx, ry, nn, s_p, u_p, t, t1, d2n2;
    int q, status, iters, done, typ, k, code, full, dep, l, inner_done;
};

// Slack (natural units), deviation sign and 1/|normal| of row `row` at the current iterate.
__device__ __forceinline__ void long_eval_row(int row, int N, int nobs, const LongShared& S, const StepConst& C,
                                              const LongScalars& sc, double& slack, double& m, double& inrm) {
    if (row < 2 * N) {
        const int k = row >> 1;
        const double dx = S.P[2 * (k + 1)] - S.P[2 * k], dy = S.P[2 * (k + 1) + 1] - S.P[2 * k + 1];
        const double c = S.rc[k], s = S.rs[k];
        if ((row & 1) == 0) { m = (c * dx + s * dy) - C.legx_mid; slack = C.legx_half - fabs(m); }
        else { m = (c * dy - s * dx) - (C.legy_mid - (double)S.ft[k] * C.foot_offset); slack = C.legy_half - fabs(m); }
        inrm = k == 0 ? 1.0 : 0.70710678118654752;
    } else if (row < 4 * N) {
        const int kk = ((row - 2 * N) >> 1) + 1;
        const double Vx = S.V[2 * kk], Vy = S.V[2 * kk + 1];
        const double c = S.rc[kk], s = S.rs[kk];
        if ((row & 1) == 0) { m = (c * Vx + s * Vy) - S.vmid[kk]; slack = S.vhalf[kk] - fabs(m); }
        else { m = ((double)S.ft[kk] * c * Vy - s * Vx) - sc.vlat_mid; slack = sc.vlat_half - fabs(m); }
        inrm = S.vinrm[kk];
    } else {
        const int j = row - 4 * N;
        const int kk = j / nobs + 1, o = j - (kk - 1) * nobs;
        slack = S.ex[o] * S.P[2 * kk] + S.ey[o] * S.P[2 * kk + 1] - S.hb[o];
        m = -1.0;
        inrm = S.eni[o];
    }
}

// v_k = (-1)^k v_0 + gtil sum_{j<k} (-1)^{k-1-j} (p_{j+1} - p_j): one thread per (state, axis), two accumulators.
template <int T>
__device__ __forceinline__ void long_velocities(int N, double gtil, const LongShared& S, const LongScalars& sc) {
    for (int e = threadIdx.x; e < 2 * (N + 1); e += T) {
        const int k = e >> 1, ax = e & 1;
        double even = 0.0, odd = 0.0;          // sums over j with k-1-j even / odd
        int j = k - 1;
        for (; j >= 1; j -= 2) {
            even += S.P[2 * (j + 1) + ax] - S.P[2 * j + ax];
            odd += S.P[2 * j + ax] - S.P[2 * (j - 1) + ax];
        }
        if (j == 0) even += S.P[2 + ax] - S.P[ax];
        const double v0 = ax ? sc.v0y : sc.v0x;
        S.V[e] = ((k & 1) ? -v0 : v0) + gtil * (even - odd);
    }
}

template <int T>
__global__ void __launch_bounds__(T) mpc_long_kernel(StepConst C, int B, int N, int max_obs, int iter_cap, StepIO io) {
    extern __shared__ __align__(16) char long_smem[];
    __shared__ LongScalars sc;
    LongShared S;
    long_carve(N, max_obs, long_smem, &S);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = 2 * N, ld = n | 1;
    const double nan = quiet_nan();

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        const int nobs = min(io.nobs[b], max_obs);
        const int nrows = 4 * N + N * nobs;
        // ------------------------------------------------------------------ setup
        if (tid == 0) {
            double4 x;
            double th0;
            load_state(io, b, x, th0);
            const double2 g = reinterpret_cast<const double2*>(io.goal)[b];
            const Limits lim = load_limits(C, io.limits, (size_t)b);
            sc.p0x = x.x; sc.v0x = x.y; sc.p0y = x.z; sc.v0y = x.w; sc.gx = g.x; sc.gy = g.y;
            sc.delta = io.delta ? io.delta[b] : 0.0;
            sc.vlat_mid = 0.5 * (lim.vmax1 + C.v_min1); sc.vlat_half = 0.5 * (lim.vmax1 - C.v_min1);
            const double phi = atan2(g.y - x.z, g.x - x.x);            // HumanoidMpc.py:137-160
            double thk = th0;
            S.th[0] = thk; S.vmid[0] = 0.0; S.vhalf[0] = 0.0; S.vinrm[0] = 1.0;
            for (int k = 0; k < N; ++k) {
                const double w = fmin(fmax(phi - thk, lim.omega_min), lim.omega_max);
                S.om[k] = w;
                thk = add_rn(thk, mul_rn(w, C.sampling_time));
                S.th[k + 1] = thk;
                const double vhi = fmin(lim.vmax0, lim.vmax0 - lim.alpha_over_pi * fabs(w));
                S.vmid[k + 1] = 0.5 * (vhi + C.v_min0);
                S.vhalf[k + 1] = 0.5 * (vhi - C.v_min0);
                S.vinrm[k + 1] = 1.0 / (C.gtil * sqrt((double)(4 * k + 1)));   // |vel row k+1| = gtil sqrt(4(k+1)-3)
            }
            sc.status = LDCBF_STATUS_SOLVED; sc.iters = 0; sc.q = 0; sc.done = 0;
        }
        if (io.state6) {
            const int f0 = io.state6[6 * (size_t)b + 5] < 0.0 ? -1 : 1;
            for (int k = tid; k <= N; k += T) S.ft[k] = (k & 1) ? -f0 : f0;
        } else {
            for (int k = tid; k <= N; k += T) S.ft[k] = io.foot[(size_t)b * (N + 1) + k];
        }
        for (int i = tid; i < n * ld; i += T) S.J[i] = 0.0;
        for (int i = tid; i < long_rsize(n); i += T) S.R[i] = 0.0;
        for (int i = tid; i < nrows; i += T) S.act[i] = 0;
        __syncthreads();
        for (int k = tid; k <= N; k += T) {
            sincos(S.th[k], &S.rs[k], &S.rc[k]);
            S.P[2 * k] = k ? sc.gx : sc.p0x;
            S.P[2 * k + 1] = k ? sc.gy : sc.p0y;
        }
        for (int i = tid; i < n; i += T) { S.J[i * ld + i] = 1.0; S.u[i] = 0.0; }
        {
            const double4* gce = reinterpret_cast<const double4*>(io.c_eta) + (size_t)b * max_obs;
            for (int o = tid; o < nobs; o += T) {
                const double4 c4 = gce[o];
                S.ex[o] = c4.z; S.ey[o] = c4.w;
                S.hb[o] = c4.z * c4.x + c4.w * c4.y + sc.delta;
                S.eni[o] = rsqrt(c4.z * c4.z + c4.w * c4.w);
                if (!(c4.z == c4.z) || !(c4.w == c4.w)) atomicMax(&sc.status, LDCBF_STATUS_DEGENERATE);
                else if (c4.z * sc.p0x + c4.w * sc.p0y - S.hb[o] < -C.eps_const_row)     // constant k = 0 row
                    atomicMax(&sc.status, LDCBF_STATUS_INFEASIBLE);
            }
        }
        __syncthreads();
        if (tid == 0 && sc.status != LDCBF_STATUS_SOLVED) sc.done = 1;
        __syncthreads();

        // ------------------------------------------------------------------ active-set loop
        while (!sc.done) {
            // velocities at the iterate
            long_velocities<T>(N, C.gtil, S, sc);
            __syncthreads();
            // most violated row, in units of distance to the row's hyperplane
            double best = INFINITY;
            int bid = INT_MAX;
            for (int row = tid; row < nrows; row += T) {
                if (S.act[row]) continue;
                double sl, m, inrm;
                long_eval_row(row, N, nobs, S, C, sc, sl, m, inrm);
                sl *= inrm;
                if (sl < best) { best = sl; bid = row; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bid, o);
                if (ov < best || (ov == best && oi < bid)) { best = ov; bid = oi; }
            }
            if (lane == 0) { S.redv[warp] = best; S.redi[warp] = bid; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < T / 32; ++w)
                    if (S.redv[w] < best || (S.redv[w] == best && S.redi[w] < bid)) { best = S.redv[w]; bid = S.redi[w]; }
                if (!(best < -C.eps_active)) sc.done = 1;
                else {
                    double sl, m, inrm;
                    long_eval_row(bid, N, nobs, S, C, sc, sl, m, inrm);
                    const bool upper = bid < 4 * N && m >= 0.0;
                    const double sg = upper ? -1.0 : 1.0;
                    int typ, k;
                    double rx, ry, kap2;
                    if (bid < 2 * N) {
                        typ = 0; k = (bid >> 1) + 1;                       // row ends at state k, heading k-1
                        const double c = S.rc[k - 1], s = S.rs[k - 1];
                        if ((bid & 1) == 0) { rx = c; ry = s; } else { rx = -s; ry = c; }
                        kap2 = k > 1 ? 2.0 : 1.0;
                    } else if (bid < 4 * N) {
                        typ = 1; k = ((bid - 2 * N) >> 1) + 1;
                        const double c = S.rc[k], s = S.rs[k];
                        if ((bid & 1) == 0) { rx = c; ry = s; } else { rx = -s; ry = (double)S.ft[k] * c; }
                        kap2 = C.gtil * C.gtil * (double)(4 * k - 3);
                    } else {
                        const int j = bid - 4 * N;
                        typ = 2; k = j / nobs + 1;
                        const int o = j - (k - 1) * nobs;
                        rx = S.ex[o]; ry = S.ey[o];
                        kap2 = 1.0;
                    }
                    sc.typ = typ; sc.k = k; sc.rx = sg * rx; sc.ry = sg * ry;
                    sc.nn = (rx * rx + ry * ry) * kap2;
                    sc.s_p = sl; sc.u_p = 0.0;
                    sc.code = 2 * bid + (upper ? 1 : 0);
                }
                sc.inner_done = 0;
            }
            __syncthreads();
            if (sc.done) break;

            // ---- add row p: partial steps (dropping a row each) until the full step fits
            while (true) {
                const int q = sc.q;
                const int typ = sc.typ, k = sc.k;
                const double rx = sc.rx, ry = sc.ry;
                // d = J^T n+ ; n+ has blocks kap_m (rx, ry) on p_m, m = m_lo..k
                const int m_lo = typ == 2 ? k : (typ == 0 ? max(k - 1, 1) : 1);
                for (int j = tid; j < n; j += T) {
                    double acc = 0.0;
                    for (int m = m_lo; m <= k; ++m) {
                        double kap;
                        if (typ == 1) kap = (m == k) ? C.gtil : (((k - m) & 1) ? -2.0 * C.gtil : 2.0 * C.gtil);
                        else kap = (m == k) ? 1.0 : -1.0;
                        const double* Jm = S.J + (size_t)(2 * (m - 1)) * ld + j;
                        acc += kap * (rx * Jm[0] + ry * Jm[ld]);
                    }
                    S.d[j] = acc;
                }
                __syncthreads();
                if (warp == 0) {
                    // r = R^-1 d[0:q): column sweep in blocks of four columns; lane owns entries lane, lane+32, lane+64
                    // (a solved entry stays in its owner's register as r_j).  Only the shuffles and the small 4x4
                    // substitution are on the dependent chain; every shared-memory load is address-independent.
                    double dl0 = lane < q ? S.d[lane] : 0.0;
                    double dl1 = lane + 32 < q ? S.d[lane + 32] : 0.0;
                    double dl2 = lane + 64 < q ? S.d[lane + 64] : 0.0;
                    const int rb0 = long_rbase(lane, n), rb1 = long_rbase(lane + 32, n), rb2 = long_rbase(lane + 64, n);
                    auto pick = [&](int j) {
                        const int e = j >> 5;
                        return __shfl_sync(0xffffffffu, e == 0 ? dl0 : (e == 1 ? dl1 : dl2), j & 31);
                    };
                    auto put = [&](int j, double v) {
                        if (lane == (j & 31)) { const int e = j >> 5; if (e == 0) dl0 = v; else if (e == 1) dl1 = v; else dl2 = v; }
                    };
                    int j = q - 1;
                    for (; j >= 3; j -= 4) {
                        const double a0 = pick(j), a1 = pick(j - 1), a2 = pick(j - 2), a3 = pick(j - 3);
                        const double* R1 = S.R + long_rbase(j - 1, n);
                        const double* R2 = S.R + long_rbase(j - 2, n);
                        const double* R3 = S.R + long_rbase(j - 3, n);
                        const double r0 = a0 * S.rdi[j];
                        const double r1 = (a1 - R1[j] * r0) * S.rdi[j - 1];
                        const double r2 = (a2 - R2[j] * r0 - R2[j - 1] * r1) * S.rdi[j - 2];
                        const double r3 = (a3 - R3[j] * r0 - R3[j - 1] * r1 - R3[j - 2] * r2) * S.rdi[j - 3];
                        put(j, r0); put(j - 1, r1); put(j - 2, r2); put(j - 3, r3);
                        if (lane < j - 3) { const double* Rr = S.R + rb0 + j; dl0 -= Rr[0] * r0 + Rr[-1] * r1 + Rr[-2] * r2 + Rr[-3] * r3; }
                        if (lane + 32 < j - 3) { const double* Rr = S.R + rb1 + j; dl1 -= Rr[0] * r0 + Rr[-1] * r1 + Rr[-2] * r2 + Rr[-3] * r3; }
                        if (lane + 64 < j - 3) { const double* Rr = S.R + rb2 + j; dl2 -= Rr[0] * r0 + Rr[-1] * r1 + Rr[-2] * r2 + Rr[-3] * r3; }
                    }
                    for (; j >= 0; --j) {
                        const double rj = pick(j) * S.rdi[j];
                        put(j, rj);
                        if (lane < j) dl0 -= S.R[rb0 + j] * rj;
                        if (lane + 32 < j) dl1 -= S.R[rb1 + j] * rj;
                        if (lane + 64 < j) dl2 -= S.R[rb2 + j] * rj;
                    }
                    // dual step length t1 = min u_j / r_j over r_j > 0 (ties: lowest position)
                    const double rtol = 1e-13 * sqrt(sc.nn);
                    double t1 = INFINITY;
                    int l1 = INT_MAX;
                    if (lane < q) { S.r[lane] = dl0; if (dl0 > rtol) { t1 = fmax(S.u[lane], 0.0) / dl0; l1 = lane; } }
                    if (lane + 32 < q) {
                        S.r[lane + 32] = dl1;
                        if (dl1 > rtol) { const double v = fmax(S.u[lane + 32], 0.0) / dl1; if (v < t1) { t1 = v; l1 = lane + 32; } }
                    }
                    if (lane + 64 < q) {
                        S.r[lane + 64] = dl2;
                        if (dl2 > rtol) { const double v = fmax(S.u[lane + 64], 0.0) / dl2; if (v < t1) { t1 = v; l1 = lane + 64; } }
                    }
                    // |d[q:)|^2
                    double acc = 0.0;
                    for (int jj = q + lane; jj < n; jj += 32) acc += S.d[jj] * S.d[jj];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        const double ov = __shfl_xor_sync(0xffffffffu, t1, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, l1, o);
                        if (ov < t1 || (ov == t1 && oi < l1)) { t1 = ov; l1 = oi; }
                    }
                    if (lane == 0) { sc.d2n2 = acc; sc.t1 = t1; sc.l = l1 == INT_MAX ? -1 : l1; }
                } else {
                    // z = J[:, q:) d[q:)
                    for (int i = tid - 32; i < n; i += T - 32) {
                        const double* Ji = S.J + (size_t)i * ld;
                        double acc = 0.0;
                        for (int j = q; j < n; ++j) acc += Ji[j] * S.d[j];
                        S.z[i] = acc;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    const int it = ++sc.iters;
                    const double d2n2 = sc.d2n2;
                    const bool dep = !(d2n2 > 1e-18 * sc.nn) || q == n;
                    const int l = sc.l;
                    const double t2 = dep ? INFINITY : (-sc.s_p) / d2n2;
                    const bool full = !dep && (l < 0 || t2 <= sc.t1);
                    sc.full = full; sc.dep = dep;
                    if (it > iter_cap) { sc.status = LDCBF_STATUS_MAX_ITER; sc.done = 1; sc.inner_done = 1; }
                    else if (!full && l < 0) { sc.status = LDCBF_STATUS_INFEASIBLE; sc.done = 1; sc.inner_done = 1; }
                    else {
                        const double t = full ? t2 : sc.t1;
                        sc.t = t;
                        sc.u_p += t;
                        if (!dep) sc.s_p += t * d2n2;
                    }
                }
                __syncthreads();
                if (sc.inner_done) break;
                const double t = sc.t;
                const bool full = sc.full, dep = sc.dep;
                const int l = sc.l;
                for (int j = tid; j < q; j += T) S.u[j] -= t * S.r[j];
                if (!dep) for (int i = tid; i < n; i += T) S.P[2 + i] += t * S.z[i];
                if (full) {
                    // Householder reflection H = I - 2 v v^T / v^T v on the free columns, v = d[q:) - alpha e_0
                    const double dq = S.d[q];
                    const double alpha = dq > 0.0 ? -sqrt(sc.d2n2) : sqrt(sc.d2n2);
                    const double v0 = dq - alpha;
                    const double vn = sc.d2n2 - dq * dq + v0 * v0;
                    const double two_over_vn = 2.0 / vn;
                    for (int i = tid; i < n; i += T) {
                        double* Ji = S.J + (size_t)i * ld;
                        const double coef = (S.z[i] - alpha * Ji[q]) * two_over_vn;     // (J v)_i * 2 / |v|^2
                        Ji[q] -= coef * v0;
                        for (int j = q + 1; j < n; ++j) Ji[j] -= coef * S.d[j];
                    }
                    for (int i = tid; i < q; i += T) S.R[long_rbase(i, n) + q] = S.d[i];
                    __syncthreads();            // every thread has read sc.q / sc.u_p before thread 0 moves on
                    if (tid == 0) {
                        S.R[long_rbase(q, n) + q] = alpha;
                        S.rdi[q] = 1.0 / alpha;
                        S.acode[q] = sc.code;
                        S.u[q] = sc.u_p;
                        S.act[sc.code >> 1] = 1;
                        sc.q = q + 1;
                    }
                    __syncthreads();
                    break;
                }
                // ---- partial step: row in position l leaves the active set
                __syncthreads();                // u, P updated before the shifts below read u
                for (int i = tid; i < q; i += T) {
                    double* Ri = S.R + long_rbase(i, n);
                    for (int c = max(l, i - 1); c < q - 1; ++c) Ri[c] = Ri[c + 1];
                    Ri[q - 1] = 0.0;
                }
                if (tid == T - 1) {
                    S.act[S.acode[l] >> 1] = 0;
                    for (int c = l; c < q - 1; ++c) { S.u[c] = S.u[c + 1]; S.acode[c] = S.acode[c + 1]; }
                    S.u[q - 1] = 0.0;
                }
                __syncthreads();
                if (warp == 0) {
                    // Rotations i = l..q-2 on rows (i, i+1) of R.  Lane owns columns l + lane + 32 e; the upper-row
                    // entry of each owned column is carried in a register from one rotation to the next, the
                    // lower-row entries are untouched until their rotation (loaded one rotation ahead), and the
                    // pivot pair travels by shuffle: the dependent chain per rotation is shuffle -> rsqrt -> two FMAs.
                    const int c0 = l + lane, c1 = c0 + 32, c2 = c0 + 64, qe = q - 1;
                    const int rbl = long_rbase(l, n);
                    double x0 = c0 < qe ? S.R[rbl + c0] : 0.0, x1 = c1 < qe ? S.R[rbl + c1] : 0.0, x2 = c2 < qe ? S.R[rbl + c2] : 0.0;
                    int rbn = l + 1 < n ? long_rbase(l + 1, n) : 0;
                    double y0 = (c0 < qe && l + 1 <= c0 + 1) ? S.R[rbn + c0] : 0.0;
                    double y1 = (c1 < qe) ? S.R[rbn + c1] : 0.0;
                    double y2 = (c2 < qe) ? S.R[rbn + c2] : 0.0;
                    for (int i = l; i < qe; ++i) {
                        const int rbi = long_rbase(i, n), rbi1 = rbn;
                        // prefetch the lower row of the next rotation (row i+2), only where it is stored (c >= i+1)
                        double yn0 = 0.0, yn1 = 0.0, yn2 = 0.0;
                        if (i + 1 < qe) {
                            rbn = long_rbase(i + 2, n);
                            if (c0 < qe && c0 >= i + 1) yn0 = S.R[rbn + c0];
                            if (c1 < qe && c1 >= i + 1) yn1 = S.R[rbn + c1];
                            if (c2 < qe && c2 >= i + 1) yn2 = S.R[rbn + c2];
                        }
                        const int e = (i - l) >> 5, src = (i - l) & 31;
                        const double a = __shfl_sync(0xffffffffu, e == 0 ? x0 : (e == 1 ? x1 : x2), src);
                        const double bb = __shfl_sync(0xffffffffu, e == 0 ? y0 : (e == 1 ? y1 : y2), src);
                        const double h2 = a * a + bb * bb;
                        const double inv = h2 > 0.0 ? rsqrt_f64(h2) : 0.0;
                        const double c_ = h2 > 0.0 ? a * inv : 1.0, s_ = bb * inv;
                        if (c0 >= i && c0 < qe) {
                            const double nx = c_ * x0 + s_ * y0;
                            S.R[rbi + c0] = nx;
                            if (c0 == i) { S.R[rbi1 + c0] = 0.0; S.cs[i] = c_; S.sn[i] = s_; S.rdi[i] = inv; }
                            x0 = c_ * y0 - s_ * x0;
                        }
                        if (c1 >= i && c1 < qe) {
                            const double nx = c_ * x1 + s_ * y1;
                            S.R[rbi + c1] = nx;
                            if (c1 == i) { S.R[rbi1 + c1] = 0.0; S.cs[i] = c_; S.sn[i] = s_; S.rdi[i] = inv; }
                            x1 = c_ * y1 - s_ * x1;
                        }
                        if (c2 >= i && c2 < qe) {
                            const double nx = c_ * x2 + s_ * y2;
                            S.R[rbi + c2] = nx;
                            if (c2 == i) { S.R[rbi1 + c2] = 0.0; S.cs[i] = c_; S.sn[i] = s_; S.rdi[i] = inv; }
                            x2 = c_ * y2 - s_ * x2;
                        }
                        y0 = yn0; y1 = yn1; y2 = yn2;
                    }
                }
                __syncthreads();
                // the same rotations on the columns of J: each thread walks its row
                for (int r0 = tid; r0 < n; r0 += T) {
                    double* Ji = S.J + (size_t)r0 * ld;
                    double x = Ji[l];
                    for (int i = l; i < q - 1; ++i) {
                        const double y = Ji[i + 1];
                        const double c_ = S.cs[i], s_ = S.sn[i];
                        Ji[i] = c_ * x + s_ * y;
                        x = c_ * y - s_ * x;
                    }
                    Ji[q - 1] = x;
                }
                if (tid == 0) sc.q = q - 1;
                __syncthreads();
            }
            __syncthreads();
        }

        // ------------------------------------------------------------------ outputs
        __syncthreads();
        const bool ok = sc.status == LDCBF_STATUS_SOLVED;
        long_velocities<T>(N, C.gtil, S, sc);
        __syncthreads();
        if (tid == 0) {
            double obj = 0.0;
            for (int k = 0; k <= N; ++k) {
                const double ex = S.P[2 * k] - sc.gx, ey = S.P[2 * k + 1] - sc.gy;
                obj += ex * ex + ey * ey;
            }
            obj = ok ? obj : nan;
            if (io.obj) io.obj[b] = obj;
            if (io.status) io.status[b] = sc.status;
            if (io.iters) io.iters[b] = sc.iters;
            if (io.next10) {
                double* o = io.next10 + 10 * (size_t)b;
                const double u0x = (S.P[2] - C.ch * S.P[0] - C.sh_over_beta * S.V[0]) * C.inv_one_m_ch;
                const double u0y = (S.P[3] - C.ch * S.P[1] - C.sh_over_beta * S.V[1]) * C.inv_one_m_ch;
                o[0] = ok ? S.P[2] : nan; o[1] = ok ? S.V[2] : nan; o[2] = ok ? S.P[3] : nan; o[3] = ok ? S.V[3] : nan;
                o[4] = S.th[1]; o[5] = ok ? u0x : nan; o[6] = ok ? u0y : nan; o[7] = S.om[0];
                o[8] = obj; o[9] = (double)sc.status;
            }
        }
        if (io.U) {
            for (int k = tid; k <= N; k += T) {
                const bool live = ok || k == 0;
                double4 xk = make_double4(S.P[2 * k], S.V[2 * k], S.P[2 * k + 1], S.V[2 * k + 1]);
                if (!live) xk = make_double4(nan, nan, nan, nan);
                reinterpret_cast<double4*>(io.X)[(size_t)b * (N + 1) + k] = xk;
                io.theta[(size_t)b * (N + 1) + k] = S.th[k];
                if (k < N) {
                    const double ux = (S.P[2 * (k + 1)] - C.ch * S.P[2 * k] - C.sh_over_beta * S.V[2 * k]) * C.inv_one_m_ch;
                    const double uy = (S.P[2 * (k + 1) + 1] - C.ch * S.P[2 * k + 1] - C.sh_over_beta * S.V[2 * k + 1]) * C.inv_one_m_ch;
                    reinterpret_cast<double2*>(io.U)[(size_t)b * N + k] = ok ? make_double2(ux, uy) : make_double2(nan, nan);
                    io.omega[(size_t)b * N + k] = S.om[k];
                }
            }
        }
    }
}

template <int T>
static int launch_long(const StepConst& C, int B, int N, int max_obs, const StepIO& io, cudaStream_t st) {
    const size_t smem = long_carve(N, max_obs, nullptr, nullptr);
    if (smem > 227 * 1024) return LDCBF_E_SHAPE;
    auto kern = mpc_long_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 2048 / T ? 2048 / T : per_sm);
    const int grid = B < 148 * per_sm ? B : 148 * per_sm;
    const int iter_cap = C.max_iter * ((N + 3) / 4);
    kern<<<grid, T, smem, st>>>(C, B, N, max_obs, iter_cap, io);
    return check_launch();
}

int launch_long_horizon(const StepConst& C, int B, int N, int max_obs, const StepIO& io, cudaStream_t st) {
    if (N > LDCBF_MAX_HORIZON_LONG) return LDCBF_E_SHAPE;
    return 2 * N <= 32 ? launch_long<64>(C, B, N, max_obs, io, st) : launch_long<128>(C, B, N, max_obs, io, st);
}

}  // namespace ldcbf
