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

// Development aid: -DLDCBF_LONG_PROFILE accumulates thread-0 cycle counts per phase of the active-set loop
// (read back with ldcbf_debug_long_profile); compiled out of the product library.
#ifdef LDCBF_LONG_PROFILE
__device__ unsigned long long g_long_prof[16];
#define PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PROF(i) { const long long now_ = clock64(); prof_acc[i] += (unsigned long long)(now_ - prof_t); prof_t = now_; }
#define PROF_FLUSH if (threadIdx.x == 0) { for (int i_ = 0; i_ < 12; ++i_) atomicAdd(&g_long_prof[i_], prof_acc[i_]); }
#else
#define PROF_DECL
#define PROF(i)
#define PROF_FLUSH
#endif

// R is stored packed: row i keeps the columns c >= i-1 (the triangle plus the one sub-diagonal that exists while a
// column is being deleted); element (i, c) lives at R[long_rbase(i, n) + c], and rbase(i+1) - rbase(i) = n - i.
__host__ __device__ __forceinline__ int long_rbase(int i, int n) { return n * i - (i * (i - 1)) / 2; }
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
    double tol;       // a row counts as violated below -tol: eps_active, eps_infeasible after the relaxed restart
    int q, status, iters, done, typ, k, code, full, dep, l, inner_done, restart;
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

// Solves rows [lo, hi) (hi - lo <= 32) of the upper-triangular system R r = rhs inside one warp: lane owns row
// lo + lane (`me`: right-hand side in, solution out), columns swept from hi-1 down, four at a time.  The dependent
// chain per group of four is shuffle -> 4x4 substitution -> one fused update; all shared-memory loads are
// address-independent of it.
__device__ __forceinline__ double long_tri32(const double* __restrict__ R, const double* __restrict__ rdi, int n, int lo,
                                             int hi, int lane, double me) {
    const double* Rme = R + long_rbase(lo + lane, n) + lo;      // my row, indexed by pivot lane
    int p = hi - lo - 1;
    for (; p >= 0 && (p & 3) != 3; --p) {
        const double rj = __shfl_sync(0xffffffffu, me, p) * rdi[lo + p];
        if (lane == p) me = rj;
        if (lane < p) me -= Rme[p] * rj;
    }
    for (; p >= 3; p -= 4) {
        const int j = lo + p;
        const double a0 = __shfl_sync(0xffffffffu, me, p), a1 = __shfl_sync(0xffffffffu, me, p - 1);
        const double a2 = __shfl_sync(0xffffffffu, me, p - 2), a3 = __shfl_sync(0xffffffffu, me, p - 3);
        const double* R1 = R + long_rbase(j - 1, n) + j;
        const double* R2 = R1 - (n - (j - 2));                  // rbase(i) - rbase(i-1) = n - (i-1) for i >= 2
        const double* R3 = R2 - (n - (j - 3));
        // partial sums advance as soon as each r is known: the chain is mul, (fma, mul) x 3, fma
        const bool upd = lane < p - 3;
        const double m0 = upd ? Rme[p] : 0.0, m1 = upd ? Rme[p - 1] : 0.0, m2 = upd ? Rme[p - 2] : 0.0, m3 = upd ? Rme[p - 3] : 0.0;
        const double r0 = a0 * rdi[j];
        double b1 = a1 - R1[0] * r0, b2 = a2 - R2[0] * r0, b3 = a3 - R3[0] * r0, mm = me - m0 * r0;
        const double r1 = b1 * rdi[j - 1];
        b2 -= R2[-1] * r1; b3 -= R3[-1] * r1; mm -= m1 * r1;
        const double r2 = b2 * rdi[j - 2];
        b3 -= R3[-2] * r2; mm -= m2 * r2;
        const double r3 = b3 * rdi[j - 3];
        mm -= m3 * r3;
        me = mm;
        if (lane == p) me = r0;
        if (lane == p - 1) me = r1;
        if (lane == p - 2) me = r2;
        if (lane == p - 3) me = r3;
    }
    return me;
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

template <int T, int MINB>
__global__ void __launch_bounds__(T, MINB) mpc_long_kernel(StepConst C, int B, int N, int max_obs, int iter_cap, StepIO io) {
    extern __shared__ __align__(16) char long_smem[];
    __shared__ LongScalars sc;
    LongShared S;
    long_carve(N, max_obs, long_smem, &S);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = 2 * N, ld = n | 1;
    const double nan = quiet_nan();

    PROF_DECL
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        PROF(0)
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
            sc.tol = C.eps_active; sc.restart = 0;
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
        PROF(1)
        while (!sc.done) {
            if (sc.restart) {
                // relaxed restart (same rule as mpc_qp.cuh:qp_trip): back to the empty active set at w = (g, .., g)
                for (int i = tid; i < n * ld; i += T) S.J[i] = 0.0;
                for (int i = tid; i < long_rsize(n); i += T) S.R[i] = 0.0;
                for (int i = tid; i < nrows; i += T) S.act[i] = 0;
                __syncthreads();
                for (int k = 1 + tid; k <= N; k += T) { S.P[2 * k] = sc.gx; S.P[2 * k + 1] = sc.gy; }
                for (int i = tid; i < n; i += T) { S.J[i * ld + i] = 1.0; S.u[i] = 0.0; }
                if (tid == 0) { sc.q = 0; sc.restart = 0; }
                __syncthreads();
            }
            // velocities at the iterate
            long_velocities<T>(N, C.gtil, S, sc);
            __syncthreads();
            PROF(2)
            // most violated row, in units of distance to the row's hyperplane
            double best = INFINITY;
            int bid = INT_MAX;
            for (int row = tid; row < 4 * N; row += T) {
                if (S.act[row]) continue;
                double sl, m, inrm;
                long_eval_row(row, N, nobs, S, C, sc, sl, m, inrm);
                sl *= inrm;
                if (sl < best) { best = sl; bid = row; }
            }
            if (nobs > 0) {                                  // LDCBF rows (k, o), o fastest; (k, o) advanced without a division
                const int dk = T / nobs, dob = T - dk * nobs;
                int kk = tid / nobs, o = tid - kk * nobs;
                while (kk < N) {
                    const int row = 4 * N + kk * nobs + o;
                    if (!S.act[row]) {
                        const double sl = (S.ex[o] * S.P[2 * (kk + 1)] + S.ey[o] * S.P[2 * (kk + 1) + 1] - S.hb[o]) * S.eni[o];
                        if (sl < best) { best = sl; bid = row; }
                    }
                    kk += dk; o += dob;
                    if (o >= nobs) { o -= nobs; ++kk; }
                }
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
                if (!(best < -sc.tol)) sc.done = 1;
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
            PROF(3)
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
                PROF(4)
                // r = R^-1 d[0:q): 32-row diagonal blocks from the bottom up, each solved by warp 0 (long_tri32), the rows
                // above a solved block updated by a thread-per-row product; the other warps form z = J[:, q:) d[q:)
                // during the first triangle
                for (int j = tid; j < q; j += T) S.r[j] = S.d[j];
                bool z_done = false;
                if (q > 0) __syncthreads();
                for (int E = (q - 1) >> 5; E >= 0 && q > 0; --E) {
                    const int lo = 32 * E, hi = min(q, lo + 32);
                    if (warp == 0) {
                        double me = lo + lane < hi ? S.r[lo + lane] : 0.0;
                        me = long_tri32(S.R, S.rdi, n, lo, hi, lane, me);
                        if (lo + lane < hi) S.r[lo + lane] = me;
                    } else if (!z_done) {
                        for (int i = tid - 32; i < n; i += T - 32) {
                            const double* Ji = S.J + (size_t)i * ld;
                            double acc = 0.0;
                            for (int j = q; j < n; ++j) acc += Ji[j] * S.d[j];
                            S.z[i] = acc;
                        }
                    }
                    z_done = true;
                    __syncthreads();
                    if (E > 0) {
                        for (int i = tid; i < lo; i += T) {
                            const double* Ri = S.R + long_rbase(i, n);
                            double acc = 0.0;
                            for (int j = lo; j < hi; ++j) acc += Ri[j] * S.r[j];
                            S.r[i] -= acc;
                        }
                        __syncthreads();
                    }
                }
                PROF(11)
                if (warp == 0) {
                    // dual step length t1 = min u_j / r_j over r_j > 0 (ties: lowest position), and |d[q:)|^2
                    const double rtol = 1e-13 * sqrt(sc.nn);
                    double t1 = INFINITY;
                    int l1 = INT_MAX;
                    for (int j = lane; j < q; j += 32) {
                        const double rj = S.r[j];
                        if (rj > rtol) { const double v = fmax(S.u[j], 0.0) / rj; if (v < t1) { t1 = v; l1 = j; } }
                    }
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
                } else if (!z_done) {
                    for (int i = tid - 32; i < n; i += T - 32) {
                        const double* Ji = S.J + (size_t)i * ld;
                        double acc = 0.0;
                        for (int j = q; j < n; ++j) acc += Ji[j] * S.d[j];
                        S.z[i] = acc;
                    }
                }
                __syncthreads();
                PROF(5)
                if (tid == 0) {
                    const int it = ++sc.iters;
                    const double d2n2 = sc.d2n2;
                    const bool dep = !(d2n2 > 1e-18 * sc.nn) || q == n;
                    const int l = sc.l;
                    const double t2 = dep ? INFINITY : (-sc.s_p) / d2n2;
                    const bool full = !dep && (l < 0 || t2 <= sc.t1);
                    sc.full = full; sc.dep = dep;
                    if (it > iter_cap) { sc.status = LDCBF_STATUS_MAX_ITER; sc.done = 1; sc.inner_done = 1; }
                    else if (!full && l < 0) {
                        if (sc.tol < C.eps_infeasible && -sc.s_p <= C.eps_infeasible) {
                            sc.tol = C.eps_infeasible; sc.restart = 1; sc.inner_done = 1;     // feasible set = a point to rounding
                        } else {
                            sc.status = LDCBF_STATUS_INFEASIBLE; sc.done = 1; sc.inner_done = 1;
                        }
                    }
                    else {
                        const double t = full ? t2 : sc.t1;
                        sc.t = t;
                        sc.u_p += t;
                        if (!dep) sc.s_p += t * d2n2;
                    }
                }
                __syncthreads();
                PROF(6)
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
                    PROF(7)
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
                PROF(8)
                // Rotations i = l..q-2 on rows (i, i+1) of R, in chunks of 32 pivot columns.  Inside a chunk warp 0 runs
                // the wavefront: lane owns column cb + lane, the upper-row entry is carried in a register from one
                // rotation to the next, the lower-row entry is loaded one rotation ahead, the pivot pair travels by
                // shuffle (chain per rotation: shuffle -> rsqrt -> two FMAs).  Columns to the right of the chunk catch
                // up afterwards, one thread per column, from the stored (c_i, s_i).
                for (int cb = l; cb < q - 1; cb += 32) {
                    const int qe = q - 1, ce = min(cb + 32, qe);
                    if (warp == 0) {
                        const int c = cb + lane;
                        const bool own = c < ce;
                        double x = own ? S.R[long_rbase(cb, n) + c] : 0.0;
                        int rbi = long_rbase(cb, n), rbi1 = long_rbase(cb + 1, n);
                        double y = own ? S.R[rbi1 + c] : 0.0;
                        for (int i = cb; i < ce; ++i) {
                            const int rbi2 = rbi1 + (n - i - 1);               // rbase(i+2)
                            double yn = 0.0;
                            if (own && c >= i + 1 && i + 2 <= qe) yn = S.R[rbi2 + c];
                            const double a = __shfl_sync(0xffffffffu, x, i - cb);
                            const double bb = __shfl_sync(0xffffffffu, y, i - cb);
                            const double h2 = a * a + bb * bb;
                            const double inv = h2 > 0.0 ? rsqrt_f64(h2) : 0.0;
                            const double c_ = h2 > 0.0 ? a * inv : 1.0, s_ = bb * inv;
                            if (own && c >= i) {
                                S.R[rbi + c] = c_ * x + s_ * y;
                                x = c_ * y - s_ * x;
                                if (c == i) { S.R[rbi1 + c] = 0.0; S.cs[i] = c_; S.sn[i] = s_; S.rdi[i] = inv; }
                            }
                            y = yn;
                            rbi = rbi1; rbi1 = rbi2;
                        }
                    }
                    __syncthreads();
                    if (ce < qe) {
                        for (int c = ce + tid; c < qe; c += T) {
                            int rbi = long_rbase(cb, n);
                            double x = S.R[rbi + c];
                            for (int i = cb; i < ce; ++i) {
                                const int rb_next = rbi + (n - i);            // rbase(i+1)
                                const double y = S.R[rb_next + c];
                                const double c_ = S.cs[i], s_ = S.sn[i];
                                S.R[rbi + c] = c_ * x + s_ * y;
                                x = c_ * y - s_ * x;
                                rbi = rb_next;
                            }
                            S.R[rbi + c] = x;
                        }
                        __syncthreads();
                    }
                }
                __syncthreads();
                PROF(9)
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
                PROF(10)
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
    PROF_FLUSH
}

template <int T, int MINB>
static int launch_long(const StepConst& C, int B, int N, int max_obs, const StepIO& io, cudaStream_t st) {
    const size_t smem = long_carve(N, max_obs, nullptr, nullptr);
    if (smem > 227 * 1024) return LDCBF_E_SHAPE;
    auto kern = mpc_long_kernel<T, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > MINB ? MINB : per_sm);
    const int grid = B < 148 * per_sm ? B : 148 * per_sm;
    const int iter_cap = C.max_iter * ((N + 3) / 4);
    kern<<<grid, T, smem, st>>>(C, B, N, max_obs, iter_cap, io);
    return check_launch();
}

#ifdef LDCBF_LONG_PROFILE
extern "C" void ldcbf_debug_long_profile(unsigned long long* out12, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out12, g_long_prof, 12 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_long_prof, z, sizeof(z)); }
}
#endif

int launch_long_horizon(const StepConst& C, int B, int N, int max_obs, const StepIO& io, cudaStream_t st) {
    if (N > LDCBF_MAX_HORIZON_LONG) return LDCBF_E_SHAPE;
    // block width and register budget by problem size: 2N <= 32 unknowns -> 64 threads, 8 blocks/SM (128 registers);
    // 2N <= 64 -> 128 threads, 4 blocks/SM (128 registers); larger -> 128 threads, 2 blocks/SM (shared-memory bound)
    if (2 * N <= 32) return launch_long<64, 8>(C, B, N, max_obs, io, st);
    if (2 * N <= 64) return launch_long<128, 4>(C, B, N, max_obs, io, st);
    return launch_long<128, 2>(C, B, N, max_obs, io, st);
}

}  // namespace ldcbf
