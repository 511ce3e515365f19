// Per-edge arithmetic of the half-plane builder, shared by the batched K1 kernel (halfplane.cu) and the
// closed-loop rollout kernel (rollout.cu) so that both produce bit-identical (c, eta).
// Every operation is an explicit round-to-nearest intrinsic in the order of oracle/halfplane.py
// (reference: Utils/ObstaclesUtils.py:77-96 and the crossing test behind :50-57).
#pragma once
#include "ldcbf_common.cuh"

namespace ldcbf {

__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) {
    return __fma_rn(a1, b1, __dmul_rn(a0, b0));   // numpy's 2-element dot on the reference build
}

// Closest point of segment AB to P: returns the distance, writes c; `cross` is incremented when the edge
// A->B toggles the crossing-number parity of P.
__device__ __forceinline__ double edge_closest(double px, double py, double2 A, double2 Bv, double& cx, double& cy,
                                               int& cross) {
    const double apx = __dsub_rn(px, A.x), apy = __dsub_rn(py, A.y);
    const double abx = __dsub_rn(Bv.x, A.x), aby = __dsub_rn(Bv.y, A.y);
    const double nrm = __dsqrt_rn(dot2(abx, aby, abx, aby));
    const double den = __dmul_rn(nrm, nrm);                          // np.power(np.linalg.norm(AB), 2)
    double t = __ddiv_rn(dot2(apx, apy, abx, aby), den);
    t = (t != t) ? 1.0 : fmax(0.0, fmin(1.0, t));                    // max(0, min(1, t)); NaN -> 1
    cx = __dadd_rn(A.x, __dmul_rn(t, abx));
    cy = __dadd_rn(A.y, __dmul_rn(t, aby));
    const double dx = __dsub_rn(cx, px), dy = __dsub_rn(cy, py);
    const bool f0 = A.y >= py, f1 = Bv.y >= py;
    if (f0 != f1) {
        const bool side = __dmul_rn(__dsub_rn(Bv.y, py), __dsub_rn(A.x, Bv.x)) >=
                          __dmul_rn(__dsub_rn(Bv.x, px), __dsub_rn(A.y, Bv.y));
        cross += (side == f1);
    }
    return __dsqrt_rn(dot2(dx, dy, dx, dy));
}

// eta = (P - c)/||P - c||, negated when P is inside (ObstaclesUtils.py:98-107)
__device__ __forceinline__ double4 finish_halfplane(double px, double py, double cx, double cy, int cross) {
    const double nx = __dsub_rn(px, cx), ny = __dsub_rn(py, cy);
    const double nn = __dsqrt_rn(dot2(nx, ny, nx, ny));
    double ex = __ddiv_rn(nx, nn), ey = __ddiv_rn(ny, nn);
    if (cross & 1) { ex = -ex; ey = -ey; }
    return make_double4(cx, cy, ex, ey);
}

// Serial version for one thread: whole ring of V vertices.
__device__ __forceinline__ double4 halfplane_serial(double px, double py, const double2* __restrict__ ring, int V) {
    double best_d = INFINITY, bcx = 0.0, bcy = 0.0;
    int cross = 0;
    double2 A = __ldg(ring);
    for (int e = 0; e < V; ++e) {
        const double2 Bv = __ldg(ring + ((e + 1 == V) ? 0 : e + 1));
        double cx, cy;
        const double d = edge_closest(px, py, A, Bv, cx, cy, cross);
        if (d < best_d) { best_d = d; bcx = cx; bcy = cy; }
        A = Bv;
    }
    return finish_halfplane(px, py, bcx, bcy, cross);
}

}  // namespace ldcbf
