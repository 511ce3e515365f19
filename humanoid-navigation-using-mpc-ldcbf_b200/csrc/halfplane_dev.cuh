// Per-edge arithmetic of the half-plane builder, shared by the batched K1 kernel (halfplane.cu) and the
// closed-loop rollout kernel (rollout.cu) so that both produce identical (c, eta).
//
// Reference: Utils/ObstaclesUtils.py:77-96 (clamped projection on every hull edge, first strict minimum of the
// distance), :98-107 (eta = (x - c)/||x - c||, negated inside) and the crossing test behind :50-57.
//
// Two arithmetic modes (template parameter EXACT):
//  * EXACT = true (default of the library): every operation is an explicit round-to-nearest intrinsic in the
//    order of oracle/halfplane.py — including numpy's fma(a1, b1, a0*b0) 2-element dot, sqrt(dot)**2 for
//    np.power(np.linalg.norm(AB), 2) and the rounded distance in the `dist < min_dist` test — so (c, eta) are
//    bit-equal to the oracle, which applies the reference's per-edge arithmetic along the hull ring (the reference
//    walks ConvexHull.simplices: another edge order, results equal to ~1e-13).  This matters in closed loop: an active LDCBF row puts the next CoM exactly on
//    an obstacle edge, where eta = (x - c)/||x - c|| is decided by the last bits of the arithmetic.
//  * EXACT = false (LDCBF_FLAG_FAST_GEOMETRY): same algorithm to ~1 ulp with one reciprocal instead of two
//    square roots and a division per edge and squared-distance comparisons; 1.8x fewer FP64 instructions.
#pragma once
#include "ldcbf_common.cuh"

namespace ldcbf {

__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) {
    return __fma_rn(a1, b1, __dmul_rn(a0, b0));   // numpy's 2-element dot on the reference build
}

// The reference keeps the first edge whose ROUNDED distance sqrt(d2) is strictly smaller (`dist < min_dist`,
// ObstaclesUtils.py:91-94).  key_less<true>(a, b) decides RN(sqrt(a)) < RN(sqrt(b)) from the squared distances
// without taking a square root in all but near-tie cases:  a >= b  =>  sqrt(a) >= sqrt(b)  =>  not smaller (this
// covers the common exact tie of two edges clamping to their shared vertex);  a < b (1 - 2^-48)  =>  sqrt(a) <
// sqrt(b) (1 - 2^-49), which two roundings of relative size 2^-53 cannot reverse or close;  in between (a relative
// gap below 2^-48: practically never) both roots are taken.  Saves one of the two square roots per edge.
template <bool EXACT>
__device__ __forceinline__ bool key_less(double a, double b) {
    if (!EXACT) return a < b;
    if (!(a < b)) return false;
    if (a < __fma_rn(-0x1p-48, b, b)) return true;
    return __dsqrt_rn(a) < __dsqrt_rn(b);
}
constexpr double KEY_NONE = 1.0e300;      // "no edge yet": finite so that the threshold above stays finite

// Closest point of segment AB to P.  Returns the squared distance d2 (EXACT: dot2 in the reference's operation
// order, compared through key_less), writes c; `cross` is incremented when the edge A->B toggles the crossing parity.
// np.power(np.linalg.norm(AB), 2) of an edge: a property of the map, not of the query point
__device__ __forceinline__ double edge_den_exact(double2 A, double2 Bv) {
    const double abx = __dsub_rn(Bv.x, A.x), aby = __dsub_rn(Bv.y, A.y);
    const double nrm = __dsqrt_rn(dot2(abx, aby, abx, aby));
    return __dmul_rn(nrm, nrm);
}

// den_pre: the edge's edge_den_exact() computed earlier (the closed-loop kernel tabulates it once per run: the square
// root and the product are then off the per-step critical path, same bits), or a negative value to compute it here.
// The crossing-parity part of edge_closest alone (same operations, same rounding).
template <bool EXACT>
__device__ __forceinline__ void edge_cross(double px, double py, double2 A, double2 Bv, int& cross) {
    const bool f0 = A.y >= py, f1 = Bv.y >= py;
    if (f0 != f1) {
        const bool side = EXACT ? __dmul_rn(__dsub_rn(Bv.y, py), __dsub_rn(A.x, Bv.x)) >=
                                  __dmul_rn(__dsub_rn(Bv.x, px), __dsub_rn(A.y, Bv.y))
                                : (Bv.y - py) * (A.x - Bv.x) >= (Bv.x - px) * (A.y - Bv.y);
        cross += (side == f1);
    }
}

template <bool EXACT, bool CROSS = true>
__device__ __forceinline__ double edge_closest(double px, double py, double2 A, double2 Bv, double& cx, double& cy,
                                               int& cross, double den_pre = -1.0) {
    const bool f0 = CROSS && A.y >= py, f1 = CROSS && Bv.y >= py;
    if (EXACT) {
        const double apx = __dsub_rn(px, A.x), apy = __dsub_rn(py, A.y);
        const double abx = __dsub_rn(Bv.x, A.x), aby = __dsub_rn(Bv.y, A.y);
        const double den = den_pre >= 0.0 ? den_pre : edge_den_exact(A, Bv);   // np.power(np.linalg.norm(AB), 2)
        double t = __ddiv_rn(dot2(apx, apy, abx, aby), den);
        t = (t != t) ? 1.0 : fmax(0.0, fmin(1.0, t));                    // max(0, min(1, t)); NaN -> 1
        cx = __dadd_rn(A.x, __dmul_rn(t, abx));
        cy = __dadd_rn(A.y, __dmul_rn(t, aby));
        const double dx = __dsub_rn(cx, px), dy = __dsub_rn(cy, py);
        if (f0 != f1) {
            const bool side = __dmul_rn(__dsub_rn(Bv.y, py), __dsub_rn(A.x, Bv.x)) >=
                              __dmul_rn(__dsub_rn(Bv.x, px), __dsub_rn(A.y, Bv.y));
            cross += (side == f1);
        }
        return dot2(dx, dy, dx, dy);
    } else {
        const double apx = px - A.x, apy = py - A.y;
        const double abx = Bv.x - A.x, aby = Bv.y - A.y;
        const double den = abx * abx + aby * aby;
        double t = (apx * abx + apy * aby) * __drcp_rn(den);
        t = (t != t) ? 1.0 : fmax(0.0, fmin(1.0, t));
        cx = A.x + t * abx;
        cy = A.y + t * aby;
        const double dx = cx - px, dy = cy - py;
        if (f0 != f1) {
            const bool side = (Bv.y - py) * (A.x - Bv.x) >= (Bv.x - px) * (A.y - Bv.y);
            cross += (side == f1);
        }
        return dx * dx + dy * dy;
    }
}

// eta = (P - c)/||P - c||, negated when P is inside (ObstaclesUtils.py:98-107); NaN when P == c (:104)
template <bool EXACT>
__device__ __forceinline__ double4 finish_halfplane(double px, double py, double cx, double cy, int cross) {
    if (EXACT) {
        const double nx = __dsub_rn(px, cx), ny = __dsub_rn(py, cy);
        const double nn = __dsqrt_rn(dot2(nx, ny, nx, ny));
        double ex = __ddiv_rn(nx, nn), ey = __ddiv_rn(ny, nn);
        if (cross & 1) { ex = -ex; ey = -ey; }
        return make_double4(cx, cy, ex, ey);
    } else {
        const double nx = px - cx, ny = py - cy;
        const double nn = nx * nx + ny * ny;
        double inv = rsqrt(nn);
        inv = inv * (1.5 - 0.5 * nn * inv * inv);                        // one Newton step: full fp64 accuracy
        if (!(nn > 0.0)) inv = __longlong_as_double(0x7ff8000000000000LL);
        if (cross & 1) inv = -inv;
        return make_double4(cx, cy, nx * inv, ny * inv);
    }
}

// Serial walk over a ring of V vertices (global or shared memory): first strict minimum, edge order 0..V-1.
template <bool EXACT>
__device__ __forceinline__ double4 halfplane_serial(double px, double py, const double2* ring, int V) {
    double best = KEY_NONE, bcx = 0.0, bcy = 0.0;
    int cross = 0;
    double2 A = ring[0];
    for (int e = 0; e < V; ++e) {
        const double2 Bv = ring[(e + 1 == V) ? 0 : e + 1];
        double cx, cy;
        const double key = edge_closest<EXACT>(px, py, A, Bv, cx, cy, cross);
        if (key_less<EXACT>(key, best)) { best = key; bcx = cx; bcy = cy; }
        A = Bv;
    }
    return finish_halfplane<EXACT>(px, py, bcx, bcy, cross);
}

// The same walk with one 256-bit load per TWO vertices (LDG.E.256, sm_100): `ring` must be 32-byte aligned and the
// slot must hold an even number of vertices (the padded layout with an even max_verts).  Why: with one thread per ring
// the 32 lanes of a warp read 32 different lines, every 16-byte load costs a wavefront of the L1 data pipe per lane,
// and ncu put that pipe at 73 % of its peak — the kernel's top limiter, above DRAM (55 %) and the FP64 pipe (44 %).
// A 32-byte load moves the whole sector a lane touches in one wavefront: half the wavefronts for the same bytes.
// Same edges in the same order with the same arithmetic as halfplane_serial: bit-identical (c, eta).
struct __align__(32) VertexPair { double2 a, b; };
__device__ __forceinline__ VertexPair ldg_pair(const double2* p) {
    VertexPair r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.a.x), "=d"(r.a.y), "=d"(r.b.x), "=d"(r.b.y) : "l"(p));
    return r;
}
template <bool EXACT>
__device__ __forceinline__ double4 halfplane_serial_wide(double px, double py, const double2* ring, int V) {
    double best = KEY_NONE, bcx = 0.0, bcy = 0.0;
    int cross = 0;
    VertexPair cur = ldg_pair(ring);
    const double2 v0 = cur.a;
    double cx, cy, key;
    int e = 0;
    for (; e + 1 < V; e += 2) {                           // edges e and e + 1: both start inside `cur`
        VertexPair nxt;                                   // in flight while the two edges are evaluated
        if (e + 2 < V) nxt = ldg_pair(ring + e + 2); else nxt.a = nxt.b = v0;     // the ring closes on vertex 0
        key = edge_closest<EXACT>(px, py, cur.a, cur.b, cx, cy, cross);
        if (key_less<EXACT>(key, best)) { best = key; bcx = cx; bcy = cy; }
        key = edge_closest<EXACT>(px, py, cur.b, nxt.a, cx, cy, cross);
        if (key_less<EXACT>(key, best)) { best = key; bcx = cx; bcy = cy; }
        cur = nxt;
    }
    if (e < V) {                                          // V odd: the last edge runs from vertex V - 1 back to vertex 0
        key = edge_closest<EXACT>(px, py, cur.a, v0, cx, cy, cross);
        if (key_less<EXACT>(key, best)) { best = key; bcx = cx; bcy = cy; }
    }
    return finish_halfplane<EXACT>(px, py, bcx, bcy, cross);
}

// G lanes share a ring: lane l takes edges l, l + G, ..., an xor butterfly merges the partial results by the rule of
// halfplane_split_kernel (order by the rounded distance, ties towards the lower edge index = the first strict minimum
// of the serial walk; crossing counts are summed).  Every lane returns the same (c, eta), bit-equal to the serial walk.
// LDG = false: `ring` may point to shared memory (the closed-loop kernel stages the scenario's map there): plain loads.
template <bool EXACT, int G, bool LDG = true>
__device__ __forceinline__ double4 halfplane_group(double px, double py, const double2* ring, int V, int lane,
                                                   unsigned gmask, const double* den = nullptr) {
    if (G == 1 && !den) return halfplane_serial<EXACT>(px, py, ring, V);
    double best = KEY_NONE, bcx = 0.0, bcy = 0.0;
    int be = 0x7fffffff, cross = 0;
    for (int e = lane; e < V; e += G) {
        const int e1 = (e + 1 == V) ? 0 : e + 1;
        const double2 A = LDG ? __ldg(ring + e) : ring[e], Bv = LDG ? __ldg(ring + e1) : ring[e1];
        double cx, cy;
        const double key = edge_closest<EXACT>(px, py, A, Bv, cx, cy, cross, (EXACT && den) ? den[e] : -1.0);
        if (key_less<EXACT>(key, best)) { best = key; bcx = cx; bcy = cy; be = e; }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const double ok = __shfl_xor_sync(gmask, best, off, G);
        const double ocx = __shfl_xor_sync(gmask, bcx, off, G), ocy = __shfl_xor_sync(gmask, bcy, off, G);
        const int oe = __shfl_xor_sync(gmask, be, off, G);
        cross += __shfl_xor_sync(gmask, cross, off, G);
        if (key_less<EXACT>(ok, best) || (!key_less<EXACT>(best, ok) && oe < be)) { best = ok; bcx = ocx; bcy = ocy; be = oe; }
    }
    return finish_halfplane<EXACT>(px, py, bcx, bcy, cross);
}

// The ring walk of the closed-loop kernel, with temporal coherence: between two steps the CoM moves a fraction of a
// metre, so the edge that was closest at the previous step (`prev`, kept by the caller per scenario and obstacle) is
// evaluated first; its squared distance U bounds the minimum from above, and every other edge whose enclosing disc
// (midpoint M, radius h = |AB| / 2) is farther than that is skipped without the division / square roots of the exact
// evaluation:  |P - X| >= |P - M| - h for X on AB, so with m2 = |P - M|^2, h2 = h^2, thr = U (1 + 1e-9):
//     m2 > h2  and  s = m2 + h2 - thr > 0  and  s^2 > 4 h2 m2 (1 + 1e-6)   =>   (|P - M| - h)^2 > thr   =>   skip.
// A skipped edge's exact squared distance exceeds U by more than 1e-9 relative (the margins dwarf the 1e-16 rounding of
// the test itself), so its rounded distance can neither beat nor tie the minimum: the result — the minimal rounded
// distance, ties to the lowest edge index, the rule the lanes are merged with anyway — is bit-identical to the full walk
// whatever `prev` is (first step: 0).  The crossing parity still visits every edge (two comparisons, rarely more).
// Config 2: 52 edges, ~4 exact evaluations per scenario and step instead of 52.
template <bool EXACT, int G>
__device__ __forceinline__ double4 halfplane_group_pruned(double px, double py, const double2* ring, int V, int lane,
                                                          unsigned gmask, const double* den, int& prev) {
    double best = KEY_NONE, bcx = 0.0, bcy = 0.0;
    int be = 0x7fffffff, cross = 0;
    const int e0 = (prev >= 0 && prev < V) ? prev : 0;
    if (lane == e0 % G) {                               // the owner of the previously closest edge evaluates it exactly
        const double2 A = ring[e0], Bv = ring[(e0 + 1 == V) ? 0 : e0 + 1];
        best = edge_closest<EXACT, false>(px, py, A, Bv, bcx, bcy, cross, (EXACT && den) ? den[e0] : -1.0);
        be = e0;
    }
    double U = best;
    if (G > 1) U = __shfl_sync(gmask, best, e0 % G, G);
    const double thr = U * (1.0 + 1e-9);
    for (int e = lane; e < V; e += G) {
        const double2 A = ring[e], Bv = ring[(e + 1 == V) ? 0 : e + 1];
        edge_cross<EXACT>(px, py, A, Bv, cross);
        if (e == e0) continue;
        const double mx = 0.5 * (A.x + Bv.x) - px, my = 0.5 * (A.y + Bv.y) - py;
        const double hx = 0.5 * (Bv.x - A.x), hy = 0.5 * (Bv.y - A.y);
        const double m2 = mx * mx + my * my, h2 = hx * hx + hy * hy;
        const double s = m2 + h2 - thr;
        if (m2 > h2 && s > 0.0 && s * s > 4.0 * h2 * m2 * (1.0 + 1e-6)) continue;
        double cx, cy;
        const double key = edge_closest<EXACT, false>(px, py, A, Bv, cx, cy, cross, (EXACT && den) ? den[e] : -1.0);
        if (key_less<EXACT>(key, best) || (!key_less<EXACT>(best, key) && e < be)) { best = key; bcx = cx; bcy = cy; be = e; }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const double ok = __shfl_xor_sync(gmask, best, off, G);
        const double ocx = __shfl_xor_sync(gmask, bcx, off, G), ocy = __shfl_xor_sync(gmask, bcy, off, G);
        const int oe = __shfl_xor_sync(gmask, be, off, G);
        cross += __shfl_xor_sync(gmask, cross, off, G);
        if (key_less<EXACT>(ok, best) || (!key_less<EXACT>(best, ok) && oe < be)) { best = ok; bcx = ocx; bcy = ocy; be = oe; }
    }
    prev = be;
    return finish_halfplane<EXACT>(px, py, bcx, bcy, cross);
}

}  // namespace ldcbf
