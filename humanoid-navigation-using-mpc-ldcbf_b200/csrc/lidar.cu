// K4 — 2-D LiDAR ray casting against convex polygons, with (obstacle, edge) hit indices.
//
// Replaces the reference's pure-Python triple loop (rays x obstacles x edges):
//   RangeFinder/range_finder_wth_polygons_dbscan.py:26-63  compute_lidar_readings
//   RangeFinder/range_finder_wth_polygons_dbscan.py:13-23  get_closest_point
//   Utils/obstacles.py:95-139                              line_polygon_intersection / compute_intersection
//
// Mapping: one 128-thread block per scan, each thread casts rays t, t+128, ...  The block first stages the scenario's
// vertex rings (16 B per vertex, one coalesced pass) into shared memory, then culls twice, both times provably
// without changing a single output bit (a culled pair yields exactly what its edge loop would have: no hit):
//   per scan: an obstacle whose bounding box is farther from the LiDAR than the range cannot be hit by any ray;
//   per ray:  an obstacle lies inside the disc around its bounding-box centre c with the half-diagonal rho as radius;
//             a ray from p in direction u can only meet that disc when  u.(c - p) >= sqrt(|c - p|^2 - rho^2),  one dot
//             product against a per-obstacle threshold (margins of 1e-9 against rounding of ~1e-16).
// The surviving obstacles go, in their original order (the reference's first-wins tie rule needs it), to a compact
// list; every thread then walks the edges of the few obstacles whose cone contains its ray, reading each edge as a
// shared-memory broadcast.  Config 3: 4 of 20 obstacles survive the scan test, 0.6 per ray the cone test.
// FP64-pipe / issue bound before the cone test (~30 flop per ray-edge test against 16*E + 24*R bytes per scan); with
// it the 24 B written per ray are a comparable share (HBM).
//
// Bit-exactness: each operation is an explicit round-to-nearest intrinsic in the oracle's order
// (oracle/lidar.py), no FMA contraction except the one inside numpy's 2-element dot (`dot2`), strict /
// non-strict comparisons and first-wins tie order as in the reference.  The ray table
// lidar_range*(cos, sin) is computed on the host with libm and passed in.  Compiled with -fmad=false.
#include "ldcbf_common.cuh"

namespace ldcbf {

constexpr int LIDAR_BLOCK = 128;

__global__ void __launch_bounds__(LIDAR_BLOCK) lidar_kernel(int R, const double2* __restrict__ ray_dirs, double lidar_range,
                                                          const double2* __restrict__ pos, int max_obs, int max_verts,
                                                          const double2* __restrict__ verts,
                                                          const int32_t* __restrict__ nverts,
                                                          const int32_t* __restrict__ nobs, int32_t* __restrict__ hit_obs,
                                                          int32_t* __restrict__ hit_edge, double2* __restrict__ hit_xy) {
    extern __shared__ double2 sv[];                                         // [max_obs][max_verts] vertex rings
    double4* cone = reinterpret_cast<double4*>(sv + (size_t)max_obs * max_verts);   // [max_obs] (c - p, threshold, -)
    int* snv = reinterpret_cast<int*>(cone + max_obs);                      // [max_obs] edge count, 0 = culled
    int* live = snv + max_obs;                                              // [max_obs] surviving obstacles, in order
    __shared__ int n_live;
    const int b = blockIdx.x;
    const int no = min(nobs[b], max_obs);
    const double2* gv = verts + (size_t)b * max_obs * max_verts;
    for (int i = threadIdx.x; i < no * max_verts; i += LIDAR_BLOCK) sv[i] = gv[i];
    for (int i = threadIdx.x; i < no; i += LIDAR_BLOCK) snv[i] = min(nverts[(size_t)b * max_obs + i], max_verts);
    __syncthreads();
    const double2 p = pos[b];
    // -- per-scan cull (bounding box out of reach) and the cone of every survivor; warp 0 compacts the list in order
    if (threadIdx.x < 32) {
        const double reach = lidar_range * (1.0 + 1e-9) + 1e-300;
        int count = 0;
        for (int base = 0; base < no; base += 32) {
            const int o = base + threadIdx.x;
            bool keep = false;
            double4 cn = make_double4(0.0, 0.0, 0.0, 0.0);
            if (o < no && snv[o] > 0) {
                const int n = snv[o];
                const double2* ring = sv + (size_t)o * max_verts;
                double lox = ring[0].x, hix = lox, loy = ring[0].y, hiy = loy;
                for (int e = 1; e < n; ++e) {
                    lox = fmin(lox, ring[e].x); hix = fmax(hix, ring[e].x);
                    loy = fmin(loy, ring[e].y); hiy = fmax(hiy, ring[e].y);
                }
                const double dx = fmax(fmax(lox - p.x, p.x - hix), 0.0), dy = fmax(fmax(loy - p.y, p.y - hiy), 0.0);
                keep = !(dx * dx + dy * dy > reach * reach);
                // disc around the box centre that contains the box (hence the polygon and every intersection point)
                const double wx = 0.5 * (lox + hix) - p.x, wy = 0.5 * (loy + hiy) - p.y;
                const double ex = hix - lox, ey = hiy - loy;
                const double rho2 = (0.25 * (ex * ex + ey * ey)) * (1.0 + 1e-9) + 1e-18;
                const double d2 = wx * wx + wy * wy;
                const double gap2 = d2 - rho2;
                // ray r passes the test when rd.w >= thr; -inf when the LiDAR is inside (or within rounding of) the disc
                const double thr = gap2 > 1e-9 * d2 ? lidar_range * (sqrt(gap2) - 1e-9 * (sqrt(d2) + 1.0)) : -INFINITY;
                cn = make_double4(wx, wy, thr, 0.0);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int k = count + __popc(m & ((1u << threadIdx.x) - 1u));
                live[k] = o;
                cone[k] = cn;
            }
            count += __popc(m);
        }
        if (threadIdx.x == 0) n_live = count;
    }
    __syncthreads();
    const int nl = n_live;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);

    for (int r = threadIdx.x; r < R; r += LIDAR_BLOCK) {
        const double2 rd = ray_dirs[r];
        const double a1x = p.x, a1y = p.y;
        const double b1x = __dadd_rn(p.x, rd.x), b1y = __dadd_rn(p.y, rd.y);     // ray end (`:39`)
        const double d1x = __dsub_rn(b1x, a1x), d1y = __dsub_rn(b1y, a1y);
        double min_distance = lidar_range;
        int ho = -1, he = -1;
        double hx = nanv, hy = nanv;
        for (int k = 0; k < nl; ++k) {
            const double4 cn = cone[k];
            if (!(rd.x * cn.x + rd.y * cn.y >= cn.z)) continue;                  // ray misses the obstacle's disc
            const int o = live[k];
            const int n = snv[o];
            const double2* ring = sv + (size_t)o * max_verts;
            double distance = lidar_range;                  // get_closest_point `:15`
            int be = -1;
            double bx = 0.0, by = 0.0;
            double2 A = ring[0];
            for (int e = 0; e < n; ++e) {
                const double2 Bv = ring[(e + 1 == n) ? 0 : e + 1];
                const double e2x = __dsub_rn(Bv.x, A.x), e2y = __dsub_rn(Bv.y, A.y);       // b2 - a2
                const double wx = __dsub_rn(a1x, A.x), wy = __dsub_rn(a1y, A.y);           // a1 - a2
                const double denom = __dsub_rn(__dmul_rn(e2y, d1x), __dmul_rn(e2x, d1y));
                if (denom != 0.0) {
                    // 0 <= ua <= 1 and 0 <= ub <= 1 decided WITHOUT the divisions: for IEEE doubles
                    //   num/denom >= 0  <=>  num == 0 or sign(num) == sign(denom)      (-0.0 >= 0 holds in Python too)
                    //   num/denom <= 1  <=>  |num| <= |denom|   (|num| > |denom| gives a quotient >= 1 + 2^-52, which
                    //                                            is representable, so rounding cannot pull it down to 1)
                    // so the two ~25-instruction divisions are only paid by the few (ray, edge) pairs that really hit.
                    const double na = __dsub_rn(__dmul_rn(e2x, wy), __dmul_rn(e2y, wx));
                    const double nb = __dsub_rn(__dmul_rn(d1x, wy), __dmul_rn(d1y, wx));
                    const double ad = fabs(denom);
                    const bool dneg = denom < 0.0;
                    const bool oka = (na == 0.0 || ((na < 0.0) == dneg)) && fabs(na) <= ad;
                    const bool okb = (nb == 0.0 || ((nb < 0.0) == dneg)) && fabs(nb) <= ad;
                    if (oka && okb) {
                        const double ua = __ddiv_rn(na, denom);
                        const double x = __dadd_rn(a1x, __dmul_rn(ua, d1x));
                        const double y = __dadd_rn(a1y, __dmul_rn(ua, d1y));
                        const double dx = __dsub_rn(x, p.x), dy = __dsub_rn(y, p.y);
                        const double curr = __dsqrt_rn(__fma_rn(dy, dy, __dmul_rn(dx, dx)));
                        if (curr < distance) { distance = curr; be = e; bx = x; by = y; }
                    }
                }
                A = Bv;
            }
            if (be >= 0 && distance <= lidar_range && distance < min_distance) {              // `:57`
                min_distance = distance; ho = o; he = be; hx = bx; hy = by;
            }
        }
        const size_t out = (size_t)b * R + r;
        hit_obs[out] = ho;
        hit_edge[out] = he;
        hit_xy[out] = make_double2(hx, hy);
    }
}

}  // namespace ldcbf

extern "C" int ldcbf_lidar_cast_f64(int B, int R, const double* ray_dirs, double lidar_range, const double* pos,
                                    int max_obs, int max_verts, const double* verts, const int32_t* nverts,
                                    const int32_t* nobs, int32_t* hit_obs, int32_t* hit_edge, double* hit_xy,
                                    void* cuda_stream) {
    using namespace ldcbf;
    if (B < 0 || R <= 0 || max_obs <= 0 || max_verts <= 0) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!ray_dirs || !pos || !verts || !nverts || !nobs || !hit_obs || !hit_edge || !hit_xy) return LDCBF_E_ARG;
    const size_t smem = (size_t)max_obs * max_verts * sizeof(double2) + (size_t)max_obs * (sizeof(double4) + 2 * sizeof(int));
    if (smem > 200 * 1024) return LDCBF_E_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(lidar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    }
    lidar_kernel<<<(unsigned)B, LIDAR_BLOCK, smem, st>>>(R, reinterpret_cast<const double2*>(ray_dirs), lidar_range,
                                              reinterpret_cast<const double2*>(pos), max_obs, max_verts,
                                              reinterpret_cast<const double2*>(verts), nverts, nobs, hit_obs,
                                              hit_edge, reinterpret_cast<double2*>(hit_xy));
    return check_launch();
}
