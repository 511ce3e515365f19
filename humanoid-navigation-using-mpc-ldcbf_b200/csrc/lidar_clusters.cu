// f1 — LiDAR post-processing on the device: noise injection, DBSCAN clustering and convex hulls of the clusters.
//
// Replaces, per scan, the host chain of the reference's unknown-environment step
//   RangeFinder/range_finder_wth_polygons_dbscan.py:161-172  Gaussian noise on valid readings (noise tensor injected)
//   RangeFinder/range_finder_wth_polygons_dbscan.py:100-116  retrieve_clusters: sklearn DBSCAN(eps=0.3, min_samples=3)
//   RangeFinder/range_finder_wth_polygons_dbscan.py:65-83    create_convex_hull: np.unique, < 3 points or rank < 2 -> None,
//                                                            scipy ConvexHull -> points[hull.vertices]
//   RangeFinder/range_finder_wth_polygons_dbscan.py:119-126  build_local_obstacles (the closing vertex it appends is
//                                                            dropped again by the ConvexHull call of
//                                                            HumanoidMPCUnknownEnvironment.py:55, so it is not emitted)
//
// One CTA per scan (R <= 512 rays).  DBSCAN on <= 512 points is done exactly as sklearn defines it:
//   * neighbourhood = points within eps (squared distances, self included); core = at least min_samples neighbours;
//   * clusters = connected components of core points, numbered by their smallest core-point index (sklearn visits
//     points in index order and opens a cluster at the first unvisited core point);
//   * a border point takes the lowest-numbered cluster among its core neighbours (it is labelled by the first
//     cluster that reaches it and never relabelled); everything else is noise (-1).
// The eps-graph is held as a bit matrix in shared memory (R x R/32 words), components by min-label propagation.
// Hulls: lexicographic bitonic sort of the cluster (np.unique order), duplicate removal, collinearity test, Andrew's
// monotone chain (counter-clockwise, strictly convex).  Qhull and the chain may disagree on which of several
// points that are collinear to 1e-16 is called a vertex; the polygons are the same to rounding.
#include "ldcbf_common.cuh"

namespace ldcbf {

constexpr int CL_RMAX = 512;
constexpr int CL_WORDS = CL_RMAX / 32;
constexpr int CL_THREADS = 256;
constexpr int CL_BIG = 0x3fffffff;

struct ClusterShared {
    double x[CL_RMAX], y[CL_RMAX];          // compacted valid points, ray order
    double sx[CL_RMAX], sy[CL_RMAX];        // sort buffer of one cluster
    unsigned adj[CL_RMAX][CL_WORDS];        // eps-graph
    unsigned coremask[CL_WORDS];
    int ray[CL_RMAX];                       // ray index of compacted point i
    int lab[CL_RMAX];                       // component label = smallest core index, CL_BIG = noise
    int cid[CL_RMAX];                       // cluster number of root i
    int hull[CL_RMAX + 2];
    int P, n_clusters, changed, cnt, hull_n;
    double red[CL_THREADS / 32][2];
};

__device__ __forceinline__ bool lex_less(double ax, double ay, double bx, double by) {
    return ax < bx || (ax == bx && ay < by);
}

__global__ void __launch_bounds__(CL_THREADS) lidar_clusters_kernel(int R, const double2* __restrict__ hit_xy,
                                                                    const double2* __restrict__ noise, double eps2,
                                                                    int min_samples, int max_hulls, int max_hull_verts,
                                                                    int32_t* __restrict__ labels,
                                                                    double2* __restrict__ hull_verts,
                                                                    int32_t* __restrict__ hull_nverts,
                                                                    int32_t* __restrict__ n_hulls,
                                                                    int32_t* __restrict__ overflow) {
    extern __shared__ unsigned char cl_raw[];
    ClusterShared& S = *reinterpret_cast<ClusterShared*>(cl_raw);
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
    const double2* scan = hit_xy + (size_t)b * R;

    // ---- 1. compact the valid readings in ray order (warp 0: ballot + popc per chunk of 32 rays)
    if (t < 32) {
        int base = 0;
        for (int c = 0; c < R; c += 32) {
            const int i = c + lane;
            double2 p = make_double2(0.0, 0.0);
            bool ok = false;
            if (i < R) {
                p = scan[i];
                ok = p.x == p.x;
                if (ok && noise) { const double2 nz = noise[(size_t)b * R + i]; p.x += nz.x; p.y += nz.y; }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int k = base + __popc(m & ((1u << lane) - 1u));
                S.x[k] = p.x; S.y[k] = p.y; S.ray[k] = i;
            }
            base += __popc(m);
        }
        if (lane == 0) { S.P = base; S.n_clusters = 0; }
    }
    for (int i = t; i < R; i += CL_THREADS) labels[(size_t)b * R + i] = -1;
    if (t < CL_WORDS) S.coremask[t] = 0u;
    __syncthreads();
    const int P = S.P;

    // ---- 2. eps-graph and core points
    for (int i = t; i < P; i += CL_THREADS) {
        const double xi = S.x[i], yi = S.y[i];
        int count = 0;
        for (int w = 0; w * 32 < P; ++w) {
            unsigned bits = 0u;
            const int jend = min(32, P - w * 32);
            for (int jj = 0; jj < jend; ++jj) {
                const double dx = S.x[w * 32 + jj] - xi, dy = S.y[w * 32 + jj] - yi;
                if (dx * dx + dy * dy <= eps2) bits |= 1u << jj;
            }
            S.adj[i][w] = bits;
            count += __popc(bits);
        }
        const bool core = count >= min_samples;
        S.lab[i] = core ? i : CL_BIG;
        if (core) atomicOr(&S.coremask[i >> 5], 1u << (i & 31));
    }
    __syncthreads();

    // ---- 3. connected components of the core points: min-label propagation to a fixed point
    for (;;) {
        if (t == 0) S.changed = 0;
        __syncthreads();
        for (int i = t; i < P; i += CL_THREADS) {
            if (!((S.coremask[i >> 5] >> (i & 31)) & 1u)) continue;
            int m = S.lab[i];
            for (int w = 0; w * 32 < P; ++w) {
                unsigned bits = S.adj[i][w] & S.coremask[w];
                while (bits) {
                    const int j = w * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    m = min(m, S.lab[j]);
                }
            }
            if (m < S.lab[i]) { S.lab[i] = m; S.changed = 1; }
        }
        __syncthreads();
        if (!S.changed) break;
        __syncthreads();
    }
    // ---- 4. border points: lowest-numbered cluster among the core neighbours
    for (int i = t; i < P; i += CL_THREADS) {
        if ((S.coremask[i >> 5] >> (i & 31)) & 1u) continue;
        int m = CL_BIG;
        for (int w = 0; w * 32 < P; ++w) {
            unsigned bits = S.adj[i][w] & S.coremask[w];
            while (bits) {
                const int j = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                m = min(m, S.lab[j]);
            }
        }
        S.lab[i] = m;
    }
    __syncthreads();
    // ---- 5. number the clusters by their smallest core index (= sklearn's label order)
    if (t < 32) {
        int base = 0;
        for (int c = 0; c < P; c += 32) {
            const int i = c + lane;
            const bool root = i < P && S.lab[i] == i;
            const unsigned m = __ballot_sync(0xffffffffu, root);
            if (root) S.cid[i] = base + __popc(m & ((1u << lane) - 1u));
            base += __popc(m);
        }
        if (lane == 0) S.n_clusters = base;
    }
    __syncthreads();
    for (int i = t; i < P; i += CL_THREADS)
        labels[(size_t)b * R + S.ray[i]] = (S.lab[i] == CL_BIG) ? -1 : S.cid[S.lab[i]];
    const int n_clusters = S.n_clusters;

    // ---- 6. convex hull of every cluster
    int n_out = 0;
    bool ovf = false;
    for (int c = 0; c < n_clusters; ++c) {
        __syncthreads();
        // gather the members (warp 0, index order)
        if (t < 32) {
            int base = 0;
            for (int q = 0; q < P; q += 32) {
                const int i = q + lane;
                const bool in = i < P && S.lab[i] != CL_BIG && S.cid[S.lab[i]] == c;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) { const int k = base + __popc(m & ((1u << lane) - 1u)); S.sx[k] = S.x[i]; S.sy[k] = S.y[i]; }
                base += __popc(m);
            }
            if (lane == 0) S.cnt = base;
        }
        __syncthreads();
        const int cnt = S.cnt;
        int n2 = 1;
        while (n2 < cnt) n2 <<= 1;
        for (int i = cnt + t; i < n2; i += CL_THREADS) { S.sx[i] = INFINITY; S.sy[i] = INFINITY; }
        __syncthreads();
        // lexicographic bitonic sort (np.unique(points, axis=0) order)
        for (int k = 2; k <= n2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = t; i < n2; i += CL_THREADS) {
                    const int l = i ^ j;
                    if (l > i) {
                        const bool up = (i & k) == 0;
                        const double ax = S.sx[i], ay = S.sy[i], bx = S.sx[l], by = S.sy[l];
                        if (lex_less(bx, by, ax, ay) == up) { S.sx[i] = bx; S.sy[i] = by; S.sx[l] = ax; S.sy[l] = ay; }
                    }
                }
                __syncthreads();
            }
        }
        // duplicates out, collinearity test, monotone chain: sequential on <= 512 points (thread 0)
        if (t == 0) {
            int u = 0;
            for (int i = 0; i < cnt; ++i)
                if (u == 0 || S.sx[i] != S.sx[u - 1] || S.sy[i] != S.sy[u - 1]) { S.sx[u] = S.sx[i]; S.sy[u] = S.sy[i]; ++u; }
            int h = 0;
            if (u >= 3) {
                // The reference drops a cluster when np.linalg.matrix_rank(points - points[0]) < 2 (`:75-76`) or when
                // Qhull finds the input flat to its roundoff bound and raises (`:81-83`; ~23 eps max|coordinate|).  Both
                // only ever trigger on readings taken on ONE straight edge (deviations ~1e-16); real corners deviate
                // by >= 1e-6.  One test covers both: largest distance from the line through the lexicographic
                // extremes <= 64 eps max|coordinate|.
                const double dxl = S.sx[u - 1] - S.sx[0], dyl = S.sy[u - 1] - S.sy[0];
                const double len = sqrt(dxl * dxl + dyl * dyl);
                double maxdev = 0.0, scale = 0.0;
                for (int i = 0; i < u; ++i) {
                    const double ex = S.sx[i] - S.sx[0], ey = S.sy[i] - S.sy[0];
                    maxdev = fmax(maxdev, fabs(dxl * ey - dyl * ex) / len);
                    scale = fmax(scale, fmax(fabs(S.sx[i]), fabs(S.sy[i])));
                }
                if (maxdev > 64.0 * 2.220446049250313e-16 * scale) {
                    // Andrew's monotone chain, counter-clockwise, collinear points dropped
                    for (int i = 0; i < u; ++i) {
                        while (h >= 2) {
                            const int a = S.hull[h - 2], bq = S.hull[h - 1];
                            const double cr = (S.sx[bq] - S.sx[a]) * (S.sy[i] - S.sy[a]) - (S.sy[bq] - S.sy[a]) * (S.sx[i] - S.sx[a]);
                            if (cr <= 0.0) --h; else break;
                        }
                        S.hull[h++] = i;
                    }
                    const int lower = h + 1;
                    for (int i = u - 2; i >= 0; --i) {
                        while (h >= lower) {
                            const int a = S.hull[h - 2], bq = S.hull[h - 1];
                            const double cr = (S.sx[bq] - S.sx[a]) * (S.sy[i] - S.sy[a]) - (S.sy[bq] - S.sy[a]) * (S.sx[i] - S.sx[a]);
                            if (cr <= 0.0) --h; else break;
                        }
                        S.hull[h++] = i;
                    }
                    --h;                       // the last point repeats the first
                    if (h < 3) h = 0;
                }
            }
            S.hull_n = h;
        }
        __syncthreads();
        const int h = S.hull_n;
        if (h >= 3) {
            if (n_out < max_hulls) {
                const int hv = min(h, max_hull_verts);
                if (h > max_hull_verts) ovf = true;
                double2* dst = hull_verts + ((size_t)b * max_hulls + n_out) * max_hull_verts;
                for (int i = t; i < max_hull_verts; i += CL_THREADS)
                    dst[i] = (i < hv) ? make_double2(S.sx[S.hull[i]], S.sy[S.hull[i]]) : make_double2(0.0, 0.0);
                if (t == 0) hull_nverts[(size_t)b * max_hulls + n_out] = hv;
                ++n_out;
            } else {
                ovf = true;
            }
        }
    }
    for (int o = n_out + t; o < max_hulls; o += CL_THREADS) hull_nverts[(size_t)b * max_hulls + o] = 0;
    if (t == 0) { n_hulls[b] = n_out; if (overflow) overflow[b] = ovf ? 1 : 0; }
}

}  // namespace ldcbf

extern "C" int ldcbf_lidar_clusters_f64(int B, int R, const double* hit_xy, const double* noise, double eps,
                                        int min_samples, int max_hulls, int max_hull_verts, int32_t* labels,
                                        double* hull_verts, int32_t* hull_nverts, int32_t* n_hulls, int32_t* overflow,
                                        void* cuda_stream) {
    using namespace ldcbf;
    if (B < 0 || R <= 0 || max_hulls <= 0 || max_hull_verts < 3 || min_samples <= 0 || !(eps > 0.0)) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!hit_xy || !labels || !hull_verts || !hull_nverts || !n_hulls) return LDCBF_E_ARG;
    if (R > CL_RMAX) return LDCBF_E_SHAPE;
    const size_t smem = sizeof(ClusterShared);
    cudaError_t e = cudaFuncSetAttribute(lidar_clusters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    lidar_clusters_kernel<<<B, CL_THREADS, smem, static_cast<cudaStream_t>(cuda_stream)>>>(
        R, reinterpret_cast<const double2*>(hit_xy), reinterpret_cast<const double2*>(noise), eps * eps, min_samples,
        max_hulls, max_hull_verts, labels, reinterpret_cast<double2*>(hull_verts), hull_nverts, n_hulls, overflow);
    return check_launch();
}
