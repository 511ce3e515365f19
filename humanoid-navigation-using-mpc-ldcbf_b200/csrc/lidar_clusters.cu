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
// One 256-thread CTA per scan (R <= 512 rays).  DBSCAN on <= 512 points is done exactly as sklearn defines it:
//   * neighbourhood = points within eps (squared distances, self included); core = at least min_samples neighbours;
//   * clusters = connected components of core points, numbered by their smallest core-point index (sklearn visits
//     points in index order and opens a cluster at the first unvisited core point);
//   * a border point takes the lowest-numbered cluster among its core neighbours (it is labelled by the first
//     cluster that reaches it and never relabelled); everything else is noise (-1).
// The eps-graph is held as a bit matrix in shared memory (R x R/32 words); components: every core point first takes its
// lowest-indexed core neighbour (one find-first-set per point), chases that pointer to a root, and min-label
// propagation sweeps then only have to confirm the fixed point (or repair the rare component that the forest split).
// Hulls: one WARP per cluster, no sort and no sequential chain.  A point p_i is a vertex of the lower hull iff every
// lexicographically smaller point a and greater point b make a strict left turn a -> p_i -> b.  All directions
// p_i - a (and b - p_i) lie in one half-plane, where the sign of a cross product orders them by angle, so it is enough
// to test the extreme pair: the largest-angle direction from the left against the smallest-angle direction to the
// right (upper hull: mirrored).  Each lane classifies its points with one pass over the cluster (n^2 / 32 cross
// products per warp, all lanes busy) where the previous version ran Andrew's monotone chain in ONE thread after a
// 36-barrier bitonic sort of the whole scan (ncu: 19 of 32 lanes, 38 % of the stall samples on the barrier behind
// the chain).  Vertex order (counter-clockwise from the lexicographic minimum: lower chain ascending, upper chain
// descending) comes from counting, per vertex, the chain vertices before it.  Same strictly convex polygon as the
// chain and as Qhull up to which of several points collinear to 1e-16 is called a vertex.  Exact duplicates
// (np.unique of the reference) are skipped by index.
#include "ldcbf_common.cuh"

namespace ldcbf {

constexpr int CL_RMAX = 512;
constexpr int CL_THREADS = 256;
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_BIG = 0x3fffffff;
constexpr int CL_MAXC = 64;

// Shared-memory layout, sized by RP = R rounded up to 32:
//   double x[RP], y[RP]            compacted valid points, ray order
//   REGION (one buffer, three lives): raw readings sx[RP], sy[RP] (step 1)  ->  eps-graph adj[RP][RP/32] (steps 2-4)
//                                     ->  member lists mem[CL_WARPS][RP], flags flg[RP], hull positions pos[RP] (hulls)
//   unsigned coremask[RP/32]
//   int ray[RP], lab[RP], cid[RP]  ray index; component label (smallest core index), later the cluster number; cluster
//                                  number of a root
//   int hn[64], slot[64], cstart[64], cn[64]   hull size, output slot, member-list start and length per cluster (at most
//                                  64 clusters are hulled)
struct ClusterLayout {
    int RP, W;
    size_t off_y, off_region, off_core, off_ray, off_lab, off_cid, off_hn, bytes;
    __host__ __device__ explicit ClusterLayout(int R) {
        RP = (R + 31) & ~31;
        W = RP / 32;
        size_t o = 0;
        o += sizeof(double) * RP; off_y = o;
        o += sizeof(double) * RP; off_region = o;
        size_t region = sizeof(double) * RP * 2;
        const size_t adj_bytes = sizeof(unsigned) * RP * W, hull_bytes = sizeof(int) * RP * (CL_WARPS + 2);
        region = adj_bytes > region ? adj_bytes : region;
        region = hull_bytes > region ? hull_bytes : region;
        o += region; off_core = o;
        o += sizeof(unsigned) * W; off_ray = o;
        o += sizeof(int) * RP; off_lab = o;
        o += sizeof(int) * RP; off_cid = o;
        o += sizeof(int) * RP; off_hn = o;
        o += sizeof(int) * CL_MAXC * 4;
        bytes = (o + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ bool lex_less(double ax, double ay, double bx, double by) {
    return ax < bx || (ax == bx && ay < by);
}
__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

__global__ void __launch_bounds__(CL_THREADS) lidar_clusters_kernel(int R, const double2* __restrict__ hit_xy,
                                                                    const double2* __restrict__ noise, double eps2,
                                                                    int min_samples, int max_hulls, int max_hull_verts,
                                                                    int32_t* __restrict__ labels,
                                                                    double2* __restrict__ hull_verts,
                                                                    int32_t* __restrict__ hull_nverts,
                                                                    int32_t* __restrict__ n_hulls,
                                                                    int32_t* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char cl_raw[];
    const ClusterLayout Lo(R);
    const int W = Lo.W;
    double* X = reinterpret_cast<double*>(cl_raw);
    double* Y = reinterpret_cast<double*>(cl_raw + Lo.off_y);
    double* SX = reinterpret_cast<double*>(cl_raw + Lo.off_region);
    double* SY = SX + Lo.RP;
    unsigned* ADJ = reinterpret_cast<unsigned*>(cl_raw + Lo.off_region);
    int* MEM = reinterpret_cast<int*>(cl_raw + Lo.off_region);          // [CL_WARPS][RP]
    int* FLG = MEM + CL_WARPS * Lo.RP;
    int* POS = FLG + Lo.RP;
    unsigned* CORE = reinterpret_cast<unsigned*>(cl_raw + Lo.off_core);
    int* RAY = reinterpret_cast<int*>(cl_raw + Lo.off_ray);
    int* LAB = reinterpret_cast<int*>(cl_raw + Lo.off_lab);
    int* CID = reinterpret_cast<int*>(cl_raw + Lo.off_cid);
    int* HN = reinterpret_cast<int*>(cl_raw + Lo.off_hn);
    int* SLOT = HN + CL_MAXC;
    int* CSTART = SLOT + CL_MAXC;
    int* CN = CSTART + CL_MAXC;
    __shared__ int sP, sNC, sChanged, sOut, sOvf;

    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
    const double2* scan = hit_xy + (size_t)b * R;

    // ---- 1. load (+ noise) into the sort buffer, then compact the valid readings in ray order (warp 0)
    for (int i = t; i < R; i += CL_THREADS) {
        double2 p = scan[i];
        if (p.x == p.x && noise) { const double2 nz = noise[(size_t)b * R + i]; p.x += nz.x; p.y += nz.y; }
        SX[i] = p.x; SY[i] = p.y;
        labels[(size_t)b * R + i] = -1;
    }
    if (t < W) CORE[t] = 0u;
    __syncthreads();
    if (t < 32) {
        int base = 0;
        for (int c = 0; c < R; c += 32) {
            const int i = c + lane;
            const double px = i < R ? SX[i] : 0.0, py = i < R ? SY[i] : 0.0;
            const bool ok = i < R && px == px;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) { const int k = base + __popc(m & ((1u << lane) - 1u)); X[k] = px; Y[k] = py; RAY[k] = i; }
            base += __popc(m);
        }
        if (lane == 0) { sP = base; sNC = 0; }
    }
    __syncthreads();
    const int P = sP;

    // ---- 2. eps-graph and core points.  Work items are (word w, point i) pairs with i fastest, so that a scan with a
    // few more than 128 valid readings does not leave 116 threads idle during a second pass over the points, and the
    // 32 partners of a word are read as broadcasts.
    {
        const int Wp = (P + 31) >> 5;
        for (int item = t; item < Wp * P; item += CL_THREADS) {
            const int w = item / P, i = item - w * P;
            const double xi = X[i], yi = Y[i];
            unsigned bits = 0u;
            const int jend = min(32, P - w * 32);
            for (int jj = 0; jj < jend; ++jj) {
                const double dx = X[w * 32 + jj] - xi, dy = Y[w * 32 + jj] - yi;
                if (dx * dx + dy * dy <= eps2) bits |= 1u << jj;
            }
            ADJ[i * W + w] = bits;
        }
        __syncthreads();
        for (int i = t; i < P; i += CL_THREADS) {
            int count = 0;
            for (int w = 0; w < Wp; ++w) count += __popc(ADJ[i * W + w]);
            const bool core = count >= min_samples;
            LAB[i] = core ? i : CL_BIG;
            if (core) atomicOr(&CORE[i >> 5], 1u << (i & 31));
        }
    }
    __syncthreads();

    // ---- 3. connected components of the core points.  First guess: the lowest-indexed core neighbour (self included:
    // one find-first-set instead of a read per neighbour), chased to its root (labels only decrease, so a concurrent
    // update still hands back a valid ancestor); then min-label propagation to the fixed point — usually one confirming
    // sweep, since readings in ray order chain to their predecessor
    for (int i = t; i < P; i += CL_THREADS) {
        if (!((CORE[i >> 5] >> (i & 31)) & 1u)) continue;
        for (int w = 0; w * 32 < P; ++w) {
            const unsigned bits = ADJ[i * W + w] & CORE[w];
            if (bits) { LAB[i] = w * 32 + __ffs(bits) - 1; break; }
        }
    }
    __syncthreads();
    for (int i = t; i < P; i += CL_THREADS) {
        if (!((CORE[i >> 5] >> (i & 31)) & 1u)) continue;
        int m = LAB[i];
        for (int l = LAB[m]; l < m; l = LAB[m]) m = l;
        LAB[i] = m;
    }
    __syncthreads();
    for (;;) {
        if (t == 0) sChanged = 0;
        __syncthreads();
        for (int i = t; i < P; i += CL_THREADS) {
            if (!((CORE[i >> 5] >> (i & 31)) & 1u)) continue;
            int m = LAB[i];
            for (int w = 0; w * 32 < P; ++w) {
                unsigned bits = ADJ[i * W + w] & CORE[w];
                while (bits) {
                    const int j = w * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    m = min(m, LAB[j]);
                }
            }
            // pointer jumping: the label of my label (labels are indices of core points of the same component and only
            // ever decrease, so a stale read is still a valid, if larger, candidate) — a chain of k readings in ray
            // order converges in ~log k sweeps instead of ~k / (neighbours per point)
            if (m != CL_BIG) { const int m2 = LAB[m]; m = min(m, m2); const int m3 = LAB[m]; m = min(m, m3); }
            if (m < LAB[i]) { LAB[i] = m; sChanged = 1; }
        }
        __syncthreads();
        if (!sChanged) break;
        __syncthreads();
    }
    // ---- 4. border points: lowest-numbered cluster among the core neighbours
    for (int i = t; i < P; i += CL_THREADS) {
        if ((CORE[i >> 5] >> (i & 31)) & 1u) continue;
        int m = CL_BIG;
        for (int w = 0; w * 32 < P; ++w) {
            unsigned bits = ADJ[i * W + w] & CORE[w];
            while (bits) {
                const int j = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                m = min(m, LAB[j]);
            }
        }
        LAB[i] = m;
    }
    __syncthreads();
    // ---- 5. number the clusters by their smallest core index (= sklearn's label order)
    if (t < 32) {
        int base = 0;
        for (int c = 0; c < P; c += 32) {
            const int i = c + lane;
            const bool root = i < P && LAB[i] == i;
            const unsigned m = __ballot_sync(0xffffffffu, root);
            if (root) CID[i] = base + __popc(m & ((1u << lane) - 1u));
            base += __popc(m);
        }
        if (lane == 0) sNC = base;
    }
    __syncthreads();
    const int n_clusters = sNC;
    const int nc = min(n_clusters, CL_MAXC);
    // labels out; from here on LAB[i] is the cluster number of point i (CL_BIG = noise)
    for (int i = t; i < P; i += CL_THREADS) {
        const int c = (LAB[i] != CL_BIG) ? CID[LAB[i]] : CL_BIG;
        labels[(size_t)b * R + RAY[i]] = (c == CL_BIG) ? -1 : c;
        LAB[i] = c;
    }
    if (t < CL_MAXC) { HN[t] = 0; SLOT[t] = -1; CN[t] = 0; CSTART[t] = 0; }
    __syncthreads();                      // the eps-graph has been read for the last time: its memory becomes MEM / FLG / POS
    for (int i = t; i < P; i += CL_THREADS) POS[i] = -1;
    __syncthreads();

    // ---- 6. hulls.  (a) one warp per cluster: member list (ray order), lexicographic extremes, flatness rule
    {
        const int warp = t >> 5;
        int* mem = MEM + warp * Lo.RP;
        const unsigned lt = (1u << lane) - 1u;
        int used = 0;                                                      // this warp's lists sit back to back in `mem`
        for (int c = warp; c < nc; c += CL_WARPS) {
            int n = 0;
            for (int base = 0; base < P; base += 32) {
                const int i = base + lane;
                const bool ok = i < P && LAB[i] == c;
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) mem[used + n + __popc(m & lt)] = i;
                n += __popc(m);
            }
            __syncwarp();
            bool keep = false;
            if (n >= 3) {
                // The reference drops a cluster when np.linalg.matrix_rank(points - points[0]) < 2 (`:75-76`) or when
                // Qhull finds the input flat to its roundoff bound and raises (`:81-83`; ~23 eps max|coordinate|).  Both
                // only ever trigger on readings taken on ONE straight edge (deviations ~1e-16); real corners deviate by
                // >= 1e-6.  One test covers both: largest distance from the line through the lexicographic extremes
                // <= 64 eps max|coordinate|.
                double lox = INFINITY, loy = INFINITY, hix = -INFINITY, hiy = -INFINITY, scale = 0.0;
                for (int k = lane; k < n; k += 32) {
                    const double px = X[mem[used + k]], py = Y[mem[used + k]];
                    if (lex_less(px, py, lox, loy)) { lox = px; loy = py; }
                    if (lex_less(hix, hiy, px, py)) { hix = px; hiy = py; }
                    scale = fmax(scale, fmax(fabs(px), fabs(py)));
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ox = __shfl_xor_sync(0xffffffffu, lox, off), oy = __shfl_xor_sync(0xffffffffu, loy, off);
                    if (lex_less(ox, oy, lox, loy)) { lox = ox; loy = oy; }
                    const double qx = __shfl_xor_sync(0xffffffffu, hix, off), qy = __shfl_xor_sync(0xffffffffu, hiy, off);
                    if (lex_less(hix, hiy, qx, qy)) { hix = qx; hiy = qy; }
                    scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, off));
                }
                const double dxl = hix - lox, dyl = hiy - loy;
                double maxcross = 0.0;
                for (int k = lane; k < n; k += 32)
                    maxcross = fmax(maxcross, fabs(cross2(dxl, dyl, X[mem[used + k]] - lox, Y[mem[used + k]] - loy)));
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) maxcross = fmax(maxcross, __shfl_xor_sync(0xffffffffu, maxcross, off));
                const double maxdev = maxcross / sqrt(dxl * dxl + dyl * dyl);      // NaN when all points coincide
                keep = maxdev > 64.0 * 2.220446049250313e-16 * scale;
            }
            if (lane == 0) { CSTART[c] = warp * Lo.RP + used; CN[c] = keep ? n : 0; }
            used += n;
        }
    }
    __syncthreads();
    // (b) one THREAD per point, all clusters at once: classify against the members of the point's own cluster (every
    // partner coordinate is read through the member list; neighbouring threads mostly share the cluster)
    for (int i = t; i < P; i += CL_THREADS) {
        const int c = LAB[i];
        int f = 0;
        if (c < nc && CN[c] > 0) {
            const int n = CN[c];
            const int* mem = MEM + CSTART[c];
            const double xi = X[i], yi = Y[i];
            bool haveL = false, haveR = false, dup = false;
            double dnx = 0.0, dny = 0.0, dxx = 0.0, dxy = 0.0;             // min / max angle of p_i - a, a < p_i
            double enx = 0.0, eny = 0.0, exx = 0.0, exy = 0.0;             // min / max angle of b - p_i, b > p_i
            for (int a = 0; a < n; ++a) {
                const int j = mem[a];
                const double xa = X[j], ya = Y[j];
                if (xa == xi && ya == yi) { dup = dup || j < i; continue; }       // np.unique: the first copy stays
                if (lex_less(xa, ya, xi, yi)) {
                    const double dx = xi - xa, dy = yi - ya;
                    if (!haveL) { dnx = dxx = dx; dny = dxy = dy; haveL = true; }
                    else {
                        if (cross2(dx, dy, dnx, dny) > 0.0) { dnx = dx; dny = dy; }
                        if (cross2(dxx, dxy, dx, dy) > 0.0) { dxx = dx; dxy = dy; }
                    }
                } else {
                    const double ex = xa - xi, ey = ya - yi;
                    if (!haveR) { enx = exx = ex; eny = exy = ey; haveR = true; }
                    else {
                        if (cross2(ex, ey, enx, eny) > 0.0) { enx = ex; eny = ey; }
                        if (cross2(exx, exy, ex, ey) > 0.0) { exx = ex; exy = ey; }
                    }
                }
            }
            const bool ext = !haveL || !haveR;
            const bool lo = !dup && (ext || cross2(dxx, dxy, enx, eny) > 0.0);
            const bool up = !dup && !lo && cross2(dnx, dny, exx, exy) < 0.0;
            f = (lo ? 1 : 0) | (up ? 2 : 0);
        }
        FLG[i] = f;
    }
    __syncthreads();
    // (c) vertex order: counter-clockwise from the lexicographic minimum = lower chain ascending, then upper chain
    // descending; every vertex counts the vertices of its chain that come before it
    for (int i = t; i < P; i += CL_THREADS) {
        const int f = FLG[i];
        if (!f) continue;
        const int c = LAB[i], n = CN[c];
        const int* mem = MEM + CSTART[c];
        const double xi = X[i], yi = Y[i];
        int before = 0, n_lo = 0, n_up = 0;
        for (int a = 0; a < n; ++a) {
            const int j = mem[a], fj = FLG[j];
            n_lo += fj & 1;
            n_up += (fj >> 1) & 1;
            if (fj == f) before += (f & 1) ? lex_less(X[j], Y[j], xi, yi) : lex_less(xi, yi, X[j], Y[j]);
        }
        if (n_lo + n_up >= 3) {
            POS[i] = (f & 1) ? before : n_lo + before;
            if ((f & 1) && before == 0) HN[c] = n_lo + n_up;               // the lexicographic minimum reports the size
        }
    }
    __syncthreads();

    // ---- 7. output slots in cluster order (clusters without a hull are skipped), then every vertex writes itself
    if (t < 32) {
        int base = 0;
        bool ovf = n_clusters > CL_MAXC;
        for (int c0 = 0; c0 < nc; c0 += 32) {
            const int c = c0 + lane;
            const int h = c < nc ? HN[c] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, h >= 3);
            if (h >= 3) {
                const int o = base + __popc(m & ((1u << lane) - 1u));
                if (o < max_hulls) {
                    SLOT[c] = o;
                    hull_nverts[(size_t)b * max_hulls + o] = min(h, max_hull_verts);
                }
                ovf = ovf || o >= max_hulls || h > max_hull_verts;
            }
            base += __popc(m);
        }
        ovf = __any_sync(0xffffffffu, ovf);
        if (lane == 0) { sOut = min(base, max_hulls); sOvf = ovf ? 1 : 0; }
    }
    __syncthreads();
    const int n_out = sOut;
    for (int i = t; i < P; i += CL_THREADS) {
        const int c = LAB[i], q = POS[i];
        if (c < nc && q >= 0 && q < max_hull_verts && SLOT[c] >= 0)
            hull_verts[((size_t)b * max_hulls + SLOT[c]) * max_hull_verts + q] = make_double2(X[i], Y[i]);
    }
    for (int c = 0; c < nc; ++c) {                                        // zero padding behind every written hull
        const int o = SLOT[c];
        if (o < 0) continue;
        double2* dst = hull_verts + ((size_t)b * max_hulls + o) * max_hull_verts;
        for (int i = min(HN[c], max_hull_verts) + t; i < max_hull_verts; i += CL_THREADS) dst[i] = make_double2(0.0, 0.0);
    }
    for (int o = n_out + t; o < max_hulls; o += CL_THREADS) hull_nverts[(size_t)b * max_hulls + o] = 0;
    if (t == 0) { n_hulls[b] = n_out; if (overflow) overflow[b] = sOvf; }
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
    const size_t smem = ClusterLayout(R).bytes + 64;
    cudaError_t e = cudaFuncSetAttribute(lidar_clusters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    lidar_clusters_kernel<<<B, CL_THREADS, smem, static_cast<cudaStream_t>(cuda_stream)>>>(
        R, reinterpret_cast<const double2*>(hit_xy), reinterpret_cast<const double2*>(noise), eps * eps, min_samples,
        max_hulls, max_hull_verts, labels, reinterpret_cast<double2*>(hull_verts), hull_nverts, n_hulls, overflow);
    return check_launch();
}
