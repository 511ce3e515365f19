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
// One 256-thread CTA per scan (R <= 512 rays; measured 128 -> 256 threads: 1.83 -> 1.72 ms for 16384 scans).  DBSCAN on <= 512 points is done exactly as sklearn defines it:
//   * neighbourhood = points within eps (squared distances, self included); core = at least min_samples neighbours;
//   * clusters = connected components of core points, numbered by their smallest core-point index (sklearn visits
//     points in index order and opens a cluster at the first unvisited core point);
//   * a border point takes the lowest-numbered cluster among its core neighbours (it is labelled by the first
//     cluster that reaches it and never relabelled); everything else is noise (-1).
// The eps-graph is held as a bit matrix in shared memory (R x R/32 words); components: every core point first takes its
// lowest-indexed core neighbour (one find-first-set), chases that pointer to a root, and min-label propagation sweeps
// then confirm the fixed point.
// Hulls: ONE bitonic sort of all clustered points by (cluster, x, y) turns every cluster into a contiguous segment in
// np.unique order; then one WARP per cluster removes duplicates (ballot compaction), tests flatness (warp reduction) and
// builds both chains of Andrew's monotone-chain hull by parallel elimination (step 7), all clusters concurrently.  Qhull
// and the chain may disagree on which of several points that are collinear to 1e-16 is called a vertex; the polygons are
// the same to rounding.
#include "ldcbf_common.cuh"

namespace ldcbf {

constexpr int CL_RMAX = 512;
constexpr int CL_THREADS = 256;
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_BIG = 0x3fffffff;

// Shared-memory layout, sized by RP = R rounded up to 32 (and the sort arrays to the next power of two NS):
//   double x[RP], y[RP]            compacted valid points, ray order
//   double sx[NS], sy[NS]          all clustered points sorted by (cluster, x, y)
//   int    sc[NS]                  cluster number of the sorted points (CL_BIG = noise / padding)
//   unsigned adj[RP][RP/32]        eps-graph
//   unsigned coremask[RP/32]
//   int ray[RP], lab[RP], cid[RP]  ray index; component label (smallest core index); cluster number of a root
//   int hull[NS + 64]              hull stacks, one region per cluster
//   int seg0[64], segn[64], hn[64] segment start / length / hull size per cluster (at most 64 clusters are hulled)
struct ClusterLayout {
    int RP, W, NS;
    size_t off_y, off_sx, off_sy, off_sc, off_adj, off_core, off_ray, off_lab, off_cid, off_hull, off_seg, bytes;
    __host__ __device__ explicit ClusterLayout(int R) {
        RP = (R + 31) & ~31;
        W = RP / 32;
        NS = 32;
        while (NS < RP) NS <<= 1;
        // The sort buffers (sx, sy, sc) and the eps-graph share one region: sx / sy stage the raw readings in step 1,
        // before the graph is built (step 2), and are not written again until the graph has been read for the last
        // time (border points, step 4).  10 KB less per block at R = 360 -> one more resident block per SM.
        size_t o = 0;
        o += sizeof(double) * RP; off_y = o;
        o += sizeof(double) * RP; off_sx = o;
        const size_t sort_bytes = sizeof(double) * NS * 2 + sizeof(int) * NS;
        const size_t adj_bytes = sizeof(unsigned) * RP * W;
        off_sy = off_sx + sizeof(double) * NS;
        off_sc = off_sy + sizeof(double) * NS;
        off_adj = off_sx;
        o += sort_bytes > adj_bytes ? sort_bytes : adj_bytes; off_core = o;
        o += sizeof(unsigned) * W; off_ray = o;
        o += sizeof(int) * RP; off_lab = o;
        o += sizeof(int) * RP; off_cid = o;
        o += sizeof(int) * RP; off_hull = o;
        o += sizeof(int) * (NS + 64); off_seg = o;
        o += sizeof(int) * 64 * 3;
        bytes = (o + 15) & ~(size_t)15;
    }
};

constexpr int CL_MAXC = 64;

__device__ __forceinline__ bool key_less(int ac, double ax, double ay, int bc, double bx, double by) {
    return ac < bc || (ac == bc && (ax < bx || (ax == bx && ay < by)));
}

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
    double* SX = reinterpret_cast<double*>(cl_raw + Lo.off_sx);
    double* SY = reinterpret_cast<double*>(cl_raw + Lo.off_sy);
    int* SC = reinterpret_cast<int*>(cl_raw + Lo.off_sc);
    unsigned* ADJ = reinterpret_cast<unsigned*>(cl_raw + Lo.off_adj);
    unsigned* CORE = reinterpret_cast<unsigned*>(cl_raw + Lo.off_core);
    int* RAY = reinterpret_cast<int*>(cl_raw + Lo.off_ray);
    int* LAB = reinterpret_cast<int*>(cl_raw + Lo.off_lab);
    int* CID = reinterpret_cast<int*>(cl_raw + Lo.off_cid);
    int* HULL = reinterpret_cast<int*>(cl_raw + Lo.off_hull);
    int* SEG0 = reinterpret_cast<int*>(cl_raw + Lo.off_seg);
    int* SEGN = SEG0 + CL_MAXC;
    int* HN = SEGN + CL_MAXC;
    __shared__ int sP, sNC, sChanged;

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
    int ns = 32;                                   // sort size: next power of two >= P
    while (ns < P) ns <<= 1;
    // labels out; sort keys (cluster, x, y) for every point, noise and padding last
    for (int i = t; i < ns; i += CL_THREADS) {
        int c = CL_BIG;
        double px = INFINITY, py = INFINITY;
        if (i < P) {
            if (LAB[i] != CL_BIG) c = CID[LAB[i]];
            labels[(size_t)b * R + RAY[i]] = (c == CL_BIG) ? -1 : c;
            px = X[i]; py = Y[i];
        }
        SC[i] = c; SX[i] = px; SY[i] = py;
    }
    if (t < CL_MAXC) { SEG0[t] = 0; SEGN[t] = 0; HN[t] = 0; }
    __syncthreads();

    // ---- 6. ONE bitonic sort by (cluster, x, y): every cluster becomes a contiguous, np.unique-ordered segment.
    // Work items are the ns / 2 compare-exchange pairs of a stage; pair q of a stage with partner distance j touches
    // i = 2j (q / j) + q % j and i + j.  For j <= 32 both lie in the 64-element chunk q / 32, which belongs to one warp:
    // those stages only need a warp barrier (3 block barriers instead of 36 at ns = 256).
    for (int k = 2; k <= ns; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = t; q < (ns >> 1); q += CL_THREADS) {
                const int i = 2 * j * (q / j) + (q % j), l = i + j;
                const bool up = (i & k) == 0;
                const int ac = SC[i], bc = SC[l];
                const double ax = SX[i], ay = SY[i], bx = SX[l], by = SY[l];
                if (key_less(bc, bx, by, ac, ax, ay) == up) {
                    SC[i] = bc; SX[i] = bx; SY[i] = by; SC[l] = ac; SX[l] = ax; SY[l] = ay;
                }
            }
            // the NEXT stage has distance j / 2 (or k for the first stage of the next k): a block barrier is needed
            // whenever this stage or the next one crosses 64-element chunks
            const int jn = (j > 1) ? (j >> 1) : k;
            if (j > 32 || jn > 32) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
    // segment boundaries
    for (int i = t; i < ns; i += CL_THREADS) {
        const int c = SC[i];
        if (c < CL_MAXC) {
            if (i == 0 || SC[i - 1] != c) SEG0[c] = i;
            if (i == ns - 1 || SC[i + 1] != c) SEGN[c] = i + 1;        // end (exclusive)
        }
    }
    __syncthreads();

    // ---- 7. one WARP per cluster: duplicates out, flatness test, both hull chains by parallel elimination
    // Round 1 ran Andrew's monotone chain in ONE thread per cluster (ncu round 2, source view: 23 k of the 56 k warp
    // instructions of a scan at 1.1 active lanes, half of the kernel's stall samples on the barrier behind it).  The same
    // chains, all lanes busy: in a list sorted along x, an interior point that does not make a strict left turn with its
    // two CURRENT neighbours (the test of the chain: cross(b - a, q - a) <= 0) lies on or below a chord between two
    // points of the set, so it is not a vertex whatever happens to its neighbours in the same round; every lane tests one
    // point, the survivors are compacted with a ballot, and the rounds stop when nothing was removed — the survivors then
    // form a strictly convex chain that contains every vertex: the hull.  A noisy arc halves per round.  Same polygon as
    // the sequential chain up to which of several points collinear to 1e-16 is kept (different triples are tested).
    const int nc = min(n_clusters, CL_MAXC);
    {
        const int warp = t >> 5;
        const unsigned lt = (1u << lane) - 1u;
        for (int cl = warp; cl < nc; cl += CL_WARPS) {
            const int s0 = SEG0[cl], cnt = SEGN[cl] - s0;
            const double* sx = SX + s0;
            const double* sy = SY + s0;
            int* uniq = LAB + s0;                 // the per-point label arrays are free after step 5 (sorted positions < P)
            int* hull = HULL + s0 + cl;           // cnt + 1 entries; regions of different clusters are disjoint
            int* ping = SC + s0;                  // the cluster keys of this segment are all `cl`: free after the boundaries
            int* pong = CID + s0;
            // np.unique: a point equal to its predecessor in the sorted segment is dropped
            int u = 0;
            for (int base = 0; base < cnt; base += 32) {
                const int k = base + lane;
                const bool keep = k < cnt && (k == 0 || sx[k] != sx[k - 1] || sy[k] != sy[k - 1]);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) uniq[u + __popc(m & lt)] = k;
                u += __popc(m);
            }
            __syncwarp();
            int h = 0;
            if (u >= 3) {
                // The reference drops a cluster when np.linalg.matrix_rank(points - points[0]) < 2 (`:75-76`) or when
                // Qhull finds the input flat to its roundoff bound and raises (`:81-83`; ~23 eps max|coordinate|).  Both
                // only ever trigger on readings taken on ONE straight edge (deviations ~1e-16); real corners deviate
                // by >= 1e-6.  One test covers both: largest distance from the line through the lexicographic
                // extremes <= 64 eps max|coordinate|.
                const double x0 = sx[uniq[0]], y0 = sy[uniq[0]];
                const double dxl = sx[uniq[u - 1]] - x0, dyl = sy[uniq[u - 1]] - y0;
                double maxcross = 0.0, scale = 0.0;
                for (int k = lane; k < u; k += 32) {
                    const double px = sx[uniq[k]], py = sy[uniq[k]];
                    maxcross = fmax(maxcross, fabs(dxl * (py - y0) - dyl * (px - x0)));
                    scale = fmax(scale, fmax(fabs(px), fabs(py)));
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    maxcross = fmax(maxcross, __shfl_xor_sync(0xffffffffu, maxcross, off));
                    scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, off));
                }
                const double maxdev = maxcross / sqrt(dxl * dxl + dyl * dyl);
                if (maxdev > 64.0 * 2.220446049250313e-16 * scale) {
                    // one elimination round: src (length m, read through `rev` for the right-to-left chain) -> dst
                    auto round = [&](const int* src, bool rev, int m, int* dst, bool& changed) {
                        int out = 0;
                        changed = false;
                        for (int base = 0; base < m; base += 32) {
                            const int j = base + lane;
                            bool keep = false;
                            int id = 0;
                            if (j < m) {
                                id = src[rev ? m - 1 - j : j];
                                keep = j == 0 || j == m - 1;
                                if (!keep) {
                                    const int ia = src[rev ? m - j : j - 1], iq = src[rev ? m - 2 - j : j + 1];
                                    const double ax = sx[ia], ay = sy[ia];
                                    keep = (sx[id] - ax) * (sy[iq] - ay) - (sy[id] - ay) * (sx[iq] - ax) > 0.0;
                                }
                            }
                            const unsigned mk = __ballot_sync(0xffffffffu, keep);
                            if (keep) dst[out + __popc(mk & lt)] = id;
                            out += __popc(mk);
                            changed = changed || mk != __ballot_sync(0xffffffffu, j < m);
                        }
                        __syncwarp();
                        return out;
                    };
                    bool changed;
                    // lower chain, left to right: uniq -> hull -> ping -> hull ...; ends in `hull`
                    int m = round(uniq, false, u, hull, changed);
                    while (changed) {
                        m = round(hull, false, m, ping, changed);
                        for (int k = lane; k < m; k += 32) hull[k] = ping[k];
                        __syncwarp();
                    }
                    const int n_lo = m;
                    // upper chain, right to left: uniq reversed -> ping -> pong -> ping ...; ends in `ping`
                    m = round(uniq, true, u, ping, changed);
                    while (changed) {
                        m = round(ping, false, m, pong, changed);
                        for (int k = lane; k < m; k += 32) ping[k] = pong[k];
                        __syncwarp();
                    }
                    // counter-clockwise: the lower chain, then the upper chain without its two end points
                    for (int k = 1 + lane; k < m - 1; k += 32) hull[n_lo + k - 1] = ping[k];
                    h = n_lo + m - 2;
                    if (h < 3) h = 0;
                    __syncwarp();
                }
            }
            if (lane == 0) HN[cl] = h;
        }
    }
    __syncthreads();

    // ---- 8. write the hulls in cluster order (clusters without a hull are skipped)
    int n_out = 0;
    bool ovf = n_clusters > CL_MAXC;
    for (int c = 0; c < nc; ++c) {
        const int h = HN[c];
        if (h < 3) continue;
        if (n_out < max_hulls) {
            const int hv = min(h, max_hull_verts);
            if (h > max_hull_verts) ovf = true;
            const int s0 = SEG0[c];
            const int* hull = HULL + s0 + c;
            double2* dst = hull_verts + ((size_t)b * max_hulls + n_out) * max_hull_verts;
            for (int i = t; i < max_hull_verts; i += CL_THREADS)
                dst[i] = (i < hv) ? make_double2(SX[s0 + hull[i]], SY[s0 + hull[i]]) : make_double2(0.0, 0.0);
            if (t == 0) hull_nverts[(size_t)b * max_hulls + n_out] = hv;
            ++n_out;
        } else {
            ovf = true;
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
    const size_t smem = ClusterLayout(R).bytes + 64;
    cudaError_t e = cudaFuncSetAttribute(lidar_clusters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    lidar_clusters_kernel<<<B, CL_THREADS, smem, static_cast<cudaStream_t>(cuda_stream)>>>(
        R, reinterpret_cast<const double2*>(hit_xy), reinterpret_cast<const double2*>(noise), eps * eps, min_samples,
        max_hulls, max_hull_verts, labels, reinterpret_cast<double2*>(hull_verts), hull_nverts, n_hulls, overflow);
    return check_launch();
}
