// f3 (map front-end of the sub-goal planner): occupancy grid of the convex obstacles, exact Euclidean distance
// transform and clearance cost, batched over scenarios.
//
// Reference semantics restated (HumanoidNavigation/MPC/HumanoidMPCVariants/HumanoidMPCWithRRT.py):
//   :31-52   frame = bounding box of hull vertices, origin and goal, padded by 3; height = ceil(width * dy / dx)
//   :55-58   world -> grid: np.round (half to even) of ((x - min) / (max - min)) * size, same operation order
//   :66-86   per obstacle: integer cells of [xmin, xmax) x [ymin, ymax) inside the Delaunay triangulation of the
//            rounded hull vertices = inside or on their convex hull; decided here with exact integer cross products
//   :103-108 scipy.ndimage.distance_transform_edt(1 - og) and exp(-d)
// Kernel A (one block per scenario): frame, rasterisation, first EDT pass (distance along x to the nearest occupied
// cell of the same grid column y, coalesced over y).  Kernel B (one block per grid row x): second pass
// d^2[x][y] = min_y' (y - y')^2 + g[x][y']^2 from shared memory, then sqrt (exact: integer argument) and exp.
#include <limits.h>

#include "ldcbf_common.cuh"

namespace ldcbf {

constexpr int EDT_INF = 1 << 14;        // larger than any grid extent; its square still fits an int32

__device__ __forceinline__ int grid_coord(double v, double lo, double hi, double size) {
    return (int)rint(__dmul_rn(__ddiv_rn(__dsub_rn(v, lo), __dsub_rn(hi, lo)), size));
}

__global__ void __launch_bounds__(256) occupancy_kernel(int B, int width, int h_cap, int max_obs, int max_verts,
                                                        const double* __restrict__ goal,
                                                        const double* __restrict__ verts,
                                                        const int32_t* __restrict__ nverts,
                                                        const int32_t* __restrict__ nobs, double* __restrict__ meta,
                                                        uint8_t* __restrict__ og, int32_t* __restrict__ work) {
    extern __shared__ int sh_int[];                 // vx[max_verts], vy[max_verts], hx[2*max_verts+2], hy[...]
    __shared__ double red[4][8];
    __shared__ double frame[4];
    __shared__ int s_h, s_hn, s_box[4];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int no = min(nobs[b], max_obs);
    const double* V = verts + (size_t)b * max_obs * max_verts * 2;
    const int32_t* NV = nverts + (size_t)b * max_obs;
    // ---- bounding box of all hull vertices
    double lox = INFINITY, loy = INFINITY, hix = -INFINITY, hiy = -INFINITY;
    for (int e = tid; e < no * max_verts; e += 256) {
        const int o = e / max_verts, i = e - o * max_verts;
        if (i < NV[o]) {
            const double2 p = reinterpret_cast<const double2*>(V)[e];
            lox = fmin(lox, p.x); hix = fmax(hix, p.x); loy = fmin(loy, p.y); hiy = fmax(hiy, p.y);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lox = fmin(lox, __shfl_xor_sync(0xffffffffu, lox, s)); loy = fmin(loy, __shfl_xor_sync(0xffffffffu, loy, s));
        hix = fmax(hix, __shfl_xor_sync(0xffffffffu, hix, s)); hiy = fmax(hiy, __shfl_xor_sync(0xffffffffu, hiy, s));
    }
    if (lane == 0) { red[0][warp] = lox; red[1][warp] = loy; red[2][warp] = hix; red[3][warp] = hiy; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            lox = fmin(lox, red[0][w]); loy = fmin(loy, red[1][w]); hix = fmax(hix, red[2][w]); hiy = fmax(hiy, red[3][w]);
        }
        const double gx = goal[2 * (size_t)b], gy = goal[2 * (size_t)b + 1];
        const double min_x = fmin(fmin(0.0, gx), lox) - 3.0, min_y = fmin(fmin(0.0, gy), loy) - 3.0;
        const double max_x = fmax(fmax(0.0, gx), hix) + 3.0, max_y = fmax(fmax(0.0, gy), hiy) + 3.0;
        const double hh = ceil(__dmul_rn((double)width, __ddiv_rn(__dsub_rn(max_y, min_y), __dsub_rn(max_x, min_x))));
        const bool ok = no > 0 && hh >= 1.0 && hh <= (double)h_cap;
        frame[0] = min_x; frame[1] = min_y; frame[2] = max_x; frame[3] = max_y;
        s_h = ok ? (int)hh : -1;
        double* m = meta + 6 * (size_t)b;
        m[0] = min_x; m[1] = min_y; m[2] = max_x; m[3] = max_y; m[4] = hh; m[5] = ok ? 0.0 : -1.0;
    }
    __syncthreads();
    const int H = s_h;
    if (H < 0) return;                               // grid does not fit h_cap (or no obstacle): flagged in meta[5]
    const int ldy = h_cap + 1;
    uint8_t* G = og + (size_t)b * (width + 1) * ldy;
    int32_t* Wk = work + (size_t)b * (width + 1) * ldy;
    for (int i = tid; i < (width + 1) * ldy; i += 256) G[i] = 0;
    int* vx = sh_int;
    int* vy = vx + max_verts;
    int* hx = vy + max_verts;
    int* hy = hx + 2 * max_verts + 2;
    int occupied = 0;
    for (int o = 0; o < no; ++o) {
        __syncthreads();
        const int nv = min(NV[o], max_verts);
        for (int i = tid; i < nv; i += 256) {
            const double2 p = reinterpret_cast<const double2*>(V)[o * max_verts + i];
            vx[i] = grid_coord(p.x, frame[0], frame[2], (double)width);
            vy[i] = grid_coord(p.y, frame[1], frame[3], (double)H);
        }
        __syncthreads();
        if (tid == 0) {
            // insertion sort by (x, y), then Andrew's monotone chain (collinear points dropped)
            int x0 = INT_MAX, x1 = INT_MIN, y0 = INT_MAX, y1 = INT_MIN;
            for (int i = 0; i < nv; ++i) {
                x0 = min(x0, vx[i]); x1 = max(x1, vx[i]); y0 = min(y0, vy[i]); y1 = max(y1, vy[i]);
                const int px = vx[i], py = vy[i];
                int j = i - 1;
                while (j >= 0 && (vx[j] > px || (vx[j] == px && vy[j] > py))) { vx[j + 1] = vx[j]; vy[j + 1] = vy[j]; --j; }
                vx[j + 1] = px; vy[j + 1] = py;
            }
            int k = 0;
            auto cross = [&](int a, int c, int px, int py) {
                return (long long)(hx[c] - hx[a]) * (py - hy[a]) - (long long)(hy[c] - hy[a]) * (px - hx[a]);
            };
            for (int i = 0; i < nv; ++i) {
                if (i > 0 && vx[i] == vx[i - 1] && vy[i] == vy[i - 1]) continue;
                while (k >= 2 && cross(k - 2, k - 1, vx[i], vy[i]) <= 0) --k;
                hx[k] = vx[i]; hy[k] = vy[i]; ++k;
            }
            const int lower = k + 1;
            for (int i = nv - 2; i >= 0; --i) {
                if (vx[i] == vx[i + 1] && vy[i] == vy[i + 1]) continue;
                while (k >= lower && cross(k - 2, k - 1, vx[i], vy[i]) <= 0) --k;
                hx[k] = vx[i]; hy[k] = vy[i]; ++k;
            }
            s_hn = k - 1;                            // last point repeats the first
            s_box[0] = x0; s_box[1] = x1; s_box[2] = y0; s_box[3] = y1;
        }
        __syncthreads();
        const int hn = s_hn;
        if (hn < 3) continue;                        // degenerate after rounding (the reference's Delaunay raises)
        const int x0 = s_box[0], nx = s_box[1] - s_box[0], y0 = s_box[2], ny = s_box[3] - s_box[2];
        for (int c = tid; c < nx * ny; c += 256) {
            const int px = x0 + c / ny, py = y0 + c % ny;
            bool in = px >= 0 && px <= width && py >= 0 && py <= H;
            for (int e = 0; e < hn && in; ++e) {
                const long long cr = (long long)(hx[e + 1] - hx[e]) * (py - hy[e]) - (long long)(hy[e + 1] - hy[e]) * (px - hx[e]);
                in = cr >= 0;
            }
            if (in) { G[(size_t)px * ldy + py] = 1; }
        }
    }
    __syncthreads();
    // ---- first EDT pass: for every grid column y, distance along x to the nearest occupied cell
    for (int y = tid; y <= H; y += 256) {
        int g = EDT_INF;
        for (int x = 0; x <= width; ++x) {
            const bool occ = G[(size_t)x * ldy + y] != 0;
            occupied += occ;
            g = occ ? 0 : min(g + 1, EDT_INF);
            Wk[(size_t)x * ldy + y] = g;
        }
        g = EDT_INF;
        for (int x = width; x >= 0; --x) {
            const int f = Wk[(size_t)x * ldy + y];
            g = f == 0 ? 0 : min(g + 1, EDT_INF);
            Wk[(size_t)x * ldy + y] = min(f, g);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) occupied += __shfl_xor_sync(0xffffffffu, occupied, s);
    if (lane == 0 && occupied) atomicAdd(&meta[6 * (size_t)b + 5], (double)occupied);
}

__global__ void __launch_bounds__(256) clearance_kernel(int width, int h_cap, const double* __restrict__ meta,
                                                        const int32_t* __restrict__ work, double* __restrict__ dist,
                                                        double* __restrict__ cost) {
    extern __shared__ int g2[];                     // squared first-pass distances of this grid row
    const int x = blockIdx.x, b = blockIdx.y;
    const double* m = meta + 6 * (size_t)b;
    if (m[5] < 0.0) return;
    const int H = (int)m[4], ldy = h_cap + 1;
    const size_t row = ((size_t)b * (width + 1) + x) * ldy;
    for (int y = threadIdx.x; y <= H; y += 256) { const int g = work[row + y]; g2[y] = g * g; }
    __syncthreads();
    for (int y = threadIdx.x; y <= H; y += 256) {
        int best = INT_MAX;
        for (int yy = 0; yy <= H; ++yy) {
            const int dy = y - yy;
            best = min(best, dy * dy + g2[yy]);
        }
        const double d = sqrt((double)best);
        dist[row + y] = d;
        if (cost) cost[row + y] = exp(-d);
    }
}

}  // namespace ldcbf

using namespace ldcbf;

extern "C" int ldcbf_clearance_grid_f64(int B, int width, int h_cap, int max_obs, int max_verts, const double* goal,
                                        const double* verts, const int32_t* nverts, const int32_t* nobs, double* meta,
                                        uint8_t* og, double* dist, double* cost, int32_t* work, void* cuda_stream) {
    if (B < 0 || width <= 0 || width >= EDT_INF || h_cap <= 0 || h_cap >= EDT_INF || max_obs <= 0 || max_verts <= 0)
        return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!goal || !verts || !nverts || !nobs || !meta || !og || !dist || !work) return LDCBF_E_ARG;
    if (B > 65535) return LDCBF_E_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t sh_a = sizeof(int) * (size_t)(6 * max_verts + 4);
    if (sh_a > 48 * 1024) return LDCBF_E_SHAPE;
    occupancy_kernel<<<B, 256, sh_a, st>>>(B, width, h_cap, max_obs, max_verts, goal, verts, nverts, nobs, meta, og, work);
    int rc = check_launch();
    if (rc != LDCBF_OK) return rc;
    const size_t sh_b = sizeof(int) * (size_t)(h_cap + 1);
    if (sh_b > 48 * 1024) return LDCBF_E_SHAPE;
    clearance_kernel<<<dim3(width + 1, B), 256, sh_b, st>>>(width, h_cap, meta, work, dist, cost);
    return check_launch();
}
