// K1 — LDCBF half-plane builder (closest point c on each obstacle and unit normal eta).
//
// Replaces, for a whole batch at once, the reference's per-obstacle Python loop
//   Utils/ObstaclesUtils.py:60-109  get_closest_point_and_normal_vector_from_obs
//   Utils/ObstaclesUtils.py:50-57   is_point_inside_polygon (matplotlib Path.contains_point)
//   MPC/HumanoidMpc.py:296-319      _get_list_c_and_eta
//
// Mapping: a group of G lanes (G = 8, 16 or 32, the smallest power of two >= max_verts, capped at 32)
// owns one (scenario, obstacle) pair; lane l evaluates edges l, l+G, ... of the CCW vertex ring, so the
// 16-byte vertex loads of a group are one contiguous run (coalesced, read once: 16*V bytes per
// obstacle).  The first strict minimum over edges (`dist < min_dist`, ObstaclesUtils.py:94) is kept by an
// xor-butterfly argmin that breaks ties towards the lower edge index; the inside test is the parity of
// a ballot.  HBM-bound: 16*E bytes in, 32 bytes out per obstacle, ~40 flop per edge.
//
// Arithmetic follows the oracle (oracle/halfplane.py) operation by operation — explicit round-to-nearest
// intrinsics so that nvcc cannot contract a*b+c into an FMA where numpy does not, and an explicit FMA
// where numpy's 2-element dot does (oracle/model.py: dot2).  This file is compiled with -fmad=false.
#include "halfplane_dev.cuh"

namespace ldcbf {

template <int G>
__global__ void __launch_bounds__(256) halfplane_kernel(int n_pairs, int max_obs, int max_verts,
                                                        const double* __restrict__ pos, int pos_stride, int y_off,
                                                        const double2* __restrict__ verts,
                                                        const int32_t* __restrict__ nverts,
                                                        const int32_t* __restrict__ nobs,
                                                        double4* __restrict__ c_eta) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = tid / G;
    const int lane = tid % G;
    if (pair >= n_pairs) return;   // whole groups exit together (blockDim is a multiple of G)
    const int b = pair / max_obs;
    const int o = pair - b * max_obs;
    // lanes of this group inside the warp
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));

    const int V = (o < nobs[b]) ? nverts[pair] : 0;
    if (V <= 0) {
        if (lane == 0) c_eta[pair] = make_double4(0.0, 0.0, 0.0, 0.0);
        return;
    }
    const double px = pos[(size_t)b * pos_stride], py = pos[(size_t)b * pos_stride + y_off];
    const double2* ring = verts + (size_t)pair * max_verts;

    double best_d = INFINITY, best_cx = 0.0, best_cy = 0.0;
    int best_e = 0x7fffffff;
    int crossings = 0;
    for (int e = lane; e < V; e += G) {
        const double2 A = __ldg(ring + e);
        const double2 Bv = __ldg(ring + ((e + 1 == V) ? 0 : e + 1));
        double cx, cy;
        const double d = edge_closest(px, py, A, Bv, cx, cy, crossings);
        if (d < best_d) { best_d = d; best_cx = cx; best_cy = cy; best_e = e; }
    }
    // first strict minimum across the group: smaller distance wins, ties go to the lower edge index
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(gmask, best_d, off, G);
        const double ocx = __shfl_xor_sync(gmask, best_cx, off, G);
        const double ocy = __shfl_xor_sync(gmask, best_cy, off, G);
        const int oe = __shfl_xor_sync(gmask, best_e, off, G);
        crossings += __shfl_xor_sync(gmask, crossings, off, G);
        if (od < best_d || (od == best_d && oe < best_e)) { best_d = od; best_cx = ocx; best_cy = ocy; best_e = oe; }
    }
    if (lane == 0) {
        c_eta[pair] = finish_halfplane(px, py, best_cx, best_cy, crossings);
    }
}

int launch_halfplanes(int B, int max_obs, int max_verts, const double* pos, int pos_stride, int y_off,
                      const double* verts, const int32_t* nverts, const int32_t* nobs, double* c_eta,
                      void* cuda_stream) {
    if (B < 0 || max_obs <= 0 || max_verts <= 0) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!pos || !verts || !nverts || !nobs || !c_eta) return LDCBF_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int n_pairs = B * max_obs;
    const int threads = 256;
    auto grid = [&](int G) { return (unsigned)(((size_t)n_pairs * G + threads - 1) / threads); };
    const double2* v2 = reinterpret_cast<const double2*>(verts);
    double4* ce = reinterpret_cast<double4*>(c_eta);
    if (max_verts <= 8)
        halfplane_kernel<8><<<grid(8), threads, 0, st>>>(n_pairs, max_obs, max_verts, pos, pos_stride, y_off, v2, nverts, nobs, ce);
    else if (max_verts <= 16)
        halfplane_kernel<16><<<grid(16), threads, 0, st>>>(n_pairs, max_obs, max_verts, pos, pos_stride, y_off, v2, nverts, nobs, ce);
    else
        halfplane_kernel<32><<<grid(32), threads, 0, st>>>(n_pairs, max_obs, max_verts, pos, pos_stride, y_off, v2, nverts, nobs, ce);
    return check_launch();
}

}  // namespace ldcbf

extern "C" int ldcbf_halfplanes_f64(int B, int max_obs, int max_verts, const double* pos, const double* verts,
                                    const int32_t* nverts, const int32_t* nobs, double* c_eta, void* cuda_stream) {
    return ldcbf::launch_halfplanes(B, max_obs, max_verts, pos, 2, 1, verts, nverts, nobs, c_eta, cuda_stream);
}
