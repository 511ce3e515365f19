// K1 — LDCBF half-plane builder (closest point c on each obstacle and unit normal eta).
//
// Replaces, for a whole batch at once, the reference's per-obstacle Python loop
//   Utils/ObstaclesUtils.py:60-109  get_closest_point_and_normal_vector_from_obs
//   Utils/ObstaclesUtils.py:50-57   is_point_inside_polygon (matplotlib Path.contains_point)
//   MPC/HumanoidMpc.py:296-319      _get_list_c_and_eta
//
// Mapping: one thread per (scenario, obstacle) pair walks its own ring of double2 vertices straight from global
// memory (16-byte __ldg loads; a 128-byte line serves 8 consecutive edges of the same thread and stays in L1 in
// between).  No shuffles, no idle lanes inside a ring, DRAM traffic = the valid vertices only.
// An earlier version staged the rings of a block in shared memory with coalesced cp.async copies; it was 1.5x
// slower (0.58 vs 0.39 ms at B = 2^20): the 400 B of shared memory per thread capped the SM at 16 warps, and this
// kernel needs warps more than it needs coalescing — the bit-exact arithmetic costs ~95 instructions per edge
// (two square roots, one division), so it sits between the HBM and the FP64-issue rooflines.
#include <cstdint>
#include <cstdlib>
#include "halfplane_dev.cuh"

namespace ldcbf {

constexpr int K1_THREADS = 128;

static bool k1_narrow_loads() {      // LDCBF_K1_NARROW=1: 16-byte loads whatever the alignment (A/B measurements, tests)
    static const bool v = [] { const char* e = getenv("LDCBF_K1_NARROW"); return e && atoi(e) != 0; }();
    return v;
}

// Thread t -> obstacle o = t / Bpad, scenario b = t % Bpad (Bpad = B rounded up to a warp): the 32 lanes of a warp walk
// the SAME obstacle index of 32 consecutive scenarios.  Rings of one index have similar sizes (config 2: 9 / 19 / 24
// vertices for o = 0 / 1 / 2), so a warp's lanes finish together; with the natural (b, o) order a warp mixed all
// sizes and ran at mean/max = 72 % lane occupancy (22 of 32 lanes in ncu).
template <bool EXACT, bool WIDE>
__global__ void __launch_bounds__(K1_THREADS) halfplane_kernel(int B, int Bpad, int max_obs, int max_verts,
                                                               const double* __restrict__ pos, int pos_stride,
                                                               int y_off, const double2* __restrict__ verts,
                                                               const int32_t* __restrict__ nverts,
                                                               const int32_t* __restrict__ nobs,
                                                               double4* __restrict__ c_eta) {
    const long long t = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
    const int o = (int)(t / Bpad), b = (int)(t - (long long)o * Bpad);
    if (o >= max_obs || b >= B) return;
    const size_t pair = (size_t)b * max_obs + o;
    const int V = (o < nobs[b]) ? min(nverts[pair], max_verts) : 0;
    if (V <= 0) { c_eta[pair] = make_double4(0.0, 0.0, 0.0, 0.0); return; }
    const double px = pos[(size_t)b * pos_stride], py = pos[(size_t)b * pos_stride + y_off];
    c_eta[pair] = WIDE ? halfplane_serial_wide<EXACT>(px, py, verts + pair * max_verts, V)
                       : halfplane_serial<EXACT>(px, py, verts + pair * max_verts, V);
}

// Small-batch variant: G lanes per (scenario, obstacle) pair, lane l takes edges l, l+G, ...  With a few thousand
// scenarios the thread-per-ring kernel above leaves most of the GPU idle and its latency is one serial walk over
// the ring (15 us at B = 4096); splitting each ring over 8 lanes cuts that walk to 3 edges.  Same per-edge
// arithmetic; the first strict minimum is kept by an xor-butterfly that breaks ties towards the lower edge index,
// the crossing parity is summed.
template <bool EXACT, int G>
__global__ void __launch_bounds__(128) halfplane_split_kernel(int n_pairs, int max_obs, int max_verts,
                                                              const double* __restrict__ pos, int pos_stride, int y_off,
                                                              const double2* __restrict__ verts,
                                                              const int32_t* __restrict__ nverts,
                                                              const int32_t* __restrict__ nobs,
                                                              double4* __restrict__ c_eta) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = tid / G, lane = tid % G;
    if (pair >= n_pairs) return;                    // whole groups leave together (blockDim is a multiple of G)
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const int b = pair / max_obs, o = pair - b * max_obs;
    const int V = (o < nobs[b]) ? min(nverts[pair], max_verts) : 0;
    if (V <= 0) {
        if (lane == 0) c_eta[pair] = make_double4(0.0, 0.0, 0.0, 0.0);
        return;
    }
    const double px = pos[(size_t)b * pos_stride], py = pos[(size_t)b * pos_stride + y_off];
    const double2* ring = verts + (size_t)pair * max_verts;
    double best = KEY_NONE, bcx = 0.0, bcy = 0.0;
    int be = 0x7fffffff, cross = 0;
    for (int e = lane; e < V; e += G) {
        const double2 A = __ldg(ring + e), Bv = __ldg(ring + ((e + 1 == V) ? 0 : e + 1));
        double cx, cy;
        const double key = edge_closest<EXACT>(px, py, A, Bv, cx, cy, cross);
        if (key_less<EXACT>(key, best)) { best = key; bcx = cx; bcy = cy; be = e; }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const double ok = __shfl_xor_sync(gmask, best, off, G);
        const double ocx = __shfl_xor_sync(gmask, bcx, off, G), ocy = __shfl_xor_sync(gmask, bcy, off, G);
        const int oe = __shfl_xor_sync(gmask, be, off, G);
        cross += __shfl_xor_sync(gmask, cross, off, G);
        // order by the rounded distance, ties towards the lower edge index (= the first strict minimum of the walk)
        if (key_less<EXACT>(ok, best) || (!key_less<EXACT>(best, ok) && oe < be)) { best = ok; bcx = ocx; bcy = ocy; be = oe; }
    }
    if (lane == 0) c_eta[pair] = finish_halfplane<EXACT>(px, py, bcx, bcy, cross);
}

int launch_halfplanes(int B, int max_obs, int max_verts, const double* pos, int pos_stride, int y_off,
                      const double* verts, const int32_t* nverts, const int32_t* nobs, double* c_eta,
                      bool fast_geometry, void* cuda_stream) {
    if (B < 0 || max_obs <= 0 || max_verts <= 0) return LDCBF_E_ARG;
    if (B == 0) return LDCBF_OK;
    if (!pos || !verts || !nverts || !nobs || !c_eta) return LDCBF_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int n_pairs = B * max_obs;
    if ((long long)n_pairs * 8 <= 148 * 2048 && max_verts > 8) {       // the GPU has idle lanes: split every ring over 8
        const unsigned grid = (unsigned)(((size_t)n_pairs * 8 + 127) / 128);
        const double2* v2 = reinterpret_cast<const double2*>(verts);
        double4* ce = reinterpret_cast<double4*>(c_eta);
        if (fast_geometry)
            halfplane_split_kernel<false, 8><<<grid, 128, 0, st>>>(n_pairs, max_obs, max_verts, pos, pos_stride, y_off, v2, nverts, nobs, ce);
        else
            halfplane_split_kernel<true, 8><<<grid, 128, 0, st>>>(n_pairs, max_obs, max_verts, pos, pos_stride, y_off, v2, nverts, nobs, ce);
        return check_launch();
    }
    const int Bpad = (B + 31) / 32 * 32;
    const unsigned grid = (unsigned)(((long long)Bpad * max_obs + K1_THREADS - 1) / K1_THREADS);
    // 256-bit vertex loads need 32-byte aligned vertex pairs: an aligned base and an even number of slots per ring
    const bool wide = (reinterpret_cast<uintptr_t>(verts) % 32 == 0) && (max_verts % 2 == 0) && !k1_narrow_loads();
    auto kern = fast_geometry ? (wide ? halfplane_kernel<false, true> : halfplane_kernel<false, false>)
                              : (wide ? halfplane_kernel<true, true> : halfplane_kernel<true, false>);
    kern<<<grid, K1_THREADS, 0, st>>>(B, Bpad, max_obs, max_verts, pos, pos_stride, y_off,
                                      reinterpret_cast<const double2*>(verts), nverts, nobs,
                                      reinterpret_cast<double4*>(c_eta));
    return check_launch();
}

}  // namespace ldcbf

extern "C" int ldcbf_halfplanes_f64(int B, int max_obs, int max_verts, const double* pos, const double* verts,
                                    const int32_t* nverts, const int32_t* nobs, double* c_eta, void* cuda_stream) {
    return ldcbf::launch_halfplanes(B, max_obs, max_verts, pos, 2, 1, verts, nverts, nobs, c_eta, false, cuda_stream);
}
