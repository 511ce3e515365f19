// FP64 FMA-chain probe: measures the FP64 pipe peak of the box, the roofline denominator of the
// register-resident solver and the ray caster (MEASURED_PEAKS.json has no fp64 entry).
#include "ldcbf_common.cuh"

namespace ldcbf {
__global__ void __launch_bounds__(256) fp64_fma_probe(int iters, double* out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double a0 = 1.0 + t * 1e-9, a1 = 1.1, a2 = 1.2, a3 = 1.3, a4 = 1.4, a5 = 1.5, a6 = 1.6, a7 = 1.7;
    const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[t] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace ldcbf

extern "C" int ldcbf_probe_fp64_fma(int blocks, int threads, int iters, double* out, void* cuda_stream) {
    using namespace ldcbf;
    if (blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0 || !out) return LDCBF_E_ARG;
    fp64_fma_probe<<<blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(iters, out);
    return check_launch();
}
