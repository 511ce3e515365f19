// Developer micro-benchmark: dependent-issue latencies (cycles) of the instructions the QP solver's critical path is made of.
#include <cstdio>
#include <cuda_runtime.h>
#define REP 512
__global__ void lat(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0, t1;
    int k = 0;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) x = fma(x, y, b);
    t1 = clock64(); cyc[k++] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) x = x * y;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) x = x + y;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 4 independent DFMA chains (throughput per warp)
    double x1 = x + 1, x2 = x + 2, x3 = x + 3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) { x = fma(x, y, b); x1 = fma(x1, y, b); x2 = fma(x2, y, b); x3 = fma(x3, y, b); }
    t1 = clock64(); cyc[k++] = t1 - t0;
    x += x1 + x2 + x3;
    // division chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) x = b / x + a;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // rsqrt (library)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) x = rsqrt(x) + a;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // rsqrt approx + 2 newton
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) {
        double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        const double hx = 0.5 * x; r = r * (1.5 - hx * r * r); r = r * (1.5 - hx * r * r); x = r + a;
    }
    t1 = clock64(); cyc[k++] = t1 - t0;
    // shuffle chain (64-bit = 2 SHFL)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + a;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // LDS chain (pointer chase through doubles)
    int idx = (int)x & 63;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) idx = (int)sm[idx & 63] & 63;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // DSETP + select chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) x = (x < y) ? x + a : y;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // FFMA chain for comparison
    float f = (float)x, g = (float)y;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) f = fmaf(f, g, 1.0f);
    t1 = clock64(); cyc[k++] = t1 - t0;
    out[threadIdx.x] = x + idx + f;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 64); cudaMallocManaged(&cyc, 8 * 16);
    for (int warm = 0; warm < 2; ++warm) { lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); }
    const char* names[] = {"DFMA dep", "DMUL dep", "DADD dep", "4x DFMA indep (per 4)", "div+add dep", "rsqrt()+add dep", "rsqrt.approx+2NR+add dep",
                           "shfl64+add dep", "LDS+cvt dep", "DSETP+sel dep", "FFMA dep"};
    for (int i = 0; i < 11; ++i) printf("%-28s %.1f cycles\n", names[i], (double)cyc[i] / REP);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
