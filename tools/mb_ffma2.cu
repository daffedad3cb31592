// Micro-benchmark: issue rate and latency of FFMA vs FFMA2 (packed fp32 FMA, sm_100) per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_ffma2 tools/mb_ffma2.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int MODE, int CHAINS>
__global__ void k(float* out, int iters, long long* cyc) {
    float2 a[CHAINS];
    const float2 w = make_float2(1.0001f + threadIdx.x * 1e-6f, 0.9999f), v = make_float2(1e-3f, 2e-3f);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = make_float2(i, -i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, w.x, v.x); a[i].y = fmaf(a[i].y, w.y, v.y); }
                else a[i] = __ffma2_rn(a[i], w, v);
            }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int CHAINS>
void run(int warps, const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<MODE, CHAINS><<<148, warps * 32>>>(out, iters, cyc);
    k<MODE, CHAINS><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double fma_lanes = (double)iters * 8 * CHAINS * 2 * warps * 32;  // scalar FMAs per SM
    printf("%-8s chains %2d warps/SM %2d: %8.0f cycles  -> %6.1f scalar-FMA lanes/cycle/SM (peak 128), %5.2f cycles per pair-update per warp\n", name, CHAINS, warps,
           (double)h[0], fma_lanes / h[0], (double)h[0] / (iters * 8.0 * CHAINS));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 8>(4, "FFMA");  run<1, 8>(4, "FFMA2");
    run<0, 8>(8, "FFMA");  run<1, 8>(8, "FFMA2");
    run<0, 8>(16, "FFMA"); run<1, 8>(16, "FFMA2");
    run<0, 1>(4, "FFMA");  run<1, 1>(4, "FFMA2");
    run<0, 2>(4, "FFMA");  run<1, 2>(4, "FFMA2");
    run<0, 4>(8, "FFMA");  run<1, 4>(8, "FFMA2");
    return 0;
}
