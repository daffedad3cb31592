// Micro-benchmark: sustained tensor-TMA box-load throughput per SM for the sweep's access pattern.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_tma tools/mb_tma.cu -L/usr/local/cuda/lib64/stubs -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int kStagesMax = 8;

__global__ void __launch_bounds__(128, 1)
tma_only(const __grid_constant__ CUtensorMap tmap, int bw, int bh, int split, int iters, int stages, long long* cyc,
         int planes, int x_align) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[kStagesMax];
    const unsigned bytes = bw * bh * 4;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        // issue `stages` loads ahead, then wait/issue in order: measures pure TMA delivery rate
        auto issue = [&](int it) {
            const int s = it % stages;
            const unsigned b = (unsigned)__cvta_generic_to_shared(&full[s]);
            const int t = blockIdx.x + (it / 21) * gridDim.x, c = it % 21;
            const int img = (t / 256) % 16, r = t % 256;
            const int x0 = (r % 16) * 32 + x_align, y0 = (r / 16) * 32, pl = (img * 21 + c) % planes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
            for (int q = 0; q < split; ++q) {  // split the box into `split` row bands
                const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + (size_t)s * bytes + (size_t)q * (bytes / split));
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(dst), "l"(&tmap), "r"(b), "r"(x0), "r"(y0 + q * (bh / split)), "r"(pl) : "memory");
            }
        };
        for (int i = 0; i < stages - 1 && i < iters; ++i) issue(i);
        for (int it = 0; it < iters; ++it) {
            if (it + stages - 1 < iters) issue(it + stages - 1);
            const unsigned b = (unsigned)__cvta_generic_to_shared(&full[it % stages]);
            const unsigned par = (it / stages) & 1;
            asm volatile("{ .reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1; @P1 bra D; bra W; D: }" ::"r"(b), "r"(par) : "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
    const int W = 560, H = 560, planes = 336;
    float* buf; cudaMalloc(&buf, (size_t)planes * W * H * 4); cudaMemset(buf, 0, (size_t)planes * W * H * 4);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(tma_only, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Cfg { int bw, bh, split, stages, swz, xal; const char* name; };
    Cfg cfgs[] = {
        {80, 80, 1, 6, 0, 0, "80x80 box, 6 stages"},
        {80, 80, 1, 3, 0, 0, "80x80 box, 3 stages"},
        {80, 80, 2, 6, 0, 0, "80x80 as 2 bands of 40 rows"},
        {80, 80, 4, 6, 0, 0, "80x80 as 4 bands of 20 rows"},
        {80, 80, 1, 6, 0, 8, "80x80 box, x0 offset +8 floats (32B aligned only)"},
        {96, 80, 1, 6, 0, 0, "96x80 box (384 B rows)"},
        {64, 80, 1, 6, 0, 0, "64x80 box (256 B rows)"},
        {32, 80, 1, 6, 3, 0, "32x80 box, swizzle 128B"},
    };
    for (auto& c : cfgs) {
        CUtensorMap tm; cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
        cuuint64_t gs[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t bx[3] = {(cuuint32_t)c.bw, (cuuint32_t)(c.bh / c.split), 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            (CUtensorMapSwizzle)c.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        const int iters = 27 * 21;
        const size_t smem = (size_t)c.stages * c.bw * c.bh * 4;
        tma_only<<<148, 128, smem>>>(tm, c.bw, c.bh, c.split, iters, c.stages, cyc, planes, c.xal);
        tma_only<<<148, 128, smem>>>(tm, c.bw, c.bh, c.split, iters, c.stages, cyc, planes, c.xal);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-52s %s %8.0f cycles/box  %6.1f B/cycle/SM  (%.2f TB/s chip @1.965GHz)\n", c.name, cudaGetErrorString(e),
               (double)mx / iters, c.bw * c.bh * 4.0 / ((double)mx / iters), 148 * c.bw * c.bh * 4.0 / ((double)mx / iters) * 1.965e9 / 1e12);
    }
    return 0;
}
