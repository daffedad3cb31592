// Micro-benchmark of the PAMR sweep inner loop in isolation (no TMA, no global traffic):
// how many cycles does one (tile, class) item cost for different thread->pixel mappings?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_sweep tools/mb_sweep.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kBox = 80, kHalo = 24;
__host__ __device__ constexpr int dil(int i) { return i == 0 ? 1 : i == 1 ? 2 : i == 2 ? 4 : i == 3 ? 8 : i == 4 ? 12 : 24; }
__host__ __device__ constexpr int tdy(int j) { return (j < 3) ? -1 : ((j < 5) ? 0 : 1); }
__host__ __device__ constexpr int tdx(int j) { return (j < 3) ? (j - 1) : ((j == 3) ? -1 : ((j == 4) ? 1 : (j - 6))); }

// MODE 0: 512 thr, rows (ty, ty+8), LDS.32      MODE 1: 512 thr, x pairs, LDS.64
// MODE 2: 256 thr, 4 px along x, LDS.128 where aligned (d>=4), LDS.64 (d=2), mixed (d=1)
// MODE 3: like 1 but loads only (no weights), MODE 4: 1024 thr, 1 px, LDS.32, 48 weights
template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? 256 : (MODE == 4 ? 1024 : 512), 1)
kern(float* out, int iters, long long* cycles, const float* __restrict__ wsrc, const float* __restrict__ gsrc, size_t gsrc_items, const CUtensorMap* tmap, float* gout) {
    extern __shared__ __align__(1024) float sm[];
    for (int i = threadIdx.x; i < kBox * kBox * 2; i += blockDim.x) sm[i] = (float)(i % 17) * 0.01f;
    __syncthreads();
    const int tid = threadIdx.x;
    float acc[4] = {0, 0, 0, 0};
    long long t0 = clock64();
    if (MODE == 0) {
        float w0[48], w1[48];
        for (int p = 0; p < 48; ++p) { w0[p] = wsrc[tid * 96 + p]; w1[p] = wsrc[tid * 96 + 48 + p]; }
        const int tx = tid & 31, wrp = tid >> 5, ty = (wrp >> 3) * 16 + (wrp & 7);
        for (int it = 0; it < iters; ++it) {
            const float* sp = sm + (it & 1) * kBox * kBox + (ty + kHalo) * kBox + tx + kHalo;
#pragma unroll
            for (int di = 0; di < 6; ++di)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int off = tdy(j) * dil(di) * kBox + tdx(j) * dil(di);
                    acc[0] = fmaf(w0[di * 8 + j], sp[off], acc[0]);
                    acc[1] = fmaf(w1[di * 8 + j], sp[off + 8 * kBox], acc[1]);
                }
        }
    } else if (MODE == 1 || MODE == 3) {
        float w0[48], w1[48];
        for (int p = 0; p < 48; ++p) { w0[p] = wsrc[tid * 96 + p]; w1[p] = wsrc[tid * 96 + 48 + p]; }
        const int tx = (tid & 15) * 2, ty = tid >> 4;
        for (int it = 0; it < iters; ++it) {
            const float* sp = sm + (it & 1) * kBox * kBox + (ty + kHalo) * kBox + tx + kHalo;
#pragma unroll
            for (int di = 0; di < 6; ++di)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int off = tdy(j) * dil(di) * kBox + tdx(j) * dil(di);
                    float m0, m1;
                    if (((tdx(j) * dil(di)) & 1) == 0) {
                        const float2 v = *reinterpret_cast<const float2*>(sp + off);
                        m0 = v.x; m1 = v.y;
                    } else {
                        m0 = reinterpret_cast<const float2*>(sp + off - 1)->y;
                        m1 = reinterpret_cast<const float2*>(sp + off + 1)->x;
                    }
                    if (MODE == 1) {
                        acc[0] = fmaf(w0[di * 8 + j], m0, acc[0]);
                        acc[1] = fmaf(w1[di * 8 + j], m1, acc[1]);
                    } else {
                        acc[0] += m0; acc[1] += m1;
                    }
                }
        }
    } else if (MODE == 2) {
        float w[4][48];
        for (int p = 0; p < 48; ++p)
            for (int q = 0; q < 4; ++q) w[q][p] = wsrc[tid * 192 + q * 48 + p];
        const int tx = (tid & 7) * 4, ty = tid >> 3;
        for (int it = 0; it < iters; ++it) {
            const float* sp = sm + (it & 1) * kBox * kBox + (ty + kHalo) * kBox + tx + kHalo;
#pragma unroll
            for (int di = 0; di < 6; ++di)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int d = dil(di), s = tdx(j) * d;
                    const int off = tdy(j) * d * kBox + s;
                    float m[4];
                    if ((s & 3) == 0) {
                        const float4 v = *reinterpret_cast<const float4*>(sp + off);
                        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
                    } else if ((s & 1) == 0) {
                        const float2 a = *reinterpret_cast<const float2*>(sp + off);
                        const float2 b = *reinterpret_cast<const float2*>(sp + off + 2);
                        m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y;
                    } else {  // s = +-1: words off..off+3 = one scalar + float2 + scalar
                        m[0] = sp[off];
                        const float2 a = *reinterpret_cast<const float2*>(sp + off + 1);
                        m[1] = a.x; m[2] = a.y;
                        m[3] = sp[off + 3];
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] = fmaf(w[q][di * 8 + j], m[q], acc[q]);
                }
        }
    } else if (MODE == 5) {
        // mode0's loop, reading from a 3-stage ring that a 1-D bulk async copy refills every iteration
        float w0[48], w1[48];
        for (int p = 0; p < 48; ++p) { w0[p] = wsrc[tid * 96 + p]; w1[p] = wsrc[tid * 96 + 48 + p]; }
        const int tx = tid & 31, wrp = tid >> 5, ty = (wrp >> 3) * 16 + (wrp & 7);
        __shared__ __align__(8) unsigned long long bar[3];
        const unsigned bytes = kBox * kBox * 4;
        if (tid == 0) {
            for (int i = 0; i < 3; ++i)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[i])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        auto issue = [&](int it) {
            const int s = it % 3;
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + s * kBox * kBox);
            const float* src = gsrc + ((size_t)(blockIdx.x * 131 + it) % gsrc_items) * (kBox * kBox);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src), "r"(bytes), "r"(b)
                         : "memory");
        };
        if (tid == 0) { issue(0); issue(1); }
        for (int it = 0; it < iters; ++it) {
            __syncthreads();
            if (tid == 0 && it + 2 < iters) issue(it + 2);
            const int s = it % 3;
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
            const unsigned par = (it / 3) & 1;
            asm volatile("{ .reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1; @P1 bra D; bra W; D: }" ::"r"(b), "r"(par) : "memory");
            const float* sp = sm + s * kBox * kBox + (ty + kHalo) * kBox + tx + kHalo;
#pragma unroll
            for (int di = 0; di < 6; ++di)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int off = tdy(j) * dil(di) * kBox + tdx(j) * dil(di);
                    acc[0] = fmaf(w0[di * 8 + j], sp[off], acc[0]);
                    acc[1] = fmaf(w1[di * 8 + j], sp[off + 8 * kBox], acc[1]);
                }
        }
    } else if (MODE == 6 || MODE == 7) {
        float w0[48], w1[48];
        for (int p = 0; p < 48; ++p) { w0[p] = wsrc[tid * 96 + p]; w1[p] = wsrc[tid * 96 + 48 + p]; }
        const int tx = tid & 31, wrp = tid >> 5, ty = (wrp >> 3) * 16 + (wrp & 7);
        __shared__ __align__(8) unsigned long long bar[3];
        const unsigned bytes = kBox * kBox * 4;
        if (tid == 0) {
            for (int i = 0; i < 3; ++i)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[i])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        // tile walk like the real kernel: tile t = blockIdx.x + k*gridDim.x over 16x16 tiles x images, 21 classes each
        auto issue = [&](int it) {
            const int s = it % 3;
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + s * kBox * kBox);
            const int k = it / 21, c = it % 21;
            const int t = blockIdx.x + k * gridDim.x;
            const int img = (t / 256) % 16, r = t % 256;
            const int x0 = (r % 16) * 32 - 24, y0 = (r / 16) * 32 - 24, pl = img * 21 + c;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(dst), "l"(tmap), "r"(b), "r"(x0), "r"(y0), "r"(pl) : "memory");
        };
        if (tid == 0) { issue(0); issue(1); }
        for (int it = 0; it < iters; ++it) {
            __syncthreads();
            if (tid == 0 && it + 2 < iters) issue(it + 2);
            const int s = it % 3;
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
            const unsigned par = (it / 3) & 1;
            asm volatile("{ .reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1; @P1 bra D; bra W; D: }" ::"r"(b), "r"(par) : "memory");
            const float* sp = sm + s * kBox * kBox + (ty + kHalo) * kBox + tx + kHalo;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int di = 0; di < 6; ++di)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int off = tdy(j) * dil(di) * kBox + tdx(j) * dil(di);
                    a0 = fmaf(w0[di * 8 + j], sp[off], a0);
                    a1 = fmaf(w1[di * 8 + j], sp[off + 8 * kBox], a1);
                }
            if (MODE == 7) {
                const int k = it / 21, c = it % 21;
                const int t = blockIdx.x + k * gridDim.x;
                const int img = (t / 256) % 16, r = t % 256;
                float* o = gout + ((size_t)(img * 21 + c) * 512 + (r / 16) * 32 + ty) * 512 + (r % 16) * 32 + tx;
                o[0] = a0; o[8 * 512] = a1;
            }
            acc[0] += a0; acc[1] += a1;
        }
    } else {  // MODE 4
        float w0[48];
        for (int p = 0; p < 48; ++p) w0[p] = wsrc[tid * 48 + p];
        const int tx = tid & 31, ty = tid >> 5;
        for (int it = 0; it < iters; ++it) {
            const float* sp = sm + (it & 1) * kBox * kBox + (ty + kHalo) * kBox + tx + kHalo;
#pragma unroll
            for (int di = 0; di < 6; ++di)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int off = tdy(j) * dil(di) * kBox + tdx(j) * dil(di);
                    acc[0] = fmaf(w0[di * 8 + j], sp[off], acc[0]);
                }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc[0] + acc[1] + acc[2] + acc[3];
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, size_t gsrc_items = 1) {
    float* out; long long* cyc; float* wsrc;
    cudaMalloc(&wsrc, 1024 * 192 * 4); cudaMemset(wsrc, 0, 1024 * 192 * 4);
    const int grid = 148, iters = (MODE >= 6) ? 27 * 21 : 2100;
    cudaMalloc(&out, grid * 1024 * 4);
    cudaMalloc(&cyc, grid * 8);
    const size_t smem = kBox * kBox * 3 * 4;
    CUtensorMap* d_tmap = nullptr; float* gout = nullptr;
    if (MODE == 6 || MODE == 7) {
        float* planes; const size_t np = 16 * 21; cudaMalloc(&planes, np * 512 * 512 * 4); cudaMemset(planes, 0, np * 512 * 512 * 4);
        cudaMalloc(&gout, np * 512 * 512 * 4);
        CUtensorMap tm; cuuint64_t gd[3] = {512, 512, np}; cuuint64_t gs[2] = {512 * 4, 512 * 512 * 4}; cuuint32_t bx[3] = {80, 80, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, planes, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
        cudaMalloc(&d_tmap, sizeof(tm)); cudaMemcpy(d_tmap, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    }
    float* gsrc; cudaMalloc(&gsrc, gsrc_items * kBox * kBox * 4); cudaMemset(gsrc, 0, gsrc_items * kBox * kBox * 4);
    cudaFuncSetAttribute(kern<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<grid, threads, smem>>>(out, iters, cyc, wsrc, gsrc, gsrc_items, d_tmap, gout);
    cudaEventRecord(e0);
    kern<MODE><<<grid, threads, smem>>>(out, iters, cyc, wsrc, gsrc, gsrc_items, d_tmap, gout);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s %s  %.3f ms  %7.0f cycles per 1024-px class item (clock64), %7.0f by events@1.965GHz\n", name,
           cudaGetErrorString(e), ms, (double)mx / iters, ms * 1e-3 * 1.965e9 / iters);
    cudaFree(out); cudaFree(cyc); cudaFree(gsrc);
}

int main() {
    run<0>("mode0: 512thr rows(y,y+8) LDS.32", 512);
    run<1>("mode1: 512thr x-pairs LDS.64", 512);
    run<3>("mode3: 512thr x-pairs LDS.64 loads only", 512);
    run<2>("mode2: 256thr 4px LDS.128/64/32", 256);
    run<4>("mode4: 1024thr 1px LDS.32", 1024);
    run<5>("mode5: mode0 + bulk copy 25.6KB/iter, L2-resident src (2 MB)", 512, 80);
    run<5>("mode5: mode0 + bulk copy 25.6KB/iter, 100 MB src", 512, 4000);
    run<5>("mode5: mode0 + bulk copy 25.6KB/iter, 2 GB src (DRAM)", 512, 80000);
    run<6>("mode6: mode0 + tensor TMA 80x80 box walk (real pattern)", 512);
    run<7>("mode7: mode6 + result stores", 512);
    return 0;
}
