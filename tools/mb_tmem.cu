// Micro-benchmark: tensor-memory (TMEM) read / write throughput and latency from CUDA cores (tcgen05.ld / tcgen05.st,
// SASS LDTM / STTM), 4 or 8 warps per SM, each warp on its own lane quadrant.  Question: can thread-private data (the PAMR
// affinity weights) be streamed from TMEM instead of living in registers?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_tmem tools/mb_tmem.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void ld32(uint32_t (&r)[32], uint32_t addr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
}
__device__ __forceinline__ void st32(const uint32_t (&r)[32], uint32_t addr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
          "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
          "r"(r[31]), "r"(addr));
}

// MODE 0: back-to-back loads, one wait at the end of each group of `depth` (throughput)
// MODE 1: load -> wait -> dependent use (latency)
// MODE 2: stores
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, long long* cyc, int check) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = base_s;
    // lane quadrant of this warp, and its half of the 512 columns when two warps share a quadrant
    const uint32_t quad = (warp & 3) * 32, col0 = (warp >> 2) * 256;
    const uint32_t addr0 = base + (quad << 16) + col0;
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) r[j] = threadIdx.x * 1000 + j;
    for (int c = 0; c < 256; c += 32) {
        for (int j = 0; j < 32; ++j) r[j] = threadIdx.x * 1000 + c + j;
        st32(r, addr0 + c);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 192; c += 32) {  // 192 columns = the 192 weights of a thread
                ld32(r, addr0 + c);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += r[j];
            }
        }
    } else if (MODE == 1) {
        uint32_t off = 0;
        for (int it = 0; it < iters; ++it) {
            ld32(r, addr0 + off);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            off = (r[0] & 1) ? 32 : 0;  // dependent address (values are even -> 0)
            acc += r[5];
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 192; c += 32) {
                r[0] = it;
                st32(r, addr0 + c);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;");
        }
    }
    const long long t1 = clock64();
    if (check && MODE == 0 && blockIdx.x == 0) {
        ld32(r, addr0 + 64);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        if (r[3] != threadIdx.x * 1000 + 64 + 3) printf("MISMATCH thread %d: %u\n", threadIdx.x, r[3]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

template <int MODE>
void run(int warps, const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(out, iters, cyc, 1);
    k<MODE><<<148, warps * 32>>>(out, iters, cyc, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double n_ops = (double)iters * (MODE == 1 ? 1 : 6);
    const double bytes = n_ops * warps * 32 * 32 * 4;
    printf("%-8s warps/SM %d: %9.0f cycles, %7.1f cycles per x32 op per warp, %7.1f B/cycle/SM\n", name, warps, (double)h[0],
           (double)h[0] / n_ops, bytes / h[0]);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>(4, "ld");  run<0>(8, "ld");
    run<1>(4, "ld-lat"); run<1>(1, "ld-lat");
    run<2>(4, "st");  run<2>(8, "st");
    return 0;
}
