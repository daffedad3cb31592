// Shared helpers for the sm_100a kernels of the CL4WSIS pseudo-label hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "cl4wsis_b200.h"

namespace cl4 {

// Thread-local message behind cl4_last_error().
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    return CL4_OK;
}

#define CL4_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            ::cl4::set_error(__VA_ARGS__); \
            return (code);                \
        }                                 \
    } while (0)

constexpr int kNumSMs = 148;  // B200

struct Dilations {
    int d[CL4_MAX_DILATIONS];
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// max_pool2d semantics of ATen: a NaN anywhere in the window wins.
__device__ __forceinline__ float nanmax(float a, float b) { return (b > a || b != b) ? b : a; }

// Tap order of the 3x3 shift stencils with the centre skipped
// (reference wss/modules.py:30-40): j -> (dy, dx).
__device__ __forceinline__ int tap_dy(int j) { return (j < 3) ? -1 : ((j < 5) ? 0 : 1); }
__device__ __forceinline__ int tap_dx(int j) {
    // j: 0 1 2 3 4 5 6 7 -> dx: -1 0 1 -1 1 -1 0 1
    return (j < 3) ? (j - 1) : ((j == 3) ? -1 : ((j == 4) ? 1 : (j - 6)));
}

}  // namespace cl4
