// Union-find primitives of the GPU connected-component kernels (ccl.cu, refine.cu): parent links
// only ever decrease (atomicMin), so a component ends up rooted at its smallest pixel index.
#pragma once
#include "common.cuh"

namespace cl4 {

__device__ __forceinline__ int ccl_find(const int* __restrict__ L, int i) {
    int p = L[i];
    while (p != i) {
        i = p;
        p = L[i];
    }
    return i;
}

__device__ __forceinline__ void ccl_union(int* L, int a, int b) {
    bool done;
    do {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

}  // namespace cl4
