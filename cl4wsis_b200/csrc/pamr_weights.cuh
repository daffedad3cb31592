// Affinity weights of one pixel from K channel windows in shared memory (reference wss/modules.py:141-146):
//   x_std = LocalStDev(x); x = -LocalAffinityAbs(x) / (1e-8 + 0.1 * x_std); x = x.mean(1); x = softmax(x, 2)
// Shared by the TMA-staged weights kernel (pamr_tma.cu) and the phase-1 prologue (phase1_fused.cu).
#pragma once
#include "pamr_sweep.cuh"

namespace cl4 {

// sp0: the pixel's cell in channel 0's window (replicate-padded, so every neighbour is an in-window read); channel k's
// window starts k * plane_stride floats later; `pitch` floats per window row.  out[p], p = dilation * 8 + tap.
template <int D, class DS>
__device__ __forceinline__ void pixel_affinity(const float* __restrict__ sp0, int K, int plane_stride, int pitch,
                                               const Dilations& dil, float (&logit)[8 * D]) {
    constexpr int P = 8 * D;
#pragma unroll
    for (int p = 0; p < P; ++p) logit[p] = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const float* sp = sp0 + k * plane_stride;
        const float c = sp[0];
        float dlt[P];  // neighbour - centre; the D centre samples of LocalStDev contribute zeros
#pragma unroll
        for (int di = 0; di < D; ++di) {
            const int d = DS::kStatic ? DS::get(di) : dil.d[di];
#pragma unroll
            for (int j = 0; j < 8; ++j) dlt[di * 8 + j] = sp[tap_dy(j) * d * pitch + tap_dx(j) * d] - c;
        }
        float s1 = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) s1 += dlt[p];
        const float mean = s1 * (1.f / (float)(9 * D));
        float ss = (float)D * mean * mean;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float t = dlt[p] - mean;
            ss = fmaf(t, t, ss);
        }
        const float sd = sqrtf(ss * (1.f / (float)(9 * D - 1)));
        const float ninv = -1.f / (1e-8f + 0.1f * sd);
#pragma unroll
        for (int p = 0; p < P; ++p) logit[p] = fmaf(fabsf(dlt[p]), ninv, logit[p]);
    }
    const float invK = 1.f / (float)K;
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        logit[p] *= invK;
        mx = fmaxf(mx, logit[p]);
    }
    float z = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        logit[p] = __expf(logit[p] - mx);
        z += logit[p];
    }
    const float rz = 1.f / z;
#pragma unroll
    for (int p = 0; p < P; ++p) logit[p] *= rz;
}

}  // namespace cl4
