// Shared pieces of the PAMR propagation kernels (pamr_tma.cu: TMA-staged sweep over HBM-resident
// planes; pamr_fused.cu: all iterations on-chip for small maps).
#pragma once
#include "common.cuh"
#include "pamr_internal.cuh"

namespace cl4 {

constexpr int kTile = 32;                       // output tile: 32 x 32 pixels
constexpr int kHalo = kPamrPad;                 // 24: largest supported dilation on these paths
constexpr int kBox = kTile + 2 * kHalo;         // 80
constexpr int kSweepThreads = 256;              // 8 warps; warp (h,q) owns rows h*16 + q + 4*i, i = 0..3
constexpr int kPx = 4;                          // pixels per thread
constexpr int kRowGap = 4;

// compile-time dilation sets get immediate LDS offsets; DilRuntime computes them from registers
struct DilVoc6 {  // PAMR's class default (wss/modules.py:125)
    static constexpr bool kStatic = true;
    __host__ __device__ static constexpr int get(int i) {
        return i == 0 ? 1 : i == 1 ? 2 : i == 2 ? 4 : i == 3 ? 8 : i == 4 ? 12 : i == 5 ? 24 : 1;
    }
};
struct DilVoc5 {  // the trainer's setting (train.py:81)
    static constexpr bool kStatic = true;
    __host__ __device__ static constexpr int get(int i) {
        return i == 0 ? 1 : i == 1 ? 2 : i == 2 ? 4 : i == 3 ? 8 : i == 4 ? 12 : 1;
    }
};
struct DilRuntime {
    static constexpr bool kStatic = false;
    __host__ __device__ static constexpr int get(int) { return 1; }
};

// One class of one tile.  kReload: refill the weight registers with the next tile's weights
// right after their last use (software-pipelined fetch, no extra registers).
template <int D, class DS, bool kReload, int PITCH = kBox>
__device__ __forceinline__ void sweep_class(float (&w)[kPx][8 * D], const float* __restrict__ sp, const Dilations& dil,
                                            const float4* __restrict__ nw, float (&acc)[kPx]) {
#pragma unroll
    for (int i = 0; i < kPx; ++i) acc[i] = 0.f;
#pragma unroll
    for (int g = 0; g < 2 * D; ++g) {  // groups of four taps
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = 4 * g + q, di = p >> 3, j = p & 7;
            const int d = DS::kStatic ? DS::get(di) : dil.d[di];
            const int dy = (j < 3) ? -1 : ((j < 5) ? 0 : 1);
            const int dx = (j < 3) ? (j - 1) : ((j == 3) ? -1 : ((j == 4) ? 1 : (j - 6)));
            const int off = dy * d * PITCH + dx * d;
#pragma unroll
            for (int i = 0; i < kPx; ++i) acc[i] = fmaf(w[i][p], sp[off + i * kRowGap * PITCH], acc[i]);
        }
        if (kReload) {
#pragma unroll
            for (int i = 0; i < kPx; ++i) {
                const float4 v = __ldg(nw + g * (kTile * kTile) + i * kRowGap * kTile);
                w[i][4 * g + 0] = v.x;
                w[i][4 * g + 1] = v.y;
                w[i][4 * g + 2] = v.z;
                w[i][4 * g + 3] = v.w;
            }
        }
    }
}

}  // namespace cl4
