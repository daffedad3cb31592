// Pieces shared by the two lattice sweeps (pamr_lattice.cu: one class per window; pamr_duo.cu: class pairs with
// packed fp32 FMAs): thread ownership of the (pixel, tap) pairs of a 32 x 32 tile, the thread-major weight layout and the
// tile walk.  See pamr_lattice.cu for the reasoning behind the ownership.
#pragma once
#include "common.cuh"
#include "pamr_internal.cuh"
#include "pamr_sweep.cuh"
#include "tma.cuh"

namespace cl4 {

#ifndef CL4_LATTICE_STAGES
#define CL4_LATTICE_STAGES 4  // 3..6 measure within 3 %; 4 stages + 2 partial buffers is the fastest (0.506 ms)
#endif

constexpr int kLThreads = 256;                       // 8 compute warps: 0-3 group A, 4-7 group B
constexpr int kLLaunchThreads = kLThreads + 128;     // + a producer warpgroup (its first thread feeds the TMA ring)
constexpr int kLGroupThreads = 128;
constexpr int kLPitch = 84;                          // window pitch in floats; 84 % 32 == 20
constexpr int kLStageFloats = kBox * kLPitch;        // 80 rows x 84 columns = 6720
constexpr int kLStageBytes = kLStageFloats * 4;      // 26880 = 210 * 128
constexpr int kLStages = CL4_LATTICE_STAGES;
constexpr int kLPartPitch = 36;                      // partial sums of group A: 32 rows x 36 floats (36 % 32 == 4)
constexpr int kLPartFloats = kTile * kLPartPitch;    // 1152
constexpr int kLPx = 8;                              // pixels per thread
constexpr int kLTaps = 24;                           // taps per thread and pixel (three dilations)
constexpr int kLW = kLPx * kLTaps;                   // 192 weight registers
constexpr int kLWeightsPerTile = kLW * kLThreads;    // 49152 floats = 48 taps x 1024 pixels
#ifndef CL4_LATTICE_PARTS
#define CL4_LATTICE_PARTS 2
#endif
constexpr int kLParts = CL4_LATTICE_PARTS;           // partial-sum buffers: how far group A may run ahead of group B
constexpr size_t kLSmem = (size_t)kLStages * kLStageBytes + kLParts * kLPartFloats * 4 + (3 * kLStages + 8 * kLParts) * 8 + 64;

// ---- who owns pixel (y, x) of a tile in each group: thread (0..127 within the group) and slot (0..7) ----
struct Owner {
    int thread, slot;
};
// group A: 2 x 4 lattice blocks of spacing 4 inside 8 x 16 super-blocks (16 threads = 4 x 4 phases)
__host__ __device__ inline Owner owner_a(int y, int x) {
    const int sby = y >> 3, ry = y & 7, sbx = x >> 4, rx = x & 15;
    return Owner{(sby * 2 + sbx) * 16 + (ry & 3) * 4 + (rx & 3), (ry >> 2) * 4 + (rx >> 2)};
}
// group B: 4 x 2 blocks of adjacent pixels, 16 blocks per row of blocks (one half-warp)
__host__ __device__ inline Owner owner_b(int y, int x) { return Owner{(y >> 2) * 16 + (x >> 1), (y & 3) * 2 + (x & 1)}; }

// reference tap order (wss/modules.py:30-40): row-major over the 3x3 neighbourhood, centre skipped
__host__ __device__ constexpr int tap_index(int dy, int dx) {
    const int idx = (dy + 1) * 3 + (dx + 1);
    return idx > 4 ? idx - 1 : idx;
}
// lattice offset (di, dj) in units of the spacing: is it a tap of step s (dilation s * spacing)?
__host__ __device__ constexpr bool is_tap(int di, int dj, int s) {
    return (di == -s || di == 0 || di == s) && (dj == -s || dj == 0 || dj == s) && !(di == 0 && dj == 0);
}

// is lattice position (r, c) read by any (pixel, tap) of an A x B block with steps 1..NS?
template <int A, int B, int NS>
__host__ __device__ constexpr bool source_needed(int r, int c) {
    for (int i = 0; i < A; ++i)
        for (int j = 0; j < B; ++j)
            for (int s = 1; s <= NS; ++s)
                if (is_tap(r - i, c - j, s)) return true;
    return false;
}

// The weights of a tile (49152 floats) as float4 groups of four consecutive taps of one pixel slot, thread-major within three
// regions so that a warp's load is one contiguous 512-byte run:
//   [group A: 48 groups][128 threads]   g = slot*6 + q                       (dilations 4, 8, 12)
//   [group B: 32 groups][128 threads]   slot*4 + q, q < 4                    (dilations 1, 2)
//   [group B: 16 groups][128 threads]   slot*2 + (q - 4), q = 4, 5           (dilation 24; one contiguous 32 KB block per tile:
//                                                                             the class-pair sweep keeps these in tensor memory)
// A thread's register index is k = slot*24 + tap-of-the-group's-three-dilations, float4 group g = k / 4 = slot*6 + q.
constexpr int kLGroupsA = 48, kLGroupsBNear = 32, kLGroupsBFar = 16;
constexpr int kLFarF4Base = (kLGroupsA + kLGroupsBNear) * kLGroupThreads;  // float4 index of the far block inside a tile
constexpr int kLFarBytes = kLGroupsBFar * kLGroupThreads * 16;              // 32768
// float4 offset of group g (0..47) of a thread of warp group G (0: A, 1: B) from that thread's base pointer
template <int G>
__host__ __device__ constexpr int weight_group_offset(int g) {
    if (G == 0) return g * kLGroupThreads;
    const int slot = g / 6, q = g % 6;
    return (q < 4 ? slot * 4 + q : kLGroupsBNear + slot * 2 + (q - 4)) * kLGroupThreads;
}
// base pointer (float4) of thread `tid` (0..255) inside a tile's weight block
__host__ __device__ inline int weight_thread_base(int tid) {
    return tid < kLGroupThreads ? tid : kLGroupsA * kLGroupThreads + (tid - kLGroupThreads);
}
template <int G>
__device__ __forceinline__ void load_weight_group(float (&w)[kLW], const float4* __restrict__ wp, const int g) {
    const float4 v = __ldg(wp + weight_group_offset<G>(g));  // (evict-first loads measured slower once the L2 prefetch is off)
    w[4 * g + 0] = v.x;
    w[4 * g + 1] = v.y;
    w[4 * g + 2] = v.z;
    w[4 * g + 3] = v.w;
}
// kSkipFar: group B without dilation 24 (the trainer's set [1,2,4,8,12], or far weights kept elsewhere) leaves taps 16..23
// of every slot unloaded
template <int G, bool kSkipFar>
__device__ __forceinline__ void load_weights(float (&w)[kLW], const float4* __restrict__ wp) {
#pragma unroll
    for (int g = 0; g < kLW / 4; ++g)
        if (!kSkipFar || (g % (kLTaps / 4)) < 4) load_weight_group<G>(w, wp, g);
}

struct LatticeOut {
    float* ptr;       // element (plane 0, y = 0, x = 0) of the output
    long long plane;  // elements between planes (cells: between pair planes)
    int pitch;        // elements between rows (even)
    int cells;        // 0: planes [B*C][H][W]; 1: pair-interleaved cells [B*ceil(C/2)][H][pitch/2][2] (what pamr_duo.cu reads)
};

struct LTile {
    int b, y0, x0;
};
__device__ __forceinline__ LTile ltile(int t, int tiles_x, int tiles_per_img) {
    LTile tc;
    tc.b = t / tiles_per_img;
    const int r = t - tc.b * tiles_per_img;
    const int tyi = r / tiles_x;
    tc.y0 = tyi * kTile;
    tc.x0 = (r - tyi * tiles_x) * kTile;
    return tc;
}

}  // namespace cl4
