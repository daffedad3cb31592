// Max-pool NMS kernels:
//   * cl4_center_nms   — find_instance_center (reference modules/utils.py:463-502), batched;
//                        ordered (torch.nonzero / row-major) compaction without atomics.
//   * cl4_peak_extract — peak_extract (reference wss/utils.py:3-25): NMS + top-K per (b,c) plane
//                        with a warp-shuffle selection network.
//
// Both share one tile routine: a (TH+2r) x (TW+2r) window of the plane is staged in
// shared memory (outside the image = -inf, ATen's implicit max-pool padding), then a
// separable k x k maximum is taken (columns, then rows).  The maximum propagates NaN
// like ATen's max_pool2d.
#include "common.cuh"

namespace cl4 {

constexpr int kTileW = 64;
constexpr int kTileH = 32;
constexpr int kNmsThreads = 256;
constexpr int kWarpsPerBlock = kNmsThreads / 32;

__host__ __device__ inline int nms_pitch(int r) { return (kTileW + 2 * r) | 1; }  // odd pitch: no bank conflicts
__host__ inline size_t nms_smem_bytes(int r) {
    // s_in: (TH+2r) x pitch, s_col: TH x pitch
    return sizeof(float) * (size_t)nms_pitch(r) * (size_t)(kTileH + 2 * r + kTileH);
}

// Where a tile's values come from.  PlaneSrc: a plane in memory.  UpsampleSrc: the plane
// F.interpolate(small, (H, W), mode="bilinear", align_corners=False) (train.py:431) evaluated on the fly from the small map
// (which stays in L1/L2), so the up-sampled heat map is never written or read: ATen's upsample_bilinear2d arithmetic,
// source index max(scale * (dst + 0.5) - 0.5, 0) with scale = in / out, the four taps combined row-wise then column-wise.
struct PlaneSrc {
    const float* p;
    int W;
    __device__ __forceinline__ float at(int y, int x) const { return __ldg(p + (size_t)y * W + x); }
};
struct UpsampleSrc {
    const float* p;
    int h, w;
    float sy, sx;
    __device__ __forceinline__ float at(int y, int x) const {
        const float fy = fmaxf(fmaf(sy, (float)y + 0.5f, -0.5f), 0.f), fx = fmaxf(fmaf(sx, (float)x + 0.5f, -0.5f), 0.f);
        const int y1 = (int)fy, x1 = (int)fx;
        const int yp = (y1 < h - 1) ? w : 0, xp = (x1 < w - 1) ? 1 : 0;
        const float ly = fy - (float)y1, lx = fx - (float)x1, hy = 1.f - ly, hx = 1.f - lx;
        const float* q = p + (size_t)y1 * w + x1;
        return hy * (hx * __ldg(q) + lx * __ldg(q + xp)) + ly * (hx * __ldg(q + yp) + lx * __ldg(q + yp + xp));
    }
};
struct PeakInput {  // h == 0: heat is [planes][H][W]; otherwise heat is [planes][h][w] and is up-sampled to H x W on the fly
    const float* heat;
    int H, W, h, w;
    float sy, sx;
};
template <class Src>
__device__ __forceinline__ Src make_src(const PeakInput& in, int plane_id);
template <>
__device__ __forceinline__ PlaneSrc make_src<PlaneSrc>(const PeakInput& in, int plane_id) {
    return PlaneSrc{in.heat + (size_t)plane_id * in.H * in.W, in.W};
}
template <>
__device__ __forceinline__ UpsampleSrc make_src<UpsampleSrc>(const PeakInput& in, int plane_id) {
    return UpsampleSrc{in.heat + (size_t)plane_id * in.h * in.w, in.h, in.w, in.sy, in.sx};
}

// Stage the tile (optionally thresholded: v <= thr -> -1, F.threshold semantics) and
// leave in s_col[ty][tx + r] ... the k x k window maximum for every tile pixel.
// Returns through s_in / s_max: centre value at s_in[(ty+r)*pitch + tx+r],
// pooled value at s_max[ty*pitch + tx].
template <bool kThreshold, class Src>
__device__ __forceinline__ void tile_maxpool(const Src& src, int H, int W, int y0, int x0, int r,
                                             float thr, float* s_in, float* s_col) {
    const int pitch = nms_pitch(r);
    const int tw = kTileW + 2 * r, th = kTileH + 2 * r;
    for (int i = threadIdx.x; i < th * tw; i += kNmsThreads) {
        const int ty = i / tw, tx = i - ty * tw;
        const int y = y0 + ty - r, x = x0 + tx - r;
        float v = -INFINITY;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            v = src.at(y, x);
            if (kThreshold) v = (v <= thr) ? -1.f : v;
        }
        s_in[ty * pitch + tx] = v;
    }
    __syncthreads();
    // vertical pass: s_col[ty][tx] = max_{j<k} s_in[ty+j][tx], tx over the widened tile
    for (int i = threadIdx.x; i < kTileH * tw; i += kNmsThreads) {
        const int ty = i / tw, tx = i - ty * tw;
        float m = -INFINITY;
        const float* p = s_in + ty * pitch + tx;
        for (int j = 0; j <= 2 * r; ++j) m = nanmax(m, p[j * pitch]);
        s_col[ty * pitch + tx] = m;
    }
    __syncthreads();
}

__device__ __forceinline__ float row_window_max(const float* s_col, int pitch, int ty, int tx, int r) {
    float m = -INFINITY;
    const float* p = s_col + ty * pitch + tx;
    for (int j = 0; j <= 2 * r; ++j) m = nanmax(m, p[j]);
    return m;
}

// ---------------------------------------------------------------------------------------------
// Fast path for the NMS kernels the reference actually uses (k = 3, 5, 15, 41; R = (k-1)/2 at
// compile time).  Values are mapped to order-preserving int32 keys (NaN -> INT_MAX, so an integer
// max IS ATen's NaN-propagating max; -0.0 folded onto +0.0 so key equality IS float equality) and
// the k-wide running maximum is taken in registers by doubling (max over 1,2,4,.. wide windows),
// first down the columns, then along the rows with 128-bit shared-memory loads.
// ---------------------------------------------------------------------------------------------
constexpr int kKeyNaN = 0x7fffffff;
constexpr int kKeyPad = (int)0x80000000;  // below every real key: ATen's implicit -inf padding

__device__ __forceinline__ int float_key(float v) {
    if (v != v) return kKeyNaN;
    int b = __float_as_int(v);
    b ^= (b >> 31) & 0x7fffffff;
    return (b == -1) ? 0 : b;  // -0.0 -> +0.0
}
__device__ __forceinline__ float key_float(int k) {  // inverse (NaN and -0.0 are not restored exactly)
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k);
}

template <int N, int K, int Wd>
__device__ __forceinline__ void window_max_level(int (&m)[N]) {
    if constexpr (2 * Wd <= K) {
#pragma unroll
        for (int i = 0; i + Wd < N; ++i) m[i] = max(m[i], m[i + Wd]);
        window_max_level<N, K, 2 * Wd>(m);
    } else if constexpr (Wd < K) {
#pragma unroll
        for (int i = 0; i + K - Wd < N; ++i) m[i] = max(m[i], m[i + K - Wd]);
    }
}
// in place: afterwards m[i] = max(m[i .. i+K-1]) for i <= N-K
template <int N, int K>
__device__ __forceinline__ void window_max(int (&m)[N]) {
    window_max_level<N, K, 1>(m);
}

template <int R>
struct FastTile {
    static constexpr int K = 2 * R + 1;
    static constexpr int kCols = kTileW + 2 * R;              // columns of the staged window
    static constexpr int kPitch = (kCols + 3) / 4 * 4 + 4;    // multiple of 4 for the 128-bit row loads
    static constexpr int kRows = kTileH + 2 * R;
    static constexpr size_t kSmem = sizeof(int) * (size_t)kPitch * (kRows + kTileH);
    static constexpr int kRun = 8;                            // rows per vertical task

    // s_key: staged keys, s_col: column-wise window maxima for all kCols columns
    template <bool kThreshold, class Src>
    __device__ static void stage_and_columns(const Src& src, int H, int W, int y0, int x0, float thr,
                                             int* s_key, int* s_col) {
        // a warp per staged row, lanes over its columns: coalesced loads, no integer division; all loads
        // of a warp are issued before the first one is consumed
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
        constexpr int kRowsPerWarp = (kRows + kWarpsPerBlock - 1) / kWarpsPerBlock, kIts = (kCols + 31) / 32;
        float v[kRowsPerWarp][kIts];
#pragma unroll
        for (int q = 0; q < kRowsPerWarp; ++q) {
            const int y = y0 + warp_ + q * kWarpsPerBlock - R;
            const bool row_in = (warp_ + q * kWarpsPerBlock < kRows) && y >= 0 && y < H;
#pragma unroll
            for (int it = 0; it < kIts; ++it) {
                const int x = x0 + it * 32 + lane_ - R;
                v[q][it] = (row_in && it * 32 + lane_ < kCols && x >= 0 && x < W) ? src.at(y, x) : -INFINITY;
            }
        }
#pragma unroll
        for (int q = 0; q < kRowsPerWarp; ++q) {
            const int ty = warp_ + q * kWarpsPerBlock;
            const int y = y0 + ty - R;
            const bool row_in = y >= 0 && y < H;
#pragma unroll
            for (int it = 0; it < kIts; ++it) {
                const int tx = it * 32 + lane_;
                const int x = x0 + tx - R;
                if (ty < kRows && tx < kCols) {
                    int k = kKeyPad;
                    if (row_in && x >= 0 && x < W) {
                        float t = v[q][it];
                        if (kThreshold) t = (t <= thr) ? -1.f : t;  // F.threshold(x, thr, -1)
                        k = float_key(t);
                    }
                    s_key[ty * kPitch + tx] = k;
                }
            }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < kCols * (kTileH / kRun); q += kNmsThreads) {
            const int run = q / kCols, col = q - run * kCols;
            int m[kRun + 2 * R];
#pragma unroll
            for (int j = 0; j < kRun + 2 * R; ++j) m[j] = s_key[(run * kRun + j) * kPitch + col];
            window_max<kRun + 2 * R, K>(m);
#pragma unroll
            for (int j = 0; j < kRun; ++j) s_col[(run * kRun + j) * kPitch + col] = m[j];
        }
        __syncthreads();
    }

    // k x k window maxima of the 4 pixels (ty, 4*q4 .. 4*q4+3) of the tile
    __device__ static void row_max4(const int* s_col, int ty, int q4, int (&out)[4]) {
        constexpr int N = (4 + 2 * R + 3) / 4 * 4;
        int m[N];
        const int4* p = reinterpret_cast<const int4*>(s_col + ty * kPitch + 4 * q4);
#pragma unroll
        for (int j = 0; j < N / 4; ++j) {
            const int4 v = p[j];
            m[4 * j] = v.x; m[4 * j + 1] = v.y; m[4 * j + 2] = v.z; m[4 * j + 3] = v.w;
        }
        window_max<N, K>(m);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = m[j];
    }
};

template <int R>
__global__ void __launch_bounds__(kNmsThreads)
center_flags_fast_kernel(const float* __restrict__ heat, float thr, float min_value, int H, int W, int words_per_row,
                         uint32_t* __restrict__ words) {
    using T = FastTile<R>;
    extern __shared__ __align__(16) int smem_i[];
    int* s_key = smem_i;
    int* s_col = smem_i + T::kPitch * T::kRows;
    const int n = blockIdx.z;
    const int y0 = blockIdx.y * kTileH, x0 = blockIdx.x * kTileW;
    // Centres are sparse: a pixel can only be kept if its own value exceeds thr (a thresholded pixel reads -1, which is never
    // > min_value when min_value >= -1).  A tile without such a pixel needs no halo and no max-pool: one coalesced pass over
    // its own 2048 pixels, 64 zero words out.
    if (min_value >= -1.f) {
        const float* plane = heat + (size_t)n * H * W;
        bool any = false;
#pragma unroll
        for (int k = 0; k < kTileH * kTileW / kNmsThreads; ++k) {
            const int i = k * kNmsThreads + threadIdx.x;
            const int y = y0 + i / kTileW, x = x0 + (i % kTileW);
            if (y < H && x < W) {
                const float v = __ldg(plane + (size_t)y * W + x);
                any |= (v > thr) && (v > min_value);
            }
        }
        if (!__syncthreads_or(any)) {
            if (threadIdx.x < kTileH * (kTileW / 32)) {
                const int y = y0 + threadIdx.x / (kTileW / 32), xw = (x0 >> 5) + threadIdx.x % (kTileW / 32);
                if (y < H && xw < words_per_row) words[((size_t)n * H + y) * words_per_row + xw] = 0u;
            }
            return;
        }
    }
    T::template stage_and_columns<true>(PlaneSrc{heat + (size_t)n * H * W, W}, H, W, y0, x0, thr, s_key, s_col);

    // thread -> (row, 4-pixel group): lanes 0-15 one row, lanes 16-31 the next; two passes cover 32 rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int ty = pass * 16 + warp * 2 + (lane >> 4), q4 = lane & 15;
        int mx[4];
        T::row_max4(s_col, ty, q4, mx);
        unsigned bits = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kv = s_key[(ty + R) * T::kPitch + R + 4 * q4 + j];
            const int y = y0 + ty, x = x0 + 4 * q4 + j;
            const bool keep = (y < H) && (x < W) && (kv == mx[j]) && (kv != kKeyNaN) && (key_float(kv) > min_value);
            bits |= (keep ? 1u : 0u) << j;
        }
        // assemble 32-pixel words: lane l holds pixels 4*(l&15).. of row (l>>4); 8 lanes make a word
        unsigned word = bits << (4 * (lane & 7));
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
        const int xw = (x0 >> 5) + ((lane >> 3) & 1);
        const int y = y0 + ty;
        if ((lane & 7) == 0 && y < H && xw < words_per_row) words[((size_t)n * H + y) * words_per_row + xw] = word;
    }
}

// ---------------------------------------------------------------------------------------------
// Centre NMS, pass 1: one 32-bit keep-mask word per 32 consecutive x.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads)
center_flags_kernel(const float* __restrict__ heat, float thr, float min_value, int r, int H, int W, int words_per_row,
                    uint32_t* __restrict__ words) {
    extern __shared__ float smem[];
    const int pitch = nms_pitch(r);
    float* s_in = smem;
    float* s_col = smem + (size_t)pitch * (kTileH + 2 * r);
    const int n = blockIdx.z;
    const int y0 = blockIdx.y * kTileH, x0 = blockIdx.x * kTileW;
    const float* plane = heat + (size_t)n * H * W;
    tile_maxpool<true>(PlaneSrc{plane, W}, H, W, y0, x0, r, thr, s_in, s_col);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // each warp owns rows warp, warp+8, ...; two 32-wide words per row
    for (int ty = warp; ty < kTileH; ty += kWarpsPerBlock) {
        const int y = y0 + ty;
#pragma unroll
        for (int half = 0; half < kTileW / 32; ++half) {
            const int tx = half * 32 + lane;
            const int x = x0 + tx;
            bool keep = false;
            if (y < H && x < W) {
                const float v = s_in[(ty + r) * pitch + tx + r];
                const float m = row_window_max(s_col, pitch, ty, tx, r);
                keep = (v == m) && (v > min_value);  // "t != pooled -> -1", then "t > 0" (:485,:492)
            }
            const uint32_t word = __ballot_sync(0xffffffffu, keep);
            const int xw = (x0 >> 5) + half;
            if (lane == 0 && y < H && xw < words_per_row) words[((size_t)n * H + y) * words_per_row + xw] = word;
        }
    }
}

// Pass 2: one CTA per map.  Row population counts -> exclusive scan -> ordered emit.
__global__ void __launch_bounds__(1024)
center_compact_kernel(const uint32_t* __restrict__ words, int H, int words_per_row, long long* __restrict__ ctr_out,
                      int* __restrict__ count_out, int max_out, int* __restrict__ row_off_scratch) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int n = blockIdx.x;
    const uint32_t* wn = words + (size_t)n * H * words_per_row;
    int* row_off = row_off_scratch + (size_t)n * H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    // exclusive scan of per-row counts, blockDim.x rows at a time
    for (int yb = 0; yb < H; yb += blockDim.x) {
        const int y = yb + threadIdx.x;
        int cnt = 0;
        if (y < H)
            for (int i = 0; i < words_per_row; ++i) cnt += __popc(wn[(size_t)y * words_per_row + i]);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int v = (lane < nwarps) ? s_warp[lane] : 0;
            int vi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, vi, o);
                if (lane >= o) vi += t;
            }
            s_warp[lane] = vi - v;  // exclusive warp offsets
        }
        __syncthreads();
        const int carry = s_carry;
        if (y < H) row_off[y] = carry + s_warp[warp] + inc - cnt;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[warp] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) count_out[n] = s_carry;
    // ordered emit: one warp per row, lanes over words in order
    long long* out = ctr_out + (size_t)n * max_out * 2;
    for (int y = warp; y < H; y += nwarps) {
        int pos = row_off[y];
        for (int wb = 0; wb < words_per_row; wb += 32) {
            const int wi = wb + lane;
            uint32_t word = (wi < words_per_row) ? wn[(size_t)y * words_per_row + wi] : 0u;
            const int c = __popc(word);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            int p = pos + inc - c;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                if (p < max_out) {
                    out[2 * (size_t)p] = y;
                    out[2 * (size_t)p + 1] = wi * 32 + b;
                }
                ++p;
            }
            pos += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// peak_extract.  Keys are 64-bit: (order-preserving score bits << 32) | ~flat_index, so that a
// plain unsigned "greater" is (score desc, index asc); NaN sorts above everything like torch.topk
// and -0.0 is folded onto +0.0 (they compare equal in the reference).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t idx) {
    uint32_t u;
    if (s != s) u = 0xffffffffu;
    else {
        if (s == 0.f) s = 0.f;  // -0.0 -> +0.0
        u = __float_as_uint(s);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ float key_score(unsigned long long key) {
    uint32_t u = (uint32_t)(key >> 32);
    if (u == 0xffffffffu) return __uint_as_float(0x7fc00000u);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long key) { return 0xffffffffu - (uint32_t)key; }

// A warp keeps the best 32*KPL keys seen so far, sorted descending; element i lives in
// register i/32 of lane i%32.  Key 0 is "empty" (no real key is 0: the index part of a
// real key is >= 1 because idx < 2^32-1).
template <int KPL>
struct WarpTopK {
    unsigned long long k[KPL];
    unsigned long long thresh;  // current K-th best (0 while the list is not full)
    int K;

    __device__ __forceinline__ void init(int K_) {
        K = K_;
#pragma unroll
        for (int i = 0; i < KPL; ++i) k[i] = 0ull;
        thresh = 0ull;
    }
    __device__ __forceinline__ void insert(unsigned long long x, int lane) {  // x is warp-uniform
        // position = number of stored keys greater than x
        int pos = 0;
#pragma unroll
        for (int i = 0; i < KPL; ++i) pos += __popc(__ballot_sync(0xffffffffu, k[i] > x));
#pragma unroll
        for (int i = KPL - 1; i >= 0; --i) {
            unsigned long long up = __shfl_up_sync(0xffffffffu, k[i], 1);
            const unsigned long long prev_last = (i > 0) ? __shfl_sync(0xffffffffu, k[i - 1 < 0 ? 0 : i - 1], 31) : 0ull;
            if (lane == 0) up = prev_last;
            const int gi = i * 32 + lane;
            if (gi == pos) k[i] = x;
            else if (gi > pos) k[i] = up;
        }
        const int last = K - 1;  // the K-th best sits in register last/32 of lane last%32
        unsigned long long t = 0ull;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const unsigned long long v = __shfl_sync(0xffffffffu, k[i], last & 31);
            if ((last >> 5) == i) t = v;
        }
        thresh = t;
    }
    // offer one candidate per lane (0 = none)
    __device__ __forceinline__ void offer(unsigned long long cand, int lane) {
        uint32_t m = __ballot_sync(0xffffffffu, cand > thresh);
        if (KPL == 1 && __popc(m) > 3) {  // many newcomers (e.g. the first fill): one sort + merge
            merge32(cand > thresh ? cand : 0ull, lane);
            return;
        }
        while (m) {
            const int src = __ffs(m) - 1;
            const unsigned long long x = __shfl_sync(0xffffffffu, cand, src);
            if (x > thresh) insert(x, lane);  // warp-uniform branch (thresh and x are uniform)
            m &= m - 1;
        }
    }
    // KPL == 1 only: list <- the 32 largest of (list, 32 candidates).  Bitonic sort of the candidates
    // (descending), elementwise max with the reversed list (a bitonic sequence holding the top 32),
    // bitonic merge back to descending order.
    __device__ __forceinline__ void merge32(unsigned long long v, int lane) {
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
                const bool desc = (kk == 32) || ((lane & kk) == 0);
                const bool lower = (lane & j) == 0;
                v = (lower == desc) ? (v > o ? v : o) : (v < o ? v : o);
            }
        }
        const unsigned long long rev = __shfl_sync(0xffffffffu, k[0], 31 - lane);
        v = v > rev ? v : rev;
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
            v = ((lane & j) == 0) ? (v > o ? v : o) : (v < o ? v : o);
        }
        k[0] = v;
        thresh = __shfl_sync(0xffffffffu, v, (K - 1) & 31);
    }
};

// Pass 1: per tile, NMS then the tile's best K keys -> cand[plane][tile][K].  `bound` (K > 256, selection in rounds of 256):
// only keys below bound[plane], the last key of the previous round, take part.
template <int KPL, class Src>
__global__ void __launch_bounds__(kNmsThreads)
peak_tile_kernel(const PeakInput in, int r, int K, int tiles_x, int tiles_per_plane, const unsigned long long* __restrict__ bound,
                 unsigned long long* __restrict__ cand) {
    const int H = in.H, W = in.W;
    extern __shared__ float smem[];
    const int pitch = nms_pitch(r);
    float* s_in = smem;
    float* s_col = smem + (size_t)pitch * (kTileH + 2 * r);
    const int plane_id = blockIdx.y;
    const int tile = blockIdx.x;
    const int y0 = (tile / tiles_x) * kTileH, x0 = (tile % tiles_x) * kTileW;
    tile_maxpool<false>(make_src<Src>(in, plane_id), H, W, y0, x0, r, 0.f, s_in, s_col);
    const unsigned long long below = bound ? bound[plane_id] : ~0ull;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpTopK<KPL> top;
    top.init(K);
    for (int ty = warp; ty < kTileH; ty += kWarpsPerBlock) {
        const int y = y0 + ty;
#pragma unroll
        for (int half = 0; half < kTileW / 32; ++half) {
            const int tx = half * 32 + lane;
            const int x = x0 + tx;
            unsigned long long key = 0ull;
            if (y < H && x < W) {
                const float v = s_in[(ty + r) * pitch + tx + r];
                const float m = row_window_max(s_col, pitch, ty, tx, r);
                const float peak = __fmul_rn(v, (m == v) ? 1.f : 0.f);  // heat * keep (wss/utils.py:11-13)
                key = make_key(peak, (uint32_t)(y * W + x));
                if (key >= below) key = 0ull;
            }
            top.offer(key, lane);
        }
    }
    // merge the 8 warp lists through shared memory (reuse s_in), then warp 0 reselects
    __syncthreads();
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem);
#pragma unroll
    for (int i = 0; i < KPL; ++i) s_keys[warp * (32 * KPL) + i * 32 + lane] = top.k[i];
    __syncthreads();
    if (warp == 0) {
        WarpTopK<KPL> fin;
        fin.init(K);
        for (int i = lane; i < kWarpsPerBlock * 32 * KPL; i += 32) fin.offer(s_keys[i], lane);
        unsigned long long* dst = cand + ((size_t)plane_id * tiles_per_plane + tile) * K;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int gi = i * 32 + lane;
            if (gi < K) dst[gi] = fin.k[i];
        }
    }
}

// Fast variant of pass 1 for compile-time radii (see FastTile).
template <int KPL, int R, class Src>
__global__ void __launch_bounds__(kNmsThreads)
peak_tile_fast_kernel(const PeakInput in, int K, int tiles_x, int tiles_per_plane, unsigned long long* __restrict__ cand) {
    using T = FastTile<R>;
    const int H = in.H, W = in.W;
    extern __shared__ __align__(16) int smem_i[];
    int* s_key = smem_i;
    int* s_col = smem_i + T::kPitch * T::kRows;
    const int plane_id = blockIdx.y;
    const int tile = blockIdx.x;
    const int y0 = (tile / tiles_x) * kTileH, x0 = (tile % tiles_x) * kTileW;
    T::template stage_and_columns<false>(make_src<Src>(in, plane_id), H, W, y0, x0, 0.f, s_key, s_col);

    // peak = heat * keep (wss/utils.py:11-13) is exactly 0 for almost every pixel (everything that is
    // not a local maximum), and all those zeros tie on the score: among them only the K lowest flat
    // indices can reach the top-K.  So only NON-ZERO peaks go through the warp selection network; the
    // zero-peak pixels are recorded as one bit each and the tile's fillers are read off the bit rows.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s_zero[kTileH][kTileW / 32];
    WarpTopK<KPL> top;
    top.init(K);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int ty = pass * 16 + warp * 2 + (lane >> 4), q4 = lane & 15;
        int mx[4];
        T::row_max4(s_col, ty, q4, mx);
        unsigned zbits = 0;
        unsigned long long keys[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kv = s_key[(ty + R) * T::kPitch + R + 4 * q4 + j];
            const int y = y0 + ty, x = x0 + 4 * q4 + j;
            keys[j] = 0ull;
            if (y < H && x < W) {
                // the value itself where it equals the window maximum; otherwise heat*0, i.e. 0 for finite
                // heat and NaN for NaN / +-inf heat
                int pk;
                if (kv == mx[j] && kv != kKeyNaN) pk = kv;
                else pk = (kv == kKeyNaN || kv == 0x7f800000 || kv == (int)0x807fffff) ? kKeyNaN : 0;
                if (pk == 0) zbits |= 1u << j;
                else keys[j] = ((unsigned long long)((unsigned)pk ^ 0x80000000u) << 32) |
                               (unsigned long long)(0xffffffffu - (unsigned)(y * W + x));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (__any_sync(0xffffffffu, keys[j] != 0ull)) top.offer(keys[j], lane);
        // 32-pixel words of the zero-peak mask: 8 lanes make a word (as in center_flags_fast_kernel)
        unsigned word = zbits << (4 * (lane & 7));
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
        if ((lane & 7) == 0) s_zero[ty][(lane >> 3) & 1] = word;
    }
    __syncthreads();
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_i);
#pragma unroll
    for (int i = 0; i < KPL; ++i) s_keys[warp * (32 * KPL) + i * 32 + lane] = top.k[i];
    __syncthreads();
    if (warp == 0) {
        WarpTopK<KPL> fin;
        fin.init(K);
        for (int i0 = 0; i0 < kWarpsPerBlock * 32 * KPL; i0 += 32) {
            const unsigned long long c = s_keys[i0 + lane];
            if (__any_sync(0xffffffffu, c != 0ull)) fin.offer(c, lane);
        }
        // order of the tile's best: positive peaks (sorted), then zero peaks by ascending index, then
        // negative peaks (sorted); the list holds the non-zero ones
        constexpr unsigned long long kZeroLo = 0x80000000ull << 32;  // smallest key with score +0.0
        int n_pos = 0;
#pragma unroll
        for (int i = 0; i < KPL; ++i) n_pos += __popc(__ballot_sync(0xffffffffu, fin.k[i] >= kZeroLo));
        // zero-peak pixels per tile row (lane = row), exclusive prefix over the rows
        const unsigned z0 = s_zero[lane][0], z1 = s_zero[lane][1];
        const int cnt = __popc(z0) + __popc(z1);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int n_zero = __shfl_sync(0xffffffffu, inc, 31);
        const int n_fill = max(0, min(K - n_pos, n_zero));
        unsigned long long* dst = cand + ((size_t)plane_id * tiles_per_plane + tile) * K;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int gi = i * 32 + lane;
            const unsigned long long kk = fin.k[i];
            const int slot = (kk >= kZeroLo) ? gi : gi + n_fill;  // negatives move behind the fillers
            if (gi < K && slot < K) dst[slot] = kk;               // kk == 0 marks an unused slot
        }
        // fillers: this row's zero-peak pixels with tile-wide rank < n_fill
        int rank = inc - cnt;
        const unsigned y = (unsigned)(y0 + lane);
        for (int half = 0; half < 2 && rank < n_fill; ++half) {
            unsigned m = half ? z1 : z0;
            while (m && rank < n_fill) {
                const int bpos = __ffs(m) - 1;
                m &= m - 1;
                const unsigned idx = y * (unsigned)W + (unsigned)(x0 + half * 32 + bpos);
                dst[n_pos + rank] = kZeroLo | (unsigned long long)(0xffffffffu - idx);
                ++rank;
            }
        }
    }
}

// Pass 2: one warp per plane merges the tile candidates and writes the sorted result.
template <int KPL>
__global__ void __launch_bounds__(32)
peak_merge_kernel(const unsigned long long* __restrict__ cand, int n_cand, int K, int W, float* __restrict__ scores,
                  int* __restrict__ ys, int* __restrict__ xs, int out_stride, int out_off, unsigned long long* __restrict__ bound) {
    const int plane_id = blockIdx.x;
    const int lane = threadIdx.x;
    const unsigned long long* src = cand + (size_t)plane_id * n_cand;
    WarpTopK<KPL> top;
    top.init(K);
    for (int i0 = 0; i0 < n_cand; i0 += 32) {
        const int i = i0 + lane;
        top.offer(i < n_cand ? src[i] : 0ull, lane);
    }
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int gi = i * 32 + lane;
        if (gi < K) {
            const unsigned long long key = top.k[i];
            const uint32_t idx = key_index(key);
            const size_t o = (size_t)plane_id * out_stride + out_off + gi;
            if (bound && gi == K - 1) bound[plane_id] = key;  // the next round selects below this key
            scores[o] = key_score(key);
            ys[o] = (int)__fdiv_rn((float)idx, (float)W);  // (inds / W).int()  (wss/utils.py:18)
            xs[o] = (int)(idx % (uint32_t)W);              // inds % W          (wss/utils.py:19)
        }
    }
}

template <int KPL, int R, class Src>
static int launch_peak_tile_fast_R(const PeakInput& in, unsigned long long* cand, int planes, int K, int tiles_x, int tiles,
                                   cudaStream_t s) {
    size_t smem = FastTile<R>::kSmem;
    const size_t need_keys = sizeof(unsigned long long) * kWarpsPerBlock * 32 * KPL;
    if (smem < need_keys) smem = need_keys;
    cudaError_t e = cudaFuncSetAttribute(peak_tile_fast_kernel<KPL, R, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("peak_extract: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    peak_tile_fast_kernel<KPL, R, Src><<<dim3(tiles, planes), kNmsThreads, smem, s>>>(in, K, tiles_x, tiles, cand);
    return check_launch("peak_tile_fast");
}
// returns 1 when the radius has no fast specialisation
template <int KPL, class Src>
static int launch_peak_tile_fast(const PeakInput& in, unsigned long long* cand, int planes, int r, int K, int tiles_x, int tiles,
                                 cudaStream_t s) {
    switch (r) {
        case 1: return launch_peak_tile_fast_R<KPL, 1, Src>(in, cand, planes, K, tiles_x, tiles, s);
        case 2: return launch_peak_tile_fast_R<KPL, 2, Src>(in, cand, planes, K, tiles_x, tiles, s);
        case 7: return launch_peak_tile_fast_R<KPL, 7, Src>(in, cand, planes, K, tiles_x, tiles, s);
        case 20: return launch_peak_tile_fast_R<KPL, 20, Src>(in, cand, planes, K, tiles_x, tiles, s);
    }
    return 1;
}

constexpr int kPeakRound = 256;  // keys one selection round can hold (WarpTopK<8>)

// K <= 256: one tile pass + one merge.  K > 256: rounds of 256 with the generic tile kernel; round j selects, among the keys
// below the last key of round j-1 (keys are unique: they contain the pixel index), the next 256.
template <int KPL, class Src>
static int launch_peak(const PeakInput& in, float* scores, int* ys, int* xs, unsigned long long* cand, int planes, int r, int K,
                       cudaStream_t s) {
    const int tiles_x = ceil_div(in.W, kTileW), tiles_y = ceil_div(in.H, kTileH);
    const int tiles = tiles_x * tiles_y;
    size_t smem = nms_smem_bytes(r);
    const size_t need_keys = sizeof(unsigned long long) * kWarpsPerBlock * 32 * KPL;
    if (smem < need_keys) smem = need_keys;
    cudaError_t e = cudaFuncSetAttribute(peak_tile_kernel<KPL, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("peak_extract: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    if (K <= kPeakRound) {
        int rc = launch_peak_tile_fast<KPL, Src>(in, cand, planes, r, K, tiles_x, tiles, s);
        if (rc == 1) {  // no compile-time specialisation for this radius: generic kernel
            peak_tile_kernel<KPL, Src><<<dim3(tiles, planes), kNmsThreads, smem, s>>>(in, r, K, tiles_x, tiles, nullptr, cand);
            rc = check_launch("peak_tile");
        }
        if (rc != CL4_OK) return rc;
        peak_merge_kernel<KPL><<<planes, 32, 0, s>>>(cand, tiles * K, K, in.W, scores, ys, xs, K, 0, nullptr);
        return check_launch("peak_merge");
    }
    unsigned long long* bound = cand + (size_t)planes * tiles * kPeakRound;
    for (int done = 0; done < K; done += kPeakRound) {
        const int k = K - done < kPeakRound ? K - done : kPeakRound;
        peak_tile_kernel<KPL, Src><<<dim3(tiles, planes), kNmsThreads, smem, s>>>(in, r, k, tiles_x, tiles, done ? bound : nullptr, cand);
        int rc = check_launch("peak_tile");
        if (rc != CL4_OK) return rc;
        peak_merge_kernel<KPL><<<planes, 32, 0, s>>>(cand, tiles * k, k, in.W, scores, ys, xs, K, done, bound);
        rc = check_launch("peak_merge");
        if (rc != CL4_OK) return rc;
    }
    return CL4_OK;
}

template <class Src>
static int launch_peak_K(const PeakInput& in, float* scores, int* ys, int* xs, unsigned long long* cand, int planes, int r, int K,
                         cudaStream_t s) {
    if (K <= 32) return launch_peak<1, Src>(in, scores, ys, xs, cand, planes, r, K, s);
    if (K <= 64) return launch_peak<2, Src>(in, scores, ys, xs, cand, planes, r, K, s);
    if (K <= 128) return launch_peak<4, Src>(in, scores, ys, xs, cand, planes, r, K, s);
    return launch_peak<8, Src>(in, scores, ys, xs, cand, planes, r, K, s);
}

template <int R>
static int launch_center_flags_fast_R(const float* heat, float thr, float min_value, int H, int W, int wpr,
                                      uint32_t* words, dim3 grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(center_flags_fast_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FastTile<R>::kSmem);
    if (e != cudaSuccess) {
        set_error("center_nms: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    center_flags_fast_kernel<R><<<grid, kNmsThreads, FastTile<R>::kSmem, s>>>(heat, thr, min_value, H, W, wpr, words);
    return check_launch("center_flags_fast");
}
static int launch_center_flags_fast(const float* heat, float thr, float min_value, int r, int H, int W, int wpr,
                                    uint32_t* words, dim3 grid, cudaStream_t s) {
    switch (r) {
        case 1: return launch_center_flags_fast_R<1>(heat, thr, min_value, H, W, wpr, words, grid, s);
        case 2: return launch_center_flags_fast_R<2>(heat, thr, min_value, H, W, wpr, words, grid, s);
        case 7: return launch_center_flags_fast_R<7>(heat, thr, min_value, H, W, wpr, words, grid, s);
        case 20: return launch_center_flags_fast_R<20>(heat, thr, min_value, H, W, wpr, words, grid, s);
    }
    return 1;
}

int launch_center_compact(const uint32_t* words, int N, int H, int words_per_row, long long* ctr_out, int* count_out,
                          int max_out, int* row_off, cudaStream_t s) {
    center_compact_kernel<<<N, 1024, 0, s>>>(words, H, words_per_row, ctr_out, count_out, max_out, row_off);
    return check_launch("center_compact");
}

}  // namespace cl4

extern "C" size_t cl4_center_nms_scratch_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    const size_t words = (size_t)N * H * cl4::ceil_div(W, 32) * sizeof(uint32_t);
    return cl4::align_up(words, 256) + (size_t)N * H * sizeof(int);
}

extern "C" int cl4_center_nms(const float* heat, float threshold, float min_value, int kernel, int N, int H, int W,
                              long long* ctr_out, int* count_out, int max_out, void* scratch, size_t scratch_bytes,
                              cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(N >= 0 && H > 0 && W > 0, CL4_EINVAL, "center_nms: bad shape N=%d H=%d W=%d", N, H, W);
    CL4_REQUIRE(kernel > 0 && (kernel & 1), CL4_EINVAL, "center_nms: nms kernel must be odd and positive, got %d", kernel);
    CL4_REQUIRE(heat && count_out && (ctr_out || max_out == 0) && max_out >= 0, CL4_EINVAL, "center_nms: null pointer");
    CL4_REQUIRE(N <= 65535, CL4_EUNSUPPORTED, "center_nms: N=%d > 65535", N);
    const int r = (kernel - 1) / 2;
    const size_t smem = nms_smem_bytes(r);
    CL4_REQUIRE(smem <= 227 * 1024, CL4_EUNSUPPORTED, "center_nms: nms kernel %d too large for the tile", kernel);
    if (N == 0) return CL4_OK;  // an empty batch needs no scratch (as every other entry point)
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_center_nms_scratch_bytes(N, H, W), CL4_ESCRATCH,
                "center_nms: scratch too small");
    const int wpr = ceil_div(W, 32);
    uint32_t* words = reinterpret_cast<uint32_t*>(scratch);
    int* row_off = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) +
                                          align_up((size_t)N * H * wpr * sizeof(uint32_t), 256));
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaFuncSetAttribute(center_flags_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CL4_REQUIRE(e == cudaSuccess, CL4_ECUDA, "center_nms: smem attribute: %s", cudaGetErrorString(e));
    dim3 grid(ceil_div(W, kTileW), ceil_div(H, kTileH), N);
    int rc = launch_center_flags_fast(heat, threshold, min_value, r, H, W, wpr, words, grid, s);
    if (rc == 1) {  // no compile-time specialisation for this radius: generic kernel
        center_flags_kernel<<<grid, kNmsThreads, smem, s>>>(heat, threshold, min_value, r, H, W, wpr, words);
        rc = check_launch("center_flags");
    }
    if (rc != CL4_OK) return rc;
    return launch_center_compact(words, N, H, wpr, ctr_out, count_out, max_out, row_off, s);
}

extern "C" size_t cl4_peak_extract_scratch_bytes(int B, int C, int H, int W, int kernel, int K) {
    (void)kernel;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    const size_t tiles = (size_t)cl4::ceil_div(W, cl4::kTileW) * cl4::ceil_div(H, cl4::kTileH);
    const size_t k = K < cl4::kPeakRound ? K : cl4::kPeakRound;  // candidates per tile and round; + one bound key per plane
    return ((size_t)B * C * tiles * k + (size_t)B * C) * sizeof(unsigned long long);
}

static int peak_extract_checked(const cl4::PeakInput& in, float* scores, int* ys, int* xs, void* scratch, size_t scratch_bytes,
                                int B, int C, int kernel, int K, cl4_stream_t stream) {
    using namespace cl4;
    const int H = in.H, W = in.W;
    CL4_REQUIRE(B >= 0 && C >= 0 && H > 0 && W > 0, CL4_EINVAL, "peak_extract: bad shape");
    CL4_REQUIRE(kernel > 0 && (kernel & 1), CL4_EINVAL, "peak_extract: kernel must be odd and positive, got %d", kernel);
    CL4_REQUIRE(K >= 1 && (long long)K <= (long long)H * W, CL4_EINVAL, "peak_extract: K=%d out of range for %dx%d", K, H, W);
    CL4_REQUIRE((long long)H * W < 0xffffffffll, CL4_EUNSUPPORTED, "peak_extract: plane too large");
    CL4_REQUIRE((long long)B * C <= 65535, CL4_EUNSUPPORTED, "peak_extract: B*C > 65535");
    if (B * C == 0) return CL4_OK;
    CL4_REQUIRE(in.heat && scores && ys && xs, CL4_EINVAL, "peak_extract: null pointer");
    const int r = (kernel - 1) / 2;
    CL4_REQUIRE(nms_smem_bytes(r) <= 227 * 1024, CL4_EUNSUPPORTED, "peak_extract: kernel %d too large", kernel);
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_peak_extract_scratch_bytes(B, C, H, W, kernel, K), CL4_ESCRATCH,
                "peak_extract: scratch too small");
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(scratch);
    cudaStream_t s = (cudaStream_t)stream;
    if (in.h == 0) return launch_peak_K<PlaneSrc>(in, scores, ys, xs, cand, B * C, r, K, s);
    return launch_peak_K<UpsampleSrc>(in, scores, ys, xs, cand, B * C, r, K, s);
}

extern "C" int cl4_peak_extract(const float* heat, float* scores, int* ys, int* xs, void* scratch, size_t scratch_bytes,
                                int B, int C, int H, int W, int kernel, int K, cl4_stream_t stream) {
    cl4::PeakInput in{heat, H, W, 0, 0, 0.f, 0.f};
    return peak_extract_checked(in, scores, ys, xs, scratch, scratch_bytes, B, C, kernel, K, stream);
}

extern "C" int cl4_peak_extract_upsampled(const float* small, int h, int w, float* scores, int* ys, int* xs, void* scratch,
                                          size_t scratch_bytes, int B, int C, int H, int W, int kernel, int K,
                                          cl4_stream_t stream) {
    CL4_REQUIRE(h > 0 && w > 0, CL4_EINVAL, "peak_extract_upsampled: bad source shape");
    // ATen's area_pixel_compute_scale for align_corners = false without a user scale factor: float(in) / out
    cl4::PeakInput in{small, H, W, h, w, H > 0 ? (float)h / (float)H : 0.f, W > 0 ? (float)w / (float)W : 0.f};
    return peak_extract_checked(in, scores, ys, xs, scratch, scratch_bytes, B, C, kernel, K, stream);
}

// ---------------------------------------------------------------------------------------------
// cam_normalize (reference wss/modules.py:425-434): relu, gating by the image-level labels, bilinear resize
// (align_corners=False) to `size`, division by (plane maximum + 1e-5).  One CTA per (b, c) plane: pass 1 takes the maximum
// of the resized plane, pass 2 writes the quotient (the resized values are recomputed: the source plane sits in L1).
// ---------------------------------------------------------------------------------------------
namespace cl4 {
__global__ void __launch_bounds__(256)
cam_normalize_kernel(const float* __restrict__ cam, const float* __restrict__ label, float* __restrict__ out, int h, int w,
                     int hs, int ws, float sy, float sx) {
    __shared__ float s_red[8];
    const int plane = blockIdx.x;
    const float* src = cam + (size_t)plane * h * w;
    const float lab = label[plane];
    const bool same = (h == hs && w == ws);  // the trainer's call (size = None): the resize is the identity
    auto value = [&](int i) -> float {
        auto tap = [&](int q) {  // F.relu keeps NaN (fmaxf would drop it), then the label gate
            const float v = __ldg(src + q);
            return __fmul_rn((v < 0.f) ? 0.f : v, lab);
        };
        if (same) return tap(i);
        const int y = i / ws, x = i - y * ws;
        const float fy = fmaxf(fmaf(sy, (float)y + 0.5f, -0.5f), 0.f), fx = fmaxf(fmaf(sx, (float)x + 0.5f, -0.5f), 0.f);
        const int y1 = (int)fy, x1 = (int)fx;
        const int yp = (y1 < h - 1) ? w : 0, xp = (x1 < w - 1) ? 1 : 0;
        const float ly = fy - (float)y1, lx = fx - (float)x1, hy = 1.f - ly, hx = 1.f - lx;
        const int q = y1 * w + x1;
        return hy * (hx * tap(q) + lx * tap(q + xp)) + ly * (hx * tap(q + yp) + lx * tap(q + yp + xp));
    };
    const int n = hs * ws;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = nanmax(m, value(i));
    for (int o = 16; o > 0; o >>= 1) m = nanmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_red[0];
    for (int i = 1; i < 8; ++i) m = nanmax(m, s_red[i]);
    const float denom = m + 1e-5f;  // F.adaptive_max_pool2d(cam, 1) + 1e-5
    float* dst = out + (size_t)plane * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __fdiv_rn(value(i), denom);
}
}  // namespace cl4

extern "C" int cl4_cam_normalize(const float* cam, const float* label, float* out, int B, int C, int h, int w, int hs, int ws,
                                 cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 0 && h > 0 && w > 0 && hs > 0 && ws > 0 && (long long)hs * ws < (1ll << 31), CL4_EINVAL,
                "cam_normalize: bad shape");
    if (B * C == 0) return CL4_OK;
    CL4_REQUIRE(cam && label && out, CL4_EINVAL, "cam_normalize: null pointer");
    cam_normalize_kernel<<<B * C, 256, 0, (cudaStream_t)stream>>>(cam, label, out, h, w, hs, ws, (float)h / (float)hs,
                                                                  (float)w / (float)ws);
    return check_launch("cam_normalize");
}
