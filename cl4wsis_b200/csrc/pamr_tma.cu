// PAMR propagation sweep, TMA-staged, four pixels x all taps per thread (reference wss/modules.py:148-149).
// The path of every dilation set other than the class default (which takes pamr_lattice.cu), D <= 6, dilations <= 24.
//
// Layout.  Between sweeps the masks live in HBM as REPLICATE-PADDED planes
// [B*C][H+48][W+48] (24 = largest dilation): the clamped neighbour of wss/modules.py:57 is then
// an ordinary in-bounds read, every 80x80 source window is a plain TMA box and the hot loop has
// no border case.  A small kernel rewrites the 24-pixel frame after each sweep.
//
// Sweep kernel.  One persistent CTA per SM walks over 32x32-pixel output tiles in image-major order
// (neighbouring CTAs share halos in L2).  A thread owns four pixels of one column, 4 rows apart, and
// keeps their 4 x 8D affinity weights in registers for all C classes of the tile (the weights are the
// only per-pixel state; 192 registers at D = 6).  Per class the 80x80 window arrives by ONE
// cp.async.bulk.tensor box load into a 4-stage shared-memory ring guarded by full/empty mbarriers
// (3 to 6 stages and 2 to 5 windows ahead measure the same).  The inner loop is LDS (immediate
// offset) + FFMA; because the four pixels sit 4 rows apart, 49 of their 192 (pixel, tap) sources
// coincide and are loaded once (143 LDS per 192 FFMA at D = 6).  It is bound by instruction latency at
// 8 warps per SM, not by a bandwidth (profiles/r01c_notes.md, r01d_notes.md).
// While the last class of a tile is computed the weight registers are refilled with the next tile's
// values; the weights are stored tile-major ([tile][p/4][32][32][4], written by the weights kernel): a
// thread's 192 weights are 48 float4 loads at immediate offsets from one pointer.
// kWS (D >= 4, default): 384 threads, the third warpgroup's first thread is the TMA producer and the compute
// warps take its registers through setmaxnreg (see pamr_sweep_tma_kernel).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "pamr_internal.cuh"
#include "pamr_sweep.cuh"
#include "pamr_weights.cuh"
#include "tma.cuh"

namespace cl4 {

#ifndef CL4_SWEEP_STAGES
#define CL4_SWEEP_STAGES 4
#endif
#ifndef CL4_SWEEP_PRODUCER
#define CL4_SWEEP_PRODUCER 0
#endif
#ifndef CL4_SWEEP_AHEAD
#define CL4_SWEEP_AHEAD (CL4_SWEEP_STAGES - 1)
#endif
#ifndef CL4_SWEEP_PREFETCH_WEIGHTS
#define CL4_SWEEP_PREFETCH_WEIGHTS 0  // bulk L2 prefetch of the next tile's weights: off, it costs 1.6 % here (0.612 -> 0.602 ms) and 3 % in the lattice kernel
#endif
#ifndef CL4_SWEEP_L2_AHEAD
#define CL4_SWEEP_L2_AHEAD 0
#endif
constexpr int kStages = CL4_SWEEP_STAGES;
constexpr int kAhead = CL4_SWEEP_AHEAD;  // windows in flight beyond the current item (< kStages)
static_assert(kAhead >= 1 && kAhead < kStages, "prefetch distance");
constexpr int kProducerTid = CL4_SWEEP_PRODUCER;  // the thread that issues the TMA loads
constexpr int kStageBytes = kBox * kBox * 4;    // 25600
constexpr int kLoadBytes = kBox * kBox * 4;
constexpr size_t kSweepSmem = (size_t)kStages * kStageBytes + 128;

struct TileCoord {
    int b, y0, x0;
};
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_x, int tiles_per_img) {
    TileCoord tc;
    tc.b = t / tiles_per_img;
    const int r = t - tc.b * tiles_per_img;
    const int tyi = r / tiles_x;
    tc.y0 = tyi * kTile;
    tc.x0 = (r - tyi * tiles_x) * kTile;
    return tc;
}

struct SweepOut {
    float* ptr;           // element (plane 0, y = 0, x = 0) of the output
    long long plane;      // elements between planes
    int pitch;            // elements between rows
};

// Tile-major weight layout read by this kernel: [tile][P/4][32 rows][32 cols][4 taps] — the four
// taps 4g..4g+3 of one pixel are one float4, a warp's load is one contiguous 512-byte run, and a
// thread fetches its 4 x P weights with P (=48 at D=6) 128-bit loads, few enough to be in flight
// at once.
__device__ __forceinline__ size_t tiled_weight_index(size_t tile, int P, int p, int row, int col) {
    return ((tile * (size_t)(P / 4) + (size_t)(p >> 2)) * (kTile * kTile) + (size_t)row * kTile + col) * 4 + (p & 3);
}

#ifndef CL4_SWEEP_CTAS_SMALL_D
#define CL4_SWEEP_CTAS_SMALL_D 1  // experiment: 2 = two CTAs per SM when D <= 3 (96 weight registers)
#endif
// kWS (warp-specialised): a third warpgroup (threads 256..383) only feeds the ring -- its first thread waits
// for released stages and issues the TMA boxes -- and gives its registers to the eight compute warps through
// setmaxnreg (24 / 240 per thread: 128*24 + 256*240 = 64512 registers).  The compute warps then never wait for
// one another: with the producer inside compute warp 0 the whole CTA advances at the pace of that one warp.
constexpr int kWsThreads = kSweepThreads + 128;
#ifdef CL4_SWEEP_DEBUG  // per CTA: [0] total cycles, [1] producer cycles in empty waits, [2..9] cycles of warp w in full waits
__device__ long long g_sweep_dbg[148 * 10];
#endif
template <int D, class DS, bool kWS>
__global__ void __launch_bounds__(kWS ? kWsThreads : kSweepThreads, (D <= 3 && !kWS ? CL4_SWEEP_CTAS_SMALL_D : 1))
pamr_sweep_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wts, SweepOut out, int C,
                      int H, int W, int tiles_x, int tiles_y, int n_tiles, Dilations dil) {
    constexpr int P = 8 * D;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kStageBytes);
    uint64_t* empty = full + kStages;

    const int tid = threadIdx.x;
    const int lane = tid & 31, wrp = tid >> 5;
    const int tx = lane;
    const int ty = (wrp >> 2) * 16 + (wrp & 3);  // rows ty + 4*i
    const int tiles_per_img = tiles_x * tiles_y;

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int n_my = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = n_my * C;  // work items = (tile, class)

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kSweepThreads / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // Staggered class phase: CTA j starts its first tile at class s0 = j*C/gridDim and finishes that
    // tile's first s0 classes at the very end.  All CTAs would otherwise cross tile boundaries in
    // lockstep and fetch 148 x 196 KB of weights at the same instant; with the stagger the weight
    // traffic is spread evenly over time.  Item i of this CTA is (tile ordinal, class) =
    // ((i + s0) / C mod n_my, (i + s0) mod C).
    const int s0 = (int)(((long long)blockIdx.x * C) / gridDim.x);

    // ---- producer (thread 0 only): p_item is the next item to issue.  The padded plane holds pixel
    // (y, x) at (y + 24, x + 24), so the window of tile (y0, x0) starts at (y0, x0).
    int p_item = 0;
#ifdef CL4_SWEEP_DEBUG
    long long dbg_wait = 0, dbg_t0 = clock64();
#endif
    auto issue_next = [&]() {
        const int v = p_item + s0;
        int pk = v / C;
        const int pc = v - pk * C;
        if (pk == n_my) pk = 0;
        const TileCoord ptc = tile_coord(blockIdx.x + pk * gridDim.x, tiles_x, tiles_per_img);
        const int s = p_item % kStages;
#ifdef CL4_SWEEP_DEBUG
        const long long tw0 = clock64();
#endif
        if (p_item >= kStages) {
            if (kWS) mbar_wait_relaxed(&empty[s], (uint32_t)((p_item / kStages - 1) & 1));  // dedicated thread: no spinning
            else mbar_wait(&empty[s], (uint32_t)((p_item / kStages - 1) & 1));
        }
#ifdef CL4_SWEEP_DEBUG
        dbg_wait += clock64() - tw0;
#endif
        mbar_arrive_expect_tx(&full[s], kLoadBytes);
        tma_load_3d(stage0 + (size_t)s * (kBox * kBox), &tmap, &full[s], ptc.x0, ptc.y0, ptc.b * C + pc);
        ++p_item;
#if CL4_SWEEP_L2_AHEAD > 0  // pull a window further ahead into L2 (no shared memory needed for it)
        if (p_item + CL4_SWEEP_L2_AHEAD - 1 < total) {
            const int v2 = p_item + CL4_SWEEP_L2_AHEAD - 1 + s0;
            int qk = v2 / C;
            const int qc = v2 - qk * C;
            if (qk >= n_my) qk -= n_my;
            const TileCoord qtc = tile_coord(blockIdx.x + qk * gridDim.x, tiles_x, tiles_per_img);
            tma_prefetch_l2_3d(&tmap, qtc.x0, qtc.y0, qtc.b * C + qc);
        }
#endif
    };
    if (kWS) {
        if (tid >= kSweepThreads) {  // producer warpgroup
            asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
            if (tid == kSweepThreads) {
                while (p_item < total) issue_next();  // blocks on the empty barrier of the stage it refills
#ifdef CL4_SWEEP_DEBUG
                if (blockIdx.x < 148) g_sweep_dbg[blockIdx.x * 10 + 1] = dbg_wait;
#endif
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    } else if (tid == kProducerTid) {
        for (int i = 0; i < kAhead && p_item < total; ++i) issue_next();
    }

    float w[kPx][P];
    float acc[kPx];
    const int sbase = (ty + kHalo) * kBox + (tx + kHalo);

    // tile-major weights: tile t holds [P][32][32]; pixels outside the image are never stored,
    // so whatever sits in their slots is loaded but unused
    auto weight_ptr = [&](int t) -> const float4* {
        return reinterpret_cast<const float4*>(wts) + (size_t)t * (P / 4 * kTile * kTile) + ty * kTile + tx;
    };

    int k = 0, c = s0;
    unsigned valid = 0u;   // bits 0-3: pixel i lies inside the image
    float* o = out.ptr;
    auto enter_tile = [&](int kk) {  // per-tile state: store pointer and validity of the four pixels
        const TileCoord tc = tile_coord(blockIdx.x + kk * gridDim.x, tiles_x, tiles_per_img);
        const int x = tc.x0 + tx;
        valid = 0u;
#pragma unroll
        for (int i = 0; i < kPx; ++i)
            if (x < W && tc.y0 + ty + i * kRowGap < H) valid |= 1u << i;
        o = out.ptr + (long long)tc.b * C * out.plane + (long long)(tc.y0 + ty) * out.pitch + x;
    };
    // (Optional, off by default.)  Pull the weights of the tile after tile ordinal kk (196 KB, contiguous) into L2
    // while tile kk is being processed, so that the on-the-fly refill hits L2 instead of HBM.
    auto prefetch_next_weights = [&](int kk) {
        int nn = kk + 1;
        if (nn == n_my) nn = 0;
        if (nn == kk) return;
        const float* base = wts + (size_t)(blockIdx.x + nn * gridDim.x) * (P * kTile * kTile);
        constexpr int kChunk = P * kTile * kTile * 4 / 8;  // 8 chunks, issued by 8 different warps
#if CL4_SWEEP_PREFETCH_WEIGHTS
        if (lane == 0) bulk_prefetch_l2(reinterpret_cast<const char*>(base) + (size_t)wrp * kChunk, kChunk);
#else
        (void)base;
        (void)kChunk;
#endif
    };
    if (total > 0) {  // first tile: plain weight fetch
        enter_tile(0);
        prefetch_next_weights(0);
        const float4* wp = weight_ptr(blockIdx.x);
#pragma unroll
        for (int g = 0; g < P / 4; ++g)
#pragma unroll
            for (int i = 0; i < kPx; ++i) {
                const float4 v = __ldg(wp + g * (kTile * kTile) + i * kRowGap * kTile);
                w[i][4 * g + 0] = v.x;
                w[i][4 * g + 1] = v.y;
                w[i][4 * g + 2] = v.z;
                w[i][4 * g + 3] = v.w;
            }
    }

    for (int item = 0; item < total; ++item) {
        if (!kWS && tid == kProducerTid && p_item < total) issue_next();  // refills the stage released by item-1

        const int s = item % kStages;
        const float* sp = stage0 + (size_t)s * (kBox * kBox) + sbase;
        int nk = k + 1;
        if (nk == n_my) nk = 0;
        // last class of this tile visit and another tile follows: refill the weights on the fly
        const bool reload = (c == C - 1) && (item + 1 < total) && (nk != k);
#ifdef CL4_SWEEP_DEBUG
        const long long tf0 = clock64();
#endif
        mbar_wait(&full[s], (uint32_t)((item / kStages) & 1));
#ifdef CL4_SWEEP_DEBUG
        if (kWS) dbg_wait += clock64() - tf0;
#endif

        if (reload)
            sweep_class<D, DS, true>(w, sp, dil, weight_ptr(blockIdx.x + nk * gridDim.x), acc);
        else
            sweep_class<D, DS, false>(w, sp, dil, nullptr, acc);

#pragma unroll
        for (int i = 0; i < kPx; ++i)
            if ((valid >> i) & 1u) o[(long long)c * out.plane + (long long)(i * kRowGap) * out.pitch] = acc[i];

        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp no longer reads the stage

        if (++c == C) {
            c = 0;
            if (nk != k) {
                k = nk;
                enter_tile(k);
                prefetch_next_weights(k);
            }
        }
    }
#ifdef CL4_SWEEP_DEBUG
    if (kWS && lane == 0 && blockIdx.x < 148) {
        g_sweep_dbg[blockIdx.x * 10 + 2 + wrp] = dbg_wait;
        if (wrp == 0) g_sweep_dbg[blockIdx.x * 10] = clock64() - dbg_t0;
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// Affinity weights from a replicate-padded image (reference wss/modules.py:141-145), one CTA per
// 32x32 tile: the K (<= 3) 80x80 channel windows arrive by TMA, every neighbour read is an LDS with
// an immediate offset, and each thread walks its four pixels one after the other (registers hold
// one pixel's 8D deviations and 8D logits).  Writes the tile-major float4 layout the sweep reads.
// ---------------------------------------------------------------------------------------------
constexpr int kWeightsMaxK = 3;

template <int D, class DS>
__global__ void __launch_bounds__(256, 2)
pamr_weights_tma_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ wts, int K, int tiles_x,
                        int tiles_per_img, Dilations dil) {
    constexpr int P = 8 * D;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* win = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWeightsMaxK * kStageBytes);
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const TileCoord tc = tile_coord(blockIdx.x, tiles_x, tiles_per_img);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(K * kStageBytes));
        for (int k = 0; k < K; ++k) tma_load_3d(win + (size_t)k * (kBox * kBox), &tmap, bar, tc.x0, tc.y0, tc.b * K + k);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    float4* o = reinterpret_cast<float4*>(wts) + (size_t)blockIdx.x * (P / 4 * kTile * kTile) + lane;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int row = wrp + 8 * i;
        const float* sp0 = win + (row + kHalo) * kBox + lane + kHalo;
        float logit[P];
        pixel_affinity<D, DS>(sp0, K, kBox * kBox, kBox, dil, logit);
#pragma unroll
        for (int g = 0; g < P / 4; ++g)
            o[(size_t)g * (kTile * kTile) + row * kTile] = make_float4(logit[4 * g], logit[4 * g + 1], logit[4 * g + 2], logit[4 * g + 3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Replicate-padded planes.
// ---------------------------------------------------------------------------------------------
// dst [planes][H+2*pad][W+2*pad] <- replicate-pad(src [planes][H][W]).  One thread per float4 of a
// padded row (W % 4 == 0 on this path, so image columns map to aligned float4 loads).
__global__ void __launch_bounds__(256)
pamr_pad_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W) {
    const int Wp = W + 2 * kHalo, Hp = H + 2 * kHalo, w4 = Wp / 4;
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int yp = blockIdx.y;
    if (x4 >= w4) return;
    const int y = clampi(yp - kHalo, 0, H - 1);
    const float* srow = src + ((size_t)blockIdx.z * H + y) * W;
    const int x = x4 * 4 - kHalo;  // image column of the first element
    float4 v;
    if (x >= 0 && x + 3 < W) {
        v = __ldg(reinterpret_cast<const float4*>(srow + x));
    } else {
        v.x = __ldg(srow + clampi(x, 0, W - 1));
        v.y = __ldg(srow + clampi(x + 1, 0, W - 1));
        v.z = __ldg(srow + clampi(x + 2, 0, W - 1));
        v.w = __ldg(srow + clampi(x + 3, 0, W - 1));
    }
    reinterpret_cast<float4*>(dst + ((size_t)blockIdx.z * Hp + yp) * Wp)[x4] = v;
}

// In place: the 24-pixel frame of every padded plane <- nearest image pixel (replicate padding,
// wss/modules.py:57, for the NEXT sweep).  Only frame cells are written and only image cells are
// read, so blocks need no ordering.  blockIdx.x < bands_blocks: the 2 x 24 full-width rows above
// and below the image, one thread per float4; the remaining blocks: one thread per image row
// writes its 24 + 24 frame columns as 6 + 6 float4.
__global__ void __launch_bounds__(256)
pamr_pad_frame_kernel(float* __restrict__ buf, int H, int W, int bands_blocks) {
    const int Wp = W + 2 * kHalo, Hp = H + 2 * kHalo, w4 = Wp / 4;
    float* pl = buf + (size_t)blockIdx.y * Hp * Wp;
    if ((int)blockIdx.x < bands_blocks) {
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i >= 2 * kHalo * w4) return;
        const int r = i / w4, x4 = i - r * w4;
        const int yp = (r < kHalo) ? r : (H + r);              // 0..23 and H+24..H+47
        const int ys = (r < kHalo) ? kHalo : (H + kHalo - 1);  // first / last image row
        const float* srow = pl + (size_t)ys * Wp;
        const int xp = x4 * 4;
        float4 v;
        if (xp >= kHalo && xp + 3 < W + kHalo) {
            v = *reinterpret_cast<const float4*>(srow + xp);
        } else {
            const float e = (xp < kHalo) ? srow[kHalo] : srow[W + kHalo - 1];
            v = make_float4(e, e, e, e);  // a frame float4 never straddles the image (24 % 4 == 0)
        }
        reinterpret_cast<float4*>(pl + (size_t)yp * Wp)[x4] = v;
    } else {
        const int row = ((int)blockIdx.x - bands_blocks) * 256 + threadIdx.x;
        if (row >= H) return;
        float* rp = pl + (size_t)(row + kHalo) * Wp;
        const float vl = rp[kHalo], vr = rp[kHalo + W - 1];
        const float4 l4 = make_float4(vl, vl, vl, vl), r4 = make_float4(vr, vr, vr, vr);
#pragma unroll
        for (int q = 0; q < kHalo / 4; ++q) {
            reinterpret_cast<float4*>(rp)[q] = l4;
            reinterpret_cast<float4*>(rp + kHalo + W)[q] = r4;
        }
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int encode_tmap_3d_f32(CUtensorMap* map, const float* base, int W, int H, long long planes, int bx, int by) {
    return encode_tmap_3d_f32_strided(map, base, W, H, planes, W, (long long)W * H, bx, by);
}

int encode_tmap_3d_f32_strided(CUtensorMap* map, const float* base, int W, int H, long long planes, int pitch,
                               long long plane_elems, int bx, int by) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t gstride[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)plane_elems * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

bool sweep_tma_applicable(int H, int W, const Dilations& dil, int D) {
    if (D > 6) return false;                       // 4 x 8D weight registers per thread
    if (W % 4 != 0) return false;                  // TMA: 16-byte global row pitch
    if ((long long)H * W < 32 * 32) return false;  // tiny maps: the register/L1 kernel is fine
    for (int i = 0; i < D; ++i)
        if (dil.d[i] > kHalo) return false;
    return true;
}

size_t tiled_weight_elems(int B, int H, int W, int D) {
    return (size_t)B * ceil_div(H, kTile) * ceil_div(W, kTile) * (size_t)(8 * D) * kTile * kTile;
}

size_t padded_plane_elems(int H, int W) { return (size_t)(H + 2 * kHalo) * (size_t)(W + 2 * kHalo); }

int launch_pad_copy(const float* src, float* dst, long long planes, int H, int W, cudaStream_t s) {
    dim3 grid(ceil_div((W + 2 * kHalo) / 4, 256), H + 2 * kHalo, (unsigned)planes);
    pamr_pad_copy_kernel<<<grid, 256, 0, s>>>(src, dst, H, W);
    return check_launch("pamr_pad_copy");
}

int launch_pad_frame(float* buf, long long planes, int H, int W, cudaStream_t s) {
    const int bands_blocks = ceil_div(2 * kHalo * ((W + 2 * kHalo) / 4), 256);
    dim3 grid(bands_blocks + ceil_div(H, 256), (unsigned)planes);
    pamr_pad_frame_kernel<<<grid, 256, 0, s>>>(buf, H, W, bands_blocks);
    return check_launch("pamr_pad_frame");
}

template <int D, class DS>
static int launch_weights_one(const CUtensorMap& tmap, float* w, int K, int tiles_x, int tiles_per_img, int n_tiles,
                              const Dilations& dil, cudaStream_t s) {
    auto kern = pamr_weights_tma_kernel<D, DS>;
    const size_t smem = (size_t)kWeightsMaxK * kStageBytes + 64;
    // set on every call: the attribute is per device, and a cached flag would be wrong for a second device
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pamr_weights_tma: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    kern<<<n_tiles, 256, smem, s>>>(tmap, w, K, tiles_x, tiles_per_img, dil);
    return check_launch("pamr_weights_tma");
}

template <int D>
static int launch_weights_D(const CUtensorMap& tmap, float* w, int K, int tiles_x, int tiles_per_img, int n_tiles,
                            const Dilations& dil, cudaStream_t s) {
    bool voc6 = (D == 6), voc5 = (D == 5);
    for (int i = 0; i < D && i < 6; ++i) {
        voc6 = voc6 && dil.d[i] == DilVoc6::get(i);
        voc5 = voc5 && dil.d[i] == DilVoc5::get(i);
    }
    if (D == 6 && voc6) return launch_weights_one<6, DilVoc6>(tmap, w, K, tiles_x, tiles_per_img, n_tiles, dil, s);
    if (D == 5 && voc5) return launch_weights_one<5, DilVoc5>(tmap, w, K, tiles_x, tiles_per_img, n_tiles, dil, s);
    return launch_weights_one<D, DilRuntime>(tmap, w, K, tiles_x, tiles_per_img, n_tiles, dil, s);
}

bool weights_tma_applicable(int K) { return K >= 1 && K <= kWeightsMaxK; }

// padded_img: [B*K][H+48][W+48] replicate-padded image; w: tile-major weights
int launch_weights_tma(const float* padded_img, float* w, int B, int K, int H, int W, const Dilations& dil, int D,
                       cudaStream_t s) {
    CUtensorMap tmap;
    const int rc = encode_tmap_3d_f32(&tmap, padded_img, W + 2 * kHalo, H + 2 * kHalo, (long long)B * K, kBox, kBox);
    if (rc != 0) {
        set_error("pamr_weights_tma: cuTensorMapEncodeTiled failed (%d)", rc);
        return CL4_ECUDA;
    }
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile);
    const int n_tiles = B * tiles_x * tiles_y;
    switch (D) {
        case 1: return launch_weights_D<1>(tmap, w, K, tiles_x, tiles_x * tiles_y, n_tiles, dil, s);
        case 2: return launch_weights_D<2>(tmap, w, K, tiles_x, tiles_x * tiles_y, n_tiles, dil, s);
        case 3: return launch_weights_D<3>(tmap, w, K, tiles_x, tiles_x * tiles_y, n_tiles, dil, s);
        case 4: return launch_weights_D<4>(tmap, w, K, tiles_x, tiles_x * tiles_y, n_tiles, dil, s);
        case 5: return launch_weights_D<5>(tmap, w, K, tiles_x, tiles_x * tiles_y, n_tiles, dil, s);
        case 6: return launch_weights_D<6>(tmap, w, K, tiles_x, tiles_x * tiles_y, n_tiles, dil, s);
    }
    set_error("pamr_weights_tma: bad D=%d", D);
    return CL4_EUNSUPPORTED;
}

template <int D, class DS, bool kWS>
static int launch_one_ws(const CUtensorMap& tmap, const float* w, const SweepOut& out, int C, int H, int W, int tiles_x,
                         int tiles_y, int n_tiles, const Dilations& dil, cudaStream_t s) {
    auto kern = pamr_sweep_tma_kernel<D, DS, kWS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem);  // per device
    if (e != cudaSuccess) {
        set_error("pamr_sweep_tma: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    const int ctas = kNumSMs * (D <= 3 ? CL4_SWEEP_CTAS_SMALL_D : 1);
    const int grid = n_tiles < ctas ? n_tiles : ctas;
    kern<<<grid, kWS ? kWsThreads : kSweepThreads, kSweepSmem, s>>>(tmap, w, out, C, H, W, tiles_x, tiles_y, n_tiles, dil);
    return check_launch("pamr_sweep_tma");
}

// CL4_SWEEP=ws / nows selects the warp-specialised / the in-warp producer (default: CL4_SWEEP_WS_DEFAULT)
#ifndef CL4_SWEEP_WS_DEFAULT
#define CL4_SWEEP_WS_DEFAULT 1
#endif
template <int D, class DS>
static int launch_one(const CUtensorMap& tmap, const float* w, const SweepOut& out, int C, int H, int W, int tiles_x,
                      int tiles_y, int n_tiles, const Dilations& dil, cudaStream_t s) {
    const char* f = getenv("CL4_SWEEP");
    const bool ws = f ? (strcmp(f, "ws") == 0 || (strcmp(f, "nows") != 0 && CL4_SWEEP_WS_DEFAULT)) : CL4_SWEEP_WS_DEFAULT;
    if (ws && D >= 4) return launch_one_ws<D, DS, true>(tmap, w, out, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
    return launch_one_ws<D, DS, false>(tmap, w, out, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
}

template <int D>
static int launch_D(const CUtensorMap& tmap, const float* w, const SweepOut& out, int C, int H, int W, int tiles_x,
                    int tiles_y, int n_tiles, const Dilations& dil, cudaStream_t s) {
    bool voc6 = (D == 6), voc5 = (D == 5);
    for (int i = 0; i < D && i < 6; ++i) {
        voc6 = voc6 && dil.d[i] == DilVoc6::get(i);
        voc5 = voc5 && dil.d[i] == DilVoc5::get(i);
    }
    if (D == 6 && voc6) return launch_one<6, DilVoc6>(tmap, w, out, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
    if (D == 5 && voc5) return launch_one<5, DilVoc5>(tmap, w, out, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
    return launch_one<D, DilRuntime>(tmap, w, out, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
}

// padded_in: [B*C][H+48][W+48].  out_padded != 0: `out` is a padded buffer of the same shape (its
// frame is rewritten afterwards by launch_pad_frame); otherwise `out` is the plain [B*C][H][W] tensor.
int launch_sweep_tma(const float* w, const float* padded_in, float* out, int out_padded, int B, int C, int H, int W,
                     const Dilations& dil, int D, cudaStream_t s) {
    CUtensorMap tmap;
    const int Wp = W + 2 * kHalo, Hp = H + 2 * kHalo;
    const int rc = encode_tmap_3d_f32(&tmap, padded_in, Wp, Hp, (long long)B * C, kBox, kBox);
    if (rc != 0) {
        set_error("pamr_sweep_tma: cuTensorMapEncodeTiled failed (%d)", rc);
        return CL4_ECUDA;
    }
    SweepOut so;
    if (out_padded) {
        so.ptr = out + (size_t)kHalo * Wp + kHalo;
        so.plane = (long long)Hp * Wp;
        so.pitch = Wp;
    } else {
        so.ptr = out;
        so.plane = (long long)H * W;
        so.pitch = W;
    }
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile);
    const int n_tiles = B * tiles_x * tiles_y;
    switch (D) {
        case 1: return launch_D<1>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 2: return launch_D<2>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 3: return launch_D<3>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 4: return launch_D<4>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 5: return launch_D<5>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 6: return launch_D<6>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
    }
    set_error("pamr_sweep_tma: bad D=%d", D);
    return CL4_EUNSUPPORTED;
}

}  // namespace cl4

#ifdef CL4_SWEEP_DEBUG
extern "C" int cl4_debug_sweep_waits(long long* out1480) {
    return (int)cudaMemcpyFromSymbol(out1480, cl4::g_sweep_dbg, sizeof(long long) * 1480);
}
#endif
