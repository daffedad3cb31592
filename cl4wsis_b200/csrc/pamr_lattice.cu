// PAMR propagation sweep, "lattice" variant for the dilation set [1,2,4,8,12,24]
// (PAMR's class default, reference wss/modules.py:125; the sweep itself is :148-149) and for the
// trainer's [1,2,4,8,12] (train.py:81; kFar = false: group B without its dilation-24 part).
//
// Why.  The TMA sweep in pamr_tma.cu is bound by shared-memory bandwidth: every FMA needs one
// weight (reused across the C classes -> registers) and one source value (reused only across
// different (pixel, tap) pairs -> shared memory), and with four pixels x 48 taps per thread only
// 49 of 192 sources coincide (143 LDS per 192 FMA).  What a thread can reuse is decided by WHICH
// 192 (pixel, tap) pairs its 192 weight registers belong to.  Here the CTA is split into two warp
// groups of four warps, each thread owning 8 pixels x 24 taps:
//   group A  dilations {4,8,12}: a 2 x 4 block of pixels on the lattice of spacing 4 (rows y0, y0+4;
//            columns x0, x0+4, x0+8, x0+12).  Its 192 (pixel, tap) pairs touch only 76 distinct
//            sources (steps 1, 2, 3 of the same lattice).
//   group B  dilations {1,2,24}: a 4 x 2 block of adjacent pixels, sources read as float2:
//            24 LDS.64 for {1,2} (an 8 x 6 patch) + 32 LDS.64 for dilation 24 = 112 wavefronts.
// The partial sums of A meet B's in shared memory (8 STS per A thread, 4 LDS.64 per B thread, double
// buffered per warp pair behind mbarriers); B adds and stores float2.  Per (tile, class):
// 4*(76+8) + 4*(112+8) = 816 LDS/STS wavefronts (the 4-pixel kernel: 1144) + 210 of TMA writes.
// Eight accumulators per thread also double the FMA-level parallelism.
//
// Producer.  The kernel launches 384 threads: a third warpgroup drops to 24 registers (setmaxnreg) and its first
// thread only waits for released stages and issues the TMA boxes, so the ring is always full and the two compute
// groups (240 registers each) never wait for one another through it.
//
// Padding.  The planes in HBM are not padded: a window's box starts 24 pixels up-left of its tile, the TMA unit zero-fills
// what lies outside the plane, and the other three warps of the producer warpgroup write the clamped neighbour
// (wss/modules.py:57, replicate padding) into those cells before the compute warps see the window.
//
// Status: the default for this dilation set (CL4_SWEEP=nolattice selects the 4-pixel kernel).  B16 C21 512^2:
// 0.506 ms per sweep and no padding / frame kernels against 0.613 + 0.025 ms; 72.9 M shared-memory wavefronts per launch against 98.5 M (ncu).  Nothing is
// saturated (issue 48 %, l1tex 59 %, DRAM 48 %): with 8 compute warps per SM -- all that 48 register-resident weights per
// pixel allow -- the loop is bound by instruction latency.  See profiles/r01c_notes.md and r01d_notes.md.
//
//   tile    32 x 32 pixels, 256 threads; window 80 rows x 84 columns per (tile, class), one
//           cp.async.bulk.tensor box into a ring of stages with full/empty mbarriers.  The pitch of
//           84 floats (= 20 mod 32) makes group A's lane pattern (4 x 4 phases x 2 super-blocks)
//           bank-conflict free; group B's half-warps read 16 adjacent float2.
//   weights thread-major [tile][k/4][thread] float4, k = slot*24 + tap (written by pamr_weights_lattice_kernel):
//           48 128-bit loads per thread, a warp's load is one 512-byte run; same size as the 32x32 tile-major layout.
#include "pamr_lattice.cuh"

namespace cl4 {

// One source value `v` at lattice position (r, c) of an A x B block whose taps are the steps 1..NS of the
// lattice: feed every (pixel, tap) that reads it.  Weight register of (pixel slot, step s, tap): slot*24 + (s-1)*8 + tap.
template <int G, int A, int B, int NS, bool kReload>
__device__ __forceinline__ void feed(float (&w)[kLW], float (&acc)[kLPx], const float v, const int r, const int c,
                                     const float4* __restrict__ nw) {
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j)
#pragma unroll
            for (int s = 1; s <= NS; ++s) {
                const int di = r - i, dj = c - j;
                if (is_tap(di, dj, s)) {
                    const int k = (i * B + j) * kLTaps + (s - 1) * 8 + tap_index(di / s, dj / s);
                    acc[i * B + j] = fmaf(w[k], v, acc[i * B + j]);
                    // sources arrive in row-major order, so taps 3 and 7 are the last uses of their float4:
                    // refill it with the next tile's weights right away (kReload: last class of a tile)
                    if (kReload && (k & 3) == 3) load_weight_group<G>(w, nw, k >> 2);
                }
            }
}

// group A: 2 x 4 lattice block of spacing 4, dilations 4, 8, 12 = steps 1, 2, 3; sp points at the block's
// pixel (0, 0) inside the window.  76 of the 8 x 10 lattice positions are read.
template <bool kReload>
__device__ __forceinline__ void group_a_class(float (&w)[kLW], float (&acc)[kLPx], const float* __restrict__ sp,
                                              const float4* __restrict__ nw) {
#pragma unroll
    for (int r = -3; r < 2 + 3; ++r)
#pragma unroll
        for (int c = -3; c < 4 + 3; ++c)
            if (source_needed<2, 4, 3>(r, c)) {
                const float v = sp[r * 4 * kLPitch + c * 4];
                feed<0, 2, 4, 3, kReload>(w, acc, v, r, c, nw);
            }
}

// group B: 4 x 2 block of adjacent pixels.  Dilations 1 and 2 (steps 1, 2 of the unit lattice): rows -2..5,
// columns -2..3 as float2.  Dilation 24 (third tap set, weights slot*24 + 16 + tap): the two pixels of a row
// share one float2 per tap.
template <bool kReload, bool kFar>
__device__ __forceinline__ void group_b_class(float (&w)[kLW], float (&acc)[kLPx], const float* __restrict__ sp,
                                              const float4* __restrict__ nw) {
#pragma unroll
    for (int r = -2; r < 4 + 2; ++r)
#pragma unroll
        for (int c = -2; c < 4; c += 2) {
            const float2 v = *reinterpret_cast<const float2*>(sp + r * kLPitch + c);
            feed<1, 4, 2, 2, kReload>(w, acc, v.x, r, c, nw);
            feed<1, 4, 2, 2, kReload>(w, acc, v.y, r, c + 1, nw);
        }
    if (kFar) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int a = -1; a <= 1; ++a)
#pragma unroll
            for (int b = -1; b <= 1; ++b) {
                if (a == 0 && b == 0) continue;
                const float2 v = *reinterpret_cast<const float2*>(sp + (i + 24 * a) * kLPitch + 24 * b);
                const int k0 = (i * 2) * kLTaps + 16 + tap_index(a, b), k1 = k0 + kLTaps;
                acc[i * 2] = fmaf(w[k0], v.x, acc[i * 2]);
                acc[i * 2 + 1] = fmaf(w[k1], v.y, acc[i * 2 + 1]);
                if (kReload && (k0 & 3) == 3) {
                    load_weight_group<1>(w, nw, k0 >> 2);
                    load_weight_group<1>(w, nw, k1 >> 2);
                }
            }
    }
}

struct LatticeCtx {
    float* stage0;
    float* part;  // [kLParts buffers][kLPartFloats]; pfull / pempty: [kLParts][4 warp pairs]
    uint64_t *full, *ready, *empty, *pfull, *pempty;  // full: TMA landed; ready: border patched (what the compute warps wait for)
    const float* wts;
    int C, H, W, tiles_x, tiles_per_img, n_my, total, s0;
};

// The producer: one thread of a third warpgroup walks the items of this CTA in order, waits until all eight compute
// warps have released the stage it is about to refill, and issues the window's TMA box.  The compute groups never
// wait for one another through the ring.  (With the producer inside a compute warp the ring depth seen by the
// leading group collapses and random CTAs ran whole launches 1.75x slower -- profiles/r01c_notes.md.)
__device__ __forceinline__ void lattice_producer(const LatticeCtx& cx, const CUtensorMap* tmap) {
    const int C = cx.C;
    // the coordinates of the next box are ready before the wait (tile geometry only changes every C items), so the TMA
    // goes out the moment the stage is released: the release -> refill latency is on the critical path of the ring
    int k = 0, c = cx.s0, s = 0;
    uint32_t phase = 1;  // parity of the previous use of the stage
    LTile tc = ltile(blockIdx.x, cx.tiles_x, cx.tiles_per_img);
    for (int p_item = 0; p_item < cx.total; ++p_item) {
        if (p_item >= kLStages) mbar_wait_relaxed(&cx.empty[s], phase);
        mbar_arrive_expect_tx(&cx.full[s], kLStageBytes);
        tma_load_3d(cx.stage0 + (size_t)s * kLStageFloats, tmap, &cx.full[s], tc.x0 - kHalo, tc.y0 - kHalo, tc.b * C + c);
        if (++s == kLStages) {
            s = 0;
            phase ^= 1u;
        }
        if (++c == C) {
            c = 0;
            const int nk = (k + 1 == cx.n_my) ? 0 : k + 1;
            if (nk != k) {
                k = nk;
                tc = ltile(blockIdx.x + k * gridDim.x, cx.tiles_x, cx.tiles_per_img);
            }
        }
    }
}

// Replicate padding inside the ring (reference wss/modules.py:57: F.pad(mode="replicate")).  The planes in HBM are
// NOT padded: the TMA box starts 24 pixels up-left of the tile and the hardware zero-fills what lies outside the plane.
// The three otherwise idle warps of the producer warpgroup turn those zeros into the clamped neighbour before the
// compute warps see the window: every cell outside the image takes the value of the nearest image pixel, which lies
// inside the same window.  Interior tiles need nothing.  This replaces the replicate-padded planes in HBM, the copy
// into them and the frame rewrite after every sweep (10 % of the step before).
__device__ __forceinline__ void lattice_patcher(const LatticeCtx& cx) {
    const int C = cx.C;
    const int lane = threadIdx.x & 31, pw = (threadIdx.x >> 5) - (kLThreads / 32 + 1);  // patch warp 0..2
    // per-tile geometry is recomputed only when the tile changes (no division on the per-item path: the time between
    // `full` and `ready` is on the critical path of the ring)
    int k = 0, c = cx.s0, s = 0;
    uint32_t phase = 0;
    int rv0 = 0, rv1 = kBox, cv0 = 0, cv1 = kLPitch;
    auto enter_tile = [&](int kk) {
        const LTile tc = ltile(blockIdx.x + kk * gridDim.x, cx.tiles_x, cx.tiles_per_img);
        // valid window rows [rv0, rv1) and columns [cv0, cv1): the part of the 80 x 84 window that lies inside the image
        rv0 = max(0, kHalo - tc.y0);
        rv1 = min(kBox, cx.H - tc.y0 + kHalo);
        cv0 = max(0, kHalo - tc.x0);
        cv1 = min(kLPitch, cx.W - tc.x0 + kHalo);
    };
    if (cx.total > 0) enter_tile(0);
    for (int item = 0; item < cx.total; ++item) {
        mbar_wait_relaxed(&cx.full[s], phase);  // suspended, not spinning
        if (rv0 > 0 || rv1 < kBox || cv0 > 0 || cv1 < kLPitch) {
            float* win = cx.stage0 + (size_t)s * kLStageFloats;
            const int t = pw * 32 + lane;  // 0..95
            // columns left / right of the image: one thread per valid window row, independent stores of one value
            const int r = rv0 + t;
            if (r < rv1) {
                float* row = win + r * kLPitch;
                if (cv0 > 0) {
                    const float v = row[cv0];
#pragma unroll 8
                    for (int c = 0; c < cv0; ++c) row[c] = v;
                }
                if (cv1 < kLPitch) {
                    const float v = row[cv1 - 1];
#pragma unroll 4
                    for (int c = cv1; c < kLPitch; ++c) row[c] = v;
                }
            }
            // rows above / below the image: one thread per window column; the source is the clamped column of the first /
            // last image row, a cell the loop above does not write
            if (t < kLPitch) {
                const int cs = min(max(t, cv0), cv1 - 1);
                if (rv0 > 0) {
                    const float v = win[rv0 * kLPitch + cs];
#pragma unroll 8
                    for (int q = 0; q < rv0; ++q) win[q * kLPitch + t] = v;
                }
                if (rv1 < kBox) {
                    const float v = win[(rv1 - 1) * kLPitch + cs];
#pragma unroll 4
                    for (int q = rv1; q < kBox; ++q) win[q * kLPitch + t] = v;
                }
            }
            fence_proxy_async_smem();  // these generic-proxy writes are ordered before the next TMA refill of the stage
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.ready[s]);
        if (++s == kLStages) {
            s = 0;
            phase ^= 1u;
        }
        if (++c == C) {
            c = 0;
            const int nk = (k + 1 == cx.n_my) ? 0 : k + 1;
            if (nk != k) {
                k = nk;
                enter_tile(k);
            }
        }
    }
}

// The item loop of one warp group (G = 0: A, G = 1: B).  Item i of this CTA is (tile ordinal, class) =
// ((i + s0) / C mod n_my, (i + s0) mod C) as in pamr_tma.cu (staggered class phase).
template <int G, bool kFar>
__device__ __forceinline__ void lattice_group(const LatticeCtx& cx, const LatticeOut& out) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int tg = tid - G * kLGroupThreads;
    const int C = cx.C;

    // thread geometry: tile-relative row / column of the thread's block origin
    int ry, rx;
    if (G == 0) {
        const int sb = tg >> 4;
        ry = (sb >> 1) * 8 + ((tg >> 2) & 3);
        rx = (sb & 1) * 16 + (tg & 3);
    } else {
        ry = (tg >> 4) * 4;
        rx = (tg & 15) * 2;
    }
    const int tb = (ry + kHalo) * kLPitch + rx + kHalo;  // window offset of the block origin
    const int pbase = ry * kLPartPitch + rx;             // partial-buffer offset

    float w[kLW];
    float acc[kLPx];
    auto weight_ptr = [&](int t) -> const float4* {
        return reinterpret_cast<const float4*>(cx.wts + (size_t)t * kLWeightsPerTile) + weight_thread_base(tid);
    };
    int k = 0, c = cx.s0;
    // group B stores the pixels: store pointer and number of valid rows of its 4 x 2 block
    float* o = nullptr;
    int nrows = 0;
    auto enter_tile = [&](int kk) {
        if (G != 1) return;
        const LTile tc = ltile(blockIdx.x + kk * gridDim.x, cx.tiles_x, cx.tiles_per_img);
        const int y = tc.y0 + ry, x = tc.x0 + rx;
        nrows = (x < cx.W) ? min(max(cx.H - y, 0), 4) : 0;
        o = out.cells ? out.ptr + (long long)tc.b * ((C + 1) >> 1) * out.plane + (long long)y * out.pitch + 2 * x
                      : out.ptr + (long long)tc.b * C * out.plane + (long long)y * out.pitch + x;
    };
    if (cx.total > 0) {
        enter_tile(0);
        load_weights<G, (G == 1 && !kFar)>(w, weight_ptr(blockIdx.x));
    }

    // loop state kept incrementally (no division per item): ring stage + phase, partial buffer + phase,
    // window pointer, hand-over barrier of this warp pair
    const uint32_t full0 = smem_u32(cx.ready), empty0 = smem_u32(cx.empty);  // "full" for the compute warps = patched
    const uint32_t pfull0 = smem_u32(cx.pfull) + 8u * ((tid >> 5) & 3), pempty0 = smem_u32(cx.pempty) + 8u * ((tid >> 5) & 3);
    int stage = 0, pb = 0;
    uint32_t full_phase = 0, part_phase = 0;
    const float* sp = cx.stage0 + tb;
    float* pp = cx.part + pbase;
    // planes: class c at o + c * plane; cells: class c is half (c & 1) of pair plane c >> 1
    auto class_ptr = [&](int cc) -> float* { return out.cells ? o + (long long)(cc >> 1) * out.plane + (cc & 1) : o + (long long)cc * out.plane; };
    float* oc = (G == 1) ? class_ptr(c) : nullptr;

    for (int item = 0; item < cx.total; ++item) {
        // last class of this tile visit and another tile follows: refill the weights on the fly
        bool reload = false;
        const float4* nw = nullptr;
        int nk = k;
        if (c == C - 1) {
            nk = (k + 1 == cx.n_my) ? 0 : k + 1;
            reload = (item + 1 < cx.total) && (nk != k);
            nw = weight_ptr(blockIdx.x + nk * gridDim.x);
        }
        mbar_wait_u32(full0 + 8u * stage, full_phase);

#pragma unroll
        for (int i = 0; i < kLPx; ++i) acc[i] = 0.f;
        if (G == 0) {
            if (reload) group_a_class<true>(w, acc, sp, nw);
            else group_a_class<false>(w, acc, sp, nw);
        } else {
            if (reload) group_b_class<true, kFar>(w, acc, sp, nw);
            else group_b_class<false, kFar>(w, acc, sp, nw);
        }
        // A's warp j and B's warp j own the same 8 rows of the tile, so the hand-over is per warp pair.  The ring stage is
        // released AFTER the stores of the accumulators, never right after the last FMA: a store needs the FMA results, which
        // need every window load to have RETURNED, and a release-arrive stays below earlier stores.  Placed before them, the
        // arrive can be scheduled above FMAs whose LDS are still in flight, and a fast refill then overwrites window rows that
        // are still being read (found with the class-pair sweep: rare wrong rows at 1024 x 1024).
        if (G == 0) {
            if (item >= kLParts) mbar_wait_u32(pempty0 + 32u * pb, part_phase ^ 1u);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) pp[i * 4 * kLPartPitch + j * 4] = acc[i * 4 + j];
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_u32(pfull0 + 32u * pb);
                mbar_arrive_u32(empty0 + 8u * stage);  // this warp no longer reads the window
            }
        } else {
            mbar_wait_u32(pfull0 + 32u * pb, part_phase);
            float2 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float2*>(pp + i * kLPartPitch);
            __syncwarp();
            if (lane == 0) mbar_arrive_u32(pempty0 + 32u * pb);
            // the output is not read again before the next sweep: evict-first stores keep L2 for windows and weights
            if (!out.cells) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < nrows)
                        __stcs(reinterpret_cast<float2*>(oc + (long long)i * out.pitch), make_float2(acc[2 * i] + a[i].x, acc[2 * i + 1] + a[i].y));
            } else if ((C & 1) && c == C - 1) {
                // odd class count: the last class shares its cells with a dummy class that has to be zero
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < nrows)
                        __stcs(reinterpret_cast<float4*>(oc + (long long)i * out.pitch), make_float4(acc[2 * i] + a[i].x, 0.f, acc[2 * i + 1] + a[i].y, 0.f));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < nrows) {
                        __stcs(oc + (long long)i * out.pitch, acc[2 * i] + a[i].x);
                        __stcs(oc + (long long)i * out.pitch + 2, acc[2 * i + 1] + a[i].y);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_u32(empty0 + 8u * stage);  // this warp no longer reads the window
        }

        sp += kLStageFloats;
        if (++stage == kLStages) {
            stage = 0;
            full_phase ^= 1u;
            sp -= (size_t)kLStages * kLStageFloats;
        }
        pp += kLPartFloats;
        if (++pb == kLParts) {
            pb = 0;
            part_phase ^= 1u;
            pp -= kLParts * kLPartFloats;
        }
        if (++c == C) {
            c = 0;
            if (nk != k) {
                k = nk;
                enter_tile(k);
            }
        }
        if (G == 1) oc = class_ptr(c);
    }
}

template <bool kFar>  // kFar: dilation 24 present ([1,2,4,8,12,24]); otherwise [1,2,4,8,12]
__global__ void __launch_bounds__(kLLaunchThreads, 1)
pamr_sweep_lattice_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wts, LatticeOut out, int C,
                          int H, int W, int tiles_x, int tiles_y, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    LatticeCtx cx;
    cx.stage0 = reinterpret_cast<float*>(smem_raw);
    cx.part = cx.stage0 + (size_t)kLStages * kLStageFloats;
    cx.full = reinterpret_cast<uint64_t*>(cx.part + kLParts * kLPartFloats);
    cx.ready = cx.full + kLStages;
    cx.empty = cx.ready + kLStages;
    cx.pfull = cx.empty + kLStages;
    cx.pempty = cx.pfull + 4 * kLParts;
    cx.wts = wts;
    cx.C = C;
    cx.H = H;
    cx.W = W;
    cx.tiles_x = tiles_x;
    cx.tiles_per_img = tiles_x * tiles_y;
    cx.n_my = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    cx.total = cx.n_my * C;
    cx.s0 = (int)(((long long)blockIdx.x * C) / gridDim.x);  // staggered class phase (worth 1.3 %, profiles/r01f_notes.md)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < kLStages; ++s) {
            mbar_init(&cx.full[s], 1);
            mbar_init(&cx.ready[s], 3);               // the three patch warps
            mbar_init(&cx.empty[s], kLThreads / 32);  // the eight compute warps
        }
        for (int s = 0; s < 4 * kLParts; ++s) {
            mbar_init(&cx.pfull[s], 1);   // warp j of group A
            mbar_init(&cx.pempty[s], 1);  // warp j of group B
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (threadIdx.x >= kLThreads) {  // producer warpgroup: hand its registers to the compute warps
        // the CTA owns 384 x 168 registers; what this warpgroup gives back ((168 - 24) x 128) is exactly what the two
        // compute warpgroups take ((240 - 168) x 256) -- a larger value here would leave their setmaxnreg.inc waiting forever
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (threadIdx.x == kLThreads) lattice_producer(cx, &tmap);
        else if (threadIdx.x >= kLThreads + 32) lattice_patcher(cx);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    if (threadIdx.x < kLGroupThreads) lattice_group<0, kFar>(cx, out);  // warp-uniform
    else lattice_group<1, kFar>(cx, out);
}

// ---------------------------------------------------------------------------------------------
// Affinity weights (reference wss/modules.py:141-145) in the lattice layout, from a replicate-padded
// image: one CTA per 32 x 32 tile, the K (<= 3) 80 x 84 channel windows arrive by TMA, each thread
// walks four pixels (lanes along x: conflict-free LDS at immediate offsets).
// ---------------------------------------------------------------------------------------------
template <int D, class DS>
__global__ void __launch_bounds__(kLThreads, 2)
pamr_weights_lattice_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ wts, int K, int tiles_x,
                            int tiles_per_img) {
    constexpr int P = 8 * D;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* win = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)3 * kLStageBytes);
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const LTile tc = ltile(blockIdx.x, tiles_x, tiles_per_img);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(K * kLStageBytes));
        for (int k = 0; k < K; ++k) tma_load_3d(win + (size_t)k * kLStageFloats, &tmap, bar, tc.x0, tc.y0, tc.b * K + k);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    float* o = wts + (size_t)blockIdx.x * kLWeightsPerTile;
    const float invK = 1.f / (float)K;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int y = wrp + 8 * i, x = lane;
        const float* sp0 = win + (y + kHalo) * kLPitch + x + kHalo;
        float logit[P];
#pragma unroll
        for (int p = 0; p < P; ++p) logit[p] = 0.f;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const float* sp = sp0 + k * kLStageFloats;
            const float c = sp[0];
            float dlt[P];  // neighbour - centre; the D centre samples of LocalStDev contribute zeros
#pragma unroll
            for (int di = 0; di < D; ++di) {
                const int d = DS::get(di);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int dy = (j < 3) ? -1 : ((j < 5) ? 0 : 1);
                    const int dx = (j < 3) ? (j - 1) : ((j == 3) ? -1 : ((j == 4) ? 1 : (j - 6)));
                    dlt[di * 8 + j] = sp[dy * d * kLPitch + dx * d] - c;
                }
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
        // group A holds dilations 4, 8, 12 (taps 16..39), group B dilations 1, 2 (taps 0..15) and 24 (taps 40..47)
        const Owner oa = owner_a(y, x), ob = owner_b(y, x);
        float4* pa = reinterpret_cast<float4*>(o) + oa.thread;                                      // + weight_group_offset<0>(g)
        float4* pb = reinterpret_cast<float4*>(o) + kLGroupsA * kLGroupThreads + ob.thread;         // + weight_group_offset<1>(g)
#pragma unroll
        for (int q = 0; q < 6; ++q)
            pa[(oa.slot * 6 + q) * kLGroupThreads] =
                make_float4(logit[16 + 4 * q] * rz, logit[17 + 4 * q] * rz, logit[18 + 4 * q] * rz, logit[19 + 4 * q] * rz);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            pb[(ob.slot * 4 + q) * kLGroupThreads] = make_float4(logit[4 * q] * rz, logit[4 * q + 1] * rz, logit[4 * q + 2] * rz, logit[4 * q + 3] * rz);
        if (D == 6) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
                pb[(kLGroupsBNear + ob.slot * 2 + q) * kLGroupThreads] =
                    make_float4(logit[P - 8 + 4 * q] * rz, logit[P - 7 + 4 * q] * rz, logit[P - 6 + 4 * q] * rz, logit[P - 5 + 4 * q] * rz);
        }
    }
}

// ------------------------------------------------------------------------------------------ host
bool sweep_lattice_applicable(int K, int H, int W, const Dilations& dil, int D) {
    if ((D != 6 && D != 5) || K < 1 || K > 3) return false;
    for (int i = 0; i < D; ++i)  // [1,2,4,8,12,24] (class default) or its first five (the trainer's set, train.py:81)
        if (dil.d[i] != DilVoc6::get(i)) return false;
    if (W % 4 != 0) return false;                   // TMA: 16-byte global row pitch; float2 stores
    if ((long long)H * W <= 64 * 64) return false;  // small maps: the fused kernel
    return true;
}

size_t lattice_weight_elems(int B, int H, int W) {
    return (size_t)B * ceil_div(H, kTile) * ceil_div(W, kTile) * (size_t)kLWeightsPerTile;
}

template <int D, class DS>
static int launch_weights_lattice_t(const CUtensorMap& tmap, float* w, int K, int tiles_x, int tiles_y, int n_tiles, cudaStream_t s) {
    auto kern = pamr_weights_lattice_kernel<D, DS>;
    const size_t smem = (size_t)3 * kLStageBytes + 64;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("pamr_weights_lattice: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    kern<<<n_tiles, kLThreads, smem, s>>>(tmap, w, K, tiles_x, tiles_x * tiles_y);
    return check_launch("pamr_weights_lattice");
}

int launch_weights_lattice(const float* padded_img, float* w, int B, int K, int H, int W, int D, cudaStream_t s) {
    CUtensorMap tmap;
    const int rc = encode_tmap_3d_f32(&tmap, padded_img, W + 2 * kHalo, H + 2 * kHalo, (long long)B * K, kLPitch, kBox);
    if (rc != 0) {
        set_error("pamr_weights_lattice: cuTensorMapEncodeTiled failed (%d)", rc);
        return CL4_ECUDA;
    }
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile);
    const int n_tiles = B * tiles_x * tiles_y;
    return D == 6 ? launch_weights_lattice_t<6, DilVoc6>(tmap, w, K, tiles_x, tiles_y, n_tiles, s)
                  : launch_weights_lattice_t<5, DilVoc5>(tmap, w, K, tiles_x, tiles_y, n_tiles, s);
}

int launch_sweep_lattice(const float* w, const float* in, int in_pitch, long long in_plane, float* out, int out_pitch,
                         long long out_plane, int out_cells, int B, int C, int H, int W, int D, cudaStream_t s) {
    CUtensorMap tmap;  // over the H x W image area of each plane; boxes may start at negative coordinates (zero fill)
    const int rc = encode_tmap_3d_f32_strided(&tmap, in, W, H, (long long)B * C, in_pitch, in_plane, kLPitch, kBox);
    if (rc != 0) {
        set_error("pamr_sweep_lattice: cuTensorMapEncodeTiled failed (%d)", rc);
        return CL4_ECUDA;
    }
    LatticeOut so;
    so.ptr = out;
    so.plane = out_plane;
    so.pitch = out_pitch;
    so.cells = out_cells ? 1 : 0;
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile);
    const int n_tiles = B * tiles_x * tiles_y;
    auto kern = D == 6 ? pamr_sweep_lattice_kernel<true> : pamr_sweep_lattice_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLSmem);
    if (e != cudaSuccess) {
        set_error("pamr_sweep_lattice: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<grid, kLLaunchThreads, kLSmem, s>>>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles);
    return check_launch("pamr_sweep_lattice");
}

}  // namespace cl4

// Layout introspection (host only, no GPU needed): which thread of which warp group holds the weights of pixel (y, x) of a
// 32 x 32 tile, and in which of its 8 pixel slots.  tests/test_abi_and_host.py checks the invariants the kernels rely on
// (a bijection per group, warp pairs owning the same rows, bank-conflict-free lane patterns).
extern "C" int cl4_lattice_owner(int group, int y, int x, int* thread_out, int* slot_out) {
    CL4_REQUIRE(group >= 0 && group <= 1 && y >= 0 && y < cl4::kTile && x >= 0 && x < cl4::kTile && thread_out && slot_out,
                CL4_EINVAL, "lattice_owner: group 0/1, pixel inside the 32 x 32 tile");
    const cl4::Owner o = group == 0 ? cl4::owner_a(y, x) : cl4::owner_b(y, x);
    *thread_out = o.thread;
    *slot_out = o.slot;
    return CL4_OK;
}
