// PAMR propagation with EVERY iteration on-chip, for maps of at most 64 x 64 pixels — the
// resolution PAMR really runs at inside the trainer (32 x 32 at crop 512 / stride 16, 56 x 56 for
// coco-voc: train.py:376-379).  Reference: wss/modules.py:147-149, num_iter times.
//
// One CTA owns `cpb` classes of one image for all iterations.  Classes are independent in PAMR
// (they only share the affinity weights), so CTAs never communicate.  Per class the CTA keeps two
// replicate-padded planes (ping-pong) in shared memory: the clamped neighbour of wss/modules.py:57
// is an ordinary in-bounds LDS at an immediate offset, as in the HBM-resident 4-pixel sweep (pamr_tma.cu),
// whose tile-major weight layout this kernel reads:
//   * a thread owns 2 pixels of one column, 4 rows apart, of a 32 x 32 tile (16 warps per CTA) and keeps their
//     2 x 8D weights in registers for all classes and — when the map is one tile — all iterations;
//   * maps of 2-4 tiles re-read the tile-major weights from L2 at each tile switch, software-
//     pipelined into the last class pass of the previous tile;
//   * the frame around the map is 12 pixels when no dilation exceeds 12 (the trainer's set, train.py:81), else 24;
//     after each iteration the frame of the new planes is rewritten from their interior.
// HBM traffic is the minimum: masks in once, masks out once, weights once per CTA (L2 hits after the
// first CTA of an image).  The bound is the shared-memory pipe (one LDS.32 wavefront per 32 FMAs; ptxas shares the ~10 of 80
// sources per class that the two pixels have in common) and, on multi-tile maps, the weight re-reads from L2
// (profiles/r02_notes.md sections 11, 12; profiles/r03c_fused_ncu_summary.txt).
#include "common.cuh"
#include "pamr_internal.cuh"
#include "pamr_sweep.cuh"

namespace cl4 {

constexpr int kFusedMaxDim = 64;

// Phase-1 epilogue (kP1): what the trainer does with PAMR's result (train.py:382-385) happens on the values of the last
// iteration while they are still in registers:
//   int_masks_soft[:, 1:] *= l1h[:, :, None, None]      -> gated values go to `mask_out`
//   pseudo_gtmask (wss/single_stage.py:18-40): plane maximum (NaN-propagating), x cutoff_bkg / cutoff_top, floored at
//   cutoff_low -> thr[b, c]; pseudo = (mask > thr), pixels claimed by more than one class cleared.
// The ambiguity rule couples all classes of a pixel, and the classes of an image are spread over several CTAs (not
// necessarily co-resident): every CTA writes the tentative pseudo labels of its own planes (mask > thr) and counts its claims
// per pixel in cnt[b] (atomicAdd), then takes a ticket from done[b]; the CTA that draws the last ticket of its image clears the
// pixels claimed more than once -- O(H x W) reads plus C stores per ambiguous pixel instead of a pass over all C x H x W
// values by one CTA (which cost 260 us at C = 81, 56 x 56).
struct Phase1Epilogue {
    const float* labels;  // [B, C-1] or nullptr
    float* pseudo;        // [B, C, H, W]
    float* thr;           // [B, C]
    int* done;            // [B], zeroed by the prologue kernel
    int* cnt;             // [B, H, W] classes claiming each pixel, zeroed by the prologue kernel
    float cutoff_top, cutoff_bkg, cutoff_low;
};

// Pixels per thread.  The 4-pixel mapping of the HBM-resident sweep (8 warps per CTA) leaves an on-chip CTA latency-bound:
// its four pixels share no source, so two pixels per thread on 16 warps cost the same LDS and FFMA and hide twice the latency.
#ifndef CL4_FUSED_PX
#define CL4_FUSED_PX 2
#endif
constexpr int kFusedPx = CL4_FUSED_PX;                     // rows ty + 4*i, i < kFusedPx
constexpr int kFusedThreads = (kTile / kFusedPx) * 32;     // 512 (256 with four pixels per thread)
static_assert(kFusedPx == 2 || kFusedPx == 4, "pixels per thread");

// One class of one tile (sweep_class of pamr_sweep.cuh with kFusedPx pixels per thread).  kReload: refill the weight registers with
// the next tile's weights right after their last use.
template <int D, class DS, bool kReload, int PITCH>
__device__ __forceinline__ void fused_class(float (&w)[kFusedPx][8 * D], const float* __restrict__ sp, const Dilations& dil,
                                            const float4* __restrict__ nw, float (&acc)[kFusedPx]) {
#pragma unroll
    for (int i = 0; i < kFusedPx; ++i) acc[i] = 0.f;
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
            for (int i = 0; i < kFusedPx; ++i) acc[i] = fmaf(w[i][p], sp[off + i * kRowGap * PITCH], acc[i]);
        }
        if (kReload) {
#pragma unroll
            for (int i = 0; i < kFusedPx; ++i) {
                const float4 v = __ldg(nw + g * (kTile * kTile) + i * kRowGap * kTile);
                w[i][4 * g + 0] = v.x;
                w[i][4 * g + 1] = v.y;
                w[i][4 * g + 2] = v.z;
                w[i][4 * g + 3] = v.w;
            }
        }
    }
}

// plane pitch = map extent (one or two tiles) + 2 x frame
constexpr int kFusedHaloSmall = 12;
__host__ __device__ constexpr int fused_pitch(int tiles, int halo) { return tiles * kTile + 2 * halo; }
__host__ __device__ constexpr int fused_halo(int pitch) {
    return (pitch == fused_pitch(1, kFusedHaloSmall) || pitch == fused_pitch(2, kFusedHaloSmall)) ? kFusedHaloSmall : kHalo;
}

template <int D, class DS, int PITCH, bool kP1>
__global__ void __launch_bounds__(kFusedThreads, 1)
pamr_fused_kernel(const float* __restrict__ wts, const float* __restrict__ mask_in, float* __restrict__ mask_out,
                  int C, int H, int W, int cpb, int num_iter, Dilations dil, Phase1Epilogue ep) {
    constexpr int P = 8 * D;
    constexpr int kPlane = PITCH * PITCH;  // PITCH rows of PITCH floats
    constexpr int HALO = fused_halo(PITCH);  // replicate frame around the map: 24, or 12 when no dilation exceeds 12
    extern __shared__ __align__(16) float smem[];

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int tx = lane;
    const int ty = (wrp >> 2) * (kRowGap * kFusedPx) + (wrp & 3);  // rows ty + 4*i, i < kFusedPx
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * cpb;
    const int nc = min(cpb, C - c0);
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile), n_tiles = tiles_x * tiles_y;
    const int Hp = H + 2 * HALO, Wp = W + 2 * HALO;
    const size_t HW = (size_t)H * W;

    // planes: [class][ping-pong][PITCH][PITCH]
    auto plane = [&](int c, int which) -> float* { return smem + (size_t)(c * 2 + which) * kPlane; };

    // ---- stage the input masks as replicate-padded planes (a warp per padded row, lanes over columns:
    // no integer division in these loops)
    for (int c = 0; c < nc; ++c) {
        const float* src = mask_in + ((size_t)b * C + c0 + c) * HW;
        float* dst = plane(c, 0);
        for (int py = wrp; py < Hp; py += kFusedThreads / 32) {
            const float* srow = src + (size_t)clampi(py - HALO, 0, H - 1) * W;
            float* drow = dst + py * PITCH;
            for (int px = lane; px < Wp; px += 32) drow[px] = __ldg(srow + clampi(px - HALO, 0, W - 1));
        }
    }

    // tile-major weights of this image: tile t holds [P/4][32][32] float4
    auto weight_ptr = [&](int t) -> const float4* {
        return reinterpret_cast<const float4*>(wts) + ((size_t)b * n_tiles + t) * (P / 4 * kTile * kTile) + ty * kTile + tx;
    };
    float w[kFusedPx][P];
    {
        const float4* wp = weight_ptr(0);
#pragma unroll
        for (int g = 0; g < P / 4; ++g)
#pragma unroll
            for (int i = 0; i < kFusedPx; ++i) {
                const float4 v = __ldg(wp + g * (kTile * kTile) + i * kRowGap * kTile);
                w[i][4 * g + 0] = v.x;
                w[i][4 * g + 1] = v.y;
                w[i][4 * g + 2] = v.z;
                w[i][4 * g + 3] = v.w;
            }
    }
    __syncthreads();

    constexpr int kCpbCap = (227 * 1024) / (2 * kPlane * 4);   // classes per CTA that fit (launch_fused_one)
    __shared__ float s_cmax[kP1 ? kCpbCap * (kFusedThreads / 32) : 1];  // running plane maxima, one slot per (class, warp)
    __shared__ int s_last;
    if (kP1)
        for (int i = tid; i < kCpbCap * (kFusedThreads / 32); i += kFusedThreads) s_cmax[i] = -INFINITY;

    float acc[kFusedPx];
    for (int it = 0; it < num_iter; ++it) {
        const int cur = it & 1, nxt = cur ^ 1;
        const bool last_it = (it == num_iter - 1);
        for (int t = 0; t < n_tiles; ++t) {
            const int tyi = t / tiles_x;
            const int y0 = tyi * kTile, x0 = (t - tyi * tiles_x) * kTile;
            const int sbase = (y0 + ty + HALO) * PITCH + (x0 + tx + HALO);
            const int x = x0 + tx;
            unsigned valid = 0u;
#pragma unroll
            for (int i = 0; i < kFusedPx; ++i)
                if (x < W && y0 + ty + i * kRowGap < H) valid |= 1u << i;
            // weights of the tile that follows (this iteration's next tile, or tile 0 of the next iteration)
            const int tn = (t + 1 == n_tiles) ? 0 : t + 1;
            const bool more = (n_tiles > 1) && !(last_it && t + 1 == n_tiles);
            for (int c = 0; c < nc; ++c) {
                const float* sp = plane(c, cur) + sbase;
                if (more && c == nc - 1)
                    fused_class<D, DS, true, PITCH>(w, sp, dil, weight_ptr(tn), acc);
                else
                    fused_class<D, DS, false, PITCH>(w, sp, dil, nullptr, acc);
                if (last_it) {
                    float* o = mask_out + ((size_t)b * C + c0 + c) * HW + (size_t)(y0 + ty) * W + x;
                    if (kP1) {
                        const int cls = c0 + c;
                        const float g = (ep.labels && cls > 0) ? __ldg(ep.labels + (size_t)b * (C - 1) + (cls - 1)) : 1.f;
                        float m = -INFINITY;
#pragma unroll
                        for (int i = 0; i < kFusedPx; ++i) {
                            if (ep.labels && cls > 0) acc[i] = __fmul_rn(acc[i], g);
                            if ((valid >> i) & 1u) m = nanmax(m, acc[i]);  // torch.max propagates NaN
                        }
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) m = nanmax(m, __shfl_xor_sync(0xffffffffu, m, sft));
                        if (lane == 0) s_cmax[c * (kFusedThreads / 32) + wrp] = nanmax(s_cmax[c * (kFusedThreads / 32) + wrp], m);
                    }
#pragma unroll
                    for (int i = 0; i < kFusedPx; ++i)
                        if ((valid >> i) & 1u) o[(size_t)(i * kRowGap) * W] = acc[i];
                    if (kP1) {  // the epilogue compares the gated values with the plane's threshold: keep them on-chip (the other
                                // ping-pong plane is free in the last iteration)
                        float* o2 = plane(c, nxt) + sbase;
#pragma unroll
                        for (int i = 0; i < kFusedPx; ++i)
                            if ((valid >> i) & 1u) o2[i * kRowGap * PITCH] = acc[i];
                    }
                } else {
                    float* o = plane(c, nxt) + sbase;
#pragma unroll
                    for (int i = 0; i < kFusedPx; ++i)
                        if ((valid >> i) & 1u) o[i * kRowGap * PITCH] = acc[i];
                }
            }
        }
        if (last_it) break;
        __syncthreads();  // interiors of the new planes complete
        // replicate frame of the new planes (wss/modules.py:57 for the next iteration): frame cells
        // read interior cells only, so no ordering between threads is needed
        for (int c = 0; c < nc; ++c) {
            float* pl = plane(c, nxt);
            // rows above and below the image: full padded width
            for (int r = wrp; r < 2 * HALO; r += kFusedThreads / 32) {
                const int py = (r < HALO) ? r : (H + r);
                const float* srow = pl + ((r < HALO) ? HALO : (H + HALO - 1)) * PITCH;
                float* drow = pl + py * PITCH;
                for (int px = lane; px < Wp; px += 32) drow[px] = srow[clampi(px, HALO, W + HALO - 1)];
            }
            // left and right bands of the image rows: lanes 0..23 left, the next 24 right
            for (int r = wrp; r < H; r += kFusedThreads / 32) {
                float* row = pl + (r + HALO) * PITCH;
                const float vl = row[HALO], vr = row[W + HALO - 1];
                if (lane < HALO) {
                    row[lane] = vl;
                    row[W + HALO + lane] = vr;
                }
            }
        }
        __syncthreads();
    }

    if (kP1) {
        __shared__ float s_thr[kCpbCap];
        __syncthreads();  // s_cmax complete; this CTA's gated planes are visible to the whole CTA
        if (tid < nc) {
            float mx = s_cmax[tid * (kFusedThreads / 32)];
            for (int wi = 1; wi < kFusedThreads / 32; ++wi) mx = nanmax(mx, s_cmax[tid * (kFusedThreads / 32) + wi]);
            const int cls = c0 + tid;
            const float scaled = __fmul_rn(mx, cls == 0 ? ep.cutoff_bkg : ep.cutoff_top);  // mask_max[:, :1] *= bkg; [:, 1:] *= top
            const float th = (ep.cutoff_low > scaled || ep.cutoff_low != ep.cutoff_low) ? ep.cutoff_low : scaled;
            s_thr[tid] = th;
            ep.thr[(size_t)b * C + cls] = th;
        }
        __syncthreads();
        int* cnt = ep.cnt + (size_t)b * HW;
        const int fin = (num_iter & 1);  // the plane the last iteration wrote: nxt of it = num_iter - 1
        for (int c = 0; c < nc; ++c) {
            const float* gm = plane(c, fin) + HALO * PITCH + HALO;
            float* po = ep.pseudo + ((size_t)b * C + c0 + c) * HW;
            const float th = s_thr[c];
            for (int y = wrp; y < H; y += kFusedThreads / 32)
                for (int x = lane; x < W; x += 32) {
                    const bool on = gm[y * PITCH + x] > th;
                    po[(size_t)y * W + x] = on ? 1.f : 0.f;
                    if (on) atomicAdd(cnt + y * W + x, 1);
                }
        }
        __threadfence();  // every thread: its pseudo labels and claims are visible device-wide before the ticket is drawn
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(ep.done + b, 1) == (int)gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            float* po = ep.pseudo + (size_t)b * C * HW;
            for (int i = tid; i < (int)HW; i += kFusedThreads)
                if (__ldcg(cnt + i) > 1)  // ambiguous=True (train.py:384): a pixel claimed by several classes belongs to none
                    for (int c = 0; c < C; ++c) po[(size_t)c * HW + i] = 0.f;
        }
    }
}

bool pamr_fused_applicable(int H, int W, const Dilations& dil, int D) {
    if (D > 6 || H > kFusedMaxDim || W > kFusedMaxDim) return false;
    for (int i = 0; i < D; ++i)
        if (dil.d[i] > kHalo) return false;
    return true;
}

// Classes per CTA: the fewest waves x classes per CTA.  On maps of more than one tile a CTA re-reads a tile's weights from L2 at
// every tile switch, i.e. once per (tile, iteration) whatever its class count, and only part of that hides behind the last class
// pass: measured as 0.45 of a class pass (B16, 56 x 56: C = 81 667 -> 629 us with two classes per CTA, C = 21 233 -> 260 us;
// gpurun_out/r03j_cpb.log), so a second class per CTA pays when it costs at most that many extra waves.
static int pick_cpb(int B, int C, int cpb_max, bool multi_tile) {
    int best = 1;
    long long best_cost = -1;
    for (int cpb = 1; cpb <= cpb_max && cpb <= C; ++cpb) {
        const long long ctas = (long long)B * ceil_div(C, cpb);
        const long long cost = ((ctas + kNumSMs - 1) / kNumSMs) * (100 * cpb + (multi_tile ? 45 : 0));
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && cpb > best)) {
            best = cpb;
            best_cost = cost;
        }
    }
    return best;
}

// set by launch_pamr_fused_phase1 around the dispatch below (same host thread)
static thread_local const Phase1Epilogue* t_ep = nullptr;

template <int D, class DS, int PITCH, bool kP1>
static int launch_fused_one_ep(const float* w, const float* mi, float* mo, int B, int C, int H, int W, int num_iter,
                               const Dilations& dil, cudaStream_t s);

template <int D, class DS, int PITCH>
static int launch_fused_one(const float* w, const float* mi, float* mo, int B, int C, int H, int W, int num_iter,
                            const Dilations& dil, cudaStream_t s) {
    if (t_ep) return launch_fused_one_ep<D, DS, PITCH, true>(w, mi, mo, B, C, H, W, num_iter, dil, s);
    return launch_fused_one_ep<D, DS, PITCH, false>(w, mi, mo, B, C, H, W, num_iter, dil, s);
}

template <int D, class DS, int PITCH, bool kP1>
static int launch_fused_one_ep(const float* w, const float* mi, float* mo, int B, int C, int H, int W, int num_iter,
                               const Dilations& dil, cudaStream_t s) {
    auto kern = pamr_fused_kernel<D, DS, PITCH, kP1>;
    constexpr size_t kPlaneBytes = (size_t)PITCH * PITCH * 4;
    constexpr int kCpbMax = (int)((227 * 1024) / (2 * kPlaneBytes));
    static_assert(kCpbMax >= 1, "plane too large for shared memory");
    const int cpb = pick_cpb(B, C, kCpbMax, H > kTile || W > kTile);
    const size_t smem = (size_t)cpb * 2 * kPlaneBytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device
    if (e != cudaSuccess) {
        set_error("pamr_fused: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    dim3 grid(ceil_div(C, cpb), B);
    kern<<<grid, kFusedThreads, smem, s>>>(w, mi, mo, C, H, W, cpb, num_iter, dil, kP1 ? *t_ep : Phase1Epilogue{});
    return check_launch("pamr_fused");
}

template <int D, int PITCH>
static int launch_fused_D(const float* w, const float* mi, float* mo, int B, int C, int H, int W, int num_iter,
                          const Dilations& dil, cudaStream_t s) {
    bool voc6 = (D == 6), voc5 = (D == 5);
    for (int i = 0; i < D && i < 6; ++i) {
        voc6 = voc6 && dil.d[i] == DilVoc6::get(i);
        voc5 = voc5 && dil.d[i] == DilVoc5::get(i);
    }
    if (D == 6 && voc6) return launch_fused_one<6, DilVoc6, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
    if (D == 5 && voc5) return launch_fused_one<5, DilVoc5, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
    return launch_fused_one<D, DilRuntime, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
}

template <int PITCH>
static int launch_fused_P(const float* w, const float* mi, float* mo, int B, int C, int H, int W, int num_iter,
                          const Dilations& dil, int D, cudaStream_t s) {
    switch (D) {
        case 1: return launch_fused_D<1, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
        case 2: return launch_fused_D<2, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
        case 3: return launch_fused_D<3, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
        case 4: return launch_fused_D<4, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
        case 5: return launch_fused_D<5, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
        case 6: return launch_fused_D<6, PITCH>(w, mi, mo, B, C, H, W, num_iter, dil, s);
    }
    set_error("pamr_fused: bad D=%d", D);
    return CL4_EUNSUPPORTED;
}

// w: tile-major weights; mask_in / mask_out: plain [B*C][H][W]; num_iter >= 1
int launch_pamr_fused(const float* w, const float* mask_in, float* mask_out, int B, int C, int H, int W, int num_iter,
                      const Dilations& dil, int D, cudaStream_t s) {
    // the trainer's dilation set [1,2,4,8,12] (train.py:81) needs a frame of 12, not 24: planes of 56^2 / 88^2 instead of 80^2 / 112^2
    // floats -- less frame to rewrite per iteration and room for three classes per CTA on two-tile maps
    bool small = true;
    for (int i = 0; i < D; ++i) small = small && dil.d[i] <= kFusedHaloSmall;
    const bool one = H <= kTile && W <= kTile;
    if (small) {
        if (one) return launch_fused_P<fused_pitch(1, kFusedHaloSmall)>(w, mask_in, mask_out, B, C, H, W, num_iter, dil, D, s);
        return launch_fused_P<fused_pitch(2, kFusedHaloSmall)>(w, mask_in, mask_out, B, C, H, W, num_iter, dil, D, s);
    }
    if (one) return launch_fused_P<fused_pitch(1, kHalo)>(w, mask_in, mask_out, B, C, H, W, num_iter, dil, D, s);
    return launch_fused_P<fused_pitch(2, kHalo)>(w, mask_in, mask_out, B, C, H, W, num_iter, dil, D, s);
}

// The same launch with the phase-1 epilogue: mask_out receives the label-gated masks, pseudo / thr as Phase1Epilogue says.
int launch_pamr_fused_phase1(const float* w, const float* mask_in, float* gated_out, float* pseudo_out, float* thr, int* done,
                             int* cnt, const float* labels, float cutoff_top, float cutoff_bkg, float cutoff_low, int B, int C, int H, int W,
                             int num_iter, const Dilations& dil, int D, cudaStream_t s) {
    Phase1Epilogue ep{labels, pseudo_out, thr, done, cnt, cutoff_top, cutoff_bkg, cutoff_low};
    t_ep = &ep;
    const int rc = launch_pamr_fused(w, mask_in, gated_out, B, C, H, W, num_iter, dil, D, s);
    t_ep = nullptr;
    return rc;
}

}  // namespace cl4
