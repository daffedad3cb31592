// The phase-1 pseudo-label step around PAMR (reference train.py:372-385) in TWO launches for feature-resolution maps
// (at most 64 x 64, D <= 6, dilations <= 24 -- the regime the trainer runs):
//
//   1. phase1_prologue_kernel, four CTAs per (image, 32 x 32 tile), each owning eight rows of the tile:
//        im = F.interpolate(denorm(images), int_masks.shape[-2:], "bilinear", align_corners=True)   train.py:376-378
//          -> evaluated straight into a replicate-padded shared-memory window (tile + 24-pixel halo), never written out;
//        affinity weights of the tile's pixels from that window (wss/modules.py:141-146)  -> tile-major scratch;
//        int_masks.softmax(dim=1) of the tile's pixels (train.py:373)                       -> scratch planes.
//   2. pamr_fused_kernel<kP1 = true> (pamr_fused.cu): all num_iter sweeps on-chip, then label gating, plane maxima,
//      thresholds and pseudo_gtmask with the ambiguity rule (train.py:382-385, wss/single_stage.py:18-40).
//
// The separate kernels of phase1.cu stay as the public pieces (and serve maps larger than 64 x 64).
#include "common.cuh"
#include "pamr_internal.cuh"
#include "pamr_weights.cuh"

namespace cl4 {

struct DenormCoef {
    float mul[3], add[3];
    int apply;
};

int launch_pamr_fused_phase1(const float* w, const float* mask_in, float* gated_out, float* pseudo_out, float* thr, int* done,
                             int* cnt, const float* labels, float cutoff_top, float cutoff_bkg, float cutoff_low, int B, int C, int H, int W,
                             int num_iter, const Dilations& dil, int D, cudaStream_t s);

constexpr int kP1Slabs = 4;                       // slabs per (image, tile): slab z owns the tile rows 8z .. 8z+7 (one row per warp);
                                                  // CTAs z < 4 do window + weights of slab z, CTAs z >= 4 the softmax of slab z - 4
constexpr int kP1SlabRows = kTile / kP1Slabs;     // 8
constexpr int kP1WinRows = kP1SlabRows + 2 * kHalo;  // 56 window rows feed a slab

template <int D, class DS>
__global__ void __launch_bounds__(256)
phase1_prologue_kernel(const float* __restrict__ images, const float* __restrict__ logits, float* __restrict__ soft,
                       float* __restrict__ wts, int* __restrict__ done, int* __restrict__ cnt, int C, int Hi, int Wi, int h, int w,
                       float sy, float sx, DenormCoef a, Dilations dil) {
    constexpr int P = 8 * D;
    extern __shared__ __align__(16) float win[];  // [3][kBox][kBox]; this CTA fills rows wy0 .. wy0+55
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int tiles_x = ceil_div(w, kTile);
    const int tyi = blockIdx.x / tiles_x;
    const int y0 = tyi * kTile, x0 = (blockIdx.x - tyi * tiles_x) * kTile;
    const int b = blockIdx.y;
    const int wy0 = (blockIdx.z % kP1Slabs) * kP1SlabRows;
    if (blockIdx.x == 0 && blockIdx.z == 0 && tid == 0) done[b] = 0;  // ticket counter of the epilogue (pamr_fused.cu)
    const size_t hw = (size_t)h * w;

    if (blockIdx.z >= kP1Slabs) {
        // ---- softmax over the classes for the slab's pixels, one per thread, in CTAs of their own: it is as long as the rest
        // of the prologue at C = 81 and independent of it
        const int y = y0 + wy0 + wrp, x = x0 + lane;
        if (y < h && x < w) {
            cnt[(size_t)b * hw + (size_t)y * w + x] = 0;  // claims per pixel, counted by the epilogue (pamr_fused.cu)
            const float* p = logits + (size_t)b * C * hw + (size_t)y * w + x;
            float* o = soft + (size_t)b * C * hw + (size_t)y * w + x;
            float mx = p[0];
#pragma unroll 8
            for (int c = 1; c < C; ++c) mx = fmaxf(mx, p[(size_t)c * hw]);
            float z = 0.f;
#pragma unroll 8
            for (int c = 0; c < C; ++c) z += expf(p[(size_t)c * hw] - mx);
#pragma unroll 8
            for (int c = 0; c < C; ++c) o[(size_t)c * hw] = expf(p[(size_t)c * hw] - mx) / z;
        }
        return;
    }

    // ---- the shrunk, denormalised image: window cell (wy, wx) = pixel (clamp(y0 + wy - 24), clamp(x0 + wx - 24)).
    // Only the DISTINCT pixels are evaluated -- the cells that lie inside the map (for a 32 x 32 map 1024 of the window's 6400
    // cells per channel) -- a warp per (channel, row), lanes along x, several rows in flight; the clamped cells are copies.
    // Arithmetic of cl4_denorm_resize_ac (phase1.cu): denorm on each of the four taps with separate roundings for the
    // multiply and the add (Tensor.mul_().add_()), then ATen's align_corners=True bilinear combination.
    const int ya = max(0, y0 + wy0 - kHalo), yb = min(h, y0 + wy0 + kP1WinRows - kHalo);  // map rows inside this slab's window
    const int xa = max(0, x0 - kHalo), xb = min(w, x0 + kBox - kHalo);
    const int ny = max(yb - ya, 0);
    // (Batching the taps of four rows before their first use was measured slower: the step is bound by L2 sectors -- a lane's
    // two x taps share one 32-byte sector of which 8 bytes are used -- not by latency.  Every CTA of a small map needs nearly
    // the whole shrunk image, so this work is repeated by the 4 slabs x tiles of an image: 58 of the prologue's 60 us at 56 x 56.)
#pragma unroll 4
    for (int r = wrp; r < 3 * ny; r += 8) {
        const int k = r / ny, y = ya + (r - k * ny);
        const float* src = images + ((size_t)b * 3 + k) * Hi * Wi;
        const float fy = __fmul_rn(sy, (float)y);
        const int y_lo = min((int)fy, Hi - 1), y_hi = y_lo + (y_lo < Hi - 1);
        const float ly1 = fy - (float)y_lo, ly0 = 1.f - ly1;
        float* dst = win + ((size_t)k * kBox + (y - y0 + kHalo)) * kBox + kHalo - x0;
        for (int x = xa + lane; x < xb; x += 32) {
            const float fx = __fmul_rn(sx, (float)x);
            const int x_lo = min((int)fx, Wi - 1), x_hi = x_lo + (x_lo < Wi - 1);
            const float lx1 = fx - (float)x_lo, lx0 = 1.f - lx1;
            auto tap = [&](int yy, int xx) {
                const float v = __ldg(src + (size_t)yy * Wi + xx);
                return a.apply ? __fadd_rn(__fmul_rn(v, a.mul[k]), a.add[k]) : v;
            };
            const float top = __fadd_rn(__fmul_rn(lx0, tap(y_lo, x_lo)), __fmul_rn(lx1, tap(y_lo, x_hi)));
            const float bot = __fadd_rn(__fmul_rn(lx0, tap(y_hi, x_lo)), __fmul_rn(lx1, tap(y_hi, x_hi)));
            dst[x] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        }
    }

    __syncthreads();
    // ---- replicate padding: every cell of the slab's 56 window rows that lies outside the map copies the clamped in-map cell
    for (int r = wrp; r < 3 * kP1WinRows; r += 8) {
        const int k = r / kP1WinRows, wy = wy0 + (r - k * kP1WinRows);
        const int sy_ = clampi(y0 + wy - kHalo, 0, h - 1) - y0 + kHalo;
        float* row = win + ((size_t)k * kBox + wy) * kBox;
        const float* srow = win + ((size_t)k * kBox + sy_) * kBox;
        for (int wx = lane; wx < kBox; wx += 32) {
            const int sx_ = clampi(x0 + wx - kHalo, 0, w - 1) - x0 + kHalo;
            if (sy_ != wy || sx_ != wx) row[wx] = srow[sx_];
        }
    }
    __syncthreads();

    // ---- affinity weights of the slab's row of this warp, tile-major [tile][P/4][32][32] float4 (what pamr_fused_kernel reads)
    float4* o = reinterpret_cast<float4*>(wts) + ((size_t)b * gridDim.x + blockIdx.x) * (P / 4 * kTile * kTile) + lane;
    const int row = wy0 + wrp;
    float logit[P];
    pixel_affinity<D, DS>(win + (row + kHalo) * kBox + lane + kHalo, 3, kBox * kBox, kBox, dil, logit);
#pragma unroll
    for (int g = 0; g < P / 4; ++g)
        o[(size_t)g * (kTile * kTile) + row * kTile] = make_float4(logit[4 * g], logit[4 * g + 1], logit[4 * g + 2], logit[4 * g + 3]);
}

template <int D, class DS>
static int launch_prologue_one(const float* images, const float* logits, float* soft, float* wts, int* done, int* cnt, int B, int C, int Hi,
                               int Wi, int h, int w, const DenormCoef& a, const Dilations& dil, cudaStream_t s) {
    auto kern = phase1_prologue_kernel<D, DS>;
    const size_t smem = sizeof(float) * 3 * kBox * kBox;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("phase1_prologue: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    const float sy = (h > 1) ? (float)(Hi - 1) / (float)(h - 1) : 0.f;
    const float sx = (w > 1) ? (float)(Wi - 1) / (float)(w - 1) : 0.f;
    dim3 grid(ceil_div(w, kTile) * ceil_div(h, kTile), B, 2 * kP1Slabs);
    kern<<<grid, 256, smem, s>>>(images, logits, soft, wts, done, cnt, C, Hi, Wi, h, w, sy, sx, a, dil);
    return check_launch("phase1_prologue");
}

template <int D>
static int launch_prologue_D(const float* images, const float* logits, float* soft, float* wts, int* done, int* cnt, int B, int C, int Hi,
                             int Wi, int h, int w, const DenormCoef& a, const Dilations& dil, cudaStream_t s) {
    bool voc6 = (D == 6), voc5 = (D == 5);
    for (int i = 0; i < D && i < 6; ++i) {
        voc6 = voc6 && dil.d[i] == DilVoc6::get(i);
        voc5 = voc5 && dil.d[i] == DilVoc5::get(i);
    }
    if (D == 6 && voc6) return launch_prologue_one<6, DilVoc6>(images, logits, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s);
    if (D == 5 && voc5) return launch_prologue_one<5, DilVoc5>(images, logits, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s);
    return launch_prologue_one<D, DilRuntime>(images, logits, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s);
}

struct Phase1Layout {
    size_t wts, soft, thr, done, cnt, bytes;
};
static Phase1Layout phase1_layout(int B, int C, int h, int w, int D) {
    Phase1Layout L;
    size_t off = 0;
    auto take = [&](size_t n) {
        const size_t o = off;
        off += align_up(n, 256);
        return o;
    };
    const size_t tiles = (size_t)ceil_div(w, kTile) * ceil_div(h, kTile);
    L.wts = take(sizeof(float) * (size_t)B * tiles * 8 * D * kTile * kTile);
    L.soft = take(sizeof(float) * (size_t)B * C * h * w);
    L.thr = take(sizeof(float) * (size_t)B * C);
    L.done = take(sizeof(int) * (size_t)B);
    L.cnt = take(sizeof(int) * (size_t)B * h * w);
    L.bytes = off;
    return L;
}

}  // namespace cl4

extern "C" size_t cl4_phase1_scratch_bytes(int B, int C, int h, int w, int D) {
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || D <= 0) return 0;
    return cl4::phase1_layout(B, C, h, w, D).bytes;
}

extern "C" int cl4_phase1_pseudo_labels(const float* images, const float* int_masks, const float* l1h, const float* mean,
                                        const float* std, const int* dilations, int D, int num_iter, float cutoff_top,
                                        float cutoff_bkg, float cutoff_low, float* soft_out, float* pseudo_out, void* scratch,
                                        size_t scratch_bytes, int B, int C, int Hi, int Wi, int h, int w, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 1 && Hi > 0 && Wi > 0 && h > 0 && w > 0, CL4_EINVAL, "phase1_pseudo_labels: bad shape");
    CL4_REQUIRE(B <= 65535, CL4_EUNSUPPORTED, "phase1_pseudo_labels: batch > 65535");
    CL4_REQUIRE(num_iter >= 1, CL4_EUNSUPPORTED, "phase1_pseudo_labels: num_iter must be >= 1 on the fused path");
    CL4_REQUIRE((mean == nullptr) == (std == nullptr), CL4_EINVAL, "phase1_pseudo_labels: mean and std go together");
    CL4_REQUIRE(dilations && D >= 1 && D <= 6, CL4_EUNSUPPORTED, "phase1_pseudo_labels: 1..6 dilations on the fused path");
    Dilations dil;
    for (int i = 0; i < CL4_MAX_DILATIONS; ++i) dil.d[i] = 1;
    for (int i = 0; i < D; ++i) {
        CL4_REQUIRE(dilations[i] >= 1, CL4_EINVAL, "phase1_pseudo_labels: dilation %d must be >= 1", dilations[i]);
        dil.d[i] = dilations[i];
    }
    CL4_REQUIRE(pamr_fused_applicable(h, w, dil, D), CL4_EUNSUPPORTED,
                "phase1_pseudo_labels: fused path needs maps <= 64 x 64 and dilations <= 24 (use the separate entry points)");
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(images && int_masks && soft_out && pseudo_out, CL4_EINVAL, "phase1_pseudo_labels: null pointer");
    const Phase1Layout L = phase1_layout(B, C, h, w, D);
    CL4_REQUIRE(scratch && scratch_bytes >= L.bytes, CL4_ESCRATCH, "phase1_pseudo_labels: scratch too small");
    char* base = reinterpret_cast<char*>(scratch);
    float* wts = reinterpret_cast<float*>(base + L.wts);
    float* soft = reinterpret_cast<float*>(base + L.soft);
    float* thr = reinterpret_cast<float*>(base + L.thr);
    int* done = reinterpret_cast<int*>(base + L.done);
    int* cnt = reinterpret_cast<int*>(base + L.cnt);
    DenormCoef a{{1.f, 1.f, 1.f}, {0.f, 0.f, 0.f}, mean ? 1 : 0};
    if (mean)
        for (int k = 0; k < 3; ++k) {
            a.mul[k] = std[k];
            a.add[k] = mean[k];
        }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = CL4_EUNSUPPORTED;
    switch (D) {
        case 1: rc = launch_prologue_D<1>(images, int_masks, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s); break;
        case 2: rc = launch_prologue_D<2>(images, int_masks, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s); break;
        case 3: rc = launch_prologue_D<3>(images, int_masks, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s); break;
        case 4: rc = launch_prologue_D<4>(images, int_masks, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s); break;
        case 5: rc = launch_prologue_D<5>(images, int_masks, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s); break;
        case 6: rc = launch_prologue_D<6>(images, int_masks, soft, wts, done, cnt, B, C, Hi, Wi, h, w, a, dil, s); break;
    }
    if (rc != CL4_OK) return rc;
    return launch_pamr_fused_phase1(wts, soft, soft_out, pseudo_out, thr, done, cnt, l1h, cutoff_top, cutoff_bkg, cutoff_low, B, C, h, w,
                                    num_iter, dil, D, s);
}
