// group_pixels: per-pixel nearest-centre argmin (reference modules/utils.py:505-542),
// batched, with the foreground multiply of get_instance_segmentation (:606) folded in.
//
// HBM-bound for the usual handful of centres (12 B in, 8 B out per pixel) and
// ALU-bound for dense scenes (Kc = 200): centres sit in shared memory as float2
// and are read with warp-broadcast LDS.64; each thread owns 4 consecutive pixels
// so offsets/ids move as 128-bit transactions.
//
// Bit-exactness: ATen's CPU norm over the 2-vector is
//   sqrt_rn(fma_rn(dx, dx, rn(dy*dy)))   with dy = float(cy) - (float(y) + off_y)
// and argmin keeps the FIRST minimum.  sqrt_rn is monotone but not injective, so
// squared distances only pre-filter; the decision is made on the rounded sqrt.
#include "common.cuh"

namespace cl4 {

constexpr int kGroupThreads = 256;
constexpr int kGroupVec = 4;
constexpr int kCtrChunk = 2048;  // centres per shared-memory chunk (16 KB)

template <bool kVec>
__global__ void __launch_bounds__(kGroupThreads)
group_pixels_kernel(const long long* __restrict__ ctr, const int* __restrict__ count_dev, int Kc_fixed,
                    int ctr_stride, const float* __restrict__ offsets, const unsigned char* __restrict__ fg,
                    long long* __restrict__ ids, int H, int W, int empty_mode) {
    __shared__ float2 s_ctr[kCtrChunk];
    const int n = blockIdx.y;
    const int HW = H * W;
    int Kc = count_dev ? min(count_dev[n], ctr_stride) : Kc_fixed;
    const long long* c_n = ctr + (size_t)n * ctr_stride * 2;
    const float* off_y = offsets + (size_t)n * 2 * HW;
    const float* off_x = off_y + HW;
    const unsigned char* fg_n = fg ? fg + (size_t)n * HW : nullptr;
    long long* ids_n = ids + (size_t)n * HW;

    const int base = (blockIdx.x * kGroupThreads + threadIdx.x) * kGroupVec;

    float ly[kGroupVec], lx[kGroupVec], best_r2[kGroupVec], best_d[kGroupVec];
    int best_k[kGroupVec];
    unsigned char keep[kGroupVec];

    if (kVec) {
        if (base < HW) {
            const float4 oy = __ldcs(reinterpret_cast<const float4*>(off_y + base));
            const float4 ox = __ldcs(reinterpret_cast<const float4*>(off_x + base));
            const float oys[4] = {oy.x, oy.y, oy.z, oy.w};
            const float oxs[4] = {ox.x, ox.y, ox.z, ox.w};
            uchar4 f = make_uchar4(1, 1, 1, 1);
            if (fg_n) f = *reinterpret_cast<const uchar4*>(fg_n + base);
            const unsigned char fs[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int v = 0; v < kGroupVec; ++v) {
                const int i = base + v;
                const int y = i / W, x = i - y * W;
                ly[v] = __fadd_rn((float)y, oys[v]);
                lx[v] = __fadd_rn((float)x, oxs[v]);
                keep[v] = fs[v];
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < kGroupVec; ++v) {
            const int i = base + v;
            if (i < HW) {
                const int y = i / W, x = i - y * W;
                ly[v] = __fadd_rn((float)y, off_y[i]);
                lx[v] = __fadd_rn((float)x, off_x[i]);
                keep[v] = fg_n ? fg_n[i] : 1;
            } else {
                ly[v] = lx[v] = 0.f;
                keep[v] = 0;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < kGroupVec; ++v) {
        best_r2[v] = 0.f;
        best_d[v] = 0.f;
        best_k[v] = 0;
    }

    for (int k0 = 0; k0 < Kc; k0 += kCtrChunk) {
        const int kn = min(kCtrChunk, Kc - k0);
        __syncthreads();
        for (int k = threadIdx.x; k < kn; k += kGroupThreads) {
            const longlong2 c = *reinterpret_cast<const longlong2*>(c_n + 2 * (size_t)(k0 + k));
            s_ctr[k] = make_float2((float)c.x, (float)c.y);  // (cy, cx): int64 -> fp32 as torch promotes
        }
        __syncthreads();
        if (base < HW) {
            int k = 0;
            if (k0 == 0) {  // seed with centre 0 so that ties and NaNs resolve to the first index
                const float2 c = s_ctr[0];
#pragma unroll
                for (int v = 0; v < kGroupVec; ++v) {
                    const float dy = __fsub_rn(c.x, ly[v]), dx = __fsub_rn(c.y, lx[v]);
                    best_r2[v] = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
                    best_d[v] = __fsqrt_rn(best_r2[v]);
                    best_k[v] = 0;
                }
                k = 1;
            }
#pragma unroll 4
            for (; k < kn; ++k) {
                const float2 c = s_ctr[k];
#pragma unroll
                for (int v = 0; v < kGroupVec; ++v) {
                    const float dy = __fsub_rn(c.x, ly[v]), dx = __fsub_rn(c.y, lx[v]);
                    const float r2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
                    if (r2 < best_r2[v]) {  // necessary for sqrt(r2) < best_d
                        const float d = __fsqrt_rn(r2);
                        if (d < best_d[v]) {
                            best_d[v] = d;
                            best_r2[v] = r2;
                            best_k[v] = k0 + k;
                        }
                    }
                }
            }
        }
    }

    if (base >= HW) return;
    long long out[kGroupVec];
#pragma unroll
    for (int v = 0; v < kGroupVec; ++v) {
        if (Kc > 0) out[v] = keep[v] ? (long long)(best_k[v] + 1) : 0ll;
        else out[v] = (empty_mode == 1 && keep[v]) ? 1ll : 0ll;
    }
    if (kVec) {
        __stcs(reinterpret_cast<longlong2*>(ids_n + base), make_longlong2(out[0], out[1]));
        __stcs(reinterpret_cast<longlong2*>(ids_n + base + 2), make_longlong2(out[2], out[3]));
    } else {
#pragma unroll
        for (int v = 0; v < kGroupVec; ++v)
            if (base + v < HW) ids_n[base + v] = out[v];
    }
}

// refine_label_generation_with_point (reference modules/utils.py:388-460): one thread per pixel.  The pixel's class is its
// gt label - 1; its centres are that class's kept points (the reference drops points with y == 0 or x == 0, :436), the
// nearest one by group_pixels' arithmetic (first minimum) gives the offset (:458-459), every such pixel gets weight 1.
// Pixels of invalid classes (:431) or of classes without a kept point (:439) stay zero.
__global__ void __launch_bounds__(256)
refine_with_point_kernel(const long long* __restrict__ gt, const float* __restrict__ label, const long long* __restrict__ pts,
                         const unsigned char* __restrict__ keep, const float* __restrict__ offsets, float* __restrict__ out_off,
                         float* __restrict__ out_w, int C, int M, int H, int W) {
    const int b = blockIdx.y, HW = H * W;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= HW) return;
    float oy = 0.f, ox = 0.f, wgt = 0.f;
    const long long g = gt[(size_t)b * HW + i];
    if (g >= 1 && g <= C && label[(size_t)b * C + (g - 1)] != 0.f) {
        const int cls = (int)g - 1;
        const long long* p = pts + ((size_t)b * C + cls) * M * 2;
        const unsigned char* kp = keep + ((size_t)b * C + cls) * M;
        const int y = i / W, x = i - y * W;
        const float ly = __fadd_rn((float)y, offsets[(size_t)b * 2 * HW + i]);
        const float lx = __fadd_rn((float)x, offsets[((size_t)b * 2 + 1) * HW + i]);
        float best = 0.f, by = 0.f, bx = 0.f;
        bool any = false;
        for (int m = 0; m < M; ++m) {
            if (!kp[m]) continue;
            const float cy = (float)p[2 * m], cx = (float)p[2 * m + 1];
            const float dy = __fsub_rn(cy, ly), dx = __fsub_rn(cx, lx);
            const float d = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            if (!any || d < best) {  // first minimum wins (and a NaN distance never replaces the first point)
                best = d;
                by = cy;
                bx = cx;
                any = true;
            }
        }
        if (any) {
            oy = __fsub_rn(by, (float)y);
            ox = __fsub_rn(bx, (float)x);
            wgt = 1.f;
        }
    }
    out_off[(size_t)b * 2 * HW + i] = oy;
    out_off[((size_t)b * 2 + 1) * HW + i] = ox;
    out_w[(size_t)b * HW + i] = wgt;
}

}  // namespace cl4

extern "C" int cl4_refine_labels_with_point(const long long* gt_seg, const float* label, const long long* points,
                                            const unsigned char* keep, const float* offsets, float* out_offset,
                                            float* out_weight, int B, int C, int M, int H, int W, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C > 0 && M >= 0 && H > 0 && W > 0, CL4_EINVAL, "refine_with_point: bad shape B=%d C=%d M=%d H=%d W=%d", B,
                C, M, H, W);
    CL4_REQUIRE(gt_seg && label && offsets && out_offset && out_weight && (M == 0 || (points && keep)), CL4_EINVAL,
                "refine_with_point: null pointer");
    CL4_REQUIRE((long long)H * W < (1ll << 31) - 256 && B <= 65535, CL4_EUNSUPPORTED, "refine_with_point: shape too large");
    if (B == 0) return CL4_OK;
    dim3 grid(ceil_div(H * W, 256), B);
    refine_with_point_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gt_seg, label, points, keep, offsets, out_offset, out_weight, C,
                                                                      M, H, W);
    return check_launch("refine_with_point");
}

extern "C" int cl4_group_pixels(const long long* ctr, const int* count_dev, int Kc, int ctr_stride,
                                const float* offsets, const unsigned char* fg, long long* ids, int N, int H, int W,
                                int empty_mode, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(N >= 0 && H > 0 && W > 0, CL4_EINVAL, "group_pixels: bad shape N=%d H=%d W=%d", N, H, W);
    CL4_REQUIRE(offsets && ids, CL4_EINVAL, "group_pixels: null offsets/ids");
    CL4_REQUIRE(empty_mode == 0 || empty_mode == 1, CL4_EINVAL, "group_pixels: empty_mode must be 0 or 1");
    CL4_REQUIRE((long long)H * W < (1ll << 31) - 4 * kGroupThreads, CL4_EUNSUPPORTED, "group_pixels: H*W too large");
    if (count_dev) {
        CL4_REQUIRE(ctr_stride > 0 && ctr, CL4_EINVAL, "group_pixels: ctr_stride must be > 0 with device counts");
    } else {
        CL4_REQUIRE(Kc >= 0 && (Kc == 0 || ctr), CL4_EINVAL, "group_pixels: bad Kc=%d / null ctr", Kc);
        if (ctr_stride < Kc) ctr_stride = Kc;
    }
    if (N == 0) return CL4_OK;
    const int HW = H * W;
    const bool vec = (HW % 4 == 0) && ((uintptr_t)offsets % 16 == 0) && ((uintptr_t)ids % 16 == 0) &&
                     (!fg || (uintptr_t)fg % 4 == 0);
    dim3 grid(ceil_div(HW, kGroupThreads * kGroupVec), N);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec)
        group_pixels_kernel<true><<<grid, kGroupThreads, 0, s>>>(ctr, count_dev, Kc, ctr_stride, offsets, fg, ids, H,
                                                                  W, empty_mode);
    else
        group_pixels_kernel<false><<<grid, kGroupThreads, 0, s>>>(ctr, count_dev, Kc, ctr_stride, offsets, fg, ids, H,
                                                                   W, empty_mode);
    return check_launch("group_pixels");
}
