// Producers and consumers of PAMR inside the phase-1 step (reference train.py:372-385), SURVEY §8f rank 2:
//   int_masks_soft = int_masks.softmax(dim=1)                                   train.py:373
//   im = F.interpolate(denorm(images), int_masks.shape[-2:], "bilinear", align_corners=True)   :376-378
//        denorm: utils/utils.py:26-41  (x * std + mean per channel, mean = (0.485, 0.456, 0.4069))
//   int_masks_soft = PAMR(im, int_masks_soft)                                   :379  (pamr*.cu)
//   int_masks_soft[:, 1:] *= l1h[:, :, None, None]                              :382
//   pseudo_gt_seg = pseudo_gtmask(int_masks_soft, cutoff_top=0.6, cutoff_bkg=0.7, cutoff_low=0.2)   :384
//        pseudo_gtmask: wss/single_stage.py:18-40
// Everything here works on feature-resolution maps (32 x 32 ... 64 x 64), so the kernels are
// latency-bound; the point is to replace a dozen framework launches by three.
#include "common.cuh"

namespace cl4 {

struct Affine3 {
    float mul[3], add[3];
};

// out[b,k,y,x] = bilinear(align_corners=True) of (in * mul[k] + add[k]); one thread per output pixel.
// The denormalisation is applied to each of the four taps (the reference denormalises the whole
// image first), with separate roundings for the multiply and the add as Tensor.mul_().add_() has.
__global__ void denorm_resize_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int K, int Hi, int Wi, int h,
                                        int w, float sy, float sx, Affine3 a, int apply) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int plane = blockIdx.z, k = plane % K;
    const float* src = in + (size_t)plane * Hi * Wi;
    const float fy = __fmul_rn(sy, (float)y), fx = __fmul_rn(sx, (float)x);
    const int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
    const int y1 = y0 + (y0 < Hi - 1), x1 = x0 + (x0 < Wi - 1);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
    auto tap = [&](int yy, int xx) {
        const float v = __ldg(src + (size_t)yy * Wi + xx);
        return apply ? __fadd_rn(__fmul_rn(v, a.mul[k]), a.add[k]) : v;
    };
    const float top = __fadd_rn(__fmul_rn(lx0, tap(y0, x0)), __fmul_rn(lx1, tap(y0, x1)));
    const float bot = __fadd_rn(__fmul_rn(lx0, tap(y1, x0)), __fmul_rn(lx1, tap(y1, x1)));
    out[((size_t)plane * h + y) * w + x] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// elementwise denorm (utils/utils.py:26-41) for callers that use it on its own
__global__ void denorm_kernel(const float* __restrict__ in, float* __restrict__ out, int K, size_t HW, Affine3 a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int plane = blockIdx.y, k = plane % K;
    out[(size_t)plane * HW + i] = __fadd_rn(__fmul_rn(in[(size_t)plane * HW + i], a.mul[k]), a.add[k]);
}

// softmax over the channel dimension of [B,C,HW]: one thread per (b, pixel), three passes over C
__global__ void softmax_channels_kernel(const float* __restrict__ in, float* __restrict__ out, int C, size_t HW) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float* p = in + (size_t)blockIdx.y * C * HW + i;
    float* o = out + (size_t)blockIdx.y * C * HW + i;
    float mx = p[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, p[(size_t)c * HW]);
    float z = 0.f;
    for (int c = 0; c < C; ++c) z += expf(p[(size_t)c * HW] - mx);
    for (int c = 0; c < C; ++c) o[(size_t)c * HW] = expf(p[(size_t)c * HW] - mx) / z;
}

// One CTA per (b, c) plane: gate the fg planes by the image-level label (train.py:382), keep the plane
// maximum and turn it into the plane's threshold (wss/single_stage.py:24-32).
__global__ void __launch_bounds__(256)
gate_and_threshold_kernel(const float* mask, const float* __restrict__ labels, float* gated, int C,  // gated may alias mask
                          int HW, float cutoff_top, float cutoff_bkg, float cutoff_low, float* __restrict__ thr) {
    __shared__ float s_max[8];
    const int c = blockIdx.x, b = blockIdx.y;
    const size_t o = ((size_t)b * C + c) * HW;
    const float g = (labels && c > 0) ? labels[(size_t)b * (C - 1) + (c - 1)] : 1.f;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        float v = mask[o + i];
        if (labels && c > 0) v = __fmul_rn(v, g);
        if (gated) gated[o + i] = v;
        mx = (v > mx || v != v) ? v : mx;  // torch.max propagates NaN
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, mx, s);
        mx = (t > mx || t != t) ? t : mx;
    }
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) {
            const float t = s_max[wi];
            mx = (t > mx || t != t) ? t : mx;
        }
        const float scaled = __fmul_rn(mx, c == 0 ? cutoff_bkg : cutoff_top);  // mask_max[:, :1] *= bkg; [:, 1:] *= top
        thr[(size_t)b * C + c] = (cutoff_low > scaled || cutoff_low != cutoff_low) ? cutoff_low : scaled;  // .max(lowest)
    }
}

// pseudo_gt = (mask > thr); pixels claimed by more than one class are cleared (wss/single_stage.py:34-38)
__global__ void pseudo_gt_kernel(const float* __restrict__ mask, const float* __restrict__ thr, float* __restrict__ out, int C,
                                 int HW, int ambiguous) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= HW) return;
    const float* m = mask + (size_t)b * C * HW + i;
    float* o = out + (size_t)b * C * HW + i;
    int n = 0;
    for (int c = 0; c < C; ++c) n += (m[(size_t)c * HW] > thr[(size_t)b * C + c]);
    const bool clear = ambiguous && n > 1;
    for (int c = 0; c < C; ++c) o[(size_t)c * HW] = (!clear && m[(size_t)c * HW] > thr[(size_t)b * C + c]) ? 1.f : 0.f;
}

}  // namespace cl4

extern "C" int cl4_denorm_resize_ac(const float* images, float* out, int B, int K, int Hi, int Wi, int h, int w,
                                    const float* mean, const float* std, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && K >= 1 && Hi > 0 && Wi > 0 && h > 0 && w > 0, CL4_EINVAL, "denorm_resize: bad shape");
    CL4_REQUIRE((long long)B * K <= 65535, CL4_EUNSUPPORTED, "denorm_resize: more than 65535 planes");
    CL4_REQUIRE((mean == nullptr) == (std == nullptr), CL4_EINVAL, "denorm_resize: mean and std go together");
    CL4_REQUIRE(!mean || K == 3, CL4_EINVAL, "denorm_resize: denorm expects RGB images [B,3,H,W]");  // utils/utils.py:37
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(images && out, CL4_EINVAL, "denorm_resize: null pointer");
    Affine3 a{{1.f, 1.f, 1.f}, {0.f, 0.f, 0.f}};
    if (mean)
        for (int k = 0; k < 3; ++k) {
            a.mul[k] = std[k];
            a.add[k] = mean[k];
        }
    const float sy = (h > 1) ? (float)(Hi - 1) / (float)(h - 1) : 0.f;
    const float sx = (w > 1) ? (float)(Wi - 1) / (float)(w - 1) : 0.f;
    dim3 grid(ceil_div(w, 32), ceil_div(h, 8), B * K);
    denorm_resize_ac_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(images, out, K, Hi, Wi, h, w, sy, sx, a,
                                                                          mean ? 1 : 0);
    return check_launch("denorm_resize_ac");
}

extern "C" int cl4_denorm(const float* images, float* out, int planes, int K, long long HW, const float* mean,
                          const float* std, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(planes >= 0 && K == 3 && HW > 0 && planes % 3 == 0, CL4_EINVAL, "denorm: expected RGB planes");
    CL4_REQUIRE(planes <= 65535, CL4_EUNSUPPORTED, "denorm: more than 65535 planes");
    if (planes == 0) return CL4_OK;
    CL4_REQUIRE(images && out && mean && std, CL4_EINVAL, "denorm: null pointer");
    Affine3 a;
    for (int k = 0; k < 3; ++k) {
        a.mul[k] = std[k];
        a.add[k] = mean[k];
    }
    dim3 grid((unsigned)((HW + 255) / 256), planes);
    denorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, out, K, (size_t)HW, a);
    return check_launch("denorm");
}

extern "C" int cl4_softmax_channels(const float* x, float* out, int B, int C, long long HW, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 1 && HW > 0, CL4_EINVAL, "softmax_channels: bad shape");
    CL4_REQUIRE(B <= 65535, CL4_EUNSUPPORTED, "softmax_channels: batch > 65535");
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(x && out, CL4_EINVAL, "softmax_channels: null pointer");
    dim3 grid((unsigned)((HW + 255) / 256), B);
    softmax_channels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, C, (size_t)HW);
    return check_launch("softmax_channels");
}

extern "C" int cl4_pseudo_gtmask(const float* mask, const float* labels, float* gated_out, float* pseudo_out,
                                 float* thr_scratch, int B, int C, int HW, float cutoff_top, float cutoff_bkg,
                                 float cutoff_low, int ambiguous, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 1 && HW > 0, CL4_EINVAL, "pseudo_gtmask: bad shape");
    CL4_REQUIRE(B <= 65535 && C <= 65535, CL4_EUNSUPPORTED, "pseudo_gtmask: B or C > 65535");
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(mask && pseudo_out && thr_scratch, CL4_EINVAL, "pseudo_gtmask: null pointer");
    CL4_REQUIRE(gated_out || !labels, CL4_EINVAL, "pseudo_gtmask: label gating needs gated_out");
    cudaStream_t s = (cudaStream_t)stream;
    gate_and_threshold_kernel<<<dim3(C, B), 256, 0, s>>>(mask, labels, gated_out, C, HW, cutoff_top, cutoff_bkg, cutoff_low,
                                                        thr_scratch);
    pseudo_gt_kernel<<<dim3(ceil_div(HW, 256), B), 256, 0, s>>>(gated_out ? gated_out : mask, thr_scratch, pseudo_out, C, HW,
                                                                ambiguous);
    return check_launch("pseudo_gtmask");
}
