// Stand-alone forwards of PAMR's helper stencils (reference wss/modules.py:17-119):
//   LocalAffinity      out[p] = x - shift_p(x)          (:26-62, kernel +1 centre / -1 neighbour)
//   LocalAffinityAbs   out[p] = |x - shift_p(x)|        (:115-119)
//   LocalAffinityCopy  out[p] = shift_p(x)              (:65-83)
//   LocalStDev         unbiased std over the 9*D samples shift9_q(x), centre included (:86-112)
// shift_p reads (y + dy*d, x + dx*d) with replicate padding = clamped coordinates (:57); the plane
// index p = dilation_index*8 + tap in the tap order of :30-40.
//
// PAMR.forward never materialises these tensors (pamr.cu / pamr_tma.cu fuse them); the kernels here
// serve callers that use the helper modules on their own.  They are HBM-write-bound: one input
// plane is read through L1 and 8*D output planes are written with coalesced 128-bit stores.
#include "common.cuh"

namespace cl4 {

enum StencilMode { kDiff = 0, kAbs = 1, kCopy = 2 };

// One thread per 4 consecutive pixels of a row; loops over the 8*D output planes.
template <int MODE>
__global__ void __launch_bounds__(128)
local_affinity_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, Dilations dil, int D) {
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x0 >= W) return;
    const size_t HW = (size_t)H * W;
    const float* pl = in + (size_t)blockIdx.z * HW;
    float* o = out + (size_t)blockIdx.z * (size_t)(8 * D) * HW + (size_t)y * W + x0;
    const int n = min(4, W - x0);
    float c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = __ldg(pl + (size_t)y * W + min(x0 + i, W - 1));
    const bool vec = (n == 4) && ((W & 3) == 0);
    for (int di = 0; di < D; ++di) {
        const int d = dil.d[di];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int yy = clampi(y + tap_dy(j) * d, 0, H - 1);
            const float* row = pl + (size_t)yy * W;
            const int sx = tap_dx(j) * d;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float nb = __ldg(row + clampi(x0 + i + sx, 0, W - 1));
                // conv2d with the (+1 centre, -1 neighbour) kernel: a single rounded subtraction
                v[i] = (MODE == kCopy) ? nb : ((MODE == kAbs) ? fabsf(__fsub_rn(c[i], nb)) : __fsub_rn(c[i], nb));
            }
            float* op = o + (size_t)(di * 8 + j) * HW;
            if (vec) {
                __stcs(reinterpret_cast<float4*>(op), make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < n) op[i] = v[i];
            }
        }
    }
}

// std over the 9*D samples (8*D neighbours + the centre once per dilation), unbiased.
__global__ void __launch_bounds__(256)
local_stdev_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, Dilations dil, int D) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;
    const float* pl = in + (size_t)blockIdx.z * HW;
    const float c = __ldg(pl + (size_t)y * W + x);
    // two passes over deviations from the centre (the D centre samples deviate by 0): accurate in fp32
    float s1 = 0.f;
    for (int di = 0; di < D; ++di) {
        const int d = dil.d[di];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            s1 += __ldg(pl + (size_t)clampi(y + tap_dy(j) * d, 0, H - 1) * W + clampi(x + tap_dx(j) * d, 0, W - 1)) - c;
    }
    const int N = 9 * D;
    const float mean = s1 / (float)N;
    float ss = (float)D * mean * mean;
    for (int di = 0; di < D; ++di) {
        const int d = dil.d[di];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t =
                __ldg(pl + (size_t)clampi(y + tap_dy(j) * d, 0, H - 1) * W + clampi(x + tap_dx(j) * d, 0, W - 1)) - c - mean;
            ss = fmaf(t, t, ss);
        }
    }
    // torch.std of a single sample (N - 1 == 0) is NaN; 9*D >= 9 here
    out[(size_t)blockIdx.z * HW + (size_t)y * W + x] = sqrtf(ss / (float)(N - 1));
}

// smoothing(heat, kernel) = avg_pool2d(kernel, stride 1, zero padding (k-1)//2, padded cells counted):
// wss/utils.py:28-32.  Window values are added in row-major order (ATen's order), then divided by k*k.
__global__ void __launch_bounds__(256)
smoothing_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int r, float area) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* pl = in + (size_t)blockIdx.z * H * W;
    float acc = 0.f;
    for (int yy = max(y - r, 0); yy <= min(y + r, H - 1); ++yy)
        for (int xx = max(x - r, 0); xx <= min(x + r, W - 1); ++xx) acc = __fadd_rn(acc, __ldg(pl + (size_t)yy * W + xx));
    out[(size_t)blockIdx.z * H * W + (size_t)y * W + x] = __fdiv_rn(acc, area);  // sum / divide_factor, as ATen
}

static int stencil_args(const char* what, long long planes, int H, int W, const int* dilations, int D, Dilations* dil) {
    CL4_REQUIRE(planes >= 0 && H > 0 && W > 0, CL4_EINVAL, "%s: bad shape", what);
    CL4_REQUIRE(planes <= 65535, CL4_EUNSUPPORTED, "%s: more than 65535 planes", what);
    CL4_REQUIRE(H <= 65535, CL4_EUNSUPPORTED, "%s: H > 65535", what);
    CL4_REQUIRE(dilations && D >= 1, CL4_EINVAL, "%s: need at least one dilation", what);
    CL4_REQUIRE(D <= CL4_MAX_DILATIONS, CL4_EUNSUPPORTED, "%s: %d dilations > %d", what, D, CL4_MAX_DILATIONS);
    for (int i = 0; i < CL4_MAX_DILATIONS; ++i) dil->d[i] = 1;
    for (int i = 0; i < D; ++i) {
        CL4_REQUIRE(dilations[i] >= 1, CL4_EINVAL, "%s: dilation %d must be >= 1", what, dilations[i]);
        dil->d[i] = dilations[i];
    }
    return CL4_OK;
}

}  // namespace cl4

extern "C" int cl4_local_affinity(const float* x, float* out, int planes, int H, int W, const int* dilations, int D,
                                  int mode, cl4_stream_t stream) {
    using namespace cl4;
    Dilations dil;
    int rc = stencil_args("local_affinity", planes, H, W, dilations, D, &dil);
    if (rc != CL4_OK) return rc;
    CL4_REQUIRE(mode >= 0 && mode <= 2, CL4_EINVAL, "local_affinity: mode %d not in {0: diff, 1: abs, 2: copy}", mode);
    if (planes == 0) return CL4_OK;
    CL4_REQUIRE(x && out, CL4_EINVAL, "local_affinity: null pointer");
    dim3 grid(ceil_div(ceil_div(W, 4), 128), H, planes);
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == kDiff) local_affinity_kernel<kDiff><<<grid, 128, 0, s>>>(x, out, H, W, dil, D);
    else if (mode == kAbs) local_affinity_kernel<kAbs><<<grid, 128, 0, s>>>(x, out, H, W, dil, D);
    else local_affinity_kernel<kCopy><<<grid, 128, 0, s>>>(x, out, H, W, dil, D);
    return check_launch("local_affinity");
}

extern "C" int cl4_smoothing(const float* x, float* out, int planes, int H, int W, int kernel, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(planes >= 0 && H > 0 && W > 0, CL4_EINVAL, "smoothing: bad shape");
    CL4_REQUIRE(planes <= 65535, CL4_EUNSUPPORTED, "smoothing: more than 65535 planes");
    CL4_REQUIRE(kernel > 0 && (kernel & 1), CL4_EINVAL, "smoothing: kernel must be odd and positive, got %d", kernel);
    if (planes == 0) return CL4_OK;
    CL4_REQUIRE(x && out && x != out, CL4_EINVAL, "smoothing: null or aliased pointers");
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8), planes);
    smoothing_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, out, H, W, (kernel - 1) / 2,
                                                                    (float)(kernel * kernel));
    return check_launch("smoothing");
}

extern "C" int cl4_local_stdev(const float* x, float* out, int planes, int H, int W, const int* dilations, int D,
                               cl4_stream_t stream) {
    using namespace cl4;
    Dilations dil;
    int rc = stencil_args("local_stdev", planes, H, W, dilations, D, &dil);
    if (rc != CL4_OK) return rc;
    if (planes == 0) return CL4_OK;
    CL4_REQUIRE(x && out, CL4_EINVAL, "local_stdev: null pointer");
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8), planes);
    local_stdev_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, out, H, W, dil, D);
    return check_launch("local_stdev");
}
