// PAMR — pixel-adaptive mask refinement (reference wss/modules.py:122-152).
//
//   w[b,p,y,x]   = softmax_p( mean_k( -|x_k - shift_p x_k| / (1e-8 + 0.1*std_k) ) )      (:141-145)
//   mask <- sum_p w[b,p] * shift_p(mask[b,c])      num_iter times                         (:147-149)
//
// shift_p reads the neighbour at (y + dy*d, x + dx*d) with replicate padding, i.e. clamped
// coordinates (:57-58); p = dilation_index*8 + tap, taps ordered as in :30-40.
//
// Data layout in HBM: image [B,K,H,W], weights [B,P,H,W] (planar per tap so that a warp's
// read of one tap is one coalesced row segment), masks [B,C,H,W]; all fp32.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "pamr_internal.cuh"

#ifndef CL4_LATTICE_DEFAULT
#define CL4_LATTICE_DEFAULT 1
#endif

namespace cl4 {

// ---------------------------------------------------------------------------------------------
// bilinear resize, align_corners=True (reference wss/modules.py:134)
// ---------------------------------------------------------------------------------------------
__global__ void resize_bilinear_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int H,
                                          int W, float sy, float sx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* src = in + (size_t)blockIdx.z * h * w;
    const float fy = __fmul_rn(sy, (float)y), fx = __fmul_rn(sx, (float)x);
    int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
    const float top = __fadd_rn(__fmul_rn(lx0, src[y0 * w + x0]), __fmul_rn(lx1, src[y0 * w + x1]));
    const float bot = __fadd_rn(__fmul_rn(lx0, src[y1 * w + x0]), __fmul_rn(lx1, src[y1 * w + x1]));
    out[((size_t)blockIdx.z * H + y) * W + x] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// ---------------------------------------------------------------------------------------------
// affinity weights: one thread per pixel.  K image channels are looped (K = 3 in the trainer).
// ---------------------------------------------------------------------------------------------
#ifndef CL4_WEIGHTS_MINBLOCKS
#define CL4_WEIGHTS_MINBLOCKS 2
#endif
template <int D>
__global__ void __launch_bounds__(256, CL4_WEIGHTS_MINBLOCKS)
pamr_weights_kernel(const float* __restrict__ img, float* __restrict__ wts, int K, int H, int W, Dilations dil,
                    int tiled) {
    constexpr int P = 8 * D;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;

    float logit[P];
#pragma unroll
    for (int p = 0; p < P; ++p) logit[p] = 0.f;

#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const float* pl = img + ((size_t)b * K + k) * HW;
        const float c = __ldg(pl + (size_t)y * W + x);
        // dlt[p] = neighbour - centre.  The std is taken over the 9*D samples of LocalStDev
        // (wss/modules.py:86-112): the 8*D neighbours plus the centre once per dilation.  Working
        // on deviations from the centre keeps the one-pass variance accurate (the D centre
        // samples contribute zeros).
        float dlt[P];
#pragma unroll
        for (int di = 0; di < D; ++di) {
            const int d = dil.d[di];
            const int ym = max(y - d, 0), yp = min(y + d, H - 1);
            const int xm = max(x - d, 0), xp = min(x + d, W - 1);
            const float* r0 = pl + (size_t)ym * W;
            const float* r1 = pl + (size_t)y * W;
            const float* r2 = pl + (size_t)yp * W;
            dlt[di * 8 + 0] = __ldg(r0 + xm) - c; dlt[di * 8 + 1] = __ldg(r0 + x) - c; dlt[di * 8 + 2] = __ldg(r0 + xp) - c;
            dlt[di * 8 + 3] = __ldg(r1 + xm) - c;                                       dlt[di * 8 + 4] = __ldg(r1 + xp) - c;
            dlt[di * 8 + 5] = __ldg(r2 + xm) - c; dlt[di * 8 + 6] = __ldg(r2 + x) - c; dlt[di * 8 + 7] = __ldg(r2 + xp) - c;
        }
        float s1 = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) s1 += dlt[p];
        const float mean = s1 * (1.f / (float)(9 * D));  // mean deviation over all 9*D samples
        // two-pass unbiased variance; the D centre samples have deviation 0
        float ss = (float)D * mean * mean;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float t = dlt[p] - mean;
            ss = fmaf(t, t, ss);
        }
        const float sd = sqrtf(ss * (1.f / (float)(9 * D - 1)));
        const float ninv = -1.f / (1e-8f + 0.1f * sd);  // -1 / (1e-8 + 0.1*std)   (wss/modules.py:143)
#pragma unroll
        for (int p = 0; p < P; ++p) logit[p] = fmaf(fabsf(dlt[p]), ninv, logit[p]);
    }
    // mean over the K channels (:144), softmax over the 8*D taps (:145)
    const float invK = 1.f / (float)K;
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
    // planar [B,P,H,W] (public layout) or tile-major [tile][P/4][32][32][4] (what the TMA sweep reads)
    if (tiled) {
        const size_t tile = ((size_t)b * (gridDim.y / 4) + (y >> 5)) * gridDim.x + blockIdx.x;
        float4* o = reinterpret_cast<float4*>(wts) + tile * ((size_t)(P / 4) * 1024) + (y & 31) * 32 + threadIdx.x;
#pragma unroll
        for (int g = 0; g < P / 4; ++g)
            o[(size_t)g * 1024] = make_float4(logit[4 * g] * rz, logit[4 * g + 1] * rz, logit[4 * g + 2] * rz,
                                              logit[4 * g + 3] * rz);
    } else {
        float* o = wts + (size_t)b * P * HW + (size_t)y * W + x;
#pragma unroll
        for (int p = 0; p < P; ++p) o[(size_t)p * HW] = logit[p] * rz;
    }
}

// ---------------------------------------------------------------------------------------------
// propagation sweep, v1: one thread per pixel, the pixel's 8*D weights and clamped neighbour
// offsets live in registers and are reused over all C classes; neighbours come through L1.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
pamr_sweep_v1_kernel(const float* __restrict__ wts, const float* __restrict__ min_, float* __restrict__ mout, int C,
                     int H, int W, Dilations dil) {
    constexpr int P = 8 * D;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;

    float w[P];
    int off[P];
    const float* wp = wts + (size_t)b * P * HW + (size_t)y * W + x;
#pragma unroll
    for (int p = 0; p < P; ++p) w[p] = __ldg(wp + (size_t)p * HW);
#pragma unroll
    for (int di = 0; di < D; ++di) {
        const int d = dil.d[di];
        const int ym = max(y - d, 0) * W, y0 = y * W, yp = min(y + d, H - 1) * W;
        const int xm = max(x - d, 0), xp = min(x + d, W - 1);
        off[di * 8 + 0] = ym + xm; off[di * 8 + 1] = ym + x; off[di * 8 + 2] = ym + xp;
        off[di * 8 + 3] = y0 + xm;                             off[di * 8 + 4] = y0 + xp;
        off[di * 8 + 5] = yp + xm; off[di * 8 + 6] = yp + x; off[di * 8 + 7] = yp + xp;
    }
    const float* src = min_ + (size_t)b * C * HW;
    float* dst = mout + (size_t)b * C * HW + (size_t)y * W + x;
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) acc = fmaf(w[p], __ldg(src + off[p]), acc);
        *dst = acc;
        src += HW;
        dst += HW;
    }
}

template <template <int> class F, typename... Args>
static int dispatch_D(int D, Args... args) {
    switch (D) {
        case 1: return F<1>::run(args...);
        case 2: return F<2>::run(args...);
        case 3: return F<3>::run(args...);
        case 4: return F<4>::run(args...);
        case 5: return F<5>::run(args...);
        case 6: return F<6>::run(args...);
        case 7: return F<7>::run(args...);
        case 8: return F<8>::run(args...);
    }
    set_error("pamr: number of dilations %d outside 1..%d", D, CL4_MAX_DILATIONS);
    return CL4_EUNSUPPORTED;
}

template <int D>
struct WeightsLauncher {
    static int run(const float* img, float* w, int B, int K, int H, int W, Dilations dil, int tiled, cudaStream_t s) {
        // grid.y covers whole 32-row tiles (4 blocks of 8 rows each) so that tile indices are exact
        dim3 grid(ceil_div(W, 32), ceil_div(H, 32) * 4, B);
        pamr_weights_kernel<D><<<grid, dim3(32, 8), 0, s>>>(img, w, K, H, W, dil, tiled);
        return check_launch("pamr_weights");
    }
};

template <int D>
struct SweepLauncher {
    static int run(const float* w, const float* mi, float* mo, int B, int C, int H, int W, Dilations dil,
                   cudaStream_t s) {
        dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B);
        pamr_sweep_v1_kernel<D><<<grid, dim3(32, 8), 0, s>>>(w, mi, mo, C, H, W, dil);
        return check_launch("pamr_sweep");
    }
};

static int make_dilations(const int* dilations, int D, Dilations* out) {
    CL4_REQUIRE(dilations && D >= 1, CL4_EINVAL, "pamr: need at least one dilation");
    CL4_REQUIRE(D <= CL4_MAX_DILATIONS, CL4_EUNSUPPORTED, "pamr: %d dilations > %d", D, CL4_MAX_DILATIONS);
    for (int i = 0; i < CL4_MAX_DILATIONS; ++i) out->d[i] = 1;
    for (int i = 0; i < D; ++i) {
        CL4_REQUIRE(dilations[i] >= 1, CL4_EINVAL, "pamr: dilation %d must be >= 1", dilations[i]);
        out->d[i] = dilations[i];
    }
    return CL4_OK;
}

// weights region of the scratch: large enough for either tiled layout
static size_t weight_scratch_elems(int B, int H, int W, int D) {
    const size_t a = tiled_weight_elems(B, H, W, D), b = (D == 6 || D == 5) ? lattice_weight_elems(B, H, W) : 0;
    return a > b ? a : b;
}

// one ping-pong mask buffer of the scratch: replicate-padded planes (4-pixel sweep) or pair cells (duo sweep), whichever is larger
static size_t mask_buffer_bytes(int B, int C, int H, int W) {
    const size_t a = (size_t)B * C * padded_plane_elems(H, W), b = duo_buffer_elems(B, C, H, W);
    return align_up(sizeof(float) * (a > b ? a : b), 256);
}

static int check_plane(const char* what, long long B, int H, int W) {
    CL4_REQUIRE(B >= 0 && H > 0 && W > 0, CL4_EINVAL, "%s: bad shape", what);
    CL4_REQUIRE((long long)H * W < (1ll << 30), CL4_EUNSUPPORTED, "%s: plane too large", what);
    CL4_REQUIRE(B <= 65535, CL4_EUNSUPPORTED, "%s: batch > 65535", what);
    return CL4_OK;
}

}  // namespace cl4

extern "C" int cl4_resize_bilinear_ac(const float* in, float* out, int planes, int h, int w, int H, int W,
                                      cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(planes >= 0 && h > 0 && w > 0 && H > 0 && W > 0, CL4_EINVAL, "resize: bad shape");
    CL4_REQUIRE(planes <= 65535, CL4_EUNSUPPORTED, "resize: more than 65535 planes");
    if (planes == 0) return CL4_OK;
    CL4_REQUIRE(in && out, CL4_EINVAL, "resize: null pointer");
    const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8), planes);
    resize_bilinear_ac_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, h, w, H, W, sy, sx);
    return check_launch("resize_bilinear_ac");
}

extern "C" int cl4_pamr_weights(const float* img, float* w, int B, int K, int H, int W, const int* dilations, int D,
                                cl4_stream_t stream) {
    using namespace cl4;
    int rc = check_plane("pamr_weights", B, H, W);
    if (rc != CL4_OK) return rc;
    CL4_REQUIRE(K >= 1, CL4_EINVAL, "pamr_weights: K must be >= 1");
    Dilations dil;
    rc = make_dilations(dilations, D, &dil);
    if (rc != CL4_OK) return rc;
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(img && w, CL4_EINVAL, "pamr_weights: null pointer");
    return dispatch_D<WeightsLauncher>(D, img, w, B, K, H, W, dil, 0, (cudaStream_t)stream);
}

extern "C" int cl4_pamr_sweep(const float* w, const float* mask_in, float* mask_out, int B, int C, int H, int W,
                              const int* dilations, int D, cl4_stream_t stream) {
    using namespace cl4;
    int rc = check_plane("pamr_sweep", B, H, W);
    if (rc != CL4_OK) return rc;
    CL4_REQUIRE(C >= 1, CL4_EINVAL, "pamr_sweep: C must be >= 1");
    CL4_REQUIRE((long long)C * H * W < (1ll << 31), CL4_EUNSUPPORTED, "pamr_sweep: C*H*W too large");
    Dilations dil;
    rc = make_dilations(dilations, D, &dil);
    if (rc != CL4_OK) return rc;
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(w && mask_in && mask_out && mask_in != mask_out, CL4_EINVAL, "pamr_sweep: null or aliased pointers");
    return dispatch_D<SweepLauncher>(D, w, mask_in, mask_out, B, C, H, W, dil, (cudaStream_t)stream);
}

extern "C" size_t cl4_pamr_scratch_bytes(int B, int K, int C, int H, int W, int D, int num_iter) {
    if (B <= 0 || K <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return 0;
    const size_t wbytes = cl4::align_up(sizeof(float) * cl4::weight_scratch_elems(B, H, W, D), 256);
    // padded path: one replicate-padded copy of the input (+ a second ping-pong buffer from 2 sweeps on);
    // generic path: one plain ping-pong buffer.  The padded layout is the larger of the two.
    const size_t padded = cl4::mask_buffer_bytes(B, C, H, W);
    const size_t padded_img = cl4::align_up(sizeof(float) * (size_t)B * K * cl4::padded_plane_elems(H, W), 256);
    return wbytes + (num_iter >= 2 ? 2 : 1) * padded + padded_img;
}

extern "C" int cl4_pamr_forward_timed(const float* img, const float* mask_in, float* mask_out, void* scratch,
                                      size_t scratch_bytes, int B, int K, int C, int H, int W, const int* dilations,
                                      int D, int num_iter, cl4_stream_t stream, cl4_event_t ev_sweeps_begin,
                                      cl4_event_t ev_sweeps_end) {
    using namespace cl4;
    int rc = check_plane("pamr_forward", B, H, W);
    if (rc != CL4_OK) return rc;
    CL4_REQUIRE(K >= 1 && C >= 1 && num_iter >= 0, CL4_EINVAL, "pamr_forward: bad K/C/num_iter");
    CL4_REQUIRE((long long)C * H * W < (1ll << 31), CL4_EUNSUPPORTED, "pamr_forward: C*H*W too large");
    Dilations dil;
    rc = make_dilations(dilations, D, &dil);
    if (rc != CL4_OK) return rc;
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(img && mask_in && mask_out && mask_in != mask_out, CL4_EINVAL, "pamr_forward: null or aliased pointers");
    CL4_REQUIRE((((uintptr_t)img | (uintptr_t)mask_in | (uintptr_t)mask_out | (uintptr_t)scratch) & 15) == 0, CL4_EINVAL,
                "pamr_forward: img, mask_in, mask_out and scratch must be 16-byte aligned (128-bit loads, TMA)");
    const size_t HW = (size_t)H * W;
    cudaStream_t s = (cudaStream_t)stream;
    auto record = [&](cl4_event_t ev) -> int {
        if (!ev) return CL4_OK;
        cudaError_t e = cudaEventRecord((cudaEvent_t)ev, s);
        CL4_REQUIRE(e == cudaSuccess, CL4_ECUDA, "pamr_forward: event record: %s", cudaGetErrorString(e));
        return CL4_OK;
    };
    if (num_iter == 0) {
        cudaError_t e = cudaMemcpyAsync(mask_out, mask_in, sizeof(float) * (size_t)B * C * HW, cudaMemcpyDeviceToDevice, s);
        CL4_REQUIRE(e == cudaSuccess, CL4_ECUDA, "pamr_forward: copy: %s", cudaGetErrorString(e));
        if ((rc = record(ev_sweeps_begin)) != CL4_OK) return rc;
        return record(ev_sweeps_end);
    }
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_pamr_scratch_bytes(B, K, C, H, W, D, num_iter), CL4_ESCRATCH,
                "pamr_forward: scratch too small");
    char* base = reinterpret_cast<char*>(scratch);
    float* wts = reinterpret_cast<float*>(base);
    const size_t wbytes = align_up(sizeof(float) * weight_scratch_elems(B, H, W, D), 256);
    const size_t padded = mask_buffer_bytes(B, C, H, W);
    float* bufA = reinterpret_cast<float*>(base + wbytes);
    float* bufB = reinterpret_cast<float*>(base + wbytes + padded);
    // CL4_SWEEP=v1 forces the register/L1 kernel, CL4_SWEEP=tma skips the fused small-map kernel
    // (A/B timing and tests of the other paths)
    // CL4_SWEEP=lattice / nolattice: take / skip the lattice sweep (pamr_lattice.cu; the default wherever it applies)
    // (read per call on purpose: the parity tests and tools/abl.sh switch paths inside one process; a getenv is ~50 ns
    // against >= 70 us for the shortest PAMR call)
    const char* force = getenv("CL4_SWEEP");
    const bool force_v1 = force && strcmp(force, "v1") == 0, force_tma = force && strcmp(force, "tma") == 0;
    const bool force_lat1 = force && strcmp(force, "lattice1") == 0;  // the one-class-per-window lattice sweep
    const bool force_lat = force_lat1 || (force && strcmp(force, "lattice") == 0), no_lat = force && strcmp(force, "nolattice") == 0;
    if (!force_v1 && !force_tma && pamr_fused_applicable(H, W, dil, D)) {
        // small maps (the trainer's feature resolution): weights, then every iteration in one launch
        rc = dispatch_D<WeightsLauncher>(D, img, wts, B, K, H, W, dil, 1, s);
        if (rc != CL4_OK) return rc;
        if ((rc = record(ev_sweeps_begin)) != CL4_OK) return rc;
        rc = launch_pamr_fused(wts, mask_in, mask_out, B, C, H, W, num_iter, dil, D, s);
        if (rc != CL4_OK) return rc;
        return record(ev_sweeps_end);
    }
    const bool use_tma = !force_v1 && sweep_tma_applicable(H, W, dil, D);
    const bool use_lattice = use_tma && !force_tma && !no_lat && sweep_lattice_applicable(K, H, W, dil, D) &&
                             (force_lat || CL4_LATTICE_DEFAULT);
    if (use_lattice) {
        float* pimg = reinterpret_cast<float*>(base + wbytes + (num_iter >= 2 ? 2 : 1) * padded);
        rc = launch_pad_copy(img, pimg, (long long)B * K, H, W, s);
        if (rc == CL4_OK) rc = launch_weights_lattice(pimg, wts, B, K, H, W, D, s);
    } else if (use_tma && weights_tma_applicable(K)) {
        // image -> replicate-padded copy (scratch, after the mask buffers) -> TMA-staged weights kernel
        float* pimg = reinterpret_cast<float*>(base + wbytes + (num_iter >= 2 ? 2 : 1) * padded);
        rc = launch_pad_copy(img, pimg, (long long)B * K, H, W, s);
        if (rc == CL4_OK) rc = launch_weights_tma(pimg, wts, B, K, H, W, dil, D, s);
    } else {
        rc = dispatch_D<WeightsLauncher>(D, img, wts, B, K, H, W, dil, use_tma ? 1 : 0, s);
    }
    if (rc != CL4_OK) return rc;
    if (use_lattice && !force_lat1 && num_iter >= 2) {
        // class-pair sweeps.  The first sweep is the one-class lattice kernel reading the caller's planar masks and writing
        // pair cells (a separate pack pass would read and write every mask once more); the last one writes planar.
        if ((rc = record(ev_sweeps_begin)) != CL4_OK) return rc;
        rc = launch_sweep_lattice(wts, mask_in, W, (long long)H * W, bufA, duo_pitch(W), (long long)duo_plane_elems(H, W), 1, B, C, H, W, D, s);
        if (rc != CL4_OK) return rc;
        float* cur = bufA;
        float* nxt = bufB;
        for (int it = 1; it < num_iter; ++it) {
            const bool last = (it == num_iter - 1);
            rc = launch_sweep_duo(wts, cur, last ? mask_out : nxt, last ? 1 : 0, B, C, H, W, D, s);
            if (rc != CL4_OK) return rc;
            float* t = cur; cur = nxt; nxt = t;
        }
        return record(ev_sweeps_end);
    }
    if (use_lattice) {
        // ping-pong in -> A -> B -> A ... -> out with no padding pass.  The scratch planes keep the (W + 48)-float row pitch
        // of the padded layout (a power-of-two pitch such as 2048 bytes makes the 80 rows of a window collide in the memory
        // system: 0.79 ms per sweep instead of 0.5 at 512 x 512); only their H x W image area is ever touched.
        if ((rc = record(ev_sweeps_begin)) != CL4_OK) return rc;
        const int sp = W + 2 * kPamrPad;
        const long long splane = (long long)padded_plane_elems(H, W);
        const float* cur = mask_in;
        int cur_pitch = W;
        long long cur_plane = (long long)H * W;
        float* nxt = bufA;
        for (int it = 0; it < num_iter; ++it) {
            const bool last = (it == num_iter - 1);
            float* dst = last ? mask_out : nxt;
            rc = launch_sweep_lattice(wts, cur, cur_pitch, cur_plane, dst, last ? W : sp, last ? (long long)H * W : splane, 0, B, C, H, W, D, s);
            if (rc != CL4_OK) return rc;
            cur = dst;
            cur_pitch = sp;
            cur_plane = splane;
            nxt = (dst == bufA) ? bufB : bufA;
        }
        return record(ev_sweeps_end);
    }
    if (use_tma) {
        // replicate-padded ping-pong: in -> A -> B -> A ... -> out (plain layout)
        const long long planes = (long long)B * C;
        rc = launch_pad_copy(mask_in, bufA, planes, H, W, s);
        if (rc != CL4_OK) return rc;
        if ((rc = record(ev_sweeps_begin)) != CL4_OK) return rc;
        float* cur = bufA;
        float* nxt = bufB;
        for (int it = 0; it < num_iter; ++it) {
            const bool last = (it == num_iter - 1);
            rc = launch_sweep_tma(wts, cur, last ? mask_out : nxt, last ? 0 : 1, B, C, H, W, dil, D, s);
            if (rc != CL4_OK) return rc;
            if (!last) {
                rc = launch_pad_frame(nxt, planes, H, W, s);
                if (rc != CL4_OK) return rc;
                float* t = cur; cur = nxt; nxt = t;
            }
        }
        return record(ev_sweeps_end);
    }
    // generic path: plain ping-pong so that the last sweep lands in mask_out
    if ((rc = record(ev_sweeps_begin)) != CL4_OK) return rc;
    const float* cur = mask_in;
    for (int it = 0; it < num_iter; ++it) {
        const int remaining = num_iter - it;  // sweeps left including this one
        float* dst = (remaining & 1) ? mask_out : bufA;
        rc = cl4_pamr_sweep(wts, cur, dst, B, C, H, W, dilations, D, stream);
        if (rc != CL4_OK) return rc;
        cur = dst;
    }
    return record(ev_sweeps_end);
}

extern "C" int cl4_pamr_forward(const float* img, const float* mask_in, float* mask_out, void* scratch,
                                size_t scratch_bytes, int B, int K, int C, int H, int W, const int* dilations, int D,
                                int num_iter, cl4_stream_t stream) {
    return cl4_pamr_forward_timed(img, mask_in, mask_out, scratch, scratch_bytes, B, K, C, H, W, dilations, D, num_iter,
                                  stream, nullptr, nullptr);
}
