// Connected-component labelling on the GPU for the centre-clustering step of
// get_instance_segmentation (reference modules/utils.py:608-632, cluster_peaks):
//   weak = (sqrt(off_x^2 + off_y^2) < thresh) & fg           numpy fp32 arithmetic
//   cv2.connectedComponentsWithStats(weak, connectivity=4)   -> area and centroid per component
// Union-find over pixels (atomicMin on parent links); a component is represented by its smallest
// pixel index, which is also OpenCV's label order for 4-connectivity (labels are numbered by the
// first pixel met in raster order — pinned against cv2 in tests/test_gpu_parity.py::test_cluster_peaks_matches_opencv).
#include "ccl.cuh"
#include "common.cuh"

namespace cl4 {

// label[i] = i for weak-offset foreground pixels, -1 elsewhere
__global__ void ccl_init_kernel(const float* __restrict__ off, const unsigned char* __restrict__ fg, float thresh,
                                int HW, int* __restrict__ label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float oy = off[i], ox = off[HW + i];
    // numpy: offset_map[1] ** 2 + offset_map[0] ** 2, then sqrt, all fp32 (modules/utils.py:619)
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)));
    label[i] = ((mag < thresh) && fg[i]) ? i : -1;
}

__global__ void ccl_merge_kernel(int* __restrict__ label, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = y * W + x;
    if (label[i] < 0) return;
    if (x > 0 && label[i - 1] >= 0) ccl_union(label, i, i - 1);
    if (y > 0 && label[i - W] >= 0) ccl_union(label, i, i - W);
}

// flatten + statistics: area, sum of x, sum of y per root; stats[0..2] of the background
// (OpenCV's label 0 = every non-component pixel) go to bg[0..2].
__global__ void ccl_stats_kernel(int* __restrict__ label, int H, int W, int* __restrict__ area,
                                 unsigned long long* __restrict__ sx, unsigned long long* __restrict__ sy,
                                 unsigned long long* __restrict__ bg) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const bool in = (x < W && y < H);
    const int i = y * W + x;
    int root = -1;
    if (in && label[i] >= 0) {
        root = ccl_find(label, i);
        atomicAdd(&area[root], 1);
        atomicAdd(&sx[root], (unsigned long long)x);
        atomicAdd(&sy[root], (unsigned long long)y);
    }
    // background: warp-aggregated
    const bool isbg = in && root < 0;
    const unsigned m = __ballot_sync(0xffffffffu, isbg);
    unsigned long long bx = isbg ? (unsigned long long)x : 0ull, by = isbg ? (unsigned long long)y : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx += __shfl_down_sync(0xffffffffu, bx, o);
        by += __shfl_down_sync(0xffffffffu, by, o);
    }
    if ((threadIdx.x & 31) == 0 && m) {
        atomicAdd(&bg[0], (unsigned long long)__popc(m));
        atomicAdd(&bg[1], bx);
        atomicAdd(&bg[2], by);
    }
}

// keep-mask words (same format as the centre NMS) of the roots whose area lies in (lo, hi)
__global__ void ccl_select_kernel(const int* __restrict__ label, const int* __restrict__ area, float lo, float hi,
                                  int H, int W, int words_per_row, uint32_t* __restrict__ words) {
    const int lane = threadIdx.x & 31;
    const int xw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int y = blockIdx.y;
    if (xw >= words_per_row) return;
    const int x = xw * 32 + lane;
    bool keep = false;
    if (x < W) {
        const int i = y * W + x;
        if (label[i] == i) {
            const float a = (float)area[i];
            keep = (lo < a) && (a < hi);  // 21 - beta < area < 21 + beta (modules/utils.py:630)
        }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) words[(size_t)y * words_per_row + xw] = word;
}

// statistics of the selected roots (tiny)
__global__ void ccl_gather_kernel(const long long* __restrict__ roots, const int* __restrict__ count, int max_out, int W,
                                  const int* __restrict__ area, const unsigned long long* __restrict__ sx,
                                  const unsigned long long* __restrict__ sy, const unsigned long long* __restrict__ bg,
                                  long long* __restrict__ stats) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {
        stats[0] = (long long)bg[0];
        stats[1] = (long long)bg[1];
        stats[2] = (long long)bg[2];
    }
    const int n = min(*count, max_out);
    if (j < n) {
        const int i = (int)roots[2 * j] * W + (int)roots[2 * j + 1];
        stats[3 * (j + 1) + 0] = area[i];
        stats[3 * (j + 1) + 1] = (long long)sx[i];
        stats[3 * (j + 1) + 2] = (long long)sy[i];
    }
}

// ordered compaction of keep-mask words, defined in nms.cu
int launch_center_compact(const uint32_t* words, int N, int H, int words_per_row, long long* ctr_out, int* count_out,
                          int max_out, int* row_off, cudaStream_t s);

}  // namespace cl4

extern "C" size_t cl4_ccl4_scratch_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    const size_t HW = (size_t)H * W;
    // label i32, area i32, sx u64, sy u64, bg 3 x u64 (+pad), words, row offsets
    return cl4::align_up(HW * 4, 256) * 2 + cl4::align_up(HW * 8, 256) * 2 + 256 +
           cl4::align_up((size_t)H * cl4::ceil_div(W, 32) * 4, 256) + cl4::align_up((size_t)H * 4, 256);
}

// offsets [2,H,W] fp32 (dy, dx), fg [H,W] u8 -> roots_out [max_out,2] int64 (y,x of each selected
// component's first pixel, raster order), stats_out [max_out+1,3] int64: row 0 = background
// (area, sum x, sum y), rows 1.. = the selected components in the same order; count_out [1] int32.
extern "C" int cl4_ccl4_components(const float* offsets, const unsigned char* fg, float thresh, float area_lo,
                                   float area_hi, int H, int W, long long* roots_out, long long* stats_out,
                                   int* count_out, int max_out, void* scratch, size_t scratch_bytes,
                                   cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 31), CL4_EINVAL, "ccl4: bad shape");
    CL4_REQUIRE(offsets && fg && roots_out && stats_out && count_out && max_out > 0, CL4_EINVAL, "ccl4: null pointer");
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_ccl4_scratch_bytes(H, W), CL4_ESCRATCH, "ccl4: scratch too small");
    const int HW = H * W;
    char* p = reinterpret_cast<char*>(scratch);
    int* label = reinterpret_cast<int*>(p); p += align_up((size_t)HW * 4, 256);
    int* area = reinterpret_cast<int*>(p); p += align_up((size_t)HW * 4, 256);
    unsigned long long* sx = reinterpret_cast<unsigned long long*>(p); p += align_up((size_t)HW * 8, 256);
    unsigned long long* sy = reinterpret_cast<unsigned long long*>(p); p += align_up((size_t)HW * 8, 256);
    unsigned long long* bg = reinterpret_cast<unsigned long long*>(p); p += 256;
    const int wpr = ceil_div(W, 32);
    uint32_t* words = reinterpret_cast<uint32_t*>(p); p += align_up((size_t)H * wpr * 4, 256);
    int* row_off = reinterpret_cast<int*>(p);
    cudaStream_t s = (cudaStream_t)stream;
    // area, sx, sy, bg are contiguous: one memset
    cudaError_t e = cudaMemsetAsync(area, 0, (char*)words - (char*)area, s);
    CL4_REQUIRE(e == cudaSuccess, CL4_ECUDA, "ccl4: memset: %s", cudaGetErrorString(e));
    ccl_init_kernel<<<ceil_div(HW, 256), 256, 0, s>>>(offsets, fg, thresh, HW, label);
    dim3 blk(32, 8), grd(ceil_div(W, 32), ceil_div(H, 8));
    ccl_merge_kernel<<<grd, blk, 0, s>>>(label, H, W);
    ccl_stats_kernel<<<grd, blk, 0, s>>>(label, H, W, area, sx, sy, bg);
    ccl_select_kernel<<<dim3(ceil_div(wpr, 8), H), 256, 0, s>>>(label, area, area_lo, area_hi, H, W, wpr, words);
    int rc = check_launch("ccl4");
    if (rc != CL4_OK) return rc;
    rc = launch_center_compact(words, 1, H, wpr, roots_out, count_out, max_out, row_off, s);
    if (rc != CL4_OK) return rc;
    ccl_gather_kernel<<<ceil_div(max_out, 256), 256, 0, s>>>(roots_out, count_out, max_out, W, area, sx, sy, bg, stats_out);
    return check_launch("ccl4_gather");
}
