// Device-resident, batched refine_label_generation (reference modules/utils.py:257-385, the live
// phase-2 caller of the hot path: train.py:492-500) and pseudo_label_generation (:179-253, train.py:451-466).
//
// The reference walks image x class x 8-connected contour on the host; for every contour it copies a
// mask to the CPU for OpenCV, calls get_instance_segmentation (centre NMS, cluster_peaks via OpenCV,
// group_pixels) and then reads several scalars back per instance — at least six host/device round trips
// per contour.  Here the whole batch runs as a fixed sequence of kernels with no host
// involvement; one status word tells the caller whether a capacity limit was hit (then the exact
// per-contour path in cl4wsis_b200/modules/utils.py takes over).
//
//  1. contours     8-connected components of equal gt label over every valid (image, class) at once
//                  (union-find, root = smallest pixel index), area and centroid sums per root;
//                  contours with area >= MINIMUM_MASK_SIZE get a slot                       (:301-315)
//  2. centre NMS   a contour pixel p is a centre iff heat(p) > thr and no pixel of the SAME contour in
//                  the k x k window is larger (= threshold + max-pool + equality of
//                  find_instance_center on the contour-masked heat, modules/utils.py:480-492); ordered
//                  compaction gives torch.nonzero order, hence the instance ids
//  3. clustering   weak = |offset| < 2.5 inside the contour, 4-connected components, area filter
//                  21-beta < a < 21+beta, centroid (int32 truncation), heat > 0.05, merge rule of
//                  modules/utils.py:569-592 (OpenCV label 0 included, in OpenCV label order)
//  4. grouping     nearest centre per contour pixel with group_pixels' exact arithmetic   (:505-542)
//  5. instances    per (contour, id): pixel count, mean seg probability, first arg-max of the (marked)
//                  heat in row-major order; confidence / centre choice of :344-362
//  6. outputs      weight and offset maps per pixel, gaussian max-splat per instance       (:364-377)
//
// Outputs are bit-exact w.r.t. the reference except `weight`, whose mean over the instance mask is a
// floating-point reduction (the reference's own value depends on ATen's summation order): the mean is
// accumulated in double here and agrees to fp32 rounding.
#include <math.h>

#include "ccl.cuh"
#include "common.cuh"

namespace cl4 {

constexpr int kRefMaxComp = 1024;  // contours (area >= min_area) per image
constexpr int kRefMaxCtr = 64;     // centres per contour
constexpr int kRefMaxInst = 8;     // instance slots per contour (ids 1..7; MAXIMUM_NUM_INST = 5)
constexpr int kRefListCap = 4096;  // NMS centres / cluster components per image

enum RefStatus {  // bits of the status word; any bit set = use the per-contour path instead
    kRefTooManyContours = 1, kRefTooManyCentres = 2, kRefListOverflow = 4, kRefTopKBranch = 8
};

struct RefComp {  // one 8-connected contour
    int root, cls, cx, cy, area;
    int n_nms, n_ctr, n_ins;
    unsigned long long marked;  // bit j: centre j is an accepted cluster centre (its heat reads 1.0, :582,:591)
    int weak_cnt;
    unsigned long long weak_sx, weak_sy;
    int ctr[kRefMaxCtr];  // pixel indices, NMS centres in row-major order, then accepted cluster centres
    int cnt[kRefMaxInst];
    double psum[kRefMaxInst];
    unsigned long long key[kRefMaxInst];
    int py[kRefMaxInst], px[kRefMaxInst];
    float conf[kRefMaxInst];
};

struct RefDims {
    int B, C, H, W, HW;
};

__device__ __forceinline__ unsigned orderable(float v) {  // monotone map float -> unsigned
    if (v == 0.f) v = 0.f;                                // -0.0 == +0.0 for argmax
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unorderable(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- 1. contours.  A warp labels 32 consecutive pixels of a row: every pixel starts out pointing at
// the first pixel of its horizontal run inside the 32-pixel segment (one ballot instead of 31 unions).
__global__ void __launch_bounds__(256)
ref_init_kernel(const long long* __restrict__ gt, const float* __restrict__ label, RefDims d, int wpr,
                int* __restrict__ root) {
    const int lane = threadIdx.x & 31;
    const int xw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int y = blockIdx.y, b = blockIdx.z;
    if (xw >= wpr) return;
    const int x = xw * 32 + lane;
    long long L = 0;
    bool valid = false;
    if (x < d.W) {
        L = gt[(size_t)b * d.HW + y * d.W + x];
        valid = L >= 1 && L <= d.C && label[(size_t)b * d.C + (L - 1)] != 0.f;  // np.nonzero(label[b])  :299
    }
    const long long Lleft = __shfl_up_sync(0xffffffffu, L, 1);
    const bool joins = lane > 0 && valid && Lleft == L;  // same label as the pixel to the left (hence valid too)
    const unsigned m = __ballot_sync(0xffffffffu, joins);
    if (x < d.W) {
        const unsigned starts = ~m & (0xffffffffu >> (31 - lane));  // run starts at or before this lane
        root[(size_t)b * d.HW + y * d.W + x] = valid ? y * d.W + xw * 32 + (31 - __clz(starts)) : -1;
    }
}

// Links between runs: with a = NW, b = N, c = NE, d = W of a pixel p (same label), every 8-connection is
// implied by these few unions — p~b is made by p only when it has no d (otherwise d's own links reach
// b's run), p~c only when there is no b, p~a only when there is neither b nor d.
__global__ void ref_merge8_kernel(const long long* __restrict__ gt, RefDims d, int* __restrict__ root_all) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    int* root = root_all + (size_t)b * d.HW;
    const long long* g = gt + (size_t)b * d.HW;
    const int i = y * d.W + x;
    if (root[i] < 0) return;
    const long long L = g[i];  // equal label => same (valid) class: connectivity 8 of (seg == cls+1), :305-307
    const bool hd = x > 0 && g[i - 1] == L;
    if (hd && (x & 31) == 0) ccl_union(root, i, i - 1);  // runs continue across 32-pixel segments
    if (y > 0) {
        const bool hb = g[i - d.W] == L;
        if (hb) {
            if (!hd) ccl_union(root, i, i - d.W);
        } else {
            if (x + 1 < d.W && g[i - d.W + 1] == L) ccl_union(root, i, i - d.W + 1);
            if (!hd && x > 0 && g[i - d.W - 1] == L) ccl_union(root, i, i - d.W - 1);
        }
    }
}

__global__ void ref_flatten_stats_kernel(RefDims d, int* __restrict__ root_all, int* __restrict__ area_all,
                                         unsigned long long* __restrict__ sx_all, unsigned long long* __restrict__ sy_all) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    const size_t o = (size_t)b * d.HW;
    const int i = y * d.W + x;
    if (root_all[o + i] < 0) return;
    const int r = ccl_find(root_all + o, i);
    root_all[o + i] = r;
    atomicAdd(&area_all[o + r], 1);
    atomicAdd(&sx_all[o + r], (unsigned long long)x);
    atomicAdd(&sy_all[o + r], (unsigned long long)y);
}

// roots with area >= min_area get a slot; comp[] receives the slot at the root position
__global__ void ref_make_comps_kernel(const long long* __restrict__ gt, RefDims d, const int* __restrict__ root,
                                      const int* __restrict__ area, const unsigned long long* __restrict__ sx,
                                      const unsigned long long* __restrict__ sy, int min_area, int* __restrict__ comp,
                                      RefComp* __restrict__ comps, int* __restrict__ ncomp, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    int slot = -1;
    if (root[o + i] == i && area[o + i] >= min_area) {  // `if size < MINIMUM_MASK_SIZE: continue`  :313
        slot = atomicAdd(&ncomp[b], 1);
        if (slot < kRefMaxComp) {
            RefComp& c = comps[(size_t)b * kRefMaxComp + slot];
            c.root = i;
            c.cls = (int)gt[o + i] - 1;
            c.area = area[o + i];
            // cx, cy = int(centroid): OpenCV divides the integer coordinate sums by the area in double
            c.cx = (int)((double)sx[o + i] / (double)area[o + i]);
            c.cy = (int)((double)sy[o + i] / (double)area[o + i]);
        } else {
            atomicOr(status, kRefTooManyContours);
            slot = -1;
        }
    }
    if (root[o + i] == i || root[o + i] < 0) comp[o + i] = slot;
}

__global__ void ref_comp_map_kernel(RefDims d, const int* __restrict__ root, int* __restrict__ comp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    const int r = root[o + i];
    if (r >= 0 && r != i) comp[o + i] = comp[o + r];
}

// ---- 2. per-contour centre NMS.  One warp per 32 consecutive pixels of a row; for every candidate
// (heat > thr) the whole warp scans the k x k window for a larger heat in the same contour.
__global__ void __launch_bounds__(256)
ref_nms_kernel(const float* __restrict__ center, RefDims d, const int* __restrict__ comp_all,
               const RefComp* __restrict__ comps, float thr, int r, int wpr, uint32_t* __restrict__ words) {
    const int lane = threadIdx.x & 31;
    const int xw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int y = blockIdx.y, b = blockIdx.z;
    if (xw >= wpr) return;
    const int* comp = comp_all + (size_t)b * d.HW;
    const int x = xw * 32 + lane;
    int s = -1;
    float h = 0.f;
    const float* plane = nullptr;
    if (x < d.W) {
        s = comp[y * d.W + x];
        if (s >= 0) {
            plane = center + ((size_t)b * d.C + comps[(size_t)b * kRefMaxComp + s].cls) * d.HW;
            h = plane[y * d.W + x];
        }
    }
    // F.threshold(x, thr, -1) then nonzero(> 0): a centre has heat > thr and heat > 0  (:480,:492)
    const bool cand = s >= 0 && h > thr && h > 0.f;
    unsigned todo = __ballot_sync(0xffffffffu, cand);
    bool keep = false;
    const int side = 2 * r + 1, cells = side * side;
    while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const int cs = __shfl_sync(0xffffffffu, s, l);
        const float ch = __shfl_sync(0xffffffffu, h, l);
        const int cx = xw * 32 + l;
        const float* cpl = reinterpret_cast<const float*>(
            __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(plane), l));
        bool beaten = false;
        for (int base = 0; base < cells; base += 32 * 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = base + u * 32 + lane;
                if (c < cells) {
                    const int wy = c / side, wx = c - wy * side;
                    const int qy = y + wy - r, qx = cx + wx - r;
                    if (qy >= 0 && qy < d.H && qx >= 0 && qx < d.W) {
                        const int q = qy * d.W + qx;
                        if (comp[q] == cs && cpl[q] > ch) beaten = true;
                    }
                }
            }
            if (__any_sync(0xffffffffu, beaten)) break;
        }
        beaten = __any_sync(0xffffffffu, beaten);
        if (lane == l) keep = !beaten;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) words[((size_t)b * d.H + y) * wpr + xw] = word;
}

// centres of the ordered per-image list -> per-contour lists (rank = earlier centres of the same contour)
__global__ void ref_assign_centres_kernel(RefDims d, const int* __restrict__ comp_all, const long long* __restrict__ list,
                                          const int* __restrict__ count, RefComp* __restrict__ comps,
                                          int* __restrict__ status) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int n = count[b];
    if (j == 0 && n > kRefListCap) atomicOr(status, kRefListOverflow);
    if (j >= min(n, kRefListCap)) return;
    const int* comp = comp_all + (size_t)b * d.HW;
    const long long* L = list + (size_t)b * kRefListCap * 2;
    const int p = (int)L[2 * j] * d.W + (int)L[2 * j + 1];
    const int s = comp[p];
    int rank = 0;
    for (int i = 0; i < j; ++i) rank += (comp[(int)L[2 * i] * d.W + (int)L[2 * i + 1]] == s);
    RefComp& c = comps[(size_t)b * kRefMaxComp + s];
    if (rank < kRefMaxCtr) c.ctr[rank] = p;
    else atomicOr(status, kRefTooManyCentres);
    atomicAdd(&c.n_nms, 1);
}

// ---- 3. centre clustering (cluster_peaks, modules/utils.py:608-632) restricted to each contour
__global__ void ref_weak_init_kernel(const float* __restrict__ offsets, RefDims d, const int* __restrict__ comp,
                                     float thresh, int* __restrict__ root2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    const float oy = offsets[(size_t)b * 2 * d.HW + i], ox = offsets[(size_t)b * 2 * d.HW + d.HW + i];
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)));  // numpy fp32, :619
    root2[o + i] = (comp[o + i] >= 0 && mag < thresh) ? i : -1;
}

__global__ void ref_weak_merge4_kernel(RefDims d, const int* __restrict__ comp_all, int* __restrict__ root2_all) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    int* root2 = root2_all + (size_t)b * d.HW;
    const int* comp = comp_all + (size_t)b * d.HW;
    const int i = y * d.W + x;
    if (root2[i] < 0) return;
    const int s = comp[i];  // the weak map is multiplied by the contour mask (:623): never across contours
    if (x > 0 && root2[i - 1] >= 0 && comp[i - 1] == s) ccl_union(root2, i, i - 1);
    if (y > 0 && root2[i - d.W] >= 0 && comp[i - d.W] == s) ccl_union(root2, i, i - d.W);
}

__global__ void ref_weak_stats_kernel(RefDims d, const int* __restrict__ comp_all, int* __restrict__ root2_all,
                                      int* __restrict__ area_all, unsigned long long* __restrict__ sx_all,
                                      unsigned long long* __restrict__ sy_all, RefComp* __restrict__ comps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    const size_t o = (size_t)b * d.HW;
    const int i = y * d.W + x;
    if (root2_all[o + i] < 0) return;
    const int r = ccl_find(root2_all + o, i);
    root2_all[o + i] = r;
    atomicAdd(&area_all[o + r], 1);
    atomicAdd(&sx_all[o + r], (unsigned long long)x);
    atomicAdd(&sy_all[o + r], (unsigned long long)y);
    RefComp& c = comps[(size_t)b * kRefMaxComp + comp_all[o + i]];
    atomicAdd(&c.weak_cnt, 1);
    atomicAdd(&c.weak_sx, (unsigned long long)x);
    atomicAdd(&c.weak_sy, (unsigned long long)y);
}

__global__ void ref_weak_select_kernel(RefDims d, const int* __restrict__ root2, const int* __restrict__ area, float lo,
                                       float hi, int wpr, uint32_t* __restrict__ words) {
    const int lane = threadIdx.x & 31;
    const int xw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int y = blockIdx.y, b = blockIdx.z;
    if (xw >= wpr) return;
    const int x = xw * 32 + lane;
    bool keep = false;
    if (x < d.W) {
        const size_t o = (size_t)b * d.HW;
        const int i = y * d.W + x;
        if (root2[o + i] == i) {
            const float a = (float)area[o + i];
            keep = (lo < a) && (a < hi);  // 21 - beta < area < 21 + beta  (:630)
        }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) words[((size_t)b * d.H + y) * wpr + xw] = word;
}

// merge of NMS and cluster centres, one thread per contour (modules/utils.py:569-592)
__global__ void ref_merge_clusters_kernel(const float* __restrict__ center, RefDims d, const int* __restrict__ comp_all,
                                          const int* __restrict__ area2_all, const unsigned long long* __restrict__ sx2_all,
                                          const unsigned long long* __restrict__ sy2_all,
                                          const long long* __restrict__ cl_list, const int* __restrict__ cl_count,
                                          int use_clusters, float lo, float hi, long long top_k,
                                          RefComp* __restrict__ comps, const int* __restrict__ ncomp,
                                          int* __restrict__ status) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (s >= min(ncomp[b], kRefMaxComp)) return;
    RefComp& c = comps[(size_t)b * kRefMaxComp + s];
    const size_t o = (size_t)b * d.HW;
    const int* comp = comp_all + o;
    const float* plane = center + ((size_t)b * d.C + c.cls) * d.HW;
    if (top_k >= 0 && (long long)c.n_nms >= top_k) atomicOr(status, kRefTopKBranch);  // degenerate branch, :498-502
    const int n = min(c.n_nms, kRefMaxCtr);
    int n_ctr = n;
    unsigned long long marked = 0ull;
    auto consider = [&](int cy, int cx) {
        const int q = cy * d.W + cx;
        const float hc = (comp[q] == s) ? plane[q] : 0.f;  // contour-masked heat (:319)
        if (!(hc > 0.05f)) return;                          // :571
        bool accept = (n == 0);                             // no NMS centre: cluster centres take over (:578-583)
        if (!accept) {                                      // farther than 100 px from every NMS centre (:586-591)
            long long best = -1;
            for (int j = 0; j < n; ++j) {
                const long long dy = c.ctr[j] / d.W - cy, dx = c.ctr[j] % d.W - cx;
                const long long d2 = dy * dy + dx * dx;     // integer coordinates: sqrt(d2) > 100 <=> d2 > 10000
                if (best < 0 || d2 < best) best = d2;
            }
            accept = best > 10000;
        }
        if (!accept) return;
        if (n_ctr < kRefMaxCtr) {
            c.ctr[n_ctr] = q;
            marked |= 1ull << n_ctr;
            ++n_ctr;
        } else {
            atomicOr(status, kRefTooManyCentres);
        }
    };
    if (use_clusters) {
        // OpenCV's label 0 = every pixel outside this contour's weak region; the reference filters it
        // like any other label (k starts at 0, :630)
        const long long bg_area = (long long)d.HW - c.weak_cnt;
        const float a0 = (float)bg_area;
        if (lo < a0 && a0 < hi) {
            const unsigned long long tot_x = (unsigned long long)d.H * ((unsigned long long)d.W * (d.W - 1) / 2);
            const unsigned long long tot_y = (unsigned long long)d.W * ((unsigned long long)d.H * (d.H - 1) / 2);
            consider((int)((double)(tot_y - c.weak_sy) / (double)bg_area), (int)((double)(tot_x - c.weak_sx) / (double)bg_area));
        }
        const int m = min(cl_count[b], kRefListCap);
        if (s == 0 && cl_count[b] > kRefListCap) atomicOr(status, kRefListOverflow);
        const long long* L = cl_list + (size_t)b * kRefListCap * 2;
        for (int i = 0; i < m; ++i) {  // raster order of the first pixel = OpenCV's label order (4-connectivity)
            const int q = (int)L[2 * i] * d.W + (int)L[2 * i + 1];
            if (comp[q] != s) continue;
            const double a = (double)area2_all[o + q];
            consider((int)((double)sy2_all[o + q] / a), (int)((double)sx2_all[o + q] / a));  // np.int32(centroid[::-1])
        }
    }
    c.n_ctr = n_ctr;
    c.marked = marked;
}

// ---- 4. grouping: nearest centre of the pixel's contour, group_pixels arithmetic (:523-540)
// empty_id: the id of every pixel of a contour without any centre -- 0 (`ignore`: zeros_like(fg)) or 1 (fg.long()), :597-600
__global__ void ref_group_kernel(const float* __restrict__ offsets, RefDims d, const int* __restrict__ comp_all,
                                 RefComp* __restrict__ comps, unsigned char* __restrict__ ids, int empty_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    const int s = comp_all[o + i];
    unsigned char id = 0;
    if (s >= 0) {
        RefComp& c = comps[(size_t)b * kRefMaxComp + s];
        const int n = c.n_ctr;
        if (n > 0) {
            const int y = i / d.W, x = i - y * d.W;
            const float ly = __fadd_rn((float)y, offsets[(size_t)b * 2 * d.HW + i]);
            const float lx = __fadd_rn((float)x, offsets[(size_t)b * 2 * d.HW + d.HW + i]);
            float best_d = 0.f;
            int best_k = 0;
            for (int k = 0; k < n; ++k) {
                const int q = c.ctr[k];
                const float dy = __fsub_rn((float)(q / d.W), ly), dx = __fsub_rn((float)(q % d.W), lx);
                const float dd = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                if (k == 0 || dd < best_d) {
                    best_d = dd;
                    best_k = k;
                }
            }
            id = (unsigned char)(best_k + 1);
            atomicMax(&c.n_ins, best_k + 1);  // n_ins = ins_seg.max()  (:332)
        } else if (empty_id) {
            id = (unsigned char)empty_id;
            if (c.n_ins < empty_id) atomicMax(&c.n_ins, empty_id);
        }
    }
    ids[o + i] = id;
}

// ---- 5. per-instance statistics
__global__ void ref_inst_stats_kernel(const float* __restrict__ seg, const float* __restrict__ center,
                                      const float* __restrict__ label, RefDims d, const int* __restrict__ comp_all,
                                      const unsigned char* __restrict__ ids, int max_inst, RefComp* __restrict__ comps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    const int id = ids[o + i];
    if (id == 0) return;
    RefComp& c = comps[(size_t)b * kRefMaxComp + comp_all[o + i]];
    if (c.n_ins > max_inst) return;  // too many centres in a single contour (:335)
    // heat of the contour with accepted cluster centres marked 1.0
    float h = center[((size_t)b * d.C + c.cls) * d.HW + i];
    if (c.marked)
        for (int j = 0; j < c.n_ctr; ++j)
            if (((c.marked >> j) & 1ull) && c.ctr[j] == i) h = 1.f;
    // softmax over the C+1 channels, fg channels gated by the image-level label (:279-280)
    const float* sp = seg + (size_t)b * (d.C + 1) * d.HW + i;
    float mx = sp[0];
    for (int k = 1; k <= d.C; ++k) mx = fmaxf(mx, sp[(size_t)k * d.HW]);
    float z = 0.f, e_cls = 0.f;
    for (int k = 0; k <= d.C; ++k) {
        const float e = expf(sp[(size_t)k * d.HW] - mx);
        z += e;
        if (k == c.cls + 1) e_cls = e;
    }
    const float prob = (e_cls / z) * label[(size_t)b * d.C + c.cls];
    atomicAdd(&c.cnt[id], 1);
    atomicAdd(&c.psum[id], (double)prob);
    // first maximum in row-major order (torch.where order + argmax, :342-345)
    atomicMax(&c.key[id], ((unsigned long long)orderable(h) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
}

__global__ void ref_finalize_kernel(RefDims d, double refine_thresh, int max_inst, RefComp* __restrict__ comps,
                                    const int* __restrict__ ncomp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int s = t / kRefMaxInst, id = t % kRefMaxInst;
    if (s >= min(ncomp[b], kRefMaxComp) || id == 0) return;
    RefComp& c = comps[(size_t)b * kRefMaxComp + s];
    if (c.n_ins > max_inst || id > c.n_ins || c.cnt[id] == 0) return;
    const unsigned long long key = c.key[id];
    const int p = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
    const float center_score = unorderable((unsigned)(key >> 32));
    const float seg_score = (float)(c.psum[id] / (double)c.cnt[id]);  // .mean() of fp32 probabilities
    int py = p / d.W, px = p % d.W;
    double conf;
    if ((double)center_score < refine_thresh) {  // weak peak: the contour's centroid stands in (:354-358)
        py = c.cy;
        px = c.cx;
        conf = (double)seg_score;
    } else {
        conf = (double)center_score * (double)seg_score;
    }
    conf = fmax(0.0, fmin(conf, 1.0));
    c.py[id] = py;
    c.px[id] = px;
    c.conf[id] = (float)conf;
}

// ---- 6. outputs
__global__ void ref_write_kernel(RefDims d, const int* __restrict__ comp_all, const unsigned char* __restrict__ ids,
                                 int max_inst, const RefComp* __restrict__ comps, float* __restrict__ out_offset,
                                 float* __restrict__ out_weight) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    const int id = ids[o + i];
    float w = 0.f, oy = 0.f, ox = 0.f;
    if (id) {
        const RefComp& c = comps[(size_t)b * kRefMaxComp + comp_all[o + i]];
        if (c.n_ins <= max_inst) {
            const int y = i / d.W, x = i - y * d.W;
            w = c.conf[id];
            oy = (float)c.py[id] - (float)y;  // py - y_coord[index]  (:375-376)
            ox = (float)c.px[id] - (float)x;
        }
    }
    out_weight[o + i] = w;
    out_offset[(size_t)b * 2 * d.HW + i] = oy;
    out_offset[(size_t)b * 2 * d.HW + d.HW + i] = ox;
}

// gaussian max-splat, one block per contour slot (center_map_gen, modules/utils.py:84-119)
__global__ void __launch_bounds__(256)
ref_splat_kernel(RefDims d, const float* __restrict__ gauss, int sigma, int max_inst, const RefComp* __restrict__ comps,
                 const int* __restrict__ ncomp, float* __restrict__ out_center) {
    const int s = blockIdx.x, b = blockIdx.y;
    if (s >= min(ncomp[b], kRefMaxComp)) return;
    const RefComp& c = comps[(size_t)b * kRefMaxComp + s];
    if (c.n_ins > max_inst) return;
    const int gs = 6 * sigma + 3;
    int* plane = reinterpret_cast<int*>(out_center + ((size_t)b * d.C + c.cls) * d.HW);
    for (int id = 1; id <= c.n_ins && id < kRefMaxInst; ++id) {
        if (c.cnt[id] == 0) continue;
        const int x0 = c.px[id] - 3 * sigma - 1, y0 = c.py[id] - 3 * sigma - 1;
        for (int t = threadIdx.x; t < gs * gs; t += blockDim.x) {
            const int gy = t / gs, gx = t - gy * gs;
            const int iy = y0 + gy, ix = x0 + gx;
            if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W)
                atomicMax(&plane[iy * d.W + ix], __float_as_int(gauss[t]));  // values >= 0: int order = float order
        }
    }
}

// copy of the contour table for the per-contour (fallback) path: [B, kRefMaxComp, 5] int32 = root, cls, cx, cy, area
__global__ void ref_export_comps_kernel(const RefComp* __restrict__ comps, const int* __restrict__ ncomp,
                                        int* __restrict__ info) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (s >= min(ncomp[b], kRefMaxComp)) return;
    const RefComp& c = comps[(size_t)b * kRefMaxComp + s];
    int* o = info + ((size_t)b * kRefMaxComp + s) * 5;
    o[0] = c.root; o[1] = c.cls; o[2] = c.cx; o[3] = c.cy; o[4] = c.area;
}

int launch_center_compact(const uint32_t* words, int N, int H, int words_per_row, long long* ctr_out, int* count_out,
                          int max_out, int* row_off, cudaStream_t s);

struct RefScratch {
    int *root, *area, *comp, *ncomp, *status, *row_off, *cnt_nms, *cnt_cl;
    unsigned long long *sx, *sy;
    unsigned char* ids;
    uint32_t* words;
    long long *list_nms, *list_cl;
    RefComp* comps;
    size_t bytes;
};

static RefScratch ref_layout(char* base, int B, int H, int W) {
    const size_t n = (size_t)B * H * W;
    const int wpr = ceil_div(W, 32);
    RefScratch r;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    // area, sx, sy are contiguous (one memset)
    r.area = reinterpret_cast<int*>(take(n * 4));
    r.sx = reinterpret_cast<unsigned long long*>(take(n * 8));
    r.sy = reinterpret_cast<unsigned long long*>(take(n * 8));
    r.root = reinterpret_cast<int*>(take(n * 4));
    r.comp = reinterpret_cast<int*>(take(n * 4));
    r.ids = reinterpret_cast<unsigned char*>(take(n));
    r.words = reinterpret_cast<uint32_t*>(take((size_t)B * H * wpr * 4));
    r.row_off = reinterpret_cast<int*>(take((size_t)B * H * 4));
    r.list_nms = reinterpret_cast<long long*>(take((size_t)B * kRefListCap * 16));
    r.list_cl = reinterpret_cast<long long*>(take((size_t)B * kRefListCap * 16));
    // comps, ncomp, status, counts are contiguous (one memset)
    r.comps = reinterpret_cast<RefComp*>(take((size_t)B * kRefMaxComp * sizeof(RefComp)));
    r.ncomp = reinterpret_cast<int*>(take((size_t)B * 4));
    r.cnt_nms = reinterpret_cast<int*>(take((size_t)B * 4));
    r.cnt_cl = reinterpret_cast<int*>(take((size_t)B * 4));
    r.status = reinterpret_cast<int*>(take(256));
    r.bytes = off;
    return r;
}

#define REF_CUDA(call, what)                                                                   \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        CL4_REQUIRE(e_ == cudaSuccess, CL4_ECUDA, "refine: %s: %s", what, cudaGetErrorString(e_)); \
    } while (0)

// steps 1: contours of the whole batch
static int run_contours(const long long* gt, const float* label, RefDims d, int min_area, const RefScratch& sc,
                        cudaStream_t s) {
    const size_t n = (size_t)d.B * d.HW;
    REF_CUDA(cudaMemsetAsync(sc.area, 0, (char*)sc.root - (char*)sc.area, s), "memset");
    REF_CUDA(cudaMemsetAsync(sc.comps, 0, (char*)sc.status + 256 - (char*)sc.comps, s), "memset");
    (void)n;
    dim3 lin(ceil_div(d.HW, 256), d.B), blk(32, 8), grd(ceil_div(d.W, 32), ceil_div(d.H, 8), d.B);
    const int wpr = ceil_div(d.W, 32);
    ref_init_kernel<<<dim3(ceil_div(wpr, 8), d.H, d.B), 256, 0, s>>>(gt, label, d, wpr, sc.root);
    ref_merge8_kernel<<<grd, blk, 0, s>>>(gt, d, sc.root);
    ref_flatten_stats_kernel<<<grd, blk, 0, s>>>(d, sc.root, sc.area, sc.sx, sc.sy);
    ref_make_comps_kernel<<<lin, 256, 0, s>>>(gt, d, sc.root, sc.area, sc.sx, sc.sy, min_area, sc.comp, sc.comps, sc.ncomp,
                                              sc.status);
    ref_comp_map_kernel<<<lin, 256, 0, s>>>(d, sc.root, sc.comp);
    return check_launch("refine contours");
}

// ---- pseudo_label_generation (modules/utils.py:179-253) for a batch: contours holding exactly one
// confident CAM peak of their class become instances centred on the contour's centroid.
__global__ void pl_match_kernel(const float* __restrict__ conf, const int* __restrict__ py, const int* __restrict__ px,
                                const float* __restrict__ label, RefDims d, int K, float thresh,
                                const int* __restrict__ comp_all, RefComp* __restrict__ comps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= d.C * K) return;
    const int c = t / K, j = t - c * K;
    if (label[(size_t)b * d.C + c] == 0.f) return;  // valid_label = np.nonzero(cls_label[b])   train.py:453
    const float* cf = conf + ((size_t)b * d.C + c) * K;
    for (int i = 0; i <= j; ++i)
        if (cf[i] < thresh) return;  // `if conf < pseudo_thresh: break` over the score-sorted peaks  train.py:456-457
    const size_t o = ((size_t)b * d.C + c) * K + j;
    const int y = py[o], x = px[o];
    if (y < 0 || y >= d.H || x < 0 || x >= d.W) return;
    const int s = comp_all[(size_t)b * d.HW + y * d.W + x];
    if (s < 0) return;
    RefComp& cc = comps[(size_t)b * kRefMaxComp + s];
    if (cc.cls == c) atomicAdd(&cc.n_nms, 1);  // points of this class inside the contour (:235-238)
}

__global__ void pl_accept_kernel(RefComp* __restrict__ comps, const int* __restrict__ ncomp, int* __restrict__ total_match) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (s >= min(ncomp[b], kRefMaxComp)) return;
    RefComp& c = comps[(size_t)b * kRefMaxComp + s];
    if (c.n_nms == 1) {  // accept: 1 contour - 1 point (:241); the centre is the contour's centroid (:245)
        c.n_ins = 1;
        c.cnt[1] = 1;
        c.px[1] = c.cx;
        c.py[1] = c.cy;
        c.conf[1] = 1.f;
        atomicAdd(&total_match[b], 1);
    } else {
        c.n_ins = 0;
    }
}

__global__ void pl_write_kernel(RefDims d, const int* __restrict__ comp_all, const RefComp* __restrict__ comps,
                                float* __restrict__ out_offset, float* __restrict__ out_weight) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= d.HW) return;
    const size_t o = (size_t)b * d.HW;
    const int s = comp_all[o + i];
    float w = 0.f, oy = 0.f, ox = 0.f;
    if (s >= 0) {
        const RefComp& c = comps[(size_t)b * kRefMaxComp + s];
        if (c.n_ins == 1) {
            const int y = i / d.W, x = i - y * d.W;
            w = 1.f;
            oy = (float)c.cy - (float)y;  // cy - y_coord[mask_index]  (:250-251)
            ox = (float)c.cx - (float)x;
        }
    }
    out_weight[o + i] = w;
    out_offset[(size_t)b * 2 * d.HW + i] = oy;
    out_offset[(size_t)b * 2 * d.HW + d.HW + i] = ox;
}

// ---------------------------------------------------------------------------------------------------------------------
// get_ins_map (dataset/utils.py:795-902, Trainer.validate -> train.py:622) for one image on the device.
//
// The reference: softmax, flip test-time augmentation, label cleaning, argmax on the device; then per class one cv2 call
// on the host, and per contour get_instance_segmentation plus several .item() reads per instance.  Here: one pass over
// the pixels (im_prepare_kernel), the contour / NMS / clustering / grouping kernels of refine_label_generation above with
// min_area 50 and the validation thresholds, per-(contour, id) statistics, and ONE block that orders the contours as
// the reference visits them (class ascending, then OpenCV's label order) and numbers the instances.  The host reads the
// instance count once, then cl4_ins_masks writes the [n, H, W] boolean masks.
constexpr int kImMaxInst = 4096;           // instances per image in the output table
constexpr int kImMaxG = 2 * kRefListCap;   // centres of all contours together (NMS centres <= kRefListCap, + cluster centres, + one
                                           // pseudo centre per contour that has none when ignore is off)

// Unlike the training path (at most 64 centres per contour, then the per-contour fallback), validation runs with small NMS
// kernels on noisy heat maps: a contour may hold hundreds of NMS centres of which only a few attract pixels.  So the
// centres of a contour are a RANGE of one per-image array: nms_sorted[nbase[s] .. + n_nms[s]) in raster order, followed by the
// contour's accepted cluster centres (RefComp::ctr, at most kRefMaxCtr); a centre's global index gbase[s] + k (k = id - 1)
// addresses the per-instance statistics.
struct ImScratch {
    float *pmax, *center_avg, *ones;
    int *cslot, *crank, *nms_sorted, *nbase, *gbase, *gcount, *gidx, *minx, *rank, *icnt;
    double* ipsum;
    unsigned long long* ikey;
    size_t bytes;
};

static ImScratch im_layout(char* base, size_t off, int C, int H, int W) {
    ImScratch r;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    const size_t n = (size_t)H * W;
    r.pmax = reinterpret_cast<float*>(take(n * 4));
    r.center_avg = reinterpret_cast<float*>(take(n * 4 * C));
    r.ones = reinterpret_cast<float*>(take((size_t)C * 4));
    r.cslot = reinterpret_cast<int*>(take((size_t)kRefListCap * 4));
    r.crank = reinterpret_cast<int*>(take((size_t)kRefListCap * 4));
    r.nms_sorted = reinterpret_cast<int*>(take((size_t)kRefListCap * 4));
    r.nbase = reinterpret_cast<int*>(take((size_t)kRefMaxComp * 4));
    r.gbase = reinterpret_cast<int*>(take((size_t)kRefMaxComp * 4));
    r.gcount = reinterpret_cast<int*>(take((size_t)kRefMaxComp * 4));
    r.gidx = reinterpret_cast<int*>(take(n * 4));
    r.rank = reinterpret_cast<int*>(take((size_t)kImMaxG * 4));
    // icnt, ipsum, ikey are contiguous (one memset); minx is filled with 0x7f bytes
    r.icnt = reinterpret_cast<int*>(take((size_t)kImMaxG * 4));
    r.ipsum = reinterpret_cast<double*>(take((size_t)kImMaxG * 8));
    r.ikey = reinterpret_cast<unsigned long long*>(take((size_t)kImMaxG * 8));
    r.minx = reinterpret_cast<int*>(take((size_t)kRefMaxComp * 4));
    r.bytes = off;
    return r;
}

// softmax over the C+1 channels of each test-time view (:818), average with the mirrored second view (:823-825), label
// cleaning (:835), argmax (:837, first maximum); the offsets of view 0 are rescaled IN PLACE as the reference does (:831-832).
__global__ void __launch_bounds__(256)
im_prepare_kernel(const float* __restrict__ seg, const float* __restrict__ center, float* offset0,
                  const float* __restrict__ cls_label, int flip, float scale_y, float scale_x, RefDims d,
                  long long* __restrict__ seg_map, float* __restrict__ pmax, float* __restrict__ center_avg,
                  float* __restrict__ ones) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.C) ones[i] = 1.f;
    if (i >= d.HW) return;
    const int y = i / d.W, x = i - y * d.W;
    const int im = y * d.W + (d.W - 1 - x);  // the same pixel in the mirrored view
    const float* s0 = seg + i;
    const float* s1 = seg + (size_t)(d.C + 1) * d.HW + im;
    float mx0 = s0[0], mx1 = flip ? s1[0] : 0.f;
    for (int k = 1; k <= d.C; ++k) {
        mx0 = fmaxf(mx0, s0[(size_t)k * d.HW]);
        if (flip) mx1 = fmaxf(mx1, s1[(size_t)k * d.HW]);
    }
    float z0 = 0.f, z1 = 0.f;
    for (int k = 0; k <= d.C; ++k) {
        z0 += expf(s0[(size_t)k * d.HW] - mx0);
        if (flip) z1 += expf(s1[(size_t)k * d.HW] - mx1);
    }
    float best = 0.f;
    int arg = 0;
    for (int k = 0; k <= d.C; ++k) {
        float p = expf(s0[(size_t)k * d.HW] - mx0) / z0;
        if (flip) p = (p + expf(s1[(size_t)k * d.HW] - mx1) / z1) / 2.f;
        if (cls_label && k >= 1) p *= cls_label[k - 1];
        if (k == 0 || p > best || (p != p && best == best)) {  // torch.argmax: first maximum, NaN counts as the largest
            best = p;
            arg = k;
        }
    }
    seg_map[i] = arg;
    pmax[i] = best;  // = seg_prob[cls + 1] at every pixel of a contour of class cls
    if (flip) {
        for (int c = 0; c < d.C; ++c)
            center_avg[(size_t)c * d.HW + i] = (center[(size_t)c * d.HW + i] + center[(size_t)(d.C + c) * d.HW + im]) / 2.f;
    }
    offset0[i] = offset0[i] * scale_y;
    offset0[d.HW + i] = offset0[d.HW + i] * scale_x;
}

// NMS centres of the ordered per-image list: contour slot and rank inside the contour (raster order)
__global__ void im_assign_kernel(RefDims d, const int* __restrict__ comp, const long long* __restrict__ list,
                                 const int* __restrict__ count, RefComp* __restrict__ comps, int* __restrict__ cslot,
                                 int* __restrict__ crank, int* __restrict__ status) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = count[0];
    if (j == 0 && n > kRefListCap) atomicOr(status, kRefListOverflow);
    if (j >= min(n, kRefListCap)) return;
    const int p = (int)list[2 * j] * d.W + (int)list[2 * j + 1];
    const int s = comp[p];
    int rank = 0;
    for (int i = 0; i < j; ++i) rank += (comp[(int)list[2 * i] * d.W + (int)list[2 * i + 1]] == s);
    cslot[j] = s;
    crank[j] = rank;
    atomicAdd(&comps[s].n_nms, 1);
}

// one block: exclusive scan over the contour slots of `what` (0: n_nms -> nbase; 1: max(n_nms + n_cl, empty_id) -> gbase, gcount)
__global__ void __launch_bounds__(kRefMaxComp)
im_scan_kernel(const RefComp* __restrict__ comps, const int* __restrict__ ncomp, int what, int empty_id, int* __restrict__ base,
               int* __restrict__ cnt_out, int* __restrict__ status) {
    __shared__ int pre[kRefMaxComp];
    const int t = threadIdx.x;
    const int n = min(ncomp[0], kRefMaxComp);
    int mine = 0;
    if (t < n) mine = what == 0 ? comps[t].n_nms : max(comps[t].n_nms + comps[t].n_ctr, empty_id);
    pre[t] = mine;
    __syncthreads();
    for (int off = 1; off < kRefMaxComp; off <<= 1) {
        const int v = (t >= off) ? pre[t - off] : 0;
        __syncthreads();
        pre[t] += v;
        __syncthreads();
    }
    base[t] = pre[t] - mine;
    if (cnt_out) cnt_out[t] = mine;
    if (t == kRefMaxComp - 1 && pre[t] > (what == 0 ? kRefListCap : kImMaxG)) atomicOr(status, kRefListOverflow);
}

__global__ void im_scatter_kernel(RefDims d, const long long* __restrict__ list, const int* __restrict__ count,
                                  const int* __restrict__ cslot, const int* __restrict__ crank, const int* __restrict__ nbase,
                                  int* __restrict__ nms_sorted) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= min(count[0], kRefListCap)) return;
    nms_sorted[nbase[cslot[j]] + crank[j]] = (int)list[2 * j] * d.W + (int)list[2 * j + 1];
}

// merge of NMS and cluster centres, one thread per contour (dataset/utils.py:728-751 = modules/utils.py:569-592); the
// accepted cluster centres go to RefComp::ctr[0 .. n_ctr)
__global__ void im_merge_clusters_kernel(const float* __restrict__ center, RefDims d, const int* __restrict__ comp,
                                         const int* __restrict__ area2, const unsigned long long* __restrict__ sx2,
                                         const unsigned long long* __restrict__ sy2, const long long* __restrict__ cl_list,
                                         const int* __restrict__ cl_count, float lo, float hi, const int* __restrict__ nms_sorted,
                                         const int* __restrict__ nbase, RefComp* __restrict__ comps,
                                         const int* __restrict__ ncomp, int* __restrict__ status) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= min(ncomp[0], kRefMaxComp)) return;
    RefComp& c = comps[s];
    const float* plane = center + (size_t)c.cls * d.HW;
    const int n = c.n_nms;
    const int* mine = nms_sorted + nbase[s];
    int n_cl = 0;
    auto consider = [&](int cy, int cx) {
        const int q = cy * d.W + cx;
        const float hc = (comp[q] == s) ? plane[q] : 0.f;  // contour-masked heat
        if (!(hc > 0.05f)) return;
        bool accept = (n == 0);
        if (!accept) {
            long long best = -1;
            for (int j = 0; j < n; ++j) {
                const long long dy = mine[j] / d.W - cy, dx = mine[j] % d.W - cx;
                const long long d2 = dy * dy + dx * dx;
                if (best < 0 || d2 < best) best = d2;
            }
            accept = best > 10000;
        }
        if (!accept) return;
        if (n_cl < kRefMaxCtr) c.ctr[n_cl++] = q;
        else atomicOr(status, kRefTooManyCentres);
    };
    const long long bg_area = (long long)d.HW - c.weak_cnt;  // OpenCV's label 0 (see ref_merge_clusters_kernel)
    const float a0 = (float)bg_area;
    if (lo < a0 && a0 < hi) {
        const unsigned long long tot_x = (unsigned long long)d.H * ((unsigned long long)d.W * (d.W - 1) / 2);
        const unsigned long long tot_y = (unsigned long long)d.W * ((unsigned long long)d.H * (d.H - 1) / 2);
        consider((int)((double)(tot_y - c.weak_sy) / (double)bg_area), (int)((double)(tot_x - c.weak_sx) / (double)bg_area));
    }
    const int m = min(cl_count[0], kRefListCap);
    if (s == 0 && cl_count[0] > kRefListCap) atomicOr(status, kRefListOverflow);
    for (int i = 0; i < m; ++i) {
        const int q = (int)cl_list[2 * i] * d.W + (int)cl_list[2 * i + 1];
        if (comp[q] != s) continue;
        const double a = (double)area2[q];
        consider((int)((double)sy2[q] / a), (int)((double)sx2[q] / a));
    }
    c.n_ctr = n_cl;
}

// grouping (group_pixels arithmetic) over the contour's NMS centres, then its cluster centres; gidx = global centre index
__global__ void im_group_kernel(const float* __restrict__ offsets, RefDims d, const int* __restrict__ comp,
                                const RefComp* __restrict__ comps, const int* __restrict__ nms_sorted,
                                const int* __restrict__ nbase, const int* __restrict__ gbase, int empty_id,
                                int* __restrict__ gidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.HW) return;
    const int s = comp[i];
    int g = -1;
    if (s >= 0) {
        const RefComp& c = comps[s];
        const int n1 = c.n_nms, n = n1 + c.n_ctr;
        if (n > 0) {
            const int* mine = nms_sorted + nbase[s];
            const int y = i / d.W, x = i - y * d.W;
            const float ly = __fadd_rn((float)y, offsets[i]);
            const float lx = __fadd_rn((float)x, offsets[d.HW + i]);
            float best_d = 0.f;
            int best_k = 0;
            for (int k = 0; k < n; ++k) {
                const int q = k < n1 ? mine[k] : c.ctr[k - n1];
                const float dy = __fsub_rn((float)(q / d.W), ly), dx = __fsub_rn((float)(q % d.W), lx);
                const float dd = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                if (k == 0 || dd < best_d) {
                    best_d = dd;
                    best_k = k;
                }
            }
            g = gbase[s] + best_k;
        } else if (empty_id) {
            g = gbase[s];  // fg.long(): the whole contour is instance 1  (dataset/utils.py:756-759)
        }
    }
    gidx[i] = g;
}

// per centre: pixel count, sum of the class probability, first arg-max of the contour's (marked) heat (:868-878)
__global__ void im_stats_kernel(const float* __restrict__ center, const float* __restrict__ pmax, RefDims d,
                                const int* __restrict__ comp, const int* __restrict__ gidx, const RefComp* __restrict__ comps,
                                int* __restrict__ icnt, double* __restrict__ ipsum, unsigned long long* __restrict__ ikey,
                                int* __restrict__ minx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.HW) return;
    const int s = comp[i];
    if (s < 0) return;
    const RefComp& c = comps[s];
    // OpenCV's label order (8-connectivity, 2 x 2 block scan): block row of the contour's first pixel, then the first block
    // column of the contour inside that block row
    if ((i / d.W) >> 1 == (c.root / d.W) >> 1) atomicMin(&minx[s], i % d.W);
    const int o = gidx[i];
    if (o < 0 || o >= kImMaxG) return;
    float h = center[(size_t)c.cls * d.HW + i];
    for (int j = 0; j < c.n_ctr; ++j)
        if (c.ctr[j] == i) h = 1.f;  // accepted cluster centres read 1.0 (marked in place by the reference)
    atomicAdd(&icnt[o], 1);
    atomicAdd(&ipsum[o], (double)pmax[i]);
    atomicMax(&ikey[o], ((unsigned long long)orderable(h) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
}

// One block: sort the contours by (class, OpenCV label order), number their non-empty instances in that order and emit the
// table (label, score; :880-887).  rank[g] = index of centre g's instance in the output, -1 if it owns no pixel.
__global__ void __launch_bounds__(kRefMaxComp)
im_order_kernel(RefDims d, const RefComp* __restrict__ comps, const int* __restrict__ ncomp, const int* __restrict__ gbase,
                const int* __restrict__ gcount, const int* __restrict__ icnt, const double* __restrict__ ipsum,
                const unsigned long long* __restrict__ ikey, const int* __restrict__ minx, int* __restrict__ rank,
                int* __restrict__ out_label, double* __restrict__ out_score, int* __restrict__ n_out, int* __restrict__ status) {
    __shared__ unsigned long long key[kRefMaxComp];
    __shared__ int pre[kRefMaxComp];
    const int t = threadIdx.x;
    const int n = min(ncomp[0], kRefMaxComp);
    unsigned long long k = ~0ull;
    if (t < n) {
        const RefComp& c = comps[t];
        k = ((unsigned long long)c.cls << 44) | ((unsigned long long)((c.root / d.W) >> 1) << 28) |
            ((unsigned long long)(minx[t] >> 1) << 12) | (unsigned long long)t;
    }
    key[t] = k;
    __syncthreads();
    for (int size = 2; size <= kRefMaxComp; size <<= 1)  // bitonic sort, ascending
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int j = t ^ stride;
            if (j > t) {
                const bool up = (t & size) == 0;
                const unsigned long long a = key[t], b = key[j];
                if ((a > b) == up) {
                    key[t] = b;
                    key[j] = a;
                }
            }
            __syncthreads();
        }
    const int slot = (t < n) ? (int)(key[t] & 0xfffull) : -1;
    int g0 = 0, g1 = 0, mine = 0;
    if (slot >= 0) {
        g0 = gbase[slot];
        g1 = min(g0 + gcount[slot], kImMaxG);
        for (int g = g0; g < g1; ++g) mine += icnt[g] > 0;  // `if mask.sum() > 0`  :866
    }
    pre[t] = mine;
    __syncthreads();
    for (int off = 1; off < kRefMaxComp; off <<= 1) {  // inclusive scan
        const int v = (t >= off) ? pre[t - off] : 0;
        __syncthreads();
        pre[t] += v;
        __syncthreads();
    }
    if (t == kRefMaxComp - 1) {
        n_out[0] = pre[t];
        if (pre[t] > kImMaxInst) atomicOr(status, kRefListOverflow);
    }
    if (slot < 0) return;
    int idx = pre[t] - mine;
    const int cls = comps[slot].cls;
    for (int g = g0; g < g1; ++g) {
        if (icnt[g] == 0) {
            rank[g] = -1;
            continue;
        }
        rank[g] = idx < kImMaxInst ? idx : -1;
        if (idx < kImMaxInst) {
            const double seg_score = (double)(float)(ipsum[g] / (double)icnt[g]);  // fp32 .mean().item()
            double center_score = (double)unorderable((unsigned)(ikey[g] >> 32));
            if (center_score >= 1.0) center_score = seg_score;  // clustered centre: conf = seg_score  (:882-883)
            out_label[idx] = cls;
            out_score[idx] = center_score * seg_score;
        }
        ++idx;
    }
}

__global__ void im_pixel_map_kernel(RefDims d, const int* __restrict__ gidx, const int* __restrict__ rank,
                                    int* __restrict__ inst_map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.HW) return;
    const int g = gidx[i];
    inst_map[i] = (g >= 0 && g < kImMaxG) ? rank[g] : -1;
}

__global__ void im_masks_kernel(const int* __restrict__ inst_map, int n, int HW, unsigned char* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int k = inst_map[i];
    for (int j = blockIdx.y; j < n; j += gridDim.y) out[(size_t)j * HW + i] = (unsigned char)(j == k);
}

}  // namespace cl4

extern "C" int cl4_pseudo_labels(const long long* seg_gt, const float* cls_label, const float* peak_conf,
                                 const int* peak_y, const int* peak_x, int K, float pseudo_thresh, const float* gauss,
                                 int sigma, int min_area, float* out_center, float* out_offset, float* out_weight,
                                 int* total_match, int* status_out, int B, int C, int H, int W, void* scratch,
                                 size_t scratch_bytes, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 1 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), CL4_EINVAL, "pseudo_labels: bad shape");
    CL4_REQUIRE(B <= 65535, CL4_EUNSUPPORTED, "pseudo_labels: batch > 65535");
    CL4_REQUIRE(K >= 1 && sigma >= 0, CL4_EINVAL, "pseudo_labels: bad K or sigma");
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(seg_gt && cls_label && peak_conf && peak_y && peak_x && gauss && out_center && out_offset && out_weight &&
                    total_match && status_out, CL4_EINVAL, "pseudo_labels: null pointer");
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_refine_scratch_bytes(B, H, W), CL4_ESCRATCH, "pseudo_labels: scratch too small");
    RefDims d{B, C, H, W, H * W};
    RefScratch sc = ref_layout(reinterpret_cast<char*>(scratch), B, H, W);
    cudaStream_t s = (cudaStream_t)stream;
    REF_CUDA(cudaMemsetAsync(out_center, 0, sizeof(float) * (size_t)B * C * d.HW, s), "memset");
    REF_CUDA(cudaMemsetAsync(total_match, 0, sizeof(int) * (size_t)B, s), "memset");
    int rc = run_contours(seg_gt, cls_label, d, min_area, sc, s);
    if (rc != CL4_OK) return rc;
    dim3 lin(ceil_div(d.HW, 256), B);
    pl_match_kernel<<<dim3(ceil_div(C * K, 128), B), 128, 0, s>>>(peak_conf, peak_y, peak_x, cls_label, d, K, pseudo_thresh,
                                                                 sc.comp, sc.comps);
    pl_accept_kernel<<<dim3(ceil_div(kRefMaxComp, 128), B), 128, 0, s>>>(sc.comps, sc.ncomp, total_match);
    pl_write_kernel<<<lin, 256, 0, s>>>(d, sc.comp, sc.comps, out_offset, out_weight);
    ref_splat_kernel<<<dim3(kRefMaxComp, B), 256, 0, s>>>(d, gauss, sigma, 1, sc.comps, sc.ncomp, out_center);
    REF_CUDA(cudaMemcpyAsync(status_out, sc.status, 4, cudaMemcpyDeviceToDevice, s), "copy");
    return check_launch("pseudo_labels");
}

extern "C" int cl4_refine_max_contours(void) { return cl4::kRefMaxComp; }

extern "C" size_t cl4_refine_scratch_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return cl4::ref_layout(nullptr, B, H, W).bytes;
}

// 8-connected contours of the valid classes of a batch of label maps (the cv2 call of
// modules/utils.py:305-307 for every (image, class) at once).
extern "C" int cl4_contours8(const long long* gt_seg, const float* label, int min_area, int B, int C, int H, int W,
                             int* comp_out, int* info_out, int* ncomp_out, int* status_out, void* scratch,
                             size_t scratch_bytes, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 1 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), CL4_EINVAL, "contours8: bad shape");
    CL4_REQUIRE(B <= 65535, CL4_EUNSUPPORTED, "contours8: batch > 65535");
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(gt_seg && label && comp_out && info_out && ncomp_out && status_out, CL4_EINVAL, "contours8: null pointer");
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_refine_scratch_bytes(B, H, W), CL4_ESCRATCH, "contours8: scratch too small");
    RefDims d{B, C, H, W, H * W};
    RefScratch sc = ref_layout(reinterpret_cast<char*>(scratch), B, H, W);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = run_contours(gt_seg, label, d, min_area, sc, s);
    if (rc != CL4_OK) return rc;
    ref_export_comps_kernel<<<dim3(ceil_div(kRefMaxComp, 256), B), 256, 0, s>>>(sc.comps, sc.ncomp, info_out);
    REF_CUDA(cudaMemcpyAsync(comp_out, sc.comp, (size_t)B * d.HW * 4, cudaMemcpyDeviceToDevice, s), "copy");
    REF_CUDA(cudaMemcpyAsync(ncomp_out, sc.ncomp, (size_t)B * 4, cudaMemcpyDeviceToDevice, s), "copy");
    REF_CUDA(cudaMemcpyAsync(status_out, sc.status, 4, cudaMemcpyDeviceToDevice, s), "copy");
    return check_launch("contours8");
}

extern "C" int cl4_refine_labels(const float* seg_logits, const float* center, const float* offsets, const float* label,
                                 const long long* gt_seg, const float* gauss, int sigma, double refine_thresh,
                                 int nms_kernel, float beta, int min_area, int max_inst, long long top_k,
                                 float* out_center, float* out_offset, float* out_weight, int* status_out, int B, int C,
                                 int H, int W, void* scratch, size_t scratch_bytes, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(B >= 0 && C >= 1 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), CL4_EINVAL, "refine: bad shape");
    CL4_REQUIRE(B <= 65535, CL4_EUNSUPPORTED, "refine: batch > 65535");
    CL4_REQUIRE(nms_kernel > 0 && (nms_kernel & 1), CL4_EINVAL, "refine: nms kernel must be odd and positive");
    CL4_REQUIRE(sigma >= 0, CL4_EINVAL, "refine: sigma must be >= 0");
    CL4_REQUIRE(max_inst >= 0 && max_inst < kRefMaxInst, CL4_EUNSUPPORTED, "refine: max_inst %d >= %d", max_inst, kRefMaxInst);
    CL4_REQUIRE(refine_thresh >= 0.0, CL4_EUNSUPPORTED, "refine: negative threshold");
    if (B == 0) return CL4_OK;
    CL4_REQUIRE(seg_logits && center && offsets && label && gt_seg && gauss && out_center && out_offset && out_weight &&
                    status_out, CL4_EINVAL, "refine: null pointer");
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_refine_scratch_bytes(B, H, W), CL4_ESCRATCH, "refine: scratch too small");
    RefDims d{B, C, H, W, H * W};
    RefScratch sc = ref_layout(reinterpret_cast<char*>(scratch), B, H, W);
    cudaStream_t s = (cudaStream_t)stream;
    const int wpr = ceil_div(W, 32);
    dim3 lin(ceil_div(d.HW, 256), B), blk(32, 8), grd(ceil_div(W, 32), ceil_div(H, 8), B);
    dim3 wgrid(ceil_div(wpr, 8), H, B);

    REF_CUDA(cudaMemsetAsync(out_center, 0, sizeof(float) * (size_t)B * C * d.HW, s), "memset");
    int rc = run_contours(gt_seg, label, d, min_area, sc, s);
    if (rc != CL4_OK) return rc;

    // centre NMS per contour, ordered centre lists
    ref_nms_kernel<<<wgrid, 256, 0, s>>>(center, d, sc.comp, sc.comps, (float)refine_thresh, (nms_kernel - 1) / 2, wpr,
                                         sc.words);
    rc = check_launch("refine nms");
    if (rc != CL4_OK) return rc;
    rc = launch_center_compact(sc.words, B, H, wpr, sc.list_nms, sc.cnt_nms, kRefListCap, sc.row_off, s);
    if (rc != CL4_OK) return rc;
    ref_assign_centres_kernel<<<dim3(ceil_div(kRefListCap, 256), B), 256, 0, s>>>(d, sc.comp, sc.list_nms, sc.cnt_nms,
                                                                                  sc.comps, sc.status);
    // centre clustering (beta > 0)
    const int use_clusters = beta > 0.f;
    const float lo = 21.f - beta, hi = 21.f + beta;
    if (use_clusters) {
        REF_CUDA(cudaMemsetAsync(sc.area, 0, (char*)sc.root - (char*)sc.area, s), "memset");
        ref_weak_init_kernel<<<lin, 256, 0, s>>>(offsets, d, sc.comp, 2.5f, sc.root);
        ref_weak_merge4_kernel<<<grd, blk, 0, s>>>(d, sc.comp, sc.root);
        ref_weak_stats_kernel<<<grd, blk, 0, s>>>(d, sc.comp, sc.root, sc.area, sc.sx, sc.sy, sc.comps);
        ref_weak_select_kernel<<<wgrid, 256, 0, s>>>(d, sc.root, sc.area, lo, hi, wpr, sc.words);
        rc = check_launch("refine clusters");
        if (rc != CL4_OK) return rc;
        rc = launch_center_compact(sc.words, B, H, wpr, sc.list_cl, sc.cnt_cl, kRefListCap, sc.row_off, s);
        if (rc != CL4_OK) return rc;
    }
    ref_merge_clusters_kernel<<<dim3(ceil_div(kRefMaxComp, 128), B), 128, 0, s>>>(
        center, d, sc.comp, sc.area, sc.sx, sc.sy, sc.list_cl, sc.cnt_cl, use_clusters, lo, hi, top_k, sc.comps, sc.ncomp,
        sc.status);
    // grouping, instance statistics, outputs
    ref_group_kernel<<<lin, 256, 0, s>>>(offsets, d, sc.comp, sc.comps, sc.ids, 0);
    ref_inst_stats_kernel<<<lin, 256, 0, s>>>(seg_logits, center, label, d, sc.comp, sc.ids, max_inst, sc.comps);
    ref_finalize_kernel<<<dim3(ceil_div(kRefMaxComp * kRefMaxInst, 256), B), 256, 0, s>>>(d, refine_thresh, max_inst,
                                                                                          sc.comps, sc.ncomp);
    ref_write_kernel<<<lin, 256, 0, s>>>(d, sc.comp, sc.ids, max_inst, sc.comps, out_offset, out_weight);
    ref_splat_kernel<<<dim3(kRefMaxComp, B), 256, 0, s>>>(d, gauss, sigma, max_inst, sc.comps, sc.ncomp, out_center);
    REF_CUDA(cudaMemcpyAsync(status_out, sc.status, 4, cudaMemcpyDeviceToDevice, s), "copy");
    return check_launch("refine");
}

extern "C" int cl4_ins_map_max_instances(void) { return cl4::kImMaxInst; }

extern "C" size_t cl4_ins_map_scratch_bytes(int C, int H, int W) {
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    const size_t base = cl4::ref_layout(nullptr, 1, H, W).bytes;
    return cl4::im_layout(nullptr, base, C, H, W).bytes;
}

extern "C" int cl4_ins_map(const float* seg_logits, const float* center, float* offset0, const float* cls_label, int flip,
                           float scale_y, float scale_x, float val_thresh, int val_kernel, float beta, int ignore, int min_area,
                           long long* seg_map_out, int* inst_map_out, int* label_out, double* score_out, int* n_out,
                           int* status_out, int C, int H, int W, void* scratch, size_t scratch_bytes, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(C >= 1 && C < (1 << 19) && H > 0 && W > 0 && (long long)H * W < (1ll << 30) && H < 65536 && W < 65536, CL4_EINVAL,
                "ins_map: bad shape");
    CL4_REQUIRE(val_kernel > 0 && (val_kernel & 1), CL4_EINVAL, "ins_map: nms kernel must be odd and positive");
    CL4_REQUIRE(val_thresh >= 0.f, CL4_EUNSUPPORTED, "ins_map: negative threshold");
    CL4_REQUIRE(seg_logits && center && offset0 && seg_map_out && inst_map_out && label_out && score_out && n_out && status_out,
                CL4_EINVAL, "ins_map: null pointer");
    CL4_REQUIRE(scratch && scratch_bytes >= cl4_ins_map_scratch_bytes(C, H, W), CL4_ESCRATCH, "ins_map: scratch too small");
    RefDims d{1, C, H, W, H * W};
    RefScratch sc = ref_layout(reinterpret_cast<char*>(scratch), 1, H, W);
    ImScratch im = im_layout(reinterpret_cast<char*>(scratch), sc.bytes, C, H, W);
    cudaStream_t s = (cudaStream_t)stream;
    const int wpr = ceil_div(W, 32);
    dim3 lin(ceil_div(d.HW, 256), 1), blk(32, 8), grd(ceil_div(W, 32), ceil_div(H, 8), 1), wgrid(ceil_div(wpr, 8), H, 1);

    im_prepare_kernel<<<ceil_div(max(d.HW, C), 256), 256, 0, s>>>(seg_logits, center, offset0, cls_label, flip ? 1 : 0, scale_y,
                                                                 scale_x, d, seg_map_out, im.pmax, im.center_avg, im.ones);
    const float* heat = flip ? im.center_avg : center;
    int rc = run_contours(seg_map_out, im.ones, d, min_area, sc, s);  // every class present in seg_map (:838)
    if (rc != CL4_OK) return rc;
    ref_nms_kernel<<<wgrid, 256, 0, s>>>(heat, d, sc.comp, sc.comps, val_thresh, (val_kernel - 1) / 2, wpr, sc.words);
    rc = check_launch("ins_map nms");
    if (rc != CL4_OK) return rc;
    rc = launch_center_compact(sc.words, 1, H, wpr, sc.list_nms, sc.cnt_nms, kRefListCap, sc.row_off, s);
    if (rc != CL4_OK) return rc;
    const int lblk = ceil_div(kRefListCap, 256);
    im_assign_kernel<<<lblk, 256, 0, s>>>(d, sc.comp, sc.list_nms, sc.cnt_nms, sc.comps, im.cslot, im.crank, sc.status);
    im_scan_kernel<<<1, kRefMaxComp, 0, s>>>(sc.comps, sc.ncomp, 0, 0, im.nbase, nullptr, sc.status);
    im_scatter_kernel<<<lblk, 256, 0, s>>>(d, sc.list_nms, sc.cnt_nms, im.cslot, im.crank, im.nbase, im.nms_sorted);
    if (beta > 0.f) {  // centre clustering (cluster_peaks, dataset/utils.py:767-793)
        const float lo = 21.f - beta, hi = 21.f + beta;
        REF_CUDA(cudaMemsetAsync(sc.area, 0, (char*)sc.root - (char*)sc.area, s), "memset");
        ref_weak_init_kernel<<<lin, 256, 0, s>>>(offset0, d, sc.comp, 2.5f, sc.root);
        ref_weak_merge4_kernel<<<grd, blk, 0, s>>>(d, sc.comp, sc.root);
        ref_weak_stats_kernel<<<grd, blk, 0, s>>>(d, sc.comp, sc.root, sc.area, sc.sx, sc.sy, sc.comps);
        ref_weak_select_kernel<<<wgrid, 256, 0, s>>>(d, sc.root, sc.area, lo, hi, wpr, sc.words);
        rc = check_launch("ins_map clusters");
        if (rc != CL4_OK) return rc;
        rc = launch_center_compact(sc.words, 1, H, wpr, sc.list_cl, sc.cnt_cl, kRefListCap, sc.row_off, s);
        if (rc != CL4_OK) return rc;
        im_merge_clusters_kernel<<<ceil_div(kRefMaxComp, 128), 128, 0, s>>>(heat, d, sc.comp, sc.area, sc.sx, sc.sy, sc.list_cl,
                                                                            sc.cnt_cl, lo, hi, im.nms_sorted, im.nbase, sc.comps,
                                                                            sc.ncomp, sc.status);
    }
    const int empty_id = ignore ? 0 : 1;
    im_scan_kernel<<<1, kRefMaxComp, 0, s>>>(sc.comps, sc.ncomp, 1, empty_id, im.gbase, im.gcount, sc.status);
    im_group_kernel<<<lin.x, 256, 0, s>>>(offset0, d, sc.comp, sc.comps, im.nms_sorted, im.nbase, im.gbase, empty_id, im.gidx);
    REF_CUDA(cudaMemsetAsync(im.icnt, 0, (char*)im.minx - (char*)im.icnt, s), "memset");
    REF_CUDA(cudaMemsetAsync(im.minx, 0x7f, (size_t)kRefMaxComp * 4, s), "memset");
    im_stats_kernel<<<lin.x, 256, 0, s>>>(heat, im.pmax, d, sc.comp, im.gidx, sc.comps, im.icnt, im.ipsum, im.ikey, im.minx);
    im_order_kernel<<<1, kRefMaxComp, 0, s>>>(d, sc.comps, sc.ncomp, im.gbase, im.gcount, im.icnt, im.ipsum, im.ikey, im.minx,
                                              im.rank, label_out, score_out, n_out, sc.status);
    im_pixel_map_kernel<<<lin.x, 256, 0, s>>>(d, im.gidx, im.rank, inst_map_out);
    REF_CUDA(cudaMemcpyAsync(status_out, sc.status, 4, cudaMemcpyDeviceToDevice, s), "copy");
    return check_launch("ins_map");
}

extern "C" int cl4_ins_masks(const int* inst_map, int n, int H, int W, unsigned char* masks_out, cl4_stream_t stream) {
    using namespace cl4;
    CL4_REQUIRE(n >= 0 && H > 0 && W > 0, CL4_EINVAL, "ins_masks: bad shape");
    if (n == 0) return CL4_OK;
    CL4_REQUIRE(inst_map && masks_out, CL4_EINVAL, "ins_masks: null pointer");
    im_masks_kernel<<<dim3(ceil_div(H * W, 256), min(n, 64)), 256, 0, (cudaStream_t)stream>>>(inst_map, n, H * W, masks_out);
    return check_launch("ins_masks");
}
