// PAMR propagation sweep, TMA-staged (reference wss/modules.py:148-149).
//
// One persistent CTA per SM walks over 32x32-pixel output tiles (image-major order, so
// neighbouring CTAs share halos in L2).  For a tile, every thread keeps the 8*D affinity
// weights of its two pixels in registers for all C classes; per class the 80x80 source window
// (tile + 24-pixel halo) is brought into shared memory by ONE cp.async.bulk.tensor (TMA) box
// load, triple-buffered on mbarriers so that the copy of class c+2 overlaps the FMAs of class c.
// TMA zero-fills outside the image; tiles that touch the border then rewrite those cells with
// the clamped (replicate-padded, wss/modules.py:57) values, which are in the same window.
// The inner loop is one LDS + one FFMA per (pixel, tap) with compile-time shared-memory offsets.
#include "common.cuh"
#include "pamr_internal.cuh"
#include "tma.cuh"

namespace cl4 {

constexpr int kTile = 32;
constexpr int kHalo = 24;                       // largest supported dilation on this path
constexpr int kBox = kTile + 2 * kHalo;         // 80
constexpr int kStages = 4;
constexpr int kSweepThreads = 512;              // 16 warps; each thread owns two pixels 8 rows apart
constexpr int kRowGap = 8;                      // (y, y+8): the d=8 and d=4 taps of the pair share 7 sources
constexpr int kStageBytes = kBox * kBox * 4;    // 25600
constexpr int kMaxFix = kBox * kBox - 1;        // out-of-image cells of a window (always < box area)
constexpr size_t kSweepSmem = (size_t)kStages * kStageBytes + (size_t)kMaxFix * 4 + 128;

// compile-time dilation sets get immediate LDS offsets; RuntimeDil keeps them in registers
struct DilVoc6 {  // PAMR's class default (wss/modules.py:125)
    static constexpr bool kStatic = true;
    __host__ __device__ static constexpr int get(int i) {
        return i == 0 ? 1 : i == 1 ? 2 : i == 2 ? 4 : i == 3 ? 8 : i == 4 ? 12 : i == 5 ? 24 : 1;
    }
};
struct DilVoc5 {  // the trainer's setting (train.py:81)
    static constexpr bool kStatic = true;
    __host__ __device__ static constexpr int get(int i) {
        return i == 0 ? 1 : i == 1 ? 2 : i == 2 ? 4 : i == 3 ? 8 : i == 4 ? 12 : 1;
    }
};
struct DilRuntime {
    static constexpr bool kStatic = false;
    __host__ __device__ static constexpr int get(int) { return 1; }
};

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

template <int D, class DS>
__global__ void __launch_bounds__(kSweepThreads, 1)
pamr_sweep_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wts,
                      float* __restrict__ mout, int C, int H, int W, int tiles_x, int tiles_y, int n_tiles,
                      Dilations dil) {
    constexpr int P = 8 * D;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint32_t* fixlist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kStages * kStageBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kStageBytes + (size_t)kMaxFix * 4 + 4);
    uint64_t* empty = full + kStages;
    __shared__ int s_nfix;

    const int tid = threadIdx.x;
    const int lane = tid & 31, wrp = tid >> 5;
    const int tx = lane;
    const int ty = (wrp >> 3) * 16 + (wrp & 7);  // rows ty and ty + kRowGap
    const size_t HW = (size_t)H * W;
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

    // ---- producer (thread 0 only): p_item is the next (tile, class) item to issue
    int p_item = 0;
    auto issue_next = [&]() {
        const int pk = p_item / C, pc = p_item - pk * C;
        const TileCoord ptc = tile_coord(blockIdx.x + pk * gridDim.x, tiles_x, tiles_per_img);
        const int s = p_item % kStages;
        if (p_item >= kStages) mbar_wait(&empty[s], (uint32_t)((p_item / kStages - 1) & 1));
        mbar_arrive_expect_tx(&full[s], kStageBytes);
        tma_load_3d(stage0 + (size_t)s * (kBox * kBox), &tmap, &full[s], ptc.x0 - kHalo, ptc.y0 - kHalo,
                    ptc.b * C + pc);
        ++p_item;
    };
    if (tid == 0) {
        for (int i = 0; i < kStages - 1 && p_item < total; ++i) issue_next();
    }

    float w0[P], w1[P];
    const int sbase = (ty + kHalo) * kBox + (tx + kHalo);
    int item = 0;

    for (int k = 0; k < n_my; ++k) {
        const TileCoord tc = tile_coord(blockIdx.x + k * gridDim.x, tiles_x, tiles_per_img);
        const int x = tc.x0 + tx, ya = tc.y0 + ty;
        const bool valid0 = (x < W) && (ya < H);
        const bool valid1 = (x < W) && (ya + kRowGap < H);
        const bool border = (tc.x0 < kHalo) || (tc.y0 < kHalo) || (tc.x0 + kTile + kHalo > W) ||
                            (tc.y0 + kTile + kHalo > H);
        {  // weights of this thread's two pixels, kept in registers for all classes of the tile
            const float* wp = wts + (size_t)tc.b * P * HW + (size_t)ya * W + x;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                w0[p] = valid0 ? __ldg(wp + (size_t)p * HW) : 0.f;
                w1[p] = valid1 ? __ldg(wp + (size_t)p * HW + (size_t)kRowGap * W) : 0.f;
            }
        }
        if (k + 1 < n_my) {  // pull the next tile's weights (P planes x 32 rows x 128 B) into L2
            const TileCoord nt = tile_coord(blockIdx.x + (k + 1) * gridDim.x, tiles_x, tiles_per_img);
            const float* nb = wts + (size_t)nt.b * P * HW + (size_t)nt.y0 * W + nt.x0;
            for (int i = tid; i < P * kTile; i += kSweepThreads) {
                const int p = i >> 5, r = i & 31;
                if (nt.y0 + r < H) prefetch_l2(nb + (size_t)p * HW + (size_t)r * W);
            }
        }
        int n_fix = 0;
        if (border) {
            // Replicate padding (wss/modules.py:57): list the window cells that fall outside the
            // image together with the in-image cell they copy; the list is reused for every class.
            if (tid == 0) s_nfix = 0;
            __syncthreads();
            for (int i = tid; i < kBox * kBox; i += kSweepThreads) {
                const int by = i / kBox, bx = i - by * kBox;
                const int gy = tc.y0 - kHalo + by, gx = tc.x0 - kHalo + bx;
                const int cy = clampi(gy, 0, H - 1), cx = clampi(gx, 0, W - 1);
                if (cy != gy || cx != gx) {
                    const int src = (cy - (tc.y0 - kHalo)) * kBox + (cx - (tc.x0 - kHalo));
                    fixlist[atomicAdd(&s_nfix, 1)] = ((uint32_t)i << 16) | (uint32_t)src;
                }
            }
            __syncthreads();
            n_fix = s_nfix;
        }
        float* out0 = mout + (size_t)tc.b * C * HW + (size_t)ya * W + x;

        for (int c = 0; c < C; ++c, ++item) {
            if (tid == 0 && p_item < total) issue_next();  // refills the stage released by item-1

            const int s = item % kStages;
            float* sm = stage0 + (size_t)s * (kBox * kBox);
            mbar_wait(&full[s], (uint32_t)((item / kStages) & 1));

            if (border) {
                for (int i = tid; i < n_fix; i += kSweepThreads) {
                    const uint32_t e = fixlist[i];
                    sm[e >> 16] = sm[e & 0xffffu];
                }
                fence_proxy_async_smem();  // these generic writes precede a later TMA refill of the stage
                __syncthreads();
            }

            float a0 = 0.f, a1 = 0.f;
            const float* sp = sm + sbase;
#pragma unroll
            for (int di = 0; di < D; ++di) {
                const int d = DS::kStatic ? DS::get(di) : dil.d[di];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int dy = (j < 3) ? -1 : ((j < 5) ? 0 : 1);
                    const int dx = (j < 3) ? (j - 1) : ((j == 3) ? -1 : ((j == 4) ? 1 : (j - 6)));
                    const int off = dy * d * kBox + dx * d;
                    a0 = fmaf(w0[di * 8 + j], sp[off], a0);
                    a1 = fmaf(w1[di * 8 + j], sp[off + kRowGap * kBox], a1);
                }
            }
            if (valid0) out0[(size_t)c * HW] = a0;
            if (valid1) out0[(size_t)c * HW + (size_t)kRowGap * W] = a1;

            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // this warp no longer reads the stage
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
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

bool sweep_tma_applicable(int C, int H, int W, const Dilations& dil, int D, const float* mask_in) {
    if (W % 4 != 0 || ((uintptr_t)mask_in & 15) != 0) return false;  // TMA: 16-byte global strides / base
    if ((long long)H * W < 32 * 32) return false;                    // tiny maps: the register kernel is fine
    (void)C;
    for (int i = 0; i < D; ++i)
        if (dil.d[i] > kHalo) return false;
    return true;
}

template <int D, class DS>
static int launch_one(const CUtensorMap& tmap, const float* w, float* mo, int C, int H, int W, int tiles_x, int tiles_y,
                      int n_tiles, const Dilations& dil, cudaStream_t s) {
    auto kern = pamr_sweep_tma_kernel<D, DS>;
    static bool attr_done = false;  // per instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem);
        if (e != cudaSuccess) {
            set_error("pamr_sweep_tma: smem attribute: %s", cudaGetErrorString(e));
            return CL4_ECUDA;
        }
        attr_done = true;
    }
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<grid, kSweepThreads, kSweepSmem, s>>>(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil);
    return check_launch("pamr_sweep_tma");
}

template <int D>
struct TmaSweepLauncher {
    static int run(const CUtensorMap& tmap, const float* w, float* mo, int C, int H, int W, int tiles_x, int tiles_y,
                   int n_tiles, const Dilations& dil, cudaStream_t s) {
        bool voc6 = (D == 6), voc5 = (D == 5);
        for (int i = 0; i < D && i < 6; ++i) {
            voc6 = voc6 && dil.d[i] == DilVoc6::get(i);
            voc5 = voc5 && dil.d[i] == DilVoc5::get(i);
        }
        if (D == 6 && voc6) return launch_one<6, DilVoc6>(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        if (D == 5 && voc5) return launch_one<5, DilVoc5>(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        return launch_one<D, DilRuntime>(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
    }
};

int launch_sweep_tma(const float* w, const float* mi, float* mo, int B, int C, int H, int W, const Dilations& dil, int D,
                     cudaStream_t s) {
    CUtensorMap tmap;
    const int rc = encode_tmap_3d_f32(&tmap, mi, W, H, (long long)B * C, kBox, kBox);
    if (rc != 0) {
        set_error("pamr_sweep_tma: cuTensorMapEncodeTiled failed (%d)", rc);
        return CL4_ECUDA;
    }
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile);
    const int n_tiles = B * tiles_x * tiles_y;
    switch (D) {
        case 1: return TmaSweepLauncher<1>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 2: return TmaSweepLauncher<2>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 3: return TmaSweepLauncher<3>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 4: return TmaSweepLauncher<4>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 5: return TmaSweepLauncher<5>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 6: return TmaSweepLauncher<6>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 7: return TmaSweepLauncher<7>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
        case 8: return TmaSweepLauncher<8>::run(tmap, w, mo, C, H, W, tiles_x, tiles_y, n_tiles, dil, s);
    }
    set_error("pamr_sweep_tma: bad D=%d", D);
    return CL4_EUNSUPPORTED;
}

}  // namespace cl4
