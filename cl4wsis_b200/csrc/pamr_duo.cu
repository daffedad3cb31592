// PAMR propagation sweep on CLASS PAIRS ("duo" sweep) for the dilation sets [1,2,4,8,12,24] (PAMR's class default,
// reference wss/modules.py:125) and [1,2,4,8,12] (train.py:81); the sweep itself is wss/modules.py:147-149.
//
// Why.  The lattice sweep of pamr_lattice.cu is bound by instruction latency at two warps per scheduler: per (tile, class)
// a warp issues 192 FFMA + 76..112 LDS + loop / barrier overhead and reaches an IPC of 0.26 (profiles/r01c_notes.md).  The
// affinity weight of a (pixel, tap) pair is the same for every class, so two classes can share every instruction of
// that stream if their values sit next to each other in shared memory:
//   * between sweeps the masks live in a pair-interleaved layout [B][ceil(C/2)][H][W][2] (classes 2q and 2q+1 of a pixel
//     are one 8-byte cell), so the TMA window of an item is 80 x 84 CELLS and one LDS.64 / LDS.128 fetches the source of
//     both classes;
//   * the FMA is the packed fp32 FFMA2 of sm_100 (PTX fma.rn.f32x2) with the weight as its scalar operand
//     (SASS `FFMA2 Rd, Rv.F32x2.HI_LO, Rw.F32, Rd.F32x2.HI_LO`): one issue slot for both classes, IEEE fp32 FMAs
//     as before.
// Per class this halves the FFMA, LDS, barrier, branch and address instructions; the shared-memory bytes per class stay
// what they were (the ownership of pamr_lattice.cu is kept: group A dilations {4,8,12} on 2 x 4 lattice blocks of
// spacing 4, group B dilations {1,2,24} on 4 x 2 blocks of adjacent pixels, A's partial sums handed to B per warp pair).
//   * tensor memory holds group B's 128 dilation-{1,2} weights per thread (tcgen05.st once per tile, tcgen05.ld per item:
//     SASS STTM / LDTM).  Why: group B is the critical path, and its dilation-24 part has no source reuse inside a 32 x 32
//     tile -- 32 LDS.128 for 64 FFMA2 -- so it needs many loads in flight, but with 192 resident weights a thread has ~30
//     free registers = 8 LDS.128.  The near weights are read once per item in a fixed order, so they can be streamed from
//     tensor memory in chunks of 16 while they are needed and occupy no registers during the dilation-24 part, which then
//     keeps all of its 32 loads in flight.  (Ablation with aliased weight registers: 0.49 -> 0.41 ms per sweep,
//     profiles/r02_notes.md.)  The columns are double buffered: the next tile's near weights are installed in the other
//     half while the current tile computes, so a tile switch costs nothing.
// An odd class count costs one dummy class (C = 21: +4.8 % work); its cells are zero and stay zero.
//
// The first sweep of a PAMR call is the one-class kernel of pamr_lattice.cu reading the caller's planar masks and writing
// cells; the last sweep writes the planar [B,C,H,W] result directly.
#include "pamr_lattice.cuh"

namespace cl4 {

#ifndef CL4_DUO_STAGES
#define CL4_DUO_STAGES 3
#endif
#ifndef CL4_DUO_PARTS
#define CL4_DUO_PARTS 2
#endif

typedef unsigned long long u64;

constexpr int kDStages = CL4_DUO_STAGES;
constexpr int kDParts = CL4_DUO_PARTS;
constexpr int kDPitch = kLPitch;                   // window pitch in cells (84: A's half-warps hit 16 different bank pairs)
constexpr int kDStageCells = kBox * kDPitch;       // 6720 cells
constexpr int kDStageBytes = kDStageCells * 8;     // 53760 = 420 * 128
constexpr int kDPartPitch = kLPartPitch;           // 36 cells
constexpr int kDPartCells = kTile * kDPartPitch;   // 1152 cells
#ifndef CL4_DUO_TMEM
#define CL4_DUO_TMEM 1  // group B's dilation-{1,2} weights live in tensor memory (0: in registers, as pamr_lattice.cu)
#endif
constexpr bool kDTmem = CL4_DUO_TMEM != 0;
constexpr int kDNearCols = 128;                    // TMEM columns of the near weights: row i * 32 + dilation h * 16 + column j * 8 + tap
constexpr int kDFarCols = 64;                      // ... of the dilation-24 weights: row i * 16 + column j * 8 + tap
constexpr int kDTileCols = 256;                    // columns per tile (192 used); two tiles: the current one and the one visited next
constexpr int kDTmemCols = 2 * kDTileCols;
constexpr int kDBatches = 6;                       // installation batches of 8 float4 groups per tile: 4 near + 2 far
constexpr size_t kDSmem = (size_t)kDStages * kDStageBytes + (size_t)kDParts * kDPartCells * 8 + (3 * kDStages + 8 * kDParts + 1) * 8 + 64;
static_assert(kDSmem <= 227 * 1024, "duo sweep: shared memory");

// acc.{lo,hi} += v.{lo,hi} * w      (two IEEE fp32 FMAs, one instruction; the weight is FFMA2's scalar operand)
__device__ __forceinline__ void ffma2(u64& acc, const u64 v, const float w) {
    asm("{\n\t.reg .b64 ww;\n\tmov.b64 ww, {%2, %2};\n\tfma.rn.f32x2 %0, %1, ww, %0;\n\t}" : "+l"(acc) : "l"(v), "f"(w));
}
__device__ __forceinline__ u64 fadd2(const u64 a, const u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float2 unpack2(const u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// ---- tensor memory as thread-private storage: lane = thread of the warp's quadrant, column = index ----
__device__ __forceinline__ void tmem_ld16(float (&r)[16], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
          "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
// the loaded registers may be read after this (they are operands so that no use is scheduled above the wait)
__device__ __forceinline__ void tmem_wait_ld(float (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]),
                   "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float4 a, const float4 b, const float4 c, const float4 d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::"f"(a.x),
        "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "f"(c.x), "f"(c.y), "f"(c.z), "f"(c.w), "f"(d.x),
        "f"(d.y), "f"(d.z), "f"(d.w), "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One source cell `v` at lattice position (r, c) of an A x B block whose taps are the steps 1..NS of the lattice: feed every
// (pixel, tap) that reads it.  Weight register of (pixel slot, step s, tap): slot*24 + (s-1)*8 + tap (as pamr_lattice.cu).
template <int G, int A, int B, int NS, bool kReload>
__device__ __forceinline__ void feed2(float (&w)[kLW], u64 (&acc)[kLPx], const u64 v, const int r, const int c,
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
                    ffma2(acc[i * B + j], v, w[k]);
                    // sources arrive in row-major order, so taps 3 and 7 are the last uses of their float4: refill it
                    // with the next tile's weights right away (kReload: last item of a tile)
                    if (kReload && (k & 3) == 3) load_weight_group<G>(w, nw, k >> 2);
                }
            }
}

// group A: 2 x 4 lattice block of spacing 4, dilations 4, 8, 12 = steps 1, 2, 3; sp points at the cell of the block's
// pixel (0, 0) inside the window.  76 of the 8 x 10 lattice positions are read (76 LDS.64 for 192 FFMA2).
template <bool kReload>
__device__ __forceinline__ void duo_a(float (&w)[kLW], u64 (&acc)[kLPx], const u64* __restrict__ sp, const float4* __restrict__ nw) {
#pragma unroll
    for (int r = -3; r < 2 + 3; ++r)
#pragma unroll
        for (int c = -3; c < 4 + 3; ++c)
            if (source_needed<2, 4, 3>(r, c)) {
                const u64 v = sp[r * 4 * kDPitch + c * 4];
                feed2<0, 2, 4, 3, kReload>(w, acc, v, r, c, nw);
            }
}

// group B: 4 x 2 block of adjacent pixels; one LDS.128 = the cells of two adjacent pixels.
// Dilation 24 (weights slot*24 + 16 + tap, always in registers): the two pixels of a row share one LDS.128 per tap (32 LDS.128
// for 64 FFMA2, no reuse).
template <bool kReload>
__device__ __forceinline__ void duo_b_far(float (&w)[kLW], u64 (&acc)[kLPx], const u64* __restrict__ sp, const float4* __restrict__ nw) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int a = -1; a <= 1; ++a)
#pragma unroll
            for (int b = -1; b <= 1; ++b) {
                if (a == 0 && b == 0) continue;
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sp + (i + 24 * a) * kDPitch + 24 * b);
                const int k0 = (i * 2) * kLTaps + 16 + tap_index(a, b), k1 = k0 + kLTaps;
                ffma2(acc[i * 2], v.x, w[k0]);
                ffma2(acc[i * 2 + 1], v.y, w[k1]);
                if (kReload && (k0 & 3) == 3) {
                    load_weight_group<1>(w, nw, k0 >> 2);
                    load_weight_group<1>(w, nw, k1 >> 2);
                }
            }
}
// Dilations 1 and 2 with the weights in registers: rows -2..5, columns -2..3 (24 LDS.128 for 128 FFMA2).
template <bool kReload>
__device__ __forceinline__ void duo_b_near(float (&w)[kLW], u64 (&acc)[kLPx], const u64* __restrict__ sp, const float4* __restrict__ nw) {
#pragma unroll
    for (int r = -2; r < 4 + 2; ++r)
#pragma unroll
        for (int c = -2; c < 4; c += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sp + r * kDPitch + c);
            feed2<1, 4, 2, 2, kReload>(w, acc, v.x, r, c, nw);
            feed2<1, 4, 2, 2, kReload>(w, acc, v.y, r, c + 1, nw);
        }
}
// The same with the weights streamed from tensor memory.  Chunk (i, h) = pixel row i of the block, dilation 2^h: 16 columns
// [pixel j = 0: taps 0..7][j = 1: taps 0..7]; the next chunk is in flight while this one computes.  A tcgen05.wait::ld keeps
// later shared-memory loads below it, so the source rows are loaded explicitly ahead of the waits: rows i-2 .. i+2 of the
// window are resident (six row slots of three LDS.128), row i+3 is fetched at the start of row i.
__device__ __forceinline__ void duo_b_near_tmem(u64 (&acc)[kLPx], const u64* __restrict__ sp, const uint32_t tnear) {
    float wf[2][16];
    ulonglong2 row[6][3];
    tmem_ld16(wf[0], tnear);
#pragma unroll
    for (int r = -2; r <= 2; ++r)
#pragma unroll
        for (int pi = 0; pi < 3; ++pi) row[r + 2][pi] = *reinterpret_cast<const ulonglong2*>(sp + r * kDPitch + 2 * pi - 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i + 3 <= 5) {
#pragma unroll
            for (int pi = 0; pi < 3; ++pi)
                row[(i + 5) % 6][pi] = *reinterpret_cast<const ulonglong2*>(sp + (i + 3) * kDPitch + 2 * pi - 2);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * i + h, s = h + 1;
            tmem_wait_ld(wf[c & 1]);
            if (c + 1 < 8) tmem_ld16(wf[(c + 1) & 1], tnear + 16 * (c + 1));
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int dy = (t < 3) ? -1 : ((t < 5) ? 0 : 1);
                const int dx = (t < 3) ? (t - 1) : ((t == 3) ? -1 : ((t == 4) ? 1 : (t - 6)));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = i + s * dy, col = j + s * dx + 2;  // window row -2..5, column 0..5 of the 6-cell row
                    const ulonglong2 pr = row[(r + 2) % 6][col >> 1];
                    ffma2(acc[i * 2 + j], (col & 1) ? pr.y : pr.x, wf[c & 1][j * 8 + t]);
                }
            }
        }
    }
}
// Dilation 24 with the weights streamed from tensor memory: chunk i = pixel row i of the block, 16 columns
// [pixel j = 0: taps 0..7][j = 1: taps 0..7].  The eight LDS.128 of a row are issued two rows ahead of their use (a
// tcgen05.wait::ld keeps later shared-memory loads below it, so the loads are placed explicitly).
__device__ __forceinline__ void duo_b_far_tmem(u64 (&acc)[kLPx], const u64* __restrict__ sp, const uint32_t tfar) {
    float wf[2][16];
    ulonglong2 src[3][8];
    tmem_ld16(wf[0], tfar);
    auto load_row = [&](ulonglong2 (&d)[8], const int i) {
#pragma unroll
        for (int a = -1; a <= 1; ++a)
#pragma unroll
            for (int b = -1; b <= 1; ++b)
                if (a != 0 || b != 0) d[tap_index(a, b)] = *reinterpret_cast<const ulonglong2*>(sp + (i + 24 * a) * kDPitch + 24 * b);
    };
    load_row(src[0], 0);
    load_row(src[1], 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i + 2 < 4) load_row(src[(i + 2) % 3], i + 2);
        tmem_wait_ld(wf[i & 1]);
        if (i + 1 < 4) tmem_ld16(wf[(i + 1) & 1], tfar + 16 * (i + 1));
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            ffma2(acc[i * 2], src[i % 3][t].x, wf[i & 1][t]);
            ffma2(acc[i * 2 + 1], src[i % 3][t].y, wf[i & 1][8 + t]);
        }
    }
}
// Installation of a tile's weights in tensor memory, in batches of eight float4 groups (32 registers in flight):
// batch b < 4: near groups 8b .. 8b+7 (slots 2b and 2b+1, four groups each) -> the chunks (b, 0), (b, 1);
// batch 4, 5:  far groups 8(b-4) .. +7 (slots 4(b-4) .. +3, two groups each) -> the far chunks 2(b-4), 2(b-4)+1.
// wp: the thread's base pointer into the tile's weight block (group B: its near region; the far region follows it).
__device__ __forceinline__ void tile_fetch(float4 (&g)[8], const float4* __restrict__ wp, const int b) {
#pragma unroll
    for (int n = 0; n < 8; ++n) g[n] = __ldg(wp + (8 * b + n) * kLGroupThreads);  // near groups 0..31, then far groups 32..47
}
__device__ __forceinline__ void tile_install(const float4 (&g)[8], const int b, const uint32_t tt) {
    if (b < 4) {
        tmem_st16(tt + 32 * b, g[0], g[1], g[4], g[5]);
        tmem_st16(tt + 32 * b + 16, g[2], g[3], g[6], g[7]);
    } else {
        tmem_st16(tt + kDNearCols + 32 * (b - 4), g[0], g[1], g[2], g[3]);
        tmem_st16(tt + kDNearCols + 32 * (b - 4) + 16, g[4], g[5], g[6], g[7]);
    }
}

struct DuoOut {
    float* ptr;       // pair layout: cell (plane 0, y = 0, x = 0); planar: element (plane 0, 0, 0) of [B*C][H][W]
    long long plane;  // floats between pair planes / class planes
    int pitch;        // floats between rows
    int planar;       // 0: pair-interleaved cells (the next sweep's input); 1: the caller's planar [B,C,H,W] (last sweep)
};

struct DuoCtx {
    u64* stage0;
    u64* part;  // [kDParts buffers][kDPartCells]; pfull / pempty: [kDParts][4 warp pairs]
    uint64_t *full, *ready, *empty, *pfull, *pempty;  // full: TMA landed; ready: border patched (what the compute warps wait for)
    uint32_t tmem_base;      // 256 columns x 128 lanes of tensor memory (group B's near weights: current tile, next tile)
    const float* wts;
    int C, Cp, H, W, tiles_x, tiles_per_img, n_tiles, n_my, total, s0;
    const int* meta;  // shared memory: {n_my, s0, number of visits}: what the visit loop of the compute warps reads instead of keeping it
};

// The producer (one thread of the third warpgroup): waits until all eight compute warps have released the stage it is about
// to refill and issues the window's TMA box.  Same scheme as pamr_lattice.cu; the box is 168 floats (84 cells) wide.
__device__ __forceinline__ void duo_producer(const DuoCtx& cx, const CUtensorMap* tmap) {
    const int Cp = cx.Cp;
    int k = 0, q = cx.s0, s = 0;
    uint32_t phase = 1;  // parity of the previous use of the stage
    LTile tc = ltile(blockIdx.x, cx.tiles_x, cx.tiles_per_img);
    for (int p_item = 0; p_item < cx.total; ++p_item) {
        if (p_item >= kDStages) mbar_wait_relaxed(&cx.empty[s], phase);
        mbar_arrive_expect_tx(&cx.full[s], kDStageBytes);
        tma_load_3d(cx.stage0 + (size_t)s * kDStageCells, tmap, &cx.full[s], 2 * (tc.x0 - kHalo), tc.y0 - kHalo, tc.b * Cp + q);
        if (++s == kDStages) {
            s = 0;
            phase ^= 1u;
        }
        if (++q == Cp) {
            q = 0;
            const int nk = (k + 1 == cx.n_my) ? 0 : k + 1;
            if (nk != k) {
                k = nk;
                tc = ltile(blockIdx.x + k * gridDim.x, cx.tiles_x, cx.tiles_per_img);
            }
        }
    }
}

// Replicate padding inside the ring (reference wss/modules.py:57: F.pad(mode="replicate")): the TMA unit zero-fills the
// cells of a window that lie outside the plane; the three other warps of the producer warpgroup overwrite them with the
// clamped neighbour (a cell of the same window) before the compute warps see the window.  Interior tiles need nothing.
__device__ __forceinline__ void duo_patcher(const DuoCtx& cx) {
    const int Cp = cx.Cp;
    const int lane = threadIdx.x & 31, pw = (threadIdx.x >> 5) - (kLThreads / 32 + 1);  // patch warp 0..2
    int k = 0, q = cx.s0, s = 0;
    uint32_t phase = 0;
    int rv0 = 0, rv1 = kBox, cv0 = 0, cv1 = kDPitch;
    auto enter_tile = [&](int kk) {
        const LTile tc = ltile(blockIdx.x + kk * gridDim.x, cx.tiles_x, cx.tiles_per_img);
        // valid window rows [rv0, rv1) and columns [cv0, cv1): the part of the 80 x 84 window that lies inside the image
        rv0 = max(0, kHalo - tc.y0);
        rv1 = min(kBox, cx.H - tc.y0 + kHalo);
        cv0 = max(0, kHalo - tc.x0);
        cv1 = min(kDPitch, cx.W - tc.x0 + kHalo);
    };
    if (cx.total > 0) enter_tile(0);
    for (int item = 0; item < cx.total; ++item) {
        mbar_wait_relaxed(&cx.full[s], phase);  // suspended, not spinning
        if (rv0 > 0 || rv1 < kBox || cv0 > 0 || cv1 < kDPitch) {
            u64* win = cx.stage0 + (size_t)s * kDStageCells;
            const int t = pw * 32 + lane;  // 0..95
            // columns left / right of the image: one thread per valid window row, independent stores of one cell
            const int r = rv0 + t;
            if (r < rv1) {
                u64* row = win + r * kDPitch;
                if (cv0 > 0) {
                    const u64 v = row[cv0];
#pragma unroll 8
                    for (int c = 0; c < cv0; ++c) row[c] = v;
                }
                if (cv1 < kDPitch) {
                    const u64 v = row[cv1 - 1];
#pragma unroll 4
                    for (int c = cv1; c < kDPitch; ++c) row[c] = v;
                }
            }
            // rows above / below the image: one thread per window column; the source is the clamped column of the first /
            // last image row, a cell the loop above does not write
            if (t < kDPitch) {
                const int cs = min(max(t, cv0), cv1 - 1);
                if (rv0 > 0) {
                    const u64 v = win[rv0 * kDPitch + cs];
#pragma unroll 8
                    for (int y = 0; y < rv0; ++y) win[y * kDPitch + t] = v;
                }
                if (rv1 < kBox) {
                    const u64 v = win[(rv1 - 1) * kDPitch + cs];
#pragma unroll 4
                    for (int y = rv1; y < kBox; ++y) win[y * kDPitch + t] = v;
                }
            }
            fence_proxy_async_smem();  // these generic-proxy writes are ordered before the next TMA refill of the stage
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.ready[s]);
        if (++s == kDStages) {
            s = 0;
            phase ^= 1u;
        }
        if (++q == Cp) {
            q = 0;
            const int nk = (k + 1 == cx.n_my) ? 0 : k + 1;
            if (nk != k) {
                k = nk;
                enter_tile(k);
            }
        }
    }
}

// The item loop of one warp group (G = 0: A, G = 1: B).  A CTA walks its tiles blockIdx.x, blockIdx.x + grid, ... in VISITS:
// a visit is a run of consecutive class pairs of one tile.  The first visit starts at pair s0 (CTAs start at staggered pair
// phases so that their weight refills do not coincide) and the walk wraps around to finish pairs 0 .. s0-1 of the first tile.
// Loop state is kept small on purpose (the compute warps run at the register cap and every spilled counter costs an exposed
// local-memory load at the item boundary): ring stage + phase; everything about the partial-sum buffers follows from the
// item counter.
template <int G, bool kFar>
__device__ __forceinline__ void duo_group(const DuoCtx& cx, const DuoOut& out) {
    static_assert((kDParts & (kDParts - 1)) == 0, "partial-sum buffers: a power of two");
    int tid = threadIdx.x;
    // kept in a register (opaque to the compiler): rematerialised, it costs an S2R plus its ~25 cycles of latency in front of
    // the weight-batch address of every item, on group B's critical path (-1.2 % per sweep)
    asm volatile("" : "+r"(tid));
    tid = __shfl_sync(0xffffffffu, tid, tid & 31);  // opaque to ptxas as well (it rematerialises special-register reads): -0.3 %
    const int lane = tid & 31;
    const int tg = tid - G * kLGroupThreads;
    constexpr bool kT = (G == 1) && kFar && kDTmem;  // near weights in tensor memory

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
    const u64* const sp0 = cx.stage0 + (ry + kHalo) * kDPitch + rx + kHalo;  // window cell of the block origin, stage 0
    u64* const pp0 = cx.part + ry * kDPartPitch + rx;                         // partial-sum cell, buffer 0
    const uint32_t ready0 = smem_u32(cx.ready), empty0 = smem_u32(cx.empty);  // "full" for the compute warps = patched
    const uint32_t pfull0 = smem_u32(cx.pfull) + 8u * ((tid >> 5) & 3), pempty0 = smem_u32(cx.pempty) + 8u * ((tid >> 5) & 3);
    const float4* const wbase = reinterpret_cast<const float4*>(cx.wts) + weight_thread_base(tid);
    auto weight_ptr = [&](int tile) -> const float4* { return wbase + (size_t)tile * (kLWeightsPerTile / 4); };

    float w[kLW];  // (group B with its weights in tensor memory never touches it)
    u64 acc[kLPx];
    // group B, weights in tensor memory: this warp's lane quadrant, columns [0,256) and [256,512) alternate between visits
    uint32_t tcur = cx.tmem_base + ((uint32_t)((tid >> 5) & 3) << 21), tnext = tcur + kDTileCols;

    // The visit sequence is a function of the visit number v alone: visit 0 = pairs s0 .. Cp-1 of the CTA's first tile, visits
    // 1 .. n_my-1 = all pairs of its other tiles, and (s0 > 0) a last visit = pairs 0 .. s0-1 of the first tile again.  The loop
    // keeps v and the item counter in registers and re-reads n_my / s0 from shared memory at every visit: with tile, pair and
    // remaining-item counters live across the item loop ptxas spilled them, and a spilled word comes back from DRAM at a visit
    // boundary (the weight stream evicts local memory from L1 and L2): ~1 000 cycles per reload, 4.5 % of group A's time.
    const volatile int* const meta = cx.meta;
    int tile = blockIdx.x;
    if (meta[2] > 0) {
        if (kT) {
#pragma unroll 1
            for (int b = 0; b < kDBatches; ++b) {
                float4 g8[8];
                tile_fetch(g8, weight_ptr(tile), b);
                tile_install(g8, b, tcur);
            }
            tmem_wait_st();
        } else {
            load_weights<G, (G == 1 && !kFar)>(w, weight_ptr(tile));
        }
    }

    uint32_t it = 0;  // items done: partial-sum buffer it % kDParts, its phase (it / kDParts) & 1
    int stage = 0;
    uint32_t full_phase = 0;
#pragma unroll 1
    for (int v = 0;; ++v) {
        const int n_my = meta[0], s0 = meta[1];
        if (v >= meta[2]) break;
        // ---- one visit: pairs q .. q + n_q - 1 of `tile`
        const int Cp = (cx.C + 1) >> 1;
        const int q = (v == 0) ? s0 : 0;
        const int n_q = (v == 0) ? Cp - s0 : ((v < n_my) ? Cp : s0);
        const int nk = (v + 1 < n_my) ? v + 1 : 0;
        const int next_tile = (int)blockIdx.x + nk * (int)gridDim.x;
        const bool switch_tile = (v + 1 < meta[2]) && (next_tile != tile);  // another tile follows: its weights are fetched during this visit
        const float4* const nw = weight_ptr(next_tile);
        // The next visit's weights into L2 a whole visit ahead.  Each group's 96 KB of a tile are one contiguous run, so one
        // thread per group issues one bulk prefetch (a prefetch.global.L2 per 128-byte line from every eighth lane cost 2.7 % of
        // group A's time in CCTL issue and LSU-queue stalls: 0.4186 -> 0.4113 ms per sweep; without any prefetch: 0.4328).
        if (switch_tile && tg == 0)
            bulk_prefetch_l2(reinterpret_cast<const float4*>(cx.wts) + (size_t)next_tile * (kLWeightsPerTile / 4) + G * kLGroupsA * kLGroupThreads,
                             (uint32_t)(((G == 0 || kFar) ? kLGroupsA : kLGroupsBNear) * kLGroupThreads * 16));
        // group B stores the pixels of its 4 x 2 block: pointer to the block origin in this visit's first plane, valid rows
        float* oc = nullptr;
        int nrows = 0;
        if (G == 1) {
            const LTile tc = ltile(tile, cx.tiles_x, cx.tiles_per_img);
            const int y = tc.y0 + ry, x = tc.x0 + rx;
            nrows = (x < cx.W) ? min(max(cx.H - y, 0), 4) : 0;
            oc = out.planar ? out.ptr + ((long long)tc.b * cx.C + 2 * q) * out.plane + (long long)y * out.pitch + x
                            : out.ptr + ((long long)tc.b * cx.Cp + q) * out.plane + (long long)y * out.pitch + 2 * x;
        }

        for (int e = 0; e < n_q; ++e) {
            const bool reload = switch_tile && (e == n_q - 1);
            const u64* const sp = sp0 + stage * kDStageCells;
            u64* const pp = pp0 + (it & (kDParts - 1)) * kDPartCells;
            const uint32_t pb8 = 32u * (it & (kDParts - 1)), part_phase = (it / kDParts) & 1u;
            mbar_wait_u32(ready0 + 8u * stage, full_phase);
            // Group A probes the partial-sum buffer it will fill at the end of the item now (test_wait, non-blocking): the buffer
            // is nearly always free, and a blocking try_wait costs ~90 cycles of latency even then (-0.4 % per sweep; the same
            // probe for the next item's window gained nothing).
            bool part_free = false;
            if (G == 0) part_free = it < (uint32_t)kDParts || mbar_test_u32(pempty0 + pb8, part_phase ^ 1u);

#pragma unroll
            for (int i = 0; i < kLPx; ++i) acc[i] = 0ull;
            if (G == 0) {
                if (reload) duo_a<true>(w, acc, sp, nw);
                else duo_a<false>(w, acc, sp, nw);
            } else if (kT) {
                // One batch of the next visit's weights per item: fetched before the dilation-24 part (which has registers to
                // spare while the loads are in flight) and installed in the other half of the columns after it.
                const bool do_batch = switch_tile && e < kDBatches;
                float4 g8[8];
                if (do_batch) tile_fetch(g8, nw, e);
                duo_b_far_tmem(acc, sp, tcur + kDNearCols);
                if (do_batch) tile_install(g8, e, tnext);
                duo_b_near_tmem(acc, sp, tcur);
            } else {
                if (kFar) {
                    if (reload) duo_b_far<true>(w, acc, sp, nw);
                    else duo_b_far<false>(w, acc, sp, nw);
                }
                if (reload) duo_b_near<true>(w, acc, sp, nw);
                else duo_b_near<false>(w, acc, sp, nw);
            }
            // A's warp j and B's warp j own the same 8 rows of the tile, so the hand-over is per warp pair
            if (G == 0) {
                if (!part_free) mbar_wait_u32(pempty0 + pb8, part_phase ^ 1u);
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) pp[i * 4 * kDPartPitch + j * 4] = acc[i * 4 + j];
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_u32(pfull0 + pb8);
                    // The ring stage is released AFTER the stores of the accumulators, never right after the last FMA: a store
                    // needs the FMA results, which need every window load to have RETURNED, and a release-arrive stays below
                    // earlier stores.  Placed before them, the arrive gets scheduled above FMAs whose LDS are still in flight
                    // and a fast refill overwrites the window rows that are read last (seen as rare wrong rows at 1024 x 1024).
                    mbar_arrive_u32(empty0 + 8u * stage);
                }
            } else {
                mbar_wait_u32(pfull0 + pb8, part_phase);
                {
                    ulonglong2 a[4];  // all four loads first: one shared-memory latency, not four
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const ulonglong2*>(pp + i * kDPartPitch);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[2 * i] = fadd2(acc[2 * i], a[i].x);
                        acc[2 * i + 1] = fadd2(acc[2 * i + 1], a[i].y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(pempty0 + pb8);
                // the output is not read again before the next sweep: evict-first stores keep L2 for windows and weights
                if (!out.planar) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r0 = unpack2(acc[2 * i]), r1 = unpack2(acc[2 * i + 1]);
                        if (i < nrows) __stcs(reinterpret_cast<float4*>(oc + (long long)i * out.pitch), make_float4(r0.x, r0.y, r1.x, r1.y));
                    }
                    oc += out.plane;
                } else {
                    const bool second = 2 * (q + e) + 1 < cx.C;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r0 = unpack2(acc[2 * i]), r1 = unpack2(acc[2 * i + 1]);
                        float* orow = oc + (long long)i * out.pitch;
                        if (i < nrows) {
                            __stcs(reinterpret_cast<float2*>(orow), make_float2(r0.x, r1.x));
                            if (second) __stcs(reinterpret_cast<float2*>(orow + out.plane), make_float2(r0.y, r1.y));
                        }
                    }
                    oc += 2 * out.plane;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(empty0 + 8u * stage);  // after the global stores (see group A above)
            }
            ++it;
            if (++stage == kDStages) {
                stage = 0;
                full_phase ^= 1u;
            }
        }

        if (kT && switch_tile) {  // finish the installation of the next visit's weights (visits shorter than six items), switch halves
#pragma unroll 1
            for (int b = n_q; b < kDBatches; ++b) {
                float4 g8[8];
                tile_fetch(g8, nw, b);
                tile_install(g8, b, tnext);
            }
            tmem_wait_st();
            const uint32_t t = tcur;
            tcur = tnext;
            tnext = t;
        }
        tile = next_tile;
    }
}

template <bool kFar>  // kFar: dilation 24 present ([1,2,4,8,12,24]); otherwise [1,2,4,8,12]
__global__ void __launch_bounds__(kLLaunchThreads, 1)
pamr_sweep_duo_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wts, DuoOut out, int C, int H, int W,
                      int tiles_x, int tiles_y, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    DuoCtx cx;
    cx.stage0 = reinterpret_cast<u64*>(smem_raw);
    cx.part = cx.stage0 + (size_t)kDStages * kDStageCells;
    cx.full = reinterpret_cast<uint64_t*>(cx.part + kDParts * kDPartCells);
    cx.ready = cx.full + kDStages;
    cx.empty = cx.ready + kDStages;
    cx.pfull = cx.empty + kDStages;
    cx.pempty = cx.pfull + 4 * kDParts;
    cx.wts = wts;
    cx.C = C;
    cx.Cp = (C + 1) >> 1;
    cx.H = H;
    cx.W = W;
    cx.tiles_x = tiles_x;
    cx.tiles_per_img = tiles_x * tiles_y;
    cx.n_tiles = n_tiles;
    cx.n_my = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    cx.total = cx.n_my * cx.Cp;
    cx.s0 = (int)(((long long)blockIdx.x * cx.Cp) / gridDim.x);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < kDStages; ++s) {
            mbar_init(&cx.full[s], 1);
            mbar_init(&cx.ready[s], 3);               // the three patch warps
            mbar_init(&cx.empty[s], kLThreads / 32);  // the eight compute warps
        }
        for (int s = 0; s < 4 * kDParts; ++s) {
            mbar_init(&cx.pfull[s], 1);   // warp j of group A
            mbar_init(&cx.pempty[s], 1);  // warp j of group B
        }
        fence_mbar_init();
    }
    constexpr bool kUseTmem = kFar && kDTmem;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cx.pempty + 4 * kDParts);
    int* meta = reinterpret_cast<int*>(tmem_slot + 2);
    cx.meta = meta;
    if (threadIdx.x == 0) {
        meta[0] = cx.n_my;
        meta[1] = cx.s0;
        meta[2] = cx.n_my == 0 ? 0 : cx.n_my + (cx.s0 > 0 ? 1 : 0);
    }
    if (kUseTmem && (threadIdx.x >> 5) == kLGroupThreads / 32) {  // first warp of group B allocates (and frees) the columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kDTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (kUseTmem) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kUseTmem) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    cx.tmem_base = kUseTmem ? *tmem_slot : 0u;

    if (threadIdx.x >= kLThreads) {  // producer warpgroup: hand its registers to the compute warps
        // the CTA owns 384 x 168 registers; what this warpgroup gives back ((168 - 24) x 128) is exactly what the two
        // compute warpgroups take ((240 - 168) x 256)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (threadIdx.x == kLThreads) duo_producer(cx, &tmap);
        else if (threadIdx.x >= kLThreads + 32) duo_patcher(cx);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    if (threadIdx.x < kLGroupThreads) {  // warp-uniform
        duo_group<0, kFar>(cx, out);
    } else {
        duo_group<1, kFar>(cx, out);
        if (kUseTmem) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");  // all of group B is done with its columns
            if ((threadIdx.x >> 5) == kLGroupThreads / 32)
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem_base), "n"(kDTmemCols) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------ host
int duo_pitch(int W) { return 2 * (W + 2 * kPamrPad); }  // floats per row of a pair plane (never a power of two, see pamr.cu)
size_t duo_plane_elems(int H, int W) { return (size_t)H * duo_pitch(W); }
size_t duo_buffer_elems(int B, int C, int H, int W) { return (size_t)B * ((C + 1) / 2) * duo_plane_elems(H, W); }

// cells_in: pair cells (duo_buffer_elems).  out_planar == 0: `out` is another pair-cell buffer; otherwise the caller's
// planar [B,C,H,W] tensor.
int launch_sweep_duo(const float* w, const float* cells_in, float* out, int out_planar, int B, int C, int H, int W, int D,
                     cudaStream_t s) {
    const int Cp = (C + 1) / 2;
    CUtensorMap tmap;  // over the H x 2W floats of each pair plane; boxes may start at negative coordinates (zero fill)
    const int rc = encode_tmap_3d_f32_strided(&tmap, cells_in, 2 * W, H, (long long)B * Cp, duo_pitch(W),
                                              (long long)duo_plane_elems(H, W), 2 * kDPitch, kBox);
    if (rc != 0) {
        set_error("pamr_sweep_duo: cuTensorMapEncodeTiled failed (%d)", rc);
        return CL4_ECUDA;
    }
    DuoOut so;
    so.ptr = out;
    so.planar = out_planar ? 1 : 0;
    so.plane = out_planar ? (long long)H * W : (long long)duo_plane_elems(H, W);
    so.pitch = out_planar ? W : duo_pitch(W);
    const int tiles_x = ceil_div(W, kTile), tiles_y = ceil_div(H, kTile);
    const int n_tiles = B * tiles_x * tiles_y;
    auto kern = D == 6 ? pamr_sweep_duo_kernel<true> : pamr_sweep_duo_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmem);
    if (e != cudaSuccess) {
        set_error("pamr_sweep_duo: smem attribute: %s", cudaGetErrorString(e));
        return CL4_ECUDA;
    }
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<grid, kLLaunchThreads, kDSmem, s>>>(tmap, w, so, C, H, W, tiles_x, tiles_y, n_tiles);
    return check_launch("pamr_sweep_duo");
}

}  // namespace cl4
