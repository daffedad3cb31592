// Internal launch interface between the PAMR translation units.
#pragma once
#include "common.cuh"

namespace cl4 {

constexpr int kPamrPad = 24;  // frame of the replicate-padded mask planes = largest dilation of the TMA path

// TMA-staged sweep over replicate-padded planes (pamr_tma.cu).  Applicable when W % 4 == 0, the
// map has at least 32x32 pixels, D <= 6 and every dilation is <= 24; otherwise the register/L1
// kernel in pamr.cu is used.
bool sweep_tma_applicable(int H, int W, const Dilations& dil, int D);
size_t padded_plane_elems(int H, int W);
size_t tiled_weight_elems(int B, int H, int W, int D);  // weights in [tile][8D][32][32] layout (>= B*8D*H*W)
int launch_pad_copy(const float* src, float* dst, long long planes, int H, int W, cudaStream_t s);
int launch_pad_frame(float* buf, long long planes, int H, int W, cudaStream_t s);  // replicate frame, in place
// Affinity weights from a replicate-padded image [B*K][H+48][W+48] into the tile-major layout; K <= 3.
bool weights_tma_applicable(int K);
int launch_weights_tma(const float* padded_img, float* w, int B, int K, int H, int W, const Dilations& dil, int D,
                       cudaStream_t s);
// w: tile-major weights (tiled_weight_elems), as written by the weights kernel in tiled mode
int launch_sweep_tma(const float* w, const float* padded_in, float* out, int out_padded, int B, int C, int H, int W,
                     const Dilations& dil, int D, cudaStream_t s);

// Lattice sweep for the dilation sets [1,2,4,8,12,24] and [1,2,4,8,12] (pamr_lattice.cu): two compute warp groups owning
// dilations {4,8,12} and {1,2,24} of a 32 x 32 tile + a producer warpgroup; weights in its own thread-major layout
// (lattice_weight_elems).
bool sweep_lattice_applicable(int K, int H, int W, const Dilations& dil, int D);
size_t lattice_weight_elems(int B, int H, int W);
int launch_weights_lattice(const float* padded_img, float* w, int B, int K, int H, int W, int D, cudaStream_t s);
// in / out: planes of H x W pixels, rows `pitch` and planes `plane` elements apart (the replicate padding happens
// inside the kernel's shared-memory ring; nothing around the image area is read or written)
// out_cells != 0: `out` is a pair-cell buffer (duo_buffer_elems; pitch = duo_pitch, plane = pair plane), the layout the
// class-pair sweep reads
int launch_sweep_lattice(const float* w, const float* in, int in_pitch, long long in_plane, float* out, int out_pitch,
                         long long out_plane, int out_cells, int B, int C, int H, int W, int D, cudaStream_t s);

// Class-pair ("duo") sweep for the same two dilation sets (pamr_duo.cu): the masks between sweeps are pair-interleaved cells
// [B][ceil(C/2)][H][W + 48][2]; packed fp32 FMAs on both classes of a cell.  Weights: the lattice layout.
int duo_pitch(int W);                                          // floats per row of a pair plane
size_t duo_plane_elems(int H, int W);                          // floats per pair plane
size_t duo_buffer_elems(int B, int C, int H, int W);           // floats per pair-cell buffer
int launch_sweep_duo(const float* w, const float* cells_in, float* out, int out_planar, int B, int C, int H, int W, int D,
                     cudaStream_t s);

// All iterations on-chip for maps up to 64 x 64 (pamr_fused.cu): D <= 6, every dilation <= 24.
// w: tile-major weights; mask_in / mask_out: plain [B*C][H][W]; num_iter >= 1.
bool pamr_fused_applicable(int H, int W, const Dilations& dil, int D);
int launch_pamr_fused(const float* w, const float* mask_in, float* mask_out, int B, int C, int H, int W, int num_iter,
                      const Dilations& dil, int D, cudaStream_t s);

}  // namespace cl4
