// Internal launch interface between the PAMR translation units.
#pragma once
#include "common.cuh"

namespace cl4 {

// TMA-staged sweep (pamr_tma.cu).  Applicable when W % 4 == 0, the map has at least 32x32
// pixels and every dilation is <= 24; otherwise the register/L1 kernel in pamr.cu is used.
bool sweep_tma_applicable(int C, int H, int W, const Dilations& dil, int D, const float* mask_in);
int launch_sweep_tma(const float* w, const float* mi, float* mo, int B, int C, int H, int W, const Dilations& dil,
                     int D, cudaStream_t s);

}  // namespace cl4
