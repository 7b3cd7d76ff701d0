// Host launchers of the BDS edge-state path (see mgpu_bds.cu).
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {

// one component (0-based comp) of one box, device pointers; temporaries from the arena
void bds_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
             const int* lo, const int* hi, int comp, bool is_cons, int ng_s, int ng_f);
size_t bds_scratch(const mgpu_params& P, const int* lo, const int* hi);

// 1 (default): the FAST build of the kernels (FMA, reciprocals; <= 1e-12); 0: the bit-identical build
void bds_set_fast(int on);

}  // namespace mgpu
