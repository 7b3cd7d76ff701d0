// Argument block of the general make_edge_scal kernels (see mgpu_edge.cu).
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {

struct EdgeArgs {
  int dm, ppm_type, slope_order;
  bool trace_forces, is_cons;
  int lo[3], hi[3];
  int bclo[3], bchi[3];
  bool velnorm[3];  // is_vel .and. comp == d: the normal velocity component gets the inflow clamp
  double dt, dx[3], rel_eps;
  Box3 tb, vb;      // lo-1:hi+1 and lo:hi
  DV s, force;      // single-component views
  DV umac[3];
  DV sedge[3];      // single-component views of the output
  DV Ip, Im, Ipf, Imf;  // dm components each, on tb
  DV simh[3];           // on tb (faces lo..hi+1 used)
  DV simht[3][3];       // [d][t], 3-D only
};

// one component (0-based comp, 1-based bccomp) of one box, device pointers
void make_edge_scal_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                        const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel,
                        bool is_cons, int ng_s, int ng_f);
// arena bytes one make_edge_scal_dev call needs
size_t make_edge_scal_scratch(const mgpu_params& P, const int* lo, const int* hi);

}  // namespace mgpu
