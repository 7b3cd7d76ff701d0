// Argument block + host launchers of mkutrans / velpred (see mgpu_velpred.cu).
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {

struct VpArgs {
  int dm, ppm_type, slope_order;
  bool trace;
  int lo[3], hi[3];
  int bclo[3][3], bchi[3][3];  // [component][direction]: adv_bc(d, side, comp)
  int plo[3], phi[3];          // phys_bc(d, side)
  double dt, dx[3], rel_eps;
  double rdx[3];  // 1/dx (FAST build)
  Box3 tb, vb;  // lo-1:hi+1 and lo:hi
  DV utilde, ufull, force;
  DV utrans[3], umac[3];
  DV utpre[3];  // k_vp_face_ut only: utrans before w0 is added (utrans[] then holds utrans + w0)
  DV UL[3], UR[3], UIMH[3];  // per face direction, dm components, on tb
  DV Q[3][3];                // [component][face direction], 3-D only
  const double* w0;
  bool ufull_otf;  // k_vp_face_ut: ufull is not in memory, the kernel forms utilde + the cell average of w0 itself
  int nr;          // cells of the domain in the radial direction (periodic image of a ghost cell)
  bool spherical;  // every direction's Riemann problem carries w0mac[d] (mkutrans.f90:601, velpred.f90:1588)
  DV w0mac[3];
};

void mkutrans_dev(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* utrans, const double* w0_dev,
                  const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u,
                  const DV* w0mac = nullptr);
void velpred_dev(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* umac, const DV* utrans, const DV& force,
                 const double* w0_dev, const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u,
                 int ng_f, const DV* w0mac = nullptr);
// advance_premac on a box without physical boundaries: face states and utrans from one reconstruction
bool velpred_premac_fusable(const mgpu_params& P, const int* phys_bc);
void velpred_premac_begin(const mgpu_params& P, const DV& utilde, const DV& ufull, const DV* utpre, const DV* utfull,
                          const double* w0_dev, const int* lo, const int* hi, const int* adv_bc, const int* phys_bc,
                          int ng_u, VpArgs* st);
void velpred_premac_finish(const mgpu_params& P, VpArgs* st, DV* umac, const DV& force, int ng_u, int ng_f);
size_t velpred_scratch(const mgpu_params& P, const int* lo, const int* hi);
// 1 (default): the FAST build of the kernels (mgpu_velpred_fast.cu: FMA, reciprocals; <= 1e-12); 0: the bit-identical one
void velpred_set_fast(int on);

}  // namespace mgpu
