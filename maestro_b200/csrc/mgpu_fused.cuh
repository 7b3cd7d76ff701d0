// Fused single-launch kernels for the headline configurations (see mgpu_fused.cu).
#pragma once
#include "mgpu_common.cuh"

#ifndef MGPU_FUSED_BX
#define MGPU_FUSED_BX 32
#endif
#ifndef MGPU_FUSED_BY
#define MGPU_FUSED_BY 8
#endif

#ifndef MGPU_FUSED2_BY
#define MGPU_FUSED2_BY 1616 /* 16x16 tile: 77 % of the threads store results (32x8: 70 %) */
#endif

namespace mgpu {

struct FusedArgs {
  int slope_order;
  int lo[3], hi[3];
  int bclo[3], bchi[3];
  bool velnorm[3];
  int kchunk;  // z planes per CTA
  bool force_zero;  // the force of this component is identically zero: do not read it (density_advance.f90:99-103)
  // on-the-fly input transforms of the upwind-first kernel (device-resident lean episode, density_advance.f90:160-171
  // and :148): the predicted quantity is s*sdiv (X = rhoX * (1/rho), sdiv = the reciprocal density, same layout as s) and/or s - ssub(k) (rho' = rho - rho0(k)), the z
  // velocity is wmac + wadd(k); ssub / wadd are indexed by the plane index inside the s / wmac fab.  nullptr: none.
  const double* sdiv;
  const double* ssub;
  const double* wadd;
  double dt, dx[3], rel_eps;
  // dt/dx, dt/(6 dx), dt/(4 dx) and dt/2, filled by fused_edge3_launch: read as constant-bank operands, no registers
  double td[3], c6[3], c4[3], dt2;
  // third design, boxes with physical boundaries: 0 every tile; 1 only the tiles no boundary rule can reach (run by the
  // plain kernel), 2 only the others (run by the boundary kernel) -- see fused_edge3_launch
  int tile_mode;
  // third design: > 0: the first and the last z chunk are kedge planes thick and the chunks between them kchunk (a box
  // with a physical boundary in z: the boundary kernel then runs thin chunks, the plain kernel everything else)
  int kedge;
  DV s, force;  // single-component views
  DV umac[3];
  DV sedge[3];  // single-component views of the output
};

// true when the fused 3-D kernel covers this case (otherwise the general path of mgpu_edge.cu runs)
bool fused_edge_supported(const mgpu_params& P, bool is_cons);
// one component (0-based comp, 1-based bccomp) of one box, device pointers
// exact: bit-identical arithmetic (-fmad=false build); otherwise the FAST build (dt/dx folded, FMA)
void fused_edge_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                    const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel, int ng_s,
                    int ng_f, int kchunk, bool exact, bool force_zero = false, const double* sdiv = nullptr,
                    const double* ssub = nullptr, const double* wadd = nullptr);
// true when fused_edge_dev would run the upwind-first kernel for this component (FAST build, all faces INTERIOR)
bool fused_edge_is_upwind_first(const mgpu_params& P, const int* adv_bc, int bccomp, bool exact);
void fused_edge_launch_exact(const FusedArgs& a, int ppm_type, bool bc, int nx, int ny, int nz);
void fused_edge_launch_fast(const FusedArgs& a, int ppm_type, bool bc, int nx, int ny, int nz);
// second design (mgpu_fused2.cu): upwind-first, FAST arithmetic only; bc: the box has physical boundaries
void fused_edge2_launch(const FusedArgs& a, int ppm_type, int nx, int ny, int nz, bool bc);
// third design (mgpu_fused3.cu): the upwind-first algorithm with TMA-staged tiles and a register-renamed plane loop
bool fused_edge3_supported(const FusedArgs& a, bool bc);
void fused_edge3_launch(const FusedArgs& a, int ppm_type, int nx, int ny, int nz, bool bc);
void fused_edge3_set_split(int on);  // boundary boxes: interior tiles through the plain kernel (default on)
void fused_edge3_set_pair_streams(int on);  // tile split: the boundary kernel on a second stream (default on)
void fused_edge3_set_thin_edge(int on);  // boundary in z: 8-plane end chunks for the boundary kernel (default on)
// 2-D (mgpu_fused2.cu, k_fused_edge2d): FAST arithmetic only, non-conservative, ppm_trace_forces = 0, no REFLECT_ODD
bool fused_edge2d_supported(const mgpu_params& P, bool is_cons, const int* adv_bc, int bccomp, bool exact);
void fused_edge2d_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                      const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel, int ng_s,
                      int ng_f, bool force_zero = false);
void fused_edge2d_launch(const FusedArgs& a, int ppm_type, int nx, int ny, bool bc);
// 0: always the literal kernel; 1 (default): the upwind-first kernels wherever they apply (the TMA design where its
// stride rules hold, else the second design); 4: the second design only
void fused_edge_set_variant(int v);
void fused_edge2d_set_tile(int t);  // 2-D kernel tile: 0 32x8, 1 16x16, 2 32x16
void fused_edge2_set_by(int by);  // tile of the upwind-first kernel: 1616 (16x16, default), 8 (32x8), 16 (32x16, plain inputs only)

}  // namespace mgpu
