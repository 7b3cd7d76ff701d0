// EOS pieces next to the advective path: see mgpu_eos.cu.
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {

// what eos_init leaves in the reference's module variables (eos.F90:26, gamma_law_general.f90:46): passed to the kernels
// by value, so every field is a constant-bank operand
struct EosDev {
  int kind, neutral, nspec;
  double gamma, k_B, m_nucleon;
  double mintemp, maxtemp, mindens, maxdens, mine, maxe, minp, maxp, minh, maxh;
  double small_temp;
  double aion[MGPU_EOS_MAXSPEC], zion[MGPU_EOS_MAXSPEC];
};
// the fields of eos_t (eos_type.f90:102) the advective path reads
struct EosState {
  double rho, T, p, e, h, cv, cp, cs, dpdT, dpdr, dedT, dedr, dhdT, mu, mu_e, abar, zbar;
};

void set_eos(const mgpu_eos* e);
bool have_eos();
const EosDev& the_eos(const char* who);  // throws "EOS: not initialized" (eos.F90:117) when unset

// eos(input, state) at n points (mgpu_eos_eval); device arrays
void eos_points_dev(int input, long n, double* state, const double* xn);

// update_scal.f90:421-447 / :559-585: rhoh of the zones with rho <= base_cutoff_density from (rho, p0_new, X)
void update_scal_eos_dev(const mgpu_params& P, const DV& sold, const DV& snew, const double* p0_new, const DV* p0_new_cart,
                         const Box3& vb);

struct HEdgeArgs {
  int dm, ept, spt, rho, rhoh, temp, spec0;
  bool sphr;
  Box3 vb;
  DV sedge[3];
  // planar: base state on cells (0:nr-1) and edges (0:nr)
  const double *rho0_old, *rhoh0_old, *t0_old, *rho0_edge_old, *rhoh0_edge_old, *t0_edge_old;
  const double *rho0_new, *rhoh0_new, *t0_new, *rho0_edge_new, *rhoh0_edge_new, *t0_edge_new;
  // spherical: half-time base state on the cell centres, 2 ghost cells
  DV rho0_cart, rhoh0_cart, t0_cart;
};
void h_from_rhot_edge_dev(const HEdgeArgs& a);

struct TempForceArgs {
  int dm, nr, rho, temp, spec0;
  bool sphr;
  double dr, dx[3];
  Box3 vb;
  DV f, s, thermal, umac[3];  // f: the temp component of temp_force
  const double *p0_old, *psi;  // planar
  DV p0_cart, psi_cart;        // spherical (1 ghost cell / valid cells)
};
void mktempforce_dev(const TempForceArgs& a);

// firstdt_2d / _3d / _3d_sphr of one box (firstdt.f90:330, :460, :599): *dt and *umax out
void firstdt_box_dev(const mgpu_params& P, const DV& u, const DV& s, const DV& force, const DV& divU, const double* p0,
                     const double* gamma1bar, const DV* gp0_cart, const int* lo, const int* hi, double cfl,
                     bool use_soundspeed, bool use_divu, double* dt, double* umax);

// makeTfromRhoH / makeTfromRhoP on the valid cells (rhoh_vs_t.f90:929-1158, :1287-1445)
void make_t_dev(const mgpu_params& P, const DV& state, const double* p0, const DV* p0_cart, bool from_p, bool flag,
                bool update_rhoh, const Box3& vb);

}  // namespace mgpu
