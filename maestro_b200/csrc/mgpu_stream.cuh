// Argument blocks + host launchers of the streaming kernels (see mgpu_stream.cu).
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {

struct FluxArgs {
  int dm, species_pred_type;
  bool evolve_base_state;
  int rho, rhoh, spec0, nspec;  // 0-based component indices
  Box3 vb;
  DV sflux[3], sedge[3], umac[3], eta;
  // base state (device pointers)
  const double *w0, *rho0_old, *rho0_edge_old, *rho0_new, *rho0_edge_new, *rho0_predicted_edge;
  const double *rhoh0_old, *rhoh0_edge_old, *rhoh0_new, *rhoh0_edge_new;
};
void mk_rhoX_flux_dev(const mgpu_params& P, FluxArgs& a, int startcomp, int endcomp);
void mk_rhoh_flux_dev(const mgpu_params& P, FluxArgs& a);

struct UpdArgs {
  int dm;
  double dt, dx[3];
  Box3 vb;
  DV sold, snew, force, sflux[3];
  // the EOS reset below the cutoff (update_scal.f90:421-447): p0_new on the radial cells (planar, device pointer) or on
  // the cell centres (spherical); nullptr / false when the caller did not pass them
  const double* p0_new = nullptr;
  DV p0_new_cart;
  bool have_p0_new_cart = false;
};
void update_scal_dev(const mgpu_params& P, UpdArgs& a, int nstart, int nstop);
// mk_rhoX_flux (species + tracers) + update_scal (species + tracers + density) of density_advance in one launch
void flux_update_all_dev(const mgpu_params& P, FluxArgs& a, UpdArgs& u, bool exact, bool force_zero = false,
                         bool skip_rho_copy = false);
// the same for the z planes [k0, k1] of the box only (lean 3-D kernel; no density pre-copy: the caller refills every
// ghost cell of snew afterwards)
bool flux_update_range_supported(const mgpu_params& P, const FluxArgs& a, const UpdArgs& u);
void flux_update_range_dev(const mgpu_params& P, FluxArgs& a, UpdArgs& u, bool force_zero, int k0, int k1);

struct VelArgs {
  int dm;
  bool do_sponge;
  double dt, dx[3];
  double rdx[3];  // 1/dx, set by update_velocity_dev (FAST kernel)
  Box3 vb;
  DV uold, unew, force, sponge, umac[3], uedge[3];
  const double* w0;
};
// fast: divisions by dx as multiplications by reciprocals (the velocity_advance episodes outside the exact build)
void update_velocity_dev(VelArgs& a, bool fast = false);

// force builders
struct RhohForceArgs {
  int dm, nr, cutoff_coord;
  bool with_psi, add_thermal;
  double dr;
  Box3 vb;
  DV f, thermal, wm;  // f: the rhoh component of scal_force; wm: umac in the radial direction
  const double *p0_1, *p0_2, *rho0_1, *rho0_2, *grav, *psi;
};
void mkrhohforce_dev(RhohForceArgs& a);
struct VelForceArgs {
  int dm, nr;
  bool is_final_update, add_utilde;
  double dr, rho_cut, omega, sin_theta, cos_theta, rotation_radius;
  Box3 vb;
  DV force, uold, gpi, rho, uedge[3];
  const double *w0, *rho0, *grav, *w0_force;
};
void mk_vel_force_dev(VelForceArgs& a);
// ufull = put_1d_array_on_cart(w0 -> radial component) (+ ghost fill by the caller) ; dst += src
void radial_cell_avg_dev(const mgpu_params& P, const DV& ufull, const double* w0_dev, const int* lo, const int* hi);
void add_dev(double* dst, const double* src, long n);
// a(comp ca) op= b(comp cb) over the valid box grown by g: op 0 '/', 1 '*'
void comp_muldiv_dev(const mgpu_params& P, const DV& a, int ca, const DV& b, int cb, int op, int g, const int* lo,
                     const int* hi);

void addw0_dev(const mgpu_params& P, DV* umac, const double* w0_dev, double mult, const int* lo, const int* hi);
void modify_scal_force_dev(const mgpu_params& P, const DV& force, const DV& s, const DV* umac, const double* s0,
                           const double* s0_edge, const double* w0, int comp, bool fullform, const int* lo,
                           const int* hi, bool fast, bool assign = false);
void convert_rhoX_to_X_dev(const mgpu_params& P, const DV& s, bool flag, const int* lo, const int* hi);
void put_in_pert_form_dev(const mgpu_params& P, const DV& s, const double* base_dev, int comp, bool flag,
                          const int* lo, const int* hi);
// Batched ghost fills: between begin and end fill_boundary_dev only records; end issues every slab exchange of the
// batch as ONE NCCL group and then the local wraps / physical BCs in the recorded order (mgpu_stream.cu).
void species_form_dev(const mgpu_params& P, const DV& s, const double* base_dev, bool convert, bool pert, bool forward,
                      const int* lo, const int* hi);
void recip_dev(double* dst, const double* src, long n);
void fill_batch_begin();
void fill_batch_end();
void fill_batch_exchange_async();  // exchanges on the communication stream, behind the compute stream's work so far
void fill_batch_finish();          // compute stream waits for them; in-box wraps and physical BCs
void fill_batch_abort();
struct FillBatch {  // RAII: a throw inside the batch drops the recorded requests
  bool done = false;
  FillBatch() { fill_batch_begin(); }
  void run() { done = true; fill_batch_end(); }
  void start_exchange() { fill_batch_exchange_async(); }
  void finish() { done = true; fill_batch_finish(); }
  ~FillBatch() { if (!done) fill_batch_abort(); }
};
// several boxes on this rank: box-to-box ghost copies (periodic images included) + per-box physical BCs
void fill_boundary_mf_dev(const mgpu_params& P, int nfabs, const DV* fabs, const int* const* lo, const int* const* hi, int ng,
                          const int* nodal, int scomp, int bccomp, int ncomp, const int* const* adv_bc, const int* pmask,
                          bool same_boundary);
void fill_boundary_dev(const mgpu_params& P, const DV& s, const int* lo, const int* hi, int ng, const int* nodal,
                       int scomp, int bccomp, int ncomp, const int* adv_bc, const int* pmask, bool same_boundary);
void sum_comps_dev(const DV& a, int dst, int c0, int ncomp);
void set_dev(double* p, double v, long n);
void copy_dev(double* dst, const double* src, long n);

}  // namespace mgpu
