// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).  Declarations of the restated kernels.
#pragma once
#include "mo_array.h"

namespace mo {

struct Box {
  int lo[3];
  int hi[3];
};

// valid box grown by g in every active dimension (dims >= dm stay [0,0])
inline Box grown(const int* lo, const int* hi, int dm, int g) {
  Box b;
  for (int d = 0; d < 3; ++d) {
    if (d < dm) {
      b.lo[d] = lo[d] - g;
      b.hi[d] = hi[d] + g;
    } else {
      b.lo[d] = b.hi[d] = 0;
    }
  }
  return b;
}

// loop nest in Fortran order; OpenMP over the outermost non-degenerate index, which is where the
// reference puts its !$OMP PARALLEL DO (k in 3-D).
template <class F>
inline void for_box(const Box& b, F f) {
  if (b.hi[2] > b.lo[2]) {
#pragma omp parallel for
    for (int k = b.lo[2]; k <= b.hi[2]; ++k)
      for (int j = b.lo[1]; j <= b.hi[1]; ++j)
        for (int i = b.lo[0]; i <= b.hi[0]; ++i) f(i, j, k);
  } else {
    for (int k = b.lo[2]; k <= b.hi[2]; ++k)
      for (int j = b.lo[1]; j <= b.hi[1]; ++j)
        for (int i = b.lo[0]; i <= b.hi[0]; ++i) f(i, j, k);
  }
}

struct Ctx {  // module variables of the reference, per call
  mgpu_params p;
};

void slope_dir(const Arr& s, Arr& slp, const int* lo, const int* hi, int dm, int d, int bclo, int bchi,
               int slope_order);

void ppm(const Arr& s, const Arr* vel, Arr& Ip, Arr& Im, const int* lo, const int* hi, int dm,
         const int bc[3][2], const double* dx, double dt, bool is_umac, int ppm_type, double rel_eps,
         int ng_s);

// make_edge_scal_2d / _3d for one component (0-based comp index into s/sedge/force; 1-based bccomp)
void make_edge_scal_box(const mgpu_params& P, const Arr& s, Arr* sedge, const Arr* umac, const Arr& force,
                        const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel,
                        bool is_conservative, int ng_s);

void mk_rhoX_flux_box(const mgpu_params& P, Arr* sflux, Arr& etarhoflux, const Arr* sedge, const Arr* umac,
                      const double* w0, const double* rho0_old, const double* rho0_edge_old,
                      const double* rho0_new, const double* rho0_edge_new, const double* rho0_predicted_edge,
                      int startcomp, int endcomp, const int* lo, const int* hi);

void mk_rhoh_flux_box(const mgpu_params& P, Arr* sflux, const Arr* sedge, const Arr* umac, const double* w0,
                      const double* rho0_old, const double* rho0_edge_old, const double* rho0_new,
                      const double* rho0_edge_new, const double* rhoh0_old, const double* rhoh0_edge_old,
                      const double* rhoh0_new, const double* rhoh0_edge_new, const int* lo, const int* hi);

void update_scal_box(const mgpu_params& P, int nstart, int nstop, const Arr& sold, Arr& snew, const Arr* sflux,
                     const Arr& force, const int* lo, const int* hi, const double* p0_new = nullptr,
                     const Arr* p0_new_cart = nullptr);

void update_velocity_box(const mgpu_params& P, const Arr& uold, Arr& unew, const Arr* umac, const Arr* uedge,
                         const Arr& force, const Arr& sponge, const double* w0, const int* lo, const int* hi);

void addw0_box(const mgpu_params& P, Arr* umac, const double* w0, double mult, const int* lo, const int* hi);

void modify_scal_force_box(const mgpu_params& P, Arr& force, const Arr& s, const Arr* umac, const double* s0,
                           const double* s0_edge, const double* w0, int comp, bool fullform, const int* lo,
                           const int* hi);

void cell_to_edge(const double* s0_cell, double* s0_edge, int nr);

void mkutrans_box(const mgpu_params& P, const Arr& utilde, const Arr& ufull, Arr* utrans, const double* w0,
                  const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u,
                  const Arr* w0mac = nullptr);

void velpred_box(const mgpu_params& P, const Arr& utilde, const Arr& ufull, Arr* umac, const Arr* utrans,
                 const Arr& force, const double* w0, const int* lo, const int* hi, const int* adv_bc,
                 const int* phys_bc, int ng_u, const Arr* w0mac = nullptr);

void bds_box(const mgpu_params& P, const Arr& s, Arr* sedge, const Arr* umac, const Arr& force, const int* lo,
             const int* hi, int comp, bool is_conservative);

void mkrhohforce_box(const mgpu_params& P, Arr& scal_force, bool is_prediction, const Arr& thermal, const Arr* umac,
                     const double* p0_1, const double* p0_2, const double* rho0_1, const double* rho0_2,
                     const double* grav, const double* psi, bool add_thermal, const int* lo, const int* hi);
void mk_vel_force_box(const mgpu_params& P, Arr& vel_force, bool is_final_update, const Arr& uold, const Arr* uedge,
                      const double* w0, const Arr& gpi, const Arr& rho, const double* rho0, const double* grav,
                      const double* w0_force, const int* lo, const int* hi, bool do_add_utilde_force);
void advance_premac_box(const mgpu_params& P, const Arr& uold, const Arr& sold, Arr* umac, const Arr& gpi,
                        const double* w0, const double* w0_force, const double* rho0_old, const double* grav_cell_old,
                        const int* lo, const int* hi, int ng_u, const int* adv_bc, const int* phys_bc, const int* pmask);
void velocity_advance_box(const mgpu_params& P, const Arr& uold, Arr& unew, const Arr& sold, const Arr& rhohalf,
                          Arr* umac, const Arr& gpi, const double* w0, const double* w0_force, const double* rho0_old,
                          const double* rho0_nph, const double* grav_cell_old, const double* grav_cell_nph,
                          const Arr& sponge, const int* lo, const int* hi, int ng_u, const int* adv_bc,
                          const int* pmask);
void enthalpy_advance_box(const mgpu_params& P, int which_step, Arr& sold, Arr& snew, Arr* sedge, Arr* sflux,
                          Arr& scal_force, const Arr& thermal, Arr* umac, const double* w0, const double* rho0_old,
                          const double* rhoh0_old, const double* rho0_new, const double* rhoh0_new,
                          const double* p0_old, const double* p0_new, const double* tempbar, const double* psi,
                          const double* grav_old, const double* grav_nph, const int* lo, const int* hi, int ng_s, int ng_f,
                          const int* adv_bc, const int* pmask);

// ghost fill of a single box covering the whole domain: periodic wrap + multifab_physbc
void fill_boundary_box(const mgpu_params& P, Arr& s, const int* lo, const int* hi, int ng, int scomp, int bccomp,
                       int ncomp, const int* adv_bc, const int* pmask);
// face-centred periodic fill (multifab_fill_boundary on a nodal multifab), 1 ghost layer
void fill_boundary_face(const mgpu_params& P, Arr& u, const int* lo, const int* hi, int ng, int dir,
                        const int* pmask);

// reductions next to the path (mo_reduce.cpp): estdt.f90:142-220 for one level, make_eta.f90:36
void estdt_level(const mgpu_params& P, int nfabs, const mgpu_fab* u, const mgpu_fab* s, const mgpu_fab* force,
                 const mgpu_fab* divU, const mgpu_fab* dSdt, const double* w0, const double* p0,
                 const double* gamma1bar, double rho_min, double cfl, double& dt, double& umax);
// put_1d_array_on_cart_3d_sphr (mo_sphr.cpp, fill_3d_data.f90:269)
void put_1d_array_on_cart_sphr(const mgpu_params& P, const mgpu_geom& g, bool edge_in, bool vec, const double* s0,
                               Arr& cart, const int* lo, const int* hi);
void estdt_sphr_level(const mgpu_params& P, const mgpu_geom& g, int nfabs, const mgpu_fab* u, const mgpu_fab* s,
                      const mgpu_fab* force, const mgpu_fab* divU, const mgpu_fab* dSdt, const mgpu_fab* const* w0mac,
                      const double* w0, const double* p0, const double* gamma1bar, double rho_min, double cfl,
                      double& dt, double& umax);
void make_etarho_planar(const mgpu_params& P, int nfabs, const mgpu_fab* etarhoflux, double* etarho_ec,
                        double* etarho_cc);

// average (average.f90:24) of one level: phibar(0:nr-1); g == nullptr for planar geometry
void average_level(const mgpu_params& P, const mgpu_geom* g, int nfabs, const mgpu_fab* phi, int incomp, int nr_irreg,
                   int drdxfac, double* phibar);

// ---- EOS pieces (mo_eos.cpp) ----
struct EosState {  // the fields of eos_t (eos_type.f90:102) the advective path reads
  double rho, T, p, e, h, cv, cp, cs, dpdT, dpdr, dedT, dedr, dhdT, mu, mu_e, abar, zbar;
};
void set_eos(const mgpu_eos* e);
bool have_eos();
const mgpu_eos& the_eos(const char* who);
void eos_call(int input, EosState& s, const double* xn);
void update_scal_eos_box(const mgpu_params& P, const Arr& sold, Arr& snew, const double* p0_new, const Arr* p0_new_cart,
                         const int* lo, const int* hi);
void h_from_rhot_edge_box(const mgpu_params& P, Arr* sedge, const double* rho0_old, const double* rhoh0_old,
                          const double* t0_old, const double* rho0_edge_old, const double* rhoh0_edge_old,
                          const double* t0_edge_old, const double* rho0_new, const double* rhoh0_new,
                          const double* t0_new, const double* rho0_edge_new, const double* rhoh0_edge_new,
                          const double* t0_edge_new, const int* lo, const int* hi);
void h_from_rhot_edge_sphr_box(const mgpu_params& P, Arr* sedge, const Arr& rho0_cart, const Arr& rhoh0_cart,
                               const Arr& t0_cart, const int* lo, const int* hi);
void mktempforce_box(const mgpu_params& P, Arr& temp_force, const Arr& s, const Arr* umac, const Arr& thermal,
                     const double* p0_old, const double* psi, const int* lo, const int* hi);
void mktempforce_sphr_box(const mgpu_params& P, Arr& temp_force, const Arr& s, const Arr* umac, const Arr& thermal,
                          const Arr& p0_cart, const Arr& psi_cart, const int* lo, const int* hi);
void firstdt_box(const mgpu_params& P, const Arr& u, const Arr& s, const Arr& force, const Arr& divU, const double* p0,
                 const double* gamma1bar, const Arr* gp0_cart, const int* lo, const int* hi, double cfl,
                 bool use_soundspeed_firstdt, bool use_divu_firstdt, double& dt, double& umax);
void make_t_box(const mgpu_params& P, Arr& state, const double* p0, const Arr* p0_cart, bool from_p, bool flag,
                bool update_rhoh, const int* lo, const int* hi);
}  // namespace mo

// spherical helpers of mo_sphr.cpp used by the EOS entry points (they live outside namespace mo there)
mo::Arr cart_with_ghosts(const mgpu_params& P, const mgpu_geom& g, const double* s0, int ng, int bccomp, const int* lo,
                     const int* hi, const int* adv_bc, const int* pmask);
void mk_vel_force_sphr_box(const mgpu_params& P, const mgpu_geom& g, mo::Arr& vel_force, bool is_final_update, const mo::Arr& uold,
                           const mo::Arr* uedge, const double* w0, const mo::Arr* w0mac, const mo::Arr& gpi, const mo::Arr& rho,
                           const mo::Arr& normal, const double* rho0, const double* grav, const mo::Arr& w0_force_cart,
                           const int* lo, const int* hi, bool do_add_utilde_force);
