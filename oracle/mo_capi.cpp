// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// extern "C" surface of the oracle, mirroring include/maestro_b200.h one to one (prefix mo_ instead
// of mgpu_) so that the parity tests call both with identical arguments.  Host pointers only.
#include <stdexcept>
#include <string>

#include <algorithm>
#include <limits>

#include "mo_kernels.h"

namespace mo {
void density_advance_box(const mgpu_params& P, int which_step, Arr& sold, Arr& snew, Arr* sedge, Arr* sflux,
                         Arr& scal_force, Arr* umac, const double* w0, Arr& etarhoflux, const double* rho0_old,
                         const double* rho0_new, const double* rho0_predicted_edge, const int* lo,
                         const int* hi, int ng_s, int ng_f, const int* adv_bc, const int* pmask);
void test_advect_run(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                     double* abs_norm, double* rel_norm, double* rho_final_out);
void test_advect_run_ex(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                        const double* opt, double* abs_norm, double* rel_norm, double* rho_final_out,
                        double* rho_init_out);
extern bool g_ppm1_no_edge_clip;
extern int g_ppm1_variant;
}  // namespace mo

using namespace mo;

std::string mo_g_err;  // shared with mo_sphr.cpp
#define g_err mo_g_err

#define MO_TRY try {
#define MO_CATCH                         \
  }                                      \
  catch (const std::exception& e) {      \
    g_err = e.what();                    \
    return 1;                            \
  }                                      \
  return 0;

static void views(const mgpu_params* p, const mgpu_fab* const* f, int i, Arr* out) {
  for (int d = 0; d < p->dm; ++d) out[d] = Arr::view(f[d][i], p->dm);
}

extern "C" {

const char* mo_last_error(void) { return g_err.c_str(); }

int mo_fill_boundary(const mgpu_params* p, mgpu_fab* s, int scomp, int bccomp, int ncomp, const int* adv_bc,
                     const int* pmask) {
  MO_TRY
  Arr a = Arr::view(*s, p->dm);
  if (s->nodal[0] || s->nodal[1] || s->nodal[2]) {
    int dir = s->nodal[0] ? 0 : (s->nodal[1] ? 1 : 2);
    fill_boundary_face(*p, a, s->lo, s->hi, s->ng, dir, pmask);
  } else {
    fill_boundary_box(*p, a, s->lo, s->hi, s->ng, scomp, bccomp, ncomp, adv_bc, pmask);
  }
  MO_CATCH
}

int mo_make_edge_scal(const mgpu_params* p, int nfabs, const mgpu_fab* s, mgpu_fab* const* sedge,
                      const mgpu_fab* const* umac, const mgpu_fab* force, const int* adv_bc, int is_vel,
                      int start_scomp, int start_bccomp, int num_comp, int is_conservative) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr sa = Arr::view(s[i], p->dm), fa = Arr::view(force[i], p->dm);
    Arr se[3], um[3];
    views(p, (const mgpu_fab* const*)sedge, i, se);
    views(p, umac, i, um);
    for (int scomp = start_scomp; scomp < start_scomp + num_comp; ++scomp) {
      int bccomp = start_bccomp + scomp - start_scomp;
      make_edge_scal_box(*p, sa, se, um, fa, s[i].lo, s[i].hi, adv_bc, scomp - 1, bccomp, is_vel != 0,
                         is_conservative != 0, s[i].ng);
    }
  }
  MO_CATCH
}

int mo_bds(const mgpu_params* p, int nfabs, const mgpu_fab* s, mgpu_fab* const* sedge,
           const mgpu_fab* const* umac, const mgpu_fab* force, const int* adv_bc, int is_vel, int start_scomp,
           int start_bccomp, int num_comp, int is_conservative) {
  MO_TRY
  (void)adv_bc; (void)is_vel; (void)start_bccomp;
  for (int i = 0; i < nfabs; ++i) {
    Arr sa = Arr::view(s[i], p->dm), fa = Arr::view(force[i], p->dm);
    Arr se[3], um[3];
    views(p, (const mgpu_fab* const*)sedge, i, se);
    views(p, umac, i, um);
    for (int scomp = start_scomp; scomp < start_scomp + num_comp; ++scomp)
      bds_box(*p, sa, se, um, fa, s[i].lo, s[i].hi, scomp - 1, is_conservative != 0);
  }
  MO_CATCH
}

int mo_mk_rhoX_flux(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, mgpu_fab* etarhoflux,
                    const mgpu_fab* const* sedge, const mgpu_fab* const* umac, const double* w0,
                    const double* rho0_old, const double* rho0_edge_old, const double* rho0_new,
                    const double* rho0_edge_new, const double* rho0_predicted_edge, int startcomp, int endcomp) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr sf[3], se[3], um[3];
    views(p, (const mgpu_fab* const*)sflux, i, sf);
    views(p, sedge, i, se);
    views(p, umac, i, um);
    Arr eta = Arr::view(etarhoflux[i], p->dm);
    mk_rhoX_flux_box(*p, sf, eta, se, um, w0, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new,
                     rho0_predicted_edge, startcomp, endcomp, umac[0][i].lo, umac[0][i].hi);
  }
  MO_CATCH
}

int mo_mk_rhoh_flux(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                    const mgpu_fab* const* umac, const double* w0, const double* rho0_old,
                    const double* rho0_edge_old, const double* rho0_new, const double* rho0_edge_new,
                    const double* rhoh0_old, const double* rhoh0_edge_old, const double* rhoh0_new,
                    const double* rhoh0_edge_new) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr sf[3], se[3], um[3];
    views(p, (const mgpu_fab* const*)sflux, i, sf);
    views(p, sedge, i, se);
    views(p, umac, i, um);
    mk_rhoh_flux_box(*p, sf, se, um, w0, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new, rhoh0_old,
                     rhoh0_edge_old, rhoh0_new, rhoh0_edge_new, umac[0][i].lo, umac[0][i].hi);
  }
  MO_CATCH
}

int mo_update_scal(const mgpu_params* p, int nfabs, int nstart, int nstop, const mgpu_fab* sold, mgpu_fab* snew,
                   const mgpu_fab* const* sflux, const mgpu_fab* force, const double* p0_new,
                   const mgpu_fab* p0_new_cart) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr so = Arr::view(sold[i], p->dm), sn = Arr::view(snew[i], p->dm), fa = Arr::view(force[i], p->dm);
    Arr sf[3];
    views(p, sflux, i, sf);
    Arr pc;
    if (p0_new_cart) pc = Arr::view(p0_new_cart[i], p->dm);
    update_scal_box(*p, nstart, nstop, so, sn, sf, fa, sold[i].lo, sold[i].hi, p0_new, p0_new_cart ? &pc : nullptr);
  }
  MO_CATCH
}

int mo_update_velocity(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                       const mgpu_fab* const* umac, const mgpu_fab* const* uedge, const mgpu_fab* force,
                       const mgpu_fab* sponge, const double* w0) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr uo = Arr::view(uold[i], p->dm), un = Arr::view(unew[i], p->dm), fa = Arr::view(force[i], p->dm);
    Arr sp = Arr::view(sponge[i], p->dm);
    Arr um[3], ue[3];
    views(p, umac, i, um);
    views(p, uedge, i, ue);
    update_velocity_box(*p, uo, un, um, ue, fa, sp, w0, uold[i].lo, uold[i].hi);
  }
  MO_CATCH
}

int mo_addw0(const mgpu_params* p, int nfabs, mgpu_fab* const* umac, const double* w0, double mult) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr um[3];
    views(p, (const mgpu_fab* const*)umac, i, um);
    addw0_box(*p, um, w0, mult, umac[0][i].lo, umac[0][i].hi);
  }
  MO_CATCH
}

int mo_mkutrans(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                mgpu_fab* const* utrans, const double* w0, const int* adv_bc, const int* phys_bc) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr ut = Arr::view(utilde[i], p->dm), uf = Arr::view(ufull[i], p->dm);
    Arr tr[3];
    views(p, (const mgpu_fab* const*)utrans, i, tr);
    mkutrans_box(*p, ut, uf, tr, w0, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng);
  }
  MO_CATCH
}

int mo_velpred(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
               mgpu_fab* const* umac, const mgpu_fab* const* utrans, const mgpu_fab* force, const double* w0,
               const int* adv_bc, const int* phys_bc) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr ut = Arr::view(utilde[i], p->dm), uf = Arr::view(ufull[i], p->dm), fa = Arr::view(force[i], p->dm);
    Arr um[3], tr[3];
    views(p, (const mgpu_fab* const*)umac, i, um);
    views(p, utrans, i, tr);
    velpred_box(*p, ut, uf, um, tr, fa, w0, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng);
  }
  MO_CATCH
}

int mo_modify_scal_force(const mgpu_params* p, int nfabs, mgpu_fab* force, const mgpu_fab* s,
                         const mgpu_fab* const* umac, const double* s0, const double* s0_edge, const double* w0,
                         int comp, int fullform) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr fa = Arr::view(force[i], p->dm), sa = Arr::view(s[i], p->dm);
    Arr um[3];
    views(p, umac, i, um);
    modify_scal_force_box(*p, fa, sa, um, s0, s0_edge, w0, comp, fullform != 0, s[i].lo, s[i].hi);
  }
  MO_CATCH
}

int mo_convert_rhoX_to_X(const mgpu_params* p, int nfabs, mgpu_fab* s, int flag) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr sa = Arr::view(s[i], p->dm);
    Box vb = grown(s[i].lo, s[i].hi, p->dm, 0);
    for (int n = 0; n < p->nspec; ++n) {
      const int c = p->spec_comp - 1 + n, r = p->rho_comp - 1;
      for_box(vb, [&](int ii, int j, int k) {
        sa(ii, j, k, c) = flag ? sa(ii, j, k, c) / sa(ii, j, k, r) : sa(ii, j, k, c) * sa(ii, j, k, r);
      });
    }
  }
  MO_CATCH
}

int mo_put_in_pert_form(const mgpu_params* p, int nfabs, mgpu_fab* s, const double* base, int comp, int flag) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr sa = Arr::view(s[i], p->dm);
    Box vb = grown(s[i].lo, s[i].hi, p->dm, 0);
    const int mult = flag ? -1 : +1, r = p->dm - 1;
    for_box(vb, [&](int ii, int j, int k) { sa(ii, j, k, comp - 1) = sa(ii, j, k, comp - 1) + mult * base[r == 1 ? j : k]; });
  }
  MO_CATCH
}

int mo_cell_to_edge(const double* s0_cell, double* s0_edge, int nr) {
  MO_TRY
  cell_to_edge(s0_cell, s0_edge, nr);
  MO_CATCH
}

int mo_density_advance(const mgpu_params* p, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                       mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                       mgpu_fab* const* umac, const double* w0, mgpu_fab* etarhoflux, const double* rho0_old,
                       const double* rho0_new, const double* p0_dummy, const double* rho0_predicted_edge,
                       const int* adv_bc, const int* pmask) {
  MO_TRY
  (void)p0_dummy;
  Arr so = Arr::view(*sold, p->dm), sn = Arr::view(*snew, p->dm), fa = Arr::view(*scal_force, p->dm);
  Arr eta = Arr::view(*etarhoflux, p->dm);
  Arr se[3], sf[3], um[3];
  views(p, (const mgpu_fab* const*)sedge, 0, se);
  views(p, (const mgpu_fab* const*)sflux, 0, sf);
  views(p, (const mgpu_fab* const*)umac, 0, um);
  density_advance_box(*p, which_step, so, sn, se, sf, fa, um, w0, eta, rho0_old, rho0_new, rho0_predicted_edge,
                      sold->lo, sold->hi, sold->ng, scal_force->ng, adv_bc, pmask);
  MO_CATCH
}

int mo_mkrhohforce(const mgpu_params* p, int nfabs, mgpu_fab* scal_force, int is_prediction, const mgpu_fab* thermal,
                   const mgpu_fab* const* umac, const double* p0_1, const double* p0_2, const double* rho0_1,
                   const double* rho0_2, const double* grav, const double* psi, int add_thermal) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr f = Arr::view(scal_force[i], p->dm), th = Arr::view(thermal[i], p->dm);
    Arr um[3];
    views(p, umac, i, um);
    mkrhohforce_box(*p, f, is_prediction != 0, th, um, p0_1, p0_2, rho0_1, rho0_2, grav, psi, add_thermal != 0,
                    scal_force[i].lo, scal_force[i].hi);
  }
  MO_CATCH
}

int mo_mk_vel_force(const mgpu_params* p, int nfabs, mgpu_fab* vel_force, int is_final_update, const mgpu_fab* uold,
                    const mgpu_fab* const* uedge, const double* w0, const mgpu_fab* gpi, const mgpu_fab* s,
                    int index_rho, const double* rho0, const double* grav, const double* w0_force,
                    int do_add_utilde_force) {
  MO_TRY
  for (int i = 0; i < nfabs; ++i) {
    Arr f = Arr::view(vel_force[i], p->dm), uo = Arr::view(uold[i], p->dm), gp = Arr::view(gpi[i], p->dm);
    Arr sa = Arr::view(s[i], p->dm);
    Arr ue[3];
    views(p, uedge, i, ue);
    mk_vel_force_box(*p, f, is_final_update != 0, uo, ue, w0, gp, sa.comp(index_rho - 1), rho0, grav, w0_force,
                     vel_force[i].lo, vel_force[i].hi, do_add_utilde_force != 0);
  }
  MO_CATCH
}

int mo_advance_premac(const mgpu_params* p, const mgpu_fab* uold, const mgpu_fab* sold, mgpu_fab* const* umac,
                      const mgpu_fab* gpi, const double* w0, const double* w0_force, const double* rho0_old,
                      const double* grav_cell_old, const int* adv_bc, const int* phys_bc, const int* pmask) {
  MO_TRY
  Arr uo = Arr::view(*uold, p->dm), so = Arr::view(*sold, p->dm), gp = Arr::view(*gpi, p->dm);
  Arr um[3];
  views(p, (const mgpu_fab* const*)umac, 0, um);
  advance_premac_box(*p, uo, so, um, gp, w0, w0_force, rho0_old, grav_cell_old, uold->lo, uold->hi, uold->ng, adv_bc,
                     phys_bc, pmask);
  MO_CATCH
}

int mo_velocity_advance(const mgpu_params* p, const mgpu_fab* uold, mgpu_fab* unew, const mgpu_fab* sold,
                        const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0,
                        const double* w0_force, const double* rho0_old, const double* rho0_nph,
                        const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                        const int* adv_bc, const int* pmask) {
  MO_TRY
  Arr uo = Arr::view(*uold, p->dm), un = Arr::view(*unew, p->dm), so = Arr::view(*sold, p->dm);
  Arr rh = Arr::view(*rhohalf, p->dm), gp = Arr::view(*gpi, p->dm), sp = Arr::view(*sponge, p->dm);
  Arr um[3];
  views(p, (const mgpu_fab* const*)umac, 0, um);
  velocity_advance_box(*p, uo, un, so, rh, um, gp, w0, w0_force, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph, sp,
                       uold->lo, uold->hi, uold->ng, adv_bc, pmask);
  MO_CATCH
}

int mo_enthalpy_advance(const mgpu_params* p, int which_step, mgpu_fab* sold, mgpu_fab* snew, mgpu_fab* const* sedge,
                        mgpu_fab* const* sflux, mgpu_fab* scal_force, const mgpu_fab* thermal, mgpu_fab* const* umac,
                        const double* w0, const double* rho0_old, const double* rhoh0_old, const double* rho0_new,
                        const double* rhoh0_new, const double* p0_old, const double* p0_new, const double* tempbar,
                        const double* psi, const double* grav_old, const double* grav_nph, const int* adv_bc,
                        const int* pmask) {
  MO_TRY
  Arr so = Arr::view(*sold, p->dm), sn = Arr::view(*snew, p->dm), fa = Arr::view(*scal_force, p->dm);
  Arr th = Arr::view(*thermal, p->dm);
  Arr se[3], sf[3], um[3];
  views(p, (const mgpu_fab* const*)sedge, 0, se);
  views(p, (const mgpu_fab* const*)sflux, 0, sf);
  views(p, (const mgpu_fab* const*)umac, 0, um);
  enthalpy_advance_box(*p, which_step, so, sn, se, sf, fa, th, um, w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new,
                       p0_old, p0_new, tempbar, psi, grav_old, grav_nph, sold->lo, sold->hi, sold->ng, scal_force->ng,
                       adv_bc, pmask);
  MO_CATCH
}

int mo_test_advect(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                   double* abs_norm, double* rel_norm, double* rho_final) {
  MO_TRY
  test_advect_run(dm, n, ppm_type, bds_type, itest_dir, cflfac, stop_time, abs_norm, rel_norm, rho_final);
  MO_CATCH
}

void mo_set_ppm1_no_edge_clip(int on) { mo::g_ppm1_no_edge_clip = on != 0; }
void mo_set_ppm1_variant(int v) { mo::g_ppm1_variant = v; }

int mo_test_advect_ex(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                      const double* opt, double* abs_norm, double* rel_norm, double* rho_final, double* rho_init) {
  MO_TRY
  test_advect_run_ex(dm, n, ppm_type, bds_type, itest_dir, cflfac, stop_time, opt, abs_norm, rel_norm, rho_final,
                     rho_init);
  MO_CATCH
}

int mo_estdt(const mgpu_params* p, int nfabs, const mgpu_fab* u, const mgpu_fab* s, const mgpu_fab* force,
             const mgpu_fab* divU, const mgpu_fab* dSdt, const double* w0, const double* p0, const double* gamma1bar,
             double rho_min, double cflfac, double* dt, double* umax) {
  MO_TRY
  estdt_level(*p, nfabs, u, s, force, divU, dSdt, w0, p0, gamma1bar, rho_min, cflfac, *dt, *umax);
  MO_CATCH
}

// multifab_min_c / multifab_max_c over the valid zones (FBoxLib multifab_f.f90 -- not in /root/reference; the call
// sites are density_advance.f90:374-402 with the multifab_div_div_c / multifab_mult_mult_c pair around the species,
// enthalpy_advance.f90:440-449, velocity_advance.f90:142-160).  Unpinned: the reference holds no numbers for them.
int mo_minmax(const mgpu_params* p, int nfabs, const mgpu_fab* s, int comp, int div_comp, double* smin, double* smax) {
  MO_TRY
  double mn = std::numeric_limits<double>::max(), mx = -std::numeric_limits<double>::max();
  for (int n = 0; n < nfabs; ++n) {
    if (comp < 1 || comp > s[n].nc || div_comp > s[n].nc) fail("mo_minmax: component out of range");
    Arr a = Arr::view(s[n], p->dm);
    const int* lo = s[n].lo;
    const int* hi = s[n].hi;
    const int k0 = p->dm == 3 ? lo[2] : 0, k1 = p->dm == 3 ? hi[2] : 0;
    for (int k = k0; k <= k1; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          double v = a(i, j, k, comp - 1);
          if (div_comp >= 1) v = v / a(i, j, k, div_comp - 1);
          mn = std::min(mn, v);
          mx = std::max(mx, v);
        }
  }
  *smin = mn;
  *smax = mx;
  MO_CATCH
}

int mo_make_etarho_planar(const mgpu_params* p, int nfabs, const mgpu_fab* etarhoflux, double* etarho_ec,
                          double* etarho_cc) {
  MO_TRY
  make_etarho_planar(*p, nfabs, etarhoflux, etarho_ec, etarho_cc);
  MO_CATCH
}

int mo_estdt_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* u, const mgpu_fab* s,
                  const mgpu_fab* force, const mgpu_fab* divU, const mgpu_fab* dSdt, const mgpu_fab* const* w0mac,
                  const double* w0, const double* p0, const double* gamma1bar, double rho_min, double cflfac, double* dt,
                  double* umax) {
  MO_TRY
  estdt_sphr_level(*p, *g, nfabs, u, s, force, divU, dSdt, w0mac, w0, p0, gamma1bar, rho_min, cflfac, *dt, *umax);
  MO_CATCH
}

}  // extern "C"
