// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
// extern "C" surface of the EOS pieces (mo_eos.cpp), mirroring the mgpu_ entry points of include/maestro_b200.h.
#include <stdexcept>
#include <string>
#include <vector>

#include "mo_kernels.h"

using namespace mo;

extern std::string mo_g_err;
#define MO_TRY try {
#define MO_CATCH                         \
  }                                      \
  catch (const std::exception& e) {      \
    mo_g_err = e.what();                 \
    return 1;                            \
  }                                      \
  return 0;

static void views(const mgpu_params* p, const mgpu_fab* const* f, int i, Arr* out) {
  for (int d = 0; d < p->dm; ++d) out[d] = Arr::view(f[d][i], p->dm);
}
static void need_geom(const mgpu_params* p, const mgpu_geom* g, const char* who) {
  if (p->spherical && (!g || p->dm != 3)) fail((std::string(who) + ": spherical geometry needs mgpu_geom (3-D)").c_str());
}

extern "C" {

int mo_set_eos(const mgpu_eos* e) {
  MO_TRY
  set_eos(e);
  MO_CATCH
}

int mo_eos_eval(int input, long n, double* state, const double* xn) {
  MO_TRY
  const mgpu_eos& E = the_eos("eos");
  for (long i = 0; i < n; ++i) {
    EosState s{};
    double x[MGPU_EOS_MAXSPEC];
    for (int q = 0; q < E.nspec; ++q) x[q] = xn[(long)q * n + i];
    s.rho = state[MGPU_EOS_Q_RHO * n + i];
    s.T = state[MGPU_EOS_Q_T * n + i];
    s.p = state[MGPU_EOS_Q_P * n + i];
    s.e = state[MGPU_EOS_Q_E * n + i];
    s.h = state[MGPU_EOS_Q_H * n + i];
    eos_call(input, s, x);
    const double out[MGPU_EOS_NQ] = {s.rho, s.T, s.p, s.e, s.h, s.cv, s.cp, s.cs, s.dpdT, s.dpdr, s.dedT, s.dedr,
                                     s.dhdT, s.mu, s.abar, s.zbar};
    for (int q = 0; q < MGPU_EOS_NQ; ++q) state[(long)q * n + i] = out[q];
  }
  MO_CATCH
}

int mo_make_h_from_rhot_edge(const mgpu_params* p, int nfabs, mgpu_fab* const* sedge, const double* rho0_old,
                             const double* rhoh0_old, const double* t0_old, const double* rho0_edge_old,
                             const double* rhoh0_edge_old, const double* t0_edge_old, const double* rho0_new,
                             const double* rhoh0_new, const double* t0_new, const double* rho0_edge_new,
                             const double* rhoh0_edge_new, const double* t0_edge_new) {
  MO_TRY
  if (p->spherical) fail("make_h_from_rhot_edge: use the _sphr entry point with spherical == 1");
  for (int i = 0; i < nfabs; ++i) {
    Arr se[3];
    views(p, (const mgpu_fab* const*)sedge, i, se);
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d) { lo[d] = sedge[0][i].lo[d]; hi[d] = sedge[0][i].hi[d]; }
    h_from_rhot_edge_box(*p, se, rho0_old, rhoh0_old, t0_old, rho0_edge_old, rhoh0_edge_old, t0_edge_old, rho0_new,
                         rhoh0_new, t0_new, rho0_edge_new, rhoh0_edge_new, t0_edge_new, lo, hi);
  }
  MO_CATCH
}

int mo_make_h_from_rhot_edge_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* const* sedge,
                                  const double* rho0_old, const double* rhoh0_old, const double* t0_old,
                                  const double* rho0_new, const double* rhoh0_new, const double* t0_new,
                                  const int* adv_bc, const int* pmask) {
  MO_TRY
  need_geom(p, g, "make_h_from_rhot_edge_sphr");
  const int nr = g->nr_fine, dm = 3;
  std::vector<double> r0h(nr), rh0h(nr), t0h(nr);
  for (int r = 0; r < nr; ++r) {  // rhoh_vs_t.f90:91-95
    r0h[r] = 0.5 * (rho0_old[r] + rho0_new[r]);
    rh0h[r] = 0.5 * (rhoh0_old[r] + rhoh0_new[r]);
    t0h[r] = 0.5 * (t0_old[r] + t0_new[r]);
  }
  for (int i = 0; i < nfabs; ++i) {
    Arr se[3];
    views(p, (const mgpu_fab* const*)sedge, i, se);
    const int* lo = sedge[0][i].lo;
    const int* hi = sedge[0][i].hi;
    Arr r0c = cart_with_ghosts(*p, *g, r0h.data(), 2, dm + p->rho_comp, lo, hi, adv_bc, pmask);
    Arr rh0c = cart_with_ghosts(*p, *g, rh0h.data(), 2, dm + p->rhoh_comp, lo, hi, adv_bc, pmask);
    Arr t0c = cart_with_ghosts(*p, *g, t0h.data(), 2, dm + p->temp_comp, lo, hi, adv_bc, pmask);
    h_from_rhot_edge_sphr_box(*p, se, r0c, rh0c, t0c, lo, hi);
  }
  MO_CATCH
}

int mo_mktempforce(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* temp_force,
                   const mgpu_fab* const* umac, const mgpu_fab* s, const mgpu_fab* thermal, const double* p0_old,
                   const double* psi, const int* adv_bc, const int* pmask) {
  MO_TRY
  need_geom(p, g, "mktempforce");
  const int ept = p->enthalpy_pred_type;
  if (!(ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H))
    fail("ERROR: should only call mkrhohforce when predicting T or T'");  // mkscalforce.f90:756 (the reference's wording)
  const int dm = p->dm, foextrap_comp = dm + p->nscal + 2;
  for (int i = 0; i < nfabs; ++i) {
    Arr f = Arr::view(temp_force[i], dm), sa = Arr::view(s[i], dm), th = Arr::view(thermal[i], dm);
    Arr um[3];
    views(p, umac, i, um);
    const int* lo = s[i].lo;
    const int* hi = s[i].hi;
    if (p->spherical) {
      Arr p0_cart = cart_with_ghosts(*p, *g, p0_old, 1, foextrap_comp, lo, hi, adv_bc, pmask);
      Arr psi_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
      put_1d_array_on_cart_sphr(*p, *g, false, false, psi, psi_cart, lo, hi);
      mktempforce_sphr_box(*p, f, sa, um, th, p0_cart, psi_cart, lo, hi);
    } else {
      mktempforce_box(*p, f, sa, um, th, p0_old, psi, lo, hi);
    }
    fill_boundary_box(*p, f, lo, hi, temp_force[i].ng, p->temp_comp, foextrap_comp, 1, adv_bc, pmask);  // :833-837
  }
  MO_CATCH
}

int mo_firstdt(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* u, const mgpu_fab* gpi,
               const mgpu_fab* s, const mgpu_fab* divU, const double* rho0, const double* p0, const double* grav,
               const double* gamma1bar, double cflfac, double init_shrink, int use_soundspeed_firstdt,
               int use_divu_firstdt, double* dt, double* umax) {
  MO_TRY
  need_geom(p, g, "firstdt");
  const int dm = p->dm;
  const int nr = p->spherical ? g->nr_fine : p->nr;
  std::vector<double> w0_dummy(nr + 1, 0.0), w0_force_dummy(nr, 0.0);
  double dt_proc = 1.e99, umax_proc = 0.0;  // firstdt.f90:134-137
  for (int i = 0; i < nfabs; ++i) {
    const int* lo = u[i].lo;
    const int* hi = u[i].hi;
    Arr ua = Arr::view(u[i], dm), gp = Arr::view(gpi[i], dm), sa = Arr::view(s[i], dm), dU = Arr::view(divU[i], dm);
    Box fb = grown(lo, hi, dm, 1);
    Arr force(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], dm);
    Arr umd[3];
    for (int d = 0; d < dm; ++d) {  // umac_dummy: zero, one ghost layer (:84-88)
      Box b = grown(lo, hi, dm, 1);
      b.hi[d] += 1;
      umd[d].alloc(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], 1);
    }
    double dt_grid, umax_grid;
    if (p->spherical) {
      Arr w0md[3];
      for (int d = 0; d < 3; ++d) w0md[d].alloc(umd[d].lo[0], umd[d].lo[0] + umd[d].n[0] - 1, umd[d].lo[1],
                                                umd[d].lo[1] + umd[d].n[1] - 1, umd[d].lo[2],
                                                umd[d].lo[2] + umd[d].n[2] - 1, 1);
      Arr normal_dummy(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], 3);
      Arr w0fc(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], 3);
      mk_vel_force_sphr_box(*p, *g, force, false, ua, umd, w0_dummy.data(), w0md, gp, sa.comp(p->rho_comp - 1),
                            normal_dummy, rho0, grav, w0fc, lo, hi, false);  // :96-100
      // gp0 on the radial edges and as a Cartesian vector (:708-721)
      std::vector<double> gp0(nr + 1);
      for (int r = 1; r < nr; ++r) {
        const double gamma1bar_p_avg = 0.5 * (gamma1bar[r] * p0[r] + gamma1bar[r - 1] * p0[r - 1]);
        gp0[r] = ((p0[r] - p0[r - 1]) / g->dr) / gamma1bar_p_avg;
      }
      gp0[nr] = gp0[nr - 1];
      gp0[0] = gp0[1];
      Arr gp0_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 3);
      put_1d_array_on_cart_sphr(*p, *g, true, true, gp0.data(), gp0_cart, lo, hi);
      firstdt_box(*p, ua, sa, force, dU, p0, gamma1bar, &gp0_cart, lo, hi, cflfac, use_soundspeed_firstdt != 0,
                  use_divu_firstdt != 0, dt_grid, umax_grid);
    } else {
      mk_vel_force_box(*p, force, false, ua, umd, w0_dummy.data(), gp, sa.comp(p->rho_comp - 1), rho0, grav,
                       w0_force_dummy.data(), lo, hi, false);
      firstdt_box(*p, ua, sa, force, dU, p0, gamma1bar, nullptr, lo, hi, cflfac, use_soundspeed_firstdt != 0,
                  use_divu_firstdt != 0, dt_grid, umax_grid);
    }
    dt_proc = dmin(dt_proc, dt_grid);
    umax_proc = dmax(umax_proc, umax_grid);
  }
  *umax = dmax(*umax, umax_proc);          // :175
  const double dt_lev = dt_proc * init_shrink;  // :182
  *dt = dmin(*dt, dt_lev);                 // :189
  MO_CATCH
}

static void make_t(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0, bool from_p,
                   bool flag, bool update_rhoh, const int* adv_bc, const int* pmask) {
  need_geom(p, g, from_p ? "makeTfromRhoP" : "makeTfromRhoH");
  const int dm = p->dm;
  for (int i = 0; i < nfabs; ++i) {
    Arr st = Arr::view(state[i], dm);
    const int* lo = state[i].lo;
    const int* hi = state[i].hi;
    if (p->spherical) {
      Arr p0_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
      put_1d_array_on_cart_sphr(*p, *g, false, false, p0, p0_cart, lo, hi);
      make_t_box(*p, st, nullptr, &p0_cart, from_p, flag, update_rhoh, lo, hi);
    } else {
      make_t_box(*p, st, p0, nullptr, from_p, flag, update_rhoh, lo, hi);
    }
    fill_boundary_box(*p, st, lo, hi, state[i].ng, p->temp_comp, dm + p->temp_comp, 1, adv_bc, pmask);
    if (from_p && update_rhoh)
      fill_boundary_box(*p, st, lo, hi, state[i].ng, p->rhoh_comp, dm + p->rhoh_comp, 1, adv_bc, pmask);
  }
}

int mo_make_t_from_rhoh(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0,
                        int use_eos_e_instead_of_h, const int* adv_bc, const int* pmask) {
  MO_TRY
  make_t(p, g, nfabs, state, p0, false, use_eos_e_instead_of_h != 0, false, adv_bc, pmask);
  MO_CATCH
}

int mo_make_t_from_rhop(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0,
                        int update_rhoh, int use_pprime_in_tfromp, const int* adv_bc, const int* pmask) {
  MO_TRY
  make_t(p, g, nfabs, state, p0, true, use_pprime_in_tfromp != 0, update_rhoh != 0, adv_bc, pmask);
  MO_CATCH
}

// ---- average / make_etarho_spherical (average.f90:24, make_eta.f90:256) -----------------------------------------------
int mo_average(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* phi, int incomp, int nr_irreg,
               int drdxfac, double* phibar) {
  MO_TRY
  average_level(*p, g, nfabs, phi, incomp, nr_irreg, drdxfac, phibar);
  MO_CATCH
}

int mo_make_etarho_spherical(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* sold,
                             const mgpu_fab* snew, const mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                             const double* rho0_old, const double* rho0_new, const mgpu_fab* normal, int nr_irreg,
                             int drdxfac, double* etarho_ec, double* etarho_cc) {
  MO_TRY
  if (!p->spherical) fail("ERROR: make_eta_spherical should not be called for plane-parallel");  // make_eta.f90:289
  need_geom(p, g, "make_etarho_spherical");
  const int nr = g->nr_fine, rho = p->rho_comp - 1;
  std::vector<double> rho0_nph(nr);
  for (int r = 0; r < nr; ++r) rho0_nph[r] = 0.5 * (rho0_old[r] + rho0_new[r]);  // :376-378
  std::vector<Arr> eta(nfabs);
  std::vector<mgpu_fab> ef(nfabs);
  for (int i = 0; i < nfabs; ++i) {  // construct_eta_cart, :345-408
    const int* lo = sold[i].lo;
    const int* hi = sold[i].hi;
    Arr so = Arr::view(sold[i], 3), sn = Arr::view(snew[i], 3), nm = Arr::view(normal[i], 3);
    Arr um[3], wm[3];
    views(p, umac, i, um);
    views(p, w0mac, i, wm);
    Arr nph(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
    put_1d_array_on_cart_sphr(*p, *g, false, false, rho0_nph.data(), nph, lo, hi);
    eta[i].alloc(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
    Arr& e = eta[i];
    for_box(grown(lo, hi, 3, 0), [&](int ii, int j, int k) {
      const double U_dot_er =
          0.5 * (um[0](ii, j, k) + um[0](ii + 1, j, k) + wm[0](ii, j, k) + wm[0](ii + 1, j, k)) * nm(ii, j, k, 0) +
          0.5 * (um[1](ii, j, k) + um[1](ii, j + 1, k) + wm[1](ii, j, k) + wm[1](ii, j + 1, k)) * nm(ii, j, k, 1) +
          0.5 * (um[2](ii, j, k) + um[2](ii, j, k + 1) + wm[2](ii, j, k) + wm[2](ii, j, k + 1)) * nm(ii, j, k, 2);
      e(ii, j, k) = (0.5 * (so(ii, j, k, rho) + sn(ii, j, k, rho)) - nph(ii, j, k)) * U_dot_er;
    });
    ef[i] = sold[i];
    ef[i].ptr = e.p;
    ef[i].ng = 0;
    ef[i].nc = 1;
  }
  average_level(*p, g, nfabs, ef.data(), 1, nr_irreg, drdxfac, etarho_cc);  // :327
  etarho_ec[0] = 0.0;                                                       // :337-343
  for (int r = 1; r < nr; ++r) etarho_ec[r] = 0.5 * (etarho_cc[r] + etarho_cc[r - 1]);
  etarho_ec[nr] = etarho_cc[nr - 1];
  MO_CATCH
}

}  // extern "C"
