// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of the L4 driver Source/density_advance.f90:20 (planar geometry, one level, one box
// covering the domain) and of the unit-test driver Exec/UNIT_TESTS/test_advect/varden.f90:16 +
// test_advect.f90:58.  Parity status: the driver's own pass/fail criterion (direction independence to
// advect_test_tol) is gated in tests/; the archived report advect_3d_report_example.out predates the driver in the
// tree and is NOT reproduced -- parity against reference-produced numbers is unpinned.
#include <stdexcept>
#include <string>

#include "mo_kernels.h"

namespace mo {

void fail(const char* msg) { throw std::runtime_error(msg); }

static void fill_umac(const mgpu_params& P, Arr* umac, const int* lo, const int* hi, const int* pmask) {
  // addw0.f90:85-93 / mkutrans.f90:105-115: multifab_fill_boundary on each face-centred component.
  // multifab_physbc_edgevel (FBoxLib) is not restated: wall face-ghosts never reach a valid output
  // ("Which ghost values reach valid outputs").
  for (int d = 0; d < P.dm; ++d) fill_boundary_face(P, umac[d], lo, hi, 1, d, pmask);
}

void density_advance_box(const mgpu_params& P, int which_step, Arr& sold, Arr& snew, Arr* sedge, Arr* sflux,
                         Arr& scal_force, Arr* umac, const double* w0, Arr& etarhoflux, const double* rho0_old,
                         const double* rho0_new, const double* rho0_predicted_edge, const int* lo,
                         const int* hi, int ng_s, int ng_f, const int* adv_bc, const int* pmask) {
  const int dm = P.dm;
  const int nr = P.nr;
  const int spt = P.species_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;  // variables.f90:119-121
  std::vector<double> rho0_edge_old(nr + 1), rho0_edge_new(nr + 1);
  cell_to_edge(rho0_old, rho0_edge_old.data(), nr);  // density_advance.f90:90-91
  cell_to_edge(rho0_new, rho0_edge_new.data(), nr);

  scal_force.fill(0.0);  // :101-103

  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :119-128
    modify_scal_force_box(P, scal_force, sold, umac, rho0_old, rho0_edge_old.data(), w0, P.rho_comp,
                          spt == MGPU_PREDICT_RHO_AND_X, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rho_comp, foextrap_comp, 1, adv_bc, pmask);
  }

  addw0_box(P, umac, w0, 1.0, lo, hi);  // :148
  fill_umac(P, umac, lo, hi, pmask);

  Box vb = grown(lo, hi, dm, 0);
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :160 convert_rhoX_to_X(.true.)
    for (int n = 0; n < P.nspec; ++n) {
      const int c = P.spec_comp - 1 + n;
      for_box(vb, [&](int i, int j, int k) { sold(i, j, k, c) = sold(i, j, k, c) / sold(i, j, k, P.rho_comp - 1); });
      fill_boundary_box(P, sold, lo, hi, ng_s, c + 1, foextrap_comp, 1, adv_bc, pmask);  // same_boundary
    }
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {  // :166 put_in_pert_form(.true.)
    const int c = P.rho_comp - 1, r = dm - 1;
    for_box(vb, [&](int i, int j, int k) { sold(i, j, k, c) = sold(i, j, k, c) + (-1) * rho0_old[r == 1 ? j : k]; });
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rho_comp, foextrap_comp, 1, adv_bc, pmask);
  }

  auto edge = [&](int scomp, int ncomp, bool cons) {
    for (int n = 0; n < ncomp; ++n) {
      if (P.bds_type == 0)
        make_edge_scal_box(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, scomp - 1 + n, dm + scomp + n, false,
                           cons, ng_s);
      else
        bds_box(P, sold, sedge, umac, scal_force, lo, hi, scomp - 1 + n, cons);
    }
  };

  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :178-186
    edge(P.spec_comp, P.nspec, false);
  } else if (spt == MGPU_PREDICT_RHOX) {  // :190-198
    edge(P.spec_comp, P.nspec, true);
  }

  if (spt == MGPU_PREDICT_RHOX) {  // :204-213
    for (int d = 0; d < dm; ++d) {
      Arr& se = sedge[d];
      Arr r = se.comp(P.rho_comp - 1);
      Arr s1 = se.comp(P.spec_comp - 1);
      for (size_t q = 0; q < r.size(); ++q) r.p[q] = s1.p[q];
      for (int n = 1; n < P.nspec; ++n) {
        Arr sn = se.comp(P.spec_comp - 1 + n);
        for (size_t q = 0; q < r.size(); ++q) r.p[q] = r.p[q] + sn.p[q];
      }
    }
  } else {  // :216-224
    edge(P.rho_comp, 1, false);
  }

  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {  // :229 put_in_pert_form(.false.)
    const int c = P.rho_comp - 1, r = dm - 1;
    for_box(vb, [&](int i, int j, int k) { sold(i, j, k, c) = sold(i, j, k, c) + (+1) * rho0_old[r == 1 ? j : k]; });
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :235 convert_rhoX_to_X(.false.)
    for (int n = 0; n < P.nspec; ++n) {
      const int c = P.spec_comp - 1 + n;
      for_box(vb, [&](int i, int j, int k) { sold(i, j, k, c) = sold(i, j, k, c) * sold(i, j, k, P.rho_comp - 1); });
    }
    fill_boundary_box(P, sold, lo, hi, ng_s, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask);
  }

  if (P.ntrac >= 1) edge(P.trac_comp, P.ntrac, false);  // :242-252

  addw0_box(P, umac, w0, -1.0, lo, hi);  // :258
  fill_umac(P, umac, lo, hi, pmask);

  const double* r0n = (which_step == 1) ? rho0_old : rho0_new;  // :280-329
  const double* r0en = (which_step == 1) ? rho0_edge_old.data() : rho0_edge_new.data();
  mk_rhoX_flux_box(P, sflux, etarhoflux, sedge, umac, w0, rho0_old, rho0_edge_old.data(), r0n, r0en,
                   rho0_predicted_edge, P.spec_comp, P.spec_comp + P.nspec - 1, lo, hi);
  if (P.ntrac >= 1)
    mk_rhoX_flux_box(P, sflux, etarhoflux, sedge, umac, w0, rho0_old, rho0_edge_old.data(), r0n, r0en,
                     rho0_predicted_edge, P.trac_comp, P.trac_comp + P.ntrac - 1, lo, hi);

  scal_force.fill(0.0);  // :349-351

  update_scal_box(P, P.spec_comp, P.spec_comp + P.nspec - 1, sold, snew, sflux, scal_force, lo, hi);  // :360
  fill_boundary_box(P, snew, lo, hi, ng_s, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask);
  fill_boundary_box(P, snew, lo, hi, ng_s, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask);
  if (P.ntrac >= 1) {
    update_scal_box(P, P.trac_comp, P.trac_comp + P.ntrac - 1, sold, snew, sflux, scal_force, lo, hi);
    fill_boundary_box(P, snew, lo, hi, ng_s, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask);
  }
}

// Exec/UNIT_TESTS/test_advect: n^dm periodic unit box, Gaussian density, unit velocity along one axis,
// dt = cfl*dx, density_advance(which_step=1) repeated to stop_time.  Returns |rho_f - rho_i|_2 and
// |(rho_f - rho_i)/rho_i|_2 as plain sqrt(sum of squares) over valid cells (FBoxLib multifab_norm_l2).
// opt (all optional, used by oracle/pin_sweep.py to look for the archived report's configuration):
//   opt[0] W  gaussian width (<= 0: the reference's single-precision 0.05)      opt[1] floor of the initial density
//   (< 0: none)   opt[2] base_cutoff_density   opt[3] species_pred_type   opt[4] 1: keep dt fixed on the last step
//   opt[5] slope_order
void test_advect_run_ex(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                        const double* opt, double* abs_norm, double* rel_norm, double* rho_final_out,
                        double* rho_init_out);
void test_advect_run(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                     double* abs_norm, double* rel_norm, double* rho_final_out) {
  const double opt[6] = {-1.0, 1.e-10, 1.e-10, (double)MGPU_PREDICT_RHOPRIME_AND_X, 0.0, 4.0};
  test_advect_run_ex(dm, n, ppm_type, bds_type, itest_dir, cflfac, stop_time, opt, abs_norm, rel_norm, rho_final_out,
                     nullptr);
}
void test_advect_run_ex(int dm, int n, int ppm_type, int bds_type, int itest_dir, double cflfac, double stop_time,
                        const double* opt, double* abs_norm, double* rel_norm, double* rho_final_out,
                        double* rho_init_out) {
  mgpu_params P;
  memset(&P, 0, sizeof(P));
  P.dm = dm;
  P.ppm_type = ppm_type;
  P.bds_type = bds_type;
  P.slope_order = (int)opt[5];
  P.ppm_trace_forces = 0;
  P.species_pred_type = (int)opt[3];
  P.enthalpy_pred_type = MGPU_PREDICT_RHOHPRIME;
  P.spherical = 0;
  P.evolve_base_state = 1;
  P.nspec = 3;  // general_null ignition.net: C12, O16, Mg24
  P.ntrac = 1;
  P.rho_comp = 1; P.rhoh_comp = 2; P.spec_comp = 3; P.temp_comp = 6; P.pi_comp = 7; P.trac_comp = 8;
  P.nscal = 8;
  P.rel_eps = 0.0;  // never set in test_advect (estdt is not called): static zero
  P.base_cutoff_density = opt[2];
  P.nr = n;
  int lo[3] = {0, 0, 0}, hi[3] = {n - 1, n - 1, dm == 3 ? n - 1 : 0};
  for (int d = 0; d < 3; ++d) { P.domlo[d] = lo[d]; P.domhi[d] = hi[d]; P.dx[d] = 1.0 / n; }
  const int ng_s = 4, ng_f = 1;  // varden.f90:144
  const int nbc = dm + P.nscal + 3;
  std::vector<int> adv_bc(dm * 2 * nbc, MGPU_BC_INTERIOR);
  int pmask[3] = {1, 1, 1};

  Box gb = grown(lo, hi, dm, ng_s);
  Arr sold(gb.lo[0], gb.hi[0], gb.lo[1], gb.hi[1], gb.lo[2], gb.hi[2], P.nscal);
  Arr snew(gb.lo[0], gb.hi[0], gb.lo[1], gb.hi[1], gb.lo[2], gb.hi[2], P.nscal);
  Box fb1 = grown(lo, hi, dm, ng_f);
  Arr scal_force(fb1.lo[0], fb1.hi[0], fb1.lo[1], fb1.hi[1], fb1.lo[2], fb1.hi[2], P.nscal);
  Arr umac[3], sedge[3], sflux[3];
  for (int d = 0; d < dm; ++d) {
    Box ub = grown(lo, hi, dm, 1); ub.hi[d] += 1;
    umac[d].alloc(ub.lo[0], ub.hi[0], ub.lo[1], ub.hi[1], ub.lo[2], ub.hi[2], 1);
    Box eb = grown(lo, hi, dm, 0); eb.hi[d] += 1;
    sedge[d].alloc(eb.lo[0], eb.hi[0], eb.lo[1], eb.hi[1], eb.lo[2], eb.hi[2], P.nscal);
    sflux[d].alloc(eb.lo[0], eb.hi[0], eb.lo[1], eb.hi[1], eb.lo[2], eb.hi[2], P.nscal);
  }
  Box etab = grown(lo, hi, dm, 0); etab.hi[dm - 1] += 1;
  Arr etarhoflux(etab.lo[0], etab.hi[0], etab.lo[1], etab.hi[1], etab.lo[2], etab.hi[2], 1);

  const int idim = (itest_dir < 0 ? -itest_dir : itest_dir) - 1;
  for (int d = 0; d < dm; ++d) umac[d].fill(d == idim ? (itest_dir > 0 ? 1.0 : -1.0) : 0.0);

  std::vector<double> rho0_old(n, 0.0), rho0_new(n, 0.0), w0(n + 1, 0.0), rho0_pe(n + 1, 0.0);

  // test_advect.f90:11 -- W is declared dp but initialised from the single-precision literal 0.05
  const double W = opt[0] > 0.0 ? opt[0] : (double)0.05f;
  Box vb = grown(lo, hi, dm, 0);
  for_box(vb, [&](int i, int j, int k) {
    double x = ((double)i + 0.5) * P.dx[0], y = ((double)j + 0.5) * P.dx[1];
    double xc = 0.5, yc = 0.5, zc = 0.5;
    double dist;
    if (dm == 3) {
      double z = ((double)k + 0.5) * P.dx[2];
      dist = std::sqrt((x - xc) * (x - xc) + (y - yc) * (y - yc) + (z - zc) * (z - zc));
    } else {
      dist = std::sqrt((x - xc) * (x - xc) + (y - yc) * (y - yc));
    }
    double rho = std::exp(-(dist * dist) / (W * W));
    if (opt[1] >= 0.0) rho = dmax(rho, opt[1]);
    sold(i, j, k, P.rho_comp - 1) = rho;
    sold(i, j, k, P.spec_comp - 1) = rho;
  });
  fill_boundary_box(P, sold, lo, hi, ng_s, P.rho_comp, dm + P.rho_comp, P.nscal, adv_bc.data(), pmask);
  Arr dens_orig(vb.lo[0], vb.hi[0], vb.lo[1], vb.hi[1], vb.lo[2], vb.hi[2], 1);
  for_box(vb, [&](int i, int j, int k) { dens_orig(i, j, k) = sold(i, j, k, P.rho_comp - 1); });

  double dt = cflfac * P.dx[0] / 1.0;
  double t = 0.0;
  while (t < stop_time) {
    P.dt = dt;
    density_advance_box(P, 1, sold, snew, sedge, sflux, scal_force, umac, w0.data(), etarhoflux, rho0_old.data(),
                        rho0_new.data(), rho0_pe.data(), lo, hi, ng_s, ng_f, adv_bc.data(), pmask);
    for (size_t q = 0; q < sold.size(); ++q) sold.p[q] = snew.p[q];
    t = t + dt;
    if (opt[4] == 0.0 && t + dt > stop_time) dt = stop_time - t;
  }
  double sa = 0.0, sr = 0.0;
  for (int k = vb.lo[2]; k <= vb.hi[2]; ++k)
    for (int j = vb.lo[1]; j <= vb.hi[1]; ++j)
      for (int i = vb.lo[0]; i <= vb.hi[0]; ++i) {
        double e = snew(i, j, k, P.rho_comp - 1) - dens_orig(i, j, k);
        sa += e * e;
        double r = e / dens_orig(i, j, k);
        sr += r * r;
        if (rho_final_out) rho_final_out[dens_orig.idx(i, j, k)] = snew(i, j, k, P.rho_comp - 1);
        if (rho_init_out) rho_init_out[dens_orig.idx(i, j, k)] = dens_orig(i, j, k);
      }
  *abs_norm = std::sqrt(sa);
  *rel_norm = std::sqrt(sr);
}

}  // namespace mo
