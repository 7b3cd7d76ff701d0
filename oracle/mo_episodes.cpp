// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of the force builders that sit inside the L4 drivers and of the drivers themselves, planar
// geometry, one level, one box covering the domain:
//   mkrhohforce      Source/mkscalforce.f90:31  (_2d :249, _3d :310)
//   mk_vel_force     Source/mkforce.f90:22      (_2d :283, _3d_cart :342)
//   put_1d_array_on_cart (vector form used for ufull) Source/fill_3d_data.f90:21 (_2d :150, _3d)
//   advance_premac   Source/advance_premac.f90:21
//   enthalpy_advance Source/enthalpy_advance.f90:16
//   velocity_advance Source/velocity_advance.f90:16
// 1-D base-state work the reference does on the host inside these routines (make_grav_cell) is an input here.
#include "mo_kernels.h"

namespace mo {

void mkrhohforce_box(const mgpu_params& P, Arr& scal_force, bool is_prediction, const Arr& thermal, const Arr* umac,
                     const double* p0_1, const double* p0_2, const double* rho0_1, const double* rho0_2,
                     const double* grav, const double* psi, bool add_thermal, const int* lo, const int* hi) {
  const int dm = P.dm, r = dm - 1, nr = P.nr;
  const int ept = P.enthalpy_pred_type;
  if (is_prediction && !(ept == MGPU_PREDICT_RHOHPRIME || ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH))
    fail("ERROR: should only call mkrhohforce when predicting rhoh', h, or rhoh");  // mkscalforce.f90:87-92
  std::vector<double> rho0(nr), p0(nr);
  for (int k = 0; k < nr; ++k) {  // :97-98
    rho0[k] = 0.5 * (rho0_1[k] + rho0_2[k]);
    p0[k] = 0.5 * (p0_1[k] + p0_2[k]);
  }
  const double dr = P.dx[r];
  Arr f = scal_force.comp(P.rhoh_comp - 1);
  const Arr& wm = umac[r];
  Box vb = grown(lo, hi, dm, 0);
  const bool with_psi = (is_prediction && ept == MGPU_PREDICT_H) || (is_prediction && ept == MGPU_PREDICT_RHOH) || !is_prediction;
  for_box(vb, [&](int i, int j, int k) {
    const int q = (r == 1) ? j : k;
    double gradp0;
    if (q < P.base_cutoff_density_coord) gradp0 = rho0[q] * grav[q];
    else if (q == nr - 1) gradp0 = (p0[q] - p0[q - 1]) / dr;
    else gradp0 = (p0[q + 1] - p0[q]) / dr;
    const double wadv = 0.5 * (wm(i, j, k) + wm.at(i, j, k, r, 1));
    double v = wadv * gradp0;
    if (with_psi) v = v + psi[q];
    if (add_thermal) v = v + thermal(i, j, k);
    f(i, j, k) = v;
  });
}

void mk_vel_force_box(const mgpu_params& P, Arr& vel_force, bool is_final_update, const Arr& uold, const Arr* uedge,
                      const double* w0, const Arr& gpi, const Arr& rho, const double* rho0, const double* grav,
                      const double* w0_force, const int* lo, const int* hi, bool do_add_utilde_force) {
  const int dm = P.dm, r = dm - 1, nr = P.nr;
  const double dr = P.dx[r];
  vel_force.fill(0.0);
  Box vb = grown(lo, hi, dm, 0);
  const double omega = P.omega, sin_theta = P.sin_theta, cos_theta = P.cos_theta;
  double cen[3];
  cen[0] = -(omega * omega) * P.rotation_radius * sin_theta * sin_theta;
  cen[1] = 0.0;
  cen[2] = (omega * omega) * P.rotation_radius * cos_theta * sin_theta - (omega * omega) * P.rotation_radius;
  for_box(vb, [&](int i, int j, int k) {
    const int q = (r == 1) ? j : k;
    double rhopert = rho(i, j, k) - rho0[q];
    if (rho(i, j, k) < P.buoyancy_cutoff_factor * P.base_cutoff_density) rhopert = 0.0;
    if (dm == 2) {  // mkforce.f90:310-321
      vel_force(i, j, k, 0) = -gpi(i, j, k, 0) / rho(i, j, k);
      vel_force(i, j, k, 1) = rhopert / rho(i, j, k) * grav[q] - gpi(i, j, k, 1) / rho(i, j, k) - w0_force[q];
    } else {  // mkforce.f90:398-452
      double cor[3];
      if (is_final_update) {
        cor[0] = -2.0 * omega * 0.5 * (uedge[1](i, j, k) + uedge[1](i, j + 1, k)) * cos_theta;
        cor[1] = 2.0 * omega * (0.5 * (uedge[2](i, j, k) + w0[k] + uedge[2](i, j, k + 1) + w0[k + 1]) * sin_theta +
                                0.5 * (uedge[0](i, j, k) + uedge[0](i + 1, j, k)) * cos_theta);
        cor[2] = -2.0 * omega * 0.5 * (uedge[1](i, j, k) + uedge[1](i, j + 1, k)) * sin_theta;
      } else {
        cor[0] = -2.0 * omega * uold(i, j, k, 1) * cos_theta;
        cor[1] = 2.0 * omega * ((uold(i, j, k, 2) + 0.5 * (w0[k] + w0[k + 1])) * sin_theta + uold(i, j, k, 0) * cos_theta);
        cor[2] = -2.0 * omega * uold(i, j, k, 1) * sin_theta;
      }
      vel_force(i, j, k, 0) = -cor[0] - cen[0] - gpi(i, j, k, 0) / rho(i, j, k);
      vel_force(i, j, k, 1) = -cor[1] - cen[1] - gpi(i, j, k, 1) / rho(i, j, k);
      vel_force(i, j, k, 2) = -cor[2] - cen[2] + (rhopert * grav[k] - gpi(i, j, k, 2)) / rho(i, j, k) - w0_force[k];
    }
  });
  if (do_add_utilde_force) {  // mkforce.f90:323-340 / :460-480
    const Arr& we = uedge[r];
    for_box(vb, [&](int i, int j, int k) {
      const int q = (r == 1) ? j : k;
      if (q <= -1 || q >= nr) return;
      vel_force(i, j, k, r) = vel_force(i, j, k, r) - (we.at(i, j, k, r, 1) + we(i, j, k)) * (w0[q + 1] - w0[q]) / (2.0 * dr);
    });
  }
}

// face-centred ghost fill of dm MAC components (addw0.f90:85-93, mkutrans.f90:105-115)
static void fill_faces(const mgpu_params& P, Arr* u, const int* lo, const int* hi, const int* pmask) {
  for (int d = 0; d < P.dm; ++d) fill_boundary_face(P, u[d], lo, hi, 1, d, pmask);
}

void advance_premac_box(const mgpu_params& P, const Arr& uold, const Arr& sold, Arr* umac, const Arr& gpi,
                        const double* w0, const double* w0_force, const double* rho0_old, const double* grav_cell_old,
                        const int* lo, const int* hi, int ng_u, const int* adv_bc, const int* phys_bc, const int* pmask) {
  const int dm = P.dm, r = dm - 1;
  const int ng_f = (P.ppm_trace_forces == 1) ? ng_u : 1;  // advance_premac.f90:62-66
  Box gb = grown(lo, hi, dm, ng_u), fb = grown(lo, hi, dm, ng_f);
  Arr ufull(gb.lo[0], gb.hi[0], gb.lo[1], gb.hi[1], gb.lo[2], gb.hi[2], dm);
  Arr force(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], dm);
  // put_1d_array_on_cart(w0, ufull, 1, edge-centred input, vector output): zero, radial component on the valid
  // rows over the whole x(,y) extent incl. ghost cells (fill_3d_data.f90:166-169), then ghost fill with comps 1..dm
  ufull.fill(0.0);
  Box rows = gb;
  rows.lo[r] = lo[r];
  rows.hi[r] = hi[r];
  for_box(rows, [&](int i, int j, int k) {
    const int q = (r == 1) ? j : k;
    ufull(i, j, k, r) = 0.5 * (w0[q] + w0[q + 1]);
  });
  fill_boundary_box(P, ufull, lo, hi, ng_u, 1, 1, dm, adv_bc, pmask);
  for (size_t q = 0; q < ufull.size(); ++q) ufull.p[q] = ufull.p[q] + uold.p[q];  // multifab_plus_plus_c, all ghosts

  std::vector<Arr> utrans(dm);
  for (int d = 0; d < dm; ++d) {
    Box b = grown(lo, hi, dm, 1);
    b.hi[d] += 1;
    utrans[d].alloc(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], 1);
  }
  mkutrans_box(P, uold, ufull, utrans.data(), w0, lo, hi, adv_bc, phys_bc, ng_u);  // :90
  fill_faces(P, utrans.data(), lo, hi, pmask);
  mk_vel_force_box(P, force, false, uold, utrans.data(), w0, gpi, sold.comp(P.rho_comp - 1), rho0_old, grav_cell_old,
                   w0_force, lo, hi, true);  // :98
  fill_boundary_box(P, force, lo, hi, ng_f, 1, 1, dm, adv_bc, pmask);
  addw0_box(P, utrans.data(), w0, 1.0, lo, hi);  // :109
  fill_faces(P, utrans.data(), lo, hi, pmask);
  velpred_box(P, uold, ufull, umac, utrans.data(), force, w0, lo, hi, adv_bc, phys_bc, ng_u);  // :116
}

void velocity_advance_box(const mgpu_params& P, const Arr& uold, Arr& unew, const Arr& sold, const Arr& rhohalf,
                          Arr* umac, const Arr& gpi, const double* w0, const double* w0_force, const double* rho0_old,
                          const double* rho0_nph, const double* grav_cell_old, const double* grav_cell_nph,
                          const Arr& sponge, const int* lo, const int* hi, int ng_u, const int* adv_bc,
                          const int* pmask) {
  const int dm = P.dm;
  const int ng_f = (P.ppm_trace_forces == 0) ? 1 : ng_u;  // velocity_advance.f90:69-75
  Box fb = grown(lo, hi, dm, ng_f);
  Arr force(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], dm);
  mk_vel_force_box(P, force, false, uold, umac, w0, gpi, sold.comp(P.rho_comp - 1), rho0_old, grav_cell_old, w0_force, lo,
                   hi, true);  // :80
  fill_boundary_box(P, force, lo, hi, ng_f, 1, 1, dm, adv_bc, pmask);
  addw0_box(P, umac, w0, 1.0, lo, hi);  // :90
  fill_faces(P, umac, lo, hi, pmask);
  std::vector<Arr> uedge(dm);
  for (int d = 0; d < dm; ++d) {
    Box b = grown(lo, hi, dm, 0);
    b.hi[d] += 1;
    uedge[d].alloc(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], dm);
  }
  for (int c = 0; c < dm; ++c) {  // :102-109
    if (P.bds_type == 0) make_edge_scal_box(P, uold, uedge.data(), umac, force, lo, hi, adv_bc, c, 1 + c, true, false, ng_u);
    else bds_box(P, uold, uedge.data(), umac, force, lo, hi, c, false);
  }
  addw0_box(P, umac, w0, -1.0, lo, hi);  // :115
  fill_faces(P, umac, lo, hi, pmask);
  mk_vel_force_box(P, force, true, uold, umac, w0, gpi, rhohalf.comp(0), rho0_nph, grav_cell_nph, w0_force, lo, hi,
                   true);  // :122
  fill_boundary_box(P, force, lo, hi, ng_f, 1, 1, dm, adv_bc, pmask);
  update_velocity_box(P, uold, unew, umac, uedge.data(), force, sponge, w0, lo, hi);  // :132
  fill_boundary_box(P, unew, lo, hi, ng_u, 1, 1, dm, adv_bc, pmask);  // update_vel.f90:121
}

void enthalpy_advance_box(const mgpu_params& P, int which_step, Arr& sold, Arr& snew, Arr* sedge, Arr* sflux,
                          Arr& scal_force, const Arr& thermal, Arr* umac, const double* w0, const double* rho0_old,
                          const double* rhoh0_old, const double* rho0_new, const double* rhoh0_new,
                          const double* p0_old, const double* p0_new, const double* tempbar, const double* psi,
                          const double* grav_old, const double* grav_nph, const int* lo, const int* hi, int ng_s, int ng_f,
                          const int* adv_bc, const int* pmask) {
  const int dm = P.dm, nr = P.nr, r = dm - 1;
  const int ept = P.enthalpy_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  const int rhoh = P.rhoh_comp - 1, rho = P.rho_comp - 1;
  if (ept == MGPU_PREDICT_HPRIME) fail("mk_rhoh_flux : predict_hprime not coded yet");
  const bool pred_T =
      (ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  const int temp = P.temp_comp - 1;
  if (pred_T && !tempbar) fail("enthalpy_advance: the temperature-based predictions need tempbar");
  std::vector<double> r0e_old(nr + 1), r0e_new(nr + 1), rh0e_old(nr + 1), rh0e_new(nr + 1), t0e(nr + 1);
  if (pred_T) cell_to_edge(tempbar, t0e.data(), nr);  // :118-119 (old and new are both tempbar)
  cell_to_edge(rho0_old, r0e_old.data(), nr);  // enthalpy_advance.f90:114-117
  cell_to_edge(rho0_new, r0e_new.data(), nr);
  cell_to_edge(rhoh0_old, rh0e_old.data(), nr);
  cell_to_edge(rhoh0_new, rh0e_new.data(), nr);
  Box vb = grown(lo, hi, dm, 0);
  auto rhoh_to_h = [&](bool flag) {  // convert_rhoh_to_h, convert_rhoX_to_X.f90:80
    for_box(vb, [&](int i, int j, int k) {
      if (flag) sold(i, j, k, rhoh) = sold(i, j, k, rhoh) / sold(i, j, k, rho);
      else sold(i, j, k, rhoh) = sold(i, j, k, rhoh) * sold(i, j, k, rho);
    });
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc, pmask);
  };
  if (ept == MGPU_PREDICT_H) rhoh_to_h(true);  // :122-126

  scal_force.fill(0.0);  // :132-134
  if (pred_T) {  // :190-195
    mktempforce_box(P, scal_force, sold, umac, thermal, p0_old, psi, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.temp_comp, foextrap_comp, 1, adv_bc, pmask);
  } else {
    mkrhohforce_box(P, scal_force, true, thermal, umac, p0_old, p0_old, rho0_old, rho0_old, grav_old, psi, true, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) {  // :153-156
    modify_scal_force_box(P, scal_force, sold, umac, rhoh0_old, rh0e_old.data(), w0, P.rhoh_comp, false, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask);
  } else if (ept == MGPU_PREDICT_H) {  // :173-178: multifab_div_div_c(scal_force, rhoh_comp, sold, rho_comp, 1, ng=1)
    Box g1 = grown(lo, hi, dm, 1);
    for_box(g1, [&](int i, int j, int k) { scal_force(i, j, k, rhoh) = scal_force(i, j, k, rhoh) / sold(i, j, k, rho); });
  }
  addw0_box(P, umac, w0, 1.0, lo, hi);  // :201
  fill_faces(P, umac, lo, hi, pmask);
  auto pert = [&](bool flag) {  // put_in_pert_form on the rhoh component with rhoh0_old
    for_box(vb, [&](int i, int j, int k) {
      sold(i, j, k, rhoh) = sold(i, j, k, rhoh) + (flag ? -1.0 : 1.0) * rhoh0_old[r == 1 ? j : k];
    });
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc, pmask);
  };
  auto pert_T = [&](bool flag) {  // put_in_pert_form on the temperature with tempbar (:214-217, :268-272)
    for_box(vb, [&](int i, int j, int k) {
      sold(i, j, k, temp) = sold(i, j, k, temp) + (flag ? -1.0 : 1.0) * tempbar[r == 1 ? j : k];
    });
    fill_boundary_box(P, sold, lo, hi, ng_s, P.temp_comp, flag ? foextrap_comp : dm + P.temp_comp, 1, adv_bc, pmask);
  };
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(true);  // :204-208
  if (ept == MGPU_PREDICT_TPRIME_THEN_H) pert_T(true);
  const bool cons = (ept == MGPU_PREDICT_RHOH);     // :232-254
  const int pc = pred_T ? temp : rhoh;              // :220-226
  if (P.bds_type == 0) make_edge_scal_box(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, pc, dm + pc + 1, false, cons, ng_s);
  else bds_box(P, sold, sedge, umac, scal_force, lo, hi, pc, cons);
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(false);  // :257-261
  if (ept == MGPU_PREDICT_TPRIME_THEN_H) pert_T(false);
  if (ept == MGPU_PREDICT_H) rhoh_to_h(false);     // :272-276
  if (pred_T)                                      // :280-286
    h_from_rhot_edge_box(P, sedge, rho0_old, rhoh0_old, tempbar, r0e_old.data(), rh0e_old.data(), t0e.data(), rho0_new,
                         rhoh0_new, tempbar, r0e_new.data(), rh0e_new.data(), t0e.data(), lo, hi);
  addw0_box(P, umac, w0, -1.0, lo, hi);             // :293
  fill_faces(P, umac, lo, hi, pmask);
  const bool s1 = (which_step == 1);  // :326 / :375
  mk_rhoh_flux_box(P, sflux, sedge, umac, w0, rho0_old, r0e_old.data(), s1 ? rho0_old : rho0_new,
                   s1 ? r0e_old.data() : r0e_new.data(), rhoh0_old, rh0e_old.data(), s1 ? rhoh0_old : rhoh0_new,
                   s1 ? rh0e_old.data() : rh0e_new.data(), lo, hi);
  scal_force.fill(0.0);  // :401-403
  mkrhohforce_box(P, scal_force, false, thermal, umac, p0_old, s1 ? p0_old : p0_new, rho0_old, s1 ? rho0_old : rho0_new,
                  s1 ? grav_old : grav_nph, psi, false, lo, hi);  // :405-416
  fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask);
  update_scal_box(P, P.rhoh_comp, P.rhoh_comp, sold, snew, sflux, scal_force, lo, hi, p0_new);  // :431
  fill_boundary_box(P, snew, lo, hi, ng_s, P.rhoh_comp, dm + P.rhoh_comp, 1, adv_bc, pmask);
}

}  // namespace mo
