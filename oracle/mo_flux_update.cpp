// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of the streaming kernels of the path (planar geometry):
//   mk_rhoX_flux_2d   Source/mkflux.f90:272    mk_rhoX_flux_3d_cart :370
//   mk_rhoh_flux_2d   Source/mkflux.f90:920    mk_rhoh_flux_3d_cart :1070
//   update_scal_2d    Source/update_scal.f90:246   update_scal_3d_cart :370
//   update_velocity_2d Source/update_vel.f90:174   update_velocity_3d :227 (spherical==0 branch)
//   addw0_2d/_3d      Source/addw0.f90:133,150
//   modify_scal_force_2d/_3d_cart  Source/modify_scal_force.f90:163,206
//   cell_to_edge      Source/cell_to_edge.f90:14 (single level, one chunk)
// The "radial" direction is the last one (y in 2-D, z in 3-D); base-state arrays are indexed by it.
#include "mo_kernels.h"

namespace mo {

void mk_rhoX_flux_box(const mgpu_params& P, Arr* sflux, Arr& etarhoflux, const Arr* sedge, const Arr* umac,
                      const double* w0, const double* rho0_old, const double* rho0_edge_old,
                      const double* rho0_new, const double* rho0_edge_new, const double* rho0_predicted_edge,
                      int startcomp, int endcomp, const int* lo, const int* hi) {
  const int dm = P.dm, r = dm - 1;
  const int rho = P.rho_comp - 1;
  Box vb = grown(lo, hi, dm, 0);
  for (int comp = startcomp; comp <= endcomp; ++comp) {
    const int c = comp - 1;
    for (int d = 0; d < dm; ++d) {
      Box fb = vb; fb.hi[d] = hi[d] + 1;
      const Arr& se = sedge[d];
      const Arr& um = umac[d];
      Arr& sf = sflux[d];
      for_box(fb, [&](int i, int j, int k) {
        const int ir = (r == 1) ? j : k;  // radial index (cell index for d != r, edge index for d == r)
        double rho0_edge, vel;
        if (d != r) {
          rho0_edge = 0.5 * (rho0_old[ir] + rho0_new[ir]);
          vel = um(i, j, k);
        } else {
          rho0_edge = 0.5 * (rho0_edge_old[ir] + rho0_edge_new[ir]);
          vel = um(i, j, k) + w0[ir];
        }
        if (P.species_pred_type == MGPU_PREDICT_RHOPRIME_AND_X) {
          sf(i, j, k, c) = vel * (rho0_edge + se(i, j, k, rho)) * se(i, j, k, c);
        } else if (P.species_pred_type == MGPU_PREDICT_RHOX) {
          sf(i, j, k, c) = vel * se(i, j, k, c);
        } else if (P.species_pred_type == MGPU_PREDICT_RHO_AND_X) {
          sf(i, j, k, c) = vel * se(i, j, k, rho) * se(i, j, k, c);
        }
        if (d == r && P.evolve_base_state) {
          if (comp >= P.spec_comp && comp <= P.spec_comp + P.nspec - 1)
            etarhoflux(i, j, k) = etarhoflux(i, j, k) + sf(i, j, k, c);
          if (comp == P.spec_comp + P.nspec - 1)
            etarhoflux(i, j, k) = etarhoflux(i, j, k) - w0[ir] * rho0_predicted_edge[ir];
        }
      });
    }
  }
}

void mk_rhoh_flux_box(const mgpu_params& P, Arr* sflux, const Arr* sedge, const Arr* umac, const double* w0,
                      const double* rho0_old, const double* rho0_edge_old, const double* rho0_new,
                      const double* rho0_edge_new, const double* rhoh0_old, const double* rhoh0_edge_old,
                      const double* rhoh0_new, const double* rhoh0_edge_new, const int* lo, const int* hi) {
  const int dm = P.dm, r = dm - 1;
  const int rho = P.rho_comp - 1, rhoh = P.rhoh_comp - 1;
  const int ept = P.enthalpy_pred_type;
  const bool have_h = (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  const bool have_hprime = (ept == MGPU_PREDICT_HPRIME);
  const bool have_rhoh = (ept == MGPU_PREDICT_RHOH);
  if (have_hprime) fail("mk_rhoh_flux : predict_hprime not coded yet");  // mkflux.f90:1167
  Box vb = grown(lo, hi, dm, 0);
  for (int d = 0; d < dm; ++d) {
    Box fb = vb; fb.hi[d] = hi[d] + 1;
    const Arr& se = sedge[d];
    const Arr& um = umac[d];
    Arr& sf = sflux[d];
    for_box(fb, [&](int i, int j, int k) {
      const int ir = (r == 1) ? j : k;
      double rho0_edge, rhoh0_edge, vel;
      if (d != r) {
        rho0_edge = 0.5 * (rho0_old[ir] + rho0_new[ir]);
        rhoh0_edge = 0.5 * (rhoh0_old[ir] + rhoh0_new[ir]);
        vel = um(i, j, k);
      } else {
        rho0_edge = 0.5 * (rho0_edge_old[ir] + rho0_edge_new[ir]);
        rhoh0_edge = 0.5 * (rhoh0_edge_old[ir] + rhoh0_edge_new[ir]);
        vel = um(i, j, k) + w0[ir];
      }
      if (have_h) {
        if (P.species_pred_type == MGPU_PREDICT_RHOPRIME_AND_X) {
          sf(i, j, k, rhoh) = vel * (rho0_edge + se(i, j, k, rho)) * se(i, j, k, rhoh);
        } else {
          sf(i, j, k, rhoh) = vel * se(i, j, k, rho) * se(i, j, k, rhoh);
        }
      } else if (have_rhoh) {
        sf(i, j, k, rhoh) = vel * se(i, j, k, rhoh);
      } else {
        sf(i, j, k, rhoh) = vel * (rhoh0_edge + se(i, j, k, rhoh));
      }
    });
  }
}

void update_scal_box(const mgpu_params& P, int nstart, int nstop, const Arr& sold, Arr& snew, const Arr* sflux,
                     const Arr& force, const int* lo, const int* hi, const double* p0_new, const Arr* p0_new_cart) {
  const int dm = P.dm;
  const double dt = P.dt;
  const double* dx = P.dx;
  Box vb = grown(lo, hi, dm, 0);
  for (int comp = nstart; comp <= nstop; ++comp) {
    const int c = comp - 1;
    for_box(vb, [&](int i, int j, int k) {
      double divterm = (sflux[0](i + 1, j, k, c) - sflux[0](i, j, k, c)) / dx[0] +
                       (sflux[1](i, j + 1, k, c) - sflux[1](i, j, k, c)) / dx[1];
      if (dm == 3) divterm = divterm + (sflux[2](i, j, k + 1, c) - sflux[2](i, j, k, c)) / dx[2];
      snew(i, j, k, c) = sold(i, j, k, c) + dt * (-divterm + force(i, j, k, c));
    });
  }
  if (P.do_eos_h_above_cutoff && nstart == P.rhoh_comp) {  // update_scal.f90:421-447
    if (have_eos()) {
      update_scal_eos_box(P, sold, snew, p0_new, p0_new_cart, lo, hi);
    } else {  // like the product: never skip the reset silently
      long below = 0;
      for (int k = vb.lo[2]; k <= vb.hi[2]; ++k)
        for (int j = vb.lo[1]; j <= vb.hi[1]; ++j)
          for (int i = vb.lo[0]; i <= vb.hi[0]; ++i)
            if (snew(i, j, k, P.rho_comp - 1) <= P.base_cutoff_density) ++below;
      if (below > 0) fail("update_scal: zones below base_cutoff_density need the EOS (mo_set_eos)");
    }
  }
  if (nstart == P.spec_comp && nstop == P.spec_comp + P.nspec - 1) {
    const int rho = P.rho_comp - 1;
    {  // snew(:,:,:,rho_comp) = sold(:,:,:,rho_comp), ghost cells included
      Arr a = snew.comp(rho), b = sold.comp(rho);
      for (size_t q = 0; q < a.size(); ++q) a.p[q] = b.p[q];
    }
    const double bcd = P.base_cutoff_density;
    for_box(vb, [&](int i, int j, int k) {
      bool has_negative_species = false;
      for (int comp = nstart; comp <= nstop; ++comp) {
        const int c = comp - 1;
        snew(i, j, k, rho) = snew(i, j, k, rho) + (snew(i, j, k, c) - sold(i, j, k, c));
        if (snew(i, j, k, c) < 0.0) has_negative_species = true;
      }
      if (snew(i, j, k, rho) < 0.5 * bcd) {
        for (int comp = nstart; comp <= nstop; ++comp) {
          const int c = comp - 1;
          snew(i, j, k, c) = snew(i, j, k, c) * 0.5 * bcd / snew(i, j, k, rho);
        }
        snew(i, j, k, rho) = 0.5 * bcd;
      }
      if (has_negative_species) {
        for (int comp = nstart; comp <= nstop; ++comp) {
          const int c = comp - 1;
          if (snew(i, j, k, c) < 0.0) {
            double delta = -snew(i, j, k, c);
            double sumX = 0.0;
            for (int comp2 = nstart; comp2 <= nstop; ++comp2) {
              const int c2 = comp2 - 1;
              if (comp2 != comp && snew(i, j, k, c2) >= 0.0) sumX = sumX + snew(i, j, k, c2);
            }
            for (int comp2 = nstart; comp2 <= nstop; ++comp2) {
              const int c2 = comp2 - 1;
              if (comp2 != comp && snew(i, j, k, c2) >= 0.0) {
                double frac = snew(i, j, k, c2) / sumX;
                snew(i, j, k, c2) = snew(i, j, k, c2) - frac * delta;
              }
            }
            snew(i, j, k, c) = 0.0;
          }
        }
      }
    });
  }
}

void update_velocity_box(const mgpu_params& P, const Arr& uold, Arr& unew, const Arr* umac, const Arr* uedge,
                         const Arr& force, const Arr& sponge, const double* w0, const int* lo, const int* hi) {
  const int dm = P.dm, r = dm - 1;
  const double dt = P.dt;
  const double* dx = P.dx;
  Box vb = grown(lo, hi, dm, 0);
  for_box(vb, [&](int i, int j, int k) {
    double bar[3];
    bar[0] = 0.5 * (umac[0](i, j, k) + umac[0](i + 1, j, k));
    bar[1] = 0.5 * (umac[1](i, j, k) + umac[1](i, j + 1, k));
    if (dm == 3) bar[2] = 0.5 * (umac[2](i, j, k) + umac[2](i, j, k + 1));
    for (int n = 0; n < dm; ++n) {
      double ugrad = bar[0] * (uedge[0](i + 1, j, k, n) - uedge[0](i, j, k, n)) / dx[0] +
                     bar[1] * (uedge[1](i, j + 1, k, n) - uedge[1](i, j, k, n)) / dx[1];
      if (dm == 3) ugrad = ugrad + bar[2] * (uedge[2](i, j, k + 1, n) - uedge[2](i, j, k, n)) / dx[2];
      unew(i, j, k, n) = uold(i, j, k, n) - dt * ugrad + dt * force(i, j, k, n);
    }
    const int ir = (r == 1) ? j : k;
    double wbar = 0.5 * (w0[ir] + w0[ir + 1]);
    for (int n = 0; n < dm; ++n) {
      double hi_e = (r == 1) ? uedge[1](i, j + 1, k, n) : uedge[2](i, j, k + 1, n);
      double lo_e = uedge[r](i, j, k, n);
      unew(i, j, k, n) = unew(i, j, k, n) - dt * wbar * (hi_e - lo_e) / dx[r];
      if (P.do_sponge) unew(i, j, k, n) = unew(i, j, k, n) * sponge(i, j, k);
    }
  });
}

void addw0_box(const mgpu_params& P, Arr* umac, const double* w0, double mult, const int* lo, const int* hi) {
  const int dm = P.dm, r = dm - 1;
  Box b = grown(lo, hi, dm, 1);
  b.lo[r] = lo[r];
  b.hi[r] = hi[r] + 1;
  Arr& wm = umac[r];
  for_box(b, [&](int i, int j, int k) {
    const int ir = (r == 1) ? j : k;
    wm(i, j, k) = wm(i, j, k) + mult * w0[ir];
  });
}

void modify_scal_force_box(const mgpu_params& P, Arr& force_full, const Arr& s_full, const Arr* umac,
                           const double* s0, const double* s0_edge, const double* w0, int comp, bool fullform,
                           const int* lo, const int* hi) {
  const int dm = P.dm, r = dm - 1;
  const double* dx = P.dx;
  Arr force = force_full.comp(comp - 1);
  Arr s = s_full.comp(comp - 1);
  Box vb = grown(lo, hi, dm, 0);
  for_box(vb, [&](int i, int j, int k) {
    const int ir = (r == 1) ? j : k;
    double divu, divs0u;
    if (dm == 2) {
      divu = (umac[0](i + 1, j, k) - umac[0](i, j, k)) / dx[0] + (umac[1](i, j + 1, k) - umac[1](i, j, k)) / dx[1];
      divu = divu + (w0[ir + 1] - w0[ir]) / dx[1];
      if (fullform) {
        force(i, j, k) = force(i, j, k) - s(i, j, k) * divu;
      } else {
        divs0u = s0[ir] * (umac[0](i + 1, j, k) - umac[0](i, j, k)) / dx[0] +
                 (umac[1](i, j + 1, k) * s0_edge[ir + 1] - umac[1](i, j, k) * s0_edge[ir]) / dx[1];
        force(i, j, k) = force(i, j, k) - (s(i, j, k) - s0[ir]) * divu - divs0u;
      }
    } else {
      divu = (umac[0](i + 1, j, k) - umac[0](i, j, k)) / dx[0] + (umac[1](i, j + 1, k) - umac[1](i, j, k)) / dx[1] +
             (umac[2](i, j, k + 1) - umac[2](i, j, k)) / dx[2];
      divu = divu + (w0[ir + 1] - w0[ir]) / dx[2];
      if (fullform) {
        force(i, j, k) = force(i, j, k) - s(i, j, k) * divu;
      } else {
        divs0u = s0[ir] * ((umac[0](i + 1, j, k) - umac[0](i, j, k)) / dx[0] +
                           (umac[1](i, j + 1, k) - umac[1](i, j, k)) / dx[1]) +
                 (umac[2](i, j, k + 1) * s0_edge[ir + 1] - umac[2](i, j, k) * s0_edge[ir]) / dx[2];
        force(i, j, k) = force(i, j, k) - (s(i, j, k) - s0[ir]) * divu - divs0u;
      }
    }
  });
}

// cell_to_edge.f90:14, one level with a single chunk r = 0..nr-1
void cell_to_edge(const double* s0_cell, double* s0_edge, int nr) {
  for (int r = 0; r <= nr; ++r) {
    if (r == 0) {
      s0_edge[r] = s0_cell[r];
    } else if (r == 1) {
      s0_edge[r] = 0.5 * (s0_cell[r - 1] + s0_cell[r]);
    } else if (r == nr) {
      s0_edge[r] = s0_cell[r - 1];
    } else if (r == nr - 1) {
      s0_edge[r] = 0.5 * (s0_cell[r] + s0_cell[r - 1]);
    } else {
      double tmp = 7.0 / 12.0 * (s0_cell[r] + s0_cell[r - 1]) - 1.0 / 12.0 * (s0_cell[r + 1] + s0_cell[r - 2]);
      double s0min = dmin(s0_cell[r], s0_cell[r - 1]);
      double s0max = dmax(s0_cell[r], s0_cell[r - 1]);
      s0_edge[r] = dmin(dmax(tmp, s0min), s0max);
    }
  }
}

}  // namespace mo
