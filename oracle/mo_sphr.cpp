// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Spherical-geometry branches of the path, restated loop for loop:
//   put_1d_array_on_cart_3d_sphr  Source/fill_3d_data.f90:269      quad_interp :535
//   make_w0mac_3d_sphr            Source/fill_3d_data.f90:621      make_s0mac_3d_sphr :1017
//   addw0_3d_sphr                 Source/addw0.f90:171
//   mk_rhoX_flux_3d_sphr          Source/mkflux.f90:509            mk_rhoh_flux_3d_sphr :1289
//   update_velocity_3d            Source/update_vel.f90:227 (spherical == 1 branch :317-360)
//   modify_scal_force_3d_sphr     Source/modify_scal_force.f90:256
//   pert_form_3d_sphr             Source/put_in_pert_form.f90:185
// Parity unpinned: the reference holds no golden vector for these routines; they are pinned by this
// restatement plus analytic checks in tests/test_oracle_cpu.py (linear profiles reproduced exactly, unit normal).
#include <stdexcept>
#include <string>

#include "mo_kernels.h"

namespace mo {

namespace {

struct I3 {
  int i, j, k;
};
inline I3 sh(int i, int j, int k, int d, int o) { return I3{i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0)}; }

inline double max3(double a, double b, double c) { return dmax(dmax(a, b), c); }
inline double min3(double a, double b, double c) { return dmin(dmin(a, b), c); }

// fill_3d_data.f90:535-546
inline double quad_interp(double x, double x0, double x1, double x2, double y0, double y1, double y2) {
  double y = y0 + (y1 - y0) / (x1 - x0) * (x - x0) +
             ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0)) / (x2 - x0) * (x - x0) * (x - x1);
  if (y > max3(y0, y1, y2)) y = max3(y0, y1, y2);
  if (y < min3(y0, y1, y2)) y = min3(y0, y1, y2);
  return y;
}

// value of an EDGE-centred 1-D array at `radius` (w0_interp_type / w0mac_interp_type semantics 1..3)
inline double interp_edge(const mgpu_geom& g, int type, const double* s0, double radius) {
  const double dr = g.dr;
  int index = (int)(radius / dr);
  if (type == 1) {  // :308-316
    const double rfac = (radius - (double)index * dr) / dr;
    return (rfac > 0.5) ? s0[index + 1] : s0[index];
  }
  if (type == 2) {  // :341-349
    const double rfac = (radius - (double)index * dr) / dr;
    if (index < g.nr_fine) return rfac * s0[index + 1] + (1.0 - rfac) * s0[index];
    return s0[g.nr_fine];
  }
  // :374-389.  QUIRK: the third test compares a distance with a position, so it is true whenever it is reached
  if (index <= 0) index = 0;
  else if (index >= g.nr_fine - 1) index = g.nr_fine - 2;
  else if (radius - g.r_edge_loc[index] < g.r_edge_loc[index + 1]) index = index - 1;
  return quad_interp(radius, g.r_edge_loc[index], g.r_edge_loc[index + 1], g.r_edge_loc[index + 2], s0[index],
                     s0[index + 1], s0[index + 2]);
}

// value of a BIN-centred 1-D array at `radius` (s0_interp_type / s0mac_interp_type semantics 1..3)
inline double interp_cc(const mgpu_geom& g, int type, const double* s0, double radius) {
  const double dr = g.dr;
  int index = (int)(radius / dr);
  const int nr = g.nr_fine;
  if (type == 1) return s0[index];  // :426-428
  if (type == 2) {                  // :453-471
    if (radius >= g.r_cc_loc[index]) {
      if (index >= nr - 1) return s0[nr - 1];
      return s0[index + 1] * (radius - g.r_cc_loc[index]) / dr + s0[index] * (g.r_cc_loc[index + 1] - radius) / dr;
    }
    if (index == 0) return s0[index];
    if (index > nr - 1) return s0[nr - 1];
    return s0[index] * (radius - g.r_cc_loc[index - 1]) / dr + s0[index - 1] * (g.r_cc_loc[index] - radius) / dr;
  }
  // :496-510
  if (index == 0) index = 1;
  else if (index >= nr - 1) index = nr - 2;
  return quad_interp(radius, g.r_cc_loc[index - 1], g.r_cc_loc[index], g.r_cc_loc[index + 1], s0[index - 1], s0[index],
                     s0[index + 1]);
}

// position of index i in direction d: cell centre (half = true) or face/node (half = false)
inline double pos(const mgpu_geom& g, const mgpu_params& P, int d, int i, bool half) {
  return g.prob_lo[d] + ((double)i + (half ? 0.5 : 0.0)) * P.dx[d] - g.center[d];
}
inline double radius_of(double x, double y, double z) { return std::sqrt(x * x + y * y + z * z); }

void check_types(const mgpu_geom& g) {
  if (g.nr_fine < 3) fail("spherical geometry: nr_fine must be at least 3");
}

}  // namespace

void put_1d_array_on_cart_sphr(const mgpu_params& P, const mgpu_geom& g, bool edge_in, bool vec, const double* s0,
                               Arr& cart, const int* lo, const int* hi) {
  check_types(g);
  const int type = edge_in ? g.w0_interp_type : g.s0_interp_type;
  if (type < 1 || type > 3) fail(edge_in ? "Error: w0_interp_type not defined" : "Error: s0_interp_type not defined");
#pragma omp parallel for
  for (int k = lo[2]; k <= hi[2]; ++k) {
    const double z = pos(g, P, 2, k, true);
    for (int j = lo[1]; j <= hi[1]; ++j) {
      const double y = pos(g, P, 1, j, true);
      for (int i = lo[0]; i <= hi[0]; ++i) {
        const double x = pos(g, P, 0, i, true);
        const double radius = radius_of(x, y, z);
        const double v = edge_in ? interp_edge(g, type, s0, radius) : interp_cc(g, type, s0, radius);
        if (vec) {
          cart(i, j, k, 0) = v * x * (1.0 / radius);
          cart(i, j, k, 1) = v * y * (1.0 / radius);
          cart(i, j, k, 2) = v * z * (1.0 / radius);
        } else {
          cart(i, j, k, 0) = v;
        }
      }
    }
  }
}

// loop bounds of the face arrays: lo-1:hi+1 transverse, lo-1:hi+2 normal (fill_3d_data.f90:651-677)
static Box mac_box(const int* lo, const int* hi, int d) {
  Box b;
  for (int q = 0; q < 3; ++q) {
    b.lo[q] = lo[q] - 1;
    b.hi[q] = hi[q] + 1 + (q == d ? 1 : 0);
  }
  return b;
}

void make_w0mac_sphr(const mgpu_params& P, const mgpu_geom& g, const double* w0, Arr* w0mac, const Arr* w0_cart,
                     const int* lo, const int* hi) {
  check_types(g);
  const int type = g.w0mac_interp_type;
  if (type < 1 || type > 4) fail("Error: w0mac_interp_type not defined");
  if (type == 1 && !w0_cart) fail("make_w0mac: w0mac_interp_type = 1 needs w0_cart");
  for (int d = 0; d < 3; ++d) {
    Arr& m = w0mac[d];
    for_box(mac_box(lo, hi, d), [&](int i, int j, int k) {
      if (type == 1) {  // :651-677
        I3 l = sh(i, j, k, d, -1);
        m(i, j, k) = 0.5 * ((*w0_cart)(l.i, l.j, l.k, d) + (*w0_cart)(i, j, k, d));
      } else if (type == 2 || type == 3) {  // :684-759, :765-850
        const double x = pos(g, P, 0, i, d != 0), y = pos(g, P, 1, j, d != 1), z = pos(g, P, 2, k, d != 2);
        const double radius = radius_of(x, y, z);
        const double v = interp_edge(g, type, w0, radius);
        const double c = (d == 0) ? x : (d == 1 ? y : z);
        m(i, j, k) = v * c / radius;
      } else {  // :852-935: nodal values averaged over the four nodes of the face
        auto nodal = [&](int ii, int jj, int kk) {
          const double x = pos(g, P, 0, ii, false), y = pos(g, P, 1, jj, false), z = pos(g, P, 2, kk, false);
          const double radius = radius_of(x, y, z);
          const double v = interp_edge(g, 2, w0, radius);
          const double c = (d == 0) ? x : (d == 1 ? y : z);
          return v * c * (1.0 / radius);
        };
        const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;  // the two transverse directions, in the source's order
        I3 a = {i, j, k}, b = sh(i, j, k, t1, 1), c2 = sh(i, j, k, t2, 1);
        I3 e = sh(b.i, b.j, b.k, t2, 1);
        m(i, j, k) = 0.25 * (nodal(a.i, a.j, a.k) + nodal(b.i, b.j, b.k) + nodal(c2.i, c2.j, c2.k) + nodal(e.i, e.j, e.k));
      }
    });
  }
}

void make_s0mac_sphr(const mgpu_params& P, const mgpu_geom& g, const double* s0, Arr* s0mac, const Arr* s0_cart,
                     const int* lo, const int* hi) {
  check_types(g);
  const int type = g.s0mac_interp_type;
  if (type < 1 || type > 3) fail("Error: s0mac_interp_type not defined");
  if (type == 1 && !s0_cart) fail("make_s0mac: s0mac_interp_type = 1 needs s0_cart");
  for (int d = 0; d < 3; ++d) {
    Arr& m = s0mac[d];
    for_box(mac_box(lo, hi, d), [&](int i, int j, int k) {
      if (type == 1) {  // :1046-1072
        I3 l = sh(i, j, k, d, -1);
        m(i, j, k) = 0.5 * ((*s0_cart)(i, j, k) + (*s0_cart)(l.i, l.j, l.k));
      } else {  // :1077-1275
        const double x = pos(g, P, 0, i, d != 0), y = pos(g, P, 1, j, d != 1), z = pos(g, P, 2, k, d != 2);
        m(i, j, k) = interp_cc(g, type, s0, radius_of(x, y, z));
      }
    });
  }
}

void addw0_sphr(Arr* umac, const Arr* w0mac, const int* lo, const int* hi, double mult) {
  for (int d = 0; d < 3; ++d) {
    Box b = grown(lo, hi, 3, 0);
    b.hi[d] += 1;
    for_box(b, [&](int i, int j, int k) { umac[d](i, j, k) = umac[d](i, j, k) + mult * w0mac[d](i, j, k); });
  }
}

void mk_rhoX_flux_sphr(const mgpu_params& P, Arr* sflux, const Arr* sedge, const Arr* umac, const Arr* w0mac,
                       const Arr* r0o, const Arr* r0n, int startcomp, int endcomp, const int* lo, const int* hi) {
  const int rho = P.rho_comp - 1;
  for (int comp = startcomp - 1; comp <= endcomp - 1; ++comp)
    for (int d = 0; d < 3; ++d) {
      Box b = grown(lo, hi, 3, 0);
      b.hi[d] += 1;
      for_box(b, [&](int i, int j, int k) {
        const double vel = umac[d](i, j, k) + w0mac[d](i, j, k);
        if (P.species_pred_type == MGPU_PREDICT_RHOPRIME_AND_X) {
          const double rho0_edge = 0.5 * (r0o[d](i, j, k) + r0n[d](i, j, k));
          sflux[d](i, j, k, comp) = vel * (rho0_edge + sedge[d](i, j, k, rho)) * sedge[d](i, j, k, comp);
        } else if (P.species_pred_type == MGPU_PREDICT_RHOX) {
          sflux[d](i, j, k, comp) = vel * sedge[d](i, j, k, comp);
        } else if (P.species_pred_type == MGPU_PREDICT_RHO_AND_X) {
          sflux[d](i, j, k, comp) = vel * sedge[d](i, j, k, rho) * sedge[d](i, j, k, comp);
        }
      });
    }
}

void mk_rhoh_flux_sphr(const mgpu_params& P, Arr* sflux, const Arr* sedge, const Arr* umac, const Arr* w0mac,
                       const Arr* r0o, const Arr* r0n, const Arr* h0o, const Arr* h0n, const int* lo, const int* hi) {
  const int rho = P.rho_comp - 1, rhoh = P.rhoh_comp - 1;
  const int ept = P.enthalpy_pred_type, spt = P.species_pred_type;
  const bool have_h = (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  const bool have_hprime = (ept == MGPU_PREDICT_HPRIME), have_rhoh = (ept == MGPU_PREDICT_RHOH);
  if (have_hprime && spt == MGPU_PREDICT_RHO_AND_X)
    fail("ERROR: predict_rho_and_X and predict_hprime not supported together");  // mkflux.f90:1408
  if (have_hprime && spt == MGPU_PREDICT_RHOX) fail("ERROR: predict_rhoX and predict_hprime not supported together");
  for (int d = 0; d < 3; ++d) {
    Box b = grown(lo, hi, 3, 0);
    b.hi[d] += 1;
    for_box(b, [&](int i, int j, int k) {
      const double vel = umac[d](i, j, k) + w0mac[d](i, j, k);
      const double erho = sedge[d](i, j, k, rho), erhoh = sedge[d](i, j, k, rhoh);
      double f;
      if (have_h) {
        if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
          const double rho0_edge = 0.5 * (r0o[d](i, j, k) + r0n[d](i, j, k));
          f = vel * (rho0_edge + erho) * erhoh;
        } else {
          f = vel * erho * erhoh;
        }
      } else if (have_hprime) {
        const double rho0_edge = 0.5 * (r0o[d](i, j, k) + r0n[d](i, j, k));
        const double h0_edge = 0.5 * (h0o[d](i, j, k) + h0n[d](i, j, k));
        f = vel * (erho + rho0_edge) * (erhoh + h0_edge);
      } else if (have_rhoh) {
        f = vel * erhoh;
      } else {
        const double rho0_edge = 0.5 * (r0o[d](i, j, k) + r0n[d](i, j, k));
        const double h0_edge = 0.5 * (h0o[d](i, j, k) + h0n[d](i, j, k));
        f = vel * (rho0_edge * h0_edge + erhoh);
      }
      sflux[d](i, j, k, rhoh) = f;
    });
  }
}

void update_velocity_sphr(const mgpu_params& P, const Arr& uold, Arr& unew, const Arr* umac, const Arr* uedge,
                          const Arr& force, const Arr& sponge, const Arr* w0mac, const int* lo, const int* hi) {
  const double dt = P.dt;
  const double* dx = P.dx;
  const Arr &ux = uedge[0], &uy = uedge[1], &uz = uedge[2];
  for_box(grown(lo, hi, 3, 0), [&](int i, int j, int k) {
    const double ubar = 0.5 * (umac[0](i, j, k) + umac[0](i + 1, j, k));
    const double vbar = 0.5 * (umac[1](i, j, k) + umac[1](i, j + 1, k));
    const double wbar = 0.5 * (umac[2](i, j, k) + umac[2](i, j, k + 1));
    for (int c = 0; c < 3; ++c) {
      const double ugrad = ubar * (ux(i + 1, j, k, c) - ux(i, j, k, c)) / dx[0] +
                           vbar * (uy(i, j + 1, k, c) - uy(i, j, k, c)) / dx[1] +
                           wbar * (uz(i, j, k + 1, c) - uz(i, j, k, c)) / dx[2];
      unew(i, j, k, c) = uold(i, j, k, c) - dt * ugrad + dt * force(i, j, k, c);
    }
  });
  for_box(grown(lo, hi, 3, 0), [&](int i, int j, int k) {
    const double w0x = 0.5 * (w0mac[0](i, j, k) + w0mac[0](i + 1, j, k));
    const double w0y = 0.5 * (w0mac[1](i, j, k) + w0mac[1](i, j + 1, k));
    const double w0z = 0.5 * (w0mac[2](i, j, k) + w0mac[2](i, j, k + 1));
    for (int c = 0; c < 3; ++c) {
      const double gx = (ux(i + 1, j, k, c) - ux(i, j, k, c)) / dx[0];
      const double gy = (uy(i, j + 1, k, c) - uy(i, j, k, c)) / dx[1];
      const double gz = (uz(i, j, k + 1, c) - uz(i, j, k, c)) / dx[2];
      const double w0_grad = gx * w0x + gy * w0y + gz * w0z;
      unew(i, j, k, c) = unew(i, j, k, c) - dt * w0_grad;
    }
    if (P.do_sponge)
      for (int c = 0; c < 3; ++c) unew(i, j, k, c) = unew(i, j, k, c) * sponge(i, j, k);
  });
}

void modify_scal_force_sphr(const mgpu_params& P, const mgpu_geom& g, Arr& force, const Arr& s, const Arr* umac,
                            const Arr& s0c, const double* w0, bool fullform, const int* lo, const int* hi) {
  check_types(g);
  const int nr = g.nr_fine;
  std::vector<double> divu(nr);
  for (int r = 0; r < nr; ++r)  // modify_scal_force.f90:293-297
    divu[r] = (g.r_edge_loc[r + 1] * g.r_edge_loc[r + 1] * w0[r + 1] - g.r_edge_loc[r] * g.r_edge_loc[r] * w0[r]) /
              (g.dr * (g.r_cc_loc[r] * g.r_cc_loc[r]));
  Arr divu_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
  put_1d_array_on_cart_sphr(P, g, false, false, divu.data(), divu_cart, lo, hi);
  const double* dx = P.dx;
  const Arr &um = umac[0], &vm = umac[1], &wm = umac[2];
  for_box(grown(lo, hi, 3, 0), [&](int i, int j, int k) {
    const double divumac = (um(i + 1, j, k) - um(i, j, k)) / dx[0] + (vm(i, j + 1, k) - vm(i, j, k)) / dx[1] +
                           (wm(i, j, k + 1) - wm(i, j, k)) / dx[2];
    if (fullform) {
      force(i, j, k) = force(i, j, k) - s(i, j, k) * (divumac + divu_cart(i, j, k));
    } else {
      const double c = s0c(i, j, k);
      const double s0_xhi = (i < P.domhi[0]) ? 0.5 * (c + s0c(i + 1, j, k)) : c;
      const double s0_xlo = (i > P.domlo[0]) ? 0.5 * (c + s0c(i - 1, j, k)) : c;
      const double s0_yhi = (j < P.domhi[1]) ? 0.5 * (c + s0c(i, j + 1, k)) : c;
      const double s0_ylo = (j > P.domlo[1]) ? 0.5 * (c + s0c(i, j - 1, k)) : c;
      const double s0_zhi = (k < P.domhi[2]) ? 0.5 * (c + s0c(i, j, k + 1)) : c;
      const double s0_zlo = (k > P.domlo[2]) ? 0.5 * (c + s0c(i, j, k - 1)) : c;
      const double divs0u = (um(i + 1, j, k) * s0_xhi - um(i, j, k) * s0_xlo) / dx[0] +
                            (vm(i, j + 1, k) * s0_yhi - vm(i, j, k) * s0_ylo) / dx[1] +
                            (wm(i, j, k + 1) * s0_zhi - wm(i, j, k) * s0_zlo) / dx[2];
      force(i, j, k) = force(i, j, k) - divs0u - (s(i, j, k) - c) * (divumac + divu_cart(i, j, k));
    }
  });
}

void pert_form_sphr(const mgpu_params& P, const mgpu_geom& g, Arr& s, const double* s0, bool flag, const int* lo,
                    const int* hi) {
  Arr s0_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
  put_1d_array_on_cart_sphr(P, g, false, false, s0, s0_cart, lo, hi);
  const int mult = flag ? -1 : 1;
  for_box(grown(lo, hi, 3, 0), [&](int i, int j, int k) { s(i, j, k) = s(i, j, k) + mult * s0_cart(i, j, k); });
}


// density_advance (Source/density_advance.f90:20) with spherical == 1: rho0_old_cart for the force (:105-111), addw0
// with w0mac (:148), spherical perturbational form (:166,229), rho0mac_old/new by make_s0mac (:268-311), no etarhoflux.
void density_advance_sphr_box(const mgpu_params& P, const mgpu_geom& g, int which_step, Arr& sold, Arr& snew, Arr* sedge,
                              Arr* sflux, Arr& scal_force, Arr* umac, const double* w0, const Arr* w0mac,
                              const double* rho0_old, const double* rho0_new, const int* lo, const int* hi, int ng_s,
                              int ng_f, const int* adv_bc, const int* pmask) {
  const int dm = 3, spt = P.species_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  auto cart_of = [&](const double* s0, int ng) {  // put_1d_array_on_cart (fill_3d_data.f90:21) incl. its ghost fill
    Arr c(lo[0] - ng, hi[0] + ng, lo[1] - ng, hi[1] + ng, lo[2] - ng, hi[2] + ng, 1);
    put_1d_array_on_cart_sphr(P, g, false, false, s0, c, lo, hi);
    // component 1 of a 1-component array filled with the BCs of the density (bc_comp = dm+rho_comp)
    fill_boundary_box(P, c, lo, hi, ng, 1, dm + P.rho_comp, 1, adv_bc, pmask);
    return c;
  };
  auto fill_umac = [&]() {
    for (int d = 0; d < dm; ++d) fill_boundary_face(P, umac[d], lo, hi, 1, d, pmask);
  };
  scal_force.fill(0.0);
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {
    Arr rho0_old_cart = cart_of(rho0_old, 1);
    Arr fo = scal_force.comp(P.rho_comp - 1), sa = sold.comp(P.rho_comp - 1);
    modify_scal_force_sphr(P, g, fo, sa, umac, rho0_old_cart, w0, spt == MGPU_PREDICT_RHO_AND_X, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rho_comp, foextrap_comp, 1, adv_bc, pmask);
  }
  addw0_sphr(umac, w0mac, lo, hi, 1.0);
  fill_umac();
  Box vb = grown(lo, hi, dm, 0);
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {
    for (int n = 0; n < P.nspec; ++n) {
      const int c = P.spec_comp - 1 + n;
      for_box(vb, [&](int i, int j, int k) { sold(i, j, k, c) = sold(i, j, k, c) / sold(i, j, k, P.rho_comp - 1); });
      fill_boundary_box(P, sold, lo, hi, ng_s, c + 1, foextrap_comp, 1, adv_bc, pmask);
    }
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
    Arr sa = sold.comp(P.rho_comp - 1);
    pert_form_sphr(P, g, sa, rho0_old, true, lo, hi);
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rho_comp, foextrap_comp, 1, adv_bc, pmask);
  }
  auto edge = [&](int scomp, int ncomp, bool cons) {
    for (int n = 0; n < ncomp; ++n) {
      if (P.bds_type == 0)
        make_edge_scal_box(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, scomp - 1 + n, dm + scomp + n, false, cons, ng_s);
      else
        bds_box(P, sold, sedge, umac, scal_force, lo, hi, scomp - 1 + n, cons);
    }
  };
  edge(P.spec_comp, P.nspec, spt == MGPU_PREDICT_RHOX);
  if (spt == MGPU_PREDICT_RHOX) {
    for (int d = 0; d < dm; ++d) {
      Arr r = sedge[d].comp(P.rho_comp - 1), s1 = sedge[d].comp(P.spec_comp - 1);
      for (size_t q = 0; q < r.size(); ++q) r.p[q] = s1.p[q];
      for (int n = 1; n < P.nspec; ++n) {
        Arr sn = sedge[d].comp(P.spec_comp - 1 + n);
        for (size_t q = 0; q < r.size(); ++q) r.p[q] = r.p[q] + sn.p[q];
      }
    }
  } else {
    edge(P.rho_comp, 1, false);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
    Arr sa = sold.comp(P.rho_comp - 1);
    pert_form_sphr(P, g, sa, rho0_old, false, lo, hi);
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {
    for (int n = 0; n < P.nspec; ++n) {
      const int c = P.spec_comp - 1 + n;
      for_box(vb, [&](int i, int j, int k) { sold(i, j, k, c) = sold(i, j, k, c) * sold(i, j, k, P.rho_comp - 1); });
    }
    fill_boundary_box(P, sold, lo, hi, ng_s, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask);
  }
  if (P.ntrac >= 1) edge(P.trac_comp, P.ntrac, false);
  addw0_sphr(umac, w0mac, lo, hi, -1.0);
  fill_umac();
  // rho0mac_old / rho0mac_new (make_s0mac, fill_3d_data.f90:942): via cell centres (ng = 2) when s0mac_interp_type = 1
  auto mac_of = [&](const double* s0, Arr* mac) {
    for (int d = 0; d < 3; ++d) {
      Box b;
      for (int q = 0; q < 3; ++q) { b.lo[q] = lo[q] - 1; b.hi[q] = hi[q] + 1 + (q == d ? 1 : 0); }
      mac[d].alloc(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], 1);
    }
    if (g.s0mac_interp_type == 1) {
      Arr c = cart_of(s0, 2);
      make_s0mac_sphr(P, g, s0, mac, &c, lo, hi);
    } else {
      make_s0mac_sphr(P, g, s0, mac, nullptr, lo, hi);
    }
  };
  Arr r0o[3], r0n[3];
  mac_of(rho0_old, r0o);
  if (which_step == 2) mac_of(rho0_new, r0n);
  const Arr* rn = (which_step == 1) ? r0o : r0n;
  mk_rhoX_flux_sphr(P, sflux, sedge, umac, w0mac, r0o, rn, P.spec_comp, P.spec_comp + P.nspec - 1, lo, hi);
  if (P.ntrac >= 1) mk_rhoX_flux_sphr(P, sflux, sedge, umac, w0mac, r0o, rn, P.trac_comp, P.trac_comp + P.ntrac - 1, lo, hi);
  scal_force.fill(0.0);
  update_scal_box(P, P.spec_comp, P.spec_comp + P.nspec - 1, sold, snew, sflux, scal_force, lo, hi);
  fill_boundary_box(P, snew, lo, hi, ng_s, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask);
  fill_boundary_box(P, snew, lo, hi, ng_s, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask);
  if (P.ntrac >= 1) {
    update_scal_box(P, P.trac_comp, P.trac_comp + P.ntrac - 1, sold, snew, sflux, scal_force, lo, hi);
    fill_boundary_box(P, snew, lo, hi, ng_s, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask);
  }
}

}  // namespace mo

// ---------------------------------------------------------------------------------------------
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

// make_normal_3d_sphr (fill_3d_data.f90:1308): unit radial vector at the cell centres, ghost cells included
void make_normal_sphr(const mgpu_params& P, const mgpu_geom& g, Arr& normal, const int* lo, const int* hi, int ng) {
  for_box(grown(lo, hi, 3, ng), [&](int i, int j, int k) {
    const double x = pos(g, P, 0, i, true), y = pos(g, P, 1, j, true), z = pos(g, P, 2, k, true);
    const double radius = std::sqrt(x * x + y * y + z * z);
    normal(i, j, k, 0) = x * (1.0 / radius);
    normal(i, j, k, 1) = y * (1.0 / radius);
    normal(i, j, k, 2) = z * (1.0 / radius);
  });
}

// mk_vel_force_3d_sphr (mkforce.f90:484) with the spherical preparation of its wrapper (mkforce.f90:92-127: w0 on the
// cell centres as a vector, the radial gradient of w0 on the cell centres).  rho is a single-component view.
void mk_vel_force_sphr_box(const mgpu_params& P, const mgpu_geom& g, Arr& vel_force, bool is_final_update, const Arr& uold,
                           const Arr* uedge, const double* w0, const Arr* w0mac, const Arr& gpi, const Arr& rho,
                           const Arr& normal, const double* rho0, const double* grav, const Arr& w0_force_cart,
                           const int* lo, const int* hi, bool do_add_utilde_force) {
  const int nr = g.nr_fine;
  Box vb = grown(lo, hi, 3, 0);
  auto cart = [&](int nc) { return Arr(vb.lo[0], vb.hi[0], vb.lo[1], vb.hi[1], vb.lo[2], vb.hi[2], nc); };
  Arr w0_cart = cart(3), gradw0_cart = cart(1), rho0_cart = cart(1), grav_cart = cart(3);
  w0_cart.fill(0.0);
  gradw0_cart.fill(0.0);
  if (P.evolve_base_state) {  // :109-126
    put_1d_array_on_cart_sphr(P, g, true, true, w0, w0_cart, lo, hi);
    if (do_add_utilde_force) {
      std::vector<double> gradw0_rad(nr);
      for (int r = 0; r < nr; ++r) gradw0_rad[r] = (w0[r + 1] - w0[r]) / g.dr;
      put_1d_array_on_cart_sphr(P, g, false, false, gradw0_rad.data(), gradw0_cart, lo, hi);
    }
  }
  vel_force.fill(0.0);  // :527
  put_1d_array_on_cart_sphr(P, g, false, false, rho0, rho0_cart, lo, hi);
  put_1d_array_on_cart_sphr(P, g, false, true, grav, grav_cart, lo, hi);
  const double omega = P.omega;
  const Arr &ue = uedge[0], &ve = uedge[1], &we = uedge[2];
  for_box(vb, [&](int i, int j, int k) {
    const double zz = g.prob_lo[2] + ((double)k + 0.5) * P.dx[2] - g.center[2];
    const double yy = g.prob_lo[1] + ((double)j + 0.5) * P.dx[1] - g.center[1];
    const double xx = g.prob_lo[0] + ((double)i + 0.5) * P.dx[0] - g.center[0];
    (void)zz;
    double rhopert = rho(i, j, k) - rho0_cart(i, j, k, 0);
    const bool outside = rho(i, j, k) < P.buoyancy_cutoff_factor * P.base_cutoff_density;
    if (outside) rhopert = 0.0;
    double cen[3] = {-omega * omega * xx, -omega * omega * yy, 0.0};
    if (outside) cen[0] = cen[1] = cen[2] = 0.0;
    double cor[3];
    if (is_final_update) {
      cor[0] = -(2.0 * omega * 0.5 * (ve(i, j, k) + w0mac[1](i, j, k) + ve(i, j + 1, k) + w0mac[1](i, j + 1, k)));
      cor[1] = 2.0 * omega * 0.5 * (ue(i, j, k) + w0mac[0](i, j, k) + ue(i + 1, j, k) + w0mac[0](i + 1, j, k));
      cor[2] = 0.0;
    } else {
      cor[0] = -(2.0 * omega * (uold(i, j, k, 1) + w0_cart(i, j, k, 1)));
      cor[1] = 2.0 * omega * (uold(i, j, k, 0) + w0_cart(i, j, k, 0));
      cor[2] = 0.0;
    }
    for (int c = 0; c < 3; ++c)
      vel_force(i, j, k, c) = -cor[c] - cen[c] + (rhopert * grav_cart(i, j, k, c) - gpi(i, j, k, c)) / rho(i, j, k) -
                              w0_force_cart(i, j, k, c);
  });
  if (do_add_utilde_force) {  // :617-640
    for_box(vb, [&](int i, int j, int k) {
      const double Ut_dot_er = 0.5 * (ue(i, j, k) + ue(i + 1, j, k)) * normal(i, j, k, 0) +
                               0.5 * (ve(i, j, k) + ve(i, j + 1, k)) * normal(i, j, k, 1) +
                               0.5 * (we(i, j, k) + we(i, j, k + 1)) * normal(i, j, k, 2);
      for (int c = 0; c < 3; ++c)
        vel_force(i, j, k, c) = vel_force(i, j, k, c) - Ut_dot_er * gradw0_cart(i, j, k, 0) * normal(i, j, k, c);
    });
  }
}

static void fill_faces3(const mgpu_params& P, Arr* u, const int* lo, const int* hi, const int* pmask) {
  for (int d = 0; d < 3; ++d) fill_boundary_face(P, u[d], lo, hi, 1, d, pmask);
}

// advance_premac (advance_premac.f90:21) with spherical == 1
void advance_premac_sphr_box(const mgpu_params& P, const mgpu_geom& g, const Arr& uold, const Arr& sold, Arr* umac,
                             const Arr& gpi, const Arr& normal, const double* w0, const Arr* w0mac,
                             const Arr& w0_force_cart, const double* rho0_old, const double* grav_cell_old, const int* lo,
                             const int* hi, int ng_u, const int* adv_bc, const int* phys_bc, const int* pmask) {
  const int dm = 3;
  const int ng_f = (P.ppm_trace_forces == 1) ? ng_u : 1;  // :62-66
  Box gb = grown(lo, hi, dm, ng_u), fb = grown(lo, hi, dm, ng_f);
  Arr ufull(gb.lo[0], gb.hi[0], gb.lo[1], gb.hi[1], gb.lo[2], gb.hi[2], dm);
  Arr force(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], dm);
  ufull.fill(0.0);
  put_1d_array_on_cart_sphr(P, g, true, true, w0, ufull, lo, hi);  // :75, then its ghost fill (fill_3d_data.f90:110-128)
  fill_boundary_box(P, ufull, lo, hi, ng_u, 1, 1, dm, adv_bc, pmask);
  for (size_t q = 0; q < ufull.size(); ++q) ufull.p[q] = ufull.p[q] + uold.p[q];  // :76-78
  std::vector<Arr> utrans(dm);
  for (int d = 0; d < dm; ++d) {
    Box b = grown(lo, hi, dm, 1);
    b.hi[d] += 1;
    utrans[d].alloc(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], 1);
  }
  mkutrans_box(P, uold, ufull, utrans.data(), nullptr, lo, hi, adv_bc, phys_bc, ng_u, w0mac);  // :90
  fill_faces3(P, utrans.data(), lo, hi, pmask);
  mk_vel_force_sphr_box(P, g, force, false, uold, utrans.data(), w0, w0mac, gpi, sold.comp(P.rho_comp - 1), normal,
                        rho0_old, grav_cell_old, w0_force_cart, lo, hi, true);  // :98
  fill_boundary_box(P, force, lo, hi, ng_f, 1, 1, dm, adv_bc, pmask);
  addw0_sphr(utrans.data(), w0mac, lo, hi, 1.0);  // :109
  fill_faces3(P, utrans.data(), lo, hi, pmask);
  velpred_box(P, uold, ufull, umac, utrans.data(), force, nullptr, lo, hi, adv_bc, phys_bc, ng_u, w0mac);  // :116
}

// velocity_advance (velocity_advance.f90:16) with spherical == 1
void velocity_advance_sphr_box(const mgpu_params& P, const mgpu_geom& g, const Arr& uold, Arr& unew, const Arr& sold,
                               const Arr& rhohalf, Arr* umac, const Arr& gpi, const Arr& normal, const double* w0,
                               const Arr* w0mac, const Arr& w0_force_cart, const double* rho0_old, const double* rho0_nph,
                               const double* grav_cell_old, const double* grav_cell_nph, const Arr& sponge, const int* lo,
                               const int* hi, int ng_u, const int* adv_bc, const int* pmask) {
  const int dm = 3;
  const int ng_f = (P.ppm_trace_forces == 0) ? 1 : ng_u;  // :69-75
  Box fb = grown(lo, hi, dm, ng_f);
  Arr force(fb.lo[0], fb.hi[0], fb.lo[1], fb.hi[1], fb.lo[2], fb.hi[2], dm);
  mk_vel_force_sphr_box(P, g, force, false, uold, umac, w0, w0mac, gpi, sold.comp(P.rho_comp - 1), normal, rho0_old,
                        grav_cell_old, w0_force_cart, lo, hi, true);  // :80
  fill_boundary_box(P, force, lo, hi, ng_f, 1, 1, dm, adv_bc, pmask);
  addw0_sphr(umac, w0mac, lo, hi, 1.0);  // :90
  fill_faces3(P, umac, lo, hi, pmask);
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
  addw0_sphr(umac, w0mac, lo, hi, -1.0);  // :115
  fill_faces3(P, umac, lo, hi, pmask);
  mk_vel_force_sphr_box(P, g, force, true, uold, umac, w0, w0mac, gpi, rhohalf.comp(0), normal, rho0_nph, grav_cell_nph,
                        w0_force_cart, lo, hi, true);  // :122
  fill_boundary_box(P, force, lo, hi, ng_f, 1, 1, dm, adv_bc, pmask);
  update_velocity_sphr(P, uold, unew, umac, uedge.data(), force, sponge, w0mac, lo, hi);  // :132
  fill_boundary_box(P, unew, lo, hi, ng_u, 1, 1, dm, adv_bc, pmask);  // update_vel.f90:121
}

// ---- enthalpy path, spherical ------------------------------------------------------------------------------------
// put_1d_array_on_cart of a bin-centred array incl. its ghost fill with the BCs of component bccomp (fill_3d_data.f90:21)
Arr cart_with_ghosts(const mgpu_params& P, const mgpu_geom& g, const double* s0, int ng, int bccomp, const int* lo,
                            const int* hi, const int* adv_bc, const int* pmask) {
  Arr c(lo[0] - ng, hi[0] + ng, lo[1] - ng, hi[1] + ng, lo[2] - ng, hi[2] + ng, 1);
  put_1d_array_on_cart_sphr(P, g, false, false, s0, c, lo, hi);
  fill_boundary_box(P, c, lo, hi, ng, 1, bccomp, 1, adv_bc, pmask);
  return c;
}
// make_s0mac (fill_3d_data.f90:942): through the cell centres (two ghost cells) when s0mac_interp_type = 1
static void s0mac_of(const mgpu_params& P, const mgpu_geom& g, const double* s0, Arr* mac, int bccomp, const int* lo,
                     const int* hi, const int* adv_bc, const int* pmask) {
  for (int d = 0; d < 3; ++d) {
    Box b;
    for (int q = 0; q < 3; ++q) { b.lo[q] = lo[q] - 1; b.hi[q] = hi[q] + 1 + (q == d ? 1 : 0); }
    mac[d].alloc(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], 1);
  }
  if (g.s0mac_interp_type == 1) {
    Arr c = cart_with_ghosts(P, g, s0, 2, bccomp, lo, hi, adv_bc, pmask);
    make_s0mac_sphr(P, g, s0, mac, &c, lo, hi);
  } else {
    make_s0mac_sphr(P, g, s0, mac, nullptr, lo, hi);
  }
}

// mkrhohforce with spherical == 1 (mkscalforce.f90:31 -> mkrhohforce_3d_sphr :388): p0 = (p0_1 + p0_2)/2 on the cell
// centres and on the faces (make_s0mac), u.grad p0 = div(u p0) - p0 div u, psi where the reference adds it, thermal.
// Writes the rhoh component of scal_force on the valid cells.
void mkrhohforce_sphr_box(const mgpu_params& P, const mgpu_geom& g, Arr& scal_force, bool is_prediction, const Arr& thermal,
                          const Arr* umac, const double* p0_1, const double* p0_2, const double* psi, bool add_thermal,
                          const int* lo, const int* hi, const int* adv_bc, const int* pmask) {
  const int ept = P.enthalpy_pred_type, nr = g.nr_fine;
  const int foextrap_comp = 3 + P.nscal + 2;
  if (is_prediction && !(ept == MGPU_PREDICT_RHOHPRIME || ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH))
    fail("ERROR: should only call mkrhohforce when predicting rhoh', h, or rhoh");  // mkscalforce.f90:87-92
  std::vector<double> p0(nr);
  for (int r = 0; r < nr; ++r) p0[r] = 0.5 * (p0_1[r] + p0_2[r]);
  Arr p0_cart = cart_with_ghosts(P, g, p0.data(), 1, foextrap_comp, lo, hi, adv_bc, pmask);
  Arr p0mac[3];
  s0mac_of(P, g, p0.data(), p0mac, foextrap_comp, lo, hi, adv_bc, pmask);
  const int rhoh = P.rhoh_comp - 1;
  const double* dx = P.dx;
  Box vb = grown(lo, hi, 3, 0);
  const Arr &um = umac[0], &vm = umac[1], &wm = umac[2];
  for_box(vb, [&](int i, int j, int k) {
    const double divup = (um(i + 1, j, k) * p0mac[0](i + 1, j, k) - um(i, j, k) * p0mac[0](i, j, k)) / dx[0] +
                         (vm(i, j + 1, k) * p0mac[1](i, j + 1, k) - vm(i, j, k) * p0mac[1](i, j, k)) / dx[1] +
                         (wm(i, j, k + 1) * p0mac[2](i, j, k + 1) - wm(i, j, k) * p0mac[2](i, j, k)) / dx[2];
    const double p0divu = ((um(i + 1, j, k) - um(i, j, k)) / dx[0] + (vm(i, j + 1, k) - vm(i, j, k)) / dx[1] +
                           (wm(i, j, k + 1) - wm(i, j, k)) / dx[2]) * p0_cart(i, j, k);
    scal_force(i, j, k, rhoh) = divup - p0divu;
  });
  if ((is_prediction && ept == MGPU_PREDICT_H) || (is_prediction && ept == MGPU_PREDICT_RHOH) || !is_prediction) {
    Arr psi_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 1);
    put_1d_array_on_cart_sphr(P, g, false, false, psi, psi_cart, lo, hi);
    for_box(vb, [&](int i, int j, int k) { scal_force(i, j, k, rhoh) = scal_force(i, j, k, rhoh) + psi_cart(i, j, k); });
  }
  if (add_thermal)
    for_box(vb, [&](int i, int j, int k) { scal_force(i, j, k, rhoh) = scal_force(i, j, k, rhoh) + thermal(i, j, k); });
}

// enthalpy_advance (enthalpy_advance.f90:16) with spherical == 1
void enthalpy_advance_sphr_box(const mgpu_params& P, const mgpu_geom& g, int which_step, Arr& sold, Arr& snew, Arr* sedge,
                               Arr* sflux, Arr& scal_force, const Arr& thermal, Arr* umac, const double* w0,
                               const Arr* w0mac, const double* rho0_old, const double* rhoh0_old, const double* rho0_new,
                               const double* rhoh0_new, const double* p0_old, const double* p0_new, const double* tempbar,
                               const double* psi, const int* lo, const int* hi, int ng_s, int ng_f, const int* adv_bc,
                               const int* pmask) {
  const int dm = 3, nr = g.nr_fine;
  const int ept = P.enthalpy_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  const int rhoh = P.rhoh_comp - 1, rho = P.rho_comp - 1;
  if (ept == MGPU_PREDICT_HPRIME) fail("mk_rhoh_flux : predict_hprime not coded yet");
  const bool pred_T =
      (ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  const int temp = P.temp_comp - 1;
  if (pred_T && !tempbar) fail("enthalpy_advance: the temperature-based predictions need tempbar");
  Box vb = grown(lo, hi, dm, 0);
  auto fill_umac = [&]() {
    for (int d = 0; d < dm; ++d) fill_boundary_face(P, umac[d], lo, hi, 1, d, pmask);
  };
  auto rhoh_to_h = [&](bool flag) {  // convert_rhoh_to_h, convert_rhoX_to_X.f90:80
    for_box(vb, [&](int i, int j, int k) {
      if (flag) sold(i, j, k, rhoh) = sold(i, j, k, rhoh) / sold(i, j, k, rho);
      else sold(i, j, k, rhoh) = sold(i, j, k, rhoh) * sold(i, j, k, rho);
    });
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc, pmask);
  };
  if (ept == MGPU_PREDICT_H) rhoh_to_h(true);  // :122-126
  scal_force.fill(0.0);                         // :132-134
  if (pred_T) {  // :190-195, mkscalforce.f90:770-776
    Arr p0_cart = cart_with_ghosts(P, g, p0_old, 1, foextrap_comp, lo, hi, adv_bc, pmask);
    Arr psi_cart(vb.lo[0], vb.hi[0], vb.lo[1], vb.hi[1], vb.lo[2], vb.hi[2], 1);
    put_1d_array_on_cart_sphr(P, g, false, false, psi, psi_cart, lo, hi);
    mktempforce_sphr_box(P, scal_force, sold, umac, thermal, p0_cart, psi_cart, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.temp_comp, foextrap_comp, 1, adv_bc, pmask);
  } else {
    mkrhohforce_sphr_box(P, g, scal_force, true, thermal, umac, p0_old, p0_old, psi, true, lo, hi, adv_bc, pmask);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) {  // :141-156
    Arr rhoh0_old_cart = cart_with_ghosts(P, g, rhoh0_old, 1, dm + P.rhoh_comp, lo, hi, adv_bc, pmask);
    Arr fo = scal_force.comp(rhoh), sa = sold.comp(rhoh);
    modify_scal_force_sphr(P, g, fo, sa, umac, rhoh0_old_cart, w0, false, lo, hi);
    fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask);
  } else if (ept == MGPU_PREDICT_H) {  // :173-178
    Box g1 = grown(lo, hi, dm, 1);
    for_box(g1, [&](int i, int j, int k) { scal_force(i, j, k, rhoh) = scal_force(i, j, k, rhoh) / sold(i, j, k, rho); });
  }
  addw0_sphr(umac, w0mac, lo, hi, 1.0);  // :201
  fill_umac();
  auto pert = [&](bool flag) {  // put_in_pert_form on the rhoh component with rhoh0_old
    Arr sa = sold.comp(rhoh);
    pert_form_sphr(P, g, sa, rhoh0_old, flag, lo, hi);
    fill_boundary_box(P, sold, lo, hi, ng_s, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc, pmask);
  };
  auto pert_T = [&](bool flag) {  // :214-217, :268-272
    Arr sa = sold.comp(temp);
    pert_form_sphr(P, g, sa, tempbar, flag, lo, hi);
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
  if (pred_T) {                                    // :280-286, rhoh_vs_t.f90:84-105
    std::vector<double> r0h(nr), rh0h(nr), t0h(nr);
    for (int r = 0; r < nr; ++r) {
      r0h[r] = 0.5 * (rho0_old[r] + rho0_new[r]);
      rh0h[r] = 0.5 * (rhoh0_old[r] + rhoh0_new[r]);
      t0h[r] = 0.5 * (tempbar[r] + tempbar[r]);
    }
    Arr r0c = cart_with_ghosts(P, g, r0h.data(), 2, dm + P.rho_comp, lo, hi, adv_bc, pmask);
    Arr rh0c = cart_with_ghosts(P, g, rh0h.data(), 2, dm + P.rhoh_comp, lo, hi, adv_bc, pmask);
    Arr t0c = cart_with_ghosts(P, g, t0h.data(), 2, dm + P.temp_comp, lo, hi, adv_bc, pmask);
    h_from_rhot_edge_sphr_box(P, sedge, r0c, rh0c, t0c, lo, hi);
  }
  addw0_sphr(umac, w0mac, lo, hi, -1.0);           // :293
  fill_umac();
  // :301-399: rho0mac and h0mac of the old (and, for which_step 2, the new) base state; rhoh0mac is built by the
  // reference too but mk_rhoh_flux_3d_sphr (mkflux.f90:1289) does not read it
  std::vector<double> h0_old(nr), h0_new(nr);
  for (int r = 0; r < nr; ++r) {
    h0_old[r] = rhoh0_old[r] / rho0_old[r];
    h0_new[r] = rhoh0_new[r] / rho0_new[r];
  }
  Arr r0o[3], r0n[3], h0o[3], h0n[3];
  s0mac_of(P, g, rho0_old, r0o, dm + P.rho_comp, lo, hi, adv_bc, pmask);
  s0mac_of(P, g, h0_old.data(), h0o, foextrap_comp, lo, hi, adv_bc, pmask);
  const bool s1 = (which_step == 1);
  if (!s1) {
    s0mac_of(P, g, rho0_new, r0n, dm + P.rho_comp, lo, hi, adv_bc, pmask);
    s0mac_of(P, g, h0_new.data(), h0n, foextrap_comp, lo, hi, adv_bc, pmask);
  }
  mk_rhoh_flux_sphr(P, sflux, sedge, umac, w0mac, r0o, s1 ? r0o : r0n, h0o, s1 ? h0o : h0n, lo, hi);
  scal_force.fill(0.0);  // :401-403
  mkrhohforce_sphr_box(P, g, scal_force, false, thermal, umac, p0_old, s1 ? p0_old : p0_new, psi, false, lo, hi, adv_bc,
                       pmask);  // :405-416
  fill_boundary_box(P, scal_force, lo, hi, ng_f, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask);
  {  // :418-431
    Arr p0_new_cart = cart_with_ghosts(P, g, p0_new, 1, foextrap_comp, lo, hi, adv_bc, pmask);
    update_scal_box(P, P.rhoh_comp, P.rhoh_comp, sold, snew, sflux, scal_force, lo, hi, nullptr, &p0_new_cart);
  }
  fill_boundary_box(P, snew, lo, hi, ng_s, P.rhoh_comp, dm + P.rhoh_comp, 1, adv_bc, pmask);
}

static void need3(const mgpu_params* p) {
  if (p->dm != 3) fail("spherical geometry is 3-D only");
}
static void views3(const mgpu_fab* const* f, int i, Arr* out) {
  for (int d = 0; d < 3; ++d) out[d] = Arr::view(f[d][i], 3);
}

extern "C" {

int mo_put_1d_array_on_cart(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0, mgpu_fab* s0_cart,
                            int is_input_edge_centered, int is_output_a_vector) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr c = Arr::view(s0_cart[i], 3);
    if (c.nc < (is_output_a_vector ? 3 : 1)) fail("put_1d_array_on_cart: s0_cart has too few components");
    put_1d_array_on_cart_sphr(*p, *g, is_input_edge_centered != 0, is_output_a_vector != 0, s0, c, s0_cart[i].lo,
                              s0_cart[i].hi);
  }
  MO_CATCH
}

int mo_make_w0mac(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* w0, mgpu_fab* const* w0mac,
                  const mgpu_fab* w0_cart) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    if (w0mac[0][i].ng != 1) fail("Error: make_w0mac_3d_sphr assumes one ghost cell");  // fill_3d_data.f90:588
    Arr m[3], c;
    views3((const mgpu_fab* const*)w0mac, i, m);
    if (w0_cart) c = Arr::view(w0_cart[i], 3);
    make_w0mac_sphr(*p, *g, w0, m, w0_cart ? &c : nullptr, w0mac[0][i].lo, w0mac[0][i].hi);
  }
  MO_CATCH
}

int mo_make_s0mac(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0, mgpu_fab* const* s0mac,
                  const mgpu_fab* s0_cart) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    if (s0mac[0][i].ng != 1) fail("Error: make_s0mac assumes one ghost cell in s0mac");  // fill_3d_data.f90:992
    Arr m[3], c;
    views3((const mgpu_fab* const*)s0mac, i, m);
    if (s0_cart) c = Arr::view(s0_cart[i], 3);
    make_s0mac_sphr(*p, *g, s0, m, s0_cart ? &c : nullptr, s0mac[0][i].lo, s0mac[0][i].hi);
  }
  MO_CATCH
}

int mo_addw0_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* umac, const mgpu_fab* const* w0mac, double mult) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr um[3], wm[3];
    views3((const mgpu_fab* const*)umac, i, um);
    views3(w0mac, i, wm);
    addw0_sphr(um, wm, umac[0][i].lo, umac[0][i].hi, mult);
  }
  MO_CATCH
}

int mo_mk_rhoX_flux_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                         const mgpu_fab* const* umac, const mgpu_fab* const* w0mac, const mgpu_fab* const* rho0mac_old,
                         const mgpu_fab* const* rho0mac_new, int startcomp, int endcomp) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr sf[3], se[3], um[3], wm[3], ro[3], rn[3];
    views3((const mgpu_fab* const*)sflux, i, sf);
    views3(sedge, i, se);
    views3(umac, i, um);
    views3(w0mac, i, wm);
    views3(rho0mac_old, i, ro);
    views3(rho0mac_new, i, rn);
    mk_rhoX_flux_sphr(*p, sf, se, um, wm, ro, rn, startcomp, endcomp, sflux[0][i].lo, sflux[0][i].hi);
  }
  MO_CATCH
}

int mo_mk_rhoh_flux_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                         const mgpu_fab* const* umac, const mgpu_fab* const* w0mac, const mgpu_fab* const* rho0mac_old,
                         const mgpu_fab* const* rho0mac_new, const mgpu_fab* const* h0mac_old,
                         const mgpu_fab* const* h0mac_new) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr sf[3], se[3], um[3], wm[3], ro[3], rn[3], ho[3], hn[3];
    views3((const mgpu_fab* const*)sflux, i, sf);
    views3(sedge, i, se);
    views3(umac, i, um);
    views3(w0mac, i, wm);
    views3(rho0mac_old, i, ro);
    views3(rho0mac_new, i, rn);
    views3(h0mac_old, i, ho);
    views3(h0mac_new, i, hn);
    mk_rhoh_flux_sphr(*p, sf, se, um, wm, ro, rn, ho, hn, sflux[0][i].lo, sflux[0][i].hi);
  }
  MO_CATCH
}

int mo_update_velocity_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                            const mgpu_fab* const* umac, const mgpu_fab* const* uedge, const mgpu_fab* force,
                            const mgpu_fab* sponge, const mgpu_fab* const* w0mac) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr uo = Arr::view(uold[i], 3), un = Arr::view(unew[i], 3), fo = Arr::view(force[i], 3), sp = Arr::view(sponge[i], 3);
    Arr um[3], ue[3], wm[3];
    views3(umac, i, um);
    views3(uedge, i, ue);
    views3(w0mac, i, wm);
    update_velocity_sphr(*p, uo, un, um, ue, fo, sp, wm, uold[i].lo, uold[i].hi);
  }
  MO_CATCH
}

int mo_mkutrans_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                     mgpu_fab* const* utrans, const mgpu_fab* const* w0mac, const int* adv_bc, const int* phys_bc) {
  MO_TRY
  need3(p);
  if (!p->spherical) fail("mkutrans_sphr: params.spherical must be 1");
  for (int i = 0; i < nfabs; ++i) {
    Arr ut = Arr::view(utilde[i], 3), uf = Arr::view(ufull[i], 3);
    Arr tr[3], wm[3];
    views3((const mgpu_fab* const*)utrans, i, tr);
    views3(w0mac, i, wm);
    mkutrans_box(*p, ut, uf, tr, nullptr, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng, wm);
  }
  MO_CATCH
}

int mo_velpred_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                    mgpu_fab* const* umac, const mgpu_fab* const* utrans, const mgpu_fab* force,
                    const mgpu_fab* const* w0mac, const int* adv_bc, const int* phys_bc) {
  MO_TRY
  need3(p);
  if (!p->spherical) fail("velpred_sphr: params.spherical must be 1");
  for (int i = 0; i < nfabs; ++i) {
    Arr ut = Arr::view(utilde[i], 3), uf = Arr::view(ufull[i], 3), fa = Arr::view(force[i], 3);
    Arr um[3], tr[3], wm[3];
    views3((const mgpu_fab* const*)umac, i, um);
    views3(utrans, i, tr);
    views3(w0mac, i, wm);
    velpred_box(*p, ut, uf, um, tr, fa, nullptr, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng, wm);
  }
  MO_CATCH
}

int mo_modify_scal_force_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* force, const mgpu_fab* s,
                              const mgpu_fab* const* umac, const mgpu_fab* s0_cart, const double* w0, int comp,
                              int fullform) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr fo = Arr::view(force[i], 3).comp(comp - 1), sa = Arr::view(s[i], 3).comp(comp - 1);
    Arr sc = Arr::view(s0_cart[i], 3);
    Arr um[3];
    views3(umac, i, um);
    modify_scal_force_sphr(*p, *g, fo, sa, um, sc, w0, fullform != 0, s[i].lo, s[i].hi);
  }
  MO_CATCH
}

int mo_put_in_pert_form_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* s, const double* s0,
                             int comp, int flag) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr sa = Arr::view(s[i], 3).comp(comp - 1);
    pert_form_sphr(*p, *g, sa, s0, flag != 0, s[i].lo, s[i].hi);
  }
  MO_CATCH
}

int mo_density_advance_sphr(const mgpu_params* p, const mgpu_geom* g, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                            mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force, mgpu_fab* const* umac,
                            const double* w0, const mgpu_fab* const* w0mac, const double* rho0_old,
                            const double* rho0_new, const int* adv_bc, const int* pmask) {
  MO_TRY
  need3(p);
  Arr so = Arr::view(*sold, 3), sn = Arr::view(*snew, 3), fo = Arr::view(*scal_force, 3);
  Arr se[3], sf[3], um[3], wm[3];
  views3((const mgpu_fab* const*)sedge, 0, se);
  views3((const mgpu_fab* const*)sflux, 0, sf);
  views3((const mgpu_fab* const*)umac, 0, um);
  views3(w0mac, 0, wm);
  density_advance_sphr_box(*p, *g, which_step, so, sn, se, sf, fo, um, w0, wm, rho0_old, rho0_new, sold->lo, sold->hi,
                           sold->ng, scal_force->ng, adv_bc, pmask);
  MO_CATCH
}

int mo_make_normal(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* normal) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr n = Arr::view(normal[i], 3);
    make_normal_sphr(*p, *g, n, normal[i].lo, normal[i].hi, normal[i].ng);
  }
  MO_CATCH
}

int mo_mk_vel_force_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* vel_force, int is_final_update,
                         const mgpu_fab* uold, const mgpu_fab* const* uedge, const double* w0, const mgpu_fab* const* w0mac,
                         const mgpu_fab* gpi, const mgpu_fab* s, int index_rho, const mgpu_fab* normal, const double* rho0,
                         const double* grav, const mgpu_fab* w0_force_cart, int do_add_utilde_force) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr f = Arr::view(vel_force[i], 3), uo = Arr::view(uold[i], 3), gp = Arr::view(gpi[i], 3), sv = Arr::view(s[i], 3);
    Arr nm = Arr::view(normal[i], 3), wf = Arr::view(w0_force_cart[i], 3);
    Arr ue[3], wm[3];
    views3(uedge, i, ue);
    views3(w0mac, i, wm);
    mk_vel_force_sphr_box(*p, *g, f, is_final_update != 0, uo, ue, w0, wm, gp, sv.comp(index_rho - 1), nm, rho0, grav, wf,
                          vel_force[i].lo, vel_force[i].hi, do_add_utilde_force != 0);
  }
  MO_CATCH
}

int mo_advance_premac_sphr(const mgpu_params* p, const mgpu_geom* g, const mgpu_fab* uold, const mgpu_fab* sold,
                           mgpu_fab* const* umac, const mgpu_fab* gpi, const mgpu_fab* normal, const double* w0,
                           const mgpu_fab* const* w0mac, const mgpu_fab* w0_force_cart, const double* rho0_old,
                           const double* grav_cell_old, const int* adv_bc, const int* phys_bc, const int* pmask) {
  MO_TRY
  need3(p);
  Arr uo = Arr::view(*uold, 3), so = Arr::view(*sold, 3), gp = Arr::view(*gpi, 3), nm = Arr::view(*normal, 3);
  Arr wf = Arr::view(*w0_force_cart, 3);
  Arr um[3], wm[3];
  views3((const mgpu_fab* const*)umac, 0, um);
  views3(w0mac, 0, wm);
  advance_premac_sphr_box(*p, *g, uo, so, um, gp, nm, w0, wm, wf, rho0_old, grav_cell_old, uold->lo, uold->hi, uold->ng,
                          adv_bc, phys_bc, pmask);
  MO_CATCH
}

int mo_velocity_advance_sphr(const mgpu_params* p, const mgpu_geom* g, const mgpu_fab* uold, mgpu_fab* unew,
                             const mgpu_fab* sold, const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi,
                             const mgpu_fab* normal, const double* w0, const mgpu_fab* const* w0mac,
                             const mgpu_fab* w0_force_cart, const double* rho0_old, const double* rho0_nph,
                             const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                             const int* adv_bc, const int* pmask) {
  MO_TRY
  need3(p);
  Arr uo = Arr::view(*uold, 3), un = Arr::view(*unew, 3), so = Arr::view(*sold, 3), rh = Arr::view(*rhohalf, 3);
  Arr gp = Arr::view(*gpi, 3), nm = Arr::view(*normal, 3), wf = Arr::view(*w0_force_cart, 3), sp = Arr::view(*sponge, 3);
  Arr um[3], wm[3];
  views3((const mgpu_fab* const*)umac, 0, um);
  views3(w0mac, 0, wm);
  velocity_advance_sphr_box(*p, *g, uo, un, so, rh, um, gp, nm, w0, wm, wf, rho0_old, rho0_nph, grav_cell_old,
                            grav_cell_nph, sp, uold->lo, uold->hi, uold->ng, adv_bc, pmask);
  MO_CATCH
}

int mo_mkrhohforce_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* scal_force, int is_prediction,
                        const mgpu_fab* thermal, const mgpu_fab* const* umac, const double* p0_1, const double* p0_2,
                        const double* psi, int add_thermal, const int* adv_bc, const int* pmask) {
  MO_TRY
  need3(p);
  for (int i = 0; i < nfabs; ++i) {
    Arr f = Arr::view(scal_force[i], 3), th = Arr::view(thermal[i], 3);
    Arr um[3];
    views3(umac, i, um);
    mkrhohforce_sphr_box(*p, *g, f, is_prediction != 0, th, um, p0_1, p0_2, psi, add_thermal != 0, scal_force[i].lo,
                         scal_force[i].hi, adv_bc, pmask);
    // ml_restrict_and_fill of the rhoh component (mkscalforce.f90:177-181)
    fill_boundary_box(*p, f, scal_force[i].lo, scal_force[i].hi, scal_force[i].ng, p->rhoh_comp, 3 + p->nscal + 2, 1, adv_bc,
                      pmask);
  }
  MO_CATCH
}

int mo_enthalpy_advance_sphr(const mgpu_params* p, const mgpu_geom* g, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                             mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force, const mgpu_fab* thermal,
                             mgpu_fab* const* umac, const double* w0, const mgpu_fab* const* w0mac, const double* rho0_old,
                             const double* rhoh0_old, const double* rho0_new, const double* rhoh0_new,
                             const double* p0_old, const double* p0_new, const double* tempbar, const double* psi,
                             const int* adv_bc, const int* pmask) {
  MO_TRY
  need3(p);
  Arr so = Arr::view(*sold, 3), sn = Arr::view(*snew, 3), fa = Arr::view(*scal_force, 3), th = Arr::view(*thermal, 3);
  Arr se[3], sf[3], um[3], wm[3];
  views3((const mgpu_fab* const*)sedge, 0, se);
  views3((const mgpu_fab* const*)sflux, 0, sf);
  views3((const mgpu_fab* const*)umac, 0, um);
  views3(w0mac, 0, wm);
  enthalpy_advance_sphr_box(*p, *g, which_step, so, sn, se, sf, fa, th, um, w0, wm, rho0_old, rhoh0_old, rho0_new, rhoh0_new,
                            p0_old, p0_new, tempbar, psi, sold->lo, sold->hi, sold->ng, scal_force->ng, adv_bc, pmask);
  MO_CATCH
}

}  // extern "C"
