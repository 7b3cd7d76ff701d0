// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
// Restatement of the reductions next to the advective path:
//   estdt_2d        Source/estdt.f90:348      estdt_3d_cart   Source/estdt.f90:467
//   sum_etarho_2d   Source/make_eta.f90:176   sum_etarho_3d   Source/make_eta.f90:213
//   estdt_3d_sphr   Source/estdt.f90:620
//   make_etarho_planar Source/make_eta.f90:36 (single level, one chunk: r_end_coord = nr-1)
//   average         Source/average.f90:24 (planar :114-163; spherical :168-362 for one level) + sum_phi_3d_sphr :564
// Same loops, same order, same expressions.  average: parity unpinned (the reference holds no numbers for it); the
// tests anchor it on fields whose average is known (tests/test_average.py).
#include <algorithm>
#include <cmath>
#include <limits>

#include "mo_kernels.h"

namespace mo {

// dt inout, umax out (estdt.f90:364-465 / :483-616)
void estdt_box(const mgpu_params& P, const Arr& u, const Arr& s, const Arr& force, const Arr& divU, const Arr& dSdt,
               const double* w0, const double* p0, const double* gamma1bar, const int* lo, const int* hi,
               double rho_min, double cfl, double& dt, double& umax) {
  const int dm = P.dm, r = dm - 1, nr = P.nr, rho = P.rho_comp - 1;
  const double eps = 1.0e-8;
  const int k0 = dm == 3 ? lo[2] : 0, k1 = dm == 3 ? hi[2] : 0;
  double spd[3] = {0.0, 0.0, 0.0}, spdr = 0.0;
  umax = 0.0;
  // Limit dt based on velocity terms
  for (int d = 0; d < dm; ++d)
    for (int k = k0; k <= k1; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          const int kr = dm == 3 ? k : j;
          if (d == r) spd[d] = std::max(spd[d], std::fabs(u(i, j, k, d) + 0.5 * (w0[kr] + w0[kr + 1])));
          else spd[d] = std::max(spd[d], std::fabs(u(i, j, k, d)));
        }
  for (int k = lo[r]; k <= hi[r]; ++k) spdr = std::max(spdr, std::fabs(w0[k]));
  for (int d = 0; d < dm; ++d) umax = std::max(umax, spd[d]);
  umax = std::max(umax, spdr);
  for (int d = 0; d < dm; ++d)
    if (spd[d] > eps) dt = std::min(dt, P.dx[d] / spd[d]);
  if (spdr > eps) dt = std::min(dt, P.dx[r] / spdr);
  dt = dt * cfl;
  // Limit dt based on forcing terms
  double f[3] = {0.0, 0.0, 0.0};
  for (int d = 0; d < dm; ++d)
    for (int k = k0; k <= k1; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) f[d] = std::max(f[d], std::fabs(force(i, j, k, d)));
  for (int d = 0; d < dm; ++d)
    if (f[d] > eps) dt = std::min(dt, std::sqrt(2.0 * P.dx[d] / f[d]));
  // divU constraint
  for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j) {
      const int kr = dm == 3 ? k : j;
      double gradp0;
      if (kr == 0) gradp0 = (p0[kr + 1] - p0[kr]) / P.dx[r];
      else if (kr == nr - 1) gradp0 = (p0[kr] - p0[kr - 1]) / P.dx[r];
      else gradp0 = 0.5 * (p0[kr + 1] - p0[kr - 1]) / P.dx[r];
      for (int i = lo[0]; i <= hi[0]; ++i) {
        const double denom = divU(i, j, k) - u(i, j, k, r) * gradp0 / (gamma1bar[kr] * p0[kr]);
        if (denom > 0.0) dt = std::min(dt, 0.4 * (1.0 - rho_min / s(i, j, k, rho)) / denom);
      }
    }
  // dS/dt constraint
  for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
      for (int i = lo[0]; i <= hi[0]; ++i)
        if (dSdt(i, j, k) > 1.e-20) {
          const double a = 0.5 * s(i, j, k, rho) * dSdt(i, j, k);
          const double b = s(i, j, k, rho) * divU(i, j, k);
          const double c = rho_min - s(i, j, k, rho);
          dt = std::min(dt, 0.4 * 2.0 * c / (-b - std::sqrt(b * b - 4.0 * a * c)));
        }
}

// estdt.f90:142-220 for one level on one rank
void estdt_level(const mgpu_params& P, int nfabs, const mgpu_fab* u, const mgpu_fab* s, const mgpu_fab* force,
                 const mgpu_fab* divU, const mgpu_fab* dSdt, const double* w0, const double* p0,
                 const double* gamma1bar, double rho_min, double cfl, double& dt, double& umax) {
  const double dt_start = 1.e99;
  double dt_proc = 1.e99, umax_proc = 0.0;
  for (int i = 0; i < nfabs; ++i) {
    Arr ua = Arr::view(u[i], P.dm), sa = Arr::view(s[i], P.dm), fa = Arr::view(force[i], P.dm);
    Arr dU = Arr::view(divU[i], P.dm), dS = Arr::view(dSdt[i], P.dm);
    double dt_grid = std::numeric_limits<double>::max(), umax_grid = 0.0;
    estdt_box(P, ua, sa, fa, dU, dS, w0, p0, gamma1bar, u[i].lo, u[i].hi, rho_min, cfl, dt_grid, umax_grid);
    dt_proc = std::min(dt_proc, dt_grid);
    umax_proc = std::max(umax_proc, umax_grid);
  }
  double dt_lev = dt_proc;
  umax = std::max(umax, umax_proc);
  if (dt_lev == dt_start) {
    dt_lev = P.dx[0];
    for (int d = 1; d < P.dm; ++d) dt_lev = std::min(dt_lev, P.dx[d]);
  }
  dt = std::min(dt, dt_lev);
}

// estdt_3d_sphr, estdt.f90:620-773 (dt inout, umax out)
void estdt_sphr_box(const mgpu_params& P, const mgpu_geom& g, const Arr& u, const Arr& s, const Arr& force,
                    const Arr& divU, const Arr& dSdt, const Arr* w0mac, const double* w0, const double* p0,
                    const double* gamma1bar, const int* lo, const int* hi, double rho_min, double cfl, double& dt,
                    double& umax) {
  const int nr_fine = g.nr_fine, rho = P.rho_comp - 1;
  const double eps = 1.0e-8;
  double spd[3] = {0.0, 0.0, 0.0}, spdr = 0.0;
  umax = 0.0;
  for (int d = 0; d < 3; ++d)
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          const double whi = w0mac[d](i + (d == 0), j + (d == 1), k + (d == 2));
          spd[d] = std::max(spd[d], std::fabs(u(i, j, k, d) + 0.5 * (w0mac[d](i, j, k) + whi)));
        }
  for (int k = 0; k <= nr_fine; ++k) spdr = std::max(spdr, std::fabs(w0[k]));
  for (int d = 0; d < 3; ++d) umax = std::max(umax, spd[d]);
  umax = std::max(umax, spdr);
  for (int d = 0; d < 3; ++d)
    if (spd[d] > eps) dt = std::min(dt, P.dx[d] / spd[d]);
  if (spdr > eps) dt = std::min(dt, g.dr / spdr);
  dt = dt * cfl;
  double f[3] = {0.0, 0.0, 0.0};
  for (int d = 0; d < 3; ++d)
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) f[d] = std::max(f[d], std::fabs(force(i, j, k, d)));
  for (int d = 0; d < 3; ++d)
    if (f[d] > eps) dt = std::min(dt, std::sqrt(2.0 * P.dx[d] / f[d]));
  std::vector<double> gp0(nr_fine + 1);
  for (int r = 1; r <= nr_fine - 1; ++r) {
    const double gamma1bar_p_avg = 0.5 * (gamma1bar[r] * p0[r] + gamma1bar[r - 1] * p0[r - 1]);
    gp0[r] = ((p0[r] - p0[r - 1]) / g.dr) / gamma1bar_p_avg;
  }
  gp0[nr_fine] = gp0[nr_fine - 1];
  gp0[0] = gp0[1];
  Arr gp0_cart(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], 3);
  put_1d_array_on_cart_sphr(P, g, true, true, gp0.data(), gp0_cart, lo, hi);
  for (int k = lo[2]; k <= hi[2]; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
      for (int i = lo[0]; i <= hi[0]; ++i) {
        const double gp_dot_u =
            u(i, j, k, 0) * gp0_cart(i, j, k, 0) + u(i, j, k, 1) * gp0_cart(i, j, k, 1) + u(i, j, k, 2) * gp0_cart(i, j, k, 2);
        const double denom = divU(i, j, k) - gp_dot_u;
        if (denom > 0.0) dt = std::min(dt, 0.4 * (1.0 - rho_min / s(i, j, k, rho)) / denom);
        if (dSdt(i, j, k) > 1.e-20) {
          const double a = 0.5 * s(i, j, k, rho) * dSdt(i, j, k);
          const double b = s(i, j, k, rho) * divU(i, j, k);
          const double c = rho_min - s(i, j, k, rho);
          dt = std::min(dt, 0.4 * 2.0 * c / (-b - std::sqrt(b * b - 4.0 * a * c)));
        }
      }
}

void estdt_sphr_level(const mgpu_params& P, const mgpu_geom& g, int nfabs, const mgpu_fab* u, const mgpu_fab* s,
                      const mgpu_fab* force, const mgpu_fab* divU, const mgpu_fab* dSdt, const mgpu_fab* const* w0mac,
                      const double* w0, const double* p0, const double* gamma1bar, double rho_min, double cfl,
                      double& dt, double& umax) {
  const double dt_start = 1.e99;
  double dt_proc = 1.e99, umax_proc = 0.0;
  for (int i = 0; i < nfabs; ++i) {
    Arr ua = Arr::view(u[i], 3), sa = Arr::view(s[i], 3), fa = Arr::view(force[i], 3);
    Arr dU = Arr::view(divU[i], 3), dS = Arr::view(dSdt[i], 3);
    Arr wm[3] = {Arr::view(w0mac[0][i], 3), Arr::view(w0mac[1][i], 3), Arr::view(w0mac[2][i], 3)};
    double dt_grid = std::numeric_limits<double>::max(), umax_grid = 0.0;
    estdt_sphr_box(P, g, ua, sa, fa, dU, dS, wm, w0, p0, gamma1bar, u[i].lo, u[i].hi, rho_min, cfl, dt_grid, umax_grid);
    dt_proc = std::min(dt_proc, dt_grid);
    umax_proc = std::max(umax_proc, umax_grid);
  }
  double dt_lev = dt_proc;
  umax = std::max(umax, umax_proc);
  if (dt_lev == dt_start) dt_lev = std::min(P.dx[0], std::min(P.dx[1], P.dx[2]));
  dt = std::min(dt, dt_lev);
}

// make_eta.f90:176-250: the reference's summation order (i fastest, then j, then k)
void sum_etarho_box(const mgpu_params& P, const Arr& e, const int* lo, const int* hi, double* etarhosum) {
  const int dm = P.dm, r = dm - 1;
  auto plane = [&](int kk) {
    if (dm == 3) {
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) etarhosum[kk] = etarhosum[kk] + e(i, j, kk);
    } else {
      for (int i = lo[0]; i <= hi[0]; ++i) etarhosum[kk] = etarhosum[kk] + e(i, kk, 0);
    }
  };
  for (int kk = lo[r]; kk <= hi[r]; ++kk) plane(kk);
  if (hi[r] == P.nr - 1) plane(hi[r] + 1);  // top edge only at the top of the domain
}

void make_etarho_planar(const mgpu_params& P, int nfabs, const mgpu_fab* etarhoflux, double* etarho_ec,
                        double* etarho_cc) {
  if (P.spherical) fail("ERROR: make_eta should not be called for spherical");
  const int nr = P.nr, r = P.dm - 1;
  std::vector<double> sum(nr + 1, 0.0);
  for (int i = 0; i < nfabs; ++i) {
    Arr e = Arr::view(etarhoflux[i], P.dm);
    sum_etarho_box(P, e, etarhoflux[i].lo, etarhoflux[i].hi, sum.data());
  }
  double ncell = 1.0;
  for (int d = 0; d < r; ++d) ncell *= (double)(P.domhi[d] - P.domlo[d] + 1);
  for (int k = 0; k <= nr; ++k) etarho_ec[k] = sum[k] / ncell;
  for (int k = 0; k < nr; ++k) etarho_cc[k] = 0.5 * (etarho_ec[k] + etarho_ec[k + 1]);
}

// ---- average (Source/average.f90:24), one level -------------------------------------------------------------------
namespace {
// quad_interp, average.f90:386-401
double avg_quad_interp(double x, double x0, double x1, double x2, double y0, double y1, double y2, bool limit) {
  double y = y0 + (y1 - y0) / (x1 - x0) * (x - x0) +
             ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0)) / (x2 - x0) * (x - x0) * (x - x1);
  if (limit) {
    if (y > dmax(dmax(y0, y1), y2)) y = dmax(dmax(y0, y1), y2);
    if (y < dmin(dmin(y0, y1), y2)) y = dmin(dmin(y0, y1), y2);
  }
  return y;
}
}  // namespace

// sum_phi_3d_sphr, average.f90:564-618 (no mask: one level); radii(0:nr_irreg+1)
void sum_phi_sphr_box(const mgpu_params& P, const mgpu_geom& g, const double* radii, int nr_irreg, const Arr& phi, int comp,
                      const int* lo, const int* hi, double* phisum, long* ncell) {
  for (int k = lo[2]; k <= hi[2]; ++k) {
    const double z = g.prob_lo[2] + ((double)k + 0.5) * P.dx[2] - g.center[2];
    for (int j = lo[1]; j <= hi[1]; ++j) {
      const double y = g.prob_lo[1] + ((double)j + 0.5) * P.dx[1] - g.center[1];
      for (int i = lo[0]; i <= hi[0]; ++i) {
        const double x = g.prob_lo[0] + ((double)i + 0.5) * P.dx[0] - g.center[0];
        const double radius = std::sqrt(x * x + y * y + z * z);
        const double q = radius / P.dx[0];
        int index = (int)((q * q - 0.75) / 2.0);
        if (index < nr_irreg) {
          if (dabs(radius - radii[index]) > dabs(radius - radii[index + 1])) index = index + 1;
        }
        phisum[index] = phisum[index] + phi(i, j, k, comp);
        ncell[index] = ncell[index] + 1;
      }
    }
  }
}

// the part of average() after the sums for spherical == 1 and nlevs = max_levs = 1 (average.f90:204-362):
// phisum / ncell / radii are indexed -1..nr_irreg(+1) in the reference -> offset 1 here
void average_sphr_tail(const mgpu_geom& g, int nr_irreg, int drdxfac, std::vector<double>& phisum,
                       std::vector<long>& ncell, std::vector<double>& radii, double* phibar) {
  auto PS = [&](int r) -> double& { return phisum[r + 1]; };
  auto NC = [&](int r) -> long& { return ncell[r + 1]; };
  auto RD = [&](int r) -> double& { return radii[r + 1]; };
  for (int r = 0; r <= nr_irreg; ++r)
    if (NC(r) != 0) PS(r) = PS(r) / (double)NC(r);  // :204-210
  PS(-1) = (11.0 / 8.0) * PS(0) - (3.0 / 8.0) * PS(1);  // :213-215
  RD(-1) = 0.0;
  NC(-1) = 1;
  // :217-283 choose the level to interpolate from: one level, which_lev(r) = 1
  int max_rcoord = nr_irreg;
  {  // :286-309 squish the list down to the radii that received a cell
    int j = 0;
    for (int r = 0; r <= nr_irreg; ++r) {
      while (NC(j) == 0) {
        j = j + 1;
        if (j > nr_irreg) break;
      }
      if (j > nr_irreg) {
        for (int q = r; q <= nr_irreg; ++q) PS(q) = 1.e99;
        for (int q = r; q <= nr_irreg + 1; ++q) RD(q) = 1.e99;
        max_rcoord = r - 1;
        break;
      }
      PS(r) = PS(j);
      RD(r) = RD(j);
      NC(r) = NC(j);
      j = j + 1;
      if (j > nr_irreg) {
        max_rcoord = r;
        break;
      }
    }
  }
  int stencil_coord = 0;  // :312-352
  for (int r = 0; r < g.nr_fine; ++r) {
    const double radius = ((double)r + 0.5) * g.dr;
    for (int j = stencil_coord; j <= max_rcoord; ++j)
      if (dabs(radius - RD(j)) < dabs(radius - RD(j + 1))) {
        stencil_coord = j;
        break;
      }
    stencil_coord = std::min(stencil_coord, max_rcoord - 1);
    const bool limit = !((double)r > (double)(g.nr_fine - 1) - (double)drdxfac * 1.0);  // 2.d0**(max_levs-1) = 1
    phibar[r] = avg_quad_interp(radius, RD(stencil_coord - 1), RD(stencil_coord), RD(stencil_coord + 1),
                                PS(stencil_coord - 1), PS(stencil_coord), PS(stencil_coord + 1), limit);
  }
}

void average_level(const mgpu_params& P, const mgpu_geom* g, int nfabs, const mgpu_fab* phi, int incomp, int nr_irreg,
                   int drdxfac, double* phibar) {
  const int dm = P.dm, comp = incomp - 1;
  for (int f = 0; f < nfabs; ++f)
    if (incomp < 1 || incomp > phi[f].nc) fail("average: incomp out of range");
  if (!P.spherical) {  // :114-163: every cell of a plane, divided by the cells of the domain's plane
    const int nr = P.nr, r = dm - 1;
    std::vector<double> sum(nr, 0.0);
    for (int f = 0; f < nfabs; ++f) {  // sum_phi_2d / _3d (:520-560): k outermost, i innermost
      Arr a = Arr::view(phi[f], dm);
      const int* lo = phi[f].lo;
      const int* hi = phi[f].hi;
      if (dm == 3) {
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) sum[k] = sum[k] + a(i, j, k, comp);
      } else {
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i) sum[j] = sum[j] + a(i, j, 0, comp);
      }
    }
    double ncell = 1.0;
    for (int d = 0; d < r; ++d) ncell *= (double)(P.domhi[d] - P.domlo[d] + 1);
    for (int k = 0; k < nr; ++k) phibar[k] = sum[k] / ncell;
    return;
  }
  if (!g || dm != 3) fail("average: spherical geometry needs mgpu_geom (3-D)");
  std::vector<double> radii(nr_irreg + 3), phisum(nr_irreg + 2, 0.0);
  std::vector<long> ncell(nr_irreg + 2, 0);
  for (int r = 0; r <= nr_irreg; ++r) radii[r + 1] = std::sqrt(0.75 + 2.0 * r) * P.dx[0];  // :92
  radii[nr_irreg + 2] = 1.e99;
  for (int f = 0; f < nfabs; ++f) {  // the reference would write past phisum(nr_irreg) here: refuse instead
    double far2 = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double a = g->prob_lo[d] + ((double)phi[f].lo[d] + 0.5) * P.dx[d] - g->center[d];
      const double b = g->prob_lo[d] + ((double)phi[f].hi[d] + 0.5) * P.dx[d] - g->center[d];
      far2 += dmax(a * a, b * b);
    }
    if ((int)((far2 / (P.dx[0] * P.dx[0]) - 0.75) / 2.0) > nr_irreg) fail("average: a cell maps beyond nr_irreg");
  }
  for (int f = 0; f < nfabs; ++f) {
    Arr a = Arr::view(phi[f], dm);
    sum_phi_sphr_box(P, *g, radii.data() + 1, nr_irreg, a, comp, phi[f].lo, phi[f].hi, phisum.data() + 1, ncell.data() + 1);
  }
  average_sphr_tail(*g, nr_irreg, drdxfac, phisum, ncell, radii, phibar);
}

}  // namespace mo
