// Spherical-geometry branches of the path on the device (3-D only), one thread per zone / face, x along the
// threads so every warp reads and writes whole lines.  Built with -fmad=false: the expression trees are those of
// the reference, so results are bit-identical to the CPU oracle (sqrt and division are IEEE-correct on the device).
//   put_1d_array_on_cart_3d_sphr  Source/fill_3d_data.f90:269      quad_interp :535
//   make_w0mac_3d_sphr            Source/fill_3d_data.f90:621      make_s0mac_3d_sphr :1017
//   addw0_3d_sphr                 Source/addw0.f90:171
//   mk_rhoX_flux_3d_sphr          Source/mkflux.f90:509            mk_rhoh_flux_3d_sphr :1289
//   update_velocity_3d            Source/update_vel.f90:227 (spherical == 1 branch :317-360)
//   modify_scal_force_3d_sphr     Source/modify_scal_force.f90:256
//   pert_form_3d_sphr             Source/put_in_pert_form.f90:185
// All are HBM-bound streaming kernels except the interpolation ones, which are bound by the fp64 sqrt/divide.
#include "mgpu_sphr.cuh"

namespace mgpu {
namespace {

__device__ __forceinline__ double max3(double a, double b, double c) {
  const double m = a > b ? a : b;
  return m > c ? m : c;
}
__device__ __forceinline__ double min3(double a, double b, double c) {
  const double m = a < b ? a : b;
  return m < c ? m : c;
}
__device__ __forceinline__ double quad_interp(double x, double x0, double x1, double x2, double y0, double y1,
                                              double y2) {
  double y = y0 + (y1 - y0) / (x1 - x0) * (x - x0) +
             ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0)) / (x2 - x0) * (x - x0) * (x - x1);
  if (y > max3(y0, y1, y2)) y = max3(y0, y1, y2);
  if (y < min3(y0, y1, y2)) y = min3(y0, y1, y2);
  return y;
}

// FAST build (Geom::fast, every option but exact = 1): the radial grid is uniform, r_cc_loc(i) = (i + 1/2) dr and
// r_edge_loc(i) = i dr, so the six divisions of quad_interp by differences of neighbouring radii (each dr or 2 dr to the
// last bit or two) and the divisions by dr of the linear forms become multiplications by 1/dr: the kernels that map the
// base state onto the grid (put_1d_array_on_cart, make_s0mac / make_w0mac, the spherical force builders) held ~10 fp64
// divisions and a square root per zone and ran at the fp64 pipe's pace.  The bin index keeps its division (a rounding
// there would pick another stencil).  <= 1e-14 relative from the exact form.
__device__ __forceinline__ double quad_interp_fast(double x, double x0, double x1, double rdr, double y0, double y1,
                                                   double y2) {
  double y = y0 + (y1 - y0) * rdr * (x - x0) + ((y2 - y1) - (y1 - y0)) * (0.5 * rdr * rdr) * (x - x0) * (x - x1);
  if (y > max3(y0, y1, y2)) y = max3(y0, y1, y2);
  if (y < min3(y0, y1, y2)) y = min3(y0, y1, y2);
  return y;
}

__device__ double interp_edge(const Geom& g, int type, const double* s0, double radius) {
  const double dr = g.dr;
  int index = (int)(radius / dr);
  if (type == 1) {
    const double rfac = (radius - (double)index * dr) / dr;
    return (rfac > 0.5) ? s0[index + 1] : s0[index];
  }
  if (type == 2) {
    const double rfac = g.fast ? (radius - (double)index * dr) * g.rdr : (radius - (double)index * dr) / dr;
    if (index < g.nr_fine) return rfac * s0[index + 1] + (1.0 - rfac) * s0[index];
    return s0[g.nr_fine];
  }
  // QUIRK fill_3d_data.f90:381: the third test compares a distance with a position
  if (index <= 0) index = 0;
  else if (index >= g.nr_fine - 1) index = g.nr_fine - 2;
  else if (radius - g.r_edge_loc[index] < g.r_edge_loc[index + 1]) index = index - 1;
  if (g.fast)
    return quad_interp_fast(radius, g.r_edge_loc[index], g.r_edge_loc[index + 1], g.rdr, s0[index], s0[index + 1],
                            s0[index + 2]);
  return quad_interp(radius, g.r_edge_loc[index], g.r_edge_loc[index + 1], g.r_edge_loc[index + 2], s0[index],
                     s0[index + 1], s0[index + 2]);
}

__device__ double interp_cc(const Geom& g, int type, const double* s0, double radius) {
  const double dr = g.dr;
  int index = (int)(radius / dr);
  const int nr = g.nr_fine;
  if (type == 1) return s0[index];
  if (type == 2) {
    if (radius >= g.r_cc_loc[index]) {
      if (index >= nr - 1) return s0[nr - 1];
      if (g.fast)
        return s0[index + 1] * (radius - g.r_cc_loc[index]) * g.rdr + s0[index] * (g.r_cc_loc[index + 1] - radius) * g.rdr;
      return s0[index + 1] * (radius - g.r_cc_loc[index]) / dr + s0[index] * (g.r_cc_loc[index + 1] - radius) / dr;
    }
    if (index == 0) return s0[index];
    if (index > nr - 1) return s0[nr - 1];
    if (g.fast)
      return s0[index] * (radius - g.r_cc_loc[index - 1]) * g.rdr + s0[index - 1] * (g.r_cc_loc[index] - radius) * g.rdr;
    return s0[index] * (radius - g.r_cc_loc[index - 1]) / dr + s0[index - 1] * (g.r_cc_loc[index] - radius) / dr;
  }
  if (index == 0) index = 1;
  else if (index >= nr - 1) index = nr - 2;
  if (g.fast)
    return quad_interp_fast(radius, g.r_cc_loc[index - 1], g.r_cc_loc[index], g.rdr, s0[index - 1], s0[index],
                            s0[index + 1]);
  return quad_interp(radius, g.r_cc_loc[index - 1], g.r_cc_loc[index], g.r_cc_loc[index + 1], s0[index - 1], s0[index],
                     s0[index + 1]);
}

__device__ __forceinline__ double pos(const Geom& g, int d, int i, bool half) {
  return g.prob_lo[d] + ((double)i + (half ? 0.5 : 0.0)) * g.dx[d] - g.center[d];
}
__device__ __forceinline__ double radius_of(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }

__global__ void k_put1d(Geom g, DV cart, Box3 vb, const double* s0, bool edge_in, bool vec, int type) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const double x = pos(g, 0, ix[0], true), y = pos(g, 1, ix[1], true), z = pos(g, 2, ix[2], true);
  const double radius = radius_of(x, y, z);
  const double v = edge_in ? interp_edge(g, type, s0, radius) : interp_cc(g, type, s0, radius);
  if (vec) {
    cart(ix[0], ix[1], ix[2], 0) = v * x * (1.0 / radius);
    cart(ix[0], ix[1], ix[2], 1) = v * y * (1.0 / radius);
    cart(ix[0], ix[1], ix[2], 2) = v * z * (1.0 / radius);
  } else {
    cart(ix[0], ix[1], ix[2], 0) = v;
  }
}

// kind 0: w0mac (edge-centred input, times the direction cosine), kind 1: s0mac (bin-centred input, scalar)
__global__ void k_make_mac(Geom g, DV m, DV cart, Box3 fb, const double* s0, int d, int kind, int type) {
  int ix[3];
  if (!decode3(fb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  if (type == 1) {
    const int il = i - (d == 0), jl = j - (d == 1), kl = k - (d == 2);
    if (kind == 0) m(i, j, k) = 0.5 * (cart(il, jl, kl, d) + cart(i, j, k, d));
    else m(i, j, k) = 0.5 * (cart(i, j, k, 0) + cart(il, jl, kl, 0));
    return;
  }
  if (type == 4) {  // w0mac only: linear interpolation to the nodes, averaged over the four nodes of the face
    const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
    double sum = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int n[3] = {i, j, k};
      if (c & 1) n[t1] += 1;
      if (c & 2) n[t2] += 1;
      const double x = pos(g, 0, n[0], false), y = pos(g, 1, n[1], false), z = pos(g, 2, n[2], false);
      const double radius = radius_of(x, y, z);
      const double v = interp_edge(g, 2, s0, radius);
      const double cc = (d == 0) ? x : (d == 1 ? y : z);
      const double nodal = v * cc * (1.0 / radius);
      sum = (c == 0) ? nodal : sum + nodal;
    }
    m(i, j, k) = 0.25 * sum;
    return;
  }
  const double x = pos(g, 0, i, d != 0), y = pos(g, 1, j, d != 1), z = pos(g, 2, k, d != 2);
  const double radius = radius_of(x, y, z);
  if (kind == 0) {
    const double v = interp_edge(g, type, s0, radius);
    const double c = (d == 0) ? x : (d == 1 ? y : z);
    m(i, j, k) = v * c / radius;
  } else {
    m(i, j, k) = interp_cc(g, type, s0, radius);
  }
}

__global__ void k_addw0_sphr(DV um, DV wm, Box3 fb, double mult) {
  int ix[3];
  if (!decode3(fb, ix)) return;
  um(ix[0], ix[1], ix[2]) = um(ix[0], ix[1], ix[2]) + mult * wm(ix[0], ix[1], ix[2]);
}

// components comp0 .. comp1 of one face direction in one pass: the velocity and the density factor of a face are
// the same for every component (mkflux.f90:560-640), the product is formed in the reference's order, vel * rho * X
__global__ void k_rhoX_flux_sphr(SphrFluxArgs a, int d, int comp0, int comp1) {
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[d] += 1;
  if (!decode3(fb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double vel = a.umac[d](i, j, k) + a.w0mac[d](i, j, k);
  const DV& se = a.sedge[d];
  double vr;
  if (a.spt == MGPU_PREDICT_RHOPRIME_AND_X) {
    const double rho0_edge = 0.5 * (a.r0o[d](i, j, k) + a.r0n[d](i, j, k));
    vr = vel * (rho0_edge + se(i, j, k, a.rho));
  } else if (a.spt == MGPU_PREDICT_RHOX) {
    vr = vel;
  } else {
    vr = vel * se(i, j, k, a.rho);
  }
  for (int comp = comp0; comp <= comp1; ++comp) a.sflux[d](i, j, k, comp) = vr * se(i, j, k, comp);
}

// mode: 0 h, 1 h', 2 rhoh, 3 (rho h)'
__global__ void k_rhoh_flux_sphr(SphrFluxArgs a, int d, int mode) {
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[d] += 1;
  if (!decode3(fb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double vel = a.umac[d](i, j, k) + a.w0mac[d](i, j, k);
  const double erho = a.sedge[d](i, j, k, a.rho), erhoh = a.sedge[d](i, j, k, a.rhoh);
  double f;
  if (mode == 0) {
    if (a.spt == MGPU_PREDICT_RHOPRIME_AND_X) {
      const double rho0_edge = 0.5 * (a.r0o[d](i, j, k) + a.r0n[d](i, j, k));
      f = vel * (rho0_edge + erho) * erhoh;
    } else {
      f = vel * erho * erhoh;
    }
  } else if (mode == 1) {
    const double rho0_edge = 0.5 * (a.r0o[d](i, j, k) + a.r0n[d](i, j, k));
    const double h0_edge = 0.5 * (a.h0o[d](i, j, k) + a.h0n[d](i, j, k));
    f = vel * (erho + rho0_edge) * (erhoh + h0_edge);
  } else if (mode == 2) {
    f = vel * erhoh;
  } else {
    const double rho0_edge = 0.5 * (a.r0o[d](i, j, k) + a.r0n[d](i, j, k));
    const double h0_edge = 0.5 * (a.h0o[d](i, j, k) + a.h0n[d](i, j, k));
    f = vel * (rho0_edge * h0_edge + erhoh);
  }
  a.sflux[d](i, j, k, a.rhoh) = f;
}

__global__ void k_update_vel_sphr(VelArgs a, DV w0x, DV w0y, DV w0z) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double ubar = 0.5 * (a.umac[0](i, j, k) + a.umac[0](i + 1, j, k));
  const double vbar = 0.5 * (a.umac[1](i, j, k) + a.umac[1](i, j + 1, k));
  const double wbar = 0.5 * (a.umac[2](i, j, k) + a.umac[2](i, j, k + 1));
  const double w0xb = 0.5 * (w0x(i, j, k) + w0x(i + 1, j, k));
  const double w0yb = 0.5 * (w0y(i, j, k) + w0y(i, j + 1, k));
  const double w0zb = 0.5 * (w0z(i, j, k) + w0z(i, j, k + 1));
  const double sp = a.do_sponge ? a.sponge(i, j, k) : 1.0;
  for (int c = 0; c < 3; ++c) {
    const double dux = a.uedge[0](i + 1, j, k, c) - a.uedge[0](i, j, k, c);
    const double duy = a.uedge[1](i, j + 1, k, c) - a.uedge[1](i, j, k, c);
    const double duz = a.uedge[2](i, j, k + 1, c) - a.uedge[2](i, j, k, c);
    const double ugrad = ubar * dux / a.dx[0] + vbar * duy / a.dx[1] + wbar * duz / a.dx[2];
    double un = a.uold(i, j, k, c) - a.dt * ugrad + a.dt * a.force(i, j, k, c);
    const double w0_grad = dux / a.dx[0] * w0xb + duy / a.dx[1] * w0yb + duz / a.dx[2] * w0zb;
    un = un - a.dt * w0_grad;
    if (a.do_sponge) un = un * sp;
    a.unew(i, j, k, c) = un;
  }
}

__global__ void k_modify_scal_force_sphr(Geom g, DV force, DV s, DV um, DV vm, DV wm, DV s0c, Box3 vb, const double* divu,
                                         int type, bool fullform, int dlo0, int dlo1, int dlo2, int dhi0, int dhi1,
                                         int dhi2) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double x = pos(g, 0, i, true), y = pos(g, 1, j, true), z = pos(g, 2, k, true);
  const double divu_cart = interp_cc(g, type, divu, radius_of(x, y, z));
  const double divumac = (um(i + 1, j, k) - um(i, j, k)) / g.dx[0] + (vm(i, j + 1, k) - vm(i, j, k)) / g.dx[1] +
                         (wm(i, j, k + 1) - wm(i, j, k)) / g.dx[2];
  double f = force(i, j, k);
  if (fullform) {
    f = f - s(i, j, k) * (divumac + divu_cart);
  } else {
    const double c = s0c(i, j, k);
    const double s0_xhi = (i < dhi0) ? 0.5 * (c + s0c(i + 1, j, k)) : c;
    const double s0_xlo = (i > dlo0) ? 0.5 * (c + s0c(i - 1, j, k)) : c;
    const double s0_yhi = (j < dhi1) ? 0.5 * (c + s0c(i, j + 1, k)) : c;
    const double s0_ylo = (j > dlo1) ? 0.5 * (c + s0c(i, j - 1, k)) : c;
    const double s0_zhi = (k < dhi2) ? 0.5 * (c + s0c(i, j, k + 1)) : c;
    const double s0_zlo = (k > dlo2) ? 0.5 * (c + s0c(i, j, k - 1)) : c;
    const double divs0u = (um(i + 1, j, k) * s0_xhi - um(i, j, k) * s0_xlo) / g.dx[0] +
                          (vm(i, j + 1, k) * s0_yhi - vm(i, j, k) * s0_ylo) / g.dx[1] +
                          (wm(i, j, k + 1) * s0_zhi - wm(i, j, k) * s0_zlo) / g.dx[2];
    f = f - divs0u - (s(i, j, k) - c) * (divumac + divu_cart);
  }
  force(i, j, k) = f;
}

__global__ void k_pert_form_sphr(Geom g, DV s, Box3 vb, const double* s0, int type, int mult) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const double x = pos(g, 0, ix[0], true), y = pos(g, 1, ix[1], true), z = pos(g, 2, ix[2], true);
  const double v = interp_cc(g, type, s0, radius_of(x, y, z));
  s(ix[0], ix[1], ix[2]) = s(ix[0], ix[1], ix[2]) + mult * v;
}

// make_normal_3d_sphr (fill_3d_data.f90:1308)
__global__ void k_make_normal(Geom g, DV n, Box3 gb) {
  int ix[3];
  if (!decode3(gb, ix)) return;
  const double x = pos(g, 0, ix[0], true), y = pos(g, 1, ix[1], true), z = pos(g, 2, ix[2], true);
  const double radius = sqrt(x * x + y * y + z * z);
  n(ix[0], ix[1], ix[2], 0) = x * (1.0 / radius);
  n(ix[0], ix[1], ix[2], 1) = y * (1.0 / radius);
  n(ix[0], ix[1], ix[2], 2) = z * (1.0 / radius);
}

// mk_vel_force_3d_sphr (mkforce.f90:484): buoyancy with the radial gravity vector, pressure gradient, Coriolis and
// centrifugal terms about the z axis, w0 force, and the utilde . grad w0 term (:617-640); expression order of the source
struct VelForceSphr {
  Geom g;
  DV force, uold, gpi, rho, normal, w0_cart, gradw0, w0f, rho0c, gravc, ue[3], w0m[2];
  Box3 vb;
  bool is_final, add_utilde;
  double omega, rho_cut;
};
__global__ void k_vel_force_sphr(VelForceSphr a) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double yy = a.g.prob_lo[1] + ((double)j + 0.5) * a.g.dx[1] - a.g.center[1];
  const double xx = a.g.prob_lo[0] + ((double)i + 0.5) * a.g.dx[0] - a.g.center[0];
  const double rho = a.rho(i, j, k);
  double rhopert = rho - a.rho0c(i, j, k, 0);
  const bool outside = rho < a.rho_cut;
  if (outside) rhopert = 0.0;
  const double omega = a.omega;
  double cen[3] = {-omega * omega * xx, -omega * omega * yy, 0.0};
  if (outside) cen[0] = cen[1] = cen[2] = 0.0;
  double cor[3];
  if (a.is_final) {
    cor[0] = -(2.0 * omega * 0.5 * (a.ue[1](i, j, k) + a.w0m[1](i, j, k) + a.ue[1](i, j + 1, k) + a.w0m[1](i, j + 1, k)));
    cor[1] = 2.0 * omega * 0.5 * (a.ue[0](i, j, k) + a.w0m[0](i, j, k) + a.ue[0](i + 1, j, k) + a.w0m[0](i + 1, j, k));
    cor[2] = 0.0;
  } else {
    cor[0] = -(2.0 * omega * (a.uold(i, j, k, 1) + a.w0_cart(i, j, k, 1)));
    cor[1] = 2.0 * omega * (a.uold(i, j, k, 0) + a.w0_cart(i, j, k, 0));
    cor[2] = 0.0;
  }
  double f[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    f[c] = -cor[c] - cen[c] + (rhopert * a.gravc(i, j, k, c) - a.gpi(i, j, k, c)) / rho - a.w0f(i, j, k, c);
  if (a.add_utilde) {
    const double ut = 0.5 * (a.ue[0](i, j, k) + a.ue[0](i + 1, j, k)) * a.normal(i, j, k, 0) +
                      0.5 * (a.ue[1](i, j, k) + a.ue[1](i, j + 1, k)) * a.normal(i, j, k, 1) +
                      0.5 * (a.ue[2](i, j, k) + a.ue[2](i, j, k + 1)) * a.normal(i, j, k, 2);
    const double gw = a.gradw0(i, j, k, 0);
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] = f[c] - ut * gw * a.normal(i, j, k, c);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) a.force(i, j, k, c) = f[c];
}

// mkrhohforce_3d_sphr (mkscalforce.f90:388): u.grad p0 = div(u p0) - p0 div(u) with p0 on the faces and on the cell
// centres, + psi (cell centres) where the reference adds it, + thermal
struct RhohForceSphr {
  DV force, um, vm, wm, thermal, p0c, p0m[3], psic;
  Box3 vb;
  double dx[3];
  bool add_psi, add_thermal;
};
__global__ void k_rhoh_force_sphr(RhohForceSphr a) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double divup = (a.um(i + 1, j, k) * a.p0m[0](i + 1, j, k) - a.um(i, j, k) * a.p0m[0](i, j, k)) / a.dx[0] +
                       (a.vm(i, j + 1, k) * a.p0m[1](i, j + 1, k) - a.vm(i, j, k) * a.p0m[1](i, j, k)) / a.dx[1] +
                       (a.wm(i, j, k + 1) * a.p0m[2](i, j, k + 1) - a.wm(i, j, k) * a.p0m[2](i, j, k)) / a.dx[2];
  const double p0divu = ((a.um(i + 1, j, k) - a.um(i, j, k)) / a.dx[0] + (a.vm(i, j + 1, k) - a.vm(i, j, k)) / a.dx[1] +
                         (a.wm(i, j, k + 1) - a.wm(i, j, k)) / a.dx[2]) * a.p0c(i, j, k);
  double f = divup - p0divu;
  if (a.add_psi) f = f + a.psic(i, j, k);
  if (a.add_thermal) f = f + a.thermal(i, j, k);
  a.force(i, j, k) = f;
}

Box3 mac_box(const int* lo, const int* hi, int d) {
  Box3 b;
  for (int q = 0; q < 3; ++q) {
    b.lo[q] = lo[q] - 1;
    b.hi[q] = hi[q] + 1 + (q == d ? 1 : 0);
  }
  return b;
}

}  // namespace

static int g_sphr_fast = 1;
void sphr_set_fast(int on) { g_sphr_fast = on != 0; }
Geom make_geom(const mgpu_params& P, const mgpu_geom& g) {
  if (P.dm != 3) throw Error("spherical geometry is 3-D only");
  if (g.nr_fine < 3) throw Error("spherical geometry: nr_fine must be at least 3");
  if (!g.r_cc_loc || !g.r_edge_loc) throw Error("spherical geometry: r_cc_loc / r_edge_loc missing");
  Geom d;
  for (int q = 0; q < 3; ++q) {
    d.center[q] = g.center[q];
    d.prob_lo[q] = g.prob_lo[q];
    d.dx[q] = P.dx[q];
  }
  d.dr = g.dr;
  d.rdr = 1.0 / g.dr;
  d.fast = g_sphr_fast;
  d.nr_fine = g.nr_fine;
  d.r_cc_loc = upload_small(g.r_cc_loc, (size_t)g.nr_fine);
  d.r_edge_loc = upload_small(g.r_edge_loc, (size_t)g.nr_fine + 1);
  return d;
}

void put_1d_array_on_cart_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const double* s0_dev,
                              const DV& cart, bool edge_in, bool vec, const int* lo, const int* hi) {
  const int type = edge_in ? g.w0_interp_type : g.s0_interp_type;
  if (type < 1 || type > 3)
    throw Error(edge_in ? "Error: w0_interp_type not defined" : "Error: s0_interp_type not defined");
  if (cart.nc < (vec ? 3 : 1)) throw Error("put_1d_array_on_cart: s0_cart has too few components");
  Box3 vb = grown(lo, hi, 3, 0);
  MGPU_TIMED(TAG_GLUE, (k_put1d<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(gd, cart, vb, s0_dev, edge_in, vec, type)));
}

void make_mac_dev(const mgpu_geom& g, const Geom& gd, const double* s0_dev, DV* mac, const DV* cart, int kind,
                  const int* lo, const int* hi) {
  const int type = (kind == 0) ? g.w0mac_interp_type : g.s0mac_interp_type;
  if (type < 1 || type > (kind == 0 ? 4 : 3))
    throw Error(kind == 0 ? "Error: w0mac_interp_type not defined" : "Error: s0mac_interp_type not defined");
  if (type == 1 && !cart) throw Error("make_w0mac/make_s0mac: interp type 1 needs the cell-centred array");
  for (int d = 0; d < 3; ++d) {
    Box3 fb = mac_box(lo, hi, d);
    MGPU_TIMED(TAG_GLUE, (k_make_mac<<<grid3(fb, 256), block3(fb, 256), 0, ctx().stream>>>(gd, mac[d], cart ? *cart : mac[d], fb,
                                                                                       s0_dev, d, kind, type)));
  }
}

// utot = umac + w0mac on the valid faces and umac itself on the ghost faces (what addw0 leaves there), the whole fab
// extent in one pass: the FAST spherical episodes predict with this copy and leave umac alone
__global__ void k_sum_faces_sphr(DV ut, DV um, DV wm, Box3 all, Box3 fb) {
  int ix[3];
  if (!decode3(all, ix)) return;
  const bool in = ix[0] >= fb.lo[0] && ix[0] <= fb.hi[0] && ix[1] >= fb.lo[1] && ix[1] <= fb.hi[1] && ix[2] >= fb.lo[2] &&
                  ix[2] <= fb.hi[2];
  const double u = um(ix[0], ix[1], ix[2]);
  ut(ix[0], ix[1], ix[2]) = in ? u + 1.0 * wm(ix[0], ix[1], ix[2]) : u;
}
void sum_faces_sphr_dev(DV* utot, const DV* umac, const DV* w0mac, const int* lo, const int* hi) {
  for (int d = 0; d < 3; ++d) {
    Box3 fb = grown(lo, hi, 3, 0), all;
    fb.hi[d] += 1;
    for (int q = 0; q < 3; ++q) {
      all.lo[q] = umac[d].lo[q];
      all.hi[q] = umac[d].lo[q] + umac[d].n[q] - 1;
    }
    MGPU_TIMED(TAG_GLUE, (k_sum_faces_sphr<<<grid3(all, 256), block3(all, 256), 0, ctx().stream>>>(utot[d], umac[d], w0mac[d], all, fb)));
  }
}

void addw0_sphr_dev(DV* umac, const DV* w0mac, double mult, const int* lo, const int* hi) {
  for (int d = 0; d < 3; ++d) {
    Box3 fb = grown(lo, hi, 3, 0);
    fb.hi[d] += 1;
    MGPU_TIMED(TAG_GLUE, (k_addw0_sphr<<<grid3(fb, 256), block3(fb, 256), 0, ctx().stream>>>(umac[d], w0mac[d], fb, mult)));
  }
}

void mk_rhoX_flux_sphr_dev(SphrFluxArgs& a, int startcomp, int endcomp) {
  for (int d = 0; d < 3; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    MGPU_TIMED(TAG_FLUX, (k_rhoX_flux_sphr<<<grid3(fb, 256), block3(fb, 256), 0, ctx().stream>>>(a, d, startcomp - 1, endcomp - 1)));
  }
}

void mk_rhoh_flux_sphr_dev(const mgpu_params& P, SphrFluxArgs& a) {
  const int ept = P.enthalpy_pred_type, spt = P.species_pred_type;
  const bool have_h = (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  const bool have_hprime = (ept == MGPU_PREDICT_HPRIME);
  if (have_hprime && spt == MGPU_PREDICT_RHO_AND_X)
    throw Error("ERROR: predict_rho_and_X and predict_hprime not supported together");  // mkflux.f90:1408
  if (have_hprime && spt == MGPU_PREDICT_RHOX) throw Error("ERROR: predict_rhoX and predict_hprime not supported together");
  const int mode = have_h ? 0 : (have_hprime ? 1 : (ept == MGPU_PREDICT_RHOH ? 2 : 3));
  for (int d = 0; d < 3; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    MGPU_TIMED(TAG_FLUX, (k_rhoh_flux_sphr<<<grid3(fb, 256), block3(fb, 256), 0, ctx().stream>>>(a, d, mode)));
  }
}

void update_velocity_sphr_dev(VelArgs& a, const DV* w0mac) {
  MGPU_TIMED(TAG_UPDATE, (k_update_vel_sphr<<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a, w0mac[0], w0mac[1], w0mac[2])));
}

void modify_scal_force_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const DV& force, const DV& s,
                                const DV* umac, const DV& s0_cart, const double* w0_host, int comp, bool fullform,
                                const int* lo, const int* hi) {
  const int type = g.s0_interp_type;
  if (type < 1 || type > 3) throw Error("Error: s0_interp_type not defined");
  // w0 contribution to div(U) on the radial grid (modify_scal_force.f90:293-297): O(nr) host work, as in the reference
  std::vector<double> divu(g.nr_fine);
  for (int r = 0; r < g.nr_fine; ++r)
    divu[r] = (g.r_edge_loc[r + 1] * g.r_edge_loc[r + 1] * w0_host[r + 1] - g.r_edge_loc[r] * g.r_edge_loc[r] * w0_host[r]) /
              (g.dr * (g.r_cc_loc[r] * g.r_cc_loc[r]));
  const double* divu_d = upload_small(divu.data(), divu.size());
  Box3 vb = grown(lo, hi, 3, 0);
  MGPU_TIMED(TAG_GLUE, (k_modify_scal_force_sphr<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
                           gd, force.comp(comp - 1), s.comp(comp - 1), umac[0], umac[1], umac[2], s0_cart, vb, divu_d, type,
                           fullform, P.domlo[0], P.domlo[1], P.domlo[2], P.domhi[0], P.domhi[1], P.domhi[2])));
}

void pert_form_sphr_dev(const mgpu_geom& g, const Geom& gd, const DV& s, const double* s0_dev, int comp, bool flag,
                        const int* lo, const int* hi) {
  const int type = g.s0_interp_type;
  if (type < 1 || type > 3) throw Error("Error: s0_interp_type not defined");
  Box3 vb = grown(lo, hi, 3, 0);
  MGPU_TIMED(TAG_GLUE, (k_pert_form_sphr<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(gd, s.comp(comp - 1), vb, s0_dev, type,
                                                                                           flag ? -1 : 1)));
}

void make_normal_dev(const Geom& gd, const DV& normal, const int* lo, const int* hi, int ng) {
  Box3 gb = grown(lo, hi, 3, ng);
  MGPU_TIMED(TAG_GLUE, (k_make_normal<<<grid3(gb, 256), block3(gb, 256), 0, ctx().stream>>>(gd, normal, gb)));
}

// the spherical branch of mk_vel_force (mkforce.f90:22): w0 and grad w0 on the cell centres (:92-127), rho0 and the
// gravity vector on the cell centres (:529-530), then the force on the valid cells.  Temporaries from the arena.
void mk_vel_force_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const DV& force, bool is_final,
                           const DV& uold, const DV* uedge, const double* w0_h, const DV* w0mac, const DV& gpi,
                           const DV& rho1, const DV& normal, const double* rho0_h, const double* grav_h,
                           const DV& w0_force_cart, const int* lo, const int* hi, bool add_utilde) {
  const int nr = g.nr_fine;
  const int z3[3] = {0, 0, 0};
  auto tmp = [&](int nc) {
    DV v = make_view(nullptr, lo, hi, 3, 0, z3, nc);
    v.p = arena_alloc((size_t)v.size());
    return v;
  };
  DV w0_cart = tmp(3), gradw0 = tmp(1), rho0c = tmp(1), gravc = tmp(3);
  set_dev(w0_cart.p, 0.0, w0_cart.size());
  set_dev(gradw0.p, 0.0, gradw0.size());
  if (P.evolve_base_state) {
    put_1d_array_on_cart_dev(P, g, gd, upload_small(w0_h, (size_t)nr + 1), w0_cart, true, true, lo, hi);
    if (add_utilde) {
      std::vector<double> gr(nr);
      for (int r = 0; r < nr; ++r) gr[r] = (w0_h[r + 1] - w0_h[r]) / g.dr;
      put_1d_array_on_cart_dev(P, g, gd, upload_small(gr.data(), (size_t)nr), gradw0, false, false, lo, hi);
    }
  }
  set_dev(force.p, 0.0, force.size());
  put_1d_array_on_cart_dev(P, g, gd, upload_small(rho0_h, (size_t)nr), rho0c, false, false, lo, hi);
  put_1d_array_on_cart_dev(P, g, gd, upload_small(grav_h, (size_t)nr), gravc, false, true, lo, hi);
  VelForceSphr a;
  a.g = gd;
  a.force = force; a.uold = uold; a.gpi = gpi; a.rho = rho1; a.normal = normal; a.w0_cart = w0_cart; a.gradw0 = gradw0;
  a.w0f = w0_force_cart; a.rho0c = rho0c; a.gravc = gravc;
  for (int d = 0; d < 3; ++d) a.ue[d] = uedge[d];
  a.w0m[0] = w0mac[0]; a.w0m[1] = w0mac[1];
  a.vb = grown(lo, hi, 3, 0);
  a.is_final = is_final; a.add_utilde = add_utilde;
  a.omega = P.omega;
  a.rho_cut = P.buoyancy_cutoff_factor * P.base_cutoff_density;
  MGPU_TIMED(TAG_GLUE, (k_vel_force_sphr<<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a)));
}

// force: single-component view of scal_force(rhoh_comp); p0c (one ghost cell, filled) and p0mac from the caller
void mkrhohforce_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const DV& force, bool is_prediction,
                          const DV& thermal, const DV* umac, const DV& p0c, const DV* p0mac, const double* psi_h,
                          bool add_thermal, const int* lo, const int* hi) {
  const int ept = P.enthalpy_pred_type;
  if (is_prediction && !(ept == MGPU_PREDICT_RHOHPRIME || ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH))
    throw Error("ERROR: should only call mkrhohforce when predicting rhoh', h, or rhoh");  // mkscalforce.f90:87-92
  RhohForceSphr a;
  a.force = force; a.um = umac[0]; a.vm = umac[1]; a.wm = umac[2]; a.thermal = thermal; a.p0c = p0c;
  for (int d = 0; d < 3; ++d) { a.p0m[d] = p0mac[d]; a.dx[d] = P.dx[d]; }
  a.vb = grown(lo, hi, 3, 0);
  a.add_psi = (is_prediction && ept == MGPU_PREDICT_H) || (is_prediction && ept == MGPU_PREDICT_RHOH) || !is_prediction;
  a.add_thermal = add_thermal;
  a.psic = p0c;
  if (a.add_psi) {
    const int z3[3] = {0, 0, 0};
    DV psic = make_view(nullptr, lo, hi, 3, 0, z3, 1);
    psic.p = arena_alloc((size_t)psic.size());
    put_1d_array_on_cart_dev(P, g, gd, upload_small(psi_h, (size_t)g.nr_fine), psic, false, false, lo, hi);
    a.psic = psic;
  }
  MGPU_TIMED(TAG_GLUE, (k_rhoh_force_sphr<<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a)));
}

}  // namespace mgpu
