// Streaming (pointwise / 1-stencil) kernels of the path, planar geometry:
//   mk_rhoX_flux   Source/mkflux.f90:272 (2-D), :370 (3-D cart)
//   mk_rhoh_flux   Source/mkflux.f90:920 (2-D), :1070 (3-D cart)
//   update_scal    Source/update_scal.f90:246 (2-D), :370 (3-D cart)
//   update_velocity Source/update_vel.f90:174 (2-D), :227 (3-D, spherical==0)
//   addw0          Source/addw0.f90:133,150
//   modify_scal_force Source/modify_scal_force.f90:163,206;  pert_form put_in_pert_form.f90:138,162
//   rhoX<->X       Source/convert_rhoX_to_X.f90:44-56
//   ghost fill     FBoxLib multifab_fill_boundary (periodic wrap) + Source/multifab_physbc.f90:150,329
// All are HBM-bound: one thread per zone/face, x-contiguous so every warp reads/writes whole lines.
#include "mgpu_halo.cuh"
#include <algorithm>

#include "mgpu_eos.cuh"
#include "mgpu_stream.cuh"

namespace mgpu {

// ------------------------------------------------------------------------------------------
__global__ void k_rhoX_flux(FluxArgs a, int d, int c) {
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[d] += 1;
  if (!decode3(fb, ix)) return;
  const int r = a.dm - 1;
  const int ir = ix[r];
  double rho0_edge, vel = a.umac[d](ix[0], ix[1], ix[2]);
  if (d != r) {
    rho0_edge = 0.5 * (a.rho0_old[ir] + a.rho0_new[ir]);
  } else {
    rho0_edge = 0.5 * (a.rho0_edge_old[ir] + a.rho0_edge_new[ir]);
    vel = vel + a.w0[ir];
  }
  const DV& se = a.sedge[d];
  const long o = se.off(ix[0], ix[1], ix[2]);
  double f;
  if (a.species_pred_type == MGPU_PREDICT_RHOPRIME_AND_X)
    f = vel * (rho0_edge + se.p[o + se.cs * a.rho]) * se.p[o + se.cs * c];
  else if (a.species_pred_type == MGPU_PREDICT_RHOX)
    f = vel * se.p[o + se.cs * c];
  else
    f = vel * se.p[o + se.cs * a.rho] * se.p[o + se.cs * c];
  a.sflux[d](ix[0], ix[1], ix[2], c) = f;
  if (d == r && a.evolve_base_state) {
    double e = a.eta(ix[0], ix[1], ix[2]);
    if (c >= a.spec0 && c <= a.spec0 + a.nspec - 1) e = e + f;
    if (c == a.spec0 + a.nspec - 1) e = e - a.w0[ir] * a.rho0_predicted_edge[ir];
    a.eta(ix[0], ix[1], ix[2]) = e;
  }
}

void mk_rhoX_flux_dev(const mgpu_params& P, FluxArgs& a, int startcomp, int endcomp) {
  Context& cx = ctx();
  for (int comp = startcomp; comp <= endcomp; ++comp)
    for (int d = 0; d < P.dm; ++d) {
      Box3 fb = a.vb;
      fb.hi[d] += 1;
      k_rhoX_flux<<<grid3(fb, 256), block3(fb, 256), 0, cx.stream>>>(a, d, comp - 1);
      MGPU_LAUNCH_CHECK();
    }
}

__global__ void k_rhoh_flux(FluxArgs a, int d, int mode) {
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[d] += 1;
  if (!decode3(fb, ix)) return;
  const int r = a.dm - 1;
  const int ir = ix[r];
  double rho0_edge, rhoh0_edge, vel = a.umac[d](ix[0], ix[1], ix[2]);
  if (d != r) {
    rho0_edge = 0.5 * (a.rho0_old[ir] + a.rho0_new[ir]);
    rhoh0_edge = 0.5 * (a.rhoh0_old[ir] + a.rhoh0_new[ir]);
  } else {
    rho0_edge = 0.5 * (a.rho0_edge_old[ir] + a.rho0_edge_new[ir]);
    rhoh0_edge = 0.5 * (a.rhoh0_edge_old[ir] + a.rhoh0_edge_new[ir]);
    vel = vel + a.w0[ir];
  }
  const DV& se = a.sedge[d];
  const long o = se.off(ix[0], ix[1], ix[2]);
  const double erho = se.p[o + se.cs * a.rho], erhoh = se.p[o + se.cs * a.rhoh];
  double f;
  if (mode == 0) {  // have_h
    if (a.species_pred_type == MGPU_PREDICT_RHOPRIME_AND_X)
      f = vel * (rho0_edge + erho) * erhoh;
    else
      f = vel * erho * erhoh;
  } else if (mode == 1) {  // have_rhoh
    f = vel * erhoh;
  } else {  // (rho h)'
    f = vel * (rhoh0_edge + erhoh);
  }
  a.sflux[d](ix[0], ix[1], ix[2], a.rhoh) = f;
}

void mk_rhoh_flux_dev(const mgpu_params& P, FluxArgs& a) {
  Context& cx = ctx();
  const int ept = P.enthalpy_pred_type;
  const bool have_h = (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  if (ept == MGPU_PREDICT_HPRIME) throw Error("mk_rhoh_flux : predict_hprime not coded yet");  // mkflux.f90:1167
  const int mode = have_h ? 0 : (ept == MGPU_PREDICT_RHOH ? 1 : 2);
  for (int d = 0; d < P.dm; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    k_rhoh_flux<<<grid3(fb, 256), block3(fb, 256), 0, cx.stream>>>(a, d, mode);
    MGPU_LAUNCH_CHECK();
  }
}

// ------------------------------------------------------------------------------------------
__global__ void k_update_scal(UpdArgs a, int c) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  double divterm = (a.sflux[0](i + 1, j, k, c) - a.sflux[0](i, j, k, c)) / a.dx[0] +
                   (a.sflux[1](i, j + 1, k, c) - a.sflux[1](i, j, k, c)) / a.dx[1];
  if (a.dm == 3) divterm = divterm + (a.sflux[2](i, j, k + 1, c) - a.sflux[2](i, j, k, c)) / a.dx[2];
  a.snew(i, j, k, c) = a.sold(i, j, k, c) + a.dt * (-divterm + a.force(i, j, k, c));
}

__global__ void k_copy(double* dst, const double* src, long n) {
  long t = MGPU_TID;
  if (t < n) dst[t] = src[t];
}
__global__ void k_set(double* dst, double v, long n) {
  long t = MGPU_TID;
  if (t < n) dst[t] = v;
}

// density from the species updates + floor + negative-species redistribution, update_scal.f90:453-505
__global__ void k_update_rho(UpdArgs a, int c0, int c1, int rho, double bcd) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const long on = a.snew.off(i, j, k), oo = a.sold.off(i, j, k);
  double* sn = a.snew.p + on;
  const double* so = a.sold.p + oo;
  const long cn = a.snew.cs, co = a.sold.cs;
  bool neg = false;
  double r = sn[cn * rho];
  for (int c = c0; c <= c1; ++c) {
    r = r + (sn[cn * c] - so[co * c]);
    if (sn[cn * c] < 0.0) neg = true;
  }
  if (r < 0.5 * bcd) {
    for (int c = c0; c <= c1; ++c) sn[cn * c] = sn[cn * c] * 0.5 * bcd / r;
    r = 0.5 * bcd;
  }
  sn[cn * rho] = r;
  if (neg) {
    for (int c = c0; c <= c1; ++c) {
      if (sn[cn * c] < 0.0) {
        double delta = -sn[cn * c];
        double sumX = 0.0;
        for (int c2 = c0; c2 <= c1; ++c2)
          if (c2 != c && sn[cn * c2] >= 0.0) sumX = sumX + sn[cn * c2];
        for (int c2 = c0; c2 <= c1; ++c2)
          if (c2 != c && sn[cn * c2] >= 0.0) {
            double frac = sn[cn * c2] / sumX;
            sn[cn * c2] = sn[cn * c2] - frac * delta;
          }
        sn[cn * c] = 0.0;
      }
    }
  }
}

// zones of the valid box with rho <= cutoff (update_scal.f90:421: the zones whose rhoh the reference resets with the EOS)
__global__ void k_count_below(DV s, Box3 vb, int rho, double cutoff, unsigned long long* count) {
  int ix[3];
  bool hit = decode3(vb, ix) && s(ix[0], ix[1], ix[2], rho) <= cutoff;
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

void update_scal_dev(const mgpu_params& P, UpdArgs& a, int nstart, int nstop) {
  Context& cx = ctx();
  const bool eos_reset = P.do_eos_h_above_cutoff && nstart == P.rhoh_comp;
  if (eos_reset && !have_eos()) {
    // The reference recomputes rhoh from the EOS at (rho, p0, X) wherever rho <= base_cutoff_density
    // (update_scal.f90:421-447, default do_eos_h_above_cutoff = T).  Without an EOS (mgpu_set_eos): if any zone of this
    // box would take that branch the call fails instead of returning a different rhoh.
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(arena_alloc(1));
    MGPU_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), cx.stream));
    k_count_below<<<grid3(a.vb, 256), block3(a.vb, 256), 0, cx.stream>>>(a.snew, a.vb, P.rho_comp - 1, P.base_cutoff_density, cnt);
    MGPU_LAUNCH_CHECK();
    unsigned long long h = 0;
    MGPU_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, cx.stream));
    MGPU_CUDA(cudaStreamSynchronize(cx.stream));
    if (h > 0)
      throw Error("update_scal: " + std::to_string(h) + " zone(s) have rho <= base_cutoff_density: the EOS reset of rhoh "
                  "(update_scal.f90:421-447) needs an EOS -- call mgpu_set_eos (gamma_law_general), or set "
                  "do_eos_h_above_cutoff = F and apply the reset on the host (INTEGRATION.md)");
  }
  const long nv = a.vb.npts();
  for (int comp = nstart; comp <= nstop; ++comp) {
    k_update_scal<<<grid3(a.vb, 256), block3(a.vb, 256), 0, cx.stream>>>(a, comp - 1);
    MGPU_LAUNCH_CHECK();
  }
  if (eos_reset && have_eos())
    update_scal_eos_dev(P, a.sold, a.snew, a.p0_new, a.have_p0_new_cart ? &a.p0_new_cart : nullptr, a.vb);
  if (nstart == P.spec_comp && nstop == P.spec_comp + P.nspec - 1) {
    const int rho = P.rho_comp - 1;
    if (a.snew.cs != a.sold.cs) throw Error("update_scal: sold and snew must have the same ghost width");
    k_copy<<<nblocks(a.snew.cs, 256), 256, 0, cx.stream>>>(a.snew.p + a.snew.cs * rho, a.sold.p + a.sold.cs * rho,
                                                          a.snew.cs);
    MGPU_LAUNCH_CHECK();
    k_update_rho<<<grid3(a.vb, 256), block3(a.vb, 256), 0, cx.stream>>>(a, nstart - 1, nstop - 1, rho, P.base_cutoff_density);
    MGPU_LAUNCH_CHECK();
  }
}

// dst comp = sum of ncomp comps starting at c0 (multifab_copy_c + multifab_plus_plus_c, density_advance.f90:204-213)
__global__ void k_sum_comps(DV a, int dst, int c0, int ncomp) {
  long t = MGPU_TID;
  if (t >= a.cs) return;
  double r = a.p[t + a.cs * c0];
  for (int n = 1; n < ncomp; ++n) r = r + a.p[t + a.cs * (c0 + n)];
  a.p[t + a.cs * dst] = r;
}
void sum_comps_dev(const DV& a, int dst, int c0, int ncomp) {
  k_sum_comps<<<nblocks(a.cs, 256), 256, 0, ctx().stream>>>(a, dst, c0, ncomp);
  MGPU_LAUNCH_CHECK();
}

void set_dev(double* p, double v, long n) {
  if (v == 0.0) {  // all-zero bit pattern: the copy engine's memset runs at full write bandwidth
    MGPU_CUDA(cudaMemsetAsync(p, 0, (size_t)n * sizeof(double), ctx().stream));  // not counted as a launch of ours
    return;
  }
  k_set<<<nblocks(n, 256), 256, 0, ctx().stream>>>(p, v, n);
  MGPU_LAUNCH_CHECK();
}
void copy_dev(double* dst, const double* src, long n) {
  k_copy<<<nblocks(n, 256), 256, 0, ctx().stream>>>(dst, src, n);
  MGPU_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// FAST: the divisions by dx as multiplications by reciprocals formed on the host (four fp64 divisions per component
// otherwise: the kernel then runs at the fp64 pipe's pace instead of HBM's); <= 1e-12 relative like the other FAST builds
template <bool FAST>
__global__ void __launch_bounds__(256) k_update_vel(VelArgs a) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const int dm = a.dm, r = dm - 1;
  double bar[3];
  bar[0] = 0.5 * (a.umac[0](i, j, k) + a.umac[0](i + 1, j, k));
  bar[1] = 0.5 * (a.umac[1](i, j, k) + a.umac[1](i, j + 1, k));
  bar[2] = (dm == 3) ? 0.5 * (a.umac[2](i, j, k) + a.umac[2](i, j, k + 1)) : 0.0;
  const int ir = ix[r];
  const double wbar = 0.5 * (a.w0[ir] + a.w0[ir + 1]);
  const double sp = a.do_sponge ? a.sponge(i, j, k) : 1.0;
  for (int n = 0; n < dm; ++n) {
    const double e0l = a.uedge[0](i, j, k, n), e0h = a.uedge[0](i + 1, j, k, n);
    const double e1l = a.uedge[1](i, j, k, n), e1h = a.uedge[1](i, j + 1, k, n);
    const double e2l = (dm == 3) ? a.uedge[2](i, j, k, n) : 0.0, e2h = (dm == 3) ? a.uedge[2](i, j, k + 1, n) : 0.0;
    double ugrad;
    if (FAST) ugrad = bar[0] * (e0h - e0l) * a.rdx[0] + bar[1] * (e1h - e1l) * a.rdx[1];
    else ugrad = bar[0] * (e0h - e0l) / a.dx[0] + bar[1] * (e1h - e1l) / a.dx[1];
    if (dm == 3) ugrad = ugrad + (FAST ? bar[2] * (e2h - e2l) * a.rdx[2] : bar[2] * (e2h - e2l) / a.dx[2]);
    double un = a.uold(i, j, k, n) - a.dt * ugrad + a.dt * a.force(i, j, k, n);
    const double hi_e = (r == 1) ? e1h : e2h;
    const double lo_e = (r == 1) ? e1l : e2l;
    if (FAST) un = un - a.dt * wbar * (hi_e - lo_e) * a.rdx[r];
    else un = un - a.dt * wbar * (hi_e - lo_e) / a.dx[r];
    if (a.do_sponge) un = un * sp;
    a.unew(i, j, k, n) = un;
  }
}
void update_velocity_dev(VelArgs& a, bool fast) {
  for (int d = 0; d < 3; ++d) a.rdx[d] = 1.0 / a.dx[d];
  if (fast) MGPU_TIMED(TAG_UPDATE, (k_update_vel<true><<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a)));
  else MGPU_TIMED(TAG_UPDATE, (k_update_vel<false><<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a)));
}

// ------------------------------------------------------------------------------------------
__global__ void k_addw0(DV wm, Box3 b, int r, const double* w0, double mult) {
  int ix[3];
  if (!decode3(b, ix)) return;
  wm(ix[0], ix[1], ix[2]) = wm(ix[0], ix[1], ix[2]) + mult * w0[ix[r]];
}
void addw0_dev(const mgpu_params& P, DV* umac, const double* w0_dev, double mult, const int* lo, const int* hi) {
  const int r = P.dm - 1;
  Box3 b = grown(lo, hi, P.dm, 1);
  b.lo[r] = lo[r];
  b.hi[r] = hi[r] + 1;
  k_addw0<<<grid3(b, 256), block3(b, 256), 0, ctx().stream>>>(umac[r], b, r, w0_dev, mult);
  MGPU_LAUNCH_CHECK();
}

template <bool FAST>
__global__ void k_modify_scal_force(DV force, DV s, DV u, DV v, DV w, Box3 vb, int dm, const double* s0,
                                    const double* s0_edge, const double* w0, double dx0, double dx1, double dx2,
                                    bool fullform, bool assign) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const int ir = ix[dm - 1];
  // FAST: dx0..dx2 hold 1/dx and the divisions become multiplications (last-bit differences)
#define MGPU_DIV(x, h) (FAST ? (x) * (h) : (x) / (h))
  double divu, divs0u, f = assign ? 0.0 : force(i, j, k);  // assign: the force is known to be zero on entry
  if (dm == 2) {
    divu = MGPU_DIV(u(i + 1, j, k) - u(i, j, k), dx0) + MGPU_DIV(v(i, j + 1, k) - v(i, j, k), dx1);
    divu = divu + MGPU_DIV(w0[ir + 1] - w0[ir], dx1);
    if (fullform) {
      f = f - s(i, j, k) * divu;
    } else {
      divs0u = MGPU_DIV(s0[ir] * (u(i + 1, j, k) - u(i, j, k)), dx0) +
               MGPU_DIV(v(i, j + 1, k) * s0_edge[ir + 1] - v(i, j, k) * s0_edge[ir], dx1);
      f = f - (s(i, j, k) - s0[ir]) * divu - divs0u;
    }
  } else {
    divu = MGPU_DIV(u(i + 1, j, k) - u(i, j, k), dx0) + MGPU_DIV(v(i, j + 1, k) - v(i, j, k), dx1) +
           MGPU_DIV(w(i, j, k + 1) - w(i, j, k), dx2);
    divu = divu + MGPU_DIV(w0[ir + 1] - w0[ir], dx2);
    if (fullform) {
      f = f - s(i, j, k) * divu;
    } else {
      divs0u = s0[ir] * (MGPU_DIV(u(i + 1, j, k) - u(i, j, k), dx0) + MGPU_DIV(v(i, j + 1, k) - v(i, j, k), dx1)) +
               MGPU_DIV(w(i, j, k + 1) * s0_edge[ir + 1] - w(i, j, k) * s0_edge[ir], dx2);
      f = f - (s(i, j, k) - s0[ir]) * divu - divs0u;
    }
  }
  force(i, j, k) = f;
#undef MGPU_DIV
}
void modify_scal_force_dev(const mgpu_params& P, const DV& force, const DV& s, const DV* umac, const double* s0,
                           const double* s0_edge, const double* w0, int comp, bool fullform, const int* lo,
                           const int* hi, bool fast, bool assign) {
  Box3 vb = grown(lo, hi, P.dm, 0);
  if (fast)
    k_modify_scal_force<true><<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
        force.comp(comp - 1), s.comp(comp - 1), umac[0], umac[1], P.dm == 3 ? umac[2] : umac[1], vb, P.dm, s0, s0_edge,
        w0, 1.0 / P.dx[0], 1.0 / P.dx[1], 1.0 / P.dx[2], fullform, assign);
  else
    k_modify_scal_force<false><<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
        force.comp(comp - 1), s.comp(comp - 1), umac[0], umac[1], P.dm == 3 ? umac[2] : umac[1], vb, P.dm, s0, s0_edge,
        w0, P.dx[0], P.dx[1], P.dx[2], fullform, assign);
  MGPU_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// mkrhohforce_2d / _3d (Source/mkscalforce.f90:249, :310)
__global__ void k_mkrhohforce(RhohForceArgs a) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int r = a.dm - 1, q = ix[r];
  double gradp0;
  if (q < a.cutoff_coord) {
    gradp0 = (0.5 * (a.rho0_1[q] + a.rho0_2[q])) * a.grav[q];
  } else if (q == a.nr - 1) {
    gradp0 = (0.5 * (a.p0_1[q] + a.p0_2[q]) - 0.5 * (a.p0_1[q - 1] + a.p0_2[q - 1])) / a.dr;
  } else {
    gradp0 = (0.5 * (a.p0_1[q + 1] + a.p0_2[q + 1]) - 0.5 * (a.p0_1[q] + a.p0_2[q])) / a.dr;
  }
  const long wo = a.wm.off(ix[0], ix[1], ix[2]);
  const double wadv = 0.5 * (a.wm.p[wo] + a.wm.p[wo + a.wm.stride(r)]);
  double v = wadv * gradp0;
  if (a.with_psi) v = v + a.psi[q];
  if (a.add_thermal) v = v + a.thermal(ix[0], ix[1], ix[2]);
  a.f(ix[0], ix[1], ix[2]) = v;
}
void mkrhohforce_dev(RhohForceArgs& a) {
  MGPU_TIMED(TAG_GLUE, (k_mkrhohforce<<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a)));
}

// mk_vel_force_2d / _3d_cart (Source/mkforce.f90:283, :342)
__global__ void k_mk_vel_force(VelForceArgs a) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const int dm = a.dm, r = dm - 1, q = ix[r];
  const double rho = a.rho(i, j, k);
  double rhopert = rho - a.rho0[q];
  if (rho < a.rho_cut) rhopert = 0.0;
  double fr;
  if (dm == 2) {
    a.force(i, j, k, 0) = -a.gpi(i, j, k, 0) / rho;
    fr = rhopert / rho * a.grav[q] - a.gpi(i, j, k, 1) / rho - a.w0_force[q];
  } else {
    const double omega = a.omega, sin_theta = a.sin_theta, cos_theta = a.cos_theta;
    const double cen0 = -(omega * omega) * a.rotation_radius * sin_theta * sin_theta;
    const double cen1 = 0.0;
    const double cen2 = (omega * omega) * a.rotation_radius * cos_theta * sin_theta - (omega * omega) * a.rotation_radius;
    double cor0, cor1, cor2;
    if (a.is_final_update) {
      cor0 = -2.0 * omega * 0.5 * (a.uedge[1](i, j, k) + a.uedge[1](i, j + 1, k)) * cos_theta;
      cor1 = 2.0 * omega * (0.5 * (a.uedge[2](i, j, k) + a.w0[k] + a.uedge[2](i, j, k + 1) + a.w0[k + 1]) * sin_theta +
                            0.5 * (a.uedge[0](i, j, k) + a.uedge[0](i + 1, j, k)) * cos_theta);
      cor2 = -2.0 * omega * 0.5 * (a.uedge[1](i, j, k) + a.uedge[1](i, j + 1, k)) * sin_theta;
    } else {
      cor0 = -2.0 * omega * a.uold(i, j, k, 1) * cos_theta;
      cor1 = 2.0 * omega * ((a.uold(i, j, k, 2) + 0.5 * (a.w0[k] + a.w0[k + 1])) * sin_theta + a.uold(i, j, k, 0) * cos_theta);
      cor2 = -2.0 * omega * a.uold(i, j, k, 1) * sin_theta;
    }
    a.force(i, j, k, 0) = -cor0 - cen0 - a.gpi(i, j, k, 0) / rho;
    a.force(i, j, k, 1) = -cor1 - cen1 - a.gpi(i, j, k, 1) / rho;
    fr = -cor2 - cen2 + (rhopert * a.grav[k] - a.gpi(i, j, k, 2)) / rho - a.w0_force[k];
  }
  if (a.add_utilde && q > -1 && q < a.nr) {
    const DV& we = a.uedge[r];
    const long o = we.off(i, j, k);
    fr = fr - (we.p[o + we.stride(r)] + we.p[o]) * (a.w0[q + 1] - a.w0[q]) / (2.0 * a.dr);
  }
  a.force(i, j, k, r) = fr;
}
void mk_vel_force_dev(VelForceArgs& a) {
  set_dev(a.force.p, 0.0, a.force.size());  // vel_force = ZERO (mkforce.f90:302 / :371)
  MGPU_TIMED(TAG_GLUE, (k_mk_vel_force<<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(a)));
}

__global__ void k_radial_cell_avg(DV u, Box3 b, int r, const double* w0) {
  int ix[3];
  if (!decode3(b, ix)) return;
  u(ix[0], ix[1], ix[2]) = 0.5 * (w0[ix[r]] + w0[ix[r] + 1]);
}
void radial_cell_avg_dev(const mgpu_params& P, const DV& ufull, const double* w0_dev, const int* lo, const int* hi) {
  const int r = P.dm - 1;
  set_dev(ufull.p, 0.0, ufull.size());  // fill_3d_data.f90:164 (s0_cart = ZERO)
  Box3 b;
  for (int d = 0; d < 3; ++d) { b.lo[d] = ufull.lo[d]; b.hi[d] = ufull.lo[d] + ufull.n[d] - 1; }
  b.lo[r] = lo[r];
  b.hi[r] = hi[r];
  k_radial_cell_avg<<<grid3(b, 256), block3(b, 256), 0, ctx().stream>>>(ufull.comp(r), b, r, w0_dev);
  MGPU_LAUNCH_CHECK();
}
__global__ void k_add(double* dst, const double* src, long n) {
  long t = MGPU_TID;
  if (t < n) dst[t] = dst[t] + src[t];
}
__global__ void k_recip(double* dst, const double* src, long n) {
  long t = MGPU_TID;
  if (t < n) dst[t] = 1.0 / src[t];
}
void recip_dev(double* dst, const double* src, long n) {
  MGPU_TIMED(TAG_GLUE, (k_recip<<<nblocks(n, 256), 256, 0, ctx().stream>>>(dst, src, n)));
}
void add_dev(double* dst, const double* src, long n) {
  k_add<<<nblocks(n, 256), 256, 0, ctx().stream>>>(dst, src, n);
  MGPU_LAUNCH_CHECK();
}

// op: 0 = a/b, 1 = a*b (b = another component), 2 = a + mult*base(ir)
__global__ void k_pointwise(DV a, DV b, Box3 vb, int op, int r, const double* base, double mult) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  double& x = a(ix[0], ix[1], ix[2]);
  if (op == 0) x = x / b(ix[0], ix[1], ix[2]);
  else if (op == 1) x = x * b(ix[0], ix[1], ix[2]);
  else x = x + mult * base[ix[r]];
}
void convert_rhoX_to_X_dev(const mgpu_params& P, const DV& s, bool flag, const int* lo, const int* hi) {
  Box3 vb = grown(lo, hi, P.dm, 0);
  for (int n = 0; n < P.nspec; ++n) {
    k_pointwise<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(s.comp(P.spec_comp - 1 + n), s.comp(P.rho_comp - 1),
                                                                  vb, flag ? 0 : 1, 0, nullptr, 0.0);
    MGPU_LAUNCH_CHECK();
  }
}
// rhoX <-> X of every species and rho <-> rho' in one pass over the valid cells (density_advance.f90:160-171 and
// :229-240): the same statements as convert_rhoX_to_X_dev + put_in_pert_form_dev, rho read once per zone.
//   forward : X = rhoX / rho (rho still the full density), then rho' = rho - rho0
//   backward: rho = rho' + rho0, then rhoX = X * rho (the restored density)
__global__ void k_species_form(DV s, Box3 vb, int rho, int spec0, int nspec, bool convert, bool pert, bool forward,
                               int r, const double* base) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const long o = s.off(ix[0], ix[1], ix[2]);
  double* pr = s.p + o + s.cs * rho;
  double rv = *pr;
  if (forward) {
    if (convert)
      for (int n = 0; n < nspec; ++n) {
        double* px = s.p + o + s.cs * (spec0 + n);
        *px = *px / rv;
      }
    if (pert) *pr = rv + (-1.0) * base[ix[r]];
  } else {
    if (pert) {
      rv = rv + 1.0 * base[ix[r]];
      *pr = rv;
    }
    if (convert)
      for (int n = 0; n < nspec; ++n) {
        double* px = s.p + o + s.cs * (spec0 + n);
        *px = *px * rv;
      }
  }
}
void species_form_dev(const mgpu_params& P, const DV& s, const double* base_dev, bool convert, bool pert, bool forward,
                      const int* lo, const int* hi) {
  if (!convert && !pert) return;
  Box3 vb = grown(lo, hi, P.dm, 0);
  MGPU_TIMED(TAG_GLUE, (k_species_form<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
                           s, vb, P.rho_comp - 1, P.spec_comp - 1, P.nspec, convert, pert, forward, P.dm - 1, base_dev)));
}
void comp_muldiv_dev(const mgpu_params& P, const DV& a, int ca, const DV& b, int cb, int op, int g, const int* lo,
                     const int* hi) {
  Box3 vb = grown(lo, hi, P.dm, g);
  k_pointwise<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(a.comp(ca), b.comp(cb), vb, op, 0, nullptr, 0.0);
  MGPU_LAUNCH_CHECK();
}
void put_in_pert_form_dev(const mgpu_params& P, const DV& s, const double* base_dev, int comp, bool flag,
                          const int* lo, const int* hi) {
  Box3 vb = grown(lo, hi, P.dm, 0);
  k_pointwise<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(s.comp(comp - 1), s.comp(comp - 1), vb, 2, P.dm - 1,
                                                                base_dev, flag ? -1.0 : 1.0);
  MGPU_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// ghost fill: periodic wraps (k_wrap_all), then physbc in the reference's order.
#define WRAP_MAX 12  // (fab, component) pairs per wrap launch: blockIdx.y selects the pair

// All periodic directions in ONE launch: a ghost cell takes the value of the cell obtained by wrapping every periodic
// coordinate that lies in its ghost range -- the same values as wrapping x, then y, then z (edges and corners
// included), without the ordering.  The ghost shell is split into disjoint regions (blockIdx.z): region r holds the
// cells that are ghost in direction r and not ghost in any periodic direction before r.
struct WrapAll {
  int n;
  DV a[WRAP_MAX];
  int lo[WRAP_MAX][3], hi[WRAP_MAX][3], nodal[WRAP_MAX][3], ng[WRAP_MAX];
  int per[3];  // wrap in this direction (periodic and not the slab direction of a multi-rank run)
};
__global__ void k_wrap_all(const __grid_constant__ WrapAll w) {
  const int e = blockIdx.y, reg = blockIdx.z;
  if (!w.per[reg]) return;
  const DV& a = w.a[e];
  const int ng = w.ng[e];
  int cnt[3], base[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d == reg) {
      cnt[d] = 2 * ng;
      base[d] = 0;
    } else if (d < reg && w.per[d]) {  // valid range only (its ghost cells belong to region d)
      base[d] = w.lo[e][d];
      cnt[d] = w.hi[e][d] + w.nodal[e][d] - w.lo[e][d] + 1;
    } else {  // the whole allocated extent
      base[d] = a.lo[d];
      cnt[d] = a.n[d];
    }
  }
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)cnt[0] * (unsigned)cnt[1] * (unsigned)cnt[2]) return;
  const unsigned q1 = t / (unsigned)cnt[0];
  int c[3];
  c[0] = (int)(t - q1 * (unsigned)cnt[0]);
  const unsigned q2 = q1 / (unsigned)cnt[1];
  c[1] = (int)(q1 - q2 * (unsigned)cnt[1]);
  c[2] = (int)q2;
  int dst[3], src[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int lo = w.lo[e][d], top = w.hi[e][d] + w.nodal[e][d], n = w.hi[e][d] - w.lo[e][d] + 1;
    if (d == reg) dst[d] = (c[d] < ng) ? lo - 1 - c[d] : top + 1 + (c[d] - ng);
    else dst[d] = base[d] + c[d];
    src[d] = dst[d];
    if (w.per[d]) {
      if (dst[d] < lo) src[d] = dst[d] + n;
      else if (dst[d] > top) src[d] = dst[d] - n;
    }
  }
  a(dst[0], dst[1], dst[2]) = a(src[0], src[1], src[2]);
}

__global__ void k_physbc(DV s, Box3 tb, int d, int side, int bc, int lo, int hi, int ng) {
  int ix[3];
  if (!decode32(tb, ix)) return;  // tb: d collapsed to a single index
  const int e = (side == 0) ? lo : hi;
  const int sg = (side == 0) ? -1 : 1;
  const long st = s.stride(d);
  int ib[3] = {ix[0], ix[1], ix[2]};
  ib[d] = e;
  double* q = s.p + s.off(ib[0], ib[1], ib[2]);  // first valid cell next to the wall
  if (bc == MGPU_BC_EXT_DIR) {
    for (int g = 1; g <= ng; ++g) q[sg * g * st] = 0.0;
  } else if (bc == MGPU_BC_FOEXTRAP) {
    for (int g = 1; g <= ng; ++g) q[sg * g * st] = q[0];
  } else if (bc == MGPU_BC_HOEXTRAP) {
    const double v = (15.0 * q[0] - 10.0 * q[-sg * st] + 3.0 * q[-2 * sg * st]) * 0.125;
    for (int g = 1; g <= ng; ++g) q[sg * g * st] = v;
  } else if (bc == MGPU_BC_REFLECT_EVEN) {
    for (int g = 1; g <= ng; ++g) q[sg * g * st] = q[-sg * (g - 1) * st];
  } else if (bc == MGPU_BC_REFLECT_ODD) {
    for (int g = 1; g <= ng; ++g) q[sg * g * st] = -q[-sg * (g - 1) * st];
  }
}

// One ghost fill = (1) slab-direction exchange with the neighbouring ranks (NCCL), (2) in-box periodic wraps and
// physical BCs.  Inside a FillBatch the requests are only recorded; the batch then issues ALL exchanges as one NCCL
// group (one launch, one rendezvous with each neighbour instead of one per field) followed by the local parts in
// the recorded order.  Requests of one batch must touch disjoint (fab, component) sets.
namespace {
struct FillReq {
  mgpu_params P;
  DV sfull;
  int lo[3], hi[3], ng, nodal[3];
  bool has_nodal;
  int scomp, bccomp, ncomp;
  const int* adv_bc;
  int pmask[3];
  bool same_boundary;
};
bool g_batching = false;
std::vector<FillReq> g_batch;
cudaStream_t g_fill_stream = nullptr;  // non-null: the exchanges go to this stream (fill_batch_exchange_async)

bool fill_exchange(const FillReq& r) {
  return halo_exchange_dev(r.P, r.sfull, r.lo, r.hi, r.ng, r.has_nodal ? r.nodal : nullptr, r.scomp - 1, r.ncomp, r.pmask,
                           g_fill_stream ? g_fill_stream : ctx().stream);
}

// periodic wraps of a set of requests: every periodic direction in one launch (k_wrap_all) for all the (fab,
// component) pairs that wrap in the same directions
void fill_wraps(const FillReq* reqs, const char* slab, size_t nreq) {
  Context& cx = ctx();
  for (int sig = 1; sig < 8; ++sig) {  // bit d: wrap in direction d
    WrapAll w;
    w.n = 0;
    for (int d = 0; d < 3; ++d) w.per[d] = (sig >> d) & 1;
    unsigned maxblocks = 0;
    auto flush = [&]() {
      if (w.n == 0) return;
      k_wrap_all<<<dim3(maxblocks, w.n, 3), 256, 0, cx.stream>>>(w);
      MGPU_LAUNCH_CHECK();
      w.n = 0;
      maxblocks = 0;
    };
    for (size_t q = 0; q < nreq; ++q) {
      const FillReq& r = reqs[q];
      int mine = 0;
      for (int d = 0; d < r.P.dm; ++d)
        if (r.pmask[d] && !(slab[q] && d == r.P.dm - 1)) mine |= 1 << d;
      if (mine != sig) continue;
      for (int c = 0; c < r.ncomp; ++c) {
        if (w.n == WRAP_MAX) flush();
        const int e = w.n++;
        w.a[e] = r.sfull.comp(r.scomp - 1 + c);
        w.ng[e] = r.ng;
        long big = 0;
        for (int d = 0; d < 3; ++d) {
          w.lo[e][d] = r.lo[d];
          w.hi[e][d] = r.hi[d];
          w.nodal[e][d] = r.has_nodal ? r.nodal[d] : 0;
        }
        for (int reg = 0; reg < 3; ++reg) {  // the largest region decides the grid
          if (!w.per[reg]) continue;
          long npts = 2L * r.ng;
          for (int d = 0; d < 3; ++d)
            if (d != reg) npts *= w.a[e].n[d];
          big = std::max(big, npts);
        }
        maxblocks = std::max(maxblocks, nblocks(big, 256));
      }
    }
    flush();
  }
}

void fill_physbc(const FillReq& r) {
  Context& cx = ctx();
  const mgpu_params& P = r.P;
  const int dm = P.dm, ng = r.ng, ncomp = r.ncomp, scomp = r.scomp;
  const int *lo = r.lo, *hi = r.hi, *adv_bc = r.adv_bc;
  const int* nodal = r.has_nodal ? r.nodal : nullptr;
  const DV& sfull = r.sfull;
  const bool is_nodal = nodal && (nodal[0] || nodal[1] || nodal[2]);
  for (int n = 0; n < ncomp; ++n) {
    DV s = sfull.comp(scomp - 1 + n);
    if (is_nodal) continue;  // multifab_physbc_edgevel (FBoxLib) is left to the caller
    const int bcc = r.same_boundary ? r.bccomp : r.bccomp + n;
    int bc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int d = 0; d < dm; ++d) {
      bc[d][0] = adv_bc[d + dm * (0 + 2 * (bcc - 1))];
      bc[d][1] = adv_bc[d + dm * (1 + 2 * (bcc - 1))];
    }
    for (int d = 0; d < dm; ++d)
      for (int side = 0; side < 2; ++side) {
        const int b = bc[d][side];
        if (b == MGPU_BC_INTERIOR || b == MGPU_BC_PERIODIC) continue;
        if (b != MGPU_BC_EXT_DIR && b != MGPU_BC_FOEXTRAP && b != MGPU_BC_HOEXTRAP && b != MGPU_BC_REFLECT_EVEN &&
            b != MGPU_BC_REFLECT_ODD)
          throw Error("physbc: bc not yet supported");
        Box3 tb;
        for (int t = 0; t < 3; ++t) {
          if (t >= dm || t == d) { tb.lo[t] = tb.hi[t] = 0; continue; }
          int glo = ng, ghi = ng;
          if (t > d && b != MGPU_BC_EXT_DIR) {
            glo = (bc[t][0] == MGPU_BC_INTERIOR) ? ng : 0;
            ghi = (bc[t][1] == MGPU_BC_INTERIOR) ? ng : 0;
          }
          tb.lo[t] = lo[t] - glo;
          tb.hi[t] = hi[t] + ghi;
        }
        k_physbc<<<nblocks(tb.npts(), 256), 256, 0, cx.stream>>>(s, tb, d, side, b, lo[d], hi[d], ng);
        MGPU_LAUNCH_CHECK();
      }
  }
}
}  // namespace

static std::vector<char> g_batch_slab;
void fill_batch_begin() {
  if (g_batching) throw Error("mgpu: nested ghost-fill batch");
  g_batching = true;
  g_batch.clear();
  g_batch_slab.clear();
}
// Overlapped form of fill_batch_end: the exchanges of the batch go to the communication stream behind everything
// already enqueued on the compute stream (the kernels that produced the planes to send), the caller then enqueues
// work that touches neither the sent nor the received planes, and fill_batch_finish() makes the compute stream
// wait for the exchange before the in-box wraps / physical BCs.
// It may be called several times while a batch is open: each call sends the requests recorded since the last one.
void fill_batch_exchange_async() {
  if (!g_batching) throw Error("mgpu: no ghost-fill batch open");
  const size_t first = g_batch_slab.size();
  g_batch_slab.resize(g_batch.size(), 0);
  if (comm_size() > 1 && first < g_batch.size()) {
    cudaStream_t cs = comm_stream();
    MGPU_CUDA(cudaEventRecord(comm_event(0), ctx().stream));
    MGPU_CUDA(cudaStreamWaitEvent(cs, comm_event(0), 0));
    g_fill_stream = cs;
    halo_group_begin();
    try {
      for (size_t q = first; q < g_batch.size(); ++q) g_batch_slab[q] = fill_exchange(g_batch[q]) ? 1 : 0;
    } catch (...) {
      g_fill_stream = nullptr;
      halo_group_abort();
      throw;
    }
    halo_group_end();
    g_fill_stream = nullptr;
    MGPU_CUDA(cudaEventRecord(comm_event(1), cs));
  }
}
void fill_batch_finish() {
  if (!g_batching) return;
  g_batching = false;
  if (g_batch_slab.size() != g_batch.size()) throw Error("mgpu: ghost-fill batch finished with unsent requests");
  if (comm_size() > 1) MGPU_CUDA(cudaStreamWaitEvent(ctx().stream, comm_event(1), 0));
  fill_wraps(g_batch.data(), g_batch_slab.data(), g_batch.size());
  for (size_t q = 0; q < g_batch.size(); ++q) fill_physbc(g_batch[q]);
  g_batch.clear();
  g_batch_slab.clear();
}
void fill_batch_end() {
  if (!g_batching) return;
  g_batching = false;
  std::vector<char> slab(g_batch.size(), 0);
  // slab-partitioned domain: the slab-direction ghost planes come from the neighbouring ranks (NCCL), first, so
  // that the in-box wraps and physical BCs below also cover the received planes (corners come out right)
  halo_group_begin();
  try {
    for (size_t q = 0; q < g_batch.size(); ++q) slab[q] = fill_exchange(g_batch[q]) ? 1 : 0;
  } catch (...) {
    halo_group_abort();
    g_batch.clear();
    throw;
  }
  halo_group_end();
  fill_wraps(g_batch.data(), slab.data(), g_batch.size());
  for (size_t q = 0; q < g_batch.size(); ++q) fill_physbc(g_batch[q]);
  g_batch.clear();
}
void fill_batch_abort() {
  halo_group_abort();
  g_batching = false;
  g_batch_slab.clear();
  g_batch.clear();
}

// ---- several boxes per rank: FBoxLib multifab_fill_boundary + multifab_physbc for a list of fabs -----------------------
// Every ghost cell of box i that lies inside the valid region of a box j (j = i included) shifted by a periodic image
// of the domain takes that box's value; then the physical BCs of each box on the faces where it touches the domain
// boundary.  One launch copies every (destination, source, image) intersection of a component range.
struct BoxCopy {
  double* dst;
  const double* src;
  int lo[3], n[3];      // intersection in destination indices
  int dlo[3], dn[2];    // destination fab origin and extents (x, y)
  int slo[3], sn[2];    // source fab origin and extents
  int shift[3];         // source index = destination index - shift
  int vlo[3], vhi[3];   // destination valid region (skipped: only ghost cells are written)
  long dcs, scs;
  int ncomp;
};
__global__ void k_box_copies(const BoxCopy* __restrict__ cs, int ncopies) {
  const BoxCopy c = cs[blockIdx.y];
  const long npts = (long)c.n[0] * c.n[1] * c.n[2];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < npts; t += (long)gridDim.x * blockDim.x) {
    const int ix = c.lo[0] + (int)(t % c.n[0]);
    const long r = t / c.n[0];
    const int iy = c.lo[1] + (int)(r % c.n[1]), iz = c.lo[2] + (int)(r / c.n[1]);
    if (ix >= c.vlo[0] && ix <= c.vhi[0] && iy >= c.vlo[1] && iy <= c.vhi[1] && iz >= c.vlo[2] && iz <= c.vhi[2]) continue;
    const long od = (ix - c.dlo[0]) + (long)c.dn[0] * ((iy - c.dlo[1]) + (long)c.dn[1] * (iz - c.dlo[2]));
    const long os = (ix - c.shift[0] - c.slo[0]) +
                    (long)c.sn[0] * ((iy - c.shift[1] - c.slo[1]) + (long)c.sn[1] * (iz - c.shift[2] - c.slo[2]));
    for (int q = 0; q < c.ncomp; ++q) c.dst[od + c.dcs * q] = c.src[os + c.scs * q];
  }
}

void fill_boundary_mf_dev(const mgpu_params& P, int nfabs, const DV* fabs, const int* const* lo, const int* const* hi, int ng,
                          const int* nodal, int scomp, int bccomp, int ncomp, const int* const* adv_bc, const int* pmask,
                          bool same_boundary) {
  if (ng == 0) return;
  if (comm_size() > 1) throw Error("fill_boundary: several boxes per rank and several ranks at once are not built");
  const int dm = P.dm;
  std::vector<BoxCopy> copies;
  int nshift[3] = {0, 0, 0};
  for (int d = 0; d < dm; ++d) nshift[d] = pmask[d] ? 1 : 0;
  for (int i = 0; i < nfabs; ++i) {
    int glo[3] = {0, 0, 0}, ghi[3] = {0, 0, 0}, vlo[3] = {0, 0, 0}, vhi[3] = {0, 0, 0};
    for (int d = 0; d < dm; ++d) {
      vlo[d] = lo[i][d];
      vhi[d] = hi[i][d] + (nodal ? nodal[d] : 0);
      glo[d] = vlo[d] - ng;
      ghi[d] = vhi[d] + ng;
    }
    for (int j = 0; j < nfabs; ++j)
      for (int sz = -nshift[2]; sz <= nshift[2]; ++sz)
        for (int sy = -nshift[1]; sy <= nshift[1]; ++sy)
          for (int sx = -nshift[0]; sx <= nshift[0]; ++sx) {
            if (i == j && sx == 0 && sy == 0 && sz == 0) continue;
            const int sh[3] = {sx * (P.domhi[0] - P.domlo[0] + 1), sy * (P.domhi[1] - P.domlo[1] + 1),
                               sz * (P.domhi[2] - P.domlo[2] + 1)};
            BoxCopy c;
            bool empty = false;
            for (int d = 0; d < 3; ++d) {
              if (d >= dm) { c.lo[d] = 0; c.n[d] = 1; c.shift[d] = 0; continue; }
              const int slo = lo[j][d] + sh[d], shi = hi[j][d] + (nodal ? nodal[d] : 0) + sh[d];
              const int a = std::max(glo[d], slo), b = std::min(ghi[d], shi);
              if (a > b) { empty = true; break; }
              c.lo[d] = a;
              c.n[d] = b - a + 1;
              c.shift[d] = sh[d];
            }
            if (empty) continue;
            c.dst = fabs[i].p + fabs[i].cs * (scomp - 1);
            c.src = fabs[j].p + fabs[j].cs * (scomp - 1);
            for (int d = 0; d < 3; ++d) {
              c.dlo[d] = fabs[i].lo[d];
              c.slo[d] = fabs[j].lo[d];
              c.vlo[d] = d < dm ? vlo[d] : 0;
              c.vhi[d] = d < dm ? vhi[d] : 0;
            }
            c.dn[0] = fabs[i].n[0]; c.dn[1] = fabs[i].n[1];
            c.sn[0] = fabs[j].n[0]; c.sn[1] = fabs[j].n[1];
            c.dcs = fabs[i].cs; c.scs = fabs[j].cs;
            c.ncomp = ncomp;
            copies.push_back(c);
          }
  }
  // Face-centred data on a periodic domain stores the face on the domain boundary twice (lo and hi+1), so a ghost
  // face can have two images.  They hold the same value in MAESTRO's data; to be deterministic whatever the data, the
  // image reached with the fewest periodic shifts is written last (a ghost face that exists unshifted in direction d
  // takes that one: the same-index rule of a single box wrapping onto itself).
  auto nsh = [](const BoxCopy& c) { return (c.shift[0] != 0) + (c.shift[1] != 0) + (c.shift[2] != 0); };
  std::stable_sort(copies.begin(), copies.end(), [&](const BoxCopy& a, const BoxCopy& b) { return nsh(a) > nsh(b); });
  if (!copies.empty()) {
    if (copies.size() > 65535) throw Error("fill_boundary: too many box intersections");
    const BoxCopy* d_c = reinterpret_cast<const BoxCopy*>(upload_small(reinterpret_cast<const double*>(copies.data()),
                                                                      (copies.size() * sizeof(BoxCopy) + 7) / 8));
    size_t first = 0;
    while (first < copies.size()) {  // one launch per number of shifted directions, most-shifted first
      size_t last = first;
      long mx = 1;
      while (last < copies.size() && nsh(copies[last]) == nsh(copies[first])) {
        mx = std::max(mx, (long)copies[last].n[0] * copies[last].n[1] * copies[last].n[2]);
        ++last;
      }
      dim3 g((unsigned)std::min<long>((mx + 255) / 256, 64), (unsigned)(last - first));
      MGPU_TIMED(TAG_FILL, (k_box_copies<<<g, 256, 0, ctx().stream>>>(d_c + first, (int)(last - first))));
      first = last;
    }
  }
  for (int i = 0; i < nfabs; ++i) {  // physical BCs where the box touches the domain boundary (per-box table)
    FillReq r;
    r.P = P;
    r.sfull = fabs[i];
    for (int d = 0; d < 3; ++d) {
      r.lo[d] = d < dm ? lo[i][d] : 0;
      r.hi[d] = d < dm ? hi[i][d] : 0;
      r.nodal[d] = nodal ? nodal[d] : 0;
      r.pmask[d] = pmask[d];
    }
    r.ng = ng;
    r.has_nodal = nodal != nullptr;
    r.scomp = scomp; r.bccomp = bccomp; r.ncomp = ncomp;
    r.adv_bc = adv_bc[i];
    r.same_boundary = same_boundary;
    fill_physbc(r);
  }
}

void fill_boundary_dev(const mgpu_params& P, const DV& sfull, const int* lo, const int* hi, int ng,
                       const int* nodal, int scomp, int bccomp, int ncomp, const int* adv_bc, const int* pmask,
                       bool same_boundary) {
  if (ng == 0) return;
  FillReq r;
  r.P = P;
  r.sfull = sfull;
  for (int d = 0; d < 3; ++d) {
    r.lo[d] = d < P.dm ? lo[d] : 0;
    r.hi[d] = d < P.dm ? hi[d] : 0;
    r.nodal[d] = nodal ? nodal[d] : 0;
    r.pmask[d] = d < P.dm ? pmask[d] : 0;
  }
  r.has_nodal = nodal != nullptr;
  r.ng = ng;
  r.scomp = scomp;
  r.bccomp = bccomp;
  r.ncomp = ncomp;
  r.adv_bc = adv_bc;
  r.same_boundary = same_boundary;
  if (g_batching) {
    g_batch.push_back(r);
    return;
  }
  const char slab = fill_exchange(r) ? 1 : 0;
  fill_wraps(&r, &slab, 1);
  fill_physbc(r);
}

}  // namespace mgpu

// ------------------------------------------------------------------------------------------
// Fused mk_rhoX_flux + update_scal for the species AND tracer ranges of density_advance in one launch
// (density_advance.f90:280-366): one thread per zone computes the fluxes on its own lo faces (written
// once; the hi+1 faces by the last cell of each line), re-evaluates the hi-face fluxes from the same
// edge states (3 multiplies; cheaper than a second pass through HBM), then the conservative update,
// the density from the species updates, the floor and the negative-species redistribution.
// Arithmetic is expression-for-expression that of k_rhoX_flux / k_update_scal / k_update_rho.
namespace mgpu {

// DM: dimensionality; FAST: multiply by 1/dx instead of dividing (last-bit differences, <= 1e-12 relative)
template <int DM, bool FAST>
__global__ void __launch_bounds__(256, 3) k_flux_update_all(FluxArgs a, UpdArgs u, int t0, int ntrac, int rho, double bcd) {
  int ix[3];
  if (!decode3(u.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  constexpr int r = DM - 1;
  const int c0 = a.spec0, nspec = a.nspec;
  const int ir = ix[r];
  const int spt = a.species_pred_type;
  // face offsets (lo face of this zone; hi face = + stride) in the edge-state and flux fabs of each direction
  long oe[DM], of[DM], se_st[DM], sf_st[DM];
  double vlo[DM], vhi[DM], r0lo[DM], r0hi[DM], rdx[DM];
  bool last[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    oe[d] = a.sedge[d].off(i, j, k);
    of[d] = a.sflux[d].off(i, j, k);
    se_st[d] = a.sedge[d].stride(d);
    sf_st[d] = a.sflux[d].stride(d);
    const long ou = a.umac[d].off(i, j, k);
    vlo[d] = __ldg(a.umac[d].p + ou);
    vhi[d] = __ldg(a.umac[d].p + ou + a.umac[d].stride(d));
    last[d] = (ix[d] == u.vb.hi[d]);
    rdx[d] = FAST ? 1.0 / u.dx[d] : u.dx[d];
    double e_lo, e_hi;
    if (d != r) {  // mkflux.f90:410
      e_lo = e_hi = 0.5 * (a.rho0_old[ir] + a.rho0_new[ir]);
    } else {  // mkflux.f90:465
      e_lo = 0.5 * (a.rho0_edge_old[ir] + a.rho0_edge_new[ir]);
      e_hi = 0.5 * (a.rho0_edge_old[ir + 1] + a.rho0_edge_new[ir + 1]);
      vlo[d] = vlo[d] + a.w0[ir];
      vhi[d] = vhi[d] + a.w0[ir + 1];
    }
    // density factor of the flux: (rho0_edge + rho'_edge), 1, or rho_edge by species_pred_type (mkflux.f90:415-431)
    if (spt == MGPU_PREDICT_RHOX) {
      r0lo[d] = r0hi[d] = 1.0;
    } else {
      const double* pr = a.sedge[d].p + oe[d] + a.sedge[d].cs * a.rho;
      if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
        r0lo[d] = e_lo + __ldg(pr);
        r0hi[d] = e_hi + __ldg(pr + se_st[d]);
      } else {
        r0lo[d] = __ldg(pr);
        r0hi[d] = __ldg(pr + se_st[d]);
      }
    }
  }
  const bool do_eta = a.evolve_base_state;
  const bool last_r = last[r];
  double eta = 0.0, eta_hi = 0.0;
  long oeta = 0, eta_st = 0;
  if (do_eta) {
    oeta = a.eta.off(i, j, k);
    eta_st = a.eta.stride(r);
    eta = a.eta.p[oeta];
    if (last_r) eta_hi = a.eta.p[oeta + eta_st];
  }
  const long oso = u.sold.off(i, j, k), osn = u.snew.off(i, j, k), ofo = u.force.off(i, j, k);
  bool neg = false;
  double rnew = u.sold.p[oso + u.sold.cs * rho];
  const int ncomp = nspec + ntrac;
  for (int n = 0; n < ncomp; ++n) {
    const int c = (n < nspec) ? c0 + n : t0 + (n - nspec);
    double flo[DM], fhi[DM];
#pragma unroll
    for (int d = 0; d < DM; ++d) {
      const double* pe = a.sedge[d].p + oe[d] + a.sedge[d].cs * c;
      // vel * rhofac * X_edge (vel * X_edge for predict_rhoX): expression order of mkflux.f90:415-431
      if (spt == MGPU_PREDICT_RHOX) {
        flo[d] = vlo[d] * __ldg(pe);
        fhi[d] = vhi[d] * __ldg(pe + se_st[d]);
      } else {
        flo[d] = vlo[d] * r0lo[d] * __ldg(pe);
        fhi[d] = vhi[d] * r0hi[d] * __ldg(pe + se_st[d]);
      }
      double* pf = a.sflux[d].p + of[d] + a.sflux[d].cs * c;
      pf[0] = flo[d];
      if (last[d]) pf[sf_st[d]] = fhi[d];
    }
    if (do_eta && n < nspec) {  // mkflux.f90:486-494
      eta = eta + flo[r];
      if (last_r) eta_hi = eta_hi + fhi[r];
      if (n == nspec - 1) {
        eta = eta - a.w0[ir] * a.rho0_predicted_edge[ir];
        if (last_r) eta_hi = eta_hi - a.w0[ir + 1] * a.rho0_predicted_edge[ir + 1];
      }
    }
    double divterm;
    if (FAST) {
      divterm = (fhi[0] - flo[0]) * rdx[0] + (fhi[1] - flo[1]) * rdx[1];
      if (DM == 3) divterm = divterm + (fhi[DM - 1] - flo[DM - 1]) * rdx[DM - 1];
    } else {
      divterm = (fhi[0] - flo[0]) / rdx[0] + (fhi[1] - flo[1]) / rdx[1];
      if (DM == 3) divterm = divterm + (fhi[DM - 1] - flo[DM - 1]) / rdx[DM - 1];
    }
    const double so = __ldg(u.sold.p + oso + u.sold.cs * c);
    const double sn = so + u.dt * (-divterm + __ldg(u.force.p + ofo + u.force.cs * c));
    u.snew.p[osn + u.snew.cs * c] = sn;
    if (n < nspec) {
      rnew = rnew + (sn - so);
      if (sn < 0.0) neg = true;
    }
  }
  if (do_eta) {
    a.eta.p[oeta] = eta;
    if (last_r) a.eta.p[oeta + eta_st] = eta_hi;
  }
  // density, floor, negative species: update_scal.f90:453-505 (same statements as k_update_rho)
  double* sn = u.snew.p + osn;
  const long cn = u.snew.cs;
  const int c1 = c0 + nspec - 1;
  if (rnew < 0.5 * bcd) {
    for (int c = c0; c <= c1; ++c) sn[cn * c] = sn[cn * c] * 0.5 * bcd / rnew;
    rnew = 0.5 * bcd;
  }
  sn[cn * rho] = rnew;
  if (neg) {
    for (int c = c0; c <= c1; ++c) {
      if (sn[cn * c] < 0.0) {
        double delta = -sn[cn * c];
        double sumX = 0.0;
        for (int c2 = c0; c2 <= c1; ++c2)
          if (c2 != c && sn[cn * c2] >= 0.0) sumX = sumX + sn[cn * c2];
        for (int c2 = c0; c2 <= c1; ++c2)
          if (c2 != c && sn[cn * c2] >= 0.0) {
            double frac = sn[cn * c2] / sumX;
            sn[cn * c2] = sn[cn * c2] - frac * delta;
          }
        sn[cn * c] = 0.0;
      }
    }
  }
}

// 3-D FAST specialisation of k_flux_update_all: 32-bit element offsets (the 64-bit index arithmetic of the
// general kernel costs it a third of its issue slots and spills), sedge and sflux share one layout per
// direction, sold/snew share one layout, two components in flight per thread.  Same statements as above.
struct FU3 {
  const double* sedge[3];
  double* sflux[3];
  const double* umac[3];
  const double *sold, *force;
  double *snew, *eta;
  int e_lo[3][3], e_n0[3], e_n01[3], e_cs[3];  // face fabs (sedge == sflux layout), per direction
  int u_lo[3][3], u_n0[3], u_n01[3];
  int s_lo[3], s_n0, s_n01, s_cs;  // sold / snew
  int f_lo[3], f_n0, f_n01, f_cs;  // force
  int t_lo[3], t_n0, t_n01;        // etarhoflux
  int lo[3], hi[3];
  int kb;  // first z plane of this launch (a launch may cover a sub-range of planes; `hi` stays the box's)
  int spt, do_eta, rho, spec0, nspec, trac0, ntrac;
  int force_zero;  // scal_force is identically zero (density_advance.f90:349-351): do not read it
  double dt, rdx[3], half_bcd;
  const double *w0, *rho0_old, *rho0_edge_old, *rho0_new, *rho0_edge_new, *rho0_predicted_edge;
};

__global__ void __launch_bounds__(256, 4) k_flux_update3_fast(const __grid_constant__ FU3 a) {
  const int i = a.lo[0] + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int j = a.lo[1] + (int)blockIdx.y, k = a.kb + (int)blockIdx.z;
  if (i > a.hi[0]) return;
  int oe[3], se[3];
  double vlo[3], vhi[3], r0lo[3], r0hi[3];
  bool last[3];
  const int ix[3] = {i, j, k};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    oe[d] = (i - a.e_lo[d][0]) + a.e_n0[d] * (j - a.e_lo[d][1]) + a.e_n01[d] * (k - a.e_lo[d][2]);
    se[d] = d == 0 ? 1 : (d == 1 ? a.e_n0[d] : a.e_n01[d]);
    const int ou = (i - a.u_lo[d][0]) + a.u_n0[d] * (j - a.u_lo[d][1]) + a.u_n01[d] * (k - a.u_lo[d][2]);
    const int su = d == 0 ? 1 : (d == 1 ? a.u_n0[d] : a.u_n01[d]);
    vlo[d] = __ldg(a.umac[d] + ou);
    vhi[d] = __ldg(a.umac[d] + ou + su);
    last[d] = (ix[d] == a.hi[d]);
  }
  {
    const double ec = 0.5 * (a.rho0_old[k] + a.rho0_new[k]);                                   // mkflux.f90:410
    const double el = 0.5 * (a.rho0_edge_old[k] + a.rho0_edge_new[k]);                         // mkflux.f90:465
    const double eh = 0.5 * (a.rho0_edge_old[k + 1] + a.rho0_edge_new[k + 1]);
    vlo[2] += a.w0[k];
    vhi[2] += a.w0[k + 1];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (a.spt == MGPU_PREDICT_RHOX) {
        r0lo[d] = r0hi[d] = 1.0;
      } else {
        const double* pr = a.sedge[d] + oe[d] + a.e_cs[d] * a.rho;
        const double add_lo = (a.spt == MGPU_PREDICT_RHOPRIME_AND_X) ? (d == 2 ? el : ec) : 0.0;
        const double add_hi = (a.spt == MGPU_PREDICT_RHOPRIME_AND_X) ? (d == 2 ? eh : ec) : 0.0;
        r0lo[d] = add_lo + __ldg(pr);
        r0hi[d] = add_hi + __ldg(pr + se[d]);
      }
      // fold the velocity in: flux = (vel * rhofac) * X_edge
      r0lo[d] *= vlo[d];
      r0hi[d] *= vhi[d];
    }
  }
  const int os = (i - a.s_lo[0]) + a.s_n0 * (j - a.s_lo[1]) + a.s_n01 * (k - a.s_lo[2]);
  const int of = (i - a.f_lo[0]) + a.f_n0 * (j - a.f_lo[1]) + a.f_n01 * (k - a.f_lo[2]);
  double eta = 0.0, eta_hi = 0.0;
  int ot = 0;
  if (a.do_eta) {
    ot = (i - a.t_lo[0]) + a.t_n0 * (j - a.t_lo[1]) + a.t_n01 * (k - a.t_lo[2]);
    eta = a.eta[ot];
    if (last[2]) eta_hi = a.eta[ot + a.t_n01];
  }
  bool neg = false;
  double rnew = __ldg(a.sold + os + a.s_cs * a.rho);
  const int ncomp = a.nspec + a.ntrac;
#pragma unroll 2
  for (int n = 0; n < ncomp; ++n) {
    const int c = (n < a.nspec) ? a.spec0 + n : a.trac0 + (n - a.nspec);
    double flo[3], fhi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double* pe = a.sedge[d] + oe[d] + a.e_cs[d] * c;
      flo[d] = r0lo[d] * __ldg(pe);
      fhi[d] = r0hi[d] * __ldg(pe + se[d]);
    }
    const double so = __ldg(a.sold + os + a.s_cs * c);
    const double fo = a.force_zero ? 0.0 : __ldg(a.force + of + a.f_cs * c);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double* pf = a.sflux[d] + oe[d] + a.e_cs[d] * c;
      pf[0] = flo[d];
      if (last[d]) pf[se[d]] = fhi[d];
    }
    if (a.do_eta && n < a.nspec) {  // mkflux.f90:486-494
      eta += flo[2];
      if (last[2]) eta_hi += fhi[2];
    }
    const double divterm = (fhi[0] - flo[0]) * a.rdx[0] + (fhi[1] - flo[1]) * a.rdx[1] + (fhi[2] - flo[2]) * a.rdx[2];
    const double sn = so + a.dt * (fo - divterm);
    a.snew[os + a.s_cs * c] = sn;
    if (n < a.nspec) {
      rnew += sn - so;
      neg = neg || (sn < 0.0);
    }
  }
  if (a.do_eta) {
    if (a.nspec > 0) {
      eta -= a.w0[k] * a.rho0_predicted_edge[k];
      eta_hi -= a.w0[k + 1] * a.rho0_predicted_edge[k + 1];
    }
    a.eta[ot] = eta;
    if (last[2]) a.eta[ot + a.t_n01] = eta_hi;
  }
  // density, floor, negative species: update_scal.f90:453-505
  double* sn = a.snew + os;
  const int cn = a.s_cs, c0 = a.spec0, c1 = a.spec0 + a.nspec - 1;
  if (rnew < a.half_bcd) {
    for (int c = c0; c <= c1; ++c) sn[cn * c] = sn[cn * c] * a.half_bcd / rnew;
    rnew = a.half_bcd;
  }
  sn[cn * a.rho] = rnew;
  if (neg) {
    for (int c = c0; c <= c1; ++c) {
      if (sn[cn * c] < 0.0) {
        const double delta = -sn[cn * c];
        double sumX = 0.0;
        for (int c2 = c0; c2 <= c1; ++c2)
          if (c2 != c && sn[cn * c2] >= 0.0) sumX = sumX + sn[cn * c2];
        for (int c2 = c0; c2 <= c1; ++c2)
          if (c2 != c && sn[cn * c2] >= 0.0) {
            const double frac = sn[cn * c2] / sumX;
            sn[cn * c2] = sn[cn * c2] - frac * delta;
          }
        sn[cn * c] = 0.0;
      }
    }
  }
}

static bool same_layout(const DV& x, const DV& y) {
  for (int d = 0; d < 3; ++d)
    if (x.lo[d] != y.lo[d] || x.n[d] != y.n[d]) return false;
  return x.cs == y.cs;
}
// true if the specialised kernel covers this call (and then launches it)
// k0 <= k1: only the z planes [k0, k1] of the box (the slab episodes update the planes next to the slab faces first
// and exchange them while the interior planes are updated); dry: only answer whether the kernel covers the call
static bool flux_update3_fast(const mgpu_params& P, const FluxArgs& a, const UpdArgs& u, bool force_zero, int k0 = 1,
                              int k1 = 0, bool dry = false) {
  if (P.dm != 3) return false;
  const long lim = 1L << 31;
  for (int d = 0; d < 3; ++d)
    if (!same_layout(a.sedge[d], a.sflux[d]) || a.sedge[d].cs * a.sedge[d].nc >= lim || a.umac[d].cs >= lim) return false;
  if (!same_layout(u.sold, u.snew) || u.sold.cs * u.sold.nc >= lim || u.force.cs * u.force.nc >= lim) return false;
  FU3 f;
  memset(&f, 0, sizeof(f));
  for (int d = 0; d < 3; ++d) {
    f.sedge[d] = a.sedge[d].p;
    f.sflux[d] = a.sflux[d].p;
    f.umac[d] = a.umac[d].p;
    for (int q = 0; q < 3; ++q) {
      f.e_lo[d][q] = a.sedge[d].lo[q];
      f.u_lo[d][q] = a.umac[d].lo[q];
    }
    f.e_n0[d] = a.sedge[d].n[0];
    f.e_n01[d] = a.sedge[d].n[0] * a.sedge[d].n[1];
    f.e_cs[d] = (int)a.sedge[d].cs;
    f.u_n0[d] = a.umac[d].n[0];
    f.u_n01[d] = a.umac[d].n[0] * a.umac[d].n[1];
    f.s_lo[d] = u.sold.lo[d];
    f.f_lo[d] = u.force.lo[d];
    f.t_lo[d] = a.eta.lo[d];
    f.lo[d] = u.vb.lo[d];
    f.hi[d] = u.vb.hi[d];
    f.rdx[d] = 1.0 / u.dx[d];
  }
  f.sold = u.sold.p; f.snew = u.snew.p; f.force = u.force.p; f.eta = a.eta.p;
  f.s_n0 = u.sold.n[0]; f.s_n01 = u.sold.n[0] * u.sold.n[1]; f.s_cs = (int)u.sold.cs;
  f.f_n0 = u.force.n[0]; f.f_n01 = u.force.n[0] * u.force.n[1]; f.f_cs = (int)u.force.cs;
  f.t_n0 = a.eta.n[0]; f.t_n01 = a.eta.n[0] * a.eta.n[1];
  f.spt = a.species_pred_type; f.do_eta = a.evolve_base_state ? 1 : 0;
  f.rho = a.rho; f.spec0 = a.spec0; f.nspec = a.nspec; f.trac0 = P.trac_comp - 1; f.ntrac = P.ntrac;
  f.dt = u.dt; f.half_bcd = 0.5 * P.base_cutoff_density;
  f.force_zero = force_zero ? 1 : 0;
  f.w0 = a.w0; f.rho0_old = a.rho0_old; f.rho0_edge_old = a.rho0_edge_old; f.rho0_new = a.rho0_new;
  f.rho0_edge_new = a.rho0_edge_new; f.rho0_predicted_edge = a.rho0_predicted_edge;
  if (dry) return true;
  dim3 g = grid3(u.vb, 256);
  const int b = block3(u.vb, 256);
  f.kb = f.lo[2];
  if (k0 <= k1) {
    f.kb = k0;
    g.z = (unsigned)(k1 - k0 + 1);
  }
  MGPU_TIMED(TAG_UPDATE, (k_flux_update3_fast<<<g, b, 0, ctx().stream>>>(f)));
  return true;
}
bool flux_update_range_supported(const mgpu_params& P, const FluxArgs& a, const UpdArgs& u) {
  return flux_update3_fast(P, a, u, true, 1, 0, true);
}
void flux_update_range_dev(const mgpu_params& P, FluxArgs& a, UpdArgs& u, bool force_zero, int k0, int k1) {
  if (k0 > k1) return;
  if (!flux_update3_fast(P, a, u, force_zero, k0, k1)) throw Error("update_scal: plane ranges need the lean kernel");
}

void flux_update_all_dev(const mgpu_params& P, FluxArgs& a, UpdArgs& u, bool exact, bool force_zero, bool skip_rho_copy) {
  Context& cx = ctx();
  const int rho = P.rho_comp - 1;
  if (u.snew.cs != u.sold.cs) throw Error("update_scal: sold and snew must have the same ghost width");
  // snew(:,:,:,rho_comp) = sold(:,:,:,rho_comp) including ghost cells (update_scal.f90:455)
  // (skipped when the caller knows that the ghost fill that follows overwrites every ghost cell of the component)
  auto rho_copy = [&]() {
    MGPU_TIMED(TAG_UPDATE, (k_copy<<<nblocks(u.snew.cs, 256), 256, 0, cx.stream>>>(u.snew.p + u.snew.cs * rho,
                                                                                   u.sold.p + u.sold.cs * rho, u.snew.cs)));
  };
  if (!skip_rho_copy) rho_copy();
  if (!exact && flux_update3_fast(P, a, u, force_zero)) return;
  // general kernel: it reads the force (zero in memory when force_zero) and the copy is always made
  if (skip_rho_copy) rho_copy();
  const dim3 g = grid3(u.vb, 256);
  const int b = block3(u.vb, 256);
  const int t0 = P.trac_comp - 1;
  const double bcd = P.base_cutoff_density;
  if (P.dm == 3) {
    if (exact) MGPU_TIMED(TAG_UPDATE, (k_flux_update_all<3, false><<<g, b, 0, cx.stream>>>(a, u, t0, P.ntrac, rho, bcd)));
    else MGPU_TIMED(TAG_UPDATE, (k_flux_update_all<3, true><<<g, b, 0, cx.stream>>>(a, u, t0, P.ntrac, rho, bcd)));
  } else {
    if (exact) MGPU_TIMED(TAG_UPDATE, (k_flux_update_all<2, false><<<g, b, 0, cx.stream>>>(a, u, t0, P.ntrac, rho, bcd)));
    else MGPU_TIMED(TAG_UPDATE, (k_flux_update_all<2, true><<<g, b, 0, cx.stream>>>(a, u, t0, P.ntrac, rho, bcd)));
  }
}

}  // namespace mgpu
