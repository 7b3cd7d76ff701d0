// General (all ppm_type / BC / conservative / is_vel variants) device path of make_edge_scal.
//
// Replaces make_edge_scal_2d (Source/make_edge_scal.f90:290) and make_edge_scal_3d (:677) together
// with the slope/ppm calls they make.  Stage structure (one launch each, temporaries in the arena):
//   k_cell_states : per cell, per direction: 1-D extrapolated states Ip/Im   (ppm.f90 / slope.f90)
//   k_simh        : normal-predictor Riemann states simh_d                   (:807-1059)
//   k_transverse  : 3-D only, the six transverse states simh_{d|t}           (:1065-1683)
//   k_final       : final edge states + BCs                                  (:1692-2040, 2-D :530-670)
// The fused single-launch kernel for the headline case lives in mgpu_fused.cu; this path covers
// everything else and is the reference implementation the fused kernel is tested against.
#include "mgpu_edge.cuh"
#include "mgpu_recon.cuh"

namespace mgpu {

__device__ __forceinline__ void lr_bc(const EdgeArgs& a, int d, int f, const double* sq, long st, int stage,
                                      double& sl, double& sr) {
  const int is = a.lo[d], ie = a.hi[d];
  const int bclo = a.bclo[d], bchi = a.bchi[d];
  const bool velnorm = a.velnorm[d];
  // QUIRK make_edge_scal.f90:421-423: 2-D x-lo REFLECT_ODD zeroes index ie+1 instead of is
  const bool quirk_odd = (a.dm == 2 && d == 0 && stage == 0 && bclo == MGPU_BC_REFLECT_ODD);
  if (f == is) {
    if (bclo == MGPU_BC_EXT_DIR) {
      // QUIRK make_edge_scal.f90:1010-1011: 3-D z-lo normal predictor takes s(ks), not s(ks-1)
      const int off = (a.dm == 3 && d == 2 && stage == 0) ? 0 : -1;
      sl = sq[off * st];
      sr = sl;
    } else if (bclo == MGPU_BC_FOEXTRAP || bclo == MGPU_BC_HOEXTRAP) {
      if (velnorm) sr = dmin2(sr, 0.0);
      sl = sr;
    } else if (bclo == MGPU_BC_REFLECT_EVEN) {
      sl = sr;
    } else if (bclo == MGPU_BC_REFLECT_ODD) {
      if (!quirk_odd) { sl = 0.0; sr = 0.0; }
    }
  }
  if (f == ie + 1) {
    if (quirk_odd) { sl = 0.0; sr = 0.0; }
    if (bchi == MGPU_BC_EXT_DIR) {
      sl = sq[0];
      sr = sl;
    } else if (bchi == MGPU_BC_FOEXTRAP || bchi == MGPU_BC_HOEXTRAP) {
      if (velnorm) sl = dmax2(sl, 0.0);
      sr = sl;
    } else if (bchi == MGPU_BC_REFLECT_EVEN) {
      sr = sl;
    } else if (bchi == MGPU_BC_REFLECT_ODD) {
      sl = 0.0;
      sr = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void k_cell_states(EdgeArgs a) {
  int ix[3];
  if (!decode3(a.tb, ix)) return;
  const long so = a.s.off(ix[0], ix[1], ix[2]);
  const long to = a.Ip.off(ix[0], ix[1], ix[2]);
  for (int d = 0; d < a.dm; ++d) {
    const long st = a.s.stride(d);
    const LineBC b = make_linebc(a.dm, d, a.lo[d], a.hi[d], a.bclo[d], a.bchi[d]);
    const DV& u = a.umac[d];
    const long uo = u.off(ix[0], ix[1], ix[2]);
    const double ulo = u.p[uo], uhi = u.p[uo + u.stride(d)];
    double Ip, Im;
    cell_states(a.ppm_type, a.slope_order, a.s.p + so, st, ix[d], b, uhi, ulo, a.dt, a.dx[d], a.rel_eps, Ip, Im);
    a.Ip.p[to + a.Ip.cs * d] = Ip;
    a.Im.p[to + a.Im.cs * d] = Im;
    if (a.trace_forces) {
      const long fo = a.force.off(ix[0], ix[1], ix[2]);
      cell_states(a.ppm_type, a.slope_order, a.force.p + fo, a.force.stride(d), ix[d], b, uhi, ulo, a.dt, a.dx[d],
                  a.rel_eps, Ip, Im);
      a.Ipf.p[to + a.Ipf.cs * d] = Ip;
      a.Imf.p[to + a.Imf.cs * d] = Im;
    }
  }
}

// left/right normal-predictor states of the d-face whose right cell is ix (BCs of stage 0 applied)
__device__ __forceinline__ void face_lr(const EdgeArgs& a, int d, const int* ix, double& sl, double& sr) {
  const long to = a.Ip.off(ix[0], ix[1], ix[2]);
  sl = a.Ip.p[to - a.Ip.stride(d) + a.Ip.cs * d];
  sr = a.Im.p[to + a.Im.cs * d];
  lr_bc(a, d, ix[d], a.s.p + a.s.off(ix[0], ix[1], ix[2]), a.s.stride(d), 0, sl, sr);
}

__global__ void k_simh(EdgeArgs a) {
  int ix[3];
  if (!decode3(a.tb, ix)) return;
  for (int d = 0; d < a.dm; ++d) {
    if (ix[d] < a.lo[d]) continue;  // faces lo..hi+1 in d, lo-1..hi+1 transverse
    double sl, sr;
    face_lr(a, d, ix, sl, sr);
    a.simh[d](ix[0], ix[1], ix[2]) = riemann(sl, sr, a.umac[d](ix[0], ix[1], ix[2]), a.rel_eps);
  }
}

// divergence used by the conservative corner coupling: every term over dx(1) (make_edge_scal.f90:768-770)
__device__ __forceinline__ double divu_of(const EdgeArgs& a, int i, int j, int k) {
  return (a.umac[0](i + 1, j, k) - a.umac[0](i, j, k) + a.umac[1](i, j + 1, k) - a.umac[1](i, j, k) +
          a.umac[2](i, j, k + 1) - a.umac[2](i, j, k)) / a.dx[0];
}

__global__ void k_transverse(EdgeArgs a) {
  int ix[3];
  if (!decode3(a.tb, ix)) return;
  const double dt3 = a.dt / 3.0, dt6 = a.dt / 6.0;
  for (int d = 0; d < 3; ++d) {
    if (ix[d] < a.lo[d]) continue;
    for (int t = 0; t < 3; ++t) {
      if (t == d) continue;
      if (ix[t] < a.lo[t] || ix[t] > a.hi[t]) continue;
      double sl, sr;
      face_lr(a, d, ix, sl, sr);
      int cl[3] = {ix[0], ix[1], ix[2]};
      cl[d] -= 1;  // cell left of the face
      const DV& vt = a.umac[t];
      const DV& sh = a.simh[t];
      const double ht = a.dx[t];
      const long vst = vt.stride(t), sst = sh.stride(t);
      const long vl = vt.off(cl[0], cl[1], cl[2]), vr = vt.off(ix[0], ix[1], ix[2]);
      const long hl = sh.off(cl[0], cl[1], cl[2]), hr = sh.off(ix[0], ix[1], ix[2]);
      double slt, srt;
      if (a.is_cons) {
        const double scl = a.s(cl[0], cl[1], cl[2]), scr = a.s(ix[0], ix[1], ix[2]);
        slt = sl - (dt3 / ht) * (sh.p[hl + sst] * vt.p[vl + vst] - sh.p[hl] * vt.p[vl]) -
              dt3 * scl * divu_of(a, cl[0], cl[1], cl[2]) + (dt3 / ht) * scl * (vt.p[vl + vst] - vt.p[vl]);
        srt = sr - (dt3 / ht) * (sh.p[hr + sst] * vt.p[vr + vst] - sh.p[hr] * vt.p[vr]) -
              dt3 * scr * divu_of(a, ix[0], ix[1], ix[2]) + (dt3 / ht) * scr * (vt.p[vr + vst] - vt.p[vr]);
      } else {
        slt = sl - (dt6 / ht) * (vt.p[vl + vst] + vt.p[vl]) * (sh.p[hl + sst] - sh.p[hl]);
        srt = sr - (dt6 / ht) * (vt.p[vr + vst] + vt.p[vr]) * (sh.p[hr + sst] - sh.p[hr]);
      }
      lr_bc(a, d, ix[d], a.s.p + a.s.off(ix[0], ix[1], ix[2]), a.s.stride(d), 1, slt, srt);
      a.simht[d][t](ix[0], ix[1], ix[2]) = riemann(slt, srt, a.umac[d](ix[0], ix[1], ix[2]), a.rel_eps);
    }
  }
}

__global__ void k_final(EdgeArgs a) {
  int ix[3];
  Box3 fb = a.vb;
  for (int d = 0; d < a.dm; ++d) fb.hi[d] += 1;
  if (!decode3(fb, ix)) return;
  const double dt2 = 0.5 * a.dt, dt4 = a.dt / 4.0;
  for (int d = 0; d < a.dm; ++d) {
    bool ok = true;
    for (int q = 0; q < a.dm; ++q)
      if (q != d && ix[q] > a.hi[q]) ok = false;
    if (!ok) continue;
    double sl, sr;
    face_lr(a, d, ix, sl, sr);
    int cl[3] = {ix[0], ix[1], ix[2]};
    cl[d] -= 1;
    const long fl_o = a.force.off(cl[0], cl[1], cl[2]), fr_o = a.force.off(ix[0], ix[1], ix[2]);
    double fl, fr;
    if (a.trace_forces) {
      fl = a.Ipf.p[a.Ipf.off(cl[0], cl[1], cl[2]) + a.Ipf.cs * d];
      fr = a.Imf.p[a.Imf.off(ix[0], ix[1], ix[2]) + a.Imf.cs * d];
    } else {
      fl = a.force.p[fl_o];
      fr = a.force.p[fr_o];
    }
    const DV& vd = a.umac[d];
    const double hd = a.dx[d];
    const double scl = a.s(cl[0], cl[1], cl[2]), scr = a.s(ix[0], ix[1], ix[2]);
    double sedgel, sedger;
    if (a.dm == 2) {
      const int t = 1 - d;
      const DV& vt = a.umac[t];
      const DV& sh = a.simh[t];
      const double ht = a.dx[t];
      const long vst = vt.stride(t), sst = sh.stride(t);
      const long vl = vt.off(cl[0], cl[1], cl[2]), vr = vt.off(ix[0], ix[1], ix[2]);
      const long hl = sh.off(cl[0], cl[1], cl[2]), hr = sh.off(ix[0], ix[1], ix[2]);
      if (a.is_cons) {
        sedgel = sl - (dt2 / ht) * (sh.p[hl + sst] * vt.p[vl + vst] - sh.p[hl] * vt.p[vl]) -
                 (dt2 / hd) * scl * (vd(ix[0], ix[1], ix[2]) - vd(cl[0], cl[1], cl[2])) + dt2 * fl;
        sedger = sr - (dt2 / ht) * (sh.p[hr + sst] * vt.p[vr + vst] - sh.p[hr] * vt.p[vr]) -
                 (dt2 / hd) * scr * (vd.p[vd.off(ix[0], ix[1], ix[2]) + vd.stride(d)] - vd(ix[0], ix[1], ix[2])) +
                 dt2 * fr;
      } else {
        sedgel = sl - (dt4 / ht) * (vt.p[vl + vst] + vt.p[vl]) * (sh.p[hl + sst] - sh.p[hl]) + dt2 * fl;
        sedger = sr - (dt4 / ht) * (vt.p[vr + vst] + vt.p[vr]) * (sh.p[hr + sst] - sh.p[hr]) + dt2 * fr;
      }
    } else {
      const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
      const DV& v1 = a.umac[t1];
      const DV& v2 = a.umac[t2];
      const DV& q1 = a.simht[t1][t2];
      const DV& q2 = a.simht[t2][t1];
      const double h1 = a.dx[t1], h2 = a.dx[t2];
      const long v1s = v1.stride(t1), v2s = v2.stride(t2), q1s = q1.stride(t1), q2s = q2.stride(t2);
      const long v1l = v1.off(cl[0], cl[1], cl[2]), v1r = v1.off(ix[0], ix[1], ix[2]);
      const long v2l = v2.off(cl[0], cl[1], cl[2]), v2r = v2.off(ix[0], ix[1], ix[2]);
      const long q1l = q1.off(cl[0], cl[1], cl[2]), q1r = q1.off(ix[0], ix[1], ix[2]);
      const long q2l = q2.off(cl[0], cl[1], cl[2]), q2r = q2.off(ix[0], ix[1], ix[2]);
      if (a.is_cons) {
        sedgel = sl - (dt2 / h1) * (q1.p[q1l + q1s] * v1.p[v1l + v1s] - q1.p[q1l] * v1.p[v1l]) -
                 (dt2 / h2) * (q2.p[q2l + q2s] * v2.p[v2l + v2s] - q2.p[q2l] * v2.p[v2l]) -
                 (dt2 / hd) * scl * (vd(ix[0], ix[1], ix[2]) - vd(cl[0], cl[1], cl[2])) + dt2 * fl;
        sedger = sr - (dt2 / h1) * (q1.p[q1r + q1s] * v1.p[v1r + v1s] - q1.p[q1r] * v1.p[v1r]) -
                 (dt2 / h2) * (q2.p[q2r + q2s] * v2.p[v2r + v2s] - q2.p[q2r] * v2.p[v2r]) -
                 (dt2 / hd) * scr * (vd.p[vd.off(ix[0], ix[1], ix[2]) + vd.stride(d)] - vd(ix[0], ix[1], ix[2])) +
                 dt2 * fr;
      } else {
        sedgel = sl - (dt4 / h1) * (v1.p[v1l + v1s] + v1.p[v1l]) * (q1.p[q1l + q1s] - q1.p[q1l]) -
                 (dt4 / h2) * (v2.p[v2l + v2s] + v2.p[v2l]) * (q2.p[q2l + q2s] - q2.p[q2l]) + dt2 * fl;
        sedger = sr - (dt4 / h1) * (v1.p[v1r + v1s] + v1.p[v1r]) * (q1.p[q1r + q1s] - q1.p[q1r]) -
                 (dt4 / h2) * (v2.p[v2r + v2s] + v2.p[v2r]) * (q2.p[q2r + q2s] - q2.p[q2r]) + dt2 * fr;
      }
    }
    double e = riemann(sedgel, sedger, vd(ix[0], ix[1], ix[2]), a.rel_eps);
    // final BCs, make_edge_scal.f90:1768-1804
    const int f = ix[d];
    if (f == a.lo[d]) {
      const int bc = a.bclo[d];
      if (bc == MGPU_BC_EXT_DIR) e = scl;
      else if (bc == MGPU_BC_FOEXTRAP || bc == MGPU_BC_HOEXTRAP) e = a.velnorm[d] ? dmin2(sedger, 0.0) : sedger;
      else if (bc == MGPU_BC_REFLECT_EVEN) e = sedger;
      else if (bc == MGPU_BC_REFLECT_ODD) e = 0.0;
    }
    if (f == a.hi[d] + 1) {
      const int bc = a.bchi[d];
      if (bc == MGPU_BC_EXT_DIR) e = scr;
      else if (bc == MGPU_BC_FOEXTRAP || bc == MGPU_BC_HOEXTRAP) e = a.velnorm[d] ? dmax2(sedgel, 0.0) : sedgel;
      else if (bc == MGPU_BC_REFLECT_EVEN) e = sedgel;
      else if (bc == MGPU_BC_REFLECT_ODD) e = 0.0;
    }
    a.sedge[d](ix[0], ix[1], ix[2]) = e;
  }
}

static void check_bc(int bc, const char* what) {
  switch (bc) {
    case MGPU_BC_EXT_DIR: case MGPU_BC_FOEXTRAP: case MGPU_BC_HOEXTRAP: case MGPU_BC_REFLECT_EVEN:
    case MGPU_BC_REFLECT_ODD: case MGPU_BC_INTERIOR:
      return;
    default:
      throw Error(std::string("make_edge_scal: invalid boundary type ") + what);
  }
}

// one component of one box; all pointers are device pointers
void make_edge_scal_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                        const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel,
                        bool is_cons, int ng_s, int ng_f) {
  Context& c = ctx();
  const int dm = P.dm;
  if (P.ppm_type == 2 && ng_s < 4) throw Error("Need 4 ghost cells for ppm_type=2");  // ppm.f90:1864-1866
  if (ng_s < 3) throw Error("make_edge_scal: need at least 3 ghost cells");
  if (P.ppm_trace_forces == 1 && (P.ppm_type == 0 || ng_f < ng_s))
    throw Error("make_edge_scal: ppm_trace_forces needs ppm_type>=1 and force with ng_s ghost cells");
  EdgeArgs a;
  a.dm = dm;
  a.ppm_type = P.ppm_type;
  a.slope_order = P.slope_order;
  a.trace_forces = (P.ppm_trace_forces == 1);
  a.is_cons = is_cons;
  a.dt = P.dt;
  a.rel_eps = P.rel_eps;
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = d < dm ? lo[d] : 0;
    a.hi[d] = d < dm ? hi[d] : 0;
    a.dx[d] = P.dx[d < dm ? d : 0];
    a.bclo[d] = a.bchi[d] = MGPU_BC_INTERIOR;
    a.velnorm[d] = false;
    if (d < dm) {
      a.bclo[d] = adv_bc[d + dm * (0 + 2 * (bccomp - 1))];
      a.bchi[d] = adv_bc[d + dm * (1 + 2 * (bccomp - 1))];
      check_bc(a.bclo[d], "adv_bc(d,1)");
      check_bc(a.bchi[d], "adv_bc(d,2)");
      a.velnorm[d] = is_vel && (comp == d);
      a.umac[d] = umac[d];
      a.sedge[d] = sedge_full[d].comp(comp);
    }
  }
  a.tb = grown(lo, hi, dm, 1);
  a.vb = grown(lo, hi, dm, 0);
  a.s = s_full.comp(comp);
  a.force = force_full.comp(comp);
  const long nt = a.tb.npts();
  int z3[3] = {0, 0, 0};
  auto tmp = [&](int nc) { return make_view(arena_alloc((size_t)nt * nc), a.tb.lo, a.tb.hi, dm, 0, z3, nc); };
  a.Ip = tmp(dm);
  a.Im = tmp(dm);
  if (a.trace_forces) { a.Ipf = tmp(dm); a.Imf = tmp(dm); }
  for (int d = 0; d < dm; ++d) a.simh[d] = tmp(1);
  if (dm == 3)
    for (int d = 0; d < 3; ++d)
      for (int t = 0; t < 3; ++t)
        if (t != d) a.simht[d][t] = tmp(1);
  const int bs = 256;
  MGPU_TIMED(TAG_EDGE_CELL, (k_cell_states<<<grid3(a.tb, bs), block3(a.tb, bs), 0, c.stream>>>(a)));
  MGPU_TIMED(TAG_EDGE_SIMH, (k_simh<<<grid3(a.tb, bs), block3(a.tb, bs), 0, c.stream>>>(a)));
  if (dm == 3) MGPU_TIMED(TAG_EDGE_TRANS, (k_transverse<<<grid3(a.tb, bs), block3(a.tb, bs), 0, c.stream>>>(a)));
  Box3 fb = a.vb;
  for (int d = 0; d < dm; ++d) fb.hi[d] += 1;
  MGPU_TIMED(TAG_EDGE_FINAL, (k_final<<<grid3(fb, bs), block3(fb, bs), 0, c.stream>>>(a)));
}

}  // namespace mgpu

namespace mgpu {
size_t make_edge_scal_scratch(const mgpu_params& P, const int* lo, const int* hi) {
  Box3 tb = grown(lo, hi, P.dm, 1);
  // Ip, Im (+Ipf, Imf): dm comps each; simh: dm; simht: 6 in 3-D; +256 B alignment slack per array
  const size_t narr = (size_t)(P.ppm_trace_forces == 1 ? 4 : 2) * P.dm + P.dm + (P.dm == 3 ? 6 : 0);
  return narr * ((size_t)tb.npts() * sizeof(double) + 256) + 4096;
}
}  // namespace mgpu
