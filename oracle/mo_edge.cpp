// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of Source/make_edge_scal.f90: make_edge_scal_2d :290, make_edge_scal_3d :677.
// Same stages, temporaries (index ranges) and expression order as the reference; the x/y/z blocks
// are permutations of each other and are produced here by one direction-generic routine, with the
// reference's asymmetries kept literally (marked QUIRK).
#include "mo_kernels.h"

namespace mo {

namespace {

struct EdgeCtx {
  int dm;
  const int* lo;
  const int* hi;
  const int* adv_bc;
  int bccomp;  // 1-based
  int comp;    // 0-based component of s (comp+1 is the Fortran comp)
  bool is_vel;
  double rel_eps;
};

// Riemann/upwind select used everywhere (make_edge_scal.f90:881-883)
inline double riemann(double l, double r, double u, double rel_eps) {
  double v = (u > 0.0) ? l : r;
  double savg = 0.5 * (l + r);
  return (dabs(u) > rel_eps) ? v : savg;
}

// BCs on left/right face states of direction d over the transverse box tbx (d collapsed).
// stage: 0 = normal predictor, 1 = transverse predictor.  make_edge_scal.f90:836-874, 1108-1146
void lr_bc(const EdgeCtx& c, int d, Arr& sl, Arr& sr, const Arr& s, Box tbx, int stage) {
  const int is = c.lo[d], ie = c.hi[d];
  const int bclo = advbc(c.adv_bc, c.dm, d + 1, 1, c.bccomp);
  const int bchi = advbc(c.adv_bc, c.dm, d + 1, 2, c.bccomp);
  const bool velnorm = c.is_vel && (c.comp == d);
  Box b = tbx;
  // ---- lo side ----
  b.lo[d] = b.hi[d] = is;
  if (bclo == MGPU_BC_EXT_DIR) {
    // QUIRK make_edge_scal.f90:1010-1011: the 3-D z-lo normal predictor takes s(ks), not s(ks-1)
    const int off = (c.dm == 3 && d == 2 && stage == 0) ? 0 : -1;
    for_box(b, [&](int i, int j, int k) {
      sl(i, j, k) = s.at(i, j, k, d, off);
      sr(i, j, k) = s.at(i, j, k, d, off);
    });
  } else if (bclo == MGPU_BC_FOEXTRAP || bclo == MGPU_BC_HOEXTRAP) {
    for_box(b, [&](int i, int j, int k) {
      if (velnorm) sr(i, j, k) = dmin(sr(i, j, k), 0.0);
      sl(i, j, k) = sr(i, j, k);
    });
  } else if (bclo == MGPU_BC_REFLECT_EVEN) {
    for_box(b, [&](int i, int j, int k) { sl(i, j, k) = sr(i, j, k); });
  } else if (bclo == MGPU_BC_REFLECT_ODD) {
    // QUIRK make_edge_scal.f90:421-423: the 2-D x-lo branch zeroes index ie+1 instead of is
    if (c.dm == 2 && d == 0 && stage == 0) b.lo[d] = b.hi[d] = ie + 1;
    for_box(b, [&](int i, int j, int k) { sl(i, j, k) = 0.0; sr(i, j, k) = 0.0; });
  } else if (bclo == MGPU_BC_INTERIOR) {
  } else {
    fail("make_edge_scal: invalid boundary type adv_bc(d,1)");
  }
  // ---- hi side ----
  b.lo[d] = b.hi[d] = ie + 1;
  if (bchi == MGPU_BC_EXT_DIR) {
    for_box(b, [&](int i, int j, int k) {
      sl(i, j, k) = s(i, j, k);
      sr(i, j, k) = s(i, j, k);
    });
  } else if (bchi == MGPU_BC_FOEXTRAP || bchi == MGPU_BC_HOEXTRAP) {
    for_box(b, [&](int i, int j, int k) {
      if (velnorm) sl(i, j, k) = dmax(sl(i, j, k), 0.0);
      sr(i, j, k) = sl(i, j, k);
    });
  } else if (bchi == MGPU_BC_REFLECT_EVEN) {
    for_box(b, [&](int i, int j, int k) { sr(i, j, k) = sl(i, j, k); });
  } else if (bchi == MGPU_BC_REFLECT_ODD) {
    for_box(b, [&](int i, int j, int k) { sl(i, j, k) = 0.0; sr(i, j, k) = 0.0; });
  } else if (bchi == MGPU_BC_INTERIOR) {
  } else {
    fail("make_edge_scal: invalid boundary type adv_bc(d,2)");
  }
}

// BCs on the final edge state, make_edge_scal.f90:1768-1804 (3-D) / 563-599 (2-D)
void final_bc(const EdgeCtx& c, int d, Arr& sedge, const Arr& sedgel, const Arr& sedger, const Arr& s, Box vb) {
  const int is = c.lo[d], ie = c.hi[d];
  const int bclo = advbc(c.adv_bc, c.dm, d + 1, 1, c.bccomp);
  const int bchi = advbc(c.adv_bc, c.dm, d + 1, 2, c.bccomp);
  const bool velnorm = c.is_vel && (c.comp == d);
  Box b = vb;
  b.lo[d] = b.hi[d] = is;
  if (bclo == MGPU_BC_EXT_DIR) {
    for_box(b, [&](int i, int j, int k) { sedge(i, j, k) = s.at(i, j, k, d, -1); });
  } else if (bclo == MGPU_BC_FOEXTRAP || bclo == MGPU_BC_HOEXTRAP) {
    for_box(b, [&](int i, int j, int k) {
      sedge(i, j, k) = velnorm ? dmin(sedger(i, j, k), 0.0) : sedger(i, j, k);
    });
  } else if (bclo == MGPU_BC_REFLECT_EVEN) {
    for_box(b, [&](int i, int j, int k) { sedge(i, j, k) = sedger(i, j, k); });
  } else if (bclo == MGPU_BC_REFLECT_ODD) {
    for_box(b, [&](int i, int j, int k) { sedge(i, j, k) = 0.0; });
  } else if (bclo == MGPU_BC_INTERIOR) {
  } else {
    fail("make_edge_scal: invalid boundary type adv_bc(d,1)");
  }
  b.lo[d] = b.hi[d] = ie + 1;
  if (bchi == MGPU_BC_EXT_DIR) {
    for_box(b, [&](int i, int j, int k) { sedge(i, j, k) = s(i, j, k); });
  } else if (bchi == MGPU_BC_FOEXTRAP || bchi == MGPU_BC_HOEXTRAP) {
    for_box(b, [&](int i, int j, int k) {
      sedge(i, j, k) = velnorm ? dmax(sedgel(i, j, k), 0.0) : sedgel(i, j, k);
    });
  } else if (bchi == MGPU_BC_REFLECT_EVEN) {
    for_box(b, [&](int i, int j, int k) { sedge(i, j, k) = sedgel(i, j, k); });
  } else if (bchi == MGPU_BC_REFLECT_ODD) {
    for_box(b, [&](int i, int j, int k) { sedge(i, j, k) = 0.0; });
  } else if (bchi == MGPU_BC_INTERIOR) {
  } else {
    fail("make_edge_scal: invalid boundary type adv_bc(d,2)");
  }
}

inline Arr mk(const Box& b) { return Arr(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2]); }

}  // namespace

void make_edge_scal_box(const mgpu_params& P, const Arr& sfull, Arr* sedge_full, const Arr* umac,
                        const Arr& force_full, const int* lo, const int* hi, const int* adv_bc, int comp,
                        int bccomp, bool is_vel, bool is_conservative, int ng_s) {
  const int dm = P.dm;
  const double dt = P.dt;
  const double* dx = P.dx;
  const double rel_eps = P.rel_eps;
  const Arr s = sfull.comp(comp);
  const Arr force = force_full.comp(comp);
  EdgeCtx c{dm, lo, hi, adv_bc, bccomp, comp, is_vel, rel_eps};

  Box tb = grown(lo, hi, dm, 1);
  Box vb = grown(lo, hi, dm, 0);

  int bc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  for (int d = 0; d < dm; ++d) {
    bc[d][0] = advbc(adv_bc, dm, d + 1, 1, bccomp);
    bc[d][1] = advbc(adv_bc, dm, d + 1, 2, bccomp);
  }

  Arr divu;
  if (is_conservative && dm == 3) {  // make_edge_scal.f90:763-774 (every term over dx(1))
    divu = mk(tb);
    for_box(tb, [&](int i, int j, int k) {
      divu(i, j, k) = (umac[0](i + 1, j, k) - umac[0](i, j, k) + umac[1](i, j + 1, k) - umac[1](i, j, k) +
                       umac[2](i, j, k + 1) - umac[2](i, j, k)) / dx[0];
    });
  }

  Arr slope[3];
  Arr Ip, Im, Ipf, Imf;
  if (P.ppm_type == 0) {
    for (int d = 0; d < dm; ++d) {
      slope[d] = mk(tb);
      slope_dir(s, slope[d], lo, hi, dm, d, bc[d][0], bc[d][1], P.slope_order);
    }
  } else {
    Ip.alloc(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2], dm);
    Im.alloc(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2], dm);
    ppm(s, umac, Ip, Im, lo, hi, dm, bc, dx, dt, true, P.ppm_type, rel_eps, ng_s);
    if (P.ppm_trace_forces == 1) {
      Ipf.alloc(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2], dm);
      Imf.alloc(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2], dm);
      ppm(force, umac, Ipf, Imf, lo, hi, dm, bc, dx, dt, true, P.ppm_type, rel_eps, ng_s);
    }
  }

  const double dt2 = 0.5 * dt, dt3 = dt / 3.0, dt4 = dt / 4.0, dt6 = dt / 6.0;

  // ---- normal predictor states: lo:hi+1 in d, lo-1:hi+1 transverse (make_edge_scal.f90:807-1059) ----
  Arr sl[3], sr[3], simh[3];
  for (int d = 0; d < dm; ++d) {
    Box fb = tb; fb.lo[d] = lo[d]; fb.hi[d] = hi[d] + 1;
    sl[d] = mk(fb); sr[d] = mk(fb); simh[d] = mk(fb);
    const double h = dx[d];
    if (P.ppm_type == 0) {
      for_box(fb, [&](int i, int j, int k) {
        sl[d](i, j, k) = s.at(i, j, k, d, -1) + (0.5 - dt2 * umac[d](i, j, k) / h) * slope[d].at(i, j, k, d, -1);
        sr[d](i, j, k) = s(i, j, k) - (0.5 + dt2 * umac[d](i, j, k) / h) * slope[d](i, j, k);
      });
    } else {
      for_box(fb, [&](int i, int j, int k) {
        sl[d](i, j, k) = Ip.p[Ip.idx(i - (d == 0), j - (d == 1), k - (d == 2), d)];
        sr[d](i, j, k) = Im(i, j, k, d);
      });
    }
    Box tbx = tb;
    lr_bc(c, d, sl[d], sr[d], s, tbx, 0);
    for_box(fb, [&](int i, int j, int k) {
      simh[d](i, j, k) = riemann(sl[d](i, j, k), sr[d](i, j, k), umac[d](i, j, k), rel_eps);
    });
  }

  auto fl_of = [&](int i, int j, int k, int d) {  // force at the cell left of face (i,j,k)
    return (P.ppm_trace_forces == 0) ? force.at(i, j, k, d, -1)
                                     : Ipf.p[Ipf.idx(i - (d == 0), j - (d == 1), k - (d == 2), d)];
  };
  auto fr_of = [&](int i, int j, int k, int d) {
    return (P.ppm_trace_forces == 0) ? force(i, j, k) : Imf(i, j, k, d);
  };

  if (dm == 2) {
    // ---- final edge states, make_edge_scal.f90:530-670 ----
    for (int d = 0; d < 2; ++d) {
      const int t = 1 - d;
      Box fb = vb; fb.hi[d] = hi[d] + 1;
      Arr sedgel = mk(fb), sedger = mk(fb);
      Arr sedge = sedge_full[d].comp(comp);
      const double hd = dx[d], ht = dx[t];
      for_box(fb, [&](int i, int j, int k) {
        const int il = i - (d == 0), jl = j - (d == 1);  // cell left of the face
        double fl = fl_of(i, j, k, d), fr = fr_of(i, j, k, d);
        const Arr& vt = umac[t];
        const Arr& vd = umac[d];
        const Arr& sh = simh[t];
        if (is_conservative) {
          sedgel(i, j, k) = sl[d](i, j, k) -
                            (dt2 / ht) * (sh.at(il, jl, k, t, 1) * vt.at(il, jl, k, t, 1) - sh(il, jl, k) * vt(il, jl, k)) -
                            (dt2 / hd) * s(il, jl, k) * (vd(i, j, k) - vd(il, jl, k)) + dt2 * fl;
          sedger(i, j, k) = sr[d](i, j, k) -
                            (dt2 / ht) * (sh.at(i, j, k, t, 1) * vt.at(i, j, k, t, 1) - sh(i, j, k) * vt(i, j, k)) -
                            (dt2 / hd) * s(i, j, k) * (vd.at(i, j, k, d, 1) - vd(i, j, k)) + dt2 * fr;
        } else {
          sedgel(i, j, k) = sl[d](i, j, k) -
                            (dt4 / ht) * (vt.at(il, jl, k, t, 1) + vt(il, jl, k)) * (sh.at(il, jl, k, t, 1) - sh(il, jl, k)) +
                            dt2 * fl;
          sedger(i, j, k) = sr[d](i, j, k) -
                            (dt4 / ht) * (vt.at(i, j, k, t, 1) + vt(i, j, k)) * (sh.at(i, j, k, t, 1) - sh(i, j, k)) +
                            dt2 * fr;
        }
        sedge(i, j, k) = riemann(sedgel(i, j, k), sedger(i, j, k), vd(i, j, k), rel_eps);
      });
      final_bc(c, d, sedge, sedgel, sedger, s, vb);
    }
    return;
  }

  // ---- 3-D transverse states simh_{d|t}: d-face state corrected by the t-derivative ----
  // ranges: lo:hi+1 in d, lo:hi in t, lo-1:hi+1 in the remaining direction (make_edge_scal.f90:1065-1683)
  Arr simht[3][3];
  for (int d = 0; d < 3; ++d) {
    for (int t = 0; t < 3; ++t) {
      if (t == d) continue;
      Box fb = tb; fb.lo[d] = lo[d]; fb.hi[d] = hi[d] + 1; fb.lo[t] = lo[t]; fb.hi[t] = hi[t];
      Arr slt = mk(fb), srt = mk(fb);
      simht[d][t] = mk(fb);
      const double ht = dx[t];
      const Arr& vt = umac[t];
      const Arr& sh = simh[t];
      for_box(fb, [&](int i, int j, int k) {
        const int il = i - (d == 0), jl = j - (d == 1), kl = k - (d == 2);
        if (is_conservative) {
          slt(i, j, k) = sl[d](i, j, k) -
                         (dt3 / ht) * (sh.at(il, jl, kl, t, 1) * vt.at(il, jl, kl, t, 1) - sh(il, jl, kl) * vt(il, jl, kl)) -
                         dt3 * s(il, jl, kl) * divu(il, jl, kl) +
                         (dt3 / ht) * s(il, jl, kl) * (vt.at(il, jl, kl, t, 1) - vt(il, jl, kl));
          srt(i, j, k) = sr[d](i, j, k) -
                         (dt3 / ht) * (sh.at(i, j, k, t, 1) * vt.at(i, j, k, t, 1) - sh(i, j, k) * vt(i, j, k)) -
                         dt3 * s(i, j, k) * divu(i, j, k) +
                         (dt3 / ht) * s(i, j, k) * (vt.at(i, j, k, t, 1) - vt(i, j, k));
        } else {
          slt(i, j, k) = sl[d](i, j, k) -
                         (dt6 / ht) * (vt.at(il, jl, kl, t, 1) + vt(il, jl, kl)) * (sh.at(il, jl, kl, t, 1) - sh(il, jl, kl));
          srt(i, j, k) = sr[d](i, j, k) -
                         (dt6 / ht) * (vt.at(i, j, k, t, 1) + vt(i, j, k)) * (sh.at(i, j, k, t, 1) - sh(i, j, k));
        }
      });
      Box tbx = fb;
      lr_bc(c, d, slt, srt, s, tbx, 1);
      for_box(fb, [&](int i, int j, int k) {
        simht[d][t](i, j, k) = riemann(slt(i, j, k), srt(i, j, k), umac[d](i, j, k), rel_eps);
      });
    }
  }

  // ---- final edge states (make_edge_scal.f90:1692-2040) ----
  for (int d = 0; d < 3; ++d) {
    const int t1 = (d == 0) ? 1 : 0;           // lower-index transverse direction
    const int t2 = (d == 2) ? 1 : 2;           // higher-index transverse direction
    Box fb = vb; fb.hi[d] = hi[d] + 1;
    Arr sedgel = mk(fb), sedger = mk(fb);
    Arr sedge = sedge_full[d].comp(comp);
    const double hd = dx[d], h1 = dx[t1], h2 = dx[t2];
    const Arr& v1 = umac[t1];
    const Arr& v2 = umac[t2];
    const Arr& vd = umac[d];
    const Arr& q1 = simht[t1][t2];  // t1-face state corrected by t2 (e.g. simhyz for d=x)
    const Arr& q2 = simht[t2][t1];
    for_box(fb, [&](int i, int j, int k) {
      const int il = i - (d == 0), jl = j - (d == 1), kl = k - (d == 2);
      double fl = fl_of(i, j, k, d), fr = fr_of(i, j, k, d);
      if (is_conservative) {
        sedgel(i, j, k) = sl[d](i, j, k) -
                          (dt2 / h1) * (q1.at(il, jl, kl, t1, 1) * v1.at(il, jl, kl, t1, 1) - q1(il, jl, kl) * v1(il, jl, kl)) -
                          (dt2 / h2) * (q2.at(il, jl, kl, t2, 1) * v2.at(il, jl, kl, t2, 1) - q2(il, jl, kl) * v2(il, jl, kl)) -
                          (dt2 / hd) * s(il, jl, kl) * (vd(i, j, k) - vd(il, jl, kl)) + dt2 * fl;
        sedger(i, j, k) = sr[d](i, j, k) -
                          (dt2 / h1) * (q1.at(i, j, k, t1, 1) * v1.at(i, j, k, t1, 1) - q1(i, j, k) * v1(i, j, k)) -
                          (dt2 / h2) * (q2.at(i, j, k, t2, 1) * v2.at(i, j, k, t2, 1) - q2(i, j, k) * v2(i, j, k)) -
                          (dt2 / hd) * s(i, j, k) * (vd.at(i, j, k, d, 1) - vd(i, j, k)) + dt2 * fr;
      } else {
        sedgel(i, j, k) = sl[d](i, j, k) -
                          (dt4 / h1) * (v1.at(il, jl, kl, t1, 1) + v1(il, jl, kl)) * (q1.at(il, jl, kl, t1, 1) - q1(il, jl, kl)) -
                          (dt4 / h2) * (v2.at(il, jl, kl, t2, 1) + v2(il, jl, kl)) * (q2.at(il, jl, kl, t2, 1) - q2(il, jl, kl)) +
                          dt2 * fl;
        sedger(i, j, k) = sr[d](i, j, k) -
                          (dt4 / h1) * (v1.at(i, j, k, t1, 1) + v1(i, j, k)) * (q1.at(i, j, k, t1, 1) - q1(i, j, k)) -
                          (dt4 / h2) * (v2.at(i, j, k, t2, 1) + v2(i, j, k)) * (q2.at(i, j, k, t2, 1) - q2(i, j, k)) +
                          dt2 * fr;
      }
      sedge(i, j, k) = riemann(sedgel(i, j, k), sedger(i, j, k), vd(i, j, k), rel_eps);
    });
    final_bc(c, d, sedge, sedgel, sedger, s, vb);
  }
}

}  // namespace mo
