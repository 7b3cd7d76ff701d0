// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of Source/mkutrans.f90 (mkutrans_2d :257, mkutrans_3d :461) and Source/velpred.f90
// (velpred_2d :266, velpred_3d :640), planar geometry.  Same stages, temporaries (index ranges) and
// expression order as the reference.  The x/y/z blocks are index permutations of one another and are
// produced by direction-generic loops; the reference's asymmetries are kept literally (QUIRK).
#include "mo_kernels.h"

namespace mo {

namespace {

inline Arr mk(const Box& b, int nc = 1) { return Arr(b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2], nc); }
inline int pbc(const int* phys_bc, int dm, int d, int side) { return phys_bc[d + dm * side]; }  // phys_bc(d+1,side+1)

// index of the cell left of face (i,j,k) in direction d, and shifts
struct I3 {
  int i, j, k;
};
inline I3 sh(int i, int j, int k, int d, int o) { return I3{i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0)}; }

// Riemann problem with the full velocity (mkutrans.f90:618-631 / :834-847, velpred.f90:1590-1621):
// w0 enters only in the radial direction of planar geometry; elsewhere the reference omits it.
inline double riemann_full(double l, double r, bool radial, double w0, double rel_eps) {
  const double uavg = 0.5 * (l + r);
  bool test;
  double v;
  if (radial) {
    test = ((l + w0 <= 0.0 && r + w0 >= 0.0) || (dabs(l + r + 2.0 * w0) < rel_eps));
    v = (uavg + w0 > 0.0) ? l : r;
  } else {
    test = ((l <= 0.0 && r >= 0.0) || (dabs(l + r) < rel_eps));
    v = (uavg > 0.0) ? l : r;
  }
  return test ? 0.0 : v;
}

// upwind by the transverse velocity (velpred.f90:895-897)
inline double upwind_trans(double l, double r, double ut, double rel_eps) {
  double v = (ut > 0.0) ? l : r;
  double uavg = 0.5 * (l + r);
  return (dabs(ut) < rel_eps) ? uavg : v;
}

inline bool is_wall3(int bc) { return bc == MGPU_BC_SLIP_WALL || bc == MGPU_BC_NO_SLIP_WALL || bc == MGPU_BC_SYMMETRY; }

void check_phys(int bc, const char* who) {
  if (bc == MGPU_BC_INLET || is_wall3(bc) || bc == MGPU_BC_OUTLET || bc == MGPU_BC_INTERIOR || bc == MGPU_BC_PERIODIC)
    return;
  fail(who);
}

void comp_bc(const int* adv_bc, int dm, int c, int bc[3][2]) {
  for (int q = 0; q < 3; ++q) bc[q][0] = bc[q][1] = MGPU_BC_INTERIOR;
  for (int q = 0; q < dm; ++q) {
    bc[q][0] = advbc(adv_bc, dm, q + 1, 1, c + 1);
    bc[q][1] = advbc(adv_bc, dm, q + 1, 2, c + 1);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
void mkutrans_box(const mgpu_params& P, const Arr& utilde, const Arr& ufull, Arr* utrans, const double* w0,
                  const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u, const Arr* w0mac) {
  const int dm = P.dm;
  if (P.spherical && !w0mac) fail("mkutrans: spherical geometry needs w0mac");
  const double dt = P.dt, dt2 = 0.5 * dt, rel_eps = P.rel_eps;
  Box tb = grown(lo, hi, dm, 1), vb = grown(lo, hi, dm, 0);
  Arr vel[3];
  for (int q = 0; q < dm; ++q) vel[q] = ufull.comp(q);

  for (int d = 0; d < dm; ++d) {
    const Arr u = utilde.comp(d);
    const Arr uf = ufull.comp(d);
    int bc[3][2];
    comp_bc(adv_bc, dm, d, bc);
    const double h = P.dx[d];
    const int is = lo[d], ie = hi[d];
    Box fb = vb;
    fb.hi[d] = ie + 1;
    Arr ul = mk(fb), ur = mk(fb);
    if (P.ppm_type == 0) {
      Arr slp = mk(tb);
      slope_dir(u, slp, lo, hi, dm, d, bc[d][0], bc[d][1], P.slope_order);
      for_box(fb, [&](int i, int j, int k) {  // mkutrans.f90:543-547
        ul(i, j, k) = u.at(i, j, k, d, -1) + (0.5 - (dt2 / h) * dmax(0.0, uf.at(i, j, k, d, -1))) * slp.at(i, j, k, d, -1);
        ur(i, j, k) = u(i, j, k) - (0.5 + (dt2 / h) * dmin(0.0, uf(i, j, k))) * slp(i, j, k);
      });
    } else {
      Arr Ip = mk(tb, dm), Im = mk(tb, dm);
      ppm(u, vel, Ip, Im, lo, hi, dm, bc, P.dx, dt, false, P.ppm_type, rel_eps, ng_u);
      for_box(fb, [&](int i, int j, int k) {
        I3 l = sh(i, j, k, d, -1);
        ul(i, j, k) = Ip(l.i, l.j, l.k, d);
        ur(i, j, k) = Im(i, j, k, d);
      });
    }
    const int plo = pbc(phys_bc, dm, d, 0), phi = pbc(phys_bc, dm, d, 1);
    check_phys(plo, "mkutrans: invalid boundary type phys_bc(d,1)");
    check_phys(phi, "mkutrans: invalid boundary type phys_bc(d,2)");
    Box b = fb;
    b.lo[d] = b.hi[d] = is;
    if (plo == MGPU_BC_INLET) {
      for_box(b, [&](int i, int j, int k) { ul(i, j, k) = u.at(i, j, k, d, -1); ur(i, j, k) = u.at(i, j, k, d, -1); });
    } else if (is_wall3(plo)) {
      for_box(b, [&](int i, int j, int k) { ul(i, j, k) = 0.0; ur(i, j, k) = 0.0; });
    } else if (plo == MGPU_BC_OUTLET) {
      for_box(b, [&](int i, int j, int k) { ul(i, j, k) = dmin(ur(i, j, k), 0.0); ur(i, j, k) = ul(i, j, k); });
    }
    b.lo[d] = b.hi[d] = ie + 1;
    if (phi == MGPU_BC_INLET) {
      for_box(b, [&](int i, int j, int k) { ul(i, j, k) = u(i, j, k); ur(i, j, k) = u(i, j, k); });
    } else if (is_wall3(phi)) {
      for_box(b, [&](int i, int j, int k) { ul(i, j, k) = 0.0; ur(i, j, k) = 0.0; });
    } else if (phi == MGPU_BC_OUTLET) {
      for_box(b, [&](int i, int j, int k) { ul(i, j, k) = dmax(ul(i, j, k), 0.0); ur(i, j, k) = ul(i, j, k); });
    }
    const bool radial = (d == dm - 1);
    Arr& out = utrans[d];
    for_box(fb, [&](int i, int j, int k) {
      const int ir = (d == 0) ? i : (d == 1 ? j : k);
      if (P.spherical)  // mkutrans.f90:601-613 (x), :709, :817: every direction carries its w0mac
        out(i, j, k) = riemann_full(ul(i, j, k), ur(i, j, k), true, w0mac[d](i, j, k), rel_eps);
      else
        out(i, j, k) = riemann_full(ul(i, j, k), ur(i, j, k), radial, radial ? w0[ir] : 0.0, rel_eps);
    });
  }
}

// ---------------------------------------------------------------------------------------------
void velpred_box(const mgpu_params& P, const Arr& utilde, const Arr& ufull, Arr* umac, const Arr* utrans,
                 const Arr& force, const double* w0, const int* lo, const int* hi, const int* adv_bc,
                 const int* phys_bc, int ng_u, const Arr* w0mac) {
  const int dm = P.dm;
  if (P.spherical && !w0mac) fail("velpred: spherical geometry needs w0mac");
  const double dt = P.dt, dt2 = 0.5 * dt, dt4 = dt / 4.0, dt6 = dt / 6.0, rel_eps = P.rel_eps;
  const bool trace = (P.ppm_trace_forces == 1) && P.ppm_type != 0;
  Box tb = grown(lo, hi, dm, 1), vb = grown(lo, hi, dm, 0);
  Arr vel[3];
  for (int q = 0; q < dm; ++q) vel[q] = ufull.comp(q);
  // ghost width of force: 2-D passes ng_f, 3-D passes ng_u to ppm_3d (velpred.f90:366-371 vs :783-791)
  const int ng_f = lo[0] - force.lo[0];

  Arr slope[3][3];      // [d][c]
  Arr Ip[3], Im[3];     // [c], dm components (directions)
  Arr Ipf[3], Imf[3];
  for (int c = 0; c < dm; ++c) {
    int bc[3][2];
    comp_bc(adv_bc, dm, c, bc);
    if (P.ppm_type == 0) {
      for (int d = 0; d < dm; ++d) {
        slope[d][c] = mk(tb);
        slope_dir(utilde.comp(c), slope[d][c], lo, hi, dm, d, bc[d][0], bc[d][1], P.slope_order);
      }
    } else {
      Ip[c] = mk(tb, dm);
      Im[c] = mk(tb, dm);
      ppm(utilde.comp(c), vel, Ip[c], Im[c], lo, hi, dm, bc, P.dx, dt, false, P.ppm_type, rel_eps, ng_u);
      if (trace) {
        Ipf[c] = mk(tb, dm);
        Imf[c] = mk(tb, dm);
        ppm(force.comp(c), vel, Ipf[c], Imf[c], lo, hi, dm, bc, P.dx, dt, false, P.ppm_type, rel_eps,
            dm == 2 ? ng_f : ng_u);
      }
    }
  }

  // ---- normal predictor states for all components on every face direction -------------------
  Arr ul[3], ur[3], uimh[3];
  for (int d = 0; d < dm; ++d) {
    Box fb = tb;
    fb.lo[d] = lo[d];
    fb.hi[d] = hi[d] + 1;
    ul[d] = mk(fb, dm);
    ur[d] = mk(fb, dm);
    uimh[d] = mk(fb, dm);
    const double h = P.dx[d];
    const Arr ufd = ufull.comp(d);
    if (P.ppm_type == 0) {
      for_box(fb, [&](int i, int j, int k) {
        I3 l = sh(i, j, k, d, -1);
        double maxu, minu;
        if (dm == 2) {  // velpred.f90:384-392
          maxu = 0.5 - (dt2 / h) * dmax(0.0, ufd(l.i, l.j, l.k));
          minu = 0.5 + (dt2 / h) * dmin(0.0, ufd(i, j, k));
        } else {  // velpred.f90:812-813
          maxu = 0.5 - dt2 * dmax(0.0, ufd(l.i, l.j, l.k)) / h;
          minu = 0.5 + dt2 * dmin(0.0, ufd(i, j, k)) / h;
        }
        for (int c = 0; c < dm; ++c) {
          ul[d](i, j, k, c) = utilde(l.i, l.j, l.k, c) + maxu * slope[d][c](l.i, l.j, l.k);
          ur[d](i, j, k, c) = utilde(i, j, k, c) - minu * slope[d][c](i, j, k);
        }
      });
    } else {
      for_box(fb, [&](int i, int j, int k) {
        I3 l = sh(i, j, k, d, -1);
        for (int c = 0; c < dm; ++c) {
          ul[d](i, j, k, c) = Ip[c](l.i, l.j, l.k, d);
          ur[d](i, j, k, c) = Im[c](i, j, k, d);
        }
      });
    }
    const int plo = pbc(phys_bc, dm, d, 0), phi = pbc(phys_bc, dm, d, 1);
    check_phys(plo, "velpred: invalid boundary type phys_bc(d,1)");
    check_phys(phi, "velpred: invalid boundary type phys_bc(d,2)");
    Box b = fb;
    b.lo[d] = b.hi[d] = lo[d];
    for_box(b, [&](int i, int j, int k) {
      I3 l = sh(i, j, k, d, -1);
      if (plo == MGPU_BC_INLET) {
        for (int c = 0; c < dm; ++c) ul[d](i, j, k, c) = ur[d](i, j, k, c) = utilde(l.i, l.j, l.k, c);
      } else if (plo == MGPU_BC_SLIP_WALL || plo == MGPU_BC_SYMMETRY) {
        for (int c = 0; c < dm; ++c) {
          if (c == d) ul[d](i, j, k, c) = ur[d](i, j, k, c) = 0.0;
          else ul[d](i, j, k, c) = ur[d](i, j, k, c);
        }
      } else if (plo == MGPU_BC_NO_SLIP_WALL) {
        for (int c = 0; c < dm; ++c) ul[d](i, j, k, c) = ur[d](i, j, k, c) = 0.0;
      } else if (plo == MGPU_BC_OUTLET) {
        ur[d](i, j, k, d) = dmin(ur[d](i, j, k, d), 0.0);
        if (d == 0 && dm == 2) {
          // QUIRK velpred.f90:415-417: the 2-D x-lo branch copies the wrong way (urx = ulx)
          for (int c = 0; c < dm; ++c) ur[d](i, j, k, c) = ul[d](i, j, k, c);
        } else if (d == 0 && dm == 3) {
          // QUIRK velpred.f90:861-862: the 3-D x-lo branch is a self-assignment (ulx = ulx)
        } else {
          for (int c = 0; c < dm; ++c) ul[d](i, j, k, c) = ur[d](i, j, k, c);
        }
      }
    });
    b.lo[d] = b.hi[d] = hi[d] + 1;
    for_box(b, [&](int i, int j, int k) {
      if (phi == MGPU_BC_INLET) {
        for (int c = 0; c < dm; ++c) ul[d](i, j, k, c) = ur[d](i, j, k, c) = utilde(i, j, k, c);
      } else if (phi == MGPU_BC_SLIP_WALL || phi == MGPU_BC_SYMMETRY) {
        for (int c = 0; c < dm; ++c) {
          if (c == d) ul[d](i, j, k, c) = ur[d](i, j, k, c) = 0.0;
          else ur[d](i, j, k, c) = ul[d](i, j, k, c);
        }
      } else if (phi == MGPU_BC_NO_SLIP_WALL) {
        for (int c = 0; c < dm; ++c) ul[d](i, j, k, c) = ur[d](i, j, k, c) = 0.0;
      } else if (phi == MGPU_BC_OUTLET) {
        ul[d](i, j, k, d) = dmax(ul[d](i, j, k, d), 0.0);
        for (int c = 0; c < dm; ++c) ur[d](i, j, k, c) = ul[d](i, j, k, c);
      }
    });
    // transverse components upwinded by utrans (which already contains w0), velpred.f90:891-908
    for_box(fb, [&](int i, int j, int k) {
      for (int c = 0; c < dm; ++c) {
        if (c == d) continue;
        uimh[d](i, j, k, c) = upwind_trans(ul[d](i, j, k, c), ur[d](i, j, k, c), utrans[d](i, j, k), rel_eps);
      }
    });
  }

  auto fl_of = [&](int i, int j, int k, int d) {  // force at the cell left of face, component d
    I3 l = sh(i, j, k, d, -1);
    return trace ? Ipf[d](l.i, l.j, l.k, d) : force(l.i, l.j, l.k, d);
  };
  auto fr_of = [&](int i, int j, int k, int d) { return trace ? Imf[d](i, j, k, d) : force(i, j, k, d); };

  // transverse difference term of cell (ci,cj,ck) in direction t for the state array q(comp qc)
  auto tterm = [&](double coef, const Arr& q, int qc, int ci, int cj, int ck, int t) {
    I3 p = sh(ci, cj, ck, t, 1);
    return coef * (utrans[t](p.i, p.j, p.k) + utrans[t](ci, cj, ck)) * (q(p.i, p.j, p.k, qc) - q(ci, cj, ck, qc));
  };

  // BCs + Riemann solve of the MAC velocity on d-faces (velpred.f90:1590-1651)
  auto finish = [&](int d, const Arr& ml, const Arr& mr, const Box& fb) {
    Arr& out = umac[d];
    const bool radial = (d == dm - 1);
    for_box(fb, [&](int i, int j, int k) {
      const int ir = (d == 0) ? i : (d == 1 ? j : k);
      if (P.spherical)  // velpred.f90:1588-1603 (x), :1687, :1786
        out(i, j, k) = riemann_full(ml(i, j, k), mr(i, j, k), true, w0mac[d](i, j, k), rel_eps);
      else
        out(i, j, k) = riemann_full(ml(i, j, k), mr(i, j, k), radial, radial ? w0[ir] : 0.0, rel_eps);
    });
    const int plo = pbc(phys_bc, dm, d, 0), phi = pbc(phys_bc, dm, d, 1);
    Box b = fb;
    b.lo[d] = b.hi[d] = lo[d];
    if (plo == MGPU_BC_INLET) for_box(b, [&](int i, int j, int k) { I3 l = sh(i, j, k, d, -1); out(i, j, k) = utilde(l.i, l.j, l.k, d); });
    else if (is_wall3(plo)) for_box(b, [&](int i, int j, int k) { out(i, j, k) = 0.0; });
    else if (plo == MGPU_BC_OUTLET) for_box(b, [&](int i, int j, int k) { out(i, j, k) = dmin(mr(i, j, k), 0.0); });
    b.lo[d] = b.hi[d] = hi[d] + 1;
    if (phi == MGPU_BC_INLET) for_box(b, [&](int i, int j, int k) { out(i, j, k) = utilde(i, j, k, d); });
    else if (is_wall3(phi)) for_box(b, [&](int i, int j, int k) { out(i, j, k) = 0.0; });
    else if (phi == MGPU_BC_OUTLET) for_box(b, [&](int i, int j, int k) { out(i, j, k) = dmax(ml(i, j, k), 0.0); });
  };

  if (dm == 2) {  // velpred.f90:528-630
    for (int d = 0; d < 2; ++d) {
      const int t = 1 - d;
      Box fb = vb;
      fb.hi[d] = hi[d] + 1;
      Arr ml = mk(fb), mr = mk(fb);
      const double ht = P.dx[t];
      for_box(fb, [&](int i, int j, int k) {
        I3 l = sh(i, j, k, d, -1);
        ml(i, j, k) = ul[d](i, j, k, d) - tterm(dt4 / ht, uimh[t], d, l.i, l.j, l.k, t) + dt2 * fl_of(i, j, k, d);
        mr(i, j, k) = ur[d](i, j, k, d) - tterm(dt4 / ht, uimh[t], d, i, j, k, t) + dt2 * fr_of(i, j, k, d);
      });
      finish(d, ml, mr, fb);
    }
    return;
  }

  // ---- 3-D corner-coupled transverse states Q[c][d]: component c on d-faces corrected by the third
  //      direction t (velpred.f90:1139-1558: uimhyz/uimhzy, vimhxz/vimhzx, wimhxy/wimhyx) -----------
  Arr Q[3][3];
  for (int c = 0; c < 3; ++c)
    for (int d = 0; d < 3; ++d) {
      if (d == c) continue;
      const int t = 3 - c - d;
      Box fb = tb;  // lo-1:hi+1 in c
      fb.lo[d] = lo[d]; fb.hi[d] = hi[d] + 1;
      fb.lo[t] = lo[t]; fb.hi[t] = hi[t];
      Arr ql = mk(fb), qr = mk(fb);
      Q[c][d] = mk(fb);
      const double ht = P.dx[t];
      for_box(fb, [&](int i, int j, int k) {
        I3 l = sh(i, j, k, d, -1);
        ql(i, j, k) = ul[d](i, j, k, c) - tterm(dt6 / ht, uimh[t], c, l.i, l.j, l.k, t);
        qr(i, j, k) = ur[d](i, j, k, c) - tterm(dt6 / ht, uimh[t], c, i, j, k, t);
      });
      const int plo = pbc(phys_bc, dm, d, 0), phi = pbc(phys_bc, dm, d, 1);
      Box b = fb;
      b.lo[d] = b.hi[d] = lo[d];
      if (plo == MGPU_BC_INLET)
        for_box(b, [&](int i, int j, int k) { I3 l = sh(i, j, k, d, -1); ql(i, j, k) = qr(i, j, k) = utilde(l.i, l.j, l.k, c); });
      else if (plo == MGPU_BC_SLIP_WALL || plo == MGPU_BC_SYMMETRY || plo == MGPU_BC_OUTLET)
        for_box(b, [&](int i, int j, int k) { ql(i, j, k) = qr(i, j, k); });
      else if (plo == MGPU_BC_NO_SLIP_WALL)
        for_box(b, [&](int i, int j, int k) { ql(i, j, k) = qr(i, j, k) = 0.0; });
      b.lo[d] = b.hi[d] = hi[d] + 1;
      if (phi == MGPU_BC_INLET)
        for_box(b, [&](int i, int j, int k) { ql(i, j, k) = qr(i, j, k) = utilde(i, j, k, c); });
      else if (phi == MGPU_BC_SLIP_WALL || phi == MGPU_BC_SYMMETRY || phi == MGPU_BC_OUTLET)
        for_box(b, [&](int i, int j, int k) { qr(i, j, k) = ql(i, j, k); });
      else if (phi == MGPU_BC_NO_SLIP_WALL)
        for_box(b, [&](int i, int j, int k) { ql(i, j, k) = qr(i, j, k) = 0.0; });
      for_box(fb, [&](int i, int j, int k) {
        Q[c][d](i, j, k) = upwind_trans(ql(i, j, k), qr(i, j, k), utrans[d](i, j, k), rel_eps);
      });
    }

  // ---- MAC velocities (velpred.f90:1562-1851) -------------------------------------------------
  for (int d = 0; d < 3; ++d) {
    const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
    Box fb = vb;
    fb.hi[d] = hi[d] + 1;
    Arr ml = mk(fb), mr = mk(fb);
    const double h1 = P.dx[t1], h2 = P.dx[t2];
    for_box(fb, [&](int i, int j, int k) {
      I3 l = sh(i, j, k, d, -1);
      ml(i, j, k) = ul[d](i, j, k, d) - tterm(dt4 / h1, Q[d][t1], 0, l.i, l.j, l.k, t1) -
                    tterm(dt4 / h2, Q[d][t2], 0, l.i, l.j, l.k, t2) + dt2 * fl_of(i, j, k, d);
      mr(i, j, k) = ur[d](i, j, k, d) - tterm(dt4 / h1, Q[d][t1], 0, i, j, k, t1) -
                    tterm(dt4 / h2, Q[d][t2], 0, i, j, k, t2) + dt2 * fr_of(i, j, k, d);
    });
    finish(d, ml, mr, fb);
  }
}

}  // namespace mo
