// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of the reference's reconstruction kernels:
//   slopes  : Source/slope.f90   slopex_2d :149, slopey_2d :293, slopez_3d :440
//   PPM     : Source/ppm.f90     ppm_2d :544, ppm_3d :1629
// The x/y/z blocks of the reference are index permutations of one another, so
// one direction-generic routine restates them; the places where the reference is NOT symmetric
// are kept as literal special cases (marked QUIRK).
#include "mo_kernels.h"

namespace mo {
// pin_sweep.py only: evaluate ppm_type 1 WITHOUT the "make sure sedge lies in between adjacent cell-centered values"
// clip of ppm.f90:1723-1727 (the variant that reproduces the ppm_type 1 line of the archived report)
bool g_ppm1_no_edge_clip = false;
int g_ppm1_variant = 0;  // pin_sweep.py only: historical ppm_type 1 candidates (0 = the source in the tree)


static const double C_CS = 1.25;  // ppm.f90:1656

// ---------------------------------------------------------------------------------------------
// slope.f90 : limited slopes in direction d on cells lo-1:hi+1 (all dims < dm)
// ---------------------------------------------------------------------------------------------
void slope_dir(const Arr& s, Arr& slp, const int* lo, const int* hi, int dm, int d, int bclo, int bchi,
               int slope_order) {
  Box tb = grown(lo, hi, dm, 1);
  const int is = lo[d], ie = hi[d];
  const bool wlo = (bclo == MGPU_BC_EXT_DIR || bclo == MGPU_BC_HOEXTRAP);
  const bool whi = (bchi == MGPU_BC_EXT_DIR || bchi == MGPU_BC_HOEXTRAP);

  if (slope_order == 0) {  // slope.f90:172-175
    slp.fill(0.0);
    return;
  }
  // iterate over transverse "lines": the box with direction d collapsed to a single index
  Box lines = tb;
  lines.lo[d] = lines.hi[d] = 0;
#pragma omp parallel for collapse(2)
  for (int k = lines.lo[2]; k <= lines.hi[2]; ++k)
    for (int j = lines.lo[1]; j <= lines.hi[1]; ++j)
      for (int i = lines.lo[0]; i <= lines.hi[0]; ++i) {
        // S(q): s along the line; SL(q): slope along the line
        auto S = [&](int q) -> double { return s.at(i, j, k, d, q); };
        auto SL = [&](int q) -> double& { return slp.at(i, j, k, d, q); };
        if (slope_order == 2) {  // slope.f90:177-220
          for (int q = is - 1; q <= ie + 1; ++q) {
            double del = 0.5 * (S(q + 1) - S(q - 1));
            double dpls = 2.0 * (S(q + 1) - S(q));
            double dmn = 2.0 * (S(q) - S(q - 1));
            double slim = dmin(dabs(dpls), dabs(dmn));
            slim = (dpls * dmn > 0.0) ? slim : 0.0;
            double sflag = sign1(del);
            SL(q) = sflag * dmin(slim, dabs(del));
          }
          if (wlo) {
            SL(is - 1) = 0.0;
            double del = (S(is + 1) + 3.0 * S(is) - 4.0 * S(is - 1)) * (1.0 / 3.0);
            double dpls = 2.0 * (S(is + 1) - S(is));
            double dmn = 2.0 * (S(is) - S(is - 1));
            double slim = dmin(dabs(dpls), dabs(dmn));
            slim = (dpls * dmn > 0.0) ? slim : 0.0;
            double sflag = sign1(del);
            SL(is) = sflag * dmin(slim, dabs(del));
          }
          if (whi) {
            SL(ie + 1) = 0.0;
            double del = -(S(ie - 1) + 3.0 * S(ie) - 4.0 * S(ie + 1)) * (1.0 / 3.0);
            double dpls = 2.0 * (S(ie) - S(ie - 1));
            double dmn = 2.0 * (S(ie + 1) - S(ie));
            double slim = dmin(dabs(dpls), dabs(dmn));
            slim = (dpls * dmn > 0.0) ? slim : 0.0;
            double sflag = sign1(del);
            SL(ie) = sflag * dmin(slim, dabs(del));
          }
        } else {  // 4th order, slope.f90:224-287
          const int n = ie - is + 5;
          std::vector<double> cen(n), lim(n), flag(n), fromm(n);
          auto X = [&](std::vector<double>& v, int q) -> double& { return v[q - (is - 2)]; };
          for (int q = is - 2; q <= ie + 2; ++q) {
            X(cen, q) = 0.5 * (S(q + 1) - S(q - 1));
            double dmn = 2.0 * (S(q) - S(q - 1));
            double dpls = 2.0 * (S(q + 1) - S(q));
            X(lim, q) = dmin(dabs(dmn), dabs(dpls));
            X(lim, q) = (dpls * dmn > 0.0) ? X(lim, q) : 0.0;
            X(flag, q) = sign1(X(cen, q));
            X(fromm, q) = X(flag, q) * dmin(X(lim, q), dabs(X(cen, q)));
          }
          const double two3rd = 2.0 / 3.0, sixth = 1.0 / 6.0;
          for (int q = is - 1; q <= ie + 1; ++q) {
            double ds = 2.0 * two3rd * X(cen, q) - sixth * (X(fromm, q + 1) + X(fromm, q - 1));
            SL(q) = X(flag, q) * dmin(dabs(ds), X(lim, q));
          }
          if (wlo) {
            SL(is - 1) = 0.0;
            double del = -16.0 / 15.0 * S(is - 1) + 0.5 * S(is) + two3rd * S(is + 1) - 0.1 * S(is + 2);
            double dmn = 2.0 * (S(is) - S(is - 1));
            double dpls = 2.0 * (S(is + 1) - S(is));
            double slim = dmin(dabs(dpls), dabs(dmn));
            slim = (dpls * dmn > 0.0) ? slim : 0.0;
            double sflag = sign1(del);
            SL(is) = sflag * dmin(slim, dabs(del));
            X(fromm, is) = SL(is);
            double ds = 2.0 * two3rd * X(cen, is + 1) - sixth * (X(fromm, is + 2) + X(fromm, is));
            SL(is + 1) = X(flag, is + 1) * dmin(dabs(ds), X(lim, is + 1));
          }
          if (whi) {
            SL(ie + 1) = 0.0;
            double del = -(-16.0 / 15.0 * S(ie + 1) + 0.5 * S(ie) + two3rd * S(ie - 1) - 0.1 * S(ie - 2));
            double dmn = 2.0 * (S(ie) - S(ie - 1));
            double dpls = 2.0 * (S(ie + 1) - S(ie));
            double slim = dmin(dabs(dpls), dabs(dmn));
            slim = (dpls * dmn > 0.0) ? slim : 0.0;
            double sflag = sign1(del);
            SL(ie) = sflag * dmin(slim, dabs(del));
            X(fromm, ie) = SL(ie);
            double ds = 2.0 * two3rd * X(cen, ie - 1) - sixth * (X(fromm, ie - 2) + X(fromm, ie));
            SL(ie - 1) = X(flag, ie - 1) * dmin(dabs(ds), X(lim, ie - 1));
          }
        }
      }
}

// ---------------------------------------------------------------------------------------------
// ppm.f90 : ppm_2d :544 / ppm_3d :1629.  Ip/Im have dm components (one per direction).
// vel[d] is cell-centred (is_umac=false) or face-centred in d (is_umac=true).
// ---------------------------------------------------------------------------------------------
void ppm(const Arr& s, const Arr* vel, Arr& Ip, Arr& Im, const int* lo, const int* hi, int dm,
         const int bc[3][2], const double* dx, double dt, bool is_umac, int ppm_type, double rel_eps,
         int ng_s) {
  if (ppm_type == 2 && ng_s < 4) fail("Need 4 ghost cells for ppm_type=2");  // ppm.f90:1864

  Box tb = grown(lo, hi, dm, 1);  // cells lo-1:hi+1
  Arr sp(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2]);
  Arr sm(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2]);

  for (int d = 0; d < dm; ++d) {
    const int is = lo[d], ie = hi[d];
    const bool wlo = (bc[d][0] == MGPU_BC_EXT_DIR || bc[d][0] == MGPU_BC_HOEXTRAP);
    const bool whi = (bc[d][1] == MGPU_BC_EXT_DIR || bc[d][1] == MGPU_BC_HOEXTRAP);
    // transverse slab at a fixed index q in direction d
    auto slab = [&](int q) { Box b = tb; b.lo[d] = b.hi[d] = q; return b; };

    // CW84 parabola limiter (ppm.f90:1742-1749) applied in place to cell (i,j,k)
    auto cw_limit = [&](int i, int j, int k) {
      double sc = s(i, j, k);
      double& P = sp(i, j, k);
      double& M = sm(i, j, k);
      if (g_ppm1_variant == 2) return;
      if (g_ppm1_variant == 4) {  // Colella & Woodward 1984 eq. 1.10 as printed
        if ((P - sc) * (sc - M) <= 0.0) { P = sc; M = sc; }
        else {
          const double dq = P - M, q6 = 6.0 * (sc - 0.5 * (M + P));
          if (dq * q6 > dq * dq) M = 3.0 * sc - 2.0 * P;
          else if (-(dq * dq) > dq * q6) P = 3.0 * sc - 2.0 * M;
        }
        return;
      }
      if ((P - sc) * (sc - M) <= 0.0) {
        P = sc;
        M = sc;
      } else if (dabs(P - sc) >= 2.0 * dabs(M - sc)) {
        P = 3.0 * sc - 2.0 * M;
      } else if (dabs(M - sc) >= 2.0 * dabs(P - sc)) {
        M = 3.0 * sc - 2.0 * P;
      }
    };

    if (ppm_type == 1) {
      Box db = tb; db.lo[d] = is - 2; db.hi[d] = ie + 2;
      Arr dsvl(db.lo[0], db.hi[0], db.lo[1], db.hi[1], db.lo[2], db.hi[2]);  // zero (ppm.f90:1693)
      Box eb = tb; eb.lo[d] = is - 1; eb.hi[d] = ie + 2;
      Arr sedge(eb.lo[0], eb.hi[0], eb.lo[1], eb.hi[1], eb.lo[2], eb.hi[2]);

      for_box(db, [&](int i, int j, int k) {  // van Leer slopes, ppm.f90:1697-1709
        double dsc = 0.5 * (s.at(i, j, k, d, 1) - s.at(i, j, k, d, -1));
        double dsl = 2.0 * (s(i, j, k) - s.at(i, j, k, d, -1));
        double dsr = 2.0 * (s.at(i, j, k, d, 1) - s(i, j, k));
        if (dsl * dsr > 0.0) dsvl(i, j, k) = sign1(dsc) * dmin(dmin(dabs(dsc), dabs(dsl)), dabs(dsr));
        if (g_ppm1_variant == 1) dsvl(i, j, k) = dsc;
        if (g_ppm1_variant == 3) dsvl(i, j, k) = (dsl * dsr > 0.0) ? sign1(dsc) * dmin(dmin(dabs(dsc), 0.5 * dabs(dsl)), 0.5 * dabs(dsr)) : 0.0;
      });
      for_box(eb, [&](int i, int j, int k) {  // ppm.f90:1713-1727
        double sl = s.at(i, j, k, d, -1), sc = s(i, j, k);
        double e = 0.5 * (sc + sl) - (1.0 / 6.0) * (dsvl(i, j, k) - dsvl.at(i, j, k, d, -1));
        if (!g_ppm1_no_edge_clip) {
          e = dmax(e, dmin(sc, sl));
          e = dmin(e, dmax(sc, sl));
        }
        sedge(i, j, k) = e;
      });
      for_box(tb, [&](int i, int j, int k) {  // ppm.f90:1731-1752
        sp(i, j, k) = sedge.at(i, j, k, d, 1);
        sm(i, j, k) = sedge(i, j, k);
        cw_limit(i, j, k);
      });
      if (wlo) {  // ppm.f90:1758-1806
        for_box(slab(is), [&](int i, int j, int k) {
          sm(i, j, k) = s.at(i, j, k, d, -1);
          double e = -0.2 * s.at(i, j, k, d, -1) + (3.0 / 4.0) * s(i, j, k) + 0.5 * s.at(i, j, k, d, 1) -
                     (1.0 / 20.0) * s.at(i, j, k, d, 2);
          e = dmax(e, dmin(s.at(i, j, k, d, 1), s(i, j, k)));
          e = dmin(e, dmax(s.at(i, j, k, d, 1), s(i, j, k)));
          sedge.at(i, j, k, d, 1) = e;
          sp(i, j, k) = e;
          sm.at(i, j, k, d, 1) = e;
          sp.at(i, j, k, d, 1) = sedge.at(i, j, k, d, 2);
        });
        for_box(slab(is + 1), [&](int i, int j, int k) { cw_limit(i, j, k); });
      }
      if (whi) {  // ppm.f90:1808-1856
        for_box(slab(ie), [&](int i, int j, int k) {
          sp(i, j, k) = s.at(i, j, k, d, 1);
          double e = -0.2 * s.at(i, j, k, d, 1) + (3.0 / 4.0) * s(i, j, k) + 0.5 * s.at(i, j, k, d, -1) -
                     (1.0 / 20.0) * s.at(i, j, k, d, -2);
          e = dmax(e, dmin(s.at(i, j, k, d, -1), s(i, j, k)));
          e = dmin(e, dmax(s.at(i, j, k, d, -1), s(i, j, k)));
          sedge(i, j, k) = e;
          sp.at(i, j, k, d, -1) = e;
          sm(i, j, k) = e;
          sm.at(i, j, k, d, -1) = sedge.at(i, j, k, d, -1);
        });
        for_box(slab(ie - 1), [&](int i, int j, int k) { cw_limit(i, j, k); });
      }
    } else {  // ppm_type == 2
      Box eb = tb; eb.lo[d] = is - 2; eb.hi[d] = ie + 3;
      Arr sedge(eb.lo[0], eb.hi[0], eb.lo[1], eb.hi[1], eb.lo[2], eb.hi[2]);

      for_box(eb, [&](int i, int j, int k) {  // ppm.f90:1872-1893
        double sm2 = s.at(i, j, k, d, -2), sm1 = s.at(i, j, k, d, -1), s0 = s(i, j, k), sp1 = s.at(i, j, k, d, 1);
        double e = (7.0 / 12.0) * (sm1 + s0) - (1.0 / 12.0) * (sm2 + sp1);
        if ((e - sm1) * (s0 - e) < 0.0) {
          double D2 = 3.0 * (sm1 - 2.0 * e + s0);
          double D2L = sm2 - 2.0 * sm1 + s0;
          double D2R = sm1 - 2.0 * s0 + sp1;
          double sgn = sign1(D2);
          double D2LIM = sgn * dmax(dmin(dmin(C_CS * sgn * D2L, C_CS * sgn * D2R), sgn * D2), 0.0);
          e = 0.5 * (sm1 + s0) - (1.0 / 6.0) * D2LIM;
        }
        sedge(i, j, k) = e;
      });
      // Colella 2008 limiter, ppm.f90:1903-1976 (and the identical re-limit loops at walls)
      auto cs_limit = [&](int i, int j, int k) {
        double sc = s(i, j, k);
        double alphap = sedge.at(i, j, k, d, 1) - sc;
        double alpham = sedge(i, j, k) - sc;
        bool bigp = dabs(alphap) > 2.0 * dabs(alpham);
        bool bigm = dabs(alpham) > 2.0 * dabs(alphap);
        bool extremum = false;
        if (alpham * alphap >= 0.0) {
          extremum = true;
        } else if (bigp || bigm) {
          double dafacem = sedge(i, j, k) - sedge.at(i, j, k, d, -1);
          double dafacep = sedge.at(i, j, k, d, 2) - sedge.at(i, j, k, d, 1);
          double dabarm = sc - s.at(i, j, k, d, -1);
          double dabarp = s.at(i, j, k, d, 1) - sc;
          double dafacemin = dmin(dabs(dafacem), dabs(dafacep));
          double dabarmin = dmin(dabs(dabarm), dabs(dabarp));
          double dachkm, dachkp;
          if (dafacemin >= dabarmin) {
            dachkm = dafacem; dachkp = dafacep;
          } else {
            dachkm = dabarm; dachkp = dabarp;
          }
          extremum = (dachkm * dachkp <= 0.0);
        }
        if (extremum) {
          double D2 = 6.0 * (alpham + alphap);
          double D2L = s.at(i, j, k, d, -2) - 2.0 * s.at(i, j, k, d, -1) + sc;
          double D2R = sc - 2.0 * s.at(i, j, k, d, 1) + s.at(i, j, k, d, 2);
          double D2C = s.at(i, j, k, d, -1) - 2.0 * sc + s.at(i, j, k, d, 1);
          double sgn = sign1(D2);
          double D2LIM = dmax(dmin(dmin(dmin(sgn * D2, C_CS * sgn * D2L), C_CS * sgn * D2R), C_CS * sgn * D2C), 0.0);
          double D2ABS = dmax(dabs(D2), 1.e-10);
          alpham = alpham * D2LIM / D2ABS;
          alphap = alphap * D2LIM / D2ABS;
        } else {
          if (bigp) {
            double sgn = sign1(alpham);
            double amax = -(alphap * alphap) / (4 * (alpham + alphap));
            double delam = s.at(i, j, k, d, -1) - sc;
            if (sgn * amax >= sgn * delam) {
              if (sgn * (delam - alpham) >= 1.e-10) {
                alphap = (-2.0 * delam - 2.0 * sgn * std::sqrt(delam * delam - delam * alpham));
              } else {
                alphap = -2.0 * alpham;
              }
            }
          }
          if (bigm) {
            double sgn = sign1(alphap);
            double amax = -(alpham * alpham) / (4 * (alpham + alphap));
            double delap = s.at(i, j, k, d, 1) - sc;
            if (sgn * amax >= sgn * delap) {
              if (sgn * (delap - alphap) >= 1.e-10) {
                alpham = (-2.0 * delap - 2.0 * sgn * std::sqrt(delap * delap - delap * alphap));
              } else {
                alpham = -2.0 * alphap;
              }
            }
          }
        }
        sm(i, j, k) = sc + alpham;
        sp(i, j, k) = sc + alphap;
      };
      for_box(tb, cs_limit);

      if (wlo) {  // ppm.f90:1983-2097
        for_box(slab(is), [&](int i, int j, int k) {
          sm(i, j, k) = s.at(i, j, k, d, -1);
          sedge(i, j, k) = s.at(i, j, k, d, -1);
          double e = -0.2 * s.at(i, j, k, d, -1) + (3.0 / 4.0) * s(i, j, k) + 0.5 * s.at(i, j, k, d, 1) -
                     (1.0 / 20.0) * s.at(i, j, k, d, 2);
          e = dmax(e, dmin(s.at(i, j, k, d, 1), s(i, j, k)));
          e = dmin(e, dmax(s.at(i, j, k, d, 1), s(i, j, k)));
          sedge.at(i, j, k, d, 1) = e;
          sp(i, j, k) = e;
        });
        // QUIRK ppm.f90:1383: the 2-D y-lo re-limit loop covers lo+1 only (x :869, 3-D :2021,2641,3268 do lo+1..lo+2)
        const int last = (dm == 2 && d == 1) ? is + 1 : is + 2;
        for (int q = is + 1; q <= last; ++q) for_box(slab(q), cs_limit);
      }
      if (whi) {  // ppm.f90:2099-2216
        for_box(slab(ie), [&](int i, int j, int k) {
          sp(i, j, k) = s.at(i, j, k, d, 1);
          // QUIRK ppm.f90:2099-2104: the 3-D x-hi block never resets sedge(hi+1); y :2728, z :3355 and 2-D :947 do
          if (!(dm == 3 && d == 0)) sedge.at(i, j, k, d, 1) = s.at(i, j, k, d, 1);
          double e = -0.2 * s.at(i, j, k, d, 1) + (3.0 / 4.0) * s(i, j, k) + 0.5 * s.at(i, j, k, d, -1) -
                     (1.0 / 20.0) * s.at(i, j, k, d, -2);
          e = dmax(e, dmin(s.at(i, j, k, d, -1), s(i, j, k)));
          e = dmin(e, dmax(s.at(i, j, k, d, -1), s(i, j, k)));
          sedge(i, j, k) = e;
          sm(i, j, k) = e;
        });
        for (int q = ie - 2; q <= ie - 1; ++q) for_box(slab(q), cs_limit);
      }
    }

    // traced integrals, ppm.f90:2225-2287
    const Arr& u = vel[d];
    const double hd = dx[d];
    Box pb = tb; pb.lo[d] = is - 1; pb.hi[d] = ie;
    for_box(pb, [&](int i, int j, int k) {
      double uu = is_umac ? u.at(i, j, k, d, 1) : u(i, j, k);
      double sigma = dabs(uu) * dt / hd;
      double s6 = 6.0 * s(i, j, k) - 3.0 * (sm(i, j, k) + sp(i, j, k));
      if (uu > rel_eps) {
        Ip(i, j, k, d) = sp(i, j, k) - (sigma / 2.0) * (sp(i, j, k) - sm(i, j, k) - (1.0 - (2.0 / 3.0) * sigma) * s6);
      } else {
        Ip(i, j, k, d) = s(i, j, k);
      }
    });
    Box mb = tb; mb.lo[d] = is; mb.hi[d] = ie + 1;
    for_box(mb, [&](int i, int j, int k) {
      double uu = u(i, j, k);
      double sigma = dabs(uu) * dt / hd;
      double s6 = 6.0 * s(i, j, k) - 3.0 * (sm(i, j, k) + sp(i, j, k));
      if (uu < -rel_eps) {
        Im(i, j, k, d) = sm(i, j, k) + (sigma / 2.0) * (sp(i, j, k) - sm(i, j, k) + (1.0 - (2.0 / 3.0) * sigma) * s6);
      } else {
        Im(i, j, k, d) = s(i, j, k);
      }
    });
  }
}

}  // namespace mo
