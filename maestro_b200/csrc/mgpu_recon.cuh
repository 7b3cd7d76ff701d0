// Per-cell reconstruction device functions (fp64) shared by every kernel of the path.
//
// These are *functional* forms of the reference's array sweeps: each returns the parabola
// end-points (or the limited slope) of ONE cell from the values of its line neighbours, including
// the modified stencils next to EXT_DIR/HOEXTRAP walls, so that a thread can evaluate any cell it
// needs without intermediate arrays in HBM.  Expression order follows the reference so that, built
// with -fmad=false, results are bit-identical to the Fortran expression trees:
//   slopes : Source/slope.f90:149-291 (x), :293 (y), :440 (z)
//   PPM    : Source/ppm.f90:1629-2290 (3-D x block; y/z and 2-D are permutations, QUIRKs below)
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {

__device__ __forceinline__ double sign1(double x) { return copysign(1.0, x); }
// compare + select (DSETP + 2 FSEL): sm_100a has no fp64 min/max instruction -- fmin()/fmax() expand to a
// NaN-aware sequence three times as long (measured: profiles/r01b_notes.md)
__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax2(double a, double b) { return a > b ? a : b; }
// (min, max) of a pair from a single compare: same values as dmin2/dmax2
__device__ __forceinline__ void dminmax(double a, double b, double& lo, double& hi) {
  const bool c = a < b;
  lo = c ? a : b;
  hi = c ? b : a;
}

// what a line needs to know about its direction
struct LineBC {
  int lo, hi;     // valid cell range of the box in this direction
  bool wlo, whi;  // EXT_DIR or HOEXTRAP at the lo / hi end (modified stencils)
  // QUIRK ppm.f90:1383: in 2-D the y-lo re-limit loop of ppm_type 2 covers lo+1 only
  int relimit_last;  // lo+2 normally, lo+1 for (dm==2, d==y)
  // QUIRK ppm.f90:2099-2104: the 3-D x-hi block of ppm_type 2 never resets sedge(hi+1)
  bool hi_reset;
};

__host__ __device__ inline LineBC make_linebc(int dm, int d, int lo, int hi, int bclo, int bchi) {
  LineBC b;
  b.lo = lo;
  b.hi = hi;
  b.wlo = (bclo == MGPU_BC_EXT_DIR || bclo == MGPU_BC_HOEXTRAP);
  b.whi = (bchi == MGPU_BC_EXT_DIR || bchi == MGPU_BC_HOEXTRAP);
  b.relimit_last = (dm == 2 && d == 1) ? lo + 1 : lo + 2;
  b.hi_reset = !(dm == 3 && d == 0);
  return b;
}

// ---------------------------------------------------------------------------------------------
// ppm_type = 0 : limited slope of cell c.   q points at s(c); st = stride of the direction.
// ---------------------------------------------------------------------------------------------
struct Fromm {
  double cen, lim, flag, fromm;
};
__device__ __forceinline__ Fromm fromm_of(const double* q, long st) {  // slope.f90:228-237
  Fromm f;
  f.cen = 0.5 * (q[st] - q[-st]);
  double dmn = 2.0 * (q[0] - q[-st]);
  double dpls = 2.0 * (q[st] - q[0]);
  f.lim = dmin2(fabs(dmn), fabs(dpls));
  f.lim = (dpls * dmn > 0.0) ? f.lim : 0.0;
  f.flag = sign1(f.cen);
  f.fromm = f.flag * dmin2(f.lim, fabs(f.cen));
  return f;
}
// one-sided slope in the first cell next to a wall; sg=+1 at lo (neighbours c+1,c+2 inside), -1 at hi
__device__ __forceinline__ double wall_slope(const double* q, long st, int order, int sg) {
  // at lo: ghost = q[-st], inside = q[+st], q[+2st];   at hi mirror
  const double g = q[-sg * st], c0 = q[0], c1 = q[sg * st];
  double del;
  if (order == 2) {  // slope.f90:195-196 / :209-210
    del = (c1 + 3.0 * c0 - 4.0 * g) * (1.0 / 3.0);
    if (sg < 0) del = -del;
  } else {  // slope.f90:249-250 / :270-271
    const double c2 = q[2 * sg * st];
    del = -16.0 / 15.0 * g + 0.5 * c0 + (2.0 / 3.0) * c1 - 0.1 * c2;
    if (sg < 0) del = -del;
  }
  double dpls = 2.0 * (q[st] - q[0]);
  double dmn = 2.0 * (q[0] - q[-st]);
  double slim = dmin2(fabs(dpls), fabs(dmn));
  slim = (dpls * dmn > 0.0) ? slim : 0.0;
  double sflag = sign1(del);
  return sflag * dmin2(slim, fabs(del));
}
__device__ __forceinline__ double slope_cell(const double* q, long st, int c, const LineBC& b, int order) {
  if (order == 0) return 0.0;
  if (b.wlo) {
    if (c == b.lo - 1) return 0.0;
    if (c == b.lo) return wall_slope(q, st, order, +1);
  }
  if (b.whi) {
    if (c == b.hi + 1) return 0.0;
    if (c == b.hi) return wall_slope(q, st, order, -1);
  }
  if (order == 2) {  // slope.f90:183-189
    double del = 0.5 * (q[st] - q[-st]);
    double dpls = 2.0 * (q[st] - q[0]);
    double dmn = 2.0 * (q[0] - q[-st]);
    double slim = dmin2(fabs(dpls), fabs(dmn));
    slim = (dpls * dmn > 0.0) ? slim : 0.0;
    double sflag = sign1(del);
    return sflag * dmin2(slim, fabs(del));
  }
  // 4th order, slope.f90:239-243, with the revised neighbour next to walls (:258-262, :279-283)
  Fromm f0 = fromm_of(q, st);
  double fp = (b.whi && c == b.hi - 1) ? wall_slope(q + st, st, order, -1) : fromm_of(q + st, st).fromm;
  double fm = (b.wlo && c == b.lo + 1) ? wall_slope(q - st, st, order, +1) : fromm_of(q - st, st).fromm;
  double ds;
  if (b.whi && c == b.hi - 1)
    ds = 2.0 * (2.0 / 3.0) * f0.cen - (1.0 / 6.0) * (fm + fp);  // (dxscr(ie-2,fromm) + dxscr(ie,fromm))
  else
    ds = 2.0 * (2.0 / 3.0) * f0.cen - (1.0 / 6.0) * (fp + fm);  // (dxscr(i+1,fromm) + dxscr(i-1,fromm))
  return f0.flag * dmin2(fabs(ds), f0.lim);
}

// ---------------------------------------------------------------------------------------------
// ppm_type = 1 (Colella & Woodward 84)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dsvl_of(const double* q, long st) {  // ppm.f90:1703-1706
  double dsc = 0.5 * (q[st] - q[-st]);
  double dsl = 2.0 * (q[0] - q[-st]);
  double dsr = 2.0 * (q[st] - q[0]);
  // sign(1,dsc)*min(...) with min(...) >= 0: multiplying by +-1 is exact, so copysign gives the same bits
  return (dsl * dsr > 0.0) ? copysign(dmin2(dmin2(fabs(dsc), fabs(dsl)), fabs(dsr)), dsc) : 0.0;
}
// edge value between cells f-1 and f; q points at s(f).  ppm.f90:1719-1724
__device__ __forceinline__ double sedge1_of(const double* q, long st) {
  double e = 0.5 * (q[0] + q[-st]) - (1.0 / 6.0) * (dsvl_of(q, st) - dsvl_of(q - st, st));
  double lo, hi;
  dminmax(q[0], q[-st], lo, hi);
  e = dmax2(e, lo);
  e = dmin2(e, hi);
  return e;
}
// modified stencil on the first interior edge next to a wall: ppm.f90:1773-1781 (lo), :1823-1831 (hi)
// qw points at the first interior cell (lo or hi), sg = +1 at lo, -1 at hi
__device__ __forceinline__ double sedge_wall(const double* qw, long st, int sg) {
  double e = -0.2 * qw[-sg * st] + (3.0 / 4.0) * qw[0] + 0.5 * qw[sg * st] - (1.0 / 20.0) * qw[2 * sg * st];
  if (sg > 0) {
    e = dmax2(e, dmin2(qw[st], qw[0]));
    e = dmin2(e, dmax2(qw[st], qw[0]));
  } else {
    e = dmax2(e, dmin2(qw[-st], qw[0]));
    e = dmin2(e, dmax2(qw[-st], qw[0]));
  }
  return e;
}
__device__ __forceinline__ void cw_limit(double sc, double& sm, double& sp) {  // ppm.f90:1742-1749
  // the reference's if / else if / else if chain written with selects (no divergent branches): same values
  const double dp = sp - sc, dm = sm - sc;
  const bool flat = dp * (sc - sm) <= 0.0;
  const bool c1 = fabs(dp) >= 2.0 * fabs(dm);
  const bool c2 = fabs(dm) >= 2.0 * fabs(dp);
  const double sp1 = 3.0 * sc - 2.0 * sm, sm1 = 3.0 * sc - 2.0 * sp;
  sp = flat ? sc : (c1 ? sp1 : sp);
  sm = flat ? sc : ((!c1 && c2) ? sm1 : sm);
}
__device__ __forceinline__ void ppm1_cell(const double* q, long st, int c, const LineBC& b, double& sm, double& sp) {
  if (b.wlo) {
    if (c == b.lo) {  // ppm.f90:1762,1785: no limiter on this cell
      sm = q[-st];
      sp = sedge_wall(q, st, +1);
      return;
    }
    if (c == b.lo + 1) {  // :1786-1801
      sm = sedge_wall(q - st, st, +1);
      sp = sedge1_of(q + st, st);
      cw_limit(q[0], sm, sp);
      return;
    }
  }
  if (b.whi) {
    if (c == b.hi) {  // :1812,1836
      sp = q[st];
      sm = sedge_wall(q, st, -1);
      return;
    }
    if (c == b.hi - 1) {  // :1835-1851
      sp = sedge_wall(q + st, st, -1);
      sm = sedge1_of(q, st);
      cw_limit(q[0], sm, sp);
      return;
    }
  }
  sm = sedge1_of(q, st);
  sp = sedge1_of(q + st, st);
  cw_limit(q[0], sm, sp);
}

// ---------------------------------------------------------------------------------------------
// ppm_type = 2 (Colella & Sekora 08 / McCorquodale & Colella)
// ---------------------------------------------------------------------------------------------
#define MGPU_CS_C 1.25
// limited 4th-order edge between cells f-1 and f; q points at s(f).  ppm.f90:1878-1890
__device__ __forceinline__ double sedge2_of(const double* q, long st) {
  const double sm2 = q[-2 * st], sm1 = q[-st], s0 = q[0], sp1 = q[st];
  double e = (7.0 / 12.0) * (sm1 + s0) - (1.0 / 12.0) * (sm2 + sp1);
  if ((e - sm1) * (s0 - e) < 0.0) {
    double D2 = 3.0 * (sm1 - 2.0 * e + s0);
    double D2L = sm2 - 2.0 * sm1 + s0;
    double D2R = sm1 - 2.0 * s0 + sp1;
    double sgn = sign1(D2);
    double D2LIM = sgn * dmax2(dmin2(dmin2(MGPU_CS_C * sgn * D2L, MGPU_CS_C * sgn * D2R), sgn * D2), 0.0);
    e = 0.5 * (sm1 + s0) - (1.0 / 6.0) * D2LIM;
  }
  return e;
}
// edge value at face f as the limiter of an "effective-stencil" cell sees it (after the wall
// modifications of ppm.f90:1987-2010 / :2103-2127); q points at s(f)
__device__ __forceinline__ double sedge2_eff(const double* q, long st, int f, const LineBC& b) {
  if (b.wlo) {
    if (f == b.lo) return q[-st];
    if (f == b.lo + 1) return sedge_wall(q - st, st, +1);
  }
  if (b.whi) {
    if (f == b.hi) return sedge_wall(q, st, -1);
    if (f == b.hi + 1 && b.hi_reset) return q[0];
  }
  return sedge2_of(q, st);
}
// Colella 2008 limiter, ppm.f90:1905-1974. e_m1..e_p2 = sedge(c-1..c+2) (outer two only read when needed)
template <class EdgeFn>
__device__ __forceinline__ void cs_limit(const double* q, long st, EdgeFn E, double& sm, double& sp) {
  const double sc = q[0];
  const double e0 = E(0), e1 = E(1);
  double alphap = e1 - sc;
  double alpham = e0 - sc;
  const bool bigp = fabs(alphap) > 2.0 * fabs(alpham);
  const bool bigm = fabs(alpham) > 2.0 * fabs(alphap);
  bool extremum = false;
  if (alpham * alphap >= 0.0) {
    extremum = true;
  } else if (bigp || bigm) {
    double dafacem = e0 - E(-1);
    double dafacep = E(2) - e1;
    double dabarm = sc - q[-st];
    double dabarp = q[st] - sc;
    double dafacemin = dmin2(fabs(dafacem), fabs(dafacep));
    double dabarmin = dmin2(fabs(dabarm), fabs(dabarp));
    double dachkm, dachkp;
    if (dafacemin >= dabarmin) {
      dachkm = dafacem;
      dachkp = dafacep;
    } else {
      dachkm = dabarm;
      dachkp = dabarp;
    }
    extremum = (dachkm * dachkp <= 0.0);
  }
  if (extremum) {
    double D2 = 6.0 * (alpham + alphap);
    double D2L = q[-2 * st] - 2.0 * q[-st] + sc;
    double D2R = sc - 2.0 * q[st] + q[2 * st];
    double D2C = q[-st] - 2.0 * sc + q[st];
    double sgn = sign1(D2);
    double D2LIM =
        dmax2(dmin2(dmin2(dmin2(sgn * D2, MGPU_CS_C * sgn * D2L), MGPU_CS_C * sgn * D2R), MGPU_CS_C * sgn * D2C), 0.0);
    double D2ABS = dmax2(fabs(D2), 1.e-10);
    alpham = alpham * D2LIM / D2ABS;
    alphap = alphap * D2LIM / D2ABS;
  } else {
    if (bigp) {
      double sgn = sign1(alpham);
      double amax = -(alphap * alphap) / (4 * (alpham + alphap));
      double delam = q[-st] - sc;
      if (sgn * amax >= sgn * delam) {
        if (sgn * (delam - alpham) >= 1.e-10) {
          alphap = (-2.0 * delam - 2.0 * sgn * sqrt(delam * delam - delam * alpham));
        } else {
          alphap = -2.0 * alpham;
        }
      }
    }
    if (bigm) {
      double sgn = sign1(alphap);
      double amax = -(alpham * alpham) / (4 * (alpham + alphap));
      double delap = q[st] - sc;
      if (sgn * amax >= sgn * delap) {
        if (sgn * (delap - alphap) >= 1.e-10) {
          alpham = (-2.0 * delap - 2.0 * sgn * sqrt(delap * delap - delap * alphap));
        } else {
          alpham = -2.0 * alphap;
        }
      }
    }
  }
  sm = sc + alpham;
  sp = sc + alphap;
}
__device__ __forceinline__ void ppm2_cell(const double* q, long st, int c, const LineBC& b, double& sm, double& sp) {
  if (b.wlo && c == b.lo) {  // ppm.f90:1987, 2010
    sm = q[-st];
    sp = sedge_wall(q, st, +1);
    return;
  }
  if (b.whi && c == b.hi) {  // :2103, 2127
    sp = q[st];
    sm = sedge_wall(q, st, -1);
    return;
  }
  const bool eff = (b.wlo && c >= b.lo + 1 && c <= b.relimit_last) || (b.whi && c >= b.hi - 2 && c <= b.hi - 1);
  if (eff) {
    cs_limit(q, st, [&](int o) { return sedge2_eff(q + o * st, st, c + o, b); }, sm, sp);
  } else {
    cs_limit(q, st, [&](int o) { return sedge2_of(q + o * st, st); }, sm, sp);
  }
}

// ---------------------------------------------------------------------------------------------
// characteristic tracing, ppm.f90:2233-2251.  up = velocity used for Ip, um = velocity used for Im
// ---------------------------------------------------------------------------------------------
// FAST: dt holds dt/h and h is ignored (no fp64 division; results differ from the reference expression
// (|u|*dt)/h in the last bit, far inside the 1e-12 parity tolerance)
template <bool FAST = false>
__device__ __forceinline__ void ppm_trace(double sc, double sm, double sp, double up, double um, double dt,
                                          double h, double rel_eps, double& Ip, double& Im) {
  const double s6 = 6.0 * sc - 3.0 * (sm + sp);
  {
    double sigma = FAST ? fabs(up) * dt : fabs(up) * dt / h;
    Ip = (up > rel_eps) ? sp - (sigma / 2.0) * (sp - sm - (1.0 - (2.0 / 3.0) * sigma) * s6) : sc;
  }
  {
    double sigma = FAST ? fabs(um) * dt : fabs(um) * dt / h;
    Im = (um < -rel_eps) ? sm + (sigma / 2.0) * (sp - sm + (1.0 - (2.0 / 3.0) * sigma) * s6) : sc;
  }
}

// Godunov upwind select used everywhere (make_edge_scal.f90:881-883)
__device__ __forceinline__ double riemann(double l, double r, double u, double rel_eps) {
  double v = (u > 0.0) ? l : r;
  double savg = 0.5 * (l + r);
  return (fabs(u) > rel_eps) ? v : savg;
}

// one cell, one direction: the two 1-D extrapolated states this cell sends to its hi face (Ip) and lo
// face (Im), for every ppm_type.  uhi/ulo are the face velocities (is_umac) or twice the cell velocity.
template <bool FAST = false>
__device__ __forceinline__ void cell_states(int ppm_type, int slope_order, const double* q, long st, int c,
                                            const LineBC& b, double uhi, double ulo, double dt, double h,
                                            double rel_eps, double& Ip, double& Im) {
  if (ppm_type == 0) {  // make_edge_scal.f90:818-819 written per cell
    const double sl = slope_cell(q, st, c, b, slope_order);
    const double dt2 = 0.5 * dt;
    Ip = q[0] + (0.5 - (FAST ? dt2 * uhi : dt2 * uhi / h)) * sl;
    Im = q[0] - (0.5 + (FAST ? dt2 * ulo : dt2 * ulo / h)) * sl;
  } else {
    double sm, sp;
    if (ppm_type == 1)
      ppm1_cell(q, st, c, b, sm, sp);
    else
      ppm2_cell(q, st, c, b, sm, sp);
    ppm_trace<FAST>(q[0], sm, sp, uhi, ulo, dt, h, rel_eps, Ip, Im);
  }
}

}  // namespace mgpu
