// Device helpers shared by the upwind-first fused edge kernels (mgpu_fused2.cu, mgpu_fused3.cu): FAST arithmetic
// (FMA contraction, dt/dx folded; <= 1e-12 relative to the reference, not bit-identical).
#pragma once
#include "mgpu_recon.cuh"

#ifndef MGPU_EDGE_CLIP
#define MGPU_EDGE_CLIP 0
#endif

namespace mgpu {
namespace {

// one traced state on a face from the parabola (a0=sm, a1=sp) / slope (a0) of its upwind cell.
// a = u dt/h (signed CFL number), up = (u > 0): the upwind cell is the face's low neighbour.
//   up : Ip = sp - s/2 (sp - sm - (1 - 2/3 s) s6),  s = |a|      (ppm.f90:2236-2241)
//  !up : Im = sm + s/2 (sp - sm + (1 - 2/3 s) s6)                (ppm.f90:2244-2249)
// both are  base - a/2 (d + m s6),  m = 2/3 a -+ 1
template <int PPM>
__device__ __forceinline__ double trace1(double a0, double a1, double sc, double a, bool up) {
  if constexpr (PPM == 0) {  // make_edge_scal.f90:818-819
    return fma((up ? 0.5 : -0.5) - 0.5 * a, a0, sc);
  } else {
    const double s6 = 6.0 * sc - 3.0 * (a0 + a1);
    const double d = a1 - a0;
    const double m = fma(2.0 / 3.0, a, up ? -1.0 : 1.0);
    return fma(-0.5 * a, fma(m, s6, d), up ? a1 : a0);
  }
}
// face with |u| <= rel_eps: the reference's 0.5*(l+r) with l = Ip(low cell), r = Im(high cell)
template <int PPM>
__device__ __forceinline__ double trace_slow(double a0l, double scl, double a0r, double scr, double a) {
  if constexpr (PPM == 0) {
    return 0.5 * (fma(0.5 - 0.5 * a, a0l, scl) + fma(-0.5 - 0.5 * a, a0r, scr));
  } else {
    return 0.5 * (scl + scr);  // Ip = Im = s when the velocity does not exceed rel_eps (ppm.f90:2240,2248)
  }
}

// FAST forms of the ppm_type=1 building blocks (ppm.f90:1697-1752): same values as dsvl_of / sedge1_of up to the
// last bit, written without data-dependent branches.
//   van Leer slope: sign(dsc) min(|dsc|, 2|dl|, 2|dr|) if dl*dr > 0 else 0
__device__ __forceinline__ double dsvl_fast(double sm, double s0, double sp) {
  const double dl = s0 - sm, dr = sp - s0;
  const double dsc = 0.5 * (sp - sm);
  const double mn = dmin2(fabs(dl), fabs(dr));
  double lim = dmin2(fabs(dsc), mn + mn);
  lim = (dl * dr > 0.0) ? lim : 0.0;
  return copysign(lim, dsc);
}
//   edge value between cells (sl | sr) with van Leer slopes (dl | dr), clipped to the two cell values
__device__ __forceinline__ double edge_fast(double sl, double sr, double dl, double dr) {
  const double e = fma(-1.0 / 6.0, dr - dl, 0.5 * (sr + sl));
#if MGPU_EDGE_CLIP
  double lo, hi;
  dminmax(sl, sr, lo, hi);
  return dmin2(dmax2(e, lo), hi);
#else
  // The clip of ppm.f90:1723-1727 cannot act on van Leer slopes: dl and dr are zero or have the sign of sr - sl and
  // |dl|, |dr| <= 2 |sr - sl|, so |dr - dl| / 6 <= |sr - sl| / 3 < |sr - sl| / 2 and e already lies between sl and sr
  // (up to the last bit when |sr - sl| is a few ulps).  FAST build only; the exact build keeps the clip.
  return e;
#endif
}

// FAST form of the Colella 2008 limiter (cs_limit, ppm.f90:1905-1974): same branches and values, except that
//   * the two one-sided blocks (|alphap| > 2 |alpham| / |alpham| > 2 |alphap|) are mutually exclusive, so one code path
//     with the roles swapped serves both (one division and one square root instead of two of each when a warp diverges);
//   * at an extremum both ends are scaled by ONE quotient D2LIM / D2ABS (the reference divides twice: last-bit difference).
template <class EdgeFn>
__device__ __forceinline__ void cs_limit_fast(const double* q, long st, EdgeFn E, double& sm, double& sp) {
  const double sc = q[0], qm = q[-st], qp = q[st];
  const double e0 = E(0), e1 = E(1);
  double alphap = e1 - sc;
  double alpham = e0 - sc;
  const bool bigp = fabs(alphap) > 2.0 * fabs(alpham);
  const bool bigm = fabs(alpham) > 2.0 * fabs(alphap);
  bool extremum = false;
  if (alpham * alphap >= 0.0) {
    extremum = true;
  } else if (bigp || bigm) {
    const double dafacem = e0 - E(-1);
    const double dafacep = E(2) - e1;
    const double dabarm = sc - qm;
    const double dabarp = qp - sc;
    const bool face = dmin2(fabs(dafacem), fabs(dafacep)) >= dmin2(fabs(dabarm), fabs(dabarp));
    const double dachkm = face ? dafacem : dabarm, dachkp = face ? dafacep : dabarp;
    extremum = (dachkm * dachkp <= 0.0);
  }
  if (extremum) {
    const double D2 = 6.0 * (alpham + alphap);
    const double D2L = q[-2 * st] - 2.0 * qm + sc;
    const double D2R = sc - 2.0 * qp + q[2 * st];
    const double D2C = qm - 2.0 * sc + qp;
    const double sgn = sign1(D2);
    const double D2LIM =
        dmax2(dmin2(dmin2(dmin2(sgn * D2, MGPU_CS_C * sgn * D2L), MGPU_CS_C * sgn * D2R), MGPU_CS_C * sgn * D2C), 0.0);
    const double r = D2LIM * __drcp_rn(dmax2(fabs(D2), 1.e-10));  // FAST build: reciprocal + multiply (<= 2 ulp)
    alpham = alpham * r;
    alphap = alphap * r;
  } else if (bigp || bigm) {
    const double a_s = bigp ? alpham : alphap;  // the small end, kept
    const double a_b = bigp ? alphap : alpham;  // the big end, limited
    const double del = (bigp ? qm : qp) - sc;
    const double sgn = sign1(a_s);
    const double amax = -(a_b * a_b) * __drcp_rn(4 * (alpham + alphap));
    double nb = a_b;
    if (sgn * amax >= sgn * del) {
      if (sgn * (del - a_s) >= 1.e-10) nb = (-2.0 * del - 2.0 * sgn * sqrt(del * del - del * a_s));
      else nb = -2.0 * a_s;
    }
    alphap = bigp ? nb : alphap;
    alpham = bigp ? alpham : nb;
  }
  sm = sc + alpham;
  sp = sc + alphap;
}

// limited parabola (PPM>=1: a0 = sm, a1 = sp) or slope (PPM==0: a0) of one cell along a line in memory
template <int PPM>
__device__ __forceinline__ void cell_par(const double* q, int st, int slope_order, const LineBC& nb, double& a0,
                                         double& a1) {
  if constexpr (PPM == 0) {
    a0 = slope_cell(q, st, 0, nb, slope_order);
    a1 = 0.0;
  } else if constexpr (PPM == 1) {
    const double m2 = q[-2 * st], m1 = q[-st], c0 = q[0], p1 = q[st], p2 = q[2 * st];
    const double dm = dsvl_fast(m2, m1, c0), d0 = dsvl_fast(m1, c0, p1), dp = dsvl_fast(c0, p1, p2);
    a0 = edge_fast(m1, c0, dm, d0);
    a1 = edge_fast(c0, p1, d0, dp);
    cw_limit(c0, a0, a1);
  } else {
    cs_limit_fast(q, st, [&](int o) { return sedge2_of(q + o * st, st); }, a0, a1);
  }
}



// Boundary-face rule of one face in upwind-first form.  At a face on a physical boundary the reference overwrites
// the left/right states so that both are equal (bc_states / final_bc of mgpu_fused.cu restate it literally); the
// Riemann problem then returns that common value whatever the velocity.  Encoded per face:
//   FB_NONE  interior face (normal upwinding)
//   FB_LEFT / FB_RIGHT  the state of the low / high cell is used whatever the sign of u (FOEXTRAP, HOEXTRAP,
//            REFLECT_EVEN), optionally clamped to inflow-free values (normal velocity component: min/max with 0)
//   FB_GHOST the value is the s of the ghost cell, no transverse / final correction (EXT_DIR)
//   FB_ZERO  the value is 0 (REFLECT_ODD)
enum { FB_NONE = 0, FB_LEFT = 1, FB_RIGHT = 2, FB_GHOST = 3, FB_ZERO = 4 };
struct FaceRule {
  int kind;   // FB_*
  int clamp;  // 0 none, 1 min(.,0) (low boundary), 2 max(.,0) (high boundary)
  bool low;   // the face is the low boundary of its direction
};
__device__ __forceinline__ FaceRule face_rule(int f, int lo, int hi, int bclo, int bchi, bool velnorm) {
  FaceRule r;
  r.kind = FB_NONE;
  r.clamp = 0;
  r.low = false;
  int bc = MGPU_BC_INTERIOR;
  bool low = false;
  if (f == lo && bclo != MGPU_BC_INTERIOR) { bc = bclo; low = true; r.low = true; }
  else if (f == hi + 1 && bchi != MGPU_BC_INTERIOR) { bc = bchi; }
  if (bc == MGPU_BC_EXT_DIR) r.kind = FB_GHOST;
  else if (bc == MGPU_BC_REFLECT_ODD) r.kind = FB_ZERO;
  else if (bc == MGPU_BC_FOEXTRAP || bc == MGPU_BC_HOEXTRAP || bc == MGPU_BC_REFLECT_EVEN) {
    r.kind = low ? FB_RIGHT : FB_LEFT;
    if (velnorm && bc != MGPU_BC_REFLECT_EVEN) r.clamp = low ? 1 : 2;
  }
  return r;
}
__device__ __forceinline__ double clamp_rule(double v, int clamp) {
  if (clamp == 1) return dmin2(v, 0.0);
  if (clamp == 2) return dmax2(v, 0.0);
  return v;
}
// stage-0 state of a forced face: Ip of the low cell / Im of the high cell exactly as ppm_trace / the slope formula
// give them (for PPM the parabola is only traced when the velocity exceeds rel_eps in that direction)
template <int PPM>
__device__ __forceinline__ double forced_state(bool left, double a0, double a1, double sc, double u, double td,
                                               double rel_eps) {
  if constexpr (PPM == 0) return trace1<0>(a0, a1, sc, u * td, left);
  const bool moving = left ? (u > rel_eps) : (u < -rel_eps);
  return moving ? trace1<PPM>(a0, a1, sc, u * td, left) : sc;
}

// limited parabola / slope of one cell with the reference's wall stencils (cell index c on a line with BCs b)
template <int PPM>
__device__ __forceinline__ void cell_par_bc(const double* q, int st, int c, int slope_order, const LineBC& b,
                                            double& a0, double& a1) {
  if constexpr (PPM == 0) {
    a0 = slope_cell(q, st, c, b, slope_order);
    a1 = 0.0;
  } else if constexpr (PPM == 1) {
    ppm1_cell(q, st, c, b, a0, a1);
  } else {  // ppm2_cell (mgpu_recon.cuh) with the FAST limiter
    if (b.wlo && c == b.lo) {  // ppm.f90:1987, 2010
      a0 = q[-st];
      a1 = sedge_wall(q, st, +1);
    } else if (b.whi && c == b.hi) {  // :2103, 2127
      a1 = q[st];
      a0 = sedge_wall(q, st, -1);
    } else {
      const bool eff = (b.wlo && c >= b.lo + 1 && c <= b.relimit_last) || (b.whi && c >= b.hi - 2 && c <= b.hi - 1);
      if (eff) cs_limit_fast(q, st, [&](int o) { return sedge2_eff(q + o * st, st, c + o, b); }, a0, a1);
      else cs_limit_fast(q, st, [&](int o) { return sedge2_of(q + o * st, st); }, a0, a1);
    }
  }
}

__device__ __forceinline__ LineBC no_wall2() {
  LineBC b;
  b.lo = -(1 << 30);
  b.hi = (1 << 30);
  b.wlo = false;
  b.whi = false;
  b.relimit_last = b.lo + 2;
  b.hi_reset = true;
  return b;
}


}  // namespace
}  // namespace mgpu
