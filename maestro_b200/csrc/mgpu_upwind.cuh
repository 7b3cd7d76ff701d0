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
    ppm2_cell(q, st, 0, nb, a0, a1);
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
  } else {
    ppm2_cell(q, st, c, b, a0, a1);
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
