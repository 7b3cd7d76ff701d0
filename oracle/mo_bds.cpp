// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Restatement of Source/bds.f90 (bds_type = 1, Bell-Dawson-Shubin):
//   bdsslope_2d :132, bdsslope_3d :285  -- limited bilinear / trilinear slopes (literal transcription)
//   bdsconc_2d  :584, bdsconc_3d  :923  -- edge states by geometric tracing
// bdsconc's x/y (and z) face blocks are index permutations of one another; one
// direction-generic routine restates them, every vector (del, p1..p4) kept in (x,y,z) order so eval_3d
// (:4215) sees the reference's operands.  QUIRK: in the z-face block of bdsconc_3d the corner tetrahedra
// sample the transverse velocity for p3 at (face index + corner offset) instead of (upwind cell), e.g.
// umac(i+1,j+joff,k) bds.f90:2842 where the x/y blocks use vmac(i+ioff,j+1,k) :1039 -- kept literally.
// The reference applies no boundary conditions in bds (is_vel, the_bc_level, start_bccomp are unused).
#include "mo_kernels.h"

namespace mo {

namespace {

inline double min4(double a, double b, double c, double d) { return dmin(dmin(dmin(a, b), c), d); }
inline double max4(double a, double b, double c, double d) { return dmax(dmax(dmax(a, b), c), d); }

// the 3-pass redistribution shared by bdsslope_2d :221-262 and bdsslope_3d :491-532
template <int N>
inline void redistribute(double* sc, const double* smin, const double* smax, double s0) {
  const double eps = 1.e-10;
  double diff[N];
  for (int ll = 0; ll < 3; ++ll) {
    double sum = sc[0];
    for (int m = 1; m < N; ++m) sum = sum + sc[m];
    double sumloc = (N == 4 ? 0.25 : 0.125) * sum;
    double sumdif = (sumloc - s0) * (double)N;
    double sgndif = sign1(sumdif);
    for (int m = 0; m < N; ++m) diff[m] = (sc[m] - s0) * sgndif;
    double kdp = 0.0;
    for (int m = 0; m < N; ++m)
      if (diff[m] > eps) kdp = kdp + 1.0;
    for (int m = 0; m < N; ++m) {
      double div = (kdp < 1.0) ? 1.0 : kdp;
      double redfac;
      if (diff[m] > eps) {
        redfac = sumdif * sgndif / div;
        kdp = kdp - 1.0;
      } else {
        redfac = 0.0;
      }
      double redmax = (sgndif > 0.0) ? sc[m] - smin[m] : smax[m] - sc[m];
      redfac = dmin(redfac, redmax);
      sumdif = sumdif - redfac * sgndif;
      sc[m] = sc[m] - redfac * sgndif;
    }
  }
}

void bdsslope_2d(const int* lo, const int* hi, const Arr& s, Arr& slope, const double* dx) {
  const double hx = dx[0], hy = dx[1];
  Arr sint(lo[0] - 1, hi[0] + 2, lo[1] - 1, hi[1] + 2, 0, 0);
  for (int j = lo[1] - 1; j <= hi[1] + 2; ++j)
    for (int i = lo[0] - 1; i <= hi[0] + 2; ++i) {
      auto S = [&](int a, int b) { return s(i + a, j + b, 0); };
      // 2-D sum order of bds.f90:163-167: sint needs (s(i-2,j-2) + s(i-2,j+1) + s(i+1,j-2) + s(i+1,j+1) - 7(...) + 49(...))/144
      sint(i, j, 0) = (S(-2, -2) + S(-2, 1) + S(1, -2) + S(1, 1) -
                       7.0 * (S(-2, -1) + S(-2, 0) + S(-1, -2) + S(0, -2) + S(-1, 1) + S(0, 1) + S(1, -1) + S(1, 0)) +
                       49.0 * (S(-1, -1) + S(0, -1) + S(-1, 0) + S(0, 0))) /
                      144.0;
    }
  for (int j = lo[1] - 1; j <= hi[1] + 1; ++j)
    for (int i = lo[0] - 1; i <= hi[0] + 1; ++i) {
      const double s0 = s(i, j, 0);
      double sx = 0.5 * (sint(i + 1, j + 1, 0) + sint(i + 1, j, 0) - sint(i, j + 1, 0) - sint(i, j, 0)) / hx;
      double sy = 0.5 * (sint(i + 1, j + 1, 0) - sint(i + 1, j, 0) + sint(i, j + 1, 0) - sint(i, j, 0)) / hy;
      double sxy = (sint(i + 1, j + 1, 0) - sint(i + 1, j, 0) - sint(i, j + 1, 0) + sint(i, j, 0)) / (hx * hy);
      double sc[4], smin[4], smax[4];
      sc[3] = s0 + 0.5 * (hx * sx + hy * sy) + 0.25 * hx * hy * sxy;
      sc[2] = s0 + 0.5 * (hx * sx - hy * sy) - 0.25 * hx * hy * sxy;
      sc[1] = s0 - 0.5 * (hx * sx - hy * sy) - 0.25 * hx * hy * sxy;
      sc[0] = s0 - 0.5 * (hx * sx + hy * sy) + 0.25 * hx * hy * sxy;
      smin[3] = min4(s0, s(i + 1, j, 0), s(i, j + 1, 0), s(i + 1, j + 1, 0));
      smax[3] = max4(s0, s(i + 1, j, 0), s(i, j + 1, 0), s(i + 1, j + 1, 0));
      smin[2] = min4(s0, s(i + 1, j, 0), s(i, j - 1, 0), s(i + 1, j - 1, 0));
      smax[2] = max4(s0, s(i + 1, j, 0), s(i, j - 1, 0), s(i + 1, j - 1, 0));
      smin[1] = min4(s0, s(i - 1, j, 0), s(i, j + 1, 0), s(i - 1, j + 1, 0));
      smax[1] = max4(s0, s(i - 1, j, 0), s(i, j + 1, 0), s(i - 1, j + 1, 0));
      smin[0] = min4(s0, s(i - 1, j, 0), s(i, j - 1, 0), s(i - 1, j - 1, 0));
      smax[0] = max4(s0, s(i - 1, j, 0), s(i, j - 1, 0), s(i - 1, j - 1, 0));
      for (int m = 0; m < 4; ++m) sc[m] = dmax(dmin(sc[m], smax[m]), smin[m]);
      // bds.f90:222: sumloc = 0.25*(sc(4)+sc(3)+sc(2)+sc(1)) -- the 2-D sum runs from sc(4) down
      {
        const double eps = 1.e-10;
        double diff[4];
        for (int ll = 0; ll < 3; ++ll) {
          double sumloc = 0.25 * (sc[3] + sc[2] + sc[1] + sc[0]);
          double sumdif = (sumloc - s0) * 4.0;
          double sgndif = sign1(sumdif);
          for (int m = 0; m < 4; ++m) diff[m] = (sc[m] - s0) * sgndif;
          double kdp = 0.0;
          for (int m = 0; m < 4; ++m)
            if (diff[m] > eps) kdp = kdp + 1.0;
          for (int m = 0; m < 4; ++m) {
            double div = (kdp < 1.0) ? 1.0 : kdp;
            double redfac;
            if (diff[m] > eps) {
              redfac = sumdif * sgndif / div;
              kdp = kdp - 1.0;
            } else {
              redfac = 0.0;
            }
            double redmax = (sgndif > 0.0) ? sc[m] - smin[m] : smax[m] - sc[m];
            redfac = dmin(redfac, redmax);
            sumdif = sumdif - redfac * sgndif;
            sc[m] = sc[m] - redfac * sgndif;
          }
        }
      }
      slope(i, j, 0, 0) = 0.5 * (sc[3] + sc[2] - sc[0] - sc[1]) / hx;
      slope(i, j, 0, 1) = 0.5 * (sc[3] + sc[1] - sc[0] - sc[2]) / hy;
      slope(i, j, 0, 2) = (sc[0] + sc[3] - sc[1] - sc[2]) / (hx * hy);
    }
}

void bdsslope_3d(const int* lo, const int* hi, const Arr& s, Arr& slope, const double* dx) {
  const double hx = dx[0], hy = dx[1], hz = dx[2];
  const double c1 = (343.0 / 1728.0), c2 = (49.0 / 1728.0), c3 = (7.0 / 1728.0), c4 = (1.0 / 1728.0);
  Arr sint(lo[0] - 1, hi[0] + 2, lo[1] - 1, hi[1] + 2, lo[2] - 1, hi[2] + 2);
#pragma omp parallel for
  for (int k = lo[2] - 1; k <= hi[2] + 2; ++k)
    for (int j = lo[1] - 1; j <= hi[1] + 2; ++j)
      for (int i = lo[0] - 1; i <= hi[0] + 2; ++i) {
        auto S = [&](int a, int b, int c) { return s(i + a, j + b, k + c); };
        sint(i, j, k) =
            c1 * (S(0, 0, 0) + S(-1, 0, 0) + S(0, -1, 0) + S(0, 0, -1) + S(-1, -1, 0) + S(-1, 0, -1) + S(0, -1, -1) +
                  S(-1, -1, -1)) -
            c2 * (S(-1, 0, 1) + S(0, 0, 1) + S(-1, -1, 1) + S(0, -1, 1) + S(-1, 1, 0) + S(0, 1, 0) + S(-2, 0, 0) +
                  S(1, 0, 0) + S(-2, -1, 0) + S(1, -1, 0) + S(-1, -2, 0) + S(0, -2, 0) + S(-1, 1, -1) + S(0, 1, -1) +
                  S(-2, 0, -1) + S(1, 0, -1) + S(-2, -1, -1) + S(1, -1, -1) + S(-1, -2, -1) + S(0, -2, -1) +
                  S(-1, 0, -2) + S(0, 0, -2) + S(-1, -1, -2) + S(0, -1, -2)) +
            c3 * (S(-1, 1, 1) + S(0, 1, 1) + S(-2, 0, 1) + S(1, 0, 1) + S(-2, -1, 1) + S(1, -1, 1) + S(-1, -2, 1) +
                  S(0, -2, 1) + S(-2, 1, 0) + S(1, 1, 0) + S(-2, -2, 0) + S(1, -2, 0) + S(-2, 1, -1) + S(1, 1, -1) +
                  S(-2, -2, -1) + S(1, -2, -1) + S(-1, 1, -2) + S(0, 1, -2) + S(-2, 0, -2) + S(1, 0, -2) +
                  S(-2, -1, -2) + S(1, -1, -2) + S(-1, -2, -2) + S(0, -2, -2)) -
            c4 * (S(-2, 1, 1) + S(1, 1, 1) + S(-2, -2, 1) + S(1, -2, 1) + S(-2, 1, -2) + S(1, 1, -2) + S(-2, -2, -2) +
                  S(1, -2, -2));
      }
#pragma omp parallel for
  for (int k = lo[2] - 1; k <= hi[2] + 1; ++k)
    for (int j = lo[1] - 1; j <= hi[1] + 1; ++j)
      for (int i = lo[0] - 1; i <= hi[0] + 1; ++i) {
        auto N = [&](int a, int b, int c) { return sint(i + a, j + b, k + c); };
        auto S = [&](int a, int b, int c) { return s(i + a, j + b, k + c); };
        const double s0 = s(i, j, k);
        double sl[7];
        sl[0] = 0.25 * ((N(1, 0, 0) + N(1, 1, 0) + N(1, 0, 1) + N(1, 1, 1)) - (N(0, 0, 0) + N(0, 1, 0) + N(0, 0, 1) + N(0, 1, 1))) / hx;
        sl[1] = 0.25 * ((N(0, 1, 0) + N(1, 1, 0) + N(0, 1, 1) + N(1, 1, 1)) - (N(0, 0, 0) + N(1, 0, 0) + N(0, 0, 1) + N(1, 0, 1))) / hy;
        sl[2] = 0.25 * ((N(0, 0, 1) + N(1, 0, 1) + N(0, 1, 1) + N(1, 1, 1)) - (N(0, 0, 0) + N(1, 0, 0) + N(0, 1, 0) + N(1, 1, 0))) / hz;
        sl[3] = 0.5 * ((N(0, 0, 0) + N(0, 0, 1) + N(1, 1, 0) + N(1, 1, 1)) - (N(1, 0, 0) + N(1, 0, 1) + N(0, 1, 0) + N(0, 1, 1))) / (hx * hy);
        sl[4] = 0.5 * ((N(0, 0, 0) + N(0, 1, 0) + N(1, 0, 1) + N(1, 1, 1)) - (N(1, 0, 0) + N(1, 1, 0) + N(0, 0, 1) + N(0, 1, 1))) / (hx * hz);
        sl[5] = 0.5 * ((N(0, 0, 0) + N(1, 0, 0) + N(0, 1, 1) + N(1, 1, 1)) - (N(0, 0, 1) + N(1, 0, 1) + N(0, 1, 0) + N(1, 1, 0))) / (hy * hz);
        sl[6] = (-N(0, 0, 0) + N(1, 0, 0) + N(0, 1, 0) + N(0, 0, 1) - N(1, 1, 0) - N(1, 0, 1) - N(0, 1, 1) + N(1, 1, 1)) / (hx * hy * hz);
        double sc[8], smin[8], smax[8];
        // corner values: sc(8) = +++, sc(7) = ++-, sc(6) = +-+, sc(5) = +--, sc(4) = -++, sc(3) = -+-, sc(2) = --+, sc(1) = ---
        sc[7] = s0 + 0.5 * (hx * sl[0] + hy * sl[1] + hz * sl[2]) + 0.25 * (hx * hy * sl[3] + hx * hz * sl[4] + hy * hz * sl[5]) + 0.125 * hx * hy * hz * sl[6];
        sc[6] = s0 + 0.5 * (hx * sl[0] + hy * sl[1] - hz * sl[2]) + 0.25 * (hx * hy * sl[3] - hx * hz * sl[4] - hy * hz * sl[5]) - 0.125 * hx * hy * hz * sl[6];
        sc[5] = s0 + 0.5 * (hx * sl[0] - hy * sl[1] + hz * sl[2]) + 0.25 * (-hx * hy * sl[3] + hx * hz * sl[4] - hy * hz * sl[5]) - 0.125 * hx * hy * hz * sl[6];
        sc[4] = s0 + 0.5 * (hx * sl[0] - hy * sl[1] - hz * sl[2]) + 0.25 * (-hx * hy * sl[3] - hx * hz * sl[4] + hy * hz * sl[5]) + 0.125 * hx * hy * hz * sl[6];
        sc[3] = s0 + 0.5 * (-hx * sl[0] + hy * sl[1] + hz * sl[2]) + 0.25 * (-hx * hy * sl[3] - hx * hz * sl[4] + hy * hz * sl[5]) - 0.125 * hx * hy * hz * sl[6];
        sc[2] = s0 + 0.5 * (-hx * sl[0] + hy * sl[1] - hz * sl[2]) + 0.25 * (-hx * hy * sl[3] + hx * hz * sl[4] - hy * hz * sl[5]) + 0.125 * hx * hy * hz * sl[6];
        sc[1] = s0 + 0.5 * (-hx * sl[0] - hy * sl[1] + hz * sl[2]) + 0.25 * (hx * hy * sl[3] - hx * hz * sl[4] - hy * hz * sl[5]) + 0.125 * hx * hy * hz * sl[6];
        sc[0] = s0 + 0.5 * (-hx * sl[0] - hy * sl[1] - hz * sl[2]) + 0.25 * (hx * hy * sl[3] + hx * hz * sl[4] + hy * hz * sl[5]) - 0.125 * hx * hy * hz * sl[6];
        // min/max over the 8 cells sharing each corner: corner m = (a,b,c) in {0,1}^3 with cells (i+a-1..i+a, ...)
        for (int m = 0; m < 8; ++m) {
          const int a = (m >> 2) & 1, b = (m >> 1) & 1, c = m & 1;  // m = 4a+2b+c matches sc(m+1)'s sign pattern
          double mn = S(a - 1, b - 1, c - 1), mx = mn;
          for (int cc = 0; cc < 2; ++cc)
            for (int bb = 0; bb < 2; ++bb)
              for (int aa = 0; aa < 2; ++aa) {
                double v = S(a - 1 + aa, b - 1 + bb, c - 1 + cc);
                mn = dmin(mn, v);
                mx = dmax(mx, v);
              }
          smin[m] = mn;
          smax[m] = mx;
        }
        for (int m = 0; m < 8; ++m) sc[m] = dmax(dmin(sc[m], smax[m]), smin[m]);
        redistribute<8>(sc, smin, smax, s0);  // bds.f90:492: sum runs sc(1)..sc(8)
        // final slopes, bds.f90:536-574 (1-based sc(n) = sc[n-1])
        slope(i, j, k, 0) = 0.25 * ((sc[4] + sc[6] + sc[5] + sc[7]) - (sc[0] + sc[2] + sc[1] + sc[3])) / hx;
        slope(i, j, k, 1) = 0.25 * ((sc[2] + sc[6] + sc[3] + sc[7]) - (sc[0] + sc[4] + sc[1] + sc[5])) / hy;
        slope(i, j, k, 2) = 0.25 * ((sc[1] + sc[5] + sc[3] + sc[7]) - (sc[0] + sc[4] + sc[2] + sc[6])) / hz;
        slope(i, j, k, 3) = 0.5 * ((sc[0] + sc[1] + sc[6] + sc[7]) - (sc[4] + sc[5] + sc[2] + sc[3])) / (hx * hy);
        slope(i, j, k, 4) = 0.5 * ((sc[0] + sc[2] + sc[5] + sc[7]) - (sc[4] + sc[6] + sc[1] + sc[3])) / (hx * hz);
        slope(i, j, k, 5) = 0.5 * ((sc[0] + sc[4] + sc[3] + sc[7]) - (sc[1] + sc[5] + sc[2] + sc[6])) / (hy * hz);
        slope(i, j, k, 6) = (-sc[0] + sc[4] + sc[2] + sc[1] - sc[6] - sc[5] - sc[3] + sc[7]) / (hx * hy * hz);
      }
}

inline double eval2(double s, const double* sl, const double* del) {  // eval_2d, bds.f90:4204
  return s + del[0] * sl[0] + del[1] * sl[1] + del[0] * del[1] * sl[2];
}
inline double eval3(double s, const double* sl, const double* del) {  // eval_3d, bds.f90:4215
  return s + del[0] * sl[0] + del[1] * sl[1] + del[2] * sl[2] + del[0] * del[1] * sl[3] + del[0] * del[2] * sl[4] +
         del[1] * del[2] * sl[5] + del[0] * del[1] * del[2] * sl[6];
}

}  // namespace

void bds_box(const mgpu_params& P, const Arr& sfull, Arr* sedge_full, const Arr* umac, const Arr& force_full,
             const int* lo, const int* hi, int comp, bool is_cons) {
  const int dm = P.dm;
  const double dt = P.dt;
  const double* h = P.dx;
  const Arr s = sfull.comp(comp);
  const Arr force = force_full.comp(comp);
  Box tb = grown(lo, hi, dm, 1), vb = grown(lo, hi, dm, 0);
  const int nsl = (dm == 2) ? 3 : 7;
  Arr slope(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2], nsl);
  if (dm == 2) bdsslope_2d(lo, hi, s, slope, h);
  else bdsslope_3d(lo, hi, s, slope, h);

  // cell-centred velocity derivatives on lo-1:hi+1 (bds.f90:621-628 / :965-975)
  Arr dv(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2], 3), divu(tb.lo[0], tb.hi[0], tb.lo[1], tb.hi[1], tb.lo[2], tb.hi[2]);
  for_box(tb, [&](int i, int j, int k) {
    dv(i, j, k, 0) = (umac[0](i + 1, j, k) - umac[0](i, j, k)) / h[0];
    dv(i, j, k, 1) = (umac[1](i, j + 1, k) - umac[1](i, j, k)) / h[1];
    if (dm == 3) {
      dv(i, j, k, 2) = (umac[2](i, j, k + 1) - umac[2](i, j, k)) / h[2];
      divu(i, j, k) = dv(i, j, k, 0) + dv(i, j, k, 1) + dv(i, j, k, 2);
    } else {
      divu(i, j, k) = dv(i, j, k, 0) + dv(i, j, k, 1);
    }
  });
  const double dt2 = dt / 2.0, dt3 = dt / 3.0, dt4 = dt / 4.0;
  const double half = 0.5, sixth = 1.0 / 6.0;

  auto slp = [&](const int* c, double* out) {
    for (int m = 0; m < nsl; ++m) out[m] = slope(c[0], c[1], c[2], m);
  };
  auto ev = [&](const int* c, const double* del) {
    double sl[7];
    slp(c, sl);
    return (dm == 2) ? eval2(s(c[0], c[1], c[2]), sl, del) : eval3(s(c[0], c[1], c[2]), sl, del);
  };
  auto U = [&](int q, const int* c) { return umac[q](c[0], c[1], c[2]); };

  for (int d = 0; d < dm; ++d) {
    Box fb = vb;
    fb.hi[d] = hi[d] + 1;
    Arr sedge = sedge_full[d].comp(comp);
    for_box(fb, [&](int i, int j, int k) {
      const int f[3] = {i, j, k};
      const double vn = U(d, f);
      double sgn[3] = {0.0, 0.0, 0.0};
      int uc[3] = {i, j, k};  // upwind cell of the face
      if (vn > 0.0) { sgn[d] = 1.0; uc[d] -= 1; } else { sgn[d] = -1.0; }
      double del[3] = {0.0, 0.0, 0.0};
      del[d] = sgn[d] * 0.5 * h[d] - 0.5 * vn * dt;
      double se = ev(uc, del);
      if (is_cons) {
        se = se * (1.0 - dt2 * dv(uc[0], uc[1], uc[2], d)) + dt2 * force(uc[0], uc[1], uc[2]);
      } else {
        double tsum;
        if (dm == 2) tsum = dv(uc[0], uc[1], uc[2], 1 - d);
        else {
          const int ta = (d == 0) ? 1 : 0, tb2 = (d == 2) ? 1 : 2;
          tsum = dv(uc[0], uc[1], uc[2], ta) + dv(uc[0], uc[1], uc[2], tb2);
        }
        se = se * (1.0 + dt2 * tsum) + dt2 * force(uc[0], uc[1], uc[2]);
      }
      for (int t = 0; t < dm; ++t) {
        if (t == d) continue;
        const int r = (dm == 3) ? 3 - d - t : -1;
        for (int side = 1; side >= 0; --side) {  // Gamma^{t+} then Gamma^{t-}
          int tf[3] = {uc[0], uc[1], uc[2]};       // t-face of the upwind cell on this side
          tf[t] += side;
          const double vt = U(t, tf);
          int toff;
          if (vt > 0.0) { sgn[t] = 1.0; toff = side ? 0 : -1; } else { sgn[t] = -1.0; toff = side ? 1 : 0; }
          int fn[3] = {i, j, k};  // the d-face shifted by toff in t
          fn[t] += toff;
          double u = 0.0;
          if (vn * U(d, fn) > 0.0) u = U(d, fn);
          double p1[3] = {0.0, 0.0, 0.0}, p2[3] = {0.0, 0.0, 0.0}, p3[3] = {0.0, 0.0, 0.0}, p4[3];
          p1[d] = sgn[d] * 0.5 * h[d];
          p1[t] = sgn[t] * 0.5 * h[t];
          p2[d] = sgn[d] * 0.5 * h[d] - vn * dt;
          p2[t] = sgn[t] * 0.5 * h[t];
          p3[d] = sgn[d] * 0.5 * h[d] - u * dt;
          p3[t] = sgn[t] * 0.5 * h[t] - vt * dt;
          int c2[3] = {uc[0], uc[1], uc[2]};
          c2[t] += toff;
          for (int l = 0; l < 3; ++l) del[l] = (p2[l] + p3[l]) / 2.0;
          const double val1 = ev(c2, del);
          for (int l = 0; l < 3; ++l) del[l] = (p1[l] + p3[l]) / 2.0;
          const double val2 = ev(c2, del);
          for (int l = 0; l < 3; ++l) del[l] = (p1[l] + p2[l]) / 2.0;
          const double val3 = ev(c2, del);
          double gamma = (val1 + val2 + val3) / 3.0;
          if (dm == 2) {
            if (is_cons) gamma = gamma * (1.0 - dt3 * divu(c2[0], c2[1], c2[2]));
          } else {
            if (is_cons) gamma = gamma * (1.0 - dt3 * (dv(c2[0], c2[1], c2[2], d) + dv(c2[0], c2[1], c2[2], t)));
            else gamma = gamma * (1.0 + dt3 * dv(c2[0], c2[1], c2[2], r));
            for (int rside = 1; rside >= 0; --rside) {  // corner corrections Gamma^{t,r+} then Gamma^{t,r-}
              int rf[3] = {c2[0], c2[1], c2[2]};
              rf[r] += rside;
              const double vr = U(r, rf);
              int roff;
              if (vr > 0.0) { sgn[r] = 1.0; roff = rside ? 0 : -1; } else { sgn[r] = -1.0; roff = rside ? 1 : 0; }
              int fnn[3] = {fn[0], fn[1], fn[2]};
              fnn[r] += roff;
              double uu = 0.0;
              if (vn * U(d, fnn) > 0.0) uu = U(d, fnn);
              int tfr[3] = {tf[0], tf[1], tf[2]};
              tfr[r] += roff;
              double vv = 0.0;
              if (vt * U(t, tfr) > 0.0) vv = U(t, tfr);
              // velocity used for p3 in the t direction
              double vt3 = vt;
              if (d == 2) {  // QUIRK bds.f90:2842,2916,3051,3125,3261,3335,3470,3544
                int q[3] = {i, j, k};
                q[t] += side;
                q[r] += roff;
                vt3 = U(t, q);
              }
              p1[d] = sgn[d] * 0.5 * h[d]; p1[t] = sgn[t] * 0.5 * h[t]; p1[r] = sgn[r] * 0.5 * h[r];
              p2[d] = sgn[d] * 0.5 * h[d] - vn * dt; p2[t] = sgn[t] * 0.5 * h[t]; p2[r] = sgn[r] * 0.5 * h[r];
              p3[d] = sgn[d] * 0.5 * h[d] - vn * dt; p3[t] = sgn[t] * 0.5 * h[t] - vt3 * dt; p3[r] = sgn[r] * 0.5 * h[r];
              p4[d] = sgn[d] * 0.5 * h[d] - uu * dt; p4[t] = sgn[t] * 0.5 * h[t] - vv * dt; p4[r] = sgn[r] * 0.5 * h[r] - vr * dt;
              int c3[3] = {c2[0], c2[1], c2[2]};
              c3[r] += roff;
              for (int l = 0; l < 3; ++l) del[l] = (p1[l] + p2[l] + p3[l] + p4[l]) / 4.0;
              const double w1 = ev(c3, del);
              for (int l = 0; l < 3; ++l) del[l] = half * p1[l] + sixth * (p2[l] + p3[l] + p4[l]);
              const double w2 = ev(c3, del);
              for (int l = 0; l < 3; ++l) del[l] = half * p2[l] + sixth * (p1[l] + p3[l] + p4[l]);
              const double w3 = ev(c3, del);
              for (int l = 0; l < 3; ++l) del[l] = half * p3[l] + sixth * (p2[l] + p1[l] + p4[l]);
              const double w4 = ev(c3, del);
              for (int l = 0; l < 3; ++l) del[l] = half * p4[l] + sixth * (p2[l] + p3[l] + p1[l]);
              const double w5 = ev(c3, del);
              double gamma2 = -0.8 * w1 + 0.45 * (w2 + w3 + w4 + w5);
              if (is_cons) gamma2 = gamma2 * (1.0 - dt4 * divu(c3[0], c3[1], c3[2]));
              gamma2 = gamma2 * vr;
              if (rside) gamma = gamma - dt * gamma2 / (3.0 * h[r]);
              else gamma = gamma + dt * gamma2 / (3.0 * h[r]);
            }
            // the triangle points of the next Gamma have zero r-components again
            p1[r] = p2[r] = p3[r] = 0.0;
          }
          gamma = gamma * vt;
          if (side) se = se - dt * gamma / (2.0 * h[t]);
          else se = se + dt * gamma / (2.0 * h[t]);
        }
      }
      sedge(i, j, k) = se;
    });
  }
}

}  // namespace mo
