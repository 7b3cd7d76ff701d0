// bds_type = 1: Bell-Dawson-Shubin edge states (Source/bds.f90:16), 2-D and 3-D.
//   k_bds_sint  : bicubic / tricubic interpolation of s to nodes           (bdsslope_2d :158-168, _3d :319-352)
//   k_bds_slope : limited multilinear slopes of every cell lo-1:hi+1        (bdsslope_2d :170-279, _3d :354-576)
//   k_bds_conc  : one launch per face direction; each thread traces one face (bdsconc_2d :584, bdsconc_3d :923)
// Node values and slopes (3 / 7 components) live in the arena, exactly the `slope` multifab of the reference.
// The reference applies no boundary conditions here (bds.f90: is_vel, the_bc_level, start_bccomp are unused).
// bdsconc's face blocks are index permutations of one another, written once over (d, t, r) = (normal, transverse,
// corner) with every vector kept in (x,y,z) order; QUIRK: the z-face block samples the transverse velocity of p3
// in the corner tetrahedra at (face + corner offset), bds.f90:2842,2916,3051,3125,3261,3335,3470,3544.
// Two builds of this file (like mgpu_fused.cu): the exact one (-fmad=false, every division the reference writes: bit-
// identical) and the FAST one (mgpu_bds_fast.cu: FMA contraction, divisions by dx, 3, 3h, 2h as multiplications by
// reciprocals formed once; <= 1e-12 from the reference), the default.  `exact = 1` (mgpu_set_option) selects the first.
#include "mgpu_bds.cuh"
#include "mgpu_recon.cuh"

#ifdef MGPU_BDS_FAST
#define BDS_OVER(x, d, rd) ((x) * (rd))
#define BDS_FN(name) name##_fast
#else
#define BDS_OVER(x, d, rd) ((x) / (d))
#define BDS_FN(name) name##_exact
#endif

namespace mgpu {

namespace {

struct BdsArgs {
  int dm;
  bool is_cons;
  int lo[3], hi[3];
  double dt, dx[3];
  double rdx[3];    // 1/dx (FAST build)
  Box3 tb, vb, nb;  // cells lo-1:hi+1, valid cells, nodes lo-1:hi+2
  DV s, force;      // single-component views
  DV umac[3];
  DV sedge[3];      // single-component views of the output
  DV sint;          // nodal, on nb
  DV slope;         // 3 or 7 components on tb
};

__global__ void k_bds_sint(BdsArgs a) {
  int ix[3];
  if (!decode3(a.nb, ix)) return;
  const double* q = a.s.p + a.s.off(ix[0], ix[1], ix[2]);
  const long sy = a.s.stride(1), sz = a.s.stride(2);
  if (a.dm == 2) {
    auto S = [&](int i, int j) { return q[i + j * sy]; };
    a.sint(ix[0], ix[1], ix[2]) =
        (S(-2, -2) + S(-2, 1) + S(1, -2) + S(1, 1) -
         7.0 * (S(-2, -1) + S(-2, 0) + S(-1, -2) + S(0, -2) + S(-1, 1) + S(0, 1) + S(1, -1) + S(1, 0)) +
         49.0 * (S(-1, -1) + S(0, -1) + S(-1, 0) + S(0, 0))) /
        144.0;
    return;
  }
  const double c1 = (343.0 / 1728.0), c2 = (49.0 / 1728.0), c3 = (7.0 / 1728.0), c4 = (1.0 / 1728.0);
  auto S = [&](int i, int j, int k) { return q[i + j * sy + k * sz]; };
  a.sint(ix[0], ix[1], ix[2]) =
      c1 * (S(0, 0, 0) + S(-1, 0, 0) + S(0, -1, 0) + S(0, 0, -1) + S(-1, -1, 0) + S(-1, 0, -1) + S(0, -1, -1) +
            S(-1, -1, -1)) -
      c2 * (S(-1, 0, 1) + S(0, 0, 1) + S(-1, -1, 1) + S(0, -1, 1) + S(-1, 1, 0) + S(0, 1, 0) + S(-2, 0, 0) + S(1, 0, 0) +
            S(-2, -1, 0) + S(1, -1, 0) + S(-1, -2, 0) + S(0, -2, 0) + S(-1, 1, -1) + S(0, 1, -1) + S(-2, 0, -1) +
            S(1, 0, -1) + S(-2, -1, -1) + S(1, -1, -1) + S(-1, -2, -1) + S(0, -2, -1) + S(-1, 0, -2) + S(0, 0, -2) +
            S(-1, -1, -2) + S(0, -1, -2)) +
      c3 * (S(-1, 1, 1) + S(0, 1, 1) + S(-2, 0, 1) + S(1, 0, 1) + S(-2, -1, 1) + S(1, -1, 1) + S(-1, -2, 1) + S(0, -2, 1) +
            S(-2, 1, 0) + S(1, 1, 0) + S(-2, -2, 0) + S(1, -2, 0) + S(-2, 1, -1) + S(1, 1, -1) + S(-2, -2, -1) +
            S(1, -2, -1) + S(-1, 1, -2) + S(0, 1, -2) + S(-2, 0, -2) + S(1, 0, -2) + S(-2, -1, -2) + S(1, -1, -2) +
            S(-1, -2, -2) + S(0, -2, -2)) -
      c4 * (S(-2, 1, 1) + S(1, 1, 1) + S(-2, -2, 1) + S(1, -2, 1) + S(-2, 1, -2) + S(1, 1, -2) + S(-2, -2, -2) +
            S(1, -2, -2));
}

// three redistribution passes, bdsslope_2d :221-262 / bdsslope_3d :491-532; `sumloc` is formed by the caller's order
template <int N>
__device__ __forceinline__ void bds_pass(double* sc, const double* smin, const double* smax, double s0, double sumloc) {
  const double eps = 1.e-10;
  double sumdif = (sumloc - s0) * (double)N;
  const double sgndif = sign1(sumdif);
#ifdef MGPU_BDS_FAST
  // the divisor is the count of corners still to be reduced (1..N): a table of reciprocals instead of N divisions
  static constexpr double RK[9] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 3.0, 1.0 / 4.0, 1.0 / 5.0, 1.0 / 6.0, 1.0 / 7.0, 1.0 / 8.0};
  int kdp = 0;
#else
  double kdp = 0.0;
#endif
  bool big[N];
#pragma unroll
  for (int m = 0; m < N; ++m) {
    big[m] = (sc[m] - s0) * sgndif > eps;
#ifdef MGPU_BDS_FAST
    if (big[m]) kdp = kdp + 1;
#else
    if (big[m]) kdp = kdp + 1.0;
#endif
  }
#pragma unroll
  for (int m = 0; m < N; ++m) {
    double redfac;
#ifdef MGPU_BDS_FAST
    if (big[m]) {
      redfac = sumdif * sgndif * RK[kdp];
      kdp = kdp - 1;
    } else {
      redfac = 0.0;
    }
#else
    const double div = (kdp < 1.0) ? 1.0 : kdp;
    if (big[m]) {
      redfac = sumdif * sgndif / div;
      kdp = kdp - 1.0;
    } else {
      redfac = 0.0;
    }
#endif
    const double redmax = (sgndif > 0.0) ? sc[m] - smin[m] : smax[m] - sc[m];
    redfac = dmin2(redfac, redmax);
    sumdif = sumdif - redfac * sgndif;
    sc[m] = sc[m] - redfac * sgndif;
  }
}

__global__ void k_bds_slope(BdsArgs a) {
  int ix[3];
  if (!decode3(a.tb, ix)) return;
  const double* q = a.s.p + a.s.off(ix[0], ix[1], ix[2]);
  const long sy = a.s.stride(1), sz = a.s.stride(2);
  const double* nq = a.sint.p + a.sint.off(ix[0], ix[1], ix[2]);
  const long ny = a.sint.stride(1), nz = a.sint.stride(2);
  const double hx = a.dx[0], hy = a.dx[1], hz = a.dx[2];
  const double rhx = a.rdx[0], rhy = a.rdx[1], rhz = a.rdx[2];
  (void)rhx; (void)rhy; (void)rhz;
  const double s0 = q[0];
  const long so = a.slope.off(ix[0], ix[1], ix[2]);
  if (a.dm == 2) {
    auto N = [&](int i, int j) { return nq[i + j * ny]; };
    auto S = [&](int i, int j) { return q[i + j * sy]; };
    const double sx = BDS_OVER(0.5 * (N(1, 1) + N(1, 0) - N(0, 1) - N(0, 0)), hx, rhx);
    const double sy_ = BDS_OVER(0.5 * (N(1, 1) - N(1, 0) + N(0, 1) - N(0, 0)), hy, rhy);
    const double sxy = BDS_OVER((N(1, 1) - N(1, 0) - N(0, 1) + N(0, 0)), (hx * hy), (rhx * rhy));
    double sc[4], smin[4], smax[4];
    sc[3] = s0 + 0.5 * (hx * sx + hy * sy_) + 0.25 * hx * hy * sxy;
    sc[2] = s0 + 0.5 * (hx * sx - hy * sy_) - 0.25 * hx * hy * sxy;
    sc[1] = s0 - 0.5 * (hx * sx - hy * sy_) - 0.25 * hx * hy * sxy;
    sc[0] = s0 - 0.5 * (hx * sx + hy * sy_) + 0.25 * hx * hy * sxy;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int di = (m >> 1) ? 1 : -1, dj = (m & 1) ? 1 : -1;  // sc(4)=++, sc(3)=+-, sc(2)=-+, sc(1)=--
      const double v1 = S(di, 0), v2 = S(0, dj), v3 = S(di, dj);
      smin[m] = dmin2(dmin2(dmin2(s0, v1), v2), v3);
      smax[m] = dmax2(dmax2(dmax2(s0, v1), v2), v3);
      sc[m] = dmax2(dmin2(sc[m], smax[m]), smin[m]);
    }
    for (int ll = 0; ll < 3; ++ll) bds_pass<4>(sc, smin, smax, s0, 0.25 * (sc[3] + sc[2] + sc[1] + sc[0]));
    a.slope.p[so] = BDS_OVER(0.5 * (sc[3] + sc[2] - sc[0] - sc[1]), hx, rhx);
    a.slope.p[so + a.slope.cs] = BDS_OVER(0.5 * (sc[3] + sc[1] - sc[0] - sc[2]), hy, rhy);
    a.slope.p[so + 2 * a.slope.cs] = BDS_OVER((sc[0] + sc[3] - sc[1] - sc[2]), (hx * hy), (rhx * rhy));
    return;
  }
  auto N = [&](int i, int j, int k) { return nq[i + j * ny + k * nz]; };
  auto S = [&](int i, int j, int k) { return q[i + j * sy + k * sz]; };
  double sl[7];
  sl[0] = BDS_OVER(0.25 * ((N(1, 0, 0) + N(1, 1, 0) + N(1, 0, 1) + N(1, 1, 1)) - (N(0, 0, 0) + N(0, 1, 0) + N(0, 0, 1) + N(0, 1, 1))), hx, rhx);
  sl[1] = BDS_OVER(0.25 * ((N(0, 1, 0) + N(1, 1, 0) + N(0, 1, 1) + N(1, 1, 1)) - (N(0, 0, 0) + N(1, 0, 0) + N(0, 0, 1) + N(1, 0, 1))), hy, rhy);
  sl[2] = BDS_OVER(0.25 * ((N(0, 0, 1) + N(1, 0, 1) + N(0, 1, 1) + N(1, 1, 1)) - (N(0, 0, 0) + N(1, 0, 0) + N(0, 1, 0) + N(1, 1, 0))), hz, rhz);
  sl[3] = BDS_OVER(0.5 * ((N(0, 0, 0) + N(0, 0, 1) + N(1, 1, 0) + N(1, 1, 1)) - (N(1, 0, 0) + N(1, 0, 1) + N(0, 1, 0) + N(0, 1, 1))), (hx * hy), (rhx * rhy));
  sl[4] = BDS_OVER(0.5 * ((N(0, 0, 0) + N(0, 1, 0) + N(1, 0, 1) + N(1, 1, 1)) - (N(1, 0, 0) + N(1, 1, 0) + N(0, 0, 1) + N(0, 1, 1))), (hx * hz), (rhx * rhz));
  sl[5] = BDS_OVER(0.5 * ((N(0, 0, 0) + N(1, 0, 0) + N(0, 1, 1) + N(1, 1, 1)) - (N(0, 0, 1) + N(1, 0, 1) + N(0, 1, 0) + N(1, 1, 0))), (hy * hz), (rhy * rhz));
  sl[6] = BDS_OVER((-N(0, 0, 0) + N(1, 0, 0) + N(0, 1, 0) + N(0, 0, 1) - N(1, 1, 0) - N(1, 0, 1) - N(0, 1, 1) + N(1, 1, 1)), (hx * hy * hz), (rhx * rhy * rhz));
  double sc[8], smin[8], smax[8];
  // sc(n), n = 1 + 4a + 2b + c with (a,b,c) = 1 for the + side in (x,y,z): bds.f90:405-452
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double ax = (m & 4) ? 1.0 : -1.0, ay = (m & 2) ? 1.0 : -1.0, az = (m & 1) ? 1.0 : -1.0;
    sc[m] = s0 + 0.5 * (ax * hx * sl[0] + ay * hy * sl[1] + az * hz * sl[2]) +
            0.25 * ((ax * ay) * hx * hy * sl[3] + (ax * az) * hx * hz * sl[4] + (ay * az) * hy * hz * sl[5]) +
            (ax * ay * az) * 0.125 * hx * hy * hz * sl[6];
    const int i0 = (m & 4) ? 0 : -1, j0 = (m & 2) ? 0 : -1, k0 = (m & 1) ? 0 : -1;
    double mn = S(i0, j0, k0), mx = mn;
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      const double v = S(i0 + (t & 1), j0 + ((t >> 1) & 1), k0 + ((t >> 2) & 1));
      mn = dmin2(mn, v);
      mx = dmax2(mx, v);
    }
    smin[m] = mn;
    smax[m] = mx;
    sc[m] = dmax2(dmin2(sc[m], smax[m]), smin[m]);
  }
  for (int ll = 0; ll < 3; ++ll)
    bds_pass<8>(sc, smin, smax, s0, 0.125 * (sc[0] + sc[1] + sc[2] + sc[3] + sc[4] + sc[5] + sc[6] + sc[7]));
  const long cs = a.slope.cs;
  a.slope.p[so] = BDS_OVER(0.25 * ((sc[4] + sc[6] + sc[5] + sc[7]) - (sc[0] + sc[2] + sc[1] + sc[3])), hx, rhx);
  a.slope.p[so + cs] = BDS_OVER(0.25 * ((sc[2] + sc[6] + sc[3] + sc[7]) - (sc[0] + sc[4] + sc[1] + sc[5])), hy, rhy);
  a.slope.p[so + 2 * cs] = BDS_OVER(0.25 * ((sc[1] + sc[5] + sc[3] + sc[7]) - (sc[0] + sc[4] + sc[2] + sc[6])), hz, rhz);
  a.slope.p[so + 3 * cs] = BDS_OVER(0.5 * ((sc[0] + sc[1] + sc[6] + sc[7]) - (sc[4] + sc[5] + sc[2] + sc[3])), (hx * hy), (rhx * rhy));
  a.slope.p[so + 4 * cs] = BDS_OVER(0.5 * ((sc[0] + sc[2] + sc[5] + sc[7]) - (sc[4] + sc[6] + sc[1] + sc[3])), (hx * hz), (rhx * rhz));
  a.slope.p[so + 5 * cs] = BDS_OVER(0.5 * ((sc[0] + sc[4] + sc[3] + sc[7]) - (sc[1] + sc[5] + sc[2] + sc[6])), (hy * hz), (rhy * rhz));
  a.slope.p[so + 6 * cs] = BDS_OVER((-sc[0] + sc[4] + sc[2] + sc[1] - sc[6] - sc[5] - sc[3] + sc[7]), (hx * hy * hz), (rhx * rhy * rhz));
}

// polynomial of cell c evaluated at offset del: eval_2d :4204 / eval_3d :4215
template <int DM>
__device__ __forceinline__ double bds_eval(const BdsArgs& a, const int* c, const double* del) {
  // (component-major slopes: a warp's eight loads touch two cache lines each; interleaving value + slopes per cell as
  // 64 B records was measured 8 % slower -- four 128-bit loads per cell, each spread over 16 lines per warp)
  const double s = a.s.p[a.s.off32(c[0], c[1], c[2])];
  const double* sl = a.slope.p + a.slope.off32(c[0], c[1], c[2]);
  const int cs = (int)a.slope.cs;
#ifdef MGPU_BDS_FAST  // nested form: 3 resp. 7 fused multiply-adds instead of 5 resp. 16 operations
  if (DM == 2) return s + del[0] * (sl[0] + del[1] * sl[2 * cs]) + del[1] * sl[cs];
  return s + del[0] * (sl[0] + del[1] * (sl[3 * cs] + del[2] * sl[6 * cs]) + del[2] * sl[4 * cs]) +
         del[1] * (sl[cs] + del[2] * sl[5 * cs]) + del[2] * sl[2 * cs];
#else
  if (DM == 2) return s + del[0] * sl[0] + del[1] * sl[cs] + del[0] * del[1] * sl[2 * cs];
  return s + del[0] * sl[0] + del[1] * sl[cs] + del[2] * sl[2 * cs] + del[0] * del[1] * sl[3 * cs] +
         del[0] * del[2] * sl[4 * cs] + del[1] * del[2] * sl[5 * cs] + del[0] * del[1] * del[2] * sl[6 * cs];
#endif
}

// d(velocity_q)/dx_q of cell c
__device__ __forceinline__ double dvel(const BdsArgs& a, int q, const int* c) {
  const DV& u = a.umac[q];
  const int o = u.off32(c[0], c[1], c[2]);
  return BDS_OVER((u.p[o + (int)u.stride(q)] - u.p[o]), a.dx[q], a.rdx[q]);
}
__device__ __forceinline__ double divu_of(const BdsArgs& a, int dm, const int* c) {
  double r = dvel(a, 0, c) + dvel(a, 1, c);
  if (dm == 3) r = r + dvel(a, 2, c);
  return r;
}

template <int DM, int D>
__global__ void __launch_bounds__(128, 4) k_bds_conc(BdsArgs a) {
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[D] += 1;
  // (bx, by, bz) thread blocks: the traced cells of a face lie within one cell of it in every direction, so a block that
  // is a few rows thick in y and z re-reads far fewer distinct rows than a single row segment does
  ix[0] = fb.lo[0] + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  ix[1] = fb.lo[1] + (int)(blockIdx.y * blockDim.y + threadIdx.y);
  ix[2] = fb.lo[2] + (int)(blockIdx.z * blockDim.z + threadIdx.z);
  if (ix[0] > fb.hi[0] || ix[1] > fb.hi[1] || ix[2] > fb.hi[2]) return;
  const double dt = a.dt;
  const double dt2 = dt / 2.0, dt3 = dt / 3.0, dt4 = dt / 4.0;
  const double half = 0.5, sixth = 1.0 / 6.0;
  const double* h = a.dx;
  auto U = [&](int q, const int* c) { return a.umac[q].p[a.umac[q].off32(c[0], c[1], c[2])]; };
  const double vn = U(D, ix);
  double sgn[3] = {0.0, 0.0, 0.0};
  int uc[3] = {ix[0], ix[1], ix[2]};
  if (vn > 0.0) { sgn[D] = 1.0; uc[D] -= 1; } else { sgn[D] = -1.0; }
  double del[3] = {0.0, 0.0, 0.0};
  del[D] = sgn[D] * 0.5 * h[D] - 0.5 * vn * dt;
  double se = bds_eval<DM>(a, uc, del);
  const double frc = a.force.p[a.force.off32(uc[0], uc[1], uc[2])];
  if (a.is_cons) {
    se = se * (1.0 - dt2 * dvel(a, D, uc)) + dt2 * frc;
  } else {
    double tsum;
    if (DM == 2) tsum = dvel(a, 1 - D, uc);
    else tsum = dvel(a, D == 0 ? 1 : 0, uc) + dvel(a, D == 2 ? 1 : 2, uc);
    se = se * (1.0 + dt2 * tsum) + dt2 * frc;
  }
#pragma unroll
  for (int t = 0; t < DM; ++t) {
    if (t == D) continue;
    const int r = (DM == 3) ? 3 - D - t : 0;
    for (int side = 1; side >= 0; --side) {
      int tf[3] = {uc[0], uc[1], uc[2]};
      tf[t] += side;
      const double vt = U(t, tf);
      int toff;
      if (vt > 0.0) { sgn[t] = 1.0; toff = side ? 0 : -1; } else { sgn[t] = -1.0; toff = side ? 1 : 0; }
      int fn[3] = {ix[0], ix[1], ix[2]};
      fn[t] += toff;
      double u = 0.0;
      {
        const double un = U(D, fn);
        if (vn * un > 0.0) u = un;
      }
      double p1[3] = {0.0, 0.0, 0.0}, p2[3] = {0.0, 0.0, 0.0}, p3[3] = {0.0, 0.0, 0.0}, p4[3] = {0.0, 0.0, 0.0};
      p1[D] = sgn[D] * 0.5 * h[D];
      p1[t] = sgn[t] * 0.5 * h[t];
      p2[D] = sgn[D] * 0.5 * h[D] - vn * dt;
      p2[t] = sgn[t] * 0.5 * h[t];
      p3[D] = sgn[D] * 0.5 * h[D] - u * dt;
      p3[t] = sgn[t] * 0.5 * h[t] - vt * dt;
      int c2[3] = {uc[0], uc[1], uc[2]};
      c2[t] += toff;
#pragma unroll
      for (int l = 0; l < 3; ++l) del[l] = (p2[l] + p3[l]) / 2.0;
      const double val1 = bds_eval<DM>(a, c2, del);
#pragma unroll
      for (int l = 0; l < 3; ++l) del[l] = (p1[l] + p3[l]) / 2.0;
      const double val2 = bds_eval<DM>(a, c2, del);
#pragma unroll
      for (int l = 0; l < 3; ++l) del[l] = (p1[l] + p2[l]) / 2.0;
      const double val3 = bds_eval<DM>(a, c2, del);
      double gamma = BDS_OVER((val1 + val2 + val3), 3.0, (1.0 / 3.0));
      if (DM == 2) {
        if (a.is_cons) gamma = gamma * (1.0 - dt3 * divu_of(a, 2, c2));
      } else {
        if (a.is_cons) gamma = gamma * (1.0 - dt3 * (dvel(a, D, c2) + dvel(a, t, c2)));
        else gamma = gamma * (1.0 + dt3 * dvel(a, r, c2));
        for (int rside = 1; rside >= 0; --rside) {
          int rf[3] = {c2[0], c2[1], c2[2]};
          rf[r] += rside;
          const double vr = U(r, rf);
          int roff;
          if (vr > 0.0) { sgn[r] = 1.0; roff = rside ? 0 : -1; } else { sgn[r] = -1.0; roff = rside ? 1 : 0; }
          int fnn[3] = {fn[0], fn[1], fn[2]};
          fnn[r] += roff;
          double uu = 0.0;
          {
            const double un = U(D, fnn);
            if (vn * un > 0.0) uu = un;
          }
          int tfr[3] = {tf[0], tf[1], tf[2]};
          tfr[r] += roff;
          double vv = 0.0;
          {
            const double un = U(t, tfr);
            if (vt * un > 0.0) vv = un;
          }
          double vt3 = vt;
          if (D == 2) {  // QUIRK (see header)
            int q[3] = {ix[0], ix[1], ix[2]};
            q[t] += side;
            q[r] += roff;
            vt3 = U(t, q);
          }
          p1[r] = sgn[r] * 0.5 * h[r];
          p2[r] = sgn[r] * 0.5 * h[r];
          p3[D] = sgn[D] * 0.5 * h[D] - vn * dt;
          p3[t] = sgn[t] * 0.5 * h[t] - vt3 * dt;
          p3[r] = sgn[r] * 0.5 * h[r];
          p4[D] = sgn[D] * 0.5 * h[D] - uu * dt;
          p4[t] = sgn[t] * 0.5 * h[t] - vv * dt;
          p4[r] = sgn[r] * 0.5 * h[r] - vr * dt;
          int c3[3] = {c2[0], c2[1], c2[2]};
          c3[r] += roff;
#pragma unroll
          for (int l = 0; l < 3; ++l) del[l] = (p1[l] + p2[l] + p3[l] + p4[l]) / 4.0;
          const double w1 = bds_eval<DM>(a, c3, del);
#pragma unroll
          for (int l = 0; l < 3; ++l) del[l] = half * p1[l] + sixth * (p2[l] + p3[l] + p4[l]);
          const double w2 = bds_eval<DM>(a, c3, del);
#pragma unroll
          for (int l = 0; l < 3; ++l) del[l] = half * p2[l] + sixth * (p1[l] + p3[l] + p4[l]);
          const double w3 = bds_eval<DM>(a, c3, del);
#pragma unroll
          for (int l = 0; l < 3; ++l) del[l] = half * p3[l] + sixth * (p2[l] + p1[l] + p4[l]);
          const double w4 = bds_eval<DM>(a, c3, del);
#pragma unroll
          for (int l = 0; l < 3; ++l) del[l] = half * p4[l] + sixth * (p2[l] + p3[l] + p1[l]);
          const double w5 = bds_eval<DM>(a, c3, del);
          double gamma2 = -0.8 * w1 + 0.45 * (w2 + w3 + w4 + w5);
          if (a.is_cons) gamma2 = gamma2 * (1.0 - dt4 * divu_of(a, 3, c3));
          gamma2 = gamma2 * vr;
          if (rside) gamma = gamma - BDS_OVER(dt * gamma2, (3.0 * h[r]), ((1.0 / 3.0) * a.rdx[r]));
          else gamma = gamma + BDS_OVER(dt * gamma2, (3.0 * h[r]), ((1.0 / 3.0) * a.rdx[r]));
        }
      }
      gamma = gamma * vt;
      if (side) se = se - BDS_OVER(dt * gamma, (2.0 * h[t]), (0.5 * a.rdx[t]));
      else se = se + BDS_OVER(dt * gamma, (2.0 * h[t]), (0.5 * a.rdx[t]));
    }
  }
  a.sedge[D](ix[0], ix[1], ix[2]) = se;
}

}  // namespace

void BDS_FN(bds_dev)(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                     const int* lo, const int* hi, int comp, bool is_cons, int ng_s, int ng_f);

#ifndef MGPU_BDS_FAST
size_t bds_scratch(const mgpu_params& P, const int* lo, const int* hi) {
  Box3 tb = grown(lo, hi, P.dm, 1);
  Box3 nb = tb;
  for (int d = 0; d < P.dm; ++d) nb.hi[d] += 1;
  return (size_t)(P.dm == 2 ? 3 : 7) * ((size_t)tb.npts() * sizeof(double) + 256) + (size_t)nb.npts() * sizeof(double) +
         4096;
}

static int g_bds_fast = 1;
void bds_set_fast(int on) { g_bds_fast = on; }
void bds_dev_fast(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                  const int* lo, const int* hi, int comp, bool is_cons, int ng_s, int ng_f);
void bds_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
             const int* lo, const int* hi, int comp, bool is_cons, int ng_s, int ng_f) {
  if (g_bds_fast) bds_dev_fast(P, s_full, sedge_full, umac, force_full, lo, hi, comp, is_cons, ng_s, ng_f);
  else bds_dev_exact(P, s_full, sedge_full, umac, force_full, lo, hi, comp, is_cons, ng_s, ng_f);
}
#endif

void BDS_FN(bds_dev)(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                     const int* lo, const int* hi, int comp, bool is_cons, int ng_s, int ng_f) {
  const int dm = P.dm;
  if (ng_s < 3) throw Error("bds: need at least 3 ghost cells");
  if (ng_f < 1) throw Error("bds: force needs at least 1 ghost cell");
  if (s_full.cs * 8 >= (1L << 31)) throw Error("bds: box too large for the 32-bit cell indices of the kernels");
  BdsArgs a;
  a.dm = dm;
  a.is_cons = is_cons;
  a.dt = P.dt;
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = d < dm ? lo[d] : 0;
    a.hi[d] = d < dm ? hi[d] : 0;
    a.dx[d] = P.dx[d < dm ? d : 0];
    a.rdx[d] = 1.0 / a.dx[d];
    if (d < dm) {
      a.umac[d] = umac[d];
      a.sedge[d] = sedge_full[d].comp(comp);
    }
  }
  a.tb = grown(lo, hi, dm, 1);
  a.vb = grown(lo, hi, dm, 0);
  a.nb = a.tb;
  for (int d = 0; d < dm; ++d) a.nb.hi[d] += 1;
  a.s = s_full.comp(comp);
  a.force = force_full.comp(comp);
  int z3[3] = {0, 0, 0};
  a.sint = make_view(arena_alloc((size_t)a.nb.npts()), a.nb.lo, a.nb.hi, dm, 0, z3, 1);
  const int nsl = dm == 2 ? 3 : 7;
  a.slope = make_view(arena_alloc((size_t)a.tb.npts() * nsl), a.tb.lo, a.tb.hi, dm, 0, z3, nsl);
  cudaStream_t st = ctx().stream;
  MGPU_TIMED(TAG_BDS, (k_bds_sint<<<grid3(a.nb, 256), block3(a.nb, 256), 0, st>>>(a)));
  MGPU_TIMED(TAG_BDS, (k_bds_slope<<<grid3(a.tb, 128), block3(a.tb, 128), 0, st>>>(a)));
  for (int d = 0; d < dm; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    int bx = 32, by = (dm == 3) ? 2 : 4, bz = (dm == 3) ? 2 : 1;
    if (const char* e = getenv("MGPU_BDS_BLOCK")) sscanf(e, "%d,%d,%d", &bx, &by, &bz);  // tuning aid
    if (dm == 2) bz = 1;
    if (bx < 1 || by < 1 || bz < 1 || bx * by * bz > 128) throw Error("bds: MGPU_BDS_BLOCK must describe 1..128 threads");
    const dim3 bt((unsigned)bx, (unsigned)by, (unsigned)bz);
    const dim3 nb((unsigned)((fb.hi[0] - fb.lo[0] + bx) / bx), (unsigned)((fb.hi[1] - fb.lo[1] + by) / by),
                  (unsigned)((fb.hi[2] - fb.lo[2] + bz) / bz));
    if (dm == 2) {
      if (d == 0) MGPU_TIMED(TAG_BDS, (k_bds_conc<2, 0><<<nb, bt, 0, st>>>(a)));
      else MGPU_TIMED(TAG_BDS, (k_bds_conc<2, 1><<<nb, bt, 0, st>>>(a)));
    } else {
      if (d == 0) MGPU_TIMED(TAG_BDS, (k_bds_conc<3, 0><<<nb, bt, 0, st>>>(a)));
      else if (d == 1) MGPU_TIMED(TAG_BDS, (k_bds_conc<3, 1><<<nb, bt, 0, st>>>(a)));
      else MGPU_TIMED(TAG_BDS, (k_bds_conc<3, 2><<<nb, bt, 0, st>>>(a)));
    }
  }
}

}  // namespace mgpu
