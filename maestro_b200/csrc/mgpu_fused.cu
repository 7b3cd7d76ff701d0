// Fused 3-D edge-state kernel: make_edge_scal_3d (Source/make_edge_scal.f90:677) + the ppm_3d /
// slope calls it makes (Source/ppm.f90:1629, Source/slope.f90) in ONE launch per component, with
// no intermediate array in HBM.  The reference streams ~45 box-sized temporaries through DRAM per
// component; here every temporary lives in shared memory or registers.
//
// Design ("2.5-D streaming"): a CTA owns a BX x BY column of cells (one thread per (i,j), one halo
// cell on each side in x and y because every stage reaches one cell sideways) and marches in z.
// The z-direction PPM stencil lives in a register window; the x/y stencils read a shared-memory
// tile of the current plane; the normal-predictor, transverse and final stages of plane q-1/q are
// exchanged between neighbouring threads through double-buffered shared-memory planes:
//
//   step q:  P1(q)   : Ip/Im in x,y,z of cell plane q          -> simhz(q)
//            S2(q)   : slx,srx,sly,sry, simhx(q), simhy(q)
//            T1(q)   : simhxy(q), simhyx(q)        T3(q): simhzx(q), simhzy(q)   (z-face q)
//            T2(q-1) : simhxz(q-1), simhyz(q-1)    (needs simhz(q-1), simhz(q))
//            F_z(q)  : sedgez on z-face q           F_xy(q-1): sedgex, sedgey of plane q-1
//
// Only cells (i,j) strictly inside the CTA's tile produce output, so each plane costs
// (BX*BY)/((BX-2)*(BY-2)) of the minimal work and nothing is recomputed along z except two planes
// per z-chunk.
//
// Template switches
//   BC   : true  -> every BC branch of the reference (EXT_DIR, FOEXTRAP/HOEXTRAP with the inflow clamp,
//                   REFLECT_EVEN/ODD, wall stencils of PPM/slopes, and its asymmetries) is evaluated from
//                   global indices, so the kernel is valid for any box;
//          false -> all six faces are INTERIOR (periodic / box-box): the BC code is compiled out.
//   FAST : false -> expression order and operations of the reference; this translation unit is built with
//                   -fmad=false, so results are bit-identical to the staged path and to the CPU oracle;
//          true  -> (mgpu_fused_fast.cu, built with -fmad=true) dt/dx is folded into one factor (no fp64
//                   division in the loop) and FMA contraction is allowed: differs from the reference in
//                   the last bits only (tests: <= 1e-12 relative, the north-star tolerance).
#include <cstring>

#include "mgpu_fused.cuh"
#include "mgpu_recon.cuh"

#ifndef MGPU_FAST
#define MGPU_FAST 0
#endif
#ifndef MGPU_FUSED_MINB
#define MGPU_FUSED_MINB 3  // resident CTAs per SM the register allocation is tuned for
#endif

namespace mgpu {
namespace {

// Shared-memory layout: the s tile, then planes of one common shape (BY+1 rows of pitch BX+1) so that every
// access is [per-thread base + compile-time constant]: six single planes, then two parity blocks of eleven planes.
template <int H, int BX, int BY>
struct FusedSmem {
  static constexpr int SP = BX + 2 * H;  // pitch of the s tile
  static constexpr int SN = (BY + 2 * H) * SP;
  static constexpr int P = BX + 1;        // pitch of every other plane
  static constexpr int PL = (BY + 1) * P; // doubles per plane
  enum { IPX = 0, IMX, IPY, IMY, XZ, YZ, NSINGLE };
  enum { Z = 0, SHX, SHY, XY, YX, ZX, ZY, U, V, W, FRC, NPAR };
  static constexpr int TOTAL = SN + (NSINGLE + 2 * NPAR) * PL;
  double buf[TOTAL];
};
// plane A of block B (a double* already offset to this thread's cell) at row offset dy, column offset dx
#define SMP(B, A, dy, dx) (B)[SM::A * SM::PL + (dy) * SM::P + (dx)]

template <bool BC>
__device__ __forceinline__ void bc_states(const FusedArgs& a, int d, int f, double s_lo_m1, double s_lo_0,
                                          double s_hi_p1, int stage, double& sl, double& sr) {
  // lr_bc of the staged path with the needed s values passed in registers:
  // s_lo_m1 = s(is-1), s_lo_0 = s(is) (only for the z-lo EXT_DIR quirk), s_hi_p1 = s(ie+1)
  if (!BC) return;
  const int is = a.lo[d], ie = a.hi[d];
  if (f == is) {
    const int bclo = a.bclo[d];
    if (bclo == MGPU_BC_EXT_DIR) {
      sl = (d == 2 && stage == 0) ? s_lo_0 : s_lo_m1;  // QUIRK make_edge_scal.f90:1010-1011
      sr = sl;
    } else if (bclo == MGPU_BC_FOEXTRAP || bclo == MGPU_BC_HOEXTRAP) {
      if (a.velnorm[d]) sr = dmin2(sr, 0.0);
      sl = sr;
    } else if (bclo == MGPU_BC_REFLECT_EVEN) {
      sl = sr;
    } else if (bclo == MGPU_BC_REFLECT_ODD) {
      sl = 0.0;
      sr = 0.0;
    }
  }
  if (f == ie + 1) {
    const int bchi = a.bchi[d];
    if (bchi == MGPU_BC_EXT_DIR) {
      sl = s_hi_p1;
      sr = sl;
    } else if (bchi == MGPU_BC_FOEXTRAP || bchi == MGPU_BC_HOEXTRAP) {
      if (a.velnorm[d]) sl = dmax2(sl, 0.0);
      sr = sl;
    } else if (bchi == MGPU_BC_REFLECT_EVEN) {
      sr = sl;
    } else if (bchi == MGPU_BC_REFLECT_ODD) {
      sl = 0.0;
      sr = 0.0;
    }
  }
}

__device__ __forceinline__ double final_bc(const FusedArgs& a, int d, int f, double e, double sedgel, double sedger,
                                           double s_left, double s_right) {
  if (f == a.lo[d]) {
    const int bc = a.bclo[d];
    if (bc == MGPU_BC_EXT_DIR) e = s_left;
    else if (bc == MGPU_BC_FOEXTRAP || bc == MGPU_BC_HOEXTRAP) e = a.velnorm[d] ? dmin2(sedger, 0.0) : sedger;
    else if (bc == MGPU_BC_REFLECT_EVEN) e = sedger;
    else if (bc == MGPU_BC_REFLECT_ODD) e = 0.0;
  }
  if (f == a.hi[d] + 1) {
    const int bc = a.bchi[d];
    if (bc == MGPU_BC_EXT_DIR) e = s_right;
    else if (bc == MGPU_BC_FOEXTRAP || bc == MGPU_BC_HOEXTRAP) e = a.velnorm[d] ? dmax2(sedgel, 0.0) : sedgel;
    else if (bc == MGPU_BC_REFLECT_EVEN) e = sedgel;
    else if (bc == MGPU_BC_REFLECT_ODD) e = 0.0;
  }
  return e;
}

__device__ __forceinline__ LineBC no_wall() {
  LineBC b;
  b.lo = -(1 << 30);
  b.hi = (1 << 30);
  b.wlo = false;
  b.whi = false;
  b.relimit_last = b.lo + 2;
  b.hi_reset = true;
  return b;
}

template <int PPM, bool BC, bool FAST, int BX, int BY>
__global__ void __launch_bounds__(BX* BY, MGPU_FUSED_MINB) k_fused_edge(FusedArgs a) {
  constexpr int H = (PPM == 2) ? 3 : 2;
  using SM = FusedSmem<H, BX, BY>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* const sS = reinterpret_cast<double*>(smem_raw);
  constexpr int SP = SM::SP;
  constexpr int NT = (SM::SN + BX * BY - 1) / (BX * BY);  // s-tile elements per thread

  const int tx = threadIdx.x, ty = threadIdx.y;
  double* const sg = sS + SM::SN + ty * SM::P + tx;  // this thread's cell in the single planes
  double* const pb = sg + SM::NSINGLE * SM::PL;      // ... in parity block 0
  const int ibase = a.lo[0] - 1 + blockIdx.x * (BX - 2);
  const int jbase = a.lo[1] - 1 + blockIdx.y * (BY - 2);
  const int i = ibase + tx, j = jbase + ty;
  const int kz0 = a.lo[2] + blockIdx.z * a.kchunk;
  const int kz1 = min(kz0 + a.kchunk - 1, a.hi[2]);
  const bool active = (i <= a.hi[0] + 1) && (j <= a.hi[1] + 1);
  // clamped coordinates for loads by inactive threads (values never used)
  const int ic = min(i, a.hi[0] + 1), jc = min(j, a.hi[1] + 1);

  const LineBC bx = BC ? make_linebc(3, 0, a.lo[0], a.hi[0], a.bclo[0], a.bchi[0]) : no_wall();
  const LineBC by = BC ? make_linebc(3, 1, a.lo[1], a.hi[1], a.bclo[1], a.bchi[1]) : no_wall();
  const LineBC bz = BC ? make_linebc(3, 2, a.lo[2], a.hi[2], a.bclo[2], a.bchi[2]) : no_wall();

  const double dt = a.dt, rel_eps = a.rel_eps;
  const double hx = a.dx[0], hy = a.dx[1], hz = a.dx[2];
  const double dt2 = 0.5 * dt;
  // transverse / final coefficients: (dt6/h) and (dt4/h) exactly as the reference forms them
  const double c6x = (dt / 6.0) / hx, c6y = (dt / 6.0) / hy, c6z = (dt / 6.0) / hz;
  const double c4x = (dt / 4.0) / hx, c4y = (dt / 4.0) / hy, c4z = (dt / 4.0) / hz;
  // tracing: FAST folds dt/h
  const double tdx = FAST ? dt / hx : dt, tdy = FAST ? dt / hy : dt, tdz = FAST ? dt / hz : dt;

  // ---- per-thread pointers (plane 0 of the fab) and plane strides, hoisted out of the march ----
  const long s_sz = a.s.stride(2), f_sz = a.force.stride(2);
  const long u_sz = a.umac[0].stride(2), v_sz = a.umac[1].stride(2), w_sz = a.umac[2].stride(2);
  const double* ps = a.s.p + a.s.off(ic, jc, 0);
  const double* pf = a.force.p + a.force.off(ic, jc, 0);
  const double* pu = a.umac[0].p + a.umac[0].off(ic, jc, 0);
  const double* pu1 = a.umac[0].p + a.umac[0].off(min(i + 1, a.hi[0] + 2), jc, 0);
  const double* pv = a.umac[1].p + a.umac[1].off(ic, jc, 0);
  const double* pv1 = a.umac[1].p + a.umac[1].off(ic, min(j + 1, a.hi[1] + 2), 0);
  const double* pw = a.umac[2].p + a.umac[2].off(ic, jc, 0);
  // note: off(.,.,0) is relative to plane index 0; plane k is reached with + k*stride (k may be negative)
  // s-tile: in-plane offsets of the elements this thread loads (clamped to the fab)
  int t_off[NT];
  {
    const int ilo = a.s.lo[0], ihi = a.s.lo[0] + a.s.n[0] - 1;
    const int jlo = a.s.lo[1], jhi = a.s.lo[1] + a.s.n[1] - 1;
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      const int t = ty * BX + tx + m * BX * BY;
      const int yy = t / SP, xx = t - yy * SP;
      int ii = ibase - H + xx, jj = jbase - H + yy;
      ii = max(ilo, min(ii, ihi));
      jj = max(jlo, min(jj, jhi));
      t_off[m] = (int)a.s.off(ii, jj, 0);
    }
  }
  const double* s0 = a.s.p;

  // z register window: sw[m] = s(i,j,q-H+m)
  double sw[2 * H + 1];
  {
    const int q0 = kz0 - 1;
#pragma unroll
    for (int m = 1; m <= 2 * H; ++m) sw[m] = ps[(long)(q0 - 1 - H + m) * s_sz];
    sw[0] = 0.0;
  }
  double Ipz_prev = 0.0;
  // carried z-direction van Leer slope of cell q and edge value on face q (PPM == 1 && !BC only)
  double dz_c = 0.0, ez_c = 0.0;
  if constexpr (PPM == 1 && !BC) {
    // before the first shift sw[1..4] = s(q0-2 .. q0+1), q0 = kz0-1
    const double dz_m = dsvl_of(&sw[2], 1);  // cell q0-1
    dz_c = dsvl_of(&sw[3], 1);               // cell q0
    double e = 0.5 * (sw[3] + sw[2]) - (1.0 / 6.0) * (dz_c - dz_m);
    double elo, ehi;
    dminmax(sw[3], sw[2], elo, ehi);
    e = dmax2(e, elo);
    ez_c = dmin2(e, ehi);
  }
  double slx_p = 0.0, srx_p = 0.0, sly_p = 0.0, sry_p = 0.0;  // plane q-1 face states
  double slz_q = 0.0, srz_q = 0.0;
  double f_p = 0.0;  // force(i,j,q-1)
  double s_p = 0.0;  // s(i,j,q-1)

  // software prefetch: the loads of plane q+1 are issued right after the first barrier of step q, so the
  // global-memory latency overlaps the compute of step q instead of stalling the smem stores of step q+1
  double pre_s, pre_S[NT], pre_u, pre_u1 = 0.0, pre_v, pre_v1 = 0.0, pre_w, pre_w1, pre_f;
  // running plane offsets (advanced by one plane per prefetch: no 64-bit multiplies in the march)
  long o_s = (long)(kz0 - 1) * s_sz, o_u = (long)(kz0 - 1) * u_sz, o_v = (long)(kz0 - 1) * v_sz,
       o_w = (long)(kz0 - 1) * w_sz, o_f = (long)(kz0 - 1) * f_sz;
  const long hs = (long)H * s_sz;
  auto prefetch = [&]() {  // loads plane q, where o_* address plane q; then advances to q+1
    pre_s = ps[o_s + hs];
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      const int t = ty * BX + tx + m * BX * BY;
      pre_S[m] = (t < SM::SN) ? s0[t_off[m] + o_s] : 0.0;
    }
    pre_u = pu[o_u];
    if (tx == BX - 1) pre_u1 = pu1[o_u];
    pre_v = pv[o_v];
    if (ty == BY - 1) pre_v1 = pv1[o_v];
    pre_w1 = pw[o_w + w_sz];
    pre_f = a.force_zero ? 0.0 : pf[o_f];
    o_s += s_sz; o_u += u_sz; o_v += v_sz; o_w += w_sz; o_f += f_sz;
  };
  pre_w = pw[o_w];
  prefetch();

  for (int q = kz0 - 1; q <= kz1 + 1; ++q) {
    const int par = q & 1;
    double* const cur = pb + par * (SM::NPAR * SM::PL);        // parity block of plane q
    double* const prv = pb + (par ^ 1) * (SM::NPAR * SM::PL);  // parity block of plane q-1
    __syncthreads();  // previous step's readers of U/V/W/FRC[par] and S are done
    // ---- S0: publish the prefetched plane q ------------------------------------------------------
#pragma unroll
    for (int m = 0; m < 2 * H; ++m) sw[m] = sw[m + 1];
    sw[2 * H] = pre_s;
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      const int t = ty * BX + tx + m * BX * BY;
      if (t < SM::SN) sS[t] = pre_S[m];
    }
    SMP(cur, U, 0, 0) = pre_u;
    if (tx == BX - 1) SMP(cur, U, 0, 1) = pre_u1;
    SMP(cur, V, 0, 0) = pre_v;
    if (ty == BY - 1) SMP(cur, V, 1, 0) = pre_v1;
    SMP(cur, W, 0, 0) = pre_w;
    const double wq1 = pre_w1;
    const double f_q = pre_f;
    SMP(cur, FRC, 0, 0) = f_q;
    pre_w = pre_w1;
    __syncthreads();
    if (q < kz1 + 1) prefetch();

    const double s_q = sw[H];
    const double uq = SMP(cur, U, 0, 0), uq1 = SMP(cur, U, 0, +1);
    const double vq = SMP(cur, V, 0, 0), vq1 = SMP(cur, V, +1, 0);
    const double wq = SMP(cur, W, 0, 0);
    const double* c = &sS[(ty + H) * SP + tx + H];

    // ---- S1: P1(q) ---------------------------------------------------------------------------
    {
      double Ip, Im;
      cell_states<FAST>(PPM, a.slope_order, c, 1, i, bx, uq1, uq, tdx, hx, rel_eps, Ip, Im);
      SMP(sg, IPX, 0, 0) = Ip;
      SMP(sg, IMX, 0, 0) = Im;
      cell_states<FAST>(PPM, a.slope_order, c, SP, j, by, vq1, vq, tdy, hy, rel_eps, Ip, Im);
      SMP(sg, IPY, 0, 0) = Ip;
      SMP(sg, IMY, 0, 0) = Im;
      if constexpr (PPM == 1 && !BC) {
        // z marches with the thread: the van Leer slope of cell q and the edge value on face q were computed by
        // the previous step (as those of cell q+1 / face q+1), so each step evaluates one slope and one edge
        // instead of four and two.  Same expressions as dsvl_of / sedge1_of => same bits.
        const double dz_n = dsvl_of(&sw[H + 1], 1);
        double e = 0.5 * (sw[H + 1] + sw[H]) - (1.0 / 6.0) * (dz_n - dz_c);
        double elo, ehi;
        dminmax(sw[H + 1], sw[H], elo, ehi);
        e = dmax2(e, elo);
        e = dmin2(e, ehi);
        double smz = ez_c, spz = e;
        cw_limit(sw[H], smz, spz);
        ppm_trace<FAST>(sw[H], smz, spz, wq1, wq, tdz, hz, rel_eps, Ip, Im);
        dz_c = dz_n;
        ez_c = e;
      } else {
        cell_states<FAST>(PPM, a.slope_order, &sw[H], 1, q, bz, wq1, wq, tdz, hz, rel_eps, Ip, Im);
      }
      // z-face q (between planes q-1 and q)
      slz_q = Ipz_prev;
      srz_q = Im;
      Ipz_prev = Ip;
      bc_states<BC>(a, 2, q, s_p, s_q, s_q, 0, slz_q, srz_q);
      SMP(cur, Z, 0, 0) = riemann(slz_q, srz_q, wq, rel_eps);
    }
    __syncthreads();

    // ---- S2: normal-predictor face states in plane q ------------------------------------------
    double slx_q = 0.0, srx_q = 0.0, sly_q = 0.0, sry_q = 0.0;
    if (tx >= 1) {
      slx_q = SMP(sg, IPX, 0, -1);
      srx_q = SMP(sg, IMX, 0, 0);
      bc_states<BC>(a, 0, i, c[-1], 0.0, s_q, 0, slx_q, srx_q);
      SMP(cur, SHX, 0, 0) = riemann(slx_q, srx_q, uq, rel_eps);
    }
    if (ty >= 1) {
      sly_q = SMP(sg, IPY, -1, 0);
      sry_q = SMP(sg, IMY, 0, 0);
      bc_states<BC>(a, 1, j, c[-SP], 0.0, s_q, 0, sly_q, sry_q);
      SMP(cur, SHY, 0, 0) = riemann(sly_q, sry_q, vq, rel_eps);
    }
    __syncthreads();

    // ---- S3: transverse states ------------------------------------------------------------------
    // T1(q): x-face corrected by y, y-face corrected by x
    if (tx >= 1 && ty <= BY - 2) {
      double l = slx_q - c6y * (SMP(cur, V, +1, -1) + SMP(cur, V, 0, -1)) *
                             (SMP(cur, SHY, +1, -1) - SMP(cur, SHY, 0, -1));
      double r = srx_q - c6y * (vq1 + vq) * (SMP(cur, SHY, +1, 0) - SMP(cur, SHY, 0, 0));
      bc_states<BC>(a, 0, i, c[-1], 0.0, s_q, 1, l, r);
      SMP(cur, XY, 0, 0) = riemann(l, r, uq, rel_eps);
    }
    if (ty >= 1 && tx <= BX - 2) {
      double l = sly_q - c6x * (SMP(cur, U, -1, +1) + SMP(cur, U, -1, 0)) *
                             (SMP(cur, SHX, -1, +1) - SMP(cur, SHX, -1, 0));
      double r = sry_q - c6x * (uq1 + uq) * (SMP(cur, SHX, 0, +1) - SMP(cur, SHX, 0, 0));
      bc_states<BC>(a, 1, j, c[-SP], 0.0, s_q, 1, l, r);
      SMP(cur, YX, 0, 0) = riemann(l, r, vq, rel_eps);
    }
    if (q >= kz0) {
      // T3(q): z-face q corrected by x / by y (left cell = plane q-1, right cell = plane q)
      if (tx <= BX - 2) {
        double l = slz_q - c6x * (SMP(prv, U, 0, +1) + SMP(prv, U, 0, 0)) *
                               (SMP(prv, SHX, 0, +1) - SMP(prv, SHX, 0, 0));
        double r = srz_q - c6x * (uq1 + uq) * (SMP(cur, SHX, 0, +1) - SMP(cur, SHX, 0, 0));
        bc_states<BC>(a, 2, q, s_p, s_q, s_q, 1, l, r);
        SMP(cur, ZX, 0, 0) = riemann(l, r, wq, rel_eps);
      }
      if (ty <= BY - 2) {
        double l = slz_q - c6y * (SMP(prv, V, +1, 0) + SMP(prv, V, 0, 0)) *
                               (SMP(prv, SHY, +1, 0) - SMP(prv, SHY, 0, 0));
        double r = srz_q - c6y * (vq1 + vq) * (SMP(cur, SHY, +1, 0) - SMP(cur, SHY, 0, 0));
        bc_states<BC>(a, 2, q, s_p, s_q, s_q, 1, l, r);
        SMP(cur, ZY, 0, 0) = riemann(l, r, wq, rel_eps);
      }
    }
    if (q >= kz0 + 1) {
      // T2(q-1): x- and y-faces of plane q-1 corrected by z.  w on z-faces q-1 (opar) and q (par).
      if (tx >= 1) {
        double l = slx_p - c6z * (SMP(cur, W, 0, -1) + SMP(prv, W, 0, -1)) *
                               (SMP(cur, Z, 0, -1) - SMP(prv, Z, 0, -1));
        double r = srx_p - c6z * (wq + SMP(prv, W, 0, 0)) * (SMP(cur, Z, 0, 0) - SMP(prv, Z, 0, 0));
        if (BC) {
          // s(is-1) of plane q-1 for EXT_DIR: re-read from global (rare branch)
          double slm = 0.0;
          if (i == a.lo[0] && a.bclo[0] == MGPU_BC_EXT_DIR) slm = a.s(i - 1, jc, q - 1);
          bc_states<BC>(a, 0, i, slm, 0.0, s_p, 1, l, r);
        }
        SMP(sg, XZ, 0, 0) = riemann(l, r, SMP(prv, U, 0, 0), rel_eps);
      }
      if (ty >= 1) {
        double l = sly_p - c6z * (SMP(cur, W, -1, 0) + SMP(prv, W, -1, 0)) *
                               (SMP(cur, Z, -1, 0) - SMP(prv, Z, -1, 0));
        double r = sry_p - c6z * (wq + SMP(prv, W, 0, 0)) * (SMP(cur, Z, 0, 0) - SMP(prv, Z, 0, 0));
        if (BC) {
          double slm = 0.0;
          if (j == a.lo[1] && a.bclo[1] == MGPU_BC_EXT_DIR) slm = a.s(ic, j - 1, q - 1);
          bc_states<BC>(a, 1, j, slm, 0.0, s_p, 1, l, r);
        }
        SMP(sg, YZ, 0, 0) = riemann(l, r, SMP(prv, V, 0, 0), rel_eps);
      }
    }
    __syncthreads();

    // ---- S4: final edge states ------------------------------------------------------------------
    const bool inx = (tx >= 1 && tx <= BX - 2 && i <= a.hi[0]);
    const bool iny = (ty >= 1 && ty <= BY - 2 && j <= a.hi[1]);
    if (q >= kz0 && inx && iny) {
      // F_z(q): transverse terms x then y, left cell plane q-1 (opar), right cell plane q (par)
      double el = slz_q - c4x * (SMP(prv, U, 0, +1) + SMP(prv, U, 0, 0)) * (SMP(prv, XY, 0, +1) - SMP(prv, XY, 0, 0)) -
                  c4y * (SMP(prv, V, +1, 0) + SMP(prv, V, 0, 0)) * (SMP(prv, YX, +1, 0) - SMP(prv, YX, 0, 0)) +
                  dt2 * f_p;
      double er = srz_q - c4x * (uq1 + uq) * (SMP(cur, XY, 0, +1) - SMP(cur, XY, 0, 0)) -
                  c4y * (vq1 + vq) * (SMP(cur, YX, +1, 0) - SMP(cur, YX, 0, 0)) + dt2 * f_q;
      double e = riemann(el, er, wq, rel_eps);
      if (BC) e = final_bc(a, 2, q, e, el, er, s_p, s_q);
      if (q <= kz1 || q == a.hi[2] + 1) a.sedge[2](i, j, q) = e;
    }
    if (q >= kz0 + 1 && active) {
      const int k = q - 1;
      // F_xy(k): x-face (i,j,k): transverse terms y (simhyz) then z (simhzy)
      if (tx >= 1 && (tx <= BX - 2 || i == a.hi[0] + 1) && iny) {
        double el = slx_p - c4y * (SMP(prv, V, +1, -1) + SMP(prv, V, 0, -1)) * (SMP(sg, YZ, +1, -1) - SMP(sg, YZ, 0, -1)) -
                    c4z * (SMP(cur, W, 0, -1) + SMP(prv, W, 0, -1)) * (SMP(cur, ZY, 0, -1) - SMP(prv, ZY, 0, -1)) +
                    dt2 * SMP(prv, FRC, 0, -1);
        double er = srx_p - c4y * (SMP(prv, V, +1, 0) + SMP(prv, V, 0, 0)) * (SMP(sg, YZ, +1, 0) - SMP(sg, YZ, 0, 0)) -
                    c4z * (wq + SMP(prv, W, 0, 0)) * (SMP(cur, ZY, 0, 0) - SMP(prv, ZY, 0, 0)) + dt2 * f_p;
        double e = riemann(el, er, SMP(prv, U, 0, 0), rel_eps);
        if (BC && (i == a.lo[0] || i == a.hi[0] + 1)) {
          const double sl_c = (a.bclo[0] == MGPU_BC_EXT_DIR && i == a.lo[0]) ? a.s(i - 1, j, k) : 0.0;
          e = final_bc(a, 0, i, e, el, er, sl_c, s_p);
        }
        a.sedge[0](i, j, k) = e;
      }
      // y-face (i,j,k): transverse terms x (simhxz) then z (simhzx)
      if (ty >= 1 && (ty <= BY - 2 || j == a.hi[1] + 1) && inx) {
        double el = sly_p - c4x * (SMP(prv, U, -1, +1) + SMP(prv, U, -1, 0)) * (SMP(sg, XZ, -1, +1) - SMP(sg, XZ, -1, 0)) -
                    c4z * (SMP(cur, W, -1, 0) + SMP(prv, W, -1, 0)) * (SMP(cur, ZX, -1, 0) - SMP(prv, ZX, -1, 0)) +
                    dt2 * SMP(prv, FRC, -1, 0);
        double er = sry_p - c4x * (SMP(prv, U, 0, +1) + SMP(prv, U, 0, 0)) * (SMP(sg, XZ, 0, +1) - SMP(sg, XZ, 0, 0)) -
                    c4z * (wq + SMP(prv, W, 0, 0)) * (SMP(cur, ZX, 0, 0) - SMP(prv, ZX, 0, 0)) + dt2 * f_p;
        double e = riemann(el, er, SMP(prv, V, 0, 0), rel_eps);
        if (BC && (j == a.lo[1] || j == a.hi[1] + 1)) {
          const double sl_c = (a.bclo[1] == MGPU_BC_EXT_DIR && j == a.lo[1]) ? a.s(i, j - 1, k) : 0.0;
          e = final_bc(a, 1, j, e, el, er, sl_c, s_p);
        }
        a.sedge[1](i, j, k) = e;
      }
    }
    // rotate plane q -> q-1
    slx_p = slx_q; srx_p = srx_q; sly_p = sly_q; sry_p = sry_q;
    f_p = f_q;
    s_p = s_q;
  }
}

template <int PPM, bool BC, int BX, int BY>
void launch_fused(const FusedArgs& a, int nx, int ny, int nz) {
  constexpr int H = (PPM == 2) ? 3 : 2;
  using SM = FusedSmem<H, BX, BY>;
  Context& c = ctx();
  static bool configured = false;
  auto kern = k_fused_edge<PPM, BC, (MGPU_FAST != 0), BX, BY>;
  if (!configured) {
    MGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
    configured = true;
  }
  dim3 block(BX, BY, 1);
  dim3 grid((nx + BX - 3) / (BX - 2), (ny + BY - 3) / (BY - 2), (nz + a.kchunk - 1) / a.kchunk);
  MGPU_TIMED(TAG_FUSED_EDGE, (kern<<<grid, block, sizeof(SM), c.stream>>>(a)));
}

}  // namespace

#if MGPU_FAST
void fused_edge_launch_fast(const FusedArgs& a, int ppm_type, bool bc, int nx, int ny, int nz)
#else
void fused_edge_launch_exact(const FusedArgs& a, int ppm_type, bool bc, int nx, int ny, int nz)
#endif
{
  constexpr int BX = MGPU_FUSED_BX, BY = MGPU_FUSED_BY;
  switch (ppm_type * 2 + (bc ? 1 : 0)) {
    case 0: launch_fused<0, false, BX, BY>(a, nx, ny, nz); break;
    case 1: launch_fused<0, true, BX, BY>(a, nx, ny, nz); break;
    case 2: launch_fused<1, false, BX, BY>(a, nx, ny, nz); break;
    case 3: launch_fused<1, true, BX, BY>(a, nx, ny, nz); break;
    case 4: launch_fused<2, false, BX, BY>(a, nx, ny, nz); break;
    case 5: launch_fused<2, true, BX, BY>(a, nx, ny, nz); break;
    default: throw Error("make_edge_scal: invalid ppm_type");
  }
}

#if !MGPU_FAST
static int g_variant = 1;
void fused_edge_set_variant(int v) { g_variant = v; }

bool fused_edge_supported(const mgpu_params& P, bool is_cons) {
  return P.dm == 3 && P.bds_type == 0 && P.ppm_trace_forces == 0 && !is_cons;
}

bool fused_edge_is_upwind_first(const mgpu_params& P, const int* adv_bc, int bccomp, bool exact) {
  if (exact || g_variant == 0 || P.dm != 3) return false;
  for (int d = 0; d < 3; ++d)
    for (int side = 0; side < 2; ++side)
      if (adv_bc[d + 3 * (side + 2 * (bccomp - 1))] != MGPU_BC_INTERIOR) return false;
  return true;
}

bool fused_edge2d_supported(const mgpu_params& P, bool is_cons, const int* adv_bc, int bccomp, bool exact) {
  if (exact || g_variant == 0 || P.dm != 2 || P.bds_type != 0 || P.ppm_trace_forces != 0 || is_cons) return false;
  for (int d = 0; d < 2; ++d)
    for (int side = 0; side < 2; ++side)
      if (adv_bc[d + 2 * (side + 2 * (bccomp - 1))] == MGPU_BC_REFLECT_ODD) return false;
  return true;
}

void fused_edge2d_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                      const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel, int ng_s,
                      int ng_f, bool force_zero) {
  if (P.ppm_type == 2 && ng_s < 4) throw Error("Need 4 ghost cells for ppm_type=2");  // ppm.f90:576-578
  if (ng_s < 3) throw Error("make_edge_scal: need at least 3 ghost cells");
  if (ng_f < 1) throw Error("make_edge_scal: force needs at least 1 ghost cell");
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.slope_order = P.slope_order;
  a.force_zero = force_zero;
  a.dt = P.dt;
  a.rel_eps = P.rel_eps;
  bool any_bc = false;
  for (int d = 0; d < 2; ++d) {
    a.lo[d] = lo[d];
    a.hi[d] = hi[d];
    a.dx[d] = P.dx[d];
    a.bclo[d] = adv_bc[d + 2 * (0 + 2 * (bccomp - 1))];
    a.bchi[d] = adv_bc[d + 2 * (1 + 2 * (bccomp - 1))];
    if (a.bclo[d] != MGPU_BC_INTERIOR || a.bchi[d] != MGPU_BC_INTERIOR) any_bc = true;
    a.velnorm[d] = is_vel && (comp == d);
    a.umac[d] = umac[d];
    a.sedge[d] = sedge_full[d].comp(comp);
  }
  a.s = s_full.comp(comp);
  a.force = force_full.comp(comp);
  fused_edge2d_launch(a, P.ppm_type, hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, any_bc);
}

void fused_edge_dev(const mgpu_params& P, const DV& s_full, DV* sedge_full, const DV* umac, const DV& force_full,
                    const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel, int ng_s,
                    int ng_f, int kchunk, bool exact, bool force_zero, const double* sdiv, const double* ssub,
                    const double* wadd) {
  if (P.ppm_type == 2 && ng_s < 4) throw Error("Need 4 ghost cells for ppm_type=2");  // ppm.f90:1864-1866
  if (ng_s < 3) throw Error("make_edge_scal: need at least 3 ghost cells");
  if (ng_f < 1) throw Error("make_edge_scal: force needs at least 1 ghost cell");
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.slope_order = P.slope_order;
  a.force_zero = force_zero;
  a.sdiv = sdiv;
  a.ssub = ssub;
  a.wadd = wadd;
  a.dt = P.dt;
  a.rel_eps = P.rel_eps;
  bool any_bc = false;
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = lo[d];
    a.hi[d] = hi[d];
    a.dx[d] = P.dx[d];
    a.bclo[d] = adv_bc[d + 3 * (0 + 2 * (bccomp - 1))];
    a.bchi[d] = adv_bc[d + 3 * (1 + 2 * (bccomp - 1))];
    if (a.bclo[d] != MGPU_BC_INTERIOR || a.bchi[d] != MGPU_BC_INTERIOR) any_bc = true;
    a.velnorm[d] = is_vel && (comp == d);
    a.umac[d] = umac[d];
    a.sedge[d] = sedge_full[d].comp(comp);
  }
  a.s = s_full.comp(comp);
  a.force = force_full.comp(comp);
  const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
  // kchunk < 0: the upwind-first kernel balances the chunk against its resident CTA slots (fused2_auto_kchunk);
  // the literal kernels keep 32 planes
  a.kchunk = kchunk > 0 ? kchunk : (kchunk < 0 ? -1 : nz);
  if (a.kchunk > nz) a.kchunk = nz;
  const bool xform = sdiv || ssub || wadd;
  if (xform && (exact || any_bc || g_variant == 0))
    throw Error("make_edge_scal: on-the-fly input transforms need the upwind-first kernel (internal error)");
  const bool to_fused2 = !exact && (g_variant == 1 || (g_variant == 2 && !any_bc) || g_variant == 3 || g_variant == 4);
  if (a.kchunk < 0 && !to_fused2) a.kchunk = nz < 32 ? nz : 32;
  if (exact)
    fused_edge_launch_exact(a, P.ppm_type, any_bc, nx, ny, nz);
  else if (to_fused2)
    // (variant 1, the default: the upwind-first kernel for every FAST prediction -- since the warp-uniform slow-face
    //  branches it also wins for ppm_type 2 on boxes with physical boundaries, profiles/r01k_final.md; variant 2
    //  keeps the literal kernel on such boxes, variant 0 everywhere)
    if (g_variant != 4 && fused_edge3_supported(a, any_bc)) fused_edge3_launch(a, P.ppm_type, nx, ny, nz, any_bc);
    else fused_edge2_launch(a, P.ppm_type, nx, ny, nz, any_bc);
  else
    fused_edge_launch_fast(a, P.ppm_type, any_bc, nx, ny, nz);
}
#endif

}  // namespace mgpu
