// mkutrans (Source/mkutrans.f90: _2d :257, _3d :461) and velpred (Source/velpred.f90: _2d :266, _3d :640),
// planar geometry, every ppm_type and every physical BC branch.
//
// Stage structure (temporaries in the arena; the per-cell PPM/slope work is re-evaluated by the two faces
// that need it instead of being stored, so no Ip/Im arrays ever reach HBM):
//   k_mkutrans   : one launch per direction: 1-D extrapolation of the normal component, BCs, Riemann with w0
//   k_vp_face    : per face direction d: u_L^d, u_R^d of every component (+BCs) and the transverse
//                  components upwinded by utrans                                 (velpred.f90:803-1129)
// Both are fp64-compute bound (a PPM reconstruction per component), so they are CELL-centred: a thread reconstructs
// its own cell once (Ip and Im together), hands Ip to the thread of the next cell through shared memory, and that
// thread finishes the face between them.  Blocks overlap by one cell along d (1/128 resp. 1/8 of the threads).
//   k_vp_trans   : 3-D only, the six corner-coupled states                      (:1139-1558)
//   k_vp_final   : umac_L/R, Riemann with w0, BCs                               (:1562-1851, 2-D :528-630)
#include "mgpu_recon.cuh"
#include "mgpu_velpred.cuh"

// Two builds of this file (like mgpu_bds.cu): the exact one (-fmad=false, the reference's divisions: bit-identical) and
// the FAST one (mgpu_velpred_fast.cu: FMA contraction, dt/h and the dt/(4h), dt/(6h) factors as multiplications by
// reciprocals formed on the host; <= 1e-12), the default.  `exact = 1` selects the first.
#ifdef MGPU_VP_FAST
#define VP_OVER(x, d, rd) ((x) * (rd))
#define VP_FN(name) name##_fast
#else
#define VP_OVER(x, d, rd) ((x) / (d))
#define VP_FN(name) name##_exact
#endif

namespace mgpu {

namespace {

__device__ __forceinline__ bool wall3(int bc) {
  return bc == MGPU_BC_SLIP_WALL || bc == MGPU_BC_NO_SLIP_WALL || bc == MGPU_BC_SYMMETRY;
}

// 1-D extrapolated states of one cell with the CELL-centred velocity (ppm is_umac=.false.)
// form3d: velpred_3d writes the slope predictor as (1/2 - dt2*max(0,u)/h); mkutrans and velpred_2d as
// (1/2 - (dt2/h)*max(0,u))  (velpred.f90:812-813 vs :384-392, mkutrans.f90:543-547)
__device__ __forceinline__ void vel_cell_states(int ppm_type, int slope_order, bool form3d, const double* q, long st,
                                                int c, const LineBC& b, double ucell, double dt, double h, double rh,
                                                double rel_eps, double& Ip, double& Im) {
  (void)rh;
  if (ppm_type == 0) {
    const double sl = slope_cell(q, st, c, b, slope_order);
    const double dt2 = 0.5 * dt;
    if (form3d) {
      Ip = q[0] + (0.5 - VP_OVER(dt2 * dmax2(0.0, ucell), h, rh)) * sl;
      Im = q[0] - (0.5 + VP_OVER(dt2 * dmin2(0.0, ucell), h, rh)) * sl;
    } else {
      Ip = q[0] + (0.5 - VP_OVER(dt2, h, rh) * dmax2(0.0, ucell)) * sl;
      Im = q[0] - (0.5 + VP_OVER(dt2, h, rh) * dmin2(0.0, ucell)) * sl;
    }
  } else {
    double sm, sp;
    if (ppm_type == 1) ppm1_cell(q, st, c, b, sm, sp);
    else ppm2_cell(q, st, c, b, sm, sp);
#ifdef MGPU_VP_FAST
    ppm_trace<true>(q[0], sm, sp, ucell, ucell, dt * rh, h, rel_eps, Ip, Im);
#else
    ppm_trace<false>(q[0], sm, sp, ucell, ucell, dt, h, rel_eps, Ip, Im);
#endif
  }
}

// Riemann problem with the full velocity (mkutrans.f90:618-631, velpred.f90:1590-1621)
__device__ __forceinline__ double riemann_full(double l, double r, bool radial, double w0, double rel_eps) {
  const double uavg = 0.5 * (l + r);
  bool test;
  double v;
  if (radial) {
    test = ((l + w0 <= 0.0 && r + w0 >= 0.0) || (fabs(l + r + 2.0 * w0) < rel_eps));
    v = (uavg + w0 > 0.0) ? l : r;
  } else {
    test = ((l <= 0.0 && r >= 0.0) || (fabs(l + r) < rel_eps));
    v = (uavg > 0.0) ? l : r;
  }
  return test ? 0.0 : v;
}

__device__ __forceinline__ double upwind_trans(double l, double r, double ut, double rel_eps) {
  const double v = (ut > 0.0) ? l : r;
  const double uavg = 0.5 * (l + r);
  return (fabs(ut) < rel_eps) ? uavg : v;
}

// ------------------------------------------------------------------------------------------
// thread -> cell of a block that overlaps its neighbour by one cell along D: x blocks of VP_NX threads for D = 0,
// (32, VP_ND) blocks otherwise.  ix: the cell; first: the block's first cell along D (it only feeds its neighbour);
// returns false for threads past the face box fb (hi + 1 along D included as a cell: the right cell of the last face)
constexpr int VP_NX = 128, VP_ND = 8;
template <int D>
__device__ __forceinline__ bool vp_cell(const Box3& fb, int* ix, bool& first, int& slot) {
  if (D == 0) {
    ix[0] = fb.lo[0] - 1 + (int)blockIdx.x * (VP_NX - 1) + (int)threadIdx.x;
    ix[1] = fb.lo[1] + (int)blockIdx.y;
    ix[2] = fb.lo[2] + (int)blockIdx.z;
    first = threadIdx.x == 0;
    slot = threadIdx.x;
    return ix[0] <= fb.hi[0];
  }
  constexpr int T = (D == 1) ? 2 : 1;  // the other transverse direction
  ix[0] = fb.lo[0] + (int)blockIdx.x * 32 + (int)threadIdx.x;
  ix[D] = fb.lo[D] - 1 + (int)blockIdx.y * (VP_ND - 1) + (int)threadIdx.y;
  ix[T] = fb.lo[T] + (int)blockIdx.z;
  first = threadIdx.y == 0;
  slot = threadIdx.y * 32 + threadIdx.x;
  return ix[0] <= fb.hi[0] && ix[D] <= fb.hi[D];
}
template <int D>
__device__ __forceinline__ int vp_left(int slot) { return D == 0 ? slot - 1 : slot - 32; }
template <int D>
inline dim3 vp_grid(const Box3& fb) {
  const int n0 = fb.hi[0] - fb.lo[0] + 1, nd = fb.hi[D] - fb.lo[D] + 1;
  if (D == 0) return dim3((unsigned)((n0 + VP_NX - 2) / (VP_NX - 1)), (unsigned)(fb.hi[1] - fb.lo[1] + 1), (unsigned)(fb.hi[2] - fb.lo[2] + 1));
  constexpr int T = (D == 1) ? 2 : 1;
  return dim3((unsigned)((n0 + 31) / 32), (unsigned)((nd + VP_ND - 2) / (VP_ND - 1)), (unsigned)(fb.hi[T] - fb.lo[T] + 1));
}
template <int D>
inline dim3 vp_block() { return D == 0 ? dim3(VP_NX, 1, 1) : dim3(32, VP_ND, 1); }

template <int D, int PPM>
__global__ void __launch_bounds__(256) k_mkutrans(VpArgs a) {
  constexpr int d = D;
  __shared__ double sh_ip[256];
  int ix[3], slot;
  bool first;
  Box3 fb = a.vb;
  fb.hi[d] += 1;
  const bool in = vp_cell<D>(fb, ix, first, slot);
  const DV u = a.utilde.comp(d), uf = a.ufull.comp(d);
  const long st = u.stride(d);
  double ul = 0.0, ur = 0.0;
  const double* q = nullptr;
  if (in) {  // this cell's two extrapolated states: Ip feeds the face above, Im the face below (this thread's face)
    const LineBC b = make_linebc(a.dm, d, a.lo[d], a.hi[d], a.bclo[d][d], a.bchi[d][d]);
    q = u.p + u.off(ix[0], ix[1], ix[2]);
    double ip;
    vel_cell_states(PPM, a.slope_order, false, q, st, ix[d], b, uf.p[uf.off(ix[0], ix[1], ix[2])], a.dt, a.dx[d],
                    a.rdx[d], a.rel_eps, ip, ur);
    sh_ip[slot] = ip;
  }
  __syncthreads();
  if (!in || first) return;
  ul = sh_ip[vp_left<D>(slot)];
  if (ix[d] == a.lo[d]) {
    const int p = a.plo[d];
    if (p == MGPU_BC_INLET) { ul = q[-st]; ur = q[-st]; }
    else if (wall3(p)) { ul = 0.0; ur = 0.0; }
    else if (p == MGPU_BC_OUTLET) { ul = dmin2(ur, 0.0); ur = ul; }
  }
  if (ix[d] == a.hi[d] + 1) {
    const int p = a.phi[d];
    if (p == MGPU_BC_INLET) { ul = q[0]; ur = q[0]; }
    else if (wall3(p)) { ul = 0.0; ur = 0.0; }
    else if (p == MGPU_BC_OUTLET) { ul = dmax2(ul, 0.0); ur = ul; }
  }
  const bool radial = a.spherical || (d == a.dm - 1);
  const double w0f = a.spherical ? a.w0mac[d](ix[0], ix[1], ix[2]) : (radial ? a.w0[ix[d]] : 0.0);
  a.utrans[d](ix[0], ix[1], ix[2]) = riemann_full(ul, ur, radial, w0f, a.rel_eps);
}

// ------------------------------------------------------------------------------------------
// UT = true (advance_premac on a box without physical boundaries, `velpred_premac_begin`): the face's transverse
// velocity is formed here as well -- mkutrans reconstructs u_d along d exactly as this kernel does (mkutrans.f90:529,
// 641,749 against velpred.f90:803,933,1063: the same ppm_3d call on the same data), so the d-component's (ul, ur) pair
// goes through mkutrans' Riemann problem (mkutrans.f90:618-631) and the kernel stores utrans (for mk_vel_force) and
// utrans + w0 (addw0, advance_premac.f90:109) on the whole face box, ghost rows included: no k_mkutrans launches, no
// addw0, no ghost fill of utrans.
template <int D, int PPM, bool UT>
__device__ __forceinline__ void vp_face_body(const VpArgs& a) {
  constexpr int d = D;
  __shared__ double sh_ip[3][256];
  int ix[3], slot;
  bool first;
  Box3 fb = a.tb;
  fb.lo[d] = a.lo[d];  // faces lo..hi+1 in d, lo-1..hi+1 transverse
  const bool in = vp_cell<D>(fb, ix, first, slot);
  const int dm = a.dm;
  const long uo = a.utilde.off(ix[0], ix[1], ix[2]);
  const long st = a.utilde.stride(d);
  double ul[3], ur[3];
  if (in) {  // this cell's states in direction d, every component: Ip feeds the face above, Im this thread's face
    double uc;
    if (UT && a.ufull_otf) {  // ufull = w0 on the cells + utilde (advance_premac.f90:75-78) formed here: the cell average
      uc = a.utilde.p[uo + a.utilde.cs * d];  // of w0 at the periodic image of a ghost cell is what the ghost fill copies
      if (d == dm - 1) {
        int kk = ix[d] % a.nr;
        if (kk < 0) kk += a.nr;
        uc = 0.5 * (a.w0[kk] + a.w0[kk + 1]) + uc;
      }
    } else {
      const DV ufd = a.ufull.comp(d);
      uc = ufd.p[ufd.off(ix[0], ix[1], ix[2])];
    }
    _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) {
      const LineBC b = make_linebc(dm, d, a.lo[d], a.hi[d], a.bclo[c][d], a.bchi[c][d]);
      const double* q = a.utilde.p + uo + a.utilde.cs * c;
      double ip;
      vel_cell_states(PPM, a.slope_order, dm == 3, q, st, ix[d], b, uc, a.dt, a.dx[d], a.rdx[d], a.rel_eps, ip, ur[c]);
      sh_ip[c][slot] = ip;
    }
  }
  __syncthreads();
  if (!in || first) return;
  _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ul[c] = sh_ip[c][vp_left<D>(slot)];
  if (!UT && ix[d] == a.lo[d]) {
    const int p = a.plo[d];
    if (p == MGPU_BC_INLET) {
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ul[c] = ur[c] = a.utilde.p[uo - st + a.utilde.cs * c];
    } else if (p == MGPU_BC_SLIP_WALL || p == MGPU_BC_SYMMETRY) {
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) {
        if (c == d) ul[c] = ur[c] = 0.0;
        else ul[c] = ur[c];
      }
    } else if (p == MGPU_BC_NO_SLIP_WALL) {
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ul[c] = ur[c] = 0.0;
    } else if (p == MGPU_BC_OUTLET) {
      ur[d] = dmin2(ur[d], 0.0);
      if (d == 0 && dm == 2) {  // QUIRK velpred.f90:415-417: copies the wrong way (urx = ulx)
        _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ur[c] = ul[c];
      } else if (d == 0 && dm == 3) {  // QUIRK velpred.f90:861-862: self-assignment
      } else {
        _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ul[c] = ur[c];
      }
    }
  }
  if (!UT && ix[d] == a.hi[d] + 1) {
    const int p = a.phi[d];
    if (p == MGPU_BC_INLET) {
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ul[c] = ur[c] = a.utilde.p[uo + a.utilde.cs * c];
    } else if (p == MGPU_BC_SLIP_WALL || p == MGPU_BC_SYMMETRY) {
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) {
        if (c == d) ul[c] = ur[c] = 0.0;
        else ur[c] = ul[c];
      }
    } else if (p == MGPU_BC_NO_SLIP_WALL) {
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ul[c] = ur[c] = 0.0;
    } else if (p == MGPU_BC_OUTLET) {
      ul[d] = dmax2(ul[d], 0.0);
      _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) ur[c] = ul[c];
    }
  }
  double ut;
  if (UT) {
    const bool radial = d == dm - 1;  // planar only (velpred_premac_fusable)
    const double w0f = radial ? a.w0[ix[d]] : 0.0;
    const double upre = riemann_full(ul[d], ur[d], radial, w0f, a.rel_eps);
    ut = radial ? upre + w0f : upre;
    a.utpre[d](ix[0], ix[1], ix[2]) = upre;
    if (radial) a.utrans[d](ix[0], ix[1], ix[2]) = ut;
  } else {
    ut = a.utrans[d](ix[0], ix[1], ix[2]);
  }
  const long to = a.UL[d].off(ix[0], ix[1], ix[2]);
  _Pragma("unroll") for (int c = 0; c < 3; ++c) if (c < dm) {
    a.UL[d].p[to + a.UL[d].cs * c] = ul[c];
    a.UR[d].p[to + a.UR[d].cs * c] = ur[c];
    if (c != d) a.UIMH[d].p[to + a.UIMH[d].cs * c] = upwind_trans(ul[c], ur[c], ut, a.rel_eps);
  }
}
template <int D, int PPM>
__global__ void __launch_bounds__(256) k_vp_face(VpArgs a) { vp_face_body<D, PPM, false>(a); }
template <int D, int PPM>
__global__ void __launch_bounds__(256) k_vp_face_ut(VpArgs a) { vp_face_body<D, PPM, true>(a); }

// coef * (trans_t(cell+e_t) + trans_t(cell)) * (q(cell+e_t) - q(cell))
__device__ __forceinline__ double tterm(double coef, const DV& tr, const DV& q, long qcomp_off, int t, int ci, int cj,
                                        int ck) {
  const long to = tr.off(ci, cj, ck), qo = q.off(ci, cj, ck) + qcomp_off;
  return coef * (tr.p[to + tr.stride(t)] + tr.p[to]) * (q.p[qo + q.stride(t)] - q.p[qo]);
}

__global__ void k_vp_trans(VpArgs a) {
  int ix[3];
  if (!decode3(a.tb, ix)) return;
  const double dt6 = a.dt / 6.0;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (d == c) continue;
      const int t = 3 - c - d;
      if (ix[d] < a.lo[d] || ix[t] < a.lo[t] || ix[t] > a.hi[t]) continue;
      int cl[3] = {ix[0], ix[1], ix[2]};
      cl[d] -= 1;
      const long fo = a.UL[d].off(ix[0], ix[1], ix[2]) + a.UL[d].cs * c;
      const long qc = a.UIMH[t].cs * c;
      double ql = a.UL[d].p[fo] - tterm(VP_OVER(dt6, a.dx[t], a.rdx[t]), a.utrans[t], a.UIMH[t], qc, t, cl[0], cl[1], cl[2]);
      double qr = a.UR[d].p[fo] - tterm(VP_OVER(dt6, a.dx[t], a.rdx[t]), a.utrans[t], a.UIMH[t], qc, t, ix[0], ix[1], ix[2]);
      if (ix[d] == a.lo[d]) {
        const int p = a.plo[d];
        if (p == MGPU_BC_INLET) ql = qr = a.utilde(cl[0], cl[1], cl[2], c);
        else if (p == MGPU_BC_SLIP_WALL || p == MGPU_BC_SYMMETRY || p == MGPU_BC_OUTLET) ql = qr;
        else if (p == MGPU_BC_NO_SLIP_WALL) ql = qr = 0.0;
      }
      if (ix[d] == a.hi[d] + 1) {
        const int p = a.phi[d];
        if (p == MGPU_BC_INLET) ql = qr = a.utilde(ix[0], ix[1], ix[2], c);
        else if (p == MGPU_BC_SLIP_WALL || p == MGPU_BC_SYMMETRY || p == MGPU_BC_OUTLET) qr = ql;
        else if (p == MGPU_BC_NO_SLIP_WALL) ql = qr = 0.0;
      }
      a.Q[c][d](ix[0], ix[1], ix[2]) = upwind_trans(ql, qr, a.utrans[d](ix[0], ix[1], ix[2]), a.rel_eps);
    }
}

template <int D, int PPM>
__global__ void k_vp_final(VpArgs a) {
  constexpr int d = D;
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[d] += 1;
  if (!decode3(fb, ix)) return;
  const int dm = a.dm;
  const double dt2 = 0.5 * a.dt, dt4 = a.dt / 4.0;
  int cl[3] = {ix[0], ix[1], ix[2]};
  cl[d] -= 1;
  double fl, fr;
  {
    const DV f = a.force.comp(d);
    const long fo = f.off(ix[0], ix[1], ix[2]);
    const long fst = f.stride(d);
    if (a.trace) {  // force traced along its own direction with the cell velocity (velpred.f90:783-793)
      const DV ufd = a.ufull.comp(d);
      const long uo = ufd.off(ix[0], ix[1], ix[2]);
      const LineBC b = make_linebc(dm, d, a.lo[d], a.hi[d], a.bclo[d][d], a.bchi[d][d]);
      double dummy;
      vel_cell_states(PPM, a.slope_order, dm == 3, f.p + fo - fst, fst, ix[d] - 1, b, ufd.p[uo - ufd.stride(d)],
                      a.dt, a.dx[d], a.rdx[d], a.rel_eps, fl, dummy);
      vel_cell_states(PPM, a.slope_order, dm == 3, f.p + fo, fst, ix[d], b, ufd.p[uo], a.dt, a.dx[d], a.rdx[d], a.rel_eps,
                      dummy, fr);
    } else {
      fl = f.p[fo - fst];
      fr = f.p[fo];
    }
  }
  const long fo = a.UL[d].off(ix[0], ix[1], ix[2]) + a.UL[d].cs * d;
  double ml, mr;
  if (dm == 2) {
    const int t = 1 - d;
    const long qc = a.UIMH[t].cs * d;
    ml = a.UL[d].p[fo] - tterm(VP_OVER(dt4, a.dx[t], a.rdx[t]), a.utrans[t], a.UIMH[t], qc, t, cl[0], cl[1], cl[2]) + dt2 * fl;
    mr = a.UR[d].p[fo] - tterm(VP_OVER(dt4, a.dx[t], a.rdx[t]), a.utrans[t], a.UIMH[t], qc, t, ix[0], ix[1], ix[2]) + dt2 * fr;
  } else {
    const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
    ml = a.UL[d].p[fo] - tterm(VP_OVER(dt4, a.dx[t1], a.rdx[t1]), a.utrans[t1], a.Q[d][t1], 0, t1, cl[0], cl[1], cl[2]) -
         tterm(VP_OVER(dt4, a.dx[t2], a.rdx[t2]), a.utrans[t2], a.Q[d][t2], 0, t2, cl[0], cl[1], cl[2]) + dt2 * fl;
    mr = a.UR[d].p[fo] - tterm(VP_OVER(dt4, a.dx[t1], a.rdx[t1]), a.utrans[t1], a.Q[d][t1], 0, t1, ix[0], ix[1], ix[2]) -
         tterm(VP_OVER(dt4, a.dx[t2], a.rdx[t2]), a.utrans[t2], a.Q[d][t2], 0, t2, ix[0], ix[1], ix[2]) + dt2 * fr;
  }
  const bool radial = a.spherical || (d == dm - 1);
  const double w0f = a.spherical ? a.w0mac[d](ix[0], ix[1], ix[2]) : (radial ? a.w0[ix[d]] : 0.0);
  double e = riemann_full(ml, mr, radial, w0f, a.rel_eps);
  if (ix[d] == a.lo[d]) {
    const int p = a.plo[d];
    if (p == MGPU_BC_INLET) e = a.utilde(cl[0], cl[1], cl[2], d);
    else if (wall3(p)) e = 0.0;
    else if (p == MGPU_BC_OUTLET) e = dmin2(mr, 0.0);
  }
  if (ix[d] == a.hi[d] + 1) {
    const int p = a.phi[d];
    if (p == MGPU_BC_INLET) e = a.utilde(ix[0], ix[1], ix[2], d);
    else if (wall3(p)) e = 0.0;
    else if (p == MGPU_BC_OUTLET) e = dmax2(ml, 0.0);
  }
  a.umac[d](ix[0], ix[1], ix[2]) = e;
}

// launches kern<D, PPM> for the runtime (d, ppm_type)
#define VP_LAUNCH(kern, d, ppm, grid, block, stream, args)                                            \
  do {                                                                                                \
    switch ((d)*3 + (ppm)) {                                                                          \
      case 0: MGPU_TIMED(TAG_VELPRED, (kern<0, 0><<<grid, block, 0, stream>>>(args))); break;         \
      case 1: MGPU_TIMED(TAG_VELPRED, (kern<0, 1><<<grid, block, 0, stream>>>(args))); break;         \
      case 2: MGPU_TIMED(TAG_VELPRED, (kern<0, 2><<<grid, block, 0, stream>>>(args))); break;         \
      case 3: MGPU_TIMED(TAG_VELPRED, (kern<1, 0><<<grid, block, 0, stream>>>(args))); break;         \
      case 4: MGPU_TIMED(TAG_VELPRED, (kern<1, 1><<<grid, block, 0, stream>>>(args))); break;         \
      case 5: MGPU_TIMED(TAG_VELPRED, (kern<1, 2><<<grid, block, 0, stream>>>(args))); break;         \
      case 6: MGPU_TIMED(TAG_VELPRED, (kern<2, 0><<<grid, block, 0, stream>>>(args))); break;         \
      case 7: MGPU_TIMED(TAG_VELPRED, (kern<2, 1><<<grid, block, 0, stream>>>(args))); break;         \
      case 8: MGPU_TIMED(TAG_VELPRED, (kern<2, 2><<<grid, block, 0, stream>>>(args))); break;         \
      default: throw Error("velpred: invalid ppm_type");                                              \
    }                                                                                                 \
  } while (0)

// the cell-centred kernels (k_mkutrans, k_vp_face): grid and block follow from the direction
#define VPC_CASE(kern, D, PPM, fb, stream, args) \
  MGPU_TIMED(TAG_VELPRED, (kern<D, PPM><<<vp_grid<D>(fb), vp_block<D>(), 0, stream>>>(args)))
#define VPC_LAUNCH(kern, d, ppm, fb, stream, args)                       \
  do {                                                                   \
    switch ((d)*3 + (ppm)) {                                             \
      case 0: VPC_CASE(kern, 0, 0, fb, stream, args); break;             \
      case 1: VPC_CASE(kern, 0, 1, fb, stream, args); break;             \
      case 2: VPC_CASE(kern, 0, 2, fb, stream, args); break;             \
      case 3: VPC_CASE(kern, 1, 0, fb, stream, args); break;             \
      case 4: VPC_CASE(kern, 1, 1, fb, stream, args); break;             \
      case 5: VPC_CASE(kern, 1, 2, fb, stream, args); break;             \
      case 6: VPC_CASE(kern, 2, 0, fb, stream, args); break;             \
      case 7: VPC_CASE(kern, 2, 1, fb, stream, args); break;             \
      case 8: VPC_CASE(kern, 2, 2, fb, stream, args); break;             \
      default: throw Error("velpred: invalid ppm_type");                 \
    }                                                                    \
  } while (0)

void check_phys(int bc, const char* who) {
  switch (bc) {
    case MGPU_BC_INLET: case MGPU_BC_OUTLET: case MGPU_BC_SYMMETRY: case MGPU_BC_SLIP_WALL: case MGPU_BC_NO_SLIP_WALL:
    case MGPU_BC_INTERIOR: case MGPU_BC_PERIODIC:
      return;
    default:
      throw Error(std::string(who) + ": invalid boundary type phys_bc");
  }
}

void fill_common(VpArgs& a, const mgpu_params& P, const DV& utilde, const DV& ufull, const double* w0_dev,
                 const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u, const char* who,
                 const DV* w0mac) {
  const int dm = P.dm;
  if (P.spherical && (!w0mac || dm != 3)) throw Error(std::string(who) + ": spherical geometry needs w0mac (3-D)");
  a.spherical = P.spherical != 0;
  for (int d = 0; d < 3; ++d) a.w0mac[d] = (a.spherical ? w0mac[d] : utilde);
  if (P.ppm_type == 2 && ng_u < 4) throw Error("Need 4 ghost cells for ppm_type=2");  // ppm.f90:1864-1866
  if (ng_u < 3) throw Error(std::string(who) + ": need at least 3 ghost cells");
  a.dm = dm;
  a.ppm_type = P.ppm_type;
  a.slope_order = P.slope_order;
  a.trace = false;
  a.dt = P.dt;
  a.rel_eps = P.rel_eps;
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = d < dm ? lo[d] : 0;
    a.hi[d] = d < dm ? hi[d] : 0;
    a.dx[d] = P.dx[d < dm ? d : 0];
    a.rdx[d] = 1.0 / a.dx[d];
    a.plo[d] = a.phi[d] = MGPU_BC_INTERIOR;
    for (int c = 0; c < 3; ++c) a.bclo[c][d] = a.bchi[c][d] = MGPU_BC_INTERIOR;
    if (d < dm) {
      a.plo[d] = phys_bc[d + dm * 0];
      a.phi[d] = phys_bc[d + dm * 1];
      check_phys(a.plo[d], who);
      check_phys(a.phi[d], who);
      for (int c = 0; c < dm; ++c) {
        a.bclo[c][d] = adv_bc[d + dm * (0 + 2 * c)];
        a.bchi[c][d] = adv_bc[d + dm * (1 + 2 * c)];
      }
    }
  }
  a.tb = grown(lo, hi, dm, 1);
  a.vb = grown(lo, hi, dm, 0);
  a.utilde = utilde;
  a.ufull = ufull;
  a.w0 = w0_dev;
  a.ufull_otf = false;
  a.nr = P.nr;
}

}  // namespace

#ifndef MGPU_VP_FAST
static int g_vp_fast = 1;
void velpred_set_fast(int on) { g_vp_fast = on; }
void mkutrans_dev_fast(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* utrans, const double* w0_dev,
                       const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u, const DV* w0mac);
void velpred_dev_fast(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* umac, const DV* utrans, const DV& force,
                      const double* w0_dev, const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u,
                      int ng_f, const DV* w0mac);
void mkutrans_dev_exact(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* utrans, const double* w0_dev,
                        const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u, const DV* w0mac);
void velpred_dev_exact(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* umac, const DV* utrans, const DV& force,
                       const double* w0_dev, const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u,
                       int ng_f, const DV* w0mac);
void velpred_premac_begin_fast(const mgpu_params& P, const DV& utilde, const DV& ufull, const DV* utpre,
                               const DV* utfull, const double* w0_dev, const int* lo, const int* hi, const int* adv_bc,
                               const int* phys_bc, int ng_u, VpArgs* st);
void velpred_premac_finish_fast(const mgpu_params& P, VpArgs* st, DV* umac, const DV& force, int ng_u, int ng_f);
void velpred_premac_begin_exact(const mgpu_params& P, const DV& utilde, const DV& ufull, const DV* utpre,
                                const DV* utfull, const double* w0_dev, const int* lo, const int* hi, const int* adv_bc,
                                const int* phys_bc, int ng_u, VpArgs* st);
void velpred_premac_finish_exact(const mgpu_params& P, VpArgs* st, DV* umac, const DV& force, int ng_u, int ng_f);
// the two reconstructions coincide for ppm_type 1 / 2 (for ppm_type 0 velpred_3d rounds its predictor differently,
// velpred.f90:812 against mkutrans.f90:543); a physical boundary hands the ghost rows of utrans to
// multifab_physbc_edgevel, which the face kernel cannot reproduce, and w0mac (spherical) has no transverse ghost rows
bool velpred_premac_fusable(const mgpu_params& P, const int* phys_bc) {
  if (P.spherical || P.ppm_type == 0 || P.dm < 2) return false;
  for (int i = 0; i < 2 * P.dm; ++i)
    if (phys_bc[i] != MGPU_BC_INTERIOR && phys_bc[i] != MGPU_BC_PERIODIC) return false;
  return true;
}
void velpred_premac_begin(const mgpu_params& P, const DV& utilde, const DV& ufull, const DV* utpre, const DV* utfull,
                          const double* w0_dev, const int* lo, const int* hi, const int* adv_bc, const int* phys_bc,
                          int ng_u, VpArgs* st) {
  if (g_vp_fast) velpred_premac_begin_fast(P, utilde, ufull, utpre, utfull, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, st);
  else velpred_premac_begin_exact(P, utilde, ufull, utpre, utfull, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, st);
}
void velpred_premac_finish(const mgpu_params& P, VpArgs* st, DV* umac, const DV& force, int ng_u, int ng_f) {
  if (g_vp_fast) velpred_premac_finish_fast(P, st, umac, force, ng_u, ng_f);
  else velpred_premac_finish_exact(P, st, umac, force, ng_u, ng_f);
}
void mkutrans_dev(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* utrans, const double* w0_dev,
                  const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u, const DV* w0mac) {
  if (g_vp_fast) mkutrans_dev_fast(P, utilde, ufull, utrans, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, w0mac);
  else mkutrans_dev_exact(P, utilde, ufull, utrans, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, w0mac);
}
void velpred_dev(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* umac, const DV* utrans, const DV& force,
                 const double* w0_dev, const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u,
                 int ng_f, const DV* w0mac) {
  if (g_vp_fast) velpred_dev_fast(P, utilde, ufull, umac, utrans, force, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, ng_f, w0mac);
  else velpred_dev_exact(P, utilde, ufull, umac, utrans, force, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, ng_f, w0mac);
}
#endif

void VP_FN(mkutrans_dev)(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* utrans, const double* w0_dev,
                         const int* lo, const int* hi, const int* adv_bc, const int* phys_bc, int ng_u, const DV* w0mac) {
  VpArgs a;
  fill_common(a, P, utilde, ufull, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, "mkutrans", w0mac);
  for (int d = 0; d < P.dm; ++d) a.utrans[d] = utrans[d];
  for (int d = 0; d < P.dm; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    VPC_LAUNCH(k_mkutrans, d, a.ppm_type, fb, ctx().stream, a);
  }
}

#ifndef MGPU_VP_FAST
size_t velpred_scratch(const mgpu_params& P, const int* lo, const int* hi) {
  Box3 tb = grown(lo, hi, P.dm, 1);
  const size_t narr = (size_t)3 * P.dm * P.dm + (P.dm == 3 ? 6 : 0);
  return narr * ((size_t)tb.npts() * sizeof(double) + 256) + 4096;
}
#endif

// temporaries of velpred on tb (arena) and the stages after the face states
static void vp_alloc_tmp(VpArgs& a) {
  const int dm = a.dm;
  const long nt = a.tb.npts();
  int z3[3] = {0, 0, 0};
  auto tmp = [&](int nc) { return make_view(arena_alloc((size_t)nt * nc), a.tb.lo, a.tb.hi, dm, 0, z3, nc); };
  for (int d = 0; d < dm; ++d) {
    a.UL[d] = tmp(dm);
    a.UR[d] = tmp(dm);
    a.UIMH[d] = tmp(dm);
  }
  if (dm == 3)
    for (int c = 0; c < 3; ++c)
      for (int d = 0; d < 3; ++d)
        if (c != d) a.Q[c][d] = tmp(1);
}
static void vp_set_force(VpArgs& a, const mgpu_params& P, const DV& force, int ng_u, int ng_f) {
  a.trace = (P.ppm_trace_forces == 1) && P.ppm_type != 0;
  if (a.trace && ng_f < ng_u) throw Error("velpred: ppm_trace_forces needs force with as many ghost cells as utilde");
  if (ng_f < 1) throw Error("velpred: force needs at least 1 ghost cell");
  a.force = force;
}
static void vp_trans_final(VpArgs& a) {
  const int dm = a.dm;
  cudaStream_t s = ctx().stream;
  if (dm == 3) MGPU_TIMED(TAG_VELPRED, (k_vp_trans<<<grid3(a.tb, 256), block3(a.tb, 256), 0, s>>>(a)));
  for (int d = 0; d < dm; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    VP_LAUNCH(k_vp_final, d, a.ppm_type, grid3(fb, 256), block3(fb, 256), s, a);
  }
}

void VP_FN(velpred_dev)(const mgpu_params& P, const DV& utilde, const DV& ufull, DV* umac, const DV* utrans,
                        const DV& force, const double* w0_dev, const int* lo, const int* hi, const int* adv_bc,
                        const int* phys_bc, int ng_u, int ng_f, const DV* w0mac) {
  VpArgs a;
  fill_common(a, P, utilde, ufull, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, "velpred", w0mac);
  const int dm = P.dm;
  vp_set_force(a, P, force, ng_u, ng_f);
  for (int d = 0; d < dm; ++d) {
    a.utrans[d] = utrans[d];
    a.umac[d] = umac[d];
  }
  vp_alloc_tmp(a);
  cudaStream_t s = ctx().stream;
  for (int d = 0; d < dm; ++d) {
    Box3 fb = a.tb;
    fb.lo[d] = a.lo[d];
    VPC_LAUNCH(k_vp_face, d, a.ppm_type, fb, s, a);
  }
  vp_trans_final(a);
}

// advance_premac without k_mkutrans (see vp_face_body<.., UT = true>): `begin` leaves utrans (utpre, what mk_vel_force
// reads) and utrans + w0 (utfull; the same fab as utpre in the directions w0 does not act on) on the whole face box
// and the face states in the arena; `finish` runs the transverse and final stages once the force exists.
void VP_FN(velpred_premac_begin)(const mgpu_params& P, const DV& utilde, const DV& ufull, const DV* utpre,
                                 const DV* utfull, const double* w0_dev, const int* lo, const int* hi,
                                 const int* adv_bc, const int* phys_bc, int ng_u, VpArgs* st) {
  VpArgs& a = *st;
  fill_common(a, P, utilde, ufull, w0_dev, lo, hi, adv_bc, phys_bc, ng_u, "advance_premac", nullptr);
  a.ufull_otf = ufull.p == nullptr;  // the caller did not build ufull (only traced forces read it after the faces)
  const int dm = P.dm;
  for (int d = 0; d < dm; ++d) {
    a.utpre[d] = utpre[d];
    a.utrans[d] = utfull[d];
  }
  vp_alloc_tmp(a);
  cudaStream_t s = ctx().stream;
  for (int d = 0; d < dm; ++d) {
    Box3 fb = a.tb;
    fb.lo[d] = a.lo[d];
    VPC_LAUNCH(k_vp_face_ut, d, a.ppm_type, fb, s, a);
  }
}
void VP_FN(velpred_premac_finish)(const mgpu_params& P, VpArgs* st, DV* umac, const DV& force, int ng_u, int ng_f) {
  VpArgs& a = *st;
  vp_set_force(a, P, force, ng_u, ng_f);
  for (int d = 0; d < P.dm; ++d) a.umac[d] = umac[d];
  vp_trans_final(a);
}

}  // namespace mgpu
