// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
//
// Ghost fill for a single box that covers the whole problem domain:
//   * periodic wrap = what FBoxLib multifab_fill_boundary does for a periodic layout (pure copies,
//     corners included).  FBoxLib is not in the reference tree: parity for this step is pinned only
//     by its copy semantics ("parity unpinned").
//   * physical boundaries = Source/multifab_physbc.f90: physbc_2d :150, physbc_3d :329.
#include "mo_kernels.h"

namespace mo {

static void periodic_wrap(Arr& a, const int* lo, const int* hi, int dm, int ng, int d, int nodal) {
  // cell-centred: ghost(lo-g) = valid(hi+1-g), ghost(hi+g) = valid(lo-1+g), g = 1..ng
  // face-centred in d (nodal): faces lo..hi are unique, face hi+1 == face lo;
  //   ghost(lo-g) = face(hi+1-g), ghost(hi+1+g) = face(lo+g); valid faces are never overwritten
  const int n = hi[d] - lo[d] + 1;
  Box b;
  for (int q = 0; q < 3; ++q) { b.lo[q] = a.lo[q]; b.hi[q] = a.lo[q] + a.n[q] - 1; }
  b.lo[d] = b.hi[d] = 0;
  for (int k = b.lo[2]; k <= b.hi[2]; ++k)
    for (int j = b.lo[1]; j <= b.hi[1]; ++j)
      for (int i = b.lo[0]; i <= b.hi[0]; ++i) {
        for (int g = 1; g <= ng; ++g) {
          a.at(i, j, k, d, lo[d] - g) = a.at(i, j, k, d, lo[d] - g + n);
        }
        for (int g = 1; g <= ng; ++g) {
          a.at(i, j, k, d, hi[d] + nodal + g) = a.at(i, j, k, d, hi[d] + nodal + g - n);
        }
      }
  (void)dm;
}

void fill_boundary_box(const mgpu_params& P, Arr& sfull, const int* lo, const int* hi, int ng, int scomp,
                       int bccomp, int ncomp, const int* adv_bc, const int* pmask) {
  const int dm = P.dm;
  if (ng == 0) return;
  for (int n = 0; n < ncomp; ++n) {
    Arr s = sfull.comp(scomp - 1 + n);
    const int bcc = bccomp + n;
    for (int d = 0; d < dm; ++d)
      if (pmask[d]) periodic_wrap(s, lo, hi, dm, ng, d, 0);
    // multifab_physbc: directions in order x,y,z; lo side then hi side
    int bc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int d = 0; d < dm; ++d) {
      bc[d][0] = advbc(adv_bc, dm, d + 1, 1, bcc);
      bc[d][1] = advbc(adv_bc, dm, d + 1, 2, bcc);
    }
    for (int d = 0; d < dm; ++d) {
      for (int side = 0; side < 2; ++side) {
        const int b = bc[d][side];
        if (b == MGPU_BC_INTERIOR || b == MGPU_BC_PERIODIC) continue;
        // transverse range: dims before d use the full ghost range, dims after d use the ghost range
        // only when that side is INTERIOR (corner ghosts are skipped), EXT_DIR always uses everything
        Box tb;
        for (int t = 0; t < 3; ++t) {
          if (t >= dm) { tb.lo[t] = tb.hi[t] = 0; continue; }
          if (t == d) { tb.lo[t] = tb.hi[t] = 0; continue; }
          int glo = ng, ghi = ng;
          if (t > d && b != MGPU_BC_EXT_DIR) {
            glo = (bc[t][0] == MGPU_BC_INTERIOR) ? ng : 0;
            ghi = (bc[t][1] == MGPU_BC_INTERIOR) ? ng : 0;
          }
          tb.lo[t] = lo[t] - glo;
          tb.hi[t] = hi[t] + ghi;
        }
        const int e = (side == 0) ? lo[d] : hi[d];  // first valid cell next to the wall
        const int sg = (side == 0) ? -1 : 1;        // outward direction
        for (int k = tb.lo[2]; k <= tb.hi[2]; ++k)
          for (int j = tb.lo[1]; j <= tb.hi[1]; ++j)
            for (int i = tb.lo[0]; i <= tb.hi[0]; ++i) {
              auto S = [&](int q) -> double& { return s.at(i, j, k, d, q); };  // absolute index q in d
              if (b == MGPU_BC_EXT_DIR) {
                for (int g = 1; g <= ng; ++g) S(e + sg * g) = 0.0;
              } else if (b == MGPU_BC_FOEXTRAP) {
                for (int g = 1; g <= ng; ++g) S(e + sg * g) = S(e);
              } else if (b == MGPU_BC_HOEXTRAP) {
                S(e + sg) = (15.0 * S(e) - 10.0 * S(e - sg) + 3.0 * S(e - 2 * sg)) * 0.125;
                for (int g = 2; g <= ng; ++g) S(e + sg * g) = S(e + sg);
              } else if (b == MGPU_BC_REFLECT_EVEN) {
                for (int g = 1; g <= ng; ++g) S(e + sg * g) = S(e - sg * (g - 1));
              } else if (b == MGPU_BC_REFLECT_ODD) {
                for (int g = 1; g <= ng; ++g) S(e + sg * g) = -S(e - sg * (g - 1));
              } else {
                fail("physbc: bc not yet supported");
              }
            }
      }
    }
  }
}

void fill_boundary_face(const mgpu_params& P, Arr& u, const int* lo, const int* hi, int ng, int dir,
                        const int* pmask) {
  const int dm = P.dm;
  for (int d = 0; d < dm; ++d)
    if (pmask[d]) periodic_wrap(u, lo, hi, dm, ng, d, d == dir ? 1 : 0);
}

}  // namespace mo
