// TEST INFRASTRUCTURE ONLY -- CPU oracle for the MAESTRO advective hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may use oracle/.
//
// Fortran-style array helper: lower bounds per dimension, x fastest, component slowest, so that
// the restated loops can keep the reference's index ranges verbatim.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/maestro_b200.h"

namespace mo {

struct Arr {
  double* p = nullptr;
  int lo[3] = {0, 0, 0};
  int n[3] = {1, 1, 1};
  int nc = 1;
  std::vector<double> own;

  Arr() {}
  // owning, zero-initialised: (l0:h0, l1:h1, l2:h2, nc)
  Arr(int l0, int h0, int l1, int h1, int l2, int h2, int ncomp = 1) { alloc(l0, h0, l1, h1, l2, h2, ncomp); }
  void alloc(int l0, int h0, int l1, int h1, int l2, int h2, int ncomp = 1) {
    lo[0] = l0; lo[1] = l1; lo[2] = l2;
    n[0] = h0 - l0 + 1; n[1] = h1 - l1 + 1; n[2] = h2 - l2 + 1;
    nc = ncomp;
    own.assign((size_t)n[0] * n[1] * n[2] * nc, 0.0);
    p = own.data();
  }
  // non-owning view of a fab as described by the C ABI
  static Arr view(const mgpu_fab& f, int dm) {
    Arr a;
    for (int d = 0; d < 3; ++d) {
      if (d < dm) {
        a.lo[d] = f.lo[d] - f.ng;
        a.n[d] = f.hi[d] - f.lo[d] + 1 + 2 * f.ng + f.nodal[d];
      } else {
        a.lo[d] = 0;
        a.n[d] = 1;
      }
    }
    a.nc = f.nc;
    a.p = f.ptr;
    return a;
  }
  // view of one component (0-based c) as a single-component array
  Arr comp(int c) const {
    Arr a;
    for (int d = 0; d < 3; ++d) { a.lo[d] = lo[d]; a.n[d] = n[d]; }
    a.nc = 1;
    a.p = p + (size_t)c * n[0] * n[1] * n[2];
    return a;
  }
  inline size_t idx(int i, int j, int k, int c = 0) const {
#ifdef MO_BOUNDS
    if (i < lo[0] || i >= lo[0] + n[0] || j < lo[1] || j >= lo[1] + n[1] || k < lo[2] ||
        k >= lo[2] + n[2] || c < 0 || c >= nc) {
      fprintf(stderr, "oracle: index (%d,%d,%d,%d) outside [%d:%d,%d:%d,%d:%d,%d]\n", i, j, k, c, lo[0],
              lo[0] + n[0] - 1, lo[1], lo[1] + n[1] - 1, lo[2], lo[2] + n[2] - 1, nc);
      abort();
    }
#endif
    return (size_t)(i - lo[0]) + (size_t)n[0] * ((size_t)(j - lo[1]) + (size_t)n[1] * ((size_t)(k - lo[2]) + (size_t)n[2] * c));
  }
  inline double& operator()(int i, int j, int k, int c = 0) const { return p[idx(i, j, k, c)]; }
  // access with an offset `o` along direction d
  inline double& at(int i, int j, int k, int d, int o) const {
    return p[idx(i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0))];
  }
  size_t size() const { return (size_t)n[0] * n[1] * n[2] * nc; }
  void fill(double v) { for (size_t q = 0; q < size(); ++q) p[q] = v; }
};

// adv_bc(d,side,bccomp) with 1-based d, side, bccomp as in the Fortran
inline int advbc(const int* adv_bc, int dm, int d, int side, int bccomp) {
  return adv_bc[(d - 1) + dm * ((side - 1) + 2 * (bccomp - 1))];
}

inline double sign1(double x) { return std::copysign(1.0, x); }  // Fortran sign(ONE,x) (IEEE sign bit, as gfortran)
inline double dmin(double a, double b) { return a < b ? a : b; }
inline double dmax(double a, double b) { return a > b ? a : b; }
inline double dabs(double a) { return a < 0 ? -a : a; }

void fail(const char* msg);

}  // namespace mo
