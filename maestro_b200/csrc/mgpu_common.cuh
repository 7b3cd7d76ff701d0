// Common device/host helpers of the B200-native MAESTRO advective path.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/maestro_b200.h"

namespace mgpu {

// ---------------------------------------------------------------------------------------------
// device view of one fab: Fortran order (x fastest, component slowest), lower bounds kept so the
// kernels index with the reference's (i,j,k,comp) coordinates.
// ---------------------------------------------------------------------------------------------
struct DV {
  double* p;
  int lo[3];
  int n[3];
  long cs;  // component stride = n0*n1*n2
  int nc;
  __host__ __device__ inline long off(int i, int j, int k) const {
    return (long)(i - lo[0]) + (long)n[0] * ((long)(j - lo[1]) + (long)n[1] * (long)(k - lo[2]));
  }
  // 32-bit element index for fabs below 2^31 elements (the caller checks): half the integer work of off()
  __host__ __device__ inline int off32(int i, int j, int k) const { return (i - lo[0]) + n[0] * ((j - lo[1]) + n[1] * (k - lo[2])); }
  __host__ __device__ inline long stride(int d) const { return d == 0 ? 1 : (d == 1 ? (long)n[0] : (long)n[0] * n[1]); }
  __device__ inline double& operator()(int i, int j, int k) const { return p[off(i, j, k)]; }
  __device__ inline double& operator()(int i, int j, int k, int c) const { return p[off(i, j, k) + cs * c]; }
  __host__ __device__ inline DV comp(int c) const {
    DV v = *this;
    v.p = p + cs * c;
    v.nc = 1;
    return v;
  }
  __host__ __device__ inline long size() const { return cs * nc; }
};

inline DV make_view(double* p, const int* lo, const int* hi, int dm, int ng, const int* nodal, int nc) {
  DV v;
  for (int d = 0; d < 3; ++d) {
    if (d < dm) {
      v.lo[d] = lo[d] - ng;
      v.n[d] = hi[d] - lo[d] + 1 + 2 * ng + (nodal ? nodal[d] : 0);
    } else {
      v.lo[d] = 0;
      v.n[d] = 1;
    }
  }
  v.cs = (long)v.n[0] * v.n[1] * v.n[2];
  v.nc = nc;
  v.p = p;
  return v;
}
inline DV make_view(const mgpu_fab& f, int dm, double* p = nullptr) {
  return make_view(p ? p : f.ptr, f.lo, f.hi, dm, f.ng, f.nodal, f.nc);
}

struct Box3 {
  int lo[3];
  int hi[3];
  __host__ __device__ inline long npts() const {
    return (long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
  }
};
inline Box3 grown(const int* lo, const int* hi, int dm, int g) {
  Box3 b;
  for (int d = 0; d < 3; ++d) {
    if (d < dm) {
      b.lo[d] = lo[d] - g;
      b.hi[d] = hi[d] + g;
    } else {
      b.lo[d] = b.hi[d] = 0;
    }
  }
  return b;
}

// ---------------------------------------------------------------------------------------------
// error handling: every failure becomes a C++ exception that the C ABI turns into a return code
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define MGPU_CUDA(call)                                                                             \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      throw mgpu::Error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + \
                        ":" + std::to_string(__LINE__));                                            \
  } while (0)

// ---------------------------------------------------------------------------------------------
// library context: one GPU per process, one stream, a grow-only scratch arena for temporaries
// (the reference allocates/deallocates ~45 box-sized temporaries per make_edge_scal call).
// ---------------------------------------------------------------------------------------------
struct Context {
  bool initialised = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long launches = 0;
  // scratch arena
  char* arena = nullptr;
  size_t arena_bytes = 0;
  size_t arena_used = 0;
  // small device buffer for base-state arrays / bc tables
  std::vector<void*> to_free;
};
Context& ctx();
void require_init();
// bump-allocate from the arena (aligned to 256 B); arena_reset() at the start of every entry point
double* arena_alloc(size_t ndoubles);
void arena_reset();
void arena_reserve(size_t bytes);
size_t arena_mark();
void arena_release(size_t mark);
inline void count_launch(int n = 1) { ctx().launches += n; }

#define MGPU_LAUNCH_CHECK()             \
  do {                                  \
    MGPU_CUDA(cudaGetLastError());      \
    mgpu::count_launch();               \
  } while (0)

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
// line): MGPU_TIMED(tag, launch-statement) brackets one launch when profiling is switched on.
enum KernelTag {
  TAG_EDGE_CELL = 0, TAG_EDGE_SIMH, TAG_EDGE_TRANS, TAG_EDGE_FINAL, TAG_FLUX, TAG_UPDATE, TAG_FILL, TAG_GLUE,
  TAG_FUSED_EDGE, TAG_VELPRED, TAG_BDS, TAG_HALO, TAG_COUNT
};
void prof_begin(int tag);
void prof_end(int tag);
#define MGPU_TIMED(tag, stmt) \
  do {                        \
    mgpu::prof_begin(tag);    \
    stmt;                     \
    MGPU_LAUNCH_CHECK();      \
    mgpu::prof_end(tag);      \
  } while (0)

#ifdef __CUDACC__
// thread id -> (i,j,k) of a box, x fastest
__device__ __forceinline__ bool decode(const Box3& b, long t, int* ix) {
  const int nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1;
  if (t >= b.npts()) return false;
  ix[0] = b.lo[0] + (int)(t % nx);
  long r = t / nx;
  ix[1] = b.lo[1] + (int)(r % ny);
  ix[2] = b.lo[2] + (int)(r / ny);
  return true;
}
#define MGPU_TID ((long)blockIdx.x * blockDim.x + threadIdx.x)
// linear thread id -> (i,j,k) with 32-bit arithmetic, for thin boxes (ghost shells) where a row per block
// would leave most lanes idle; requires npts < 2^32
__device__ __forceinline__ bool decode32(const Box3& b, int* ix) {
  const unsigned nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1, nz = b.hi[2] - b.lo[2] + 1;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx * ny * nz) return false;
  const unsigned r = t / nx;
  ix[0] = b.lo[0] + (int)(t - r * nx);
  const unsigned q = r / ny;
  ix[1] = b.lo[1] + (int)(r - q * ny);
  ix[2] = b.lo[2] + (int)q;
  return true;
}
// (i,j,k) of a box from a 3-D launch (grid3/block3 below): x along the threads of a block, one (j,k) row
// segment per block -- no integer division per thread, every warp reads/writes one contiguous x line
__device__ __forceinline__ bool decode3(const Box3& b, int* ix) {
  ix[0] = b.lo[0] + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  ix[1] = b.lo[1] + (int)blockIdx.y;
  ix[2] = b.lo[2] + (int)blockIdx.z;
  return ix[0] <= b.hi[0];
}
#endif
inline int block3(const Box3& b, int bs_max) {
  const int nx = b.hi[0] - b.lo[0] + 1;
  int bs = 32;
  while (bs < nx && bs < bs_max) bs *= 2;
  return bs;
}
inline dim3 grid3(const Box3& b, int bs_max) {
  const int nx = b.hi[0] - b.lo[0] + 1, bs = block3(b, bs_max);
  return dim3((unsigned)((nx + bs - 1) / bs), (unsigned)(b.hi[1] - b.lo[1] + 1), (unsigned)(b.hi[2] - b.lo[2] + 1));
}
inline unsigned nblocks(long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// upload a small host array (base state, BC table) into the arena; returns device pointer
template <class T>
T* upload_small(const T* h, size_t n);

}  // namespace mgpu
