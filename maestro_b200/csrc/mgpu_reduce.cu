// Reductions next to the advective path, planar geometry:
//   estdt_box_dev      the per-box time-step limits of Source/estdt.f90:348 (estdt_2d) and :467 (estdt_3d_cart)
//   plane_sums_dev     the per-plane sums of etarhoflux, Source/make_eta.f90:176 (sum_etarho_2d) / :213 (_3d)
// Max / min reductions are exact whatever the order, so estdt is bit-identical to the reference; the plane sums are
// floating-point sums in a different order (tolerance 1e-12 relative).
#include "mgpu_reduce.cuh"

namespace mgpu {
namespace {

// slots of the estdt partial results (one set per CTA): maxima first, minima after
enum { E_SPD0 = 0, E_SPD1, E_SPD2, E_F0, E_F1, E_F2, E_DT_DIVU, E_DT_DSDT, E_N };

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct EstdtArgs {
  int dm, nr, rho;  // rho: 0-based density component of s
  Box3 vb;          // valid box
  DV u, s, force, divU, dSdt;
  const double *w0, *p0, *gamma1bar;
  double dx_r;  // dx of the radial (last) direction
  double rho_min;
  double* part;  // E_N doubles per CTA
  // spherical (estdt_3d_sphr, estdt.f90:620): w0 on the faces of every direction, grad(p0)/(gamma1bar p0) as a
  // Cartesian vector field on the valid cells (put_1d_array_on_cart of gp0, :741)
  DV w0mac[3], gp0c;
};

// One pass over the valid cells; every thread keeps the eight running extrema of estdt_2d / estdt_3d_cart.
// The expressions are the reference's (estdt.f90:503-515 speeds, :538-552 forces, :566-585 divU, :600-606 dS/dt).
template <int DM, bool SPHR>
__global__ void __launch_bounds__(256) k_estdt(EstdtArgs a) {
  const int nx = a.vb.hi[0] - a.vb.lo[0] + 1, ny = a.vb.hi[1] - a.vb.lo[1] + 1;
  const long npts = a.vb.npts();
  double m[E_N];
#pragma unroll
  for (int q = 0; q < E_DT_DIVU; ++q) m[q] = 0.0;
  m[E_DT_DIVU] = m[E_DT_DSDT] = INFINITY;
  const int r = DM - 1;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < npts; t += (long)gridDim.x * blockDim.x) {
    const int i = a.vb.lo[0] + (int)(t % nx);
    const int j = a.vb.lo[1] + (int)((t / nx) % ny);
    const int k = a.vb.lo[2] + (int)(t / ((long)nx * ny));
    const int kr = (DM == 3) ? k : j;  // radial index
    const long ou = a.u.off(i, j, k), of = a.force.off(i, j, k);
    const double ur = a.u.p[ou + a.u.cs * r];
#pragma unroll
    for (int d = 0; d < DM; ++d) {
      double v = a.u.p[ou + a.u.cs * d];
      if constexpr (SPHR) {  // :664-680
        const long ow = a.w0mac[d].off(i, j, k);
        v = v + 0.5 * (a.w0mac[d].p[ow] + a.w0mac[d].p[ow + a.w0mac[d].stride(d)]);
      } else if (d == r) {
        v = v + 0.5 * (a.w0[kr] + a.w0[kr + 1]);
      }
      m[E_SPD0 + d] = fmax(m[E_SPD0 + d], fabs(v));
      m[E_F0 + d] = fmax(m[E_F0 + d], fabs(a.force.p[of + a.force.cs * d]));
    }
    const double rho = a.s(i, j, k, a.rho);
    const double dU = a.divU(i, j, k);
    double denom;
    if constexpr (SPHR) {  // :747-751
      const long og = a.gp0c.off(i, j, k);
      const double gp_dot_u = a.u.p[ou] * a.gp0c.p[og] + a.u.p[ou + a.u.cs] * a.gp0c.p[og + a.gp0c.cs] +
                              a.u.p[ou + 2 * a.u.cs] * a.gp0c.p[og + 2 * a.gp0c.cs];
      denom = dU - gp_dot_u;
    } else {
      double gradp0;
      if (kr == 0) gradp0 = (a.p0[kr + 1] - a.p0[kr]) / a.dx_r;
      else if (kr == a.nr - 1) gradp0 = (a.p0[kr] - a.p0[kr - 1]) / a.dx_r;
      else gradp0 = 0.5 * (a.p0[kr + 1] - a.p0[kr - 1]) / a.dx_r;
      denom = dU - ur * gradp0 / (a.gamma1bar[kr] * a.p0[kr]);
    }
    if (denom > 0.0) m[E_DT_DIVU] = fmin(m[E_DT_DIVU], 0.4 * (1.0 - a.rho_min / rho) / denom);
    const double dS = a.dSdt(i, j, k);
    if (dS > 1.e-20) {
      const double qa = 0.5 * rho * dS, qb = rho * dU, qc = a.rho_min - rho;
      m[E_DT_DSDT] = fmin(m[E_DT_DSDT], 0.4 * 2.0 * qc / (-qb - sqrt(qb * qb - 4.0 * qa * qc)));
    }
  }
  __shared__ double sh[8][E_N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < E_N; ++q) {
    const double v = q < E_DT_DIVU ? warp_max(m[q]) : warp_min(m[q]);
    if (lane == 0) sh[w][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < E_N) {
    const int q = threadIdx.x;
    double v = sh[0][q];
    for (int ww = 1; ww < 8; ++ww) v = q < E_DT_DIVU ? fmax(v, sh[ww][q]) : fmin(v, sh[ww][q]);
    a.part[(long)blockIdx.x * E_N + q] = v;
  }
}

// sum of one plane (fixed slab-direction index) of a single-component fab over the valid transverse cells; one
// CTA per plane, fixed summation tree (deterministic from run to run)
__global__ void __launch_bounds__(256) k_plane_sums(DV f, Box3 vb, int dm, int k0, double* out) {
  const int kk = k0 + blockIdx.x;
  const int nx = vb.hi[0] - vb.lo[0] + 1;
  const int ny = (dm == 3) ? vb.hi[1] - vb.lo[1] + 1 : 1;
  double acc = 0.0;
  for (int t = threadIdx.x; t < nx * ny; t += blockDim.x) {
    const int i = vb.lo[0] + t % nx, j2 = t / nx;
    acc += (dm == 3) ? f(i, vb.lo[1] + j2, kk) : f(i, kk, 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[w];
    out[blockIdx.x] = v;
  }
}

}  // namespace

void estdt_box_dev(const mgpu_params& P, const DV& u, const DV& s, const DV& force, const DV& divU, const DV& dSdt,
                   const double* w0, const double* w0_h, const double* p0, const double* gamma1bar, const int* lo,
                   const int* hi, double rho_min, double cfl, double* dt, double* umax, const DV* w0mac,
                   const DV* gp0_cart, double dr, int nr_fine) {
  Context& cx = ctx();
  const int dm = P.dm, r = dm - 1;
  EstdtArgs a;
  a.dm = dm;
  a.nr = P.nr;
  a.rho = P.rho_comp - 1;
  a.vb = grown(lo, hi, dm, 0);
  a.u = u; a.s = s; a.force = force; a.divU = divU; a.dSdt = dSdt;
  a.w0 = w0; a.p0 = p0; a.gamma1bar = gamma1bar;
  a.dx_r = P.dx[r];
  a.rho_min = rho_min;
  const unsigned nb = std::min<unsigned>(nblocks(a.vb.npts(), 256), 148u * 8u);
  a.part = arena_alloc((size_t)nb * E_N);
  const bool sphr = w0mac != nullptr;
  if (sphr) {
    for (int d = 0; d < 3; ++d) a.w0mac[d] = w0mac[d];
    a.gp0c = *gp0_cart;
    k_estdt<3, true><<<nb, 256, 0, cx.stream>>>(a);
  } else if (dm == 3) {
    k_estdt<3, false><<<nb, 256, 0, cx.stream>>>(a);
  } else {
    k_estdt<2, false><<<nb, 256, 0, cx.stream>>>(a);
  }
  MGPU_LAUNCH_CHECK();
  std::vector<double> part((size_t)nb * E_N);
  MGPU_CUDA(cudaMemcpyAsync(part.data(), a.part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
  MGPU_CUDA(cudaStreamSynchronize(cx.stream));
  double m[E_N];
  for (int q = 0; q < E_N; ++q) m[q] = part[q];
  for (unsigned b = 1; b < nb; ++b)
    for (int q = 0; q < E_N; ++q)
      m[q] = q < E_DT_DIVU ? std::max(m[q], part[(size_t)b * E_N + q]) : std::min(m[q], part[(size_t)b * E_N + q]);
  // the scalar tail of estdt_2d / estdt_3d_cart, statement by statement
  const double eps = 1.0e-8;
  double spdr = 0.0;
  if (sphr) {
    for (int k = 0; k <= nr_fine; ++k) spdr = std::max(spdr, std::fabs(w0_h[k]));  // :682-684
  } else {
    for (int k = lo[r]; k <= hi[r]; ++k) spdr = std::max(spdr, std::fabs(w0_h[k]));  // :517-519
  }
  double um = 0.0;
  for (int d = 0; d < dm; ++d) um = std::max(um, m[E_SPD0 + d]);
  *umax = std::max(um, spdr);  // :521
  double t = *dt;
  for (int d = 0; d < dm; ++d)
    if (m[E_SPD0 + d] > eps) t = std::min(t, P.dx[d] / m[E_SPD0 + d]);  // :523-525
  if (spdr > eps) t = std::min(t, (sphr ? dr : P.dx[r]) / spdr);       // :526 / :691
  t = t * cfl;                                                          // :528
  for (int d = 0; d < dm; ++d)
    if (m[E_F0 + d] > eps) t = std::min(t, std::sqrt(2.0 * P.dx[d] / m[E_F0 + d]));  // :554-561
  t = std::min(t, m[E_DT_DIVU]);
  t = std::min(t, m[E_DT_DSDT]);
  *dt = t;
}

void plane_sums_dev(const mgpu_params& P, const DV& f, const int* lo, const int* hi, int k0, int k1, double* sums_h) {
  if (k1 < k0) return;
  Context& cx = ctx();
  const int np = k1 - k0 + 1;
  double* d = arena_alloc((size_t)np);
  k_plane_sums<<<np, 256, 0, cx.stream>>>(f, grown(lo, hi, P.dm, 0), P.dm, k0, d);
  MGPU_LAUNCH_CHECK();
  MGPU_CUDA(cudaMemcpyAsync(sums_h, d, (size_t)np * sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
  MGPU_CUDA(cudaStreamSynchronize(cx.stream));
}

}  // namespace mgpu
