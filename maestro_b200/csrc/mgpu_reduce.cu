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

// sum_phi_3d_sphr (average.f90:564-618, no mask: one level): every cell adds its value and a count to the radial bin its
// centre maps into.  Floating-point atomics: the order of the additions inside a bin differs from the reference's loop
// (and from run to run), parity is 1e-12 relative like the plane sums.
__global__ void __launch_bounds__(256) k_sum_phi_sphr(DV phi, Box3 vb, double cx, double cy, double cz, double plx,
                                                      double ply, double plz, double dx0, double dx1, double dx2,
                                                      const double* radii, int nr_irreg, double* phisum,
                                                      unsigned long long* ncell) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double z = plz + ((double)k + 0.5) * dx2 - cz;
  const double y = ply + ((double)j + 0.5) * dx1 - cy;
  const double x = plx + ((double)i + 0.5) * dx0 - cx;
  const double radius = sqrt(x * x + y * y + z * z);
  const double q = radius / dx0;
  int index = (int)((q * q - 0.75) / 2.0);
  if (index < nr_irreg) {
    if (fabs(radius - radii[index]) > fabs(radius - radii[index + 1])) index = index + 1;
  }
  atomicAdd(&phisum[index], phi(i, j, k));
  atomicAdd(&ncell[index], 1ull);
}

// construct_eta_cart (make_eta.f90:345-408): [rho' (U . e_r)] at the half time on the valid cells
__global__ void k_eta_cart(DV eta, Box3 vb, DV rho_old, DV rho_new, DV um, DV vm, DV wm, DV wx, DV wy, DV wz, DV normal,
                           DV nph) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const double U_dot_er = 0.5 * (um(i, j, k) + um(i + 1, j, k) + wx(i, j, k) + wx(i + 1, j, k)) * normal(i, j, k, 0) +
                          0.5 * (vm(i, j, k) + vm(i, j + 1, k) + wy(i, j, k) + wy(i, j + 1, k)) * normal(i, j, k, 1) +
                          0.5 * (wm(i, j, k) + wm(i, j, k + 1) + wz(i, j, k) + wz(i, j, k + 1)) * normal(i, j, k, 2);
  eta(i, j, k) = (0.5 * (rho_old(i, j, k) + rho_new(i, j, k)) - nph(i, j, k)) * U_dot_er;
}

// quad_interp, average.f90:386-401
double avg_quad_interp(double x, double x0, double x1, double x2, double y0, double y1, double y2, bool limit) {
  double y = y0 + (y1 - y0) / (x1 - x0) * (x - x0) +
             ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0)) / (x2 - x0) * (x - x0) * (x - x1);
  if (limit) {
    const double hi = std::max(std::max(y0, y1), y2), lo = std::min(std::min(y0, y1), y2);
    if (y > hi) y = hi;
    if (y < lo) y = lo;
  }
  return y;
}

}  // namespace

void sum_phi_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const DV& phi1, const int* lo, const int* hi,
                      const double* radii_dev, int nr_irreg, double* phisum_dev, unsigned long long* ncell_dev) {
  const Box3 vb = grown(lo, hi, 3, 0);
  MGPU_TIMED(TAG_GLUE, (k_sum_phi_sphr<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
                           phi1, vb, g.center[0], g.center[1], g.center[2], g.prob_lo[0], g.prob_lo[1], g.prob_lo[2],
                           P.dx[0], P.dx[1], P.dx[2], radii_dev, nr_irreg, phisum_dev, ncell_dev)));
}

void eta_cart_dev(const DV& eta, const DV& rho_old, const DV& rho_new, const DV* umac, const DV* w0mac, const DV& normal,
                  const DV& rho0_nph_cart, const int* lo, const int* hi) {
  const Box3 vb = grown(lo, hi, 3, 0);
  MGPU_TIMED(TAG_GLUE, (k_eta_cart<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
                           eta, vb, rho_old, rho_new, umac[0], umac[1], umac[2], w0mac[0], w0mac[1], w0mac[2], normal,
                           rho0_nph_cart)));
}

// the part of average() after the sums for spherical == 1 and one level (average.f90:204-362); the three arrays are
// indexed -1.. in the reference: element [r + 1] here
void average_sphr_tail(double dr, int nr_fine, int nr_irreg, int drdxfac, std::vector<double>& phisum,
                       std::vector<long>& ncell, std::vector<double>& radii, double* phibar) {
  double* PS = phisum.data() + 1;
  long* NC = ncell.data() + 1;
  double* RD = radii.data() + 1;
  for (int r = 0; r <= nr_irreg; ++r)
    if (NC[r] != 0) PS[r] = PS[r] / (double)NC[r];      // :204-210
  PS[-1] = (11.0 / 8.0) * PS[0] - (3.0 / 8.0) * PS[1];  // :213-215
  RD[-1] = 0.0;
  NC[-1] = 1;
  int max_rcoord = nr_irreg;  // :286-309: drop the radii no cell maps into (one level: which_lev = 1 everywhere)
  for (int r = 0, j = 0; r <= nr_irreg; ++r) {
    while (j <= nr_irreg && NC[j] == 0) ++j;
    if (j > nr_irreg) {
      for (int q = r; q <= nr_irreg; ++q) PS[q] = 1.e99;
      for (int q = r; q <= nr_irreg + 1; ++q) RD[q] = 1.e99;
      max_rcoord = r - 1;
      break;
    }
    PS[r] = PS[j];
    RD[r] = RD[j];
    NC[r] = NC[j];
    ++j;
    if (j > nr_irreg) {
      max_rcoord = r;
      break;
    }
  }
  int sc = 0;  // :312-352
  for (int r = 0; r < nr_fine; ++r) {
    const double radius = ((double)r + 0.5) * dr;
    for (int j = sc; j <= max_rcoord; ++j)
      if (std::fabs(radius - RD[j]) < std::fabs(radius - RD[j + 1])) {
        sc = j;
        break;
      }
    sc = std::min(sc, max_rcoord - 1);
    const bool limit = !((double)r > (double)(nr_fine - 1) - (double)drdxfac);
    phibar[r] = avg_quad_interp(radius, RD[sc - 1], RD[sc], RD[sc + 1], PS[sc - 1], PS[sc], PS[sc + 1], limit);
  }
}

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

// multifab_min_c / multifab_max_c of one box (the `verbose >= 1` lines of density_advance.f90:374-402,
// enthalpy_advance.f90:440-449, velocity_advance.f90:142-160): grid-stride running min / max per thread, warp shuffles,
// per-CTA partials, the last few hundred partials on the host (min and max are exact in any order: bit-identical).
// dcomp >= 0: the value is s(comp) / s(dcomp) (multifab_div_div_c before the reduction; the state stays as it is --
// the reference multiplies it back, one rounding away from what it had).
__global__ void __launch_bounds__(256) k_minmax(DV s, Box3 vb, int comp, int dcomp, double* out) {
  const long nx = vb.hi[0] - vb.lo[0] + 1, ny = vb.hi[1] - vb.lo[1] + 1, nz = vb.hi[2] - vb.lo[2] + 1;
  const long rows = ny * nz;
  double lo = 1.7976931348623157e308, hi = -1.7976931348623157e308;
  for (long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int j = vb.lo[1] + (int)(r % ny), k = vb.lo[2] + (int)(r / ny);
    for (long t = threadIdx.x; t < nx; t += blockDim.x) {
      const int i = vb.lo[0] + (int)t;
      double v = s(i, j, k, comp);
      if (dcomp >= 0) v = v / s(i, j, k, dcomp);
      lo = fmin(lo, v);
      hi = fmax(hi, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double shl[8], shh[8];
  if ((threadIdx.x & 31) == 0) { shl[threadIdx.x >> 5] = lo; shh[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fmin(lo, shl[w]); hi = fmax(hi, shh[w]); }
    out[2 * blockIdx.x] = lo;
    out[2 * blockIdx.x + 1] = hi;
  }
}
void minmax_box_dev(const mgpu_params& P, const DV& s, const int* lo, const int* hi, int comp, int dcomp, double* mn,
                    double* mx) {
  Context& cx = ctx();
  const Box3 vb = grown(lo, hi, P.dm, 0);
  const long rows = (long)(vb.hi[1] - vb.lo[1] + 1) * (vb.hi[2] - vb.lo[2] + 1);
  const int nb = (int)std::min<long>(rows, 148 * 8);
  double* d = arena_alloc((size_t)2 * nb);
  MGPU_TIMED(TAG_GLUE, (k_minmax<<<nb, 256, 0, cx.stream>>>(s, vb, comp, dcomp, d)));
  std::vector<double> h((size_t)2 * nb);
  MGPU_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
  MGPU_CUDA(cudaStreamSynchronize(cx.stream));
  for (int b = 0; b < nb; ++b) {
    *mn = std::min(*mn, h[2 * b]);
    *mx = std::max(*mx, h[2 * b + 1]);
  }
}

}  // namespace mgpu
