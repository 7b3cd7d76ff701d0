// The EOS pieces next to the advective path, device-resident:
//   eos front end      Microphysics/EOS/eos.F90:99 (composition eos_type.f90:157, reset_inputs eos.F90:129, eos_reset :291)
//   gamma_law_general  Microphysics/EOS/gamma_law_general/gamma_law_general.f90:60
//   update_scal reset  Source/update_scal.f90:421-447 (2-D :296-318, spherical :559-585)
//   makeHfromRhoT_edge Source/rhoh_vs_t.f90:252 (_2d), :392 (_3d_cart), :596 (_3d_sphr)
//   mktempforce        Source/mkscalforce.f90:896 (_2d), :954 (_3d), :1026 (_3d_sphr)
//   firstdt            Source/firstdt.f90:330 (_2d), :460 (_3d), :599 (_3d_sphr)
//   makeTfromRhoH/P    Source/rhoh_vs_t.f90:929-1158 / :1287-1445
// All of it is streaming work: one thread per zone / face, a closed-form EOS of ~20 flops, so every kernel is bound by
// the HBM traffic of the state it reads (nspec + 2 doubles per zone) and the one double it writes.  This file is built
// without FMA contraction: results are bit-identical to the restated reference.
#include "mgpu_eos.cuh"

#include <algorithm>
#include <cmath>

namespace mgpu {

static EosDev g_eos = {};

void set_eos(const mgpu_eos* e) {
  if (!e) {
    g_eos = EosDev{};
    return;
  }
  if (e->kind != MGPU_EOS_GAMMA_LAW) throw Error("mgpu_set_eos: only gamma_law_general is available");
  if (e->nspec < 1 || e->nspec > MGPU_EOS_MAXSPEC) throw Error("mgpu_set_eos: nspec out of range");
  if (!(e->gamma > 0.0)) throw Error("eos_gamma cannot be < 0");  // gamma_law_general.f90:52
  EosDev d{};
  d.kind = e->kind;
  d.neutral = e->assume_neutral != 0;
  d.nspec = e->nspec;
  d.gamma = e->gamma;
  d.k_B = e->k_B;
  d.m_nucleon = 1.0 / e->n_A;  // gamma_law_general.f90:69
  d.mintemp = e->mintemp; d.maxtemp = e->maxtemp; d.mindens = e->mindens; d.maxdens = e->maxdens;
  d.mine = e->mine; d.maxe = e->maxe; d.minp = e->minp; d.maxp = e->maxp; d.minh = e->minh; d.maxh = e->maxh;
  d.small_temp = e->small_temp;
  for (int n = 0; n < e->nspec; ++n) { d.aion[n] = e->aion[n]; d.zion[n] = e->zion[n]; }
  g_eos = d;
}
bool have_eos() { return g_eos.kind != MGPU_EOS_NONE; }
const EosDev& the_eos(const char* who) {
  if (g_eos.kind == MGPU_EOS_NONE)
    throw Error(std::string(who) + ": EOS: not initialized (mgpu_set_eos)");  // eos.F90:117
  return g_eos;
}

namespace {

__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax2(double a, double b) { return a > b ? a : b; }

// actual_eos, gamma_law_general.f90:60-227 (the entropy and the electron quantities are not carried)
template <class XN>
__device__ __forceinline__ void actual_eos(const EosDev& E, int input, EosState& s, XN xn) {
  if (E.neutral) {
    s.mu = s.abar;
  } else {
    double sum = 0.0;
    for (int n = 0; n < E.nspec; ++n) sum = sum + (1.0 + E.zion[n]) * xn(n) / E.aion[n];
    s.mu = 1.0 / sum;
  }
  const double m_nucleon = E.m_nucleon, k_B = E.k_B, gamma_const = E.gamma;
  if (input == MGPU_EOS_INPUT_RH) s.T = (s.h * s.mu * m_nucleon / k_B) * (gamma_const - 1.0) / gamma_const;  // :104
  else if (input == MGPU_EOS_INPUT_TP) s.rho = s.p * s.mu * m_nucleon / (k_B * s.T);                          // :113
  else if (input == MGPU_EOS_INPUT_RP) s.T = s.p * s.mu * m_nucleon / (k_B * s.rho);                          // :122
  else if (input == MGPU_EOS_INPUT_RE) s.T = s.e * s.mu * m_nucleon * (gamma_const - 1.0) / k_B;              // :131
  const double Tinv = 1.0 / s.T, rhoinv = 1.0 / s.rho;
  s.p = s.rho * k_B * s.T / (s.mu * m_nucleon);  // :180
  s.e = s.p / (gamma_const - 1.0) * rhoinv;
  s.h = s.e + s.p * rhoinv;
  s.dpdT = s.p * Tinv;
  s.dpdr = s.p * rhoinv;
  s.dedT = s.e * Tinv;
  s.dedr = 0.0;
  s.dhdT = s.dedT + s.dpdT * rhoinv;
  s.cv = s.dedT;
  s.cp = gamma_const * s.cv;
  s.cs = sqrt(gamma_const * s.p * rhoinv);
}

__device__ __forceinline__ bool outside(double v, double lo, double hi) { return v < lo || v > hi; }

// eos(input, state), eos.F90:99-127; xn(n): mass fraction of species n (0-based)
template <class XN>
__device__ __forceinline__ void eos_call(const EosDev& E, int input, EosState& s, XN xn) {
  double sz = 0.0, sa = 0.0;  // composition, eos_type.f90:174-178
  for (int n = 0; n < E.nspec; ++n) sz = sz + xn(n) * E.zion[n] / E.aion[n];
  for (int n = 0; n < E.nspec; ++n) sa = sa + xn(n) / E.aion[n];
  s.mu_e = 1.0 / sz;
  s.abar = 1.0 / sa;
  s.zbar = s.abar / s.mu_e;
  bool reset = false;  // reset_inputs, eos.F90:129-183
  const bool clamp_rho = input == MGPU_EOS_INPUT_RT || input == MGPU_EOS_INPUT_RH || input == MGPU_EOS_INPUT_RP ||
                         input == MGPU_EOS_INPUT_RE;
  const bool clamp_T = input == MGPU_EOS_INPUT_RT || input == MGPU_EOS_INPUT_TP;
  if (clamp_rho) s.rho = dmin2(E.maxdens, dmax2(E.mindens, s.rho));
  if (clamp_T) s.T = dmin2(E.maxtemp, dmax2(E.mintemp, s.T));
  if (input == MGPU_EOS_INPUT_RH) reset = outside(s.h, E.minh, E.maxh);
  else if (input == MGPU_EOS_INPUT_TP || input == MGPU_EOS_INPUT_RP) reset = outside(s.p, E.minp, E.maxp);
  else if (input == MGPU_EOS_INPUT_RE) reset = outside(s.e, E.mine, E.maxe);
  if (reset) {  // eos_reset, eos.F90:291-308
    s.T = dmin2(E.maxtemp, dmax2(E.mintemp, s.T));
    s.rho = dmin2(E.maxdens, dmax2(E.mindens, s.rho));
    actual_eos(E, MGPU_EOS_INPUT_RT, s, xn);
  } else {
    actual_eos(E, input, s, xn);
  }
}

// species of one zone / face of a fab, divided by `div` or not
struct XnFab {
  const double* p;  // address of the first species at this point
  long cs;
  double div;
  bool divide;
  __device__ __forceinline__ double operator()(int n) const { return divide ? p[cs * n] / div : p[cs * n]; }
};

__global__ void k_eos_points(EosDev E, int input, long n, double* state, const double* xn) {
  const long i = MGPU_TID;
  if (i >= n) return;
  EosState s{};
  s.rho = state[MGPU_EOS_Q_RHO * n + i];
  s.T = state[MGPU_EOS_Q_T * n + i];
  s.p = state[MGPU_EOS_Q_P * n + i];
  s.e = state[MGPU_EOS_Q_E * n + i];
  s.h = state[MGPU_EOS_Q_H * n + i];
  eos_call(E, input, s, XnFab{xn + i, n, 1.0, false});
  const double out[MGPU_EOS_NQ] = {s.rho, s.T, s.p, s.e, s.h, s.cv, s.cp, s.cs, s.dpdT, s.dpdr, s.dedT, s.dedr,
                                   s.dhdT, s.mu, s.abar, s.zbar};
#pragma unroll
  for (int q = 0; q < MGPU_EOS_NQ; ++q) state[(long)q * n + i] = out[q];
}

__global__ void k_update_scal_eos(EosDev E, DV sold, DV snew, const double* p0_new, DV p0c, bool sphr, int dm, Box3 vb,
                                  int rho, int rhoh, int temp, int spec0, double cutoff) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const long on = snew.off(i, j, k);
  const double r = snew.p[on + snew.cs * rho];
  if (!(r <= cutoff)) return;
  EosState s{};
  s.rho = r;
  s.T = sold(i, j, k, temp);
  s.p = sphr ? p0c(i, j, k) : p0_new[dm == 3 ? k : j];
  eos_call(E, MGPU_EOS_INPUT_RP, s, XnFab{snew.p + on + snew.cs * spec0, snew.cs, s.rho, true});
  snew.p[on + snew.cs * rhoh] = r * s.h;
}

template <int D>
__global__ void k_h_from_rhot_edge(EosDev E, HEdgeArgs a) {
  int ix[3];
  Box3 fb = a.vb;
  fb.hi[D] += 1;
  if (!decode3(fb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const DV& se = a.sedge[D];
  const long o = se.off(i, j, k);
  const int r = (a.dm == 3) ? k : j;
  const bool radial = (D == a.dm - 1);
  EosState s{};
  double rhoh0_edge = 0.0;
  if (a.sphr) {  // rhoh_vs_t.f90:596
    const int im = i - (D == 0), jm = j - (D == 1), km = k - (D == 2);
    if (a.ept == MGPU_PREDICT_TPRIME_THEN_H) {
      const double t0_edge = 0.5 * (a.t0_cart(im, jm, km) + a.t0_cart(i, j, k));
      s.T = dmax2(se.p[o + se.cs * a.temp] + t0_edge, E.small_temp);
    } else {
      s.T = dmax2(se.p[o + se.cs * a.temp], E.small_temp);
    }
    s.rho = se.p[o + se.cs * a.rho];
    if (a.spt == MGPU_PREDICT_RHOPRIME_AND_X) s.rho = s.rho + 0.5 * (a.rho0_cart(im, jm, km) + a.rho0_cart(i, j, k));
    if (a.ept == MGPU_PREDICT_T_THEN_RHOHPRIME) rhoh0_edge = 0.5 * (a.rhoh0_cart(im, jm, km) + a.rhoh0_cart(i, j, k));
  } else {
    if (a.ept == MGPU_PREDICT_TPRIME_THEN_H) {
      const double t0_edge = radial ? 0.5 * (a.t0_edge_old[r] + a.t0_edge_new[r]) : 0.5 * (a.t0_old[r] + a.t0_new[r]);
      // QUIRK rhoh_vs_t.f90:487, :543: the 3-D Cartesian y and z faces read T' from the x-face array
      const double tp = (a.dm == 3) ? a.sedge[0](i, j, k, a.temp) : se.p[o + se.cs * a.temp];
      s.T = dmax2(tp + t0_edge, E.small_temp);
    } else {
      s.T = dmax2(se.p[o + se.cs * a.temp], E.small_temp);
    }
    s.rho = se.p[o + se.cs * a.rho];
    if (a.spt == MGPU_PREDICT_RHOPRIME_AND_X)
      s.rho = s.rho + (radial ? 0.5 * (a.rho0_edge_old[r] + a.rho0_edge_new[r]) : 0.5 * (a.rho0_old[r] + a.rho0_new[r]));
    if (a.ept == MGPU_PREDICT_T_THEN_RHOHPRIME)
      rhoh0_edge = radial ? 0.5 * (a.rhoh0_edge_old[r] + a.rhoh0_edge_new[r]) : 0.5 * (a.rhoh0_old[r] + a.rhoh0_new[r]);
  }
  eos_call(E, MGPU_EOS_INPUT_RT, s, XnFab{se.p + o + se.cs * a.spec0, se.cs, s.rho, a.spt == MGPU_PREDICT_RHOX});
  if (a.ept == MGPU_PREDICT_T_THEN_H || a.ept == MGPU_PREDICT_TPRIME_THEN_H) se.p[o + se.cs * a.rhoh] = s.h;
  else if (a.ept == MGPU_PREDICT_T_THEN_RHOHPRIME) se.p[o + se.cs * a.rhoh] = s.rho * s.h - rhoh0_edge;
}

__global__ void k_mktempforce(EosDev E, TempForceArgs a) {
  int ix[3];
  if (!decode3(a.vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const long o = a.s.off(i, j, k);
  const double sr = a.s.p[o + a.s.cs * a.rho];
  EosState st{};
  st.T = a.s.p[o + a.s.cs * a.temp];
  st.rho = sr;
  eos_call(E, MGPU_EOS_INPUT_RT, st, XnFab{a.s.p + o + a.s.cs * a.spec0, a.s.cs, sr, true});
  const double dhdp = 1.0 / sr + (sr * st.dedr - st.p / sr) / (sr * st.dpdr);
  double adv;
  if (a.sphr) {  // mkscalforce.f90:1087-1106
    const DV& pc = a.p0_cart;
    const double c = pc(i, j, k);
    const double p0_lox = 0.5 * (c + pc(i - 1, j, k)), p0_hix = 0.5 * (c + pc(i + 1, j, k));
    const double p0_loy = 0.5 * (c + pc(i, j - 1, k)), p0_hiy = 0.5 * (c + pc(i, j + 1, k));
    const double p0_loz = 0.5 * (c + pc(i, j, k - 1)), p0_hiz = 0.5 * (c + pc(i, j, k + 1));
    const DV &um = a.umac[0], &vm = a.umac[1], &wm = a.umac[2];
    const double divup = (um(i + 1, j, k) * p0_hix - um(i, j, k) * p0_lox) / a.dx[0] +
                         (vm(i, j + 1, k) * p0_hiy - vm(i, j, k) * p0_loy) / a.dx[1] +
                         (wm(i, j, k + 1) * p0_hiz - wm(i, j, k) * p0_loz) / a.dx[2];
    const double p0divu = ((um(i + 1, j, k) - um(i, j, k)) / a.dx[0] + (vm(i, j + 1, k) - vm(i, j, k)) / a.dx[1] +
                           (wm(i, j, k + 1) - wm(i, j, k)) / a.dx[2]) * c;
    const double ugradp = divup - p0divu;
    adv = ugradp + a.psi_cart(i, j, k);
  } else {
    const int r = (a.dm == 3) ? k : j;
    double gradp0;
    if (r == 0) gradp0 = (a.p0_old[r + 1] - a.p0_old[r]) / a.dr;
    else if (r == a.nr - 1) gradp0 = (a.p0_old[r] - a.p0_old[r - 1]) / a.dr;
    else gradp0 = 0.5 * (a.p0_old[r + 1] - a.p0_old[r - 1]) / a.dr;
    const DV& wm = a.umac[a.dm - 1];
    const double wadv = (a.dm == 3) ? 0.5 * (wm(i, j, k + 1) + wm(i, j, k)) : 0.5 * (wm(i, j, k) + wm(i, j + 1, k));
    adv = wadv * gradp0 + a.psi[r];
  }
  double f = a.thermal(i, j, k) + (1.0 - sr * dhdp) * adv;
  f = f / (st.cp * sr);
  a.f(i, j, k) = f;
}

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

// slots of the firstdt partial results (one set per CTA): maxima first, the minimum last
enum { F_CS = 0, F_F0, F_F1, F_F2, F_U0, F_U1, F_U2, F_DT_DIVU, F_N };
struct FirstdtArgs {
  int dm, nr, rho, temp, spec0;
  bool sphr, use_divu;
  double dx_r, rho_min;
  Box3 vb;
  DV u, s, force, divU, gp0c;
  const double *p0, *gamma1bar;
  double* part;
};
// one pass over the valid cells: the running extrema of firstdt.f90:503-533 (sound speed, force, velocity) and the divU
// constraint (:577-598 / :725-743); max / min are exact whatever the order
__global__ void __launch_bounds__(256) k_firstdt(EosDev E, FirstdtArgs a) {
  const int nx = a.vb.hi[0] - a.vb.lo[0] + 1, ny = a.vb.hi[1] - a.vb.lo[1] + 1;
  const long npts = a.vb.npts();
  double m[F_N];
#pragma unroll
  for (int q = 0; q < F_DT_DIVU; ++q) m[q] = 0.0;
  m[F_DT_DIVU] = 1.e99;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < npts; t += (long)gridDim.x * blockDim.x) {
    const int i = a.vb.lo[0] + (int)(t % nx);
    const int j = a.vb.lo[1] + (int)((t / nx) % ny);
    const int k = a.vb.lo[2] + (int)(t / ((long)nx * ny));
    const long os = a.s.off(i, j, k), ou = a.u.off(i, j, k), of = a.force.off(i, j, k);
    EosState st{};
    st.rho = a.s.p[os + a.s.cs * a.rho];
    st.T = a.s.p[os + a.s.cs * a.temp];
    const double rho_in = st.rho;
    eos_call(E, MGPU_EOS_INPUT_RT, st, XnFab{a.s.p + os + a.s.cs * a.spec0, a.s.cs, rho_in, true});
    m[F_CS] = fmax(m[F_CS], st.cs);
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d < a.dm) {
        m[F_F0 + d] = fmax(m[F_F0 + d], fabs(a.force.p[of + a.force.cs * d]));
        m[F_U0 + d] = fmax(m[F_U0 + d], fabs(a.u.p[ou + a.u.cs * d]));
      }
    if (a.use_divu) {
      double denom;
      if (a.sphr) {
        const long og = a.gp0c.off(i, j, k);
        const double gp_dot_u = a.u.p[ou] * a.gp0c.p[og] + a.u.p[ou + a.u.cs] * a.gp0c.p[og + a.gp0c.cs] +
                                a.u.p[ou + 2 * a.u.cs] * a.gp0c.p[og + 2 * a.gp0c.cs];
        denom = a.divU(i, j, k) - gp_dot_u;
      } else {
        const int r = (a.dm == 3) ? k : j;
        double gradp0;
        if (r == 0) gradp0 = (a.p0[r + 1] - a.p0[r]) / a.dx_r;
        else if (r == a.nr - 1) gradp0 = (a.p0[r] - a.p0[r - 1]) / a.dx_r;
        else gradp0 = 0.5 * (a.p0[r + 1] - a.p0[r - 1]) / a.dx_r;
        denom = a.divU(i, j, k) - a.u.p[ou + a.u.cs * (a.dm - 1)] * gradp0 / (a.gamma1bar[r] * a.p0[r]);
      }
      if (denom > 0.0) m[F_DT_DIVU] = fmin(m[F_DT_DIVU], 0.4 * (1.0 - a.rho_min / rho_in) / denom);
    }
  }
  __shared__ double sh[8][F_N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < F_N; ++q) {
    const double v = q < F_DT_DIVU ? warp_max(m[q]) : warp_min(m[q]);
    if (lane == 0) sh[w][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < F_N) {
    const int q = threadIdx.x;
    double v = sh[0][q];
    for (int ww = 1; ww < 8; ++ww) v = q < F_DT_DIVU ? fmax(v, sh[ww][q]) : fmin(v, sh[ww][q]);
    a.part[(long)blockIdx.x * F_N + q] = v;
  }
}

__global__ void k_make_t(EosDev E, DV state, const double* p0, DV p0c, bool sphr, int dm, Box3 vb, int rho, int rhoh,
                         int temp, int spec0, int pi, bool from_p, bool flag, bool update_rhoh) {
  int ix[3];
  if (!decode3(vb, ix)) return;
  const int i = ix[0], j = ix[1], k = ix[2];
  const long o = state.off(i, j, k);
  const double p0v = sphr ? p0c(i, j, k) : p0[dm == 3 ? k : j];
  const double r = state.p[o + state.cs * rho];
  EosState st{};
  st.rho = r;
  st.T = state.p[o + state.cs * temp];
  const XnFab xn{state.p + o + state.cs * spec0, state.cs, r, true};
  if (from_p) {
    st.p = flag ? p0v + state.p[o + state.cs * pi] : p0v;  // use_pprime_in_tfromp
    eos_call(E, MGPU_EOS_INPUT_RP, st, xn);
    state.p[o + state.cs * temp] = st.T;
    if (update_rhoh) state.p[o + state.cs * rhoh] = st.rho * st.h;
  } else if (flag) {  // use_eos_e_instead_of_h
    st.e = state.p[o + state.cs * rhoh] / r - p0v / r;
    eos_call(E, MGPU_EOS_INPUT_RE, st, xn);
    state.p[o + state.cs * temp] = st.T;
  } else {
    st.h = state.p[o + state.cs * rhoh] / r;
    eos_call(E, MGPU_EOS_INPUT_RH, st, xn);
    state.p[o + state.cs * temp] = st.T;
  }
}

void check_input(int input) {
  switch (input) {
    case MGPU_EOS_INPUT_RT: case MGPU_EOS_INPUT_RH: case MGPU_EOS_INPUT_TP: case MGPU_EOS_INPUT_RP:
    case MGPU_EOS_INPUT_RE:
      return;
    case MGPU_EOS_INPUT_PS: throw Error("EOS: eos_input_ps needs the entropy, which is not evaluated here");
    case MGPU_EOS_INPUT_PH: throw Error("EOS: eos_input_ph not implemented");                            // :151
    case MGPU_EOS_INPUT_TH: throw Error("EOS: eos_input_th is not a valid input for the gamma law EOS"); // :160
    default: throw Error("EOS: invalid input");
  }
}

}  // namespace

void eos_points_dev(int input, long n, double* state, const double* xn) {
  const EosDev& E = the_eos("eos");
  check_input(input);
  if (n <= 0) return;
  MGPU_TIMED(TAG_GLUE, (k_eos_points<<<nblocks(n, 256), 256, 0, ctx().stream>>>(E, input, n, state, xn)));
}

void update_scal_eos_dev(const mgpu_params& P, const DV& sold, const DV& snew, const double* p0_new, const DV* p0_new_cart,
                         const Box3& vb) {
  const EosDev& E = the_eos("update_scal");
  const bool sphr = P.spherical != 0;
  if (sphr ? !p0_new_cart : !p0_new) throw Error("update_scal: the EOS reset below the cutoff needs p0_new");
  MGPU_TIMED(TAG_UPDATE, (k_update_scal_eos<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
                             E, sold, snew, p0_new, sphr ? *p0_new_cart : snew, sphr, P.dm, vb, P.rho_comp - 1,
                             P.rhoh_comp - 1, P.temp_comp - 1, P.spec_comp - 1, P.base_cutoff_density)));
}

void h_from_rhot_edge_dev(const HEdgeArgs& a) {
  const EosDev& E = the_eos("makeHfromRhoT_edge");
  cudaStream_t s = ctx().stream;
  for (int d = 0; d < a.dm; ++d) {
    Box3 fb = a.vb;
    fb.hi[d] += 1;
    const dim3 g = grid3(fb, 256);
    const int b = block3(fb, 256);
    if (d == 0) MGPU_TIMED(TAG_GLUE, (k_h_from_rhot_edge<0><<<g, b, 0, s>>>(E, a)));
    else if (d == 1) MGPU_TIMED(TAG_GLUE, (k_h_from_rhot_edge<1><<<g, b, 0, s>>>(E, a)));
    else MGPU_TIMED(TAG_GLUE, (k_h_from_rhot_edge<2><<<g, b, 0, s>>>(E, a)));
  }
}

void mktempforce_dev(const TempForceArgs& a) {
  const EosDev& E = the_eos("mktempforce");
  MGPU_TIMED(TAG_GLUE, (k_mktempforce<<<grid3(a.vb, 256), block3(a.vb, 256), 0, ctx().stream>>>(E, a)));
}

void firstdt_box_dev(const mgpu_params& P, const DV& u, const DV& s, const DV& force, const DV& divU, const double* p0,
                     const double* gamma1bar, const DV* gp0_cart, const int* lo, const int* hi, double cfl,
                     bool use_soundspeed, bool use_divu, double* dt, double* umax) {
  const EosDev& E = the_eos("firstdt");
  Context& cx = ctx();
  const int dm = P.dm;
  FirstdtArgs a;
  a.dm = dm;
  a.nr = P.nr;
  a.rho = P.rho_comp - 1; a.temp = P.temp_comp - 1; a.spec0 = P.spec_comp - 1;
  a.sphr = gp0_cart != nullptr;
  a.use_divu = use_divu;
  a.dx_r = P.dx[dm - 1];
  a.rho_min = 1.e-20;  // firstdt.f90:493
  a.vb = grown(lo, hi, dm, 0);
  a.u = u; a.s = s; a.force = force; a.divU = divU;
  a.gp0c = gp0_cart ? *gp0_cart : u;
  a.p0 = p0; a.gamma1bar = gamma1bar;
  const unsigned nb = std::min<unsigned>(nblocks(a.vb.npts(), 256), 148u * 8u);
  a.part = arena_alloc((size_t)nb * F_N);
  MGPU_TIMED(TAG_GLUE, (k_firstdt<<<nb, 256, 0, cx.stream>>>(E, a)));
  std::vector<double> part((size_t)nb * F_N);
  MGPU_CUDA(cudaMemcpyAsync(part.data(), a.part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
  MGPU_CUDA(cudaStreamSynchronize(cx.stream));
  double m[F_N];
  for (int q = 0; q < F_N; ++q) m[q] = part[q];
  for (unsigned b = 1; b < nb; ++b)
    for (int q = 0; q < F_N; ++q)
      m[q] = q < F_DT_DIVU ? std::max(m[q], part[(size_t)b * F_N + q]) : std::min(m[q], part[(size_t)b * F_N + q]);
  // the scalar tail of firstdt_2d / _3d, statement by statement (firstdt.f90:535-600)
  const double eps = 1.0e-8;
  double t = 1.e99, um = 0.0;
  for (int d = 0; d < dm; ++d) um = std::max(um, m[F_U0 + d]);
  *umax = um;
  double ud[3], sd[3], umx = 0.0, smx = 0.0;
  bool any_u = false, all_spd = true, no_spd = true;
  for (int d = 0; d < dm; ++d) {
    ud[d] = m[F_U0 + d] / P.dx[d];
    sd[d] = m[F_CS] / P.dx[d];
    if (ud[d] != 0.0) any_u = true;
    if (sd[d] == 0.0) all_spd = false; else no_spd = false;
    umx = std::max(umx, ud[d]);
    smx = std::max(smx, sd[d]);
  }
  if (any_u) t = cfl / umx;
  else if (all_spd) t = cfl / smx;
  if (use_soundspeed) t = std::min(t, no_spd ? 1.e99 : cfl / smx);
  for (int d = 0; d < dm; ++d)
    if (m[F_F0 + d] > eps) t = std::min(t, std::sqrt(2.0 * P.dx[d] / m[F_F0 + d]));
  if (use_divu) t = std::min(t, m[F_DT_DIVU]);
  *dt = t;
}

void make_t_dev(const mgpu_params& P, const DV& state, const double* p0, const DV* p0_cart, bool from_p, bool flag,
                bool update_rhoh, const Box3& vb) {
  const EosDev& E = the_eos(from_p ? "makeTfromRhoP" : "makeTfromRhoH");
  MGPU_TIMED(TAG_GLUE, (k_make_t<<<grid3(vb, 256), block3(vb, 256), 0, ctx().stream>>>(
                           E, state, p0, p0_cart ? *p0_cart : state, p0_cart != nullptr, P.dm, vb, P.rho_comp - 1,
                           P.rhoh_comp - 1, P.temp_comp - 1, P.spec_comp - 1, P.pi_comp - 1, from_p, flag, update_rhoh)));
}

}  // namespace mgpu
