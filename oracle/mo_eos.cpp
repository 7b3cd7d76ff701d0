// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
// The EOS pieces next to the advective path, restated from the reference:
//   eos front end      Microphysics/EOS/eos.F90:99 (composition eos_type.f90:157, reset_inputs eos.F90:129, eos_reset :291)
//   gamma_law_general  Microphysics/EOS/gamma_law_general/gamma_law_general.f90:60
//   update_scal reset  Source/update_scal.f90:421-447 (2-D :296-318, spherical :559-585)
//   makeHfromRhoT_edge Source/rhoh_vs_t.f90:20 (_2d :252, _3d_cart :392, _3d_sphr :596)
//   mktempforce        Source/mkscalforce.f90:719 (_2d :896, _3d :954, _3d_sphr :1026)
//   firstdt            Source/firstdt.f90:25 (_2d :330, _3d :460, _3d_sphr :599)
//   makeTfromRhoH/P    Source/rhoh_vs_t.f90:800 / :1165
// parity unpinned for this file: the reference holds no golden numbers for its EOS (test_eos writes plotfiles only); the
// tests anchor it on the ideal-gas identities instead (tests/test_eos_cpu.py).
#include <algorithm>
#include <string>

#include "mo_kernels.h"

namespace mo {

static mgpu_eos g_eos = {};

void set_eos(const mgpu_eos* e) {
  if (!e) {
    g_eos = mgpu_eos{};
    return;
  }
  if (e->kind != MGPU_EOS_GAMMA_LAW) fail("mgpu_set_eos: only gamma_law_general is available");
  if (e->nspec < 1 || e->nspec > MGPU_EOS_MAXSPEC) fail("mgpu_set_eos: nspec out of range");
  if (!(e->gamma > 0.0)) fail("eos_gamma cannot be < 0");  // gamma_law_general.f90:52
  g_eos = *e;
}
bool have_eos() { return g_eos.kind != MGPU_EOS_NONE; }
const mgpu_eos& the_eos(const char* who) {
  if (g_eos.kind == MGPU_EOS_NONE) fail((std::string(who) + ": EOS: not initialized").c_str());  // eos.F90:117
  return g_eos;
}

namespace {

// actual_eos, gamma_law_general.f90:60-227 (entropy and the electron quantities are not carried)
void actual_eos(const mgpu_eos& E, int input, EosState& s, const double* xn) {
  const double m_nucleon = 1.0 / E.n_A;  // :69
  const double k_B = E.k_B;
  const double gamma_const = E.gamma;
  if (E.assume_neutral) {
    s.mu = s.abar;
  } else {
    double sum = 0.0;
    for (int n = 0; n < E.nspec; ++n) sum = sum + (1.0 + E.zion[n]) * xn[n] / E.aion[n];
    s.mu = 1.0 / sum;
  }
  switch (input) {
    case MGPU_EOS_INPUT_RT: break;
    case MGPU_EOS_INPUT_RH: s.T = (s.h * s.mu * m_nucleon / k_B) * (gamma_const - 1.0) / gamma_const; break;  // :104
    case MGPU_EOS_INPUT_TP: s.rho = s.p * s.mu * m_nucleon / (k_B * s.T); break;                              // :113
    case MGPU_EOS_INPUT_RP: s.T = s.p * s.mu * m_nucleon / (k_B * s.rho); break;                              // :122
    case MGPU_EOS_INPUT_RE: s.T = s.e * s.mu * m_nucleon * (gamma_const - 1.0) / k_B; break;                  // :131
    case MGPU_EOS_INPUT_PS: fail("EOS: eos_input_ps needs the entropy, which is not evaluated here");
    case MGPU_EOS_INPUT_PH: fail("EOS: eos_input_ph not implemented");                                        // :151
    case MGPU_EOS_INPUT_TH: fail("EOS: eos_input_th is not a valid input for the gamma law EOS");             // :160
    default: fail("EOS: invalid input");
  }
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
  s.cs = std::sqrt(gamma_const * s.p * rhoinv);
}

inline bool outside(double v, double lo, double hi) { return v < lo || v > hi; }

}  // namespace

// eos(input, state), eos.F90:99-127
void eos_call(int input, EosState& s, const double* xn) {
  const mgpu_eos& E = the_eos("eos");
  // the modes gamma_law_general cannot serve raise up front (in the reference an out-of-range input would take the
  // eos_reset path first and hide the error; the product checks the mode on the host before it launches)
  if (input == MGPU_EOS_INPUT_PS) fail("EOS: eos_input_ps needs the entropy, which is not evaluated here");
  if (input == MGPU_EOS_INPUT_PH) fail("EOS: eos_input_ph not implemented");
  if (input == MGPU_EOS_INPUT_TH) fail("EOS: eos_input_th is not a valid input for the gamma law EOS");
  if (input < MGPU_EOS_INPUT_RT || input > MGPU_EOS_INPUT_TH) fail("EOS: invalid input");
  double sz = 0.0, sa = 0.0;  // composition, eos_type.f90:174-178
  for (int n = 0; n < E.nspec; ++n) sz = sz + xn[n] * E.zion[n] / E.aion[n];
  for (int n = 0; n < E.nspec; ++n) sa = sa + xn[n] / E.aion[n];
  s.mu_e = 1.0 / sz;
  s.abar = 1.0 / sa;
  s.zbar = s.abar / s.mu_e;
  auto reset_rho = [&] { s.rho = dmin(E.maxdens, dmax(E.mindens, s.rho)); };
  auto reset_T = [&] { s.T = dmin(E.maxtemp, dmax(E.mintemp, s.T)); };
  bool has_been_reset = false;
  auto eos_reset = [&] {  // eos.F90:291-308
    reset_T();
    reset_rho();
    actual_eos(E, MGPU_EOS_INPUT_RT, s, xn);
    has_been_reset = true;
  };
  switch (input) {  // reset_inputs, eos.F90:129-183
    case MGPU_EOS_INPUT_RT: reset_rho(); reset_T(); break;
    case MGPU_EOS_INPUT_RH: reset_rho(); if (outside(s.h, E.minh, E.maxh)) eos_reset(); break;
    case MGPU_EOS_INPUT_TP: reset_T(); if (outside(s.p, E.minp, E.maxp)) eos_reset(); break;
    case MGPU_EOS_INPUT_RP: reset_rho(); if (outside(s.p, E.minp, E.maxp)) eos_reset(); break;
    case MGPU_EOS_INPUT_RE: reset_rho(); if (outside(s.e, E.mine, E.maxe)) eos_reset(); break;
    case MGPU_EOS_INPUT_PS: fail("EOS: eos_input_ps needs the entropy, which is not evaluated here");
    case MGPU_EOS_INPUT_PH:
      if (outside(s.p, E.minp, E.maxp)) eos_reset();
      if (outside(s.h, E.minh, E.maxh)) eos_reset();
      break;
    case MGPU_EOS_INPUT_TH: reset_T(); if (outside(s.h, E.minh, E.maxh)) eos_reset(); break;
    default: break;
  }
  if (!has_been_reset) actual_eos(E, input, s, xn);
}

// the state of one zone / face: species from the fab, divided by the density or not
static void load_xn(const Arr& a, int i, int j, int k, int spec0, int nspec, double div, bool divide, double* xn) {
  for (int n = 0; n < nspec; ++n) xn[n] = divide ? a(i, j, k, spec0 + n) / div : a(i, j, k, spec0 + n);
}

// update_scal.f90:421-447 (p0 from the 1-D array) / :559-585 (p0_new_cart)
void update_scal_eos_box(const mgpu_params& P, const Arr& sold, Arr& snew, const double* p0_new, const Arr* p0_new_cart,
                         const int* lo, const int* hi) {
  const mgpu_eos& E = the_eos("update_scal");
  const int dm = P.dm, rho = P.rho_comp - 1, rhoh = P.rhoh_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1;
  if (P.spherical ? !p0_new_cart : !p0_new) fail("update_scal: the EOS reset below the cutoff needs p0_new");
  for_box(grown(lo, hi, dm, 0), [&](int i, int j, int k) {
    if (snew(i, j, k, rho) <= P.base_cutoff_density) {
      EosState s{};
      double xn[MGPU_EOS_MAXSPEC];
      s.rho = snew(i, j, k, rho);
      s.T = sold(i, j, k, temp);
      s.p = P.spherical ? (*p0_new_cart)(i, j, k, 0) : p0_new[dm == 3 ? k : j];
      load_xn(snew, i, j, k, spec0, E.nspec, s.rho, true, xn);
      eos_call(MGPU_EOS_INPUT_RP, s, xn);
      snew(i, j, k, rhoh) = snew(i, j, k, rho) * s.h;
    }
  });
}

// makeHfromRhoT_edge_2d / _3d_cart (rhoh_vs_t.f90:252, :392)
void h_from_rhot_edge_box(const mgpu_params& P, Arr* sedge, const double* rho0_old, const double* rhoh0_old,
                          const double* t0_old, const double* rho0_edge_old, const double* rhoh0_edge_old,
                          const double* t0_edge_old, const double* rho0_new, const double* rhoh0_new,
                          const double* t0_new, const double* rho0_edge_new, const double* rhoh0_edge_new,
                          const double* t0_edge_new, const int* lo, const int* hi) {
  const mgpu_eos& E = the_eos("makeHfromRhoT_edge");
  const int dm = P.dm, rho = P.rho_comp - 1, rhoh = P.rhoh_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1;
  const int ept = P.enthalpy_pred_type, spt = P.species_pred_type;
  for (int d = 0; d < dm; ++d) {
    Box b = grown(lo, hi, dm, 0);
    b.hi[d] += 1;
    const bool radial = (d == dm - 1);
    const Arr& se = sedge[d];
    // QUIRK rhoh_vs_t.f90:487, :543: the 3-D Cartesian y and z faces read T' from the x-face array
    const Arr& sT = (dm == 3 && ept == MGPU_PREDICT_TPRIME_THEN_H) ? sedge[0] : se;
    for_box(b, [&](int i, int j, int k) {
      const int r = (dm == 3) ? k : j;
      EosState s{};
      double xn[MGPU_EOS_MAXSPEC];
      if (ept == MGPU_PREDICT_TPRIME_THEN_H) {
        const double t0_edge = radial ? 0.5 * (t0_edge_old[r] + t0_edge_new[r]) : 0.5 * (t0_old[r] + t0_new[r]);
        s.T = dmax(sT(i, j, k, temp) + t0_edge, E.small_temp);
      } else {
        s.T = dmax(se(i, j, k, temp), E.small_temp);
      }
      if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
        s.rho = se(i, j, k, rho) + (radial ? 0.5 * (rho0_edge_old[r] + rho0_edge_new[r]) : 0.5 * (rho0_old[r] + rho0_new[r]));
        load_xn(se, i, j, k, spec0, E.nspec, 1.0, false, xn);
      } else if (spt == MGPU_PREDICT_RHOX) {
        s.rho = se(i, j, k, rho);
        load_xn(se, i, j, k, spec0, E.nspec, s.rho, true, xn);
      } else {
        s.rho = se(i, j, k, rho);
        load_xn(se, i, j, k, spec0, E.nspec, 1.0, false, xn);
      }
      eos_call(MGPU_EOS_INPUT_RT, s, xn);
      if (ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H) {
        se(i, j, k, rhoh) = s.h;
      } else if (ept == MGPU_PREDICT_T_THEN_RHOHPRIME) {
        se(i, j, k, rhoh) = s.rho * s.h - (radial ? 0.5 * (rhoh0_edge_old[r] + rhoh0_edge_new[r])
                                                  : 0.5 * (rhoh0_old[r] + rhoh0_new[r]));
      }
    });
  }
}

// makeHfromRhoT_edge_3d_sphr (rhoh_vs_t.f90:596): the three Cartesian half-time base-state arrays have 2 ghost cells
void h_from_rhot_edge_sphr_box(const mgpu_params& P, Arr* sedge, const Arr& rho0_cart, const Arr& rhoh0_cart,
                               const Arr& t0_cart, const int* lo, const int* hi) {
  const mgpu_eos& E = the_eos("makeHfromRhoT_edge");
  const int rho = P.rho_comp - 1, rhoh = P.rhoh_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1;
  const int ept = P.enthalpy_pred_type, spt = P.species_pred_type;
  for (int d = 0; d < 3; ++d) {
    Box b = grown(lo, hi, 3, 0);
    b.hi[d] += 1;
    const Arr& se = sedge[d];
    for_box(b, [&](int i, int j, int k) {
      const int im = i - (d == 0), jm = j - (d == 1), km = k - (d == 2);
      EosState s{};
      double xn[MGPU_EOS_MAXSPEC];
      if (ept == MGPU_PREDICT_TPRIME_THEN_H) {
        const double t0_edge = 0.5 * (t0_cart(im, jm, km) + t0_cart(i, j, k));
        s.T = dmax(se(i, j, k, temp) + t0_edge, E.small_temp);
      } else {
        s.T = dmax(se(i, j, k, temp), E.small_temp);
      }
      if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
        const double rho0_edge = 0.5 * (rho0_cart(im, jm, km) + rho0_cart(i, j, k));
        s.rho = se(i, j, k, rho) + rho0_edge;
        load_xn(se, i, j, k, spec0, E.nspec, 1.0, false, xn);
      } else if (spt == MGPU_PREDICT_RHOX) {
        s.rho = se(i, j, k, rho);
        load_xn(se, i, j, k, spec0, E.nspec, s.rho, true, xn);
      } else {
        s.rho = se(i, j, k, rho);
        load_xn(se, i, j, k, spec0, E.nspec, 1.0, false, xn);
      }
      eos_call(MGPU_EOS_INPUT_RT, s, xn);
      if (ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H) {
        se(i, j, k, rhoh) = s.h;
      } else if (ept == MGPU_PREDICT_T_THEN_RHOHPRIME) {
        const double rhoh0_edge = 0.5 * (rhoh0_cart(im, jm, km) + rhoh0_cart(i, j, k));
        se(i, j, k, rhoh) = s.rho * s.h - rhoh0_edge;
      }
    });
  }
}

// mktempforce_2d / _3d (mkscalforce.f90:896, :954): valid cells of comp temp_comp
void mktempforce_box(const mgpu_params& P, Arr& temp_force, const Arr& s, const Arr* umac, const Arr& thermal,
                     const double* p0_old, const double* psi, const int* lo, const int* hi) {
  const mgpu_eos& E = the_eos("mktempforce");
  const int dm = P.dm, rho = P.rho_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1, nr = P.nr;
  const double dr = P.dx[dm - 1];
  const Arr& wmac = umac[dm - 1];
  for_box(grown(lo, hi, dm, 0), [&](int i, int j, int k) {
    const int r = (dm == 3) ? k : j;
    double gradp0;
    if (r == 0) gradp0 = (p0_old[r + 1] - p0_old[r]) / dr;
    else if (r == nr - 1) gradp0 = (p0_old[r] - p0_old[r - 1]) / dr;
    else gradp0 = 0.5 * (p0_old[r + 1] - p0_old[r - 1]) / dr;
    EosState st{};
    double xn[MGPU_EOS_MAXSPEC];
    st.T = s(i, j, k, temp);
    st.rho = s(i, j, k, rho);
    load_xn(s, i, j, k, spec0, E.nspec, s(i, j, k, rho), true, xn);
    eos_call(MGPU_EOS_INPUT_RT, st, xn);
    const double sr = s(i, j, k, rho);
    const double dhdp = 1.0 / sr + (sr * st.dedr - st.p / sr) / (sr * st.dpdr);
    const double wadv = (dm == 3) ? 0.5 * (wmac(i, j, k + 1) + wmac(i, j, k)) : 0.5 * (wmac(i, j, k) + wmac(i, j + 1, k));
    double f = thermal(i, j, k) + (1.0 - sr * dhdp) * (wadv * gradp0 + psi[r]);
    f = f / (st.cp * sr);
    temp_force(i, j, k, temp) = f;
  });
}

// mktempforce_3d_sphr (mkscalforce.f90:1026): p0_cart has one ghost cell, psi_cart covers the valid cells
void mktempforce_sphr_box(const mgpu_params& P, Arr& temp_force, const Arr& s, const Arr* umac, const Arr& thermal,
                          const Arr& p0_cart, const Arr& psi_cart, const int* lo, const int* hi) {
  const mgpu_eos& E = the_eos("mktempforce");
  const int rho = P.rho_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1;
  const double* dx = P.dx;
  const Arr &um = umac[0], &vm = umac[1], &wm = umac[2];
  for_box(grown(lo, hi, 3, 0), [&](int i, int j, int k) {
    EosState st{};
    double xn[MGPU_EOS_MAXSPEC];
    st.T = s(i, j, k, temp);
    st.rho = s(i, j, k, rho);
    load_xn(s, i, j, k, spec0, E.nspec, s(i, j, k, rho), true, xn);
    eos_call(MGPU_EOS_INPUT_RT, st, xn);
    const double sr = s(i, j, k, rho);
    const double dhdp = 1.0 / sr + (sr * st.dedr - st.p / sr) / (sr * st.dpdr);
    const double p0_lox = 0.5 * (p0_cart(i, j, k) + p0_cart(i - 1, j, k));
    const double p0_hix = 0.5 * (p0_cart(i, j, k) + p0_cart(i + 1, j, k));
    const double p0_loy = 0.5 * (p0_cart(i, j, k) + p0_cart(i, j - 1, k));
    const double p0_hiy = 0.5 * (p0_cart(i, j, k) + p0_cart(i, j + 1, k));
    const double p0_loz = 0.5 * (p0_cart(i, j, k) + p0_cart(i, j, k - 1));
    const double p0_hiz = 0.5 * (p0_cart(i, j, k) + p0_cart(i, j, k + 1));
    const double divup = (um(i + 1, j, k) * p0_hix - um(i, j, k) * p0_lox) / dx[0] +
                         (vm(i, j + 1, k) * p0_hiy - vm(i, j, k) * p0_loy) / dx[1] +
                         (wm(i, j, k + 1) * p0_hiz - wm(i, j, k) * p0_loz) / dx[2];
    const double p0divu = ((um(i + 1, j, k) - um(i, j, k)) / dx[0] + (vm(i, j + 1, k) - vm(i, j, k)) / dx[1] +
                           (wm(i, j, k + 1) - wm(i, j, k)) / dx[2]) * p0_cart(i, j, k);
    const double ugradp = divup - p0divu;
    double f = thermal(i, j, k) + (1.0 - sr * dhdp) * (ugradp + psi_cart(i, j, k));
    f = f / (st.cp * sr);
    temp_force(i, j, k, temp) = f;
  });
}

// firstdt_2d / _3d / _3d_sphr (firstdt.f90:330, :460, :599) of one box; gp0_cart (sphr): the vector field of :735
void firstdt_box(const mgpu_params& P, const Arr& u, const Arr& s, const Arr& force, const Arr& divU, const double* p0,
                 const double* gamma1bar, const Arr* gp0_cart, const int* lo, const int* hi, double cfl,
                 bool use_soundspeed_firstdt, bool use_divu_firstdt, double& dt, double& umax) {
  const mgpu_eos& E = the_eos("firstdt");
  const int dm = P.dm, rho = P.rho_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1, nr = P.nr;
  const double eps = 1.0e-8, rho_min = 1.e-20;
  double spd = 0.0, pforce[3] = {0.0, 0.0, 0.0}, uu[3] = {0.0, 0.0, 0.0};
  dt = 1.e99;
  umax = 0.0;
  const Box vb = grown(lo, hi, dm, 0);
  for (int k = vb.lo[2]; k <= vb.hi[2]; ++k)
    for (int j = vb.lo[1]; j <= vb.hi[1]; ++j)
      for (int i = vb.lo[0]; i <= vb.hi[0]; ++i) {
        EosState st{};
        double xn[MGPU_EOS_MAXSPEC];
        st.rho = s(i, j, k, rho);
        st.T = s(i, j, k, temp);
        load_xn(s, i, j, k, spec0, E.nspec, st.rho, true, xn);
        eos_call(MGPU_EOS_INPUT_RT, st, xn);
        spd = dmax(spd, st.cs);
        for (int d = 0; d < dm; ++d) {
          pforce[d] = dmax(pforce[d], dabs(force(i, j, k, d)));
          uu[d] = dmax(uu[d], dabs(u(i, j, k, d)));
        }
      }
  for (int d = 0; d < dm; ++d) umax = dmax(umax, uu[d]);
  double spdd[3], um = 0.0, sm = 0.0;
  bool any_u = false, all_spd = true, no_spd = true;
  for (int d = 0; d < dm; ++d) {
    uu[d] = uu[d] / P.dx[d];
    spdd[d] = spd / P.dx[d];
    if (uu[d] != 0.0) any_u = true;
    if (spdd[d] == 0.0) all_spd = false; else no_spd = false;
    um = dmax(um, uu[d]);
    sm = dmax(sm, spdd[d]);
  }
  if (any_u) dt = cfl / um;
  else if (all_spd) dt = cfl / sm;
  if (use_soundspeed_firstdt) {
    const double dt_sound = no_spd ? 1.e99 : cfl / sm;
    dt = dmin(dt, dt_sound);
  }
  for (int d = 0; d < dm; ++d)
    if (pforce[d] > eps) dt = dmin(dt, std::sqrt(2.0 * P.dx[d] / pforce[d]));
  if (use_divu_firstdt) {
    double dt_divu = 1.e99;
    for (int k = vb.lo[2]; k <= vb.hi[2]; ++k)
      for (int j = vb.lo[1]; j <= vb.hi[1]; ++j)
        for (int i = vb.lo[0]; i <= vb.hi[0]; ++i) {
          double denom;
          if (gp0_cart) {
            const double gp_dot_u = u(i, j, k, 0) * (*gp0_cart)(i, j, k, 0) + u(i, j, k, 1) * (*gp0_cart)(i, j, k, 1) +
                                    u(i, j, k, 2) * (*gp0_cart)(i, j, k, 2);
            denom = divU(i, j, k) - gp_dot_u;
          } else {
            const int r = (dm == 3) ? k : j;
            const double h = P.dx[dm - 1];
            double gradp0;
            if (r == 0) gradp0 = (p0[r + 1] - p0[r]) / h;
            else if (r == nr - 1) gradp0 = (p0[r] - p0[r - 1]) / h;
            else gradp0 = 0.5 * (p0[r + 1] - p0[r - 1]) / h;
            denom = divU(i, j, k) - u(i, j, k, dm - 1) * gradp0 / (gamma1bar[r] * p0[r]);
          }
          if (denom > 0.0) dt_divu = dmin(dt_divu, 0.4 * (1.0 - rho_min / s(i, j, k, rho)) / denom);
        }
    dt = dmin(dt, dt_divu);
  }
}

// makeTfromRhoH_2d/_3d/_3d_sphr (rhoh_vs_t.f90:929, :999, :1077) and makeTfromRhoP (:1287, :1336, :1390): valid cells
void make_t_box(const mgpu_params& P, Arr& state, const double* p0, const Arr* p0_cart, bool from_p, bool flag,
                bool update_rhoh, const int* lo, const int* hi) {
  const mgpu_eos& E = the_eos(from_p ? "makeTfromRhoP" : "makeTfromRhoH");
  const int dm = P.dm, rho = P.rho_comp - 1, rhoh = P.rhoh_comp - 1, temp = P.temp_comp - 1, spec0 = P.spec_comp - 1;
  const int pi = P.pi_comp - 1;
  for_box(grown(lo, hi, dm, 0), [&](int i, int j, int k) {
    const double p0v = p0_cart ? (*p0_cart)(i, j, k) : p0[dm == 3 ? k : j];
    EosState st{};
    double xn[MGPU_EOS_MAXSPEC];
    st.rho = state(i, j, k, rho);
    st.T = state(i, j, k, temp);
    load_xn(state, i, j, k, spec0, E.nspec, st.rho, true, xn);
    if (from_p) {
      st.p = flag ? p0v + state(i, j, k, pi) : p0v;  // use_pprime_in_tfromp
      eos_call(MGPU_EOS_INPUT_RP, st, xn);
      state(i, j, k, temp) = st.T;
      if (update_rhoh) state(i, j, k, rhoh) = st.rho * st.h;
    } else if (flag) {  // use_eos_e_instead_of_h
      st.e = state(i, j, k, rhoh) / state(i, j, k, rho) - p0v / state(i, j, k, rho);
      eos_call(MGPU_EOS_INPUT_RE, st, xn);
      state(i, j, k, temp) = st.T;
    } else {
      st.h = state(i, j, k, rhoh) / state(i, j, k, rho);
      eos_call(MGPU_EOS_INPUT_RH, st, xn);
      state(i, j, k, temp) = st.T;
    }
  });
}

}  // namespace mo
