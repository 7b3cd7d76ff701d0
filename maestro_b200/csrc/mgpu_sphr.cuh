// Spherical-geometry operators (see mgpu_sphr.cu).
#pragma once
#include "mgpu_common.cuh"
#include "mgpu_stream.cuh"

namespace mgpu {

// device copy of mgpu_geom (+ dx): the two radial coordinate arrays live in the arena for the duration of a call
struct Geom {
  double center[3], prob_lo[3], dx[3];
  double dr;
  double rdr;  // 1 / dr
  int fast;    // interpolation with multiplications by 1/dr (see quad_interp_fast); 0 in the exact build
  int nr_fine;
  const double* r_cc_loc;
  const double* r_edge_loc;
};
Geom make_geom(const mgpu_params& P, const mgpu_geom& g);
void sphr_set_fast(int on);  // 1 (default): Geom::fast = 1 for every geometry made afterwards

struct SphrFluxArgs {
  int spt, rho, rhoh;
  Box3 vb;
  DV sflux[3], sedge[3], umac[3], w0mac[3], r0o[3], r0n[3], h0o[3], h0n[3];
};

void put_1d_array_on_cart_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const double* s0_dev,
                              const DV& cart, bool edge_in, bool vec, const int* lo, const int* hi);
// kind 0: make_w0mac, kind 1: make_s0mac
void make_mac_dev(const mgpu_geom& g, const Geom& gd, const double* s0_dev, DV* mac, const DV* cart, int kind,
                  const int* lo, const int* hi);
// utot (same layout as umac) = umac + w0mac on the valid faces, umac on the ghost faces
void sum_faces_sphr_dev(DV* utot, const DV* umac, const DV* w0mac, const int* lo, const int* hi);
void addw0_sphr_dev(DV* umac, const DV* w0mac, double mult, const int* lo, const int* hi);
void mk_rhoX_flux_sphr_dev(SphrFluxArgs& a, int startcomp, int endcomp);
void mk_rhoh_flux_sphr_dev(const mgpu_params& P, SphrFluxArgs& a);
void update_velocity_sphr_dev(VelArgs& a, const DV* w0mac);
void modify_scal_force_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const DV& force, const DV& s,
                                const DV* umac, const DV& s0_cart, const double* w0_host, int comp, bool fullform,
                                const int* lo, const int* hi);
void pert_form_sphr_dev(const mgpu_geom& g, const Geom& gd, const DV& s, const double* s0_dev, int comp, bool flag,
                        const int* lo, const int* hi);

void mkrhohforce_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const DV& force, bool is_prediction,
                          const DV& thermal, const DV* umac, const DV& p0c, const DV* p0mac, const double* psi_h,
                          bool add_thermal, const int* lo, const int* hi);
void make_normal_dev(const Geom& gd, const DV& normal, const int* lo, const int* hi, int ng);
void mk_vel_force_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, const DV& force, bool is_final,
                           const DV& uold, const DV* uedge, const double* w0_h, const DV* w0mac, const DV& gpi,
                           const DV& rho1, const DV& normal, const double* rho0_h, const double* grav_h,
                           const DV& w0_force_cart, const int* lo, const int* hi, bool add_utilde);

}  // namespace mgpu
