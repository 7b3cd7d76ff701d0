// Reductions next to the advective path (see mgpu_reduce.cu).
#pragma once
#include <algorithm>
#include <cmath>

#include "mgpu_common.cuh"

namespace mgpu {

// estdt_2d / estdt_3d_cart (Source/estdt.f90:348, :467) of one box: *dt in/out (the reference enters with HUGE),
// *umax out.  w0: device copy, w0_h: the caller's host array (the spdr loop over the box's radial range)
void estdt_box_dev(const mgpu_params& P, const DV& u, const DV& s, const DV& force, const DV& divU, const DV& dSdt,
                   const double* w0, const double* w0_h, const double* p0, const double* gamma1bar, const int* lo,
                   const int* hi, double rho_min, double cfl, double* dt, double* umax, const DV* w0mac = nullptr,
                   const DV* gp0_cart = nullptr, double dr = 0.0, int nr_fine = 0);
// (w0mac != nullptr: estdt_3d_sphr, estdt.f90:620 -- w0 on the faces, gp0_cart = grad(p0)/(gamma1bar p0) as a vector)
// sums of the slab-direction planes k0..k1 of a single-component fab over the valid transverse cells -> host
void plane_sums_dev(const mgpu_params& P, const DV& f, const int* lo, const int* hi, int k0, int k1, double* sums_h);

// min / max of component comp (0-based; divided by component dcomp when dcomp >= 0) over the valid cells of one box,
// folded into *mn / *mx
void minmax_box_dev(const mgpu_params& P, const DV& s, const int* lo, const int* hi, int comp, int dcomp, double* mn,
                    double* mx);

// average() with spherical == 1 (average.f90:168-362), one level: the binning of one box (atomics into phisum / ncell,
// nr_irreg + 1 bins, radii(0:nr_irreg+1) on the device) and the host tail (normalise, drop the empty radii, interpolate)
void sum_phi_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const DV& phi1, const int* lo, const int* hi,
                      const double* radii_dev, int nr_irreg, double* phisum_dev, unsigned long long* ncell_dev);
void average_sphr_tail(double dr, int nr_fine, int nr_irreg, int drdxfac, std::vector<double>& phisum,
                       std::vector<long>& ncell, std::vector<double>& radii, double* phibar);
// construct_eta_cart (make_eta.f90:345): eta and rho0_nph_cart cover the valid cells
void eta_cart_dev(const DV& eta, const DV& rho_old, const DV& rho_new, const DV* umac, const DV* w0mac, const DV& normal,
                  const DV& rho0_nph_cart, const int* lo, const int* hi);

}  // namespace mgpu
