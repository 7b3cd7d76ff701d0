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

}  // namespace mgpu
