! maestro_b200_shim.f90 -- ISO_C_BINDING shim between MAESTRO's multifab-level (L3) operators and
! libmaestro_b200.so (include/maestro_b200.h).
!
! How it is used (see INTEGRATION.md): this file is added to Source/ (GPackage.mak), and the bodies
! of the public L3 routines -- make_edge_scal (Source/make_edge_scal.f90:26), bds (bds.f90:16),
! mkutrans (mkutrans.f90:17), velpred (velpred.f90:21), mk_rhoX_flux / mk_rhoh_flux
! (mkflux.f90:48,652), update_scal (update_scal.f90:16), update_velocity (update_vel.f90:15),
! addw0 (addw0.f90:19) -- replace their `do i=1,nfabs ... call <kernel>_3d(...)` loops by one call
! of the matching mgpu_* wrapper below.  Public names and argument lists of the L3 routines and of
! the L4 drivers (advance_premac, density_advance, enthalpy_advance, velocity_advance) do not change.
!
! NOT compiled in the build container (no Fortran compiler, FBoxLib not vendored); it is written
! against the FBoxLib API the reference uses (dataptr, get_box, lwb, upb, nghost, ncomp, nfabs,
! nodal_flags) and the C structs of include/maestro_b200.h, field for field.
module maestro_b200_shim

  use iso_c_binding
  use bl_types
  use bl_error_module
  use multifab_module
  use define_bc_module

  implicit none
  private

  ! ---- mirrors of the C structs (include/maestro_b200.h) -------------------------------------
  type, bind(C), public :: mgpu_fab
     type(c_ptr)    :: ptr
     integer(c_int) :: lo(3), hi(3)
     integer(c_int) :: ng, nc
     integer(c_int) :: nodal(3)
  end type mgpu_fab

  type, bind(C), public :: mgpu_params
     integer(c_int) :: dm, mem_space, ppm_type, bds_type, slope_order, ppm_trace_forces
     integer(c_int) :: species_pred_type, enthalpy_pred_type, spherical, evolve_base_state
     integer(c_int) :: do_sponge, do_eos_h_above_cutoff
     integer(c_int) :: rho_comp, rhoh_comp, spec_comp, temp_comp, pi_comp, trac_comp
     integer(c_int) :: nspec, ntrac, nscal
     integer(c_int) :: domlo(3), domhi(3)
     integer(c_int) :: nr
     real(c_double) :: dt, dx(3), rel_eps, base_cutoff_density
     integer(c_int) :: base_cutoff_density_coord
     real(c_double) :: buoyancy_cutoff_factor, omega, sin_theta, cos_theta, rotation_radius
  end type mgpu_params

  ! spherical geometry (geometry module + probin interpolation switches), include/maestro_b200.h mgpu_geom
  type, bind(C), public :: mgpu_geom
     real(c_double) :: center(3), prob_lo(3), dr
     integer(c_int) :: nr_fine
     type(c_ptr)    :: r_cc_loc, r_edge_loc      ! c_loc(r_cc_loc(1,0)), c_loc(r_edge_loc(1,0))
     integer(c_int) :: s0_interp_type, w0_interp_type, s0mac_interp_type, w0mac_interp_type
  end type mgpu_geom

  ! the EOS state of the library (include/maestro_b200.h mgpu_eos): what eos_init leaves in the module variables of
  ! eos_type_module / actual_eos_module, plus the network's aion / zion and the constants of constants_cgs.f90
  type, bind(C), public :: mgpu_eos
     integer(c_int) :: kind, assume_neutral, nspec, pad_
     real(c_double) :: gamma, k_B, n_A
     real(c_double) :: mintemp, maxtemp, mindens, maxdens, mine, maxe, minp, maxp, minh, maxh
     real(c_double) :: small_temp
     real(c_double) :: aion(32), zion(32)
  end type mgpu_eos

  integer(c_int), parameter :: MGPU_HOST = 0

  interface
     integer(c_int) function mgpu_init(device) bind(C, name="mgpu_init")
       import :: c_int
       integer(c_int), value :: device
     end function mgpu_init
     integer(c_int) function mgpu_finalize() bind(C, name="mgpu_finalize")
       import :: c_int
     end function mgpu_finalize
     type(c_ptr) function mgpu_last_error() bind(C, name="mgpu_last_error")
       import :: c_ptr
     end function mgpu_last_error
     ! slab runs: rank 0 makes the NCCL id, every rank joins (INTEGRATION.md, section 2)
     integer(c_int) function mgpu_comm_unique_id(id128) bind(C, name="mgpu_comm_unique_id")
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id128(128)
     end function mgpu_comm_unique_id
     integer(c_int) function mgpu_comm_init(rank, nranks, id128) bind(C, name="mgpu_comm_init")
       import :: c_int, c_char
       integer(c_int), value :: rank, nranks
       character(kind=c_char), intent(in) :: id128(128)
     end function mgpu_comm_init
     integer(c_int) function mgpu_comm_finalize() bind(C, name="mgpu_comm_finalize")
       import :: c_int
     end function mgpu_comm_finalize
     integer(c_int) function mgpu_set_option(key, value) bind(C, name="mgpu_set_option")
       import :: c_int, c_char
       character(kind=c_char), intent(in) :: key(*)   ! null-terminated, e.g. "overlap"//c_null_char
       integer(c_int), value :: value
     end function mgpu_set_option
     ! device residency (params%mem_space = MGPU_DEVICE): device buffers of n doubles and copies
     integer(c_int) function mgpu_malloc(dptr, n) bind(C, name="mgpu_malloc")
       import :: c_int, c_long, c_ptr
       type(c_ptr), intent(out) :: dptr
       integer(c_long), value :: n
     end function mgpu_malloc
     integer(c_int) function mgpu_free(dptr) bind(C, name="mgpu_free")
       import :: c_int, c_ptr
       type(c_ptr), value :: dptr
     end function mgpu_free
     integer(c_int) function mgpu_memcpy_h2d(dst, src, n) bind(C, name="mgpu_memcpy_h2d")
       import :: c_int, c_long, c_ptr
       type(c_ptr), value :: dst, src
       integer(c_long), value :: n
     end function mgpu_memcpy_h2d
     integer(c_int) function mgpu_memcpy_d2h(dst, src, n) bind(C, name="mgpu_memcpy_d2h")
       import :: c_int, c_long, c_ptr
       type(c_ptr), value :: dst, src
       integer(c_long), value :: n
     end function mgpu_memcpy_d2h
     integer(c_int) function mgpu_host_register(hptr, n) bind(C, name="mgpu_host_register")
       import :: c_int, c_long, c_ptr
       type(c_ptr), value :: hptr
       integer(c_long), value :: n
     end function mgpu_host_register
     ! residency registry: the multifabs of a step stay on the device between episodes (maestro_b200.h)
     integer(c_int) function mgpu_register(hptr, n, pin) bind(C, name="mgpu_register")
       import :: c_int, c_long, c_ptr
       type(c_ptr), value :: hptr
       integer(c_long), value :: n
       integer(c_int), value :: pin
     end function mgpu_register
     integer(c_int) function mgpu_unregister(hptr) bind(C, name="mgpu_unregister")
       import :: c_int, c_ptr
       type(c_ptr), value :: hptr
     end function mgpu_unregister
     integer(c_int) function mgpu_invalidate(hptr, comp0, ncomp) bind(C, name="mgpu_invalidate")
       import :: c_int, c_ptr
       type(c_ptr), value :: hptr
       integer(c_int), value :: comp0, ncomp
     end function mgpu_invalidate
     integer(c_int) function mgpu_download(hptr, comp0, ncomp) bind(C, name="mgpu_download")
       import :: c_int, c_ptr
       type(c_ptr), value :: hptr
       integer(c_int), value :: comp0, ncomp
     end function mgpu_download
     integer(c_int) function mgpu_upload(hptr, comp0, ncomp) bind(C, name="mgpu_upload")
       import :: c_int, c_ptr
       type(c_ptr), value :: hptr
       integer(c_int), value :: comp0, ncomp
     end function mgpu_upload

     integer(c_int) function mgpu_make_edge_scal_c(p, nfabs, s, sedge, umac, force, adv_bc, is_vel, &
          start_scomp, start_bccomp, num_comp, is_conservative) bind(C, name="mgpu_make_edge_scal")
       import :: c_int, c_ptr, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, is_vel, start_scomp, start_bccomp, num_comp, is_conservative
       type(mgpu_fab), intent(in) :: s(*), force(*)
       type(c_ptr), intent(in) :: sedge(*), umac(*)      ! dm pointers to arrays of nfabs mgpu_fab
       integer(c_int), intent(in) :: adv_bc(*)
     end function mgpu_make_edge_scal_c

     integer(c_int) function mgpu_bds_c(p, nfabs, s, sedge, umac, force, adv_bc, is_vel, &
          start_scomp, start_bccomp, num_comp, is_conservative) bind(C, name="mgpu_bds")
       import :: c_int, c_ptr, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, is_vel, start_scomp, start_bccomp, num_comp, is_conservative
       type(mgpu_fab), intent(in) :: s(*), force(*)
       type(c_ptr), intent(in) :: sedge(*), umac(*)
       integer(c_int), intent(in) :: adv_bc(*)
     end function mgpu_bds_c

     integer(c_int) function mgpu_mk_rhoX_flux_c(p, nfabs, sflux, etarhoflux, sedge, umac, w0, rho0_old, &
          rho0_edge_old, rho0_new, rho0_edge_new, rho0_predicted_edge, startcomp, endcomp) &
          bind(C, name="mgpu_mk_rhoX_flux")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, startcomp, endcomp
       type(c_ptr), intent(in) :: sflux(*), sedge(*), umac(*)
       type(mgpu_fab), intent(in) :: etarhoflux(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), rho0_edge_old(*), rho0_new(*), rho0_edge_new(*), &
                                     rho0_predicted_edge(*)
     end function mgpu_mk_rhoX_flux_c

     integer(c_int) function mgpu_mk_rhoh_flux_c(p, nfabs, sflux, sedge, umac, w0, rho0_old, rho0_edge_old, &
          rho0_new, rho0_edge_new, rhoh0_old, rhoh0_edge_old, rhoh0_new, rhoh0_edge_new) &
          bind(C, name="mgpu_mk_rhoh_flux")
       import :: c_int, c_ptr, c_double, mgpu_params
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(c_ptr), intent(in) :: sflux(*), sedge(*), umac(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), rho0_edge_old(*), rho0_new(*), rho0_edge_new(*), &
                                     rhoh0_old(*), rhoh0_edge_old(*), rhoh0_new(*), rhoh0_edge_new(*)
     end function mgpu_mk_rhoh_flux_c

     ! p0_new / p0_new_cart: c_loc of the level's 1-D array / of an array of nfabs mgpu_fab (c_null_ptr when unused)
     integer(c_int) function mgpu_update_scal_c(p, nfabs, nstart, nstop, sold, snew, sflux, force, p0_new, p0_new_cart) &
          bind(C, name="mgpu_update_scal")
       import :: c_int, c_ptr, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, nstart, nstop
       type(mgpu_fab), intent(in) :: sold(*), snew(*), force(*)
       type(c_ptr), intent(in) :: sflux(*)
       type(c_ptr), value :: p0_new, p0_new_cart
     end function mgpu_update_scal_c

     integer(c_int) function mgpu_update_velocity_c(p, nfabs, uold, unew, umac, uedge, force, sponge, w0) &
          bind(C, name="mgpu_update_velocity")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: uold(*), unew(*), force(*), sponge(*)
       type(c_ptr), intent(in) :: umac(*), uedge(*)
       real(c_double), intent(in) :: w0(*)
     end function mgpu_update_velocity_c

     integer(c_int) function mgpu_addw0_c(p, nfabs, umac, w0, mult) bind(C, name="mgpu_addw0")
       import :: c_int, c_ptr, c_double, mgpu_params
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*)
       real(c_double), value :: mult
     end function mgpu_addw0_c

     integer(c_int) function mgpu_mkutrans_c(p, nfabs, utilde, ufull, utrans, w0, adv_bc, phys_bc) &
          bind(C, name="mgpu_mkutrans")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: utilde(*), ufull(*)
       type(c_ptr), intent(in) :: utrans(*)
       real(c_double), intent(in) :: w0(*)
       integer(c_int), intent(in) :: adv_bc(*), phys_bc(*)
     end function mgpu_mkutrans_c

     integer(c_int) function mgpu_velpred_c(p, nfabs, utilde, ufull, umac, utrans, force, w0, adv_bc, phys_bc) &
          bind(C, name="mgpu_velpred")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: utilde(*), ufull(*), force(*)
       type(c_ptr), intent(in) :: umac(*), utrans(*)
       real(c_double), intent(in) :: w0(*)
       integer(c_int), intent(in) :: adv_bc(*), phys_bc(*)
     end function mgpu_velpred_c

     ! ---- spherical geometry: the *_3d_sphr branches (fill_3d_data.f90:269,621,1017; addw0.f90:171;
     !      mkflux.f90:509,1289; update_vel.f90:317; mkutrans.f90:601; velpred.f90:1588;
     !      modify_scal_force.f90:256; put_in_pert_form.f90:185)
     integer(c_int) function mgpu_put_1d_array_on_cart_c(p, g, nfabs, s0, s0_cart, is_input_edge_centered, &
          is_output_a_vector) bind(C, name="mgpu_put_1d_array_on_cart")
       import :: c_int, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs, is_input_edge_centered, is_output_a_vector
       real(c_double), intent(in) :: s0(*)
       type(mgpu_fab), intent(in) :: s0_cart(*)
     end function mgpu_put_1d_array_on_cart_c

     integer(c_int) function mgpu_make_w0mac_c(p, g, nfabs, w0, w0mac, w0_cart) bind(C, name="mgpu_make_w0mac")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs
       real(c_double), intent(in) :: w0(*)
       type(c_ptr), intent(in) :: w0mac(*)
       type(mgpu_fab), intent(in) :: w0_cart(*)
     end function mgpu_make_w0mac_c

     integer(c_int) function mgpu_make_s0mac_c(p, g, nfabs, s0, s0mac, s0_cart) bind(C, name="mgpu_make_s0mac")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs
       real(c_double), intent(in) :: s0(*)
       type(c_ptr), intent(in) :: s0mac(*)
       type(mgpu_fab), intent(in) :: s0_cart(*)
     end function mgpu_make_s0mac_c

     integer(c_int) function mgpu_addw0_sphr_c(p, nfabs, umac, w0mac, mult) bind(C, name="mgpu_addw0_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(c_ptr), intent(in) :: umac(*), w0mac(*)
       real(c_double), value :: mult
     end function mgpu_addw0_sphr_c

     integer(c_int) function mgpu_mk_rhoX_flux_sphr_c(p, nfabs, sflux, sedge, umac, w0mac, rho0mac_old, &
          rho0mac_new, startcomp, endcomp) bind(C, name="mgpu_mk_rhoX_flux_sphr")
       import :: c_int, c_ptr, mgpu_params
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, startcomp, endcomp
       type(c_ptr), intent(in) :: sflux(*), sedge(*), umac(*), w0mac(*), rho0mac_old(*), rho0mac_new(*)
     end function mgpu_mk_rhoX_flux_sphr_c

     integer(c_int) function mgpu_mk_rhoh_flux_sphr_c(p, nfabs, sflux, sedge, umac, w0mac, rho0mac_old, &
          rho0mac_new, h0mac_old, h0mac_new) bind(C, name="mgpu_mk_rhoh_flux_sphr")
       import :: c_int, c_ptr, mgpu_params
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(c_ptr), intent(in) :: sflux(*), sedge(*), umac(*), w0mac(*), rho0mac_old(*), rho0mac_new(*)
       type(c_ptr), intent(in) :: h0mac_old(*), h0mac_new(*)
     end function mgpu_mk_rhoh_flux_sphr_c

     integer(c_int) function mgpu_update_velocity_sphr_c(p, nfabs, uold, unew, umac, uedge, force, sponge, w0mac) &
          bind(C, name="mgpu_update_velocity_sphr")
       import :: c_int, c_ptr, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: uold(*), unew(*), force(*), sponge(*)
       type(c_ptr), intent(in) :: umac(*), uedge(*), w0mac(*)
     end function mgpu_update_velocity_sphr_c

     integer(c_int) function mgpu_mkutrans_sphr_c(p, nfabs, utilde, ufull, utrans, w0mac, adv_bc, phys_bc) &
          bind(C, name="mgpu_mkutrans_sphr")
       import :: c_int, c_ptr, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: utilde(*), ufull(*)
       type(c_ptr), intent(in) :: utrans(*), w0mac(*)
       integer(c_int), intent(in) :: adv_bc(*), phys_bc(*)
     end function mgpu_mkutrans_sphr_c

     integer(c_int) function mgpu_velpred_sphr_c(p, nfabs, utilde, ufull, umac, utrans, force, w0mac, adv_bc, &
          phys_bc) bind(C, name="mgpu_velpred_sphr")
       import :: c_int, c_ptr, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: utilde(*), ufull(*), force(*)
       type(c_ptr), intent(in) :: umac(*), utrans(*), w0mac(*)
       integer(c_int), intent(in) :: adv_bc(*), phys_bc(*)
     end function mgpu_velpred_sphr_c

     integer(c_int) function mgpu_modify_scal_force_sphr_c(p, g, nfabs, force, s, umac, s0_cart, w0, comp, &
          fullform) bind(C, name="mgpu_modify_scal_force_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs, comp, fullform
       type(mgpu_fab), intent(in) :: force(*), s(*), s0_cart(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*)
     end function mgpu_modify_scal_force_sphr_c

     integer(c_int) function mgpu_put_in_pert_form_sphr_c(p, g, nfabs, s, s0, comp, flag) &
          bind(C, name="mgpu_put_in_pert_form_sphr")
       import :: c_int, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs, comp, flag
       type(mgpu_fab), intent(in) :: s(*)
       real(c_double), intent(in) :: s0(*)
     end function mgpu_put_in_pert_form_sphr_c

     ! multifab_fill_boundary + multifab_physbc (multifab_physbc.f90:16)
     integer(c_int) function mgpu_fill_boundary_c(p, s, scomp, bccomp, ncomp, adv_bc, pmask) &
          bind(C, name="mgpu_fill_boundary")
       import :: c_int, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_fab), intent(inout) :: s(*)
       integer(c_int), value :: scomp
       integer(c_int), value :: bccomp
       integer(c_int), value :: ncomp
       integer(c_int), intent(in) :: adv_bc(*)
       integer(c_int), intent(in) :: pmask(*)
     end function mgpu_fill_boundary_c

     ! convert_rhoX_to_X.f90:20
     integer(c_int) function mgpu_convert_rhoX_to_X_c(p, nfabs, s, flag) &
          bind(C, name="mgpu_convert_rhoX_to_X")
       import :: c_int, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: s(*)
       integer(c_int), value :: flag
     end function mgpu_convert_rhoX_to_X_c

     ! modify_scal_force.f90:15
     integer(c_int) function mgpu_modify_scal_force_c(p, nfabs, force, s, umac, s0, s0_edge, w0, comp, &
          fullform) &
          bind(C, name="mgpu_modify_scal_force")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: force(*)
       type(mgpu_fab), intent(in) :: s(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: s0(*)
       real(c_double), intent(in) :: s0_edge(*)
       real(c_double), intent(in) :: w0(*)
       integer(c_int), value :: comp
       integer(c_int), value :: fullform
     end function mgpu_modify_scal_force_c

     ! put_in_pert_form.f90:22
     integer(c_int) function mgpu_put_in_pert_form_c(p, nfabs, s, base, comp, flag) &
          bind(C, name="mgpu_put_in_pert_form")
       import :: c_int, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: s(*)
       real(c_double), intent(in) :: base(*)
       integer(c_int), value :: comp
       integer(c_int), value :: flag
     end function mgpu_put_in_pert_form_c

     ! mkscalforce.f90:31
     integer(c_int) function mgpu_mkrhohforce_c(p, nfabs, scal_force, is_prediction, thermal, umac, p0_1, p0_2, &
          rho0_1, rho0_2, grav, psi, add_thermal) &
          bind(C, name="mgpu_mkrhohforce")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: scal_force(*)
       integer(c_int), value :: is_prediction
       type(mgpu_fab), intent(in) :: thermal(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: p0_1(*)
       real(c_double), intent(in) :: p0_2(*)
       real(c_double), intent(in) :: rho0_1(*)
       real(c_double), intent(in) :: rho0_2(*)
       real(c_double), intent(in) :: grav(*)
       real(c_double), intent(in) :: psi(*)
       integer(c_int), value :: add_thermal
     end function mgpu_mkrhohforce_c

     ! mkforce.f90:22
     integer(c_int) function mgpu_mk_vel_force_c(p, nfabs, vel_force, is_final_update, uold, uedge, w0, gpi, s, &
          index_rho, rho0, grav, w0_force, do_add_utilde_force) &
          bind(C, name="mgpu_mk_vel_force")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: vel_force(*)
       integer(c_int), value :: is_final_update
       type(mgpu_fab), intent(in) :: uold(*)
       type(c_ptr), intent(in) :: uedge(*)
       real(c_double), intent(in) :: w0(*)
       type(mgpu_fab), intent(in) :: gpi(*)
       type(mgpu_fab), intent(in) :: s(*)
       integer(c_int), value :: index_rho
       real(c_double), intent(in) :: rho0(*)
       real(c_double), intent(in) :: grav(*)
       real(c_double), intent(in) :: w0_force(*)
       integer(c_int), value :: do_add_utilde_force
     end function mgpu_mk_vel_force_c

     ! density_advance.f90:20, the whole episode on the device
     integer(c_int) function mgpu_density_advance_c(p, which_step, sold, snew, sedge, sflux, scal_force, umac, &
          w0, etarhoflux, rho0_old, rho0_new, p0_dummy, rho0_predicted_edge, adv_bc, pmask) &
          bind(C, name="mgpu_density_advance")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: which_step
       type(mgpu_fab), intent(inout) :: sold(*)
       type(mgpu_fab), intent(inout) :: snew(*)
       type(c_ptr), intent(in) :: sedge(*)
       type(c_ptr), intent(in) :: sflux(*)
       type(mgpu_fab), intent(inout) :: scal_force(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*)
       type(mgpu_fab), intent(inout) :: etarhoflux(*)
       real(c_double), intent(in) :: rho0_old(*)
       real(c_double), intent(in) :: rho0_new(*)
       real(c_double), intent(in) :: p0_dummy(*)
       real(c_double), intent(in) :: rho0_predicted_edge(*)
       integer(c_int), intent(in) :: adv_bc(*)
       integer(c_int), intent(in) :: pmask(*)
     end function mgpu_density_advance_c

     ! density_advance.f90:20 over the nfabs boxes of this rank's multifab (adv_bc: the domain's table)
     integer(c_int) function mgpu_density_advance_mf_c(p, which_step, nfabs, sold, snew, sedge, sflux, scal_force, &
          umac, w0, etarhoflux, rho0_old, rho0_new, p0_dummy, rho0_predicted_edge, adv_bc, pmask) &
          bind(C, name="mgpu_density_advance_mf")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: which_step, nfabs
       type(mgpu_fab), intent(inout) :: sold(*), snew(*), scal_force(*), etarhoflux(*)
       type(c_ptr), intent(in) :: sedge(*), sflux(*), umac(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), rho0_new(*), p0_dummy(*), rho0_predicted_edge(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_density_advance_mf_c

     ! velocity_advance.f90:16 / advance_premac.f90:21 / enthalpy_advance.f90:16 over the nfabs boxes of this rank
     integer(c_int) function mgpu_velocity_advance_mf_c(p, nfabs, uold, unew, sold, rhohalf, umac, gpi, w0, w0_force, &
          rho0_old, rho0_nph, grav_cell_old, grav_cell_nph, sponge, adv_bc, pmask) bind(C, name="mgpu_velocity_advance_mf")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: uold(*), sold(*), rhohalf(*), gpi(*), sponge(*)
       type(mgpu_fab), intent(inout) :: unew(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*), w0_force(*), rho0_old(*), rho0_nph(*), grav_cell_old(*), grav_cell_nph(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_velocity_advance_mf_c

     integer(c_int) function mgpu_advance_premac_mf_c(p, nfabs, uold, sold, umac, gpi, w0, w0_force, rho0_old, &
          grav_cell_old, adv_bc, phys_bc, pmask) bind(C, name="mgpu_advance_premac_mf")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: uold(*), sold(*), gpi(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*), w0_force(*), rho0_old(*), grav_cell_old(*)
       integer(c_int), intent(in) :: adv_bc(*), phys_bc(*), pmask(*)
     end function mgpu_advance_premac_mf_c

     integer(c_int) function mgpu_enthalpy_advance_mf_c(p, which_step, nfabs, sold, snew, sedge, sflux, scal_force, &
          thermal, umac, w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, psi, grav_old, grav_nph, adv_bc, &
          pmask) bind(C, name="mgpu_enthalpy_advance_mf")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: which_step, nfabs
       type(mgpu_fab), intent(inout) :: sold(*), snew(*), scal_force(*)
       type(mgpu_fab), intent(in) :: thermal(*)
       type(c_ptr), intent(in) :: sedge(*), sflux(*), umac(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), rhoh0_old(*), rho0_new(*), rhoh0_new(*), p0_old(*), p0_new(*)
       real(c_double), intent(in) :: psi(*), grav_old(*), grav_nph(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_enthalpy_advance_mf_c

     ! multifab_fill_boundary + multifab_physbc over the nfabs boxes of this rank's multifab
     integer(c_int) function mgpu_fill_boundary_mf_c(p, nfabs, s, scomp, bccomp, ncomp, adv_bc, pmask) &
          bind(C, name="mgpu_fill_boundary_mf")
       import :: c_int, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, scomp, bccomp, ncomp
       type(mgpu_fab), intent(inout) :: s(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_fill_boundary_mf_c

     ! mkscalforce.f90:31 with spherical == 1 (mkrhohforce_3d_sphr :388)
     integer(c_int) function mgpu_mkrhohforce_sphr_c(p, g, nfabs, scal_force, is_prediction, thermal, umac, p0_1, p0_2, &
          psi, add_thermal, adv_bc, pmask) bind(C, name="mgpu_mkrhohforce_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs, is_prediction, add_thermal
       type(mgpu_fab), intent(inout) :: scal_force(*)
       type(mgpu_fab), intent(in) :: thermal(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: p0_1(*), p0_2(*), psi(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_mkrhohforce_sphr_c

     ! enthalpy_advance.f90:16, spherical
     integer(c_int) function mgpu_enthalpy_advance_sphr_c(p, g, which_step, sold, snew, sedge, sflux, scal_force, thermal, &
          umac, w0, w0mac, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, tempbar, psi, adv_bc, pmask) &
          bind(C, name="mgpu_enthalpy_advance_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: which_step
       type(mgpu_fab), intent(inout) :: sold(*), snew(*), scal_force(*)
       type(mgpu_fab), intent(in) :: thermal(*)
       type(c_ptr), intent(in) :: sedge(*), sflux(*), umac(*), w0mac(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), rhoh0_old(*), rho0_new(*), rhoh0_new(*), p0_old(*), p0_new(*)
       real(c_double), intent(in) :: tempbar(*), psi(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_enthalpy_advance_sphr_c

     ! fill_3d_data.f90:1280
     integer(c_int) function mgpu_make_normal_c(p, g, nfabs, normal) bind(C, name="mgpu_make_normal")
       import :: c_int, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: normal(*)
     end function mgpu_make_normal_c

     ! mkforce.f90:22 with spherical == 1 (mk_vel_force_3d_sphr :484)
     integer(c_int) function mgpu_mk_vel_force_sphr_c(p, g, nfabs, vel_force, is_final_update, uold, uedge, w0, w0mac, &
          gpi, s, index_rho, normal, rho0, grav, w0_force_cart, do_add_utilde_force) &
          bind(C, name="mgpu_mk_vel_force_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs, is_final_update, index_rho, do_add_utilde_force
       type(mgpu_fab), intent(inout) :: vel_force(*)
       type(mgpu_fab), intent(in) :: uold(*), gpi(*), s(*), normal(*), w0_force_cart(*)
       type(c_ptr), intent(in) :: uedge(*), w0mac(*)
       real(c_double), intent(in) :: w0(*), rho0(*), grav(*)
     end function mgpu_mk_vel_force_sphr_c

     ! advance_premac.f90:21, spherical
     integer(c_int) function mgpu_advance_premac_sphr_c(p, g, uold, sold, umac, gpi, normal, w0, w0mac, w0_force_cart, &
          rho0_old, grav_cell_old, adv_bc, phys_bc, pmask) bind(C, name="mgpu_advance_premac_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       type(mgpu_fab), intent(in) :: uold(*), sold(*), gpi(*), normal(*), w0_force_cart(*)
       type(c_ptr), intent(in) :: umac(*), w0mac(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), grav_cell_old(*)
       integer(c_int), intent(in) :: adv_bc(*), phys_bc(*), pmask(*)
     end function mgpu_advance_premac_sphr_c

     ! velocity_advance.f90:16, spherical
     integer(c_int) function mgpu_velocity_advance_sphr_c(p, g, uold, unew, sold, rhohalf, umac, gpi, normal, w0, w0mac, &
          w0_force_cart, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph, sponge, adv_bc, pmask) &
          bind(C, name="mgpu_velocity_advance_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       type(mgpu_fab), intent(in) :: uold(*), sold(*), rhohalf(*), gpi(*), normal(*), w0_force_cart(*), sponge(*)
       type(mgpu_fab), intent(inout) :: unew(*)
       type(c_ptr), intent(in) :: umac(*), w0mac(*)
       real(c_double), intent(in) :: w0(*), rho0_old(*), rho0_nph(*), grav_cell_old(*), grav_cell_nph(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_velocity_advance_sphr_c

     ! density_advance.f90:20, spherical
     integer(c_int) function mgpu_density_advance_sphr_c(p, g, which_step, sold, snew, sedge, sflux, &
          scal_force, umac, w0, w0mac, rho0_old, rho0_new, adv_bc, pmask) &
          bind(C, name="mgpu_density_advance_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: which_step
       type(mgpu_fab), intent(inout) :: sold(*)
       type(mgpu_fab), intent(inout) :: snew(*)
       type(c_ptr), intent(in) :: sedge(*)
       type(c_ptr), intent(in) :: sflux(*)
       type(mgpu_fab), intent(inout) :: scal_force(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*)
       type(c_ptr), intent(in) :: w0mac(*)
       real(c_double), intent(in) :: rho0_old(*)
       real(c_double), intent(in) :: rho0_new(*)
       integer(c_int), intent(in) :: adv_bc(*)
       integer(c_int), intent(in) :: pmask(*)
     end function mgpu_density_advance_sphr_c

     ! enthalpy_advance.f90:16
     integer(c_int) function mgpu_enthalpy_advance_c(p, which_step, sold, snew, sedge, sflux, scal_force, &
          thermal, umac, w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, tempbar, psi, grav_old, &
          grav_nph, adv_bc, pmask) &
          bind(C, name="mgpu_enthalpy_advance")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: which_step
       type(mgpu_fab), intent(inout) :: sold(*)
       type(mgpu_fab), intent(inout) :: snew(*)
       type(c_ptr), intent(in) :: sedge(*)
       type(c_ptr), intent(in) :: sflux(*)
       type(mgpu_fab), intent(inout) :: scal_force(*)
       type(mgpu_fab), intent(in) :: thermal(*)
       type(c_ptr), intent(in) :: umac(*)
       real(c_double), intent(in) :: w0(*)
       real(c_double), intent(in) :: rho0_old(*)
       real(c_double), intent(in) :: rhoh0_old(*)
       real(c_double), intent(in) :: rho0_new(*)
       real(c_double), intent(in) :: rhoh0_new(*)
       real(c_double), intent(in) :: p0_old(*)
       real(c_double), intent(in) :: p0_new(*)
       real(c_double), intent(in) :: tempbar(*)
       real(c_double), intent(in) :: psi(*)
       real(c_double), intent(in) :: grav_old(*)
       real(c_double), intent(in) :: grav_nph(*)
       integer(c_int), intent(in) :: adv_bc(*)
       integer(c_int), intent(in) :: pmask(*)
     end function mgpu_enthalpy_advance_c

     ! velocity_advance.f90:16
     integer(c_int) function mgpu_velocity_advance_c(p, uold, unew, sold, rhohalf, umac, gpi, w0, w0_force, &
          rho0_old, rho0_nph, grav_cell_old, grav_cell_nph, sponge, adv_bc, pmask) &
          bind(C, name="mgpu_velocity_advance")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_fab), intent(in) :: uold(*)
       type(mgpu_fab), intent(inout) :: unew(*)
       type(mgpu_fab), intent(in) :: sold(*)
       type(mgpu_fab), intent(in) :: rhohalf(*)
       type(c_ptr), intent(in) :: umac(*)
       type(mgpu_fab), intent(in) :: gpi(*)
       real(c_double), intent(in) :: w0(*)
       real(c_double), intent(in) :: w0_force(*)
       real(c_double), intent(in) :: rho0_old(*)
       real(c_double), intent(in) :: rho0_nph(*)
       real(c_double), intent(in) :: grav_cell_old(*)
       real(c_double), intent(in) :: grav_cell_nph(*)
       type(mgpu_fab), intent(in) :: sponge(*)
       integer(c_int), intent(in) :: adv_bc(*)
       integer(c_int), intent(in) :: pmask(*)
     end function mgpu_velocity_advance_c

     ! advance_premac.f90:21
     integer(c_int) function mgpu_advance_premac_c(p, uold, sold, umac, gpi, w0, w0_force, rho0_old, &
          grav_cell_old, adv_bc, phys_bc, pmask) &
          bind(C, name="mgpu_advance_premac")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_fab), intent(in) :: uold(*)
       type(mgpu_fab), intent(in) :: sold(*)
       type(c_ptr), intent(in) :: umac(*)
       type(mgpu_fab), intent(in) :: gpi(*)
       real(c_double), intent(in) :: w0(*)
       real(c_double), intent(in) :: w0_force(*)
       real(c_double), intent(in) :: rho0_old(*)
       real(c_double), intent(in) :: grav_cell_old(*)
       integer(c_int), intent(in) :: adv_bc(*)
       integer(c_int), intent(in) :: phys_bc(*)
       integer(c_int), intent(in) :: pmask(*)
     end function mgpu_advance_premac_c
     ! estdt (Source/estdt.f90:29) for one level: force is the multifab mk_vel_force filled (:117-120)
     integer(c_int) function mgpu_estdt_c(p, nfabs, u, s, force, divU, dSdt, w0, p0, gamma1bar, rho_min, cflfac, &
          dt, umax) bind(C, name="mgpu_estdt")
       import :: c_int, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: u(*), s(*), force(*), divU(*), dSdt(*)
       real(c_double), intent(in) :: w0(*), p0(*), gamma1bar(*)
       real(c_double), value :: rho_min, cflfac
       real(c_double), intent(inout) :: dt, umax
     end function mgpu_estdt_c

     ! multifab_min_c / multifab_max_c of one component over the valid zones of a level, all ranks: the verbose >= 1
     ! lines of density_advance.f90:374-402 (div_comp = rho_comp for the species), enthalpy_advance.f90:440-449,
     ! velocity_advance.f90:142-160; the caller keeps its write(6, ...) statements
     integer(c_int) function mgpu_minmax_c(p, nfabs, s, comp, div_comp, smin, smax) bind(C, name="mgpu_minmax")
       import :: c_int, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs, comp, div_comp
       type(mgpu_fab), intent(in) :: s(*)
       real(c_double), intent(out) :: smin, smax
     end function mgpu_minmax_c

     ! estdt_3d_sphr (Source/estdt.f90:620) for one level; w0mac: the three face multifabs of make_w0mac
     integer(c_int) function mgpu_estdt_sphr_c(p, g, nfabs, u, s, force, divU, dSdt, w0mac, w0, p0, gamma1bar, &
          rho_min, cflfac, dt, umax) bind(C, name="mgpu_estdt_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: u(*), s(*), force(*), divU(*), dSdt(*)
       type(c_ptr), intent(in) :: w0mac(*)
       real(c_double), intent(in) :: w0(*), p0(*), gamma1bar(*)
       real(c_double), value :: rho_min, cflfac
       real(c_double), intent(inout) :: dt, umax
     end function mgpu_estdt_sphr_c

     ! make_etarho_planar (Source/make_eta.f90:36): etarho_ec(0:nr), etarho_cc(0:nr-1) of one level
     integer(c_int) function mgpu_make_etarho_planar_c(p, nfabs, etarhoflux, etarho_ec, etarho_cc) &
          bind(C, name="mgpu_make_etarho_planar")
       import :: c_int, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(in) :: etarhoflux(*)
       real(c_double), intent(out) :: etarho_ec(*), etarho_cc(*)
     end function mgpu_make_etarho_planar_c
     ! ---- the EOS and the pieces of the path that call it (SURVEY 8 f4 / f1 / f3) ----
     ! eos_init (Microphysics/EOS/eos.F90:26): called once after eos_init / network_init with their values
     integer(c_int) function mgpu_set_eos_c(e) bind(C, name="mgpu_set_eos")
       import :: c_int, mgpu_eos
       type(mgpu_eos), intent(in) :: e
     end function mgpu_set_eos_c

     ! eos(input, state) at n points: state(n, MGPU_EOS_NQ), xn(n, nspec), Fortran order
     integer(c_int) function mgpu_eos_eval_c(input, n, state, xn) bind(C, name="mgpu_eos_eval")
       import :: c_int, c_long, c_double
       integer(c_int), value :: input
       integer(c_long), value :: n
       real(c_double), intent(inout) :: state(*)
       real(c_double), intent(in) :: xn(*)
     end function mgpu_eos_eval_c

     ! rhoh_vs_t.f90:20 (planar)
     integer(c_int) function mgpu_make_h_from_rhot_edge_c(p, nfabs, sedge, rho0_old, rhoh0_old, t0_old, rho0_edge_old, &
          rhoh0_edge_old, t0_edge_old, rho0_new, rhoh0_new, t0_new, rho0_edge_new, rhoh0_edge_new, t0_edge_new) &
          bind(C, name="mgpu_make_h_from_rhot_edge")
       import :: c_int, c_ptr, c_double, mgpu_params
       type(mgpu_params), intent(in) :: p
       integer(c_int), value :: nfabs
       type(c_ptr), intent(in) :: sedge(*)
       real(c_double), intent(in) :: rho0_old(*), rhoh0_old(*), t0_old(*), rho0_edge_old(*), rhoh0_edge_old(*), t0_edge_old(*)
       real(c_double), intent(in) :: rho0_new(*), rhoh0_new(*), t0_new(*), rho0_edge_new(*), rhoh0_edge_new(*), t0_edge_new(*)
     end function mgpu_make_h_from_rhot_edge_c

     ! rhoh_vs_t.f90:20 with spherical == 1
     integer(c_int) function mgpu_make_h_from_rhot_edge_sphr_c(p, g, nfabs, sedge, rho0_old, rhoh0_old, t0_old, rho0_new, &
          rhoh0_new, t0_new, adv_bc, pmask) bind(C, name="mgpu_make_h_from_rhot_edge_sphr")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs
       type(c_ptr), intent(in) :: sedge(*)
       real(c_double), intent(in) :: rho0_old(*), rhoh0_old(*), t0_old(*), rho0_new(*), rhoh0_new(*), t0_new(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_make_h_from_rhot_edge_sphr_c

     ! mkscalforce.f90:719; g: c_loc of an mgpu_geom, c_null_ptr for planar geometry
     integer(c_int) function mgpu_mktempforce_c(p, g, nfabs, temp_force, umac, s, thermal, p0_old, psi, adv_bc, pmask) &
          bind(C, name="mgpu_mktempforce")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(c_ptr), value :: g
       integer(c_int), value :: nfabs
       type(mgpu_fab), intent(inout) :: temp_force(*)
       type(c_ptr), intent(in) :: umac(*)
       type(mgpu_fab), intent(in) :: s(*), thermal(*)
       real(c_double), intent(in) :: p0_old(*), psi(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_mktempforce_c

     ! firstdt.f90:25
     integer(c_int) function mgpu_firstdt_c(p, g, nfabs, u, gpi, s, divU, rho0, p0, grav, gamma1bar, cflfac, init_shrink, &
          use_soundspeed_firstdt, use_divu_firstdt, dt, umax) bind(C, name="mgpu_firstdt")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(c_ptr), value :: g
       integer(c_int), value :: nfabs, use_soundspeed_firstdt, use_divu_firstdt
       type(mgpu_fab), intent(in) :: u(*), gpi(*), s(*), divU(*)
       real(c_double), intent(in) :: rho0(*), p0(*), grav(*), gamma1bar(*)
       real(c_double), value :: cflfac, init_shrink
       real(c_double), intent(inout) :: dt, umax
     end function mgpu_firstdt_c

     ! rhoh_vs_t.f90:800
     integer(c_int) function mgpu_make_t_from_rhoh_c(p, g, nfabs, state, p0, use_eos_e_instead_of_h, adv_bc, pmask) &
          bind(C, name="mgpu_make_t_from_rhoh")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(c_ptr), value :: g
       integer(c_int), value :: nfabs, use_eos_e_instead_of_h
       type(mgpu_fab), intent(inout) :: state(*)
       real(c_double), intent(in) :: p0(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_make_t_from_rhoh_c

     ! rhoh_vs_t.f90:1165
     integer(c_int) function mgpu_make_t_from_rhop_c(p, g, nfabs, state, p0, update_rhoh, use_pprime_in_tfromp, adv_bc, &
          pmask) bind(C, name="mgpu_make_t_from_rhop")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(c_ptr), value :: g
       integer(c_int), value :: nfabs, update_rhoh, use_pprime_in_tfromp
       type(mgpu_fab), intent(inout) :: state(*)
       real(c_double), intent(in) :: p0(*)
       integer(c_int), intent(in) :: adv_bc(*), pmask(*)
     end function mgpu_make_t_from_rhop_c

     ! average.f90:24 (g: c_loc of an mgpu_geom, c_null_ptr for planar geometry)
     integer(c_int) function mgpu_average_c(p, g, nfabs, phi, incomp, nr_irreg, drdxfac, phibar) &
          bind(C, name="mgpu_average")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(c_ptr), value :: g
       integer(c_int), value :: nfabs, incomp, nr_irreg, drdxfac
       type(mgpu_fab), intent(in) :: phi(*)
       real(c_double), intent(inout) :: phibar(*)
     end function mgpu_average_c

     ! make_eta.f90:256
     integer(c_int) function mgpu_make_etarho_spherical_c(p, g, nfabs, sold, snew, umac, w0mac, rho0_old, rho0_new, &
          normal, nr_irreg, drdxfac, etarho_ec, etarho_cc) bind(C, name="mgpu_make_etarho_spherical")
       import :: c_int, c_ptr, c_double, mgpu_params, mgpu_geom, mgpu_fab
       type(mgpu_params), intent(in) :: p
       type(mgpu_geom), intent(in) :: g
       integer(c_int), value :: nfabs, nr_irreg, drdxfac
       type(mgpu_fab), intent(in) :: sold(*), snew(*), normal(*)
       type(c_ptr), intent(in) :: umac(*), w0mac(*)
       real(c_double), intent(in) :: rho0_old(*), rho0_new(*)
       real(c_double), intent(inout) :: etarho_ec(*), etarho_cc(*)
     end function mgpu_make_etarho_spherical_c
  end interface

  public :: mgpu_average_c, mgpu_make_etarho_spherical_c
  public :: mgpu_set_eos_c, mgpu_eos_eval_c, mgpu_make_h_from_rhot_edge_c, mgpu_make_h_from_rhot_edge_sphr_c
  public :: mgpu_mktempforce_c, mgpu_firstdt_c, mgpu_make_t_from_rhoh_c, mgpu_make_t_from_rhop_c
  public :: mgpu_startup, mgpu_shutdown, mgpu_fill_params, mgpu_describe, mgpu_describe_edges, mgpu_check
  public :: mgpu_make_edge_scal_c, mgpu_bds_c, mgpu_mk_rhoX_flux_c, mgpu_mk_rhoh_flux_c, mgpu_update_scal_c
  public :: mgpu_update_velocity_c, mgpu_addw0_c, mgpu_mkutrans_c, mgpu_velpred_c
  public :: mgpu_put_1d_array_on_cart_c, mgpu_make_w0mac_c, mgpu_make_s0mac_c, mgpu_addw0_sphr_c
  public :: mgpu_mk_rhoX_flux_sphr_c, mgpu_mk_rhoh_flux_sphr_c, mgpu_update_velocity_sphr_c
  public :: mgpu_mkutrans_sphr_c, mgpu_velpred_sphr_c, mgpu_modify_scal_force_sphr_c, mgpu_put_in_pert_form_sphr_c
  public :: mgpu_fill_boundary_c, mgpu_convert_rhoX_to_X_c, mgpu_modify_scal_force_c, mgpu_put_in_pert_form_c, mgpu_mkrhohforce_c, mgpu_mk_vel_force_c
  public :: mgpu_density_advance_mf_c, mgpu_fill_boundary_mf_c
  public :: mgpu_velocity_advance_mf_c, mgpu_advance_premac_mf_c, mgpu_enthalpy_advance_mf_c
  public :: mgpu_mkrhohforce_sphr_c, mgpu_enthalpy_advance_sphr_c
  public :: mgpu_make_normal_c, mgpu_mk_vel_force_sphr_c, mgpu_advance_premac_sphr_c, mgpu_velocity_advance_sphr_c
  public :: mgpu_density_advance_c, mgpu_density_advance_sphr_c, mgpu_enthalpy_advance_c, mgpu_velocity_advance_c, mgpu_advance_premac_c
  public :: mgpu_comm_unique_id, mgpu_comm_init, mgpu_comm_finalize, mgpu_set_option
  public :: mgpu_malloc, mgpu_free, mgpu_memcpy_h2d, mgpu_memcpy_d2h
  public :: mgpu_fill_geom, mgpu_estdt_c, mgpu_estdt_sphr_c, mgpu_make_etarho_planar_c, mgpu_minmax_c
  public :: make_edge_scal_gpu

contains

  ! one MPI rank <-> one GPU: call once after boxlib_initialize (Source/main.f90:14)
  subroutine mgpu_startup()
    use parallel, only: parallel_myproc
    integer :: ngpu_per_node
    ngpu_per_node = 8
    call mgpu_check(mgpu_init(int(mod(parallel_myproc(), ngpu_per_node), c_int)))
  end subroutine mgpu_startup

  subroutine mgpu_shutdown()
    call mgpu_check(mgpu_finalize())
  end subroutine mgpu_shutdown

  ! reference convention: every failure is bl_error -> abort (e.g. make_edge_scal.f90:853)
  subroutine mgpu_check(rc)
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: cmsg(:)
    character(len=512) :: msg
    integer :: k
    if (rc == 0) return
    call c_f_pointer(mgpu_last_error(), cmsg, [512])
    msg = ' '
    do k = 1, 512
       if (cmsg(k) == c_null_char) exit
       msg(k:k) = cmsg(k)
    end do
    call bl_error(trim(msg))
  end subroutine mgpu_check

  ! module variables of the reference -> the per-call parameter block
  subroutine mgpu_fill_params(p, dm, dx, dt, domlo, domhi)
    use probin_module, only: ppm_type, bds_type, slope_order, ppm_trace_forces, species_pred_type, &
         enthalpy_pred_type, evolve_base_state, do_sponge, do_eos_h_above_cutoff, base_cutoff_density, &
         buoyancy_cutoff_factor, rotation_radius
    use geometry, only: spherical, nr_fine, base_cutoff_density_coord, omega, sin_theta, cos_theta
    use variables, only: rho_comp, rhoh_comp, spec_comp, temp_comp, pi_comp, trac_comp, nscal, ntrac, rel_eps
    use network, only: nspec
    type(mgpu_params), intent(out) :: p
    integer, intent(in) :: dm, domlo(:), domhi(:)
    real(dp_t), intent(in) :: dx(:), dt
    p%dm = dm; p%mem_space = MGPU_HOST
    p%ppm_type = ppm_type; p%bds_type = bds_type; p%slope_order = slope_order
    p%ppm_trace_forces = ppm_trace_forces
    p%species_pred_type = species_pred_type; p%enthalpy_pred_type = enthalpy_pred_type
    p%spherical = spherical
    p%evolve_base_state = merge(1, 0, evolve_base_state)
    p%do_sponge = merge(1, 0, do_sponge)
    p%do_eos_h_above_cutoff = merge(1, 0, do_eos_h_above_cutoff)
    p%rho_comp = rho_comp; p%rhoh_comp = rhoh_comp; p%spec_comp = spec_comp
    p%temp_comp = temp_comp; p%pi_comp = pi_comp; p%trac_comp = trac_comp
    p%nspec = nspec; p%ntrac = ntrac; p%nscal = nscal
    p%domlo = 0; p%domhi = 0; p%dx = 0.d0
    p%domlo(1:dm) = domlo(1:dm); p%domhi(1:dm) = domhi(1:dm)
    p%nr = nr_fine
    p%dt = dt; p%dx(1:dm) = dx(1:dm)
    p%rel_eps = rel_eps; p%base_cutoff_density = base_cutoff_density
    p%base_cutoff_density_coord = base_cutoff_density_coord(1)
    p%buoyancy_cutoff_factor = buoyancy_cutoff_factor
    p%omega = omega; p%sin_theta = sin_theta; p%cos_theta = cos_theta; p%rotation_radius = rotation_radius
  end subroutine mgpu_fill_params

  ! what dataptr/get_box/nghost give the reference kernels (make_edge_scal.f90:70-76), per local fab
  ! spherical geometry from the geometry module and probin (Source/geometry.f90, Source/_parameters:632-651)
  subroutine mgpu_fill_geom(g)
    use geometry, only: center, dr, nr_fine, r_cc_loc, r_edge_loc
    use probin_module, only: prob_lo, s0_interp_type, w0_interp_type, s0mac_interp_type, w0mac_interp_type
    type(mgpu_geom), intent(out) :: g
    g%center = center(1:3)
    g%prob_lo = prob_lo(1:3)
    g%dr = dr(1)
    g%nr_fine = nr_fine
    g%r_cc_loc = c_loc(r_cc_loc(1,0))
    g%r_edge_loc = c_loc(r_edge_loc(1,0))
    g%s0_interp_type = s0_interp_type
    g%w0_interp_type = w0_interp_type
    g%s0mac_interp_type = s0mac_interp_type
    g%w0mac_interp_type = w0mac_interp_type
  end subroutine mgpu_fill_geom

  subroutine mgpu_describe(mf, d)
    type(multifab), intent(in) :: mf
    type(mgpu_fab), intent(out) :: d(:)
    real(dp_t), pointer :: fp(:,:,:,:)
    integer :: i, dm
    logical :: nod(mf%dim)
    dm = mf%dim
    nod = nodal_flags(mf)
    do i = 1, nfabs(mf)
       fp => dataptr(mf, i)
       d(i)%ptr = c_loc(fp(lbound(fp,1), lbound(fp,2), lbound(fp,3), 1))
       d(i)%lo = 0; d(i)%hi = 0; d(i)%nodal = 0
       d(i)%lo(1:dm) = lwb(get_box(mf, i))
       d(i)%hi(1:dm) = upb(get_box(mf, i))
       d(i)%ng = nghost(mf)
       d(i)%nc = ncomp(mf)
       d(i)%nodal(1:dm) = merge(1, 0, nod)
    end do
  end subroutine mgpu_describe

  ! dm face-centred multifabs -> dm descriptor arrays + the array of C pointers to them
  subroutine mgpu_describe_edges(mfs, d, dp)
    type(multifab), intent(in) :: mfs(:)
    type(mgpu_fab), intent(out), target :: d(:,:)   ! (nfabs, dm)
    type(c_ptr), intent(out) :: dp(:)
    integer :: comp
    do comp = 1, size(mfs)
       call mgpu_describe(mfs(comp), d(:,comp))
       dp(comp) = c_loc(d(1,comp))
    end do
  end subroutine mgpu_describe_edges

  ! Body of make_edge_scal (Source/make_edge_scal.f90:26) for one level n: replaces :69-121.
  subroutine make_edge_scal_gpu(s, sedge, umac, force, dx, dt, is_vel, bc_level, start_scomp, &
                                start_bccomp, num_comp, is_conservative, domlo, domhi)
    type(multifab), intent(in)    :: s, umac(:), force
    type(multifab), intent(inout) :: sedge(:)
    real(dp_t),     intent(in)    :: dx(:), dt
    logical,        intent(in)    :: is_vel, is_conservative
    type(bc_level), intent(in)    :: bc_level
    integer,        intent(in)    :: start_scomp, start_bccomp, num_comp, domlo(:), domhi(:)

    type(mgpu_params) :: p
    type(mgpu_fab), allocatable, target :: ds(:), df(:), dse(:,:), dum(:,:)
    type(c_ptr) :: pse(3), pum(3)
    integer :: nf, dm, i

    dm = s%dim
    nf = nfabs(s)
    allocate(ds(nf), df(nf), dse(nf,dm), dum(nf,dm))
    call mgpu_fill_params(p, dm, dx, dt, domlo, domhi)
    call mgpu_describe(s, ds)
    call mgpu_describe(force, df)
    call mgpu_describe_edges(sedge, dse, pse)
    call mgpu_describe_edges(umac, dum, pum)
    ! adv_bc_level_array(i,:,:,:) is contiguous in (d, side, comp) for fixed i only after a copy:
    do i = 1, nf
       call mgpu_check(mgpu_make_edge_scal_c(p, 1_c_int, ds(i:i), fab_slice(pse, dse, i, dm), &
            fab_slice(pum, dum, i, dm), df(i:i), &
            reshape(bc_level%adv_bc_level_array(i,:,:,:), [size(bc_level%adv_bc_level_array(i,:,:,:))]), &
            merge(1_c_int, 0_c_int, is_vel), int(start_scomp, c_int), int(start_bccomp, c_int), &
            int(num_comp, c_int), merge(1_c_int, 0_c_int, is_conservative)))
    end do
  contains
    function fab_slice(pp, dd, i, dm) result(q)
      type(c_ptr), intent(in) :: pp(:)
      type(mgpu_fab), intent(in), target :: dd(:,:)
      integer, intent(in) :: i, dm
      type(c_ptr) :: q(3)
      integer :: c
      q = c_null_ptr
      do c = 1, dm
         q(c) = c_loc(dd(i,c))
      end do
    end function fab_slice
  end subroutine make_edge_scal_gpu

end module maestro_b200_shim
