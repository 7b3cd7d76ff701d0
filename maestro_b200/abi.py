"""ctypes mirror of include/maestro_b200.h (structs, constants, prototypes).

Only declarations live here; `maestro_b200.lib` loads the CUDA library.  The same declarations are
applied to the CPU oracle's `mo_*` symbols by the tests (the oracle mirrors the ABI one to one).
"""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

# bc_module constants (FBoxLib), see include/maestro_b200.h
PERIODIC, INTERIOR = -1, 0
INLET, OUTLET, SYMMETRY, SLIP_WALL, NO_SLIP_WALL = 11, 12, 13, 14, 15
REFLECT_ODD, REFLECT_EVEN, FOEXTRAP, EXT_DIR, HOEXTRAP = 20, 21, 22, 23, 24
HOST, DEVICE = 0, 1
PREDICT_RHOPRIME_AND_X, PREDICT_RHOX, PREDICT_RHO_AND_X = 1, 2, 3
PREDICT_RHOH, PREDICT_RHOHPRIME, PREDICT_H = 0, 1, 2
PREDICT_T_THEN_RHOHPRIME, PREDICT_T_THEN_H, PREDICT_HPRIME, PREDICT_TPRIME_THEN_H = 3, 4, 5, 6


class mgpu_fab(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("lo", C.c_int * 3),
        ("hi", C.c_int * 3),
        ("ng", C.c_int),
        ("nc", C.c_int),
        ("nodal", C.c_int * 3),
    ]


class mgpu_params(C.Structure):
    _fields_ = [
        ("dm", C.c_int),
        ("mem_space", C.c_int),
        ("ppm_type", C.c_int),
        ("bds_type", C.c_int),
        ("slope_order", C.c_int),
        ("ppm_trace_forces", C.c_int),
        ("species_pred_type", C.c_int),
        ("enthalpy_pred_type", C.c_int),
        ("spherical", C.c_int),
        ("evolve_base_state", C.c_int),
        ("do_sponge", C.c_int),
        ("do_eos_h_above_cutoff", C.c_int),
        ("rho_comp", C.c_int),
        ("rhoh_comp", C.c_int),
        ("spec_comp", C.c_int),
        ("temp_comp", C.c_int),
        ("pi_comp", C.c_int),
        ("trac_comp", C.c_int),
        ("nspec", C.c_int),
        ("ntrac", C.c_int),
        ("nscal", C.c_int),
        ("domlo", C.c_int * 3),
        ("domhi", C.c_int * 3),
        ("nr", C.c_int),
        ("dt", C.c_double),
        ("dx", C.c_double * 3),
        ("rel_eps", C.c_double),
        ("base_cutoff_density", C.c_double),
        ("base_cutoff_density_coord", C.c_int),
        ("buoyancy_cutoff_factor", C.c_double),
        ("omega", C.c_double),
        ("sin_theta", C.c_double),
        ("cos_theta", C.c_double),
        ("rotation_radius", C.c_double),
    ]


class mgpu_halo_plan(C.Structure):
    _fields_ = [
        ("dir", C.c_int),
        ("up_rank", C.c_int),
        ("dn_rank", C.c_int),
        ("nplanes", C.c_int),
        ("plane_doubles", C.c_long),
        ("send_up_k0", C.c_int),
        ("send_dn_k0", C.c_int),
        ("recv_lo_k0", C.c_int),
        ("recv_hi_k0", C.c_int),
    ]


class mgpu_geom(C.Structure):
    """spherical geometry (geometry module + interpolation switches of probin), include/maestro_b200.h"""
    _fields_ = [
        ("center", C.c_double * 3),
        ("prob_lo", C.c_double * 3),
        ("dr", C.c_double),
        ("nr_fine", C.c_int),
        ("r_cc_loc", c_double_p),
        ("r_edge_loc", c_double_p),
        ("s0_interp_type", C.c_int),
        ("w0_interp_type", C.c_int),
        ("s0mac_interp_type", C.c_int),
        ("w0mac_interp_type", C.c_int),
    ]


P_ = C.POINTER(mgpu_params)
G_ = C.POINTER(mgpu_geom)
F_ = C.POINTER(mgpu_fab)
FF_ = C.POINTER(F_)  # array of dm pointers, each to an array of nfabs fabs

# name -> (restype, argtypes) for every operator symbol shared by the library (mgpu_) and the oracle (mo_)
OPERATORS = {
    "fill_boundary": (C.c_int, [P_, F_, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p]),
    "make_edge_scal": (C.c_int, [P_, C.c_int, F_, FF_, FF_, F_, c_int_p] + [C.c_int] * 5),
    "bds": (C.c_int, [P_, C.c_int, F_, FF_, FF_, F_, c_int_p] + [C.c_int] * 5),
    "mk_rhoX_flux": (C.c_int, [P_, C.c_int, FF_, F_, FF_, FF_] + [c_double_p] * 6 + [C.c_int] * 2),
    "mk_rhoh_flux": (C.c_int, [P_, C.c_int, FF_, FF_, FF_] + [c_double_p] * 9),
    "update_scal": (C.c_int, [P_, C.c_int, C.c_int, C.c_int, F_, F_, FF_, F_, c_double_p, F_]),
    "update_velocity": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, F_, F_, c_double_p]),
    "addw0": (C.c_int, [P_, C.c_int, FF_, c_double_p, C.c_double]),
    "mkutrans": (C.c_int, [P_, C.c_int, F_, F_, FF_, c_double_p, c_int_p, c_int_p]),
    "velpred": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, F_, c_double_p, c_int_p, c_int_p]),
    "modify_scal_force": (C.c_int, [P_, C.c_int, F_, F_, FF_] + [c_double_p] * 3 + [C.c_int] * 2),
    "convert_rhoX_to_X": (C.c_int, [P_, C.c_int, F_, C.c_int]),
    "put_in_pert_form": (C.c_int, [P_, C.c_int, F_, c_double_p, C.c_int, C.c_int]),
    "mkrhohforce": (C.c_int, [P_, C.c_int, F_, C.c_int, F_, FF_] + [c_double_p] * 6 + [C.c_int]),
    "mk_vel_force": (C.c_int, [P_, C.c_int, F_, C.c_int, F_, FF_, c_double_p, F_, F_, C.c_int] + [c_double_p] * 3
                     + [C.c_int]),
    "advance_premac": (C.c_int, [P_, F_, F_, FF_, F_] + [c_double_p] * 4 + [c_int_p] * 3),
    "velocity_advance": (C.c_int, [P_, F_, F_, F_, F_, FF_, F_] + [c_double_p] * 6 + [F_] + [c_int_p] * 2),
    "enthalpy_advance": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, F_, F_, FF_] + [c_double_p] * 11 + [c_int_p] * 2),
    "estdt": (C.c_int, [P_, C.c_int, F_, F_, F_, F_, F_] + [c_double_p] * 3 + [C.c_double] * 2 + [c_double_p] * 2),
    "minmax": (C.c_int, [P_, C.c_int, F_, C.c_int, C.c_int, c_double_p, c_double_p]),
    "make_etarho_planar": (C.c_int, [P_, C.c_int, F_, c_double_p, c_double_p]),
    "estdt_sphr": (C.c_int, [P_, G_, C.c_int, F_, F_, F_, F_, F_, FF_] + [c_double_p] * 3 + [C.c_double] * 2
                   + [c_double_p] * 2),
    # spherical geometry
    "put_1d_array_on_cart": (C.c_int, [P_, G_, C.c_int, c_double_p, F_, C.c_int, C.c_int]),
    "make_w0mac": (C.c_int, [P_, G_, C.c_int, c_double_p, FF_, F_]),
    "make_s0mac": (C.c_int, [P_, G_, C.c_int, c_double_p, FF_, F_]),
    "addw0_sphr": (C.c_int, [P_, C.c_int, FF_, FF_, C.c_double]),
    "mk_rhoX_flux_sphr": (C.c_int, [P_, C.c_int, FF_, FF_, FF_, FF_, FF_, FF_, C.c_int, C.c_int]),
    "mk_rhoh_flux_sphr": (C.c_int, [P_, C.c_int] + [FF_] * 8),
    "update_velocity_sphr": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, F_, F_, FF_]),
    "mkutrans_sphr": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, c_int_p, c_int_p]),
    "velpred_sphr": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, F_, FF_, c_int_p, c_int_p]),
    "modify_scal_force_sphr": (C.c_int, [P_, G_, C.c_int, F_, F_, FF_, F_, c_double_p, C.c_int, C.c_int]),
    "put_in_pert_form_sphr": (C.c_int, [P_, G_, C.c_int, F_, c_double_p, C.c_int, C.c_int]),
    "mkrhohforce_sphr": (C.c_int, [P_, G_, C.c_int, F_, C.c_int, F_, FF_, c_double_p, c_double_p, c_double_p, C.c_int,
                                   c_int_p, c_int_p]),
    "enthalpy_advance_sphr": (C.c_int, [P_, G_, C.c_int, F_, F_, FF_, FF_, F_, F_, FF_, c_double_p, FF_] + [c_double_p] * 8
                              + [c_int_p] * 2),
    "make_normal": (C.c_int, [P_, G_, C.c_int, F_]),
    "mk_vel_force_sphr": (C.c_int, [P_, G_, C.c_int, F_, C.c_int, F_, FF_, c_double_p, FF_, F_, F_, C.c_int, F_, c_double_p,
                                    c_double_p, F_, C.c_int]),
    "advance_premac_sphr": (C.c_int, [P_, G_, F_, F_, FF_, F_, F_, c_double_p, FF_, F_, c_double_p, c_double_p]
                            + [c_int_p] * 3),
    "velocity_advance_sphr": (C.c_int, [P_, G_, F_, F_, F_, F_, FF_, F_, F_, c_double_p, FF_, F_] + [c_double_p] * 4
                              + [F_] + [c_int_p] * 2),
    "density_advance_sphr": (C.c_int, [P_, G_, C.c_int, F_, F_, FF_, FF_, F_, FF_, c_double_p, FF_, c_double_p, c_double_p,
                                       c_int_p, c_int_p]),
    "density_advance": (C.c_int, [P_, C.c_int, F_, F_, FF_, FF_, F_, FF_, c_double_p, F_] + [c_double_p] * 4
                        + [c_int_p] * 2),
}

# the EOS and the pieces of the path that call it
E_ = None  # set below (POINTER(mgpu_eos))


class mgpu_eos(C.Structure):
    """gamma_law_general + the generic front end's limits (include/maestro_b200.h)"""
    MAXSPEC = 32
    _fields_ = [
        ("kind", C.c_int),
        ("assume_neutral", C.c_int),
        ("nspec", C.c_int),
        ("pad_", C.c_int),
        ("gamma", C.c_double),
        ("k_B", C.c_double),
        ("n_A", C.c_double),
        ("mintemp", C.c_double), ("maxtemp", C.c_double), ("mindens", C.c_double), ("maxdens", C.c_double),
        ("mine", C.c_double), ("maxe", C.c_double), ("minp", C.c_double), ("maxp", C.c_double),
        ("minh", C.c_double), ("maxh", C.c_double),
        ("small_temp", C.c_double),
        ("aion", C.c_double * 32),
        ("zion", C.c_double * 32),
    ]


E_ = C.POINTER(mgpu_eos)
EOS_NONE, EOS_GAMMA_LAW = 0, 1
EOS_INPUT_RT, EOS_INPUT_RH, EOS_INPUT_TP, EOS_INPUT_RP, EOS_INPUT_RE, EOS_INPUT_PS, EOS_INPUT_PH, EOS_INPUT_TH = range(1, 9)
EOS_Q = ["rho", "T", "p", "e", "h", "cv", "cp", "cs", "dpdT", "dpdr", "dedT", "dedr", "dhdT", "mu", "abar", "zbar"]
OPERATORS.update({
    "set_eos": (C.c_int, [E_]),
    "eos_eval": (C.c_int, [C.c_int, C.c_long, c_double_p, c_double_p]),
    "make_h_from_rhot_edge": (C.c_int, [P_, C.c_int, FF_] + [c_double_p] * 12),
    "make_h_from_rhot_edge_sphr": (C.c_int, [P_, G_, C.c_int, FF_] + [c_double_p] * 6 + [c_int_p] * 2),
    "mktempforce": (C.c_int, [P_, G_, C.c_int, F_, FF_, F_, F_, c_double_p, c_double_p, c_int_p, c_int_p]),
    "firstdt": (C.c_int, [P_, G_, C.c_int, F_, F_, F_, F_] + [c_double_p] * 4 + [C.c_double] * 2 + [C.c_int] * 2
                + [c_double_p] * 2),
    "make_t_from_rhoh": (C.c_int, [P_, G_, C.c_int, F_, c_double_p, C.c_int, c_int_p, c_int_p]),
    "make_t_from_rhop": (C.c_int, [P_, G_, C.c_int, F_, c_double_p, C.c_int, C.c_int, c_int_p, c_int_p]),
    "average": (C.c_int, [P_, G_, C.c_int, F_, C.c_int, C.c_int, C.c_int, c_double_p]),
    "make_etarho_spherical": (C.c_int, [P_, G_, C.c_int, F_, F_, FF_, FF_, c_double_p, c_double_p, F_, C.c_int, C.c_int,
                                        c_double_p, c_double_p]),
})

# several boxes per rank: the CUDA library only (the oracle's multifab is one box covering the domain)
MULTIBOX = {
    "mgpu_fill_boundary_mf": (C.c_int, [P_, C.c_int, F_, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p]),
    "mgpu_density_advance_mf": (C.c_int, [P_, C.c_int, C.c_int, F_, F_, FF_, FF_, F_, FF_, c_double_p, F_]
                                + [c_double_p] * 4 + [c_int_p] * 2),
    "mgpu_velocity_advance_mf": (C.c_int, [P_, C.c_int, F_, F_, F_, F_, FF_, F_] + [c_double_p] * 6 + [F_] + [c_int_p] * 2),
    "mgpu_advance_premac_mf": (C.c_int, [P_, C.c_int, F_, F_, FF_, F_] + [c_double_p] * 4 + [c_int_p] * 3),
    "mgpu_enthalpy_advance_mf": (C.c_int, [P_, C.c_int, C.c_int, F_, F_, FF_, FF_, F_, F_, FF_] + [c_double_p] * 10
                                 + [c_int_p] * 2),
}

# symbols only the library has
LIFECYCLE = {
    "mgpu_init": (C.c_int, [C.c_int]),
    "mgpu_finalize": (C.c_int, []),
    "mgpu_synchronize": (C.c_int, []),
    "mgpu_last_error": (C.c_char_p, []),
    "mgpu_version": (C.c_char_p, []),
    "mgpu_launch_count": (C.c_long, [C.c_int]),
    "mgpu_stream": (C.c_void_p, []),
    "mgpu_copy_bytes": (C.c_int, [C.POINTER(C.c_long), C.POINTER(C.c_long), C.c_int]),
    "mgpu_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "mgpu_profile": (C.c_int, [C.c_int]),
    "mgpu_profile_get": (C.c_int, [C.c_int, c_double_p, C.POINTER(C.c_long)]),
    "mgpu_set_stream": (C.c_int, [C.c_void_p]),
    "mgpu_host_register": (C.c_int, [C.c_void_p, C.c_long]),
    "mgpu_host_unregister": (C.c_int, [C.c_void_p]),
    "mgpu_register": (C.c_int, [C.c_void_p, C.c_long, C.c_int]),
    "mgpu_unregister": (C.c_int, [C.c_void_p]),
    "mgpu_invalidate": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mgpu_download": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mgpu_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mgpu_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_long]),
    "mgpu_free": (C.c_int, [C.c_void_p]),
    "mgpu_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long]),
    "mgpu_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long]),
    "mgpu_halo_plan_make": (C.c_int, [P_, F_, c_int_p, C.c_int, C.c_int, C.POINTER(mgpu_halo_plan)]),
    "mgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mgpu_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "mgpu_comm_finalize": (C.c_int, []),
}


def declare(lib, prefix):
    """Attach prototypes for every operator symbol `prefix+name` of OPERATORS to a ctypes library."""
    for name, (res, args) in OPERATORS.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
    return lib


def all_symbols():
    return ["mgpu_" + n for n in OPERATORS] + list(LIFECYCLE) + list(MULTIBOX)
