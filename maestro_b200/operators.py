"""Host-side mirror of the reference's multifab-level (L3/L4) operators for this path.

Same names, argument order and meaning as the Fortran routines they stand for, so that parity tests
read like the reference's own drivers:

    make_edge_scal   Source/make_edge_scal.f90:26      mk_rhoX_flux  Source/mkflux.f90:48
    bds              Source/bds.f90:16                 mk_rhoh_flux  Source/mkflux.f90:652
    update_scal      Source/update_scal.f90:16         update_velocity Source/update_vel.f90:15
    addw0            Source/addw0.f90:19               mkutrans / velpred  mkutrans.f90:17 / velpred.f90:21
    modify_scal_force / convert_rhoX_to_X / put_in_pert_form (glue)
    density_advance  Source/density_advance.f90:20

`Operators(lib, prefix)` binds the wrappers to any library exporting the ABI; the product binds the
CUDA library (prefix "mgpu_"), the tests additionally bind the CPU oracle (prefix "mo_").
Errors follow the reference convention (`bl_error` -> abort): a nonzero return raises.
"""
import ctypes as C

import numpy as np

from . import abi
from .fab import as_double_p, as_int_p, fab_pp, fab_ptr


class MaestroError(RuntimeError):
    pass


class Operators:
    def __init__(self, lib, prefix):
        self.lib = abi.declare(lib, prefix)
        self.prefix = prefix
        self._err = getattr(lib, prefix + "last_error")
        self._err.restype = C.c_char_p

    def _call(self, name, *args):
        rc = getattr(self.lib, self.prefix + name)(*args)
        if rc != 0:
            raise MaestroError("%s%s: %s" % (self.prefix, name, self._err().decode()))

    # ---- ghost fill ----------------------------------------------------------------------------
    def fill_boundary(self, p, s, scomp, bccomp, ncomp, adv_bc, pmask):
        bc, bcp = as_int_p(adv_bc)
        pm, pmp = as_int_p(pmask)
        f = fab_ptr(s)
        self._call("fill_boundary", C.byref(p), f, scomp, bccomp, ncomp, bcp, pmp)

    # ---- edge states ---------------------------------------------------------------------------
    def make_edge_scal(self, p, s, sedge, umac, force, adv_bc, is_vel, start_scomp, start_bccomp, num_comp,
                       is_conservative):
        bc, bcp = as_int_p(adv_bc)
        sf, ff = fab_ptr(s), fab_ptr(force)
        se, k1 = fab_pp(sedge)
        um, k2 = fab_pp(umac)
        name = "bds" if p.bds_type == 1 else "make_edge_scal"
        self._call(name, C.byref(p), 1, sf, se, um, ff, bcp, int(is_vel), start_scomp, start_bccomp, num_comp,
                   int(is_conservative))

    bds = make_edge_scal

    # ---- fluxes --------------------------------------------------------------------------------
    def mk_rhoX_flux(self, p, sflux, etarhoflux, sedge, umac, w0, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new,
                     rho0_predicted_edge, startcomp, endcomp):
        keep = [as_double_p(x) for x in (w0, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new, rho0_predicted_edge)]
        sf, k1 = fab_pp(sflux)
        se, k2 = fab_pp(sedge)
        um, k3 = fab_pp(umac)
        eta = fab_ptr(etarhoflux)
        self._call("mk_rhoX_flux", C.byref(p), 1, sf, eta, se, um, *[k[1] for k in keep], startcomp, endcomp)

    def mk_rhoh_flux(self, p, sflux, sedge, umac, w0, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new, rhoh0_old,
                     rhoh0_edge_old, rhoh0_new, rhoh0_edge_new):
        keep = [as_double_p(x) for x in (w0, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new, rhoh0_old,
                                         rhoh0_edge_old, rhoh0_new, rhoh0_edge_new)]
        sf, k1 = fab_pp(sflux)
        se, k2 = fab_pp(sedge)
        um, k3 = fab_pp(umac)
        self._call("mk_rhoh_flux", C.byref(p), 1, sf, se, um, *[k[1] for k in keep])

    # ---- updates -------------------------------------------------------------------------------
    def update_scal(self, p, nstart, nstop, sold, snew, sflux, scal_force, p0_new=None, p0_new_cart=None):
        """Source/update_scal.f90:16; p0_new (planar) / p0_new_cart (spherical) feed the EOS reset below the cutoff."""
        sf, k1 = fab_pp(sflux)
        pk, pp = as_double_p(p0_new) if p0_new is not None else (None, None)
        self._call("update_scal", C.byref(p), 1, nstart, nstop, fab_ptr(sold), fab_ptr(snew), sf, fab_ptr(scal_force),
                   pp, fab_ptr(p0_new_cart) if p0_new_cart is not None else None)

    def update_velocity(self, p, uold, unew, umac, uedge, force, sponge, w0):
        w, wp = as_double_p(w0)
        um, k1 = fab_pp(umac)
        ue, k2 = fab_pp(uedge)
        self._call("update_velocity", C.byref(p), 1, fab_ptr(uold), fab_ptr(unew), um, ue, fab_ptr(force),
                   fab_ptr(sponge), wp)

    def addw0(self, p, umac, w0, mult):
        w, wp = as_double_p(w0)
        um, k1 = fab_pp(umac)
        self._call("addw0", C.byref(p), 1, um, wp, float(mult))

    # ---- velocity prediction -------------------------------------------------------------------
    def mkutrans(self, p, utilde, ufull, utrans, w0, adv_bc, phys_bc):
        w, wp = as_double_p(w0)
        bc, bcp = as_int_p(adv_bc)
        pb, pbp = as_int_p(phys_bc)
        ut, k1 = fab_pp(utrans)
        self._call("mkutrans", C.byref(p), 1, fab_ptr(utilde), fab_ptr(ufull), ut, wp, bcp, pbp)

    def velpred(self, p, utilde, ufull, umac, utrans, force, w0, adv_bc, phys_bc):
        w, wp = as_double_p(w0)
        bc, bcp = as_int_p(adv_bc)
        pb, pbp = as_int_p(phys_bc)
        um, k1 = fab_pp(umac)
        ut, k2 = fab_pp(utrans)
        self._call("velpred", C.byref(p), 1, fab_ptr(utilde), fab_ptr(ufull), um, ut, fab_ptr(force), wp, bcp, pbp)

    # ---- glue ----------------------------------------------------------------------------------
    def modify_scal_force(self, p, force, s, umac, s0, s0_edge, w0, comp, fullform=False):
        keep = [as_double_p(x) for x in (s0, s0_edge, w0)]
        um, k1 = fab_pp(umac)
        self._call("modify_scal_force", C.byref(p), 1, fab_ptr(force), fab_ptr(s), um, *[k[1] for k in keep], comp,
                   int(fullform))

    def convert_rhoX_to_X(self, p, s, flag):
        self._call("convert_rhoX_to_X", C.byref(p), 1, fab_ptr(s), int(flag))

    def put_in_pert_form(self, p, s, base, comp, flag):
        b, bp = as_double_p(base)
        self._call("put_in_pert_form", C.byref(p), 1, fab_ptr(s), bp, comp, int(flag))

    # ---- spherical geometry (the reference's *_3d_sphr branches; fill_3d_data.f90) ----------------
    def put_1d_array_on_cart(self, p, geom, s0, s0_cart, is_input_edge_centered, is_output_a_vector):
        a, ap = as_double_p(s0)
        self._call("put_1d_array_on_cart", C.byref(p), C.byref(geom.c), 1, ap, fab_ptr(s0_cart),
                   int(is_input_edge_centered), int(is_output_a_vector))

    def make_w0mac(self, p, geom, w0, w0mac, w0_cart=None):
        a, ap = as_double_p(w0)
        wm, k1 = fab_pp(w0mac)
        self._call("make_w0mac", C.byref(p), C.byref(geom.c), 1, ap, wm, fab_ptr(w0_cart) if w0_cart is not None else None)

    def make_s0mac(self, p, geom, s0, s0mac, s0_cart=None):
        a, ap = as_double_p(s0)
        sm, k1 = fab_pp(s0mac)
        self._call("make_s0mac", C.byref(p), C.byref(geom.c), 1, ap, sm, fab_ptr(s0_cart) if s0_cart is not None else None)

    def addw0_sphr(self, p, umac, w0mac, mult):
        um, k1 = fab_pp(umac)
        wm, k2 = fab_pp(w0mac)
        self._call("addw0_sphr", C.byref(p), 1, um, wm, float(mult))

    def mk_rhoX_flux_sphr(self, p, sflux, sedge, umac, w0mac, rho0mac_old, rho0mac_new, startcomp, endcomp):
        ptrs = [fab_pp(x) for x in (sflux, sedge, umac, w0mac, rho0mac_old, rho0mac_new)]
        self._call("mk_rhoX_flux_sphr", C.byref(p), 1, *[q[0] for q in ptrs], startcomp, endcomp)

    def mk_rhoh_flux_sphr(self, p, sflux, sedge, umac, w0mac, rho0mac_old, rho0mac_new, h0mac_old, h0mac_new):
        ptrs = [fab_pp(x) for x in (sflux, sedge, umac, w0mac, rho0mac_old, rho0mac_new, h0mac_old, h0mac_new)]
        self._call("mk_rhoh_flux_sphr", C.byref(p), 1, *[q[0] for q in ptrs])

    def update_velocity_sphr(self, p, uold, unew, umac, uedge, force, sponge, w0mac):
        ptrs = [fab_pp(x) for x in (umac, uedge, w0mac)]
        self._call("update_velocity_sphr", C.byref(p), 1, fab_ptr(uold), fab_ptr(unew), ptrs[0][0], ptrs[1][0],
                   fab_ptr(force), fab_ptr(sponge), ptrs[2][0])

    def mkutrans_sphr(self, p, utilde, ufull, utrans, w0mac, adv_bc, phys_bc):
        bc, bcp = as_int_p(adv_bc)
        pb, pbp = as_int_p(phys_bc)
        ut, k1 = fab_pp(utrans)
        wm, k2 = fab_pp(w0mac)
        self._call("mkutrans_sphr", C.byref(p), 1, fab_ptr(utilde), fab_ptr(ufull), ut, wm, bcp, pbp)

    def velpred_sphr(self, p, utilde, ufull, umac, utrans, force, w0mac, adv_bc, phys_bc):
        bc, bcp = as_int_p(adv_bc)
        pb, pbp = as_int_p(phys_bc)
        um, k1 = fab_pp(umac)
        ut, k2 = fab_pp(utrans)
        wm, k3 = fab_pp(w0mac)
        self._call("velpred_sphr", C.byref(p), 1, fab_ptr(utilde), fab_ptr(ufull), um, ut, fab_ptr(force), wm, bcp, pbp)

    def modify_scal_force_sphr(self, p, geom, force, s, umac, s0_cart, w0, comp, fullform=False):
        w, wp = as_double_p(w0)
        um, k1 = fab_pp(umac)
        self._call("modify_scal_force_sphr", C.byref(p), C.byref(geom.c), 1, fab_ptr(force), fab_ptr(s), um,
                   fab_ptr(s0_cart), wp, comp, int(fullform))

    def put_in_pert_form_sphr(self, p, geom, s, s0, comp, flag):
        b, bp = as_double_p(s0)
        self._call("put_in_pert_form_sphr", C.byref(p), C.byref(geom.c), 1, fab_ptr(s), bp, comp, int(flag))

    def density_advance_sphr(self, p, geom, which_step, sold, snew, sedge, sflux, scal_force, umac, w0, w0mac, rho0_old,
                             rho0_new, adv_bc, pmask):
        keep = [as_double_p(x) for x in (w0, rho0_old, rho0_new)]
        bc, bcp = as_int_p(adv_bc)
        pm, pmp = as_int_p(pmask)
        ptrs = [fab_pp(x) for x in (sedge, sflux, umac, w0mac)]
        self._call("density_advance_sphr", C.byref(p), C.byref(geom.c), which_step, fab_ptr(sold), fab_ptr(snew),
                   ptrs[0][0], ptrs[1][0], fab_ptr(scal_force), ptrs[2][0], keep[0][1], ptrs[3][0], keep[1][1],
                   keep[2][1], bcp, pmp)

    def mkrhohforce_sphr(self, p, geom, scal_force, is_prediction, thermal, umac, p0_1, p0_2, psi, add_thermal, adv_bc,
                         pmask):
        keep = [as_double_p(x) for x in (p0_1, p0_2, psi)]
        bc, bcp = as_int_p(adv_bc)
        pm, pmp = as_int_p(pmask)
        um, k1 = fab_pp(umac)
        self._call("mkrhohforce_sphr", C.byref(p), C.byref(geom.c), 1, fab_ptr(scal_force), int(is_prediction),
                   fab_ptr(thermal), um, keep[0][1], keep[1][1], keep[2][1], int(add_thermal), bcp, pmp)

    def enthalpy_advance_sphr(self, p, geom, which_step, sold, snew, sedge, sflux, scal_force, thermal, umac, w0, w0mac,
                              rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, psi, adv_bc, pmask, tempbar=None):
        if tempbar is None:
            tempbar = np.zeros(geom.c.nr_fine)
        keep = [as_double_p(x) for x in (w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, tempbar, psi)]
        bc, bcp = as_int_p(adv_bc)
        pm, pmp = as_int_p(pmask)
        se, k1 = fab_pp(sedge)
        sf, k2 = fab_pp(sflux)
        um, k3 = fab_pp(umac)
        wm, k4 = fab_pp(w0mac)
        self._call("enthalpy_advance_sphr", C.byref(p), C.byref(geom.c), which_step, fab_ptr(sold), fab_ptr(snew), se, sf,
                   fab_ptr(scal_force), fab_ptr(thermal), um, keep[0][1], wm, *[k[1] for k in keep[1:]], bcp, pmp)

    def make_normal(self, p, geom, normal):
        self._call("make_normal", C.byref(p), C.byref(geom.c), 1, fab_ptr(normal))

    def mk_vel_force_sphr(self, p, geom, vel_force, is_final_update, uold, uedge, w0, w0mac, gpi, s, index_rho, normal,
                          rho0, grav, w0_force_cart, do_add_utilde_force):
        keep = [as_double_p(x) for x in (w0, rho0, grav)]
        ue, k1 = fab_pp(uedge)
        wm, k2 = fab_pp(w0mac)
        self._call("mk_vel_force_sphr", C.byref(p), C.byref(geom.c), 1, fab_ptr(vel_force), int(is_final_update),
                   fab_ptr(uold), ue, keep[0][1], wm, fab_ptr(gpi), fab_ptr(s), index_rho, fab_ptr(normal), keep[1][1],
                   keep[2][1], fab_ptr(w0_force_cart), int(do_add_utilde_force))

    def advance_premac_sphr(self, p, geom, uold, sold, umac, gpi, normal, w0, w0mac, w0_force_cart, rho0_old,
                            grav_cell_old, adv_bc, phys_bc, pmask):
        keep = [as_double_p(x) for x in (w0, rho0_old, grav_cell_old)]
        bc, bcp = as_int_p(adv_bc)
        pb, pbp = as_int_p(phys_bc)
        pm, pmp = as_int_p(pmask)
        um, k1 = fab_pp(umac)
        wm, k2 = fab_pp(w0mac)
        self._call("advance_premac_sphr", C.byref(p), C.byref(geom.c), fab_ptr(uold), fab_ptr(sold), um, fab_ptr(gpi),
                   fab_ptr(normal), keep[0][1], wm, fab_ptr(w0_force_cart), keep[1][1], keep[2][1], bcp, pbp, pmp)

    def velocity_advance_sphr(self, p, geom, uold, unew, sold, rhohalf, umac, gpi, normal, w0, w0mac, w0_force_cart,
                              rho0_old, rho0_nph, grav_cell_old, grav_cell_nph, sponge, adv_bc, pmask):
        keep = [as_double_p(x) for x in (w0, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph)]
        bc, bcp = as_int_p(adv_bc)
        pm, pmp = as_int_p(pmask)
        um, k1 = fab_pp(umac)
        wm, k2 = fab_pp(w0mac)
        self._call("velocity_advance_sphr", C.byref(p), C.byref(geom.c), fab_ptr(uold), fab_ptr(unew), fab_ptr(sold),
                   fab_ptr(rhohalf), um, fab_ptr(gpi), fab_ptr(normal), keep[0][1], wm, fab_ptr(w0_force_cart),
                   keep[1][1], keep[2][1], keep[3][1], keep[4][1], fab_ptr(sponge), bcp, pmp)

    # ---- L4 driver -----------------------------------------------------------------------------
    def density_advance(self, p, which_step, sold, snew, sedge, sflux, scal_force, umac, w0, etarhoflux, rho0_old,
                        rho0_new, p0_dummy, rho0_predicted_edge, adv_bc, pmask):
        keep = [as_double_p(x) for x in (w0, rho0_old, rho0_new, p0_dummy, rho0_predicted_edge)]
        bc, bcp = as_int_p(adv_bc)
        pm, pmp = as_int_p(pmask)
        se, k1 = fab_pp(sedge)
        sf, k2 = fab_pp(sflux)
        um, k3 = fab_pp(umac)
        self._call("density_advance", C.byref(p), which_step, fab_ptr(sold), fab_ptr(snew), se, sf,
                   fab_ptr(scal_force), um, keep[0][1], fab_ptr(etarhoflux), keep[1][1], keep[2][1], keep[3][1],
                   keep[4][1], bcp, pmp)

    # ---- force builders -------------------------------------------------------------------
    def mkrhohforce(self, p, scal_force, is_prediction, thermal, umac, p0_1, p0_2, rho0_1, rho0_2, grav, psi,
                    add_thermal):
        keep = [as_double_p(x) for x in (p0_1, p0_2, rho0_1, rho0_2, grav, psi)]
        um, k1 = fab_pp(umac)
        self._call("mkrhohforce", C.byref(p), 1, fab_ptr(scal_force), int(is_prediction), fab_ptr(thermal), um,
                   *[k[1] for k in keep], int(add_thermal))

    def mk_vel_force(self, p, vel_force, is_final_update, uold, uedge, w0, gpi, s, index_rho, rho0, grav, w0_force,
                     do_add_utilde_force=True):
        keep = [as_double_p(x) for x in (w0, rho0, grav, w0_force)]
        ue, k1 = fab_pp(uedge)
        self._call("mk_vel_force", C.byref(p), 1, fab_ptr(vel_force), int(is_final_update), fab_ptr(uold), ue,
                   keep[0][1], fab_ptr(gpi), fab_ptr(s), index_rho, keep[1][1], keep[2][1], keep[3][1],
                   int(do_add_utilde_force))

    # ---- reductions next to the path -------------------------------------------------
    def estdt(self, p, u, s, force, divU, dSdt, w0, p0, gamma1bar, cflfac, dt, umax=0.0, rho_min=1.0e-20):
        """Source/estdt.f90:29 for one level: returns (min(dt, dt_lev), max(umax, umax_lev)); `force` is the
        velocity force the caller built with mk_vel_force (:117-120); rel_eps = 1e-8 * umax is the caller's (:229)."""
        keep = [as_double_p(x) for x in (w0, p0, gamma1bar)]
        dt_c, um_c = C.c_double(dt), C.c_double(umax)
        self._call("estdt", C.byref(p), 1, fab_ptr(u), fab_ptr(s), fab_ptr(force), fab_ptr(divU), fab_ptr(dSdt),
                   keep[0][1], keep[1][1], keep[2][1], float(rho_min), float(cflfac), C.byref(dt_c), C.byref(um_c))
        return dt_c.value, um_c.value

    def minmax(self, p, s, comp, div_comp=0):
        """multifab_min_c / multifab_max_c of component `comp` (1-based) over the valid zones, all ranks; div_comp > 0:
        of s(comp) / s(div_comp) (the X = rhoX / rho of density_advance.f90:378-386). Returns (smin, smax)."""
        lo, hi = C.c_double(0.0), C.c_double(0.0)
        self._call("minmax", C.byref(p), 1, fab_ptr(s), int(comp), int(div_comp), C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def verbose_report(self, p, episode, snew, spec_names=None, level=1):
        """The `verbose >= 1` lines the reference prints at the end of an episode (density_advance.f90:374-402 formats
        1999-2003, enthalpy_advance.f90:440-453 formats 1999/2001/2004, velocity_advance.f90:142-160 formats 999-1004),
        as a list of strings (the I/O processor prints them)."""
        lines = []
        if episode == "density_advance":
            lines.append("... Level %1d update:" % level)
            for n in range(p.nspec):
                name = (spec_names[n] if spec_names else "X(%d)" % (n + 1))[:16].ljust(16)
                mn, mx = self.minmax(p, snew, p.spec_comp + n, p.rho_comp)
                lines.append("... new min/max : %s  %s  %s" % (name, fortran_e(mn), fortran_e(mx)))
            mn, mx = self.minmax(p, snew, p.rho_comp)
            lines.append("... new min/max : density           %s  %s" % (fortran_e(mn), fortran_e(mx)))
            if p.ntrac >= 1:
                mn, mx = self.minmax(p, snew, p.trac_comp)
                lines.append("... new min/max : tracer            %s  %s" % (fortran_e(mn), fortran_e(mx)))
        elif episode == "enthalpy_advance":
            mn, mx = self.minmax(p, snew, p.rhoh_comp)
            lines.append("... Level %1d update:" % level)
            lines.append("... new min/max : rho * H           %s  %s" % (fortran_e(mn), fortran_e(mx)))
            lines.append(" ")
        elif episode == "velocity_advance":
            lines.append("... Level %1d update:" % level)
            for d in range(p.dm):
                mn, mx = self.minmax(p, snew, d + 1)
                lines.append("... new min/max : %s-velocity       %s  %s" % ("xyz"[d], fortran_e(mn), fortran_e(mx)))
            lines.append(" ")
        else:
            raise ValueError("verbose_report: unknown episode " + episode)
        return lines

    def estdt_sphr(self, p, geom, u, s, force, divU, dSdt, w0mac, w0, p0, gamma1bar, cflfac, dt, umax=0.0,
                   rho_min=1.0e-20):
        """estdt_3d_sphr (Source/estdt.f90:620) for one level: returns (min(dt, dt_lev), max(umax, umax_lev))."""
        keep = [as_double_p(x) for x in (w0, p0, gamma1bar)]
        wm, k1 = fab_pp(w0mac)
        dt_c, um_c = C.c_double(dt), C.c_double(umax)
        self._call("estdt_sphr", C.byref(p), C.byref(geom.c), 1, fab_ptr(u), fab_ptr(s), fab_ptr(force), fab_ptr(divU),
                   fab_ptr(dSdt), wm, keep[0][1], keep[1][1], keep[2][1], float(rho_min), float(cflfac),
                   C.byref(dt_c), C.byref(um_c))
        return dt_c.value, um_c.value

    def make_etarho_planar(self, p, etarhoflux):
        """Source/make_eta.f90:36: returns (etarho_ec(0:nr), etarho_cc(0:nr-1))."""
        ec, cc = np.zeros(p.nr + 1), np.zeros(p.nr)
        k1, k2 = as_double_p(ec), as_double_p(cc)
        self._call("make_etarho_planar", C.byref(p), 1, fab_ptr(etarhoflux), k1[1], k2[1])
        return k1[0].copy(), k2[0].copy()

    # ---- the other L4 drivers ---------------------------------------------------------------------------
    def advance_premac(self, p, uold, sold, umac, gpi, w0, w0_force, rho0_old, grav_cell_old, adv_bc, phys_bc, pmask):
        keep = [as_double_p(x) for x in (w0, w0_force, rho0_old, grav_cell_old)]
        ints = [as_int_p(x) for x in (adv_bc, phys_bc, pmask)]
        um, k1 = fab_pp(umac)
        self._call("advance_premac", C.byref(p), fab_ptr(uold), fab_ptr(sold), um, fab_ptr(gpi),
                   *[k[1] for k in keep], *[k[1] for k in ints])

    def velocity_advance(self, p, uold, unew, sold, rhohalf, umac, gpi, w0, w0_force, rho0_old, rho0_nph,
                         grav_cell_old, grav_cell_nph, sponge, adv_bc, pmask):
        keep = [as_double_p(x) for x in (w0, w0_force, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph)]
        ints = [as_int_p(x) for x in (adv_bc, pmask)]
        um, k1 = fab_pp(umac)
        self._call("velocity_advance", C.byref(p), fab_ptr(uold), fab_ptr(unew), fab_ptr(sold), fab_ptr(rhohalf), um,
                   fab_ptr(gpi), *[k[1] for k in keep], fab_ptr(sponge), *[k[1] for k in ints])

    def enthalpy_advance(self, p, which_step, sold, snew, sedge, sflux, scal_force, thermal, umac, w0, rho0_old,
                         rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, psi, grav_old, grav_nph, adv_bc, pmask,
                         tempbar=None):
        """Source/enthalpy_advance.f90:16; tempbar (the reference passes it after p0_new) is only read by the
        temperature-based predictions."""
        if tempbar is None:
            tempbar = np.zeros(p.nr)
        keep = [as_double_p(x) for x in (w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, tempbar, psi,
                                         grav_old, grav_nph)]
        ints = [as_int_p(x) for x in (adv_bc, pmask)]
        se, k1 = fab_pp(sedge)
        sf, k2 = fab_pp(sflux)
        um, k3 = fab_pp(umac)
        self._call("enthalpy_advance", C.byref(p), which_step, fab_ptr(sold), fab_ptr(snew), se, sf,
                   fab_ptr(scal_force), fab_ptr(thermal), um, *[k[1] for k in keep], *[k[1] for k in ints])

    # ---- the EOS and the pieces of the path that call it ----------------------------------
    def set_eos(self, eos):
        """eos_init (Microphysics/EOS/eos.F90:26): process state; None unsets it."""
        self._call("set_eos", C.byref(eos) if eos is not None else None)

    def eos_eval(self, eos_input, state, xn):
        """eos(input, state) (eos.F90:99) at n points: `state` a dict of 1-D arrays (rho, T, p, e, h: the ones the input
        mode names are read), `xn` (n, nspec); returns a dict with every column of abi.EOS_Q."""
        from .abi import EOS_Q
        xn = np.ascontiguousarray(np.asarray(xn, dtype=np.float64))
        n = xn.shape[0]
        buf = np.zeros((len(EOS_Q), n))
        for q, name in enumerate(EOS_Q):
            if name in state:
                buf[q] = state[name]
        xt = np.ascontiguousarray(xn.T)  # point-fastest
        self._call("eos_eval", int(eos_input), n, buf.ctypes.data_as(C.POINTER(C.c_double)),
                   xt.ctypes.data_as(C.POINTER(C.c_double)))
        return {name: buf[q].copy() for q, name in enumerate(EOS_Q)}

    def make_h_from_rhot_edge(self, p, sedge, rho0_old, rhoh0_old, t0_old, rho0_edge_old, rhoh0_edge_old, t0_edge_old,
                              rho0_new, rhoh0_new, t0_new, rho0_edge_new, rhoh0_edge_new, t0_edge_new):
        """makeHfromRhoT_edge (Source/rhoh_vs_t.f90:20), planar."""
        keep = [as_double_p(x) for x in (rho0_old, rhoh0_old, t0_old, rho0_edge_old, rhoh0_edge_old, t0_edge_old, rho0_new,
                                         rhoh0_new, t0_new, rho0_edge_new, rhoh0_edge_new, t0_edge_new)]
        se, k1 = fab_pp(sedge)
        self._call("make_h_from_rhot_edge", C.byref(p), 1, se, *[k[1] for k in keep])

    def make_h_from_rhot_edge_sphr(self, p, geom, sedge, rho0_old, rhoh0_old, t0_old, rho0_new, rhoh0_new, t0_new, adv_bc,
                                   pmask):
        keep = [as_double_p(x) for x in (rho0_old, rhoh0_old, t0_old, rho0_new, rhoh0_new, t0_new)]
        ints = [as_int_p(x) for x in (adv_bc, pmask)]
        se, k1 = fab_pp(sedge)
        self._call("make_h_from_rhot_edge_sphr", C.byref(p), C.byref(geom.c), 1, se, *[k[1] for k in keep],
                   *[k[1] for k in ints])

    def mktempforce(self, p, temp_force, umac, s, thermal, p0_old, psi, adv_bc, pmask, geom=None):
        """mktempforce (Source/mkscalforce.f90:719)."""
        keep = [as_double_p(x) for x in (p0_old, psi)]
        ints = [as_int_p(x) for x in (adv_bc, pmask)]
        um, k1 = fab_pp(umac)
        self._call("mktempforce", C.byref(p), C.byref(geom.c) if geom is not None else None, 1, fab_ptr(temp_force), um,
                   fab_ptr(s), fab_ptr(thermal), keep[0][1], keep[1][1], ints[0][1], ints[1][1])

    def firstdt(self, p, u, gpi, s, divU, rho0, p0, grav, gamma1bar, cflfac, init_shrink, dt, umax=0.0,
                use_soundspeed_firstdt=False, use_divu_firstdt=False, geom=None):
        """firstdt (Source/firstdt.f90:25) for one level: returns (min(dt, dt_lev*init_shrink), max(umax, umax_lev))."""
        keep = [as_double_p(x) for x in (rho0, p0, grav, gamma1bar)]
        dt_c, um_c = C.c_double(dt), C.c_double(umax)
        self._call("firstdt", C.byref(p), C.byref(geom.c) if geom is not None else None, 1, fab_ptr(u), fab_ptr(gpi),
                   fab_ptr(s), fab_ptr(divU), *[k[1] for k in keep], float(cflfac), float(init_shrink),
                   int(use_soundspeed_firstdt), int(use_divu_firstdt), C.byref(dt_c), C.byref(um_c))
        return dt_c.value, um_c.value

    def make_t_from_rhoh(self, p, state, p0, adv_bc, pmask, use_eos_e_instead_of_h=False, geom=None):
        """makeTfromRhoH (Source/rhoh_vs_t.f90:800)."""
        pk, pp = as_double_p(p0)
        ints = [as_int_p(x) for x in (adv_bc, pmask)]
        self._call("make_t_from_rhoh", C.byref(p), C.byref(geom.c) if geom is not None else None, 1, fab_ptr(state), pp,
                   int(use_eos_e_instead_of_h), ints[0][1], ints[1][1])

    def make_t_from_rhop(self, p, state, p0, adv_bc, pmask, update_rhoh=False, use_pprime_in_tfromp=False, geom=None):
        """makeTfromRhoP (Source/rhoh_vs_t.f90:1165)."""
        pk, pp = as_double_p(p0)
        ints = [as_int_p(x) for x in (adv_bc, pmask)]
        self._call("make_t_from_rhop", C.byref(p), C.byref(geom.c) if geom is not None else None, 1, fab_ptr(state), pp,
                   int(update_rhoh), int(use_pprime_in_tfromp), ints[0][1], ints[1][1])

    # ---- averages -----------------------------------------------------------------------------------
    def average(self, p, phi, incomp, geom=None, nr_irreg=0, drdxfac=1):
        """average (Source/average.f90:24) of one level: returns phibar(0:nr-1)."""
        nr = geom.c.nr_fine if geom is not None else p.nr
        out = np.zeros(nr)
        self._call("average", C.byref(p), C.byref(geom.c) if geom is not None else None, 1, fab_ptr(phi), int(incomp),
                   int(nr_irreg), int(drdxfac), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def make_etarho_spherical(self, p, geom, sold, snew, umac, w0mac, rho0_old, rho0_new, normal, nr_irreg, drdxfac=1):
        """make_etarho_spherical (Source/make_eta.f90:256): returns (etarho_ec(0:nr_fine), etarho_cc(0:nr_fine-1))."""
        nr = geom.c.nr_fine
        ec, cc = np.zeros(nr + 1), np.zeros(nr)
        keep = [as_double_p(x) for x in (rho0_old, rho0_new)]
        um, k1 = fab_pp(umac)
        wm, k2 = fab_pp(w0mac)
        self._call("make_etarho_spherical", C.byref(p), C.byref(geom.c), 1, fab_ptr(sold), fab_ptr(snew), um, wm,
                   keep[0][1], keep[1][1], fab_ptr(normal), int(nr_irreg), int(drdxfac),
                   ec.ctypes.data_as(C.POINTER(C.c_double)), cc.ctypes.data_as(C.POINTER(C.c_double)))
        return ec, cc


def fortran_e(x, width=17, digits=10):
    """Fortran's e<width>.<digits> edit descriptor: 0.dddddddddE+xx, right-justified"""
    if x == 0.0:
        m, e = 0.0, 0
    else:
        e = int(np.floor(np.log10(abs(x)))) + 1
        m = x / 10.0 ** e
        if abs(round(m, digits)) >= 1.0:
            m, e = m / 10.0, e + 1
    body = "%.*f" % (digits, m)
    if abs(e) > 99:
        tail = "%+04d" % e
    else:
        tail = "E%+03d" % e
    return (body + tail).rjust(width)
