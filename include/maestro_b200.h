/*
 * maestro_b200.h -- C ABI of the B200-native MAESTRO advective hot path.
 *
 * This is the drop-in boundary: every entry point replaces the per-box ("L2") kernels that
 * one public multifab-level ("L3") routine of the reference calls, and takes exactly what that
 * routine hands to its kernels: raw fp64 arrays (Fortran order, x fastest, component slowest,
 * ghost cells included), the valid box lo/hi, ghost widths, dx, dt, integer BC tables and the
 * runtime parameters that live in Fortran modules (`probin_module`, `variables`, `network`).
 * The reference has no FFI of its own; the cut is the L3->L2 call in each file cited below.
 * A Fortran ISO_C_BINDING module that binds these symbols is in shim/maestro_b200_shim.f90 and
 * described in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no torch/CUDA types.  Pointers are HOST pointers when params->mem_space ==
 *    MGPU_HOST (the library stages host<->device itself, synchronously) or DEVICE pointers when
 *    MGPU_DEVICE (device-resident episode; asynchronous on the library stream until
 *    mgpu_synchronize()).
 *  - component numbers (rho_comp, scomp, bccomp ...) are 1-based exactly as in the Fortran.
 *  - adv_bc is the contiguous copy of `the_bc_level(n)%adv_bc_level_array(i,:,:,:)`, i.e.
 *    adv_bc[(d-1) + dm*((side-1) + 2*(bccomp-1))], d=1..dm, side=1(lo),2(hi).
 *  - base-state arrays (w0, rho0_old, ...) are the 1-D slice for the level: element r = 0..nr.
 *  - every function returns 0 on success; nonzero => mgpu_last_error() holds the message and the
 *    Fortran wrapper calls bl_error (reference convention, e.g. make_edge_scal.f90:853).
 */
#ifndef MAESTRO_B200_H
#define MAESTRO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* bc_module constants of FBoxLib (external to the reference tree; treated as opaque enums that the
 * Fortran shim passes through unchanged). */
enum {
  MGPU_BC_PERIODIC = -1,
  MGPU_BC_INTERIOR = 0,
  MGPU_BC_INLET = 11,
  MGPU_BC_OUTLET = 12,
  MGPU_BC_SYMMETRY = 13,
  MGPU_BC_SLIP_WALL = 14,
  MGPU_BC_NO_SLIP_WALL = 15,
  MGPU_BC_REFLECT_ODD = 20,
  MGPU_BC_REFLECT_EVEN = 21,
  MGPU_BC_FOEXTRAP = 22,
  MGPU_BC_EXT_DIR = 23,
  MGPU_BC_HOEXTRAP = 24
};

enum { MGPU_HOST = 0, MGPU_DEVICE = 1 };

/* species_pred_type / enthalpy_pred_type: Source/pred_parameters.f90:5-17 */
enum { MGPU_PREDICT_RHOPRIME_AND_X = 1, MGPU_PREDICT_RHOX = 2, MGPU_PREDICT_RHO_AND_X = 3 };
enum {
  MGPU_PREDICT_RHOH = 0,
  MGPU_PREDICT_RHOHPRIME = 1,
  MGPU_PREDICT_H = 2,
  MGPU_PREDICT_T_THEN_RHOHPRIME = 3,
  MGPU_PREDICT_T_THEN_H = 4,
  MGPU_PREDICT_HPRIME = 5,
  MGPU_PREDICT_TPRIME_THEN_H = 6
};

/* One fab = what `dataptr(mf,i)` + `get_box(mf,i)` + `nghost(mf)` give the reference kernels
 * (e.g. make_edge_scal.f90:70-76).  ptr addresses element (lo-ng, lo-ng, lo-ng, comp 1).
 * Extent in dim d (< dm) is hi[d]-lo[d]+1 + 2*ng + nodal[d]; dims >= dm have extent 1. */
typedef struct {
  double* ptr;
  int lo[3];
  int hi[3];
  int ng;
  int nc;
  int nodal[3];
} mgpu_fab;

/* Runtime parameters the reference keeps in module variables (probin_module: Source/_parameters;
 * variables: Source/variables.f90:100-124; network: nspec).  Passed on every call because the
 * reference mutates ppm_type/bds_type at run time (Exec/UNIT_TESTS/test_advect/varden.f90:278,287). */
typedef struct {
  int dm;                 /* 2 or 3 */
  int mem_space;          /* MGPU_HOST or MGPU_DEVICE */
  int ppm_type;           /* 0,1,2  (_parameters:500) */
  int bds_type;           /* 0,1    (_parameters:504) */
  int slope_order;        /* 0,2,4  (_parameters:492) */
  int ppm_trace_forces;   /* 0,1    (_parameters:509) */
  int species_pred_type;  /* 1,2,3 */
  int enthalpy_pred_type; /* 0..6 */
  int spherical;          /* 0 planar, 1 spherical */
  int evolve_base_state;  /* logical */
  int do_sponge;          /* logical */
  int do_eos_h_above_cutoff; /* logical; EOS work itself stays with the Fortran caller */
  int rho_comp, rhoh_comp, spec_comp, temp_comp, pi_comp, trac_comp; /* 1-based */
  int nspec, ntrac, nscal;
  int domlo[3], domhi[3]; /* problem domain, cell-centred */
  int nr;                 /* nr_fine: base state arrays have r = 0..nr (edge) / 0..nr-1 (cell) */
  double dt;
  double dx[3];
  double rel_eps;            /* variables.f90:17, set in estdt.f90:231 */
  double base_cutoff_density;
  /* used by the force builders of the L4 episodes (mkscalforce.f90 / mkforce.f90) */
  int base_cutoff_density_coord;  /* geometry: first r index with rho0 <= base_cutoff_density */
  double buoyancy_cutoff_factor;  /* _parameters */
  double omega, sin_theta, cos_theta, rotation_radius; /* geometry / probin: plane-parallel rotation */
} mgpu_params;

/* ---- lifecycle ------------------------------------------------------------------------- */
int mgpu_init(int device);            /* binds the calling rank to one GPU, creates streams/pool */
int mgpu_finalize(void);
int mgpu_synchronize(void);
const char* mgpu_last_error(void);
const char* mgpu_version(void);
/* number of kernel launches issued by the library since the last reset (bench "gpu_launches") */
long mgpu_launch_count(int reset);
/* tuning/testing switches: "fused" (1: fused 3-D edge kernel when applicable, 0: staged general path),
 * "kchunk" (z planes per CTA of the fused kernel), "exact" (1: bit-identical fp64 expression trees everywhere;
 * 0 (default): the FAST builds of the edge-state, BDS and mkutrans / velpred kernels -- dt/dx folded, divisions as
 * multiplications by reciprocals, FMA contraction -- <= 1e-12 relative from the reference), "premac_fuse" (1: on a box
 * without physical boundaries advance_premac forms utrans inside velpred's face kernel), "defaults" (value ignored:
 * every switch back to its initial value) */
int mgpu_set_option(const char* key, int value);
/* per-kernel-class device timing with CUDA events on the launching stream (bench roofline line):
 * mgpu_profile(1) starts/reset, mgpu_profile_get(tag,...) returns accumulated ms and launch count.
 * tags: 0 edge:cell-states 1 edge:simh 2 edge:transverse 3 edge:final 4 flux 5 update 6 fill 7 glue
 *       8 fused edge+flux+update 9 velpred 10 bds 11 halo */
int mgpu_profile(int on);
int mgpu_profile_get(int tag, double* ms, long* launches);
/* bytes moved host->device / device->host by host-pointer calls since the last reset (bench e2e accounting) */
int mgpu_copy_bytes(long* h2d, long* d2h, int reset);
/* raw cudaStream_t used for all launches (for CUDA-event timing by the caller) */
void* mgpu_stream(void);
/* run on a caller-owned stream instead (e.g. the framework's current stream) */
int mgpu_set_stream(void* stream);
/* pin caller-owned host memory (multifab data) so host-pointer calls copy at full PCIe rate */
int mgpu_host_register(double* hptr, long n);
int mgpu_host_unregister(double* hptr);

/* ---- residency registry (the multifabs of a MAESTRO step stay on the device between episodes) ----
 * A registered host fab (hptr = its first element, n = doubles, ghost cells and all components included) keeps one
 * device mirror.  Host-pointer calls then upload only the components whose host copy is newer and leave their
 * results on the device; mgpu_download brings the components a call wrote back to the host, mgpu_invalidate tells the
 * library that the host changed them (0-based comp0, ncomp < 0: every component).  Replaces what FBoxLib does
 * implicitly by owning the data: the Fortran driver calls mgpu_invalidate after it writes a multifab on the host and
 * mgpu_download before it reads one the device wrote.  pin != 0 also page-locks the host memory. */
int mgpu_register(double* hptr, long n, int pin);
int mgpu_unregister(double* hptr);
int mgpu_invalidate(double* hptr, int comp0, int ncomp);
int mgpu_download(double* hptr, int comp0, int ncomp);
int mgpu_upload(double* hptr, int comp0, int ncomp);

/* device memory helpers for device-resident episodes (tests/bench own their buffers) */
int mgpu_malloc(double** dptr, long n);
int mgpu_free(double* dptr);
int mgpu_memcpy_h2d(double* dst, const double* src, long n);
int mgpu_memcpy_d2h(double* dst, const double* src, long n);

/* ---- ghost fill (replaces FBoxLib multifab_fill_boundary + Source/multifab_physbc.f90:16) ---- */
/* Single-box-per-rank fill: periodic wrap inside the box for dims with pmask[d]=1, then physical
 * BCs per adv_bc (physbc_2d :150, physbc_3d :329) for comps scomp..scomp+ncomp-1. */
int mgpu_fill_boundary(const mgpu_params* p, mgpu_fab* s, int scomp, int bccomp, int ncomp,
                       const int* adv_bc, const int* pmask);
/* The same for a multifab of nfabs boxes on this rank (FBoxLib multifab_fill_boundary: same-level box-to-box copies,
 * periodic images included, then multifab_physbc on each box where it touches the domain boundary).  adv_bc is the
 * DOMAIN's table (bc_level index 0, define_bc_tower.f90:150-197); the per-box tables are derived from it. */
int mgpu_fill_boundary_mf(const mgpu_params* p, int nfabs, mgpu_fab* s, int scomp, int bccomp, int ncomp,
                          const int* adv_bc, const int* pmask);

/* ---- multi-GPU: one slab per rank along the slowest index (z in 3-D, y in 2-D); replaces the MPI ghost
 * exchange inside FBoxLib's multifab_fill_boundary (call sites: SURVEY.md section 2d) by NCCL send/recv between
 * slab neighbours.  Rank r owns the r-th slab; every fab carries its global lo/hi, params->domlo/domhi the
 * problem domain.  After mgpu_comm_init every ghost fill of the library (mgpu_fill_boundary and the fills
 * inside the L4 episodes) exchanges the slab-direction ghost planes with ranks r-1 / r+1 (periodic wrap where
 * pmask says so) and applies physical BCs only where adv_bc is not INTERIOR. ---- */
typedef struct {
  int dir;                    /* slab direction (dm-1) */
  int up_rank, dn_rank;       /* neighbour owning the next / previous slab; -1: none (physical boundary) */
  int nplanes;                /* ghost planes exchanged per side (= ng) */
  long plane_doubles;         /* doubles in one (x[,y]) plane of one component, ghost cells included */
  int send_up_k0, send_dn_k0; /* first plane index (slab direction, global) sent to up_rank / dn_rank */
  int recv_lo_k0, recv_hi_k0; /* first ghost plane index received from dn_rank / up_rank */
} mgpu_halo_plan;
/* pure host arithmetic (no GPU): the exchange plan of one fab */
int mgpu_halo_plan_make(const mgpu_params* p, const mgpu_fab* f, const int* pmask, int rank, int nranks,
                        mgpu_halo_plan* out);
int mgpu_comm_unique_id(void* out128);                              /* ncclGetUniqueId (rank 0) */
int mgpu_comm_init(int rank, int nranks, const void* unique_id128); /* ncclCommInitRank on the bound GPU */
int mgpu_comm_finalize(void);

/* ---- L3 operators ---------------------------------------------------------------------- */
/* make_edge_scal (Source/make_edge_scal.f90:26): edge states of comps start_scomp.. of s on all
 * faces; sedge[d] nodal in d, written in comps start_scomp...  umac[d] nodal in d with 1 ghost. */
int mgpu_make_edge_scal(const mgpu_params* p, int nfabs, const mgpu_fab* s, mgpu_fab* const* sedge,
                        const mgpu_fab* const* umac, const mgpu_fab* force, const int* adv_bc,
                        int is_vel, int start_scomp, int start_bccomp, int num_comp,
                        int is_conservative);

/* bds (Source/bds.f90:16): same contract as make_edge_scal, Bell-Dawson-Shubin reconstruction. */
int mgpu_bds(const mgpu_params* p, int nfabs, const mgpu_fab* s, mgpu_fab* const* sedge,
             const mgpu_fab* const* umac, const mgpu_fab* force, const int* adv_bc, int is_vel,
             int start_scomp, int start_bccomp, int num_comp, int is_conservative);

/* mk_rhoX_flux (Source/mkflux.f90:48), planar geometry (_2d :272, _3d_cart :370). */
int mgpu_mk_rhoX_flux(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, mgpu_fab* etarhoflux,
                      const mgpu_fab* const* sedge, const mgpu_fab* const* umac, const double* w0,
                      const double* rho0_old, const double* rho0_edge_old, const double* rho0_new,
                      const double* rho0_edge_new, const double* rho0_predicted_edge, int startcomp,
                      int endcomp);

/* mk_rhoh_flux (Source/mkflux.f90:652), planar geometry (_2d :920, _3d_cart :1070). */
int mgpu_mk_rhoh_flux(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux,
                      const mgpu_fab* const* sedge, const mgpu_fab* const* umac, const double* w0,
                      const double* rho0_old, const double* rho0_edge_old, const double* rho0_new,
                      const double* rho0_edge_new, const double* rhoh0_old,
                      const double* rhoh0_edge_old, const double* rhoh0_new,
                      const double* rhoh0_edge_new);

/* update_scal (Source/update_scal.f90:16), planar (_2d :246, _3d_cart :370) and spherical (_3d_sphr :508); valid
 * cells only, the caller follows with mgpu_fill_boundary (update_scal.f90:110-122).  With do_eos_h_above_cutoff and
 * nstart == rhoh_comp the zones with rho <= base_cutoff_density get rhoh from the EOS at (rho, p0_new, X)
 * (:421-447): p0_new is the level's 1-D array (planar) or p0_new_cart the fabs of put_1d_array_on_cart (spherical);
 * needs mgpu_set_eos -- without an EOS the call FAILS when a zone is below the cutoff (it never skips the reset
 * silently).  Both may be NULL for the other component ranges. */
int mgpu_update_scal(const mgpu_params* p, int nfabs, int nstart, int nstop, const mgpu_fab* sold,
                     mgpu_fab* snew, const mgpu_fab* const* sflux, const mgpu_fab* force, const double* p0_new,
                     const mgpu_fab* p0_new_cart);

/* update_velocity (Source/update_vel.f90:15), planar (_2d :174, _3d :227). */
int mgpu_update_velocity(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                         const mgpu_fab* const* umac, const mgpu_fab* const* uedge,
                         const mgpu_fab* force, const mgpu_fab* sponge, const double* w0);

/* addw0 (Source/addw0.f90:19), planar: umac[dm-1] += mult*w0 (no ghost exchange). */
int mgpu_addw0(const mgpu_params* p, int nfabs, mgpu_fab* const* umac, const double* w0, double mult);

/* mkutrans (Source/mkutrans.f90:17) planar (_2d :257, _3d :461). */
int mgpu_mkutrans(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                  mgpu_fab* const* utrans, const double* w0, const int* adv_bc, const int* phys_bc);

/* velpred (Source/velpred.f90:21) planar (_2d :266, _3d :640). */
int mgpu_velpred(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                 mgpu_fab* const* umac, const mgpu_fab* const* utrans, const mgpu_fab* force,
                 const double* w0, const int* adv_bc, const int* phys_bc);

/* glue used inside the drivers: modify_scal_force (modify_scal_force.f90:15, planar),
 * convert_rhoX_to_X (convert_rhoX_to_X.f90:20), put_in_pert_form (put_in_pert_form.f90:22, planar);
 * valid cells only, the caller follows with mgpu_fill_boundary. */
int mgpu_modify_scal_force(const mgpu_params* p, int nfabs, mgpu_fab* force, const mgpu_fab* s,
                           const mgpu_fab* const* umac, const double* s0, const double* s0_edge,
                           const double* w0, int comp, int fullform);
int mgpu_convert_rhoX_to_X(const mgpu_params* p, int nfabs, mgpu_fab* s, int flag);
int mgpu_put_in_pert_form(const mgpu_params* p, int nfabs, mgpu_fab* s, const double* base, int comp,
                          int flag);

/* ---- spherical geometry (SURVEY a3/a4/a7-a12 "_3d_sphr" branches, and 8f2) ------------------------------
 * geometry module (Source/geometry.f90: center, dr(1), nr_fine, r_cc_loc(1,:), r_edge_loc(1,:)) and the
 * interpolation switches of probin (Source/_parameters:632-651), passed per call like mgpu_params. */
typedef struct {
  double center[3];
  double prob_lo[3];
  double dr;                /* dr(1) = dx/drdxfac */
  int nr_fine;
  const double* r_cc_loc;   /* (0:nr_fine-1) host */
  const double* r_edge_loc; /* (0:nr_fine)   host */
  int s0_interp_type;       /* 1 constant, 2 linear, 3 quadratic: bin-centred 1-D array -> cell centres */
  int w0_interp_type;       /* same for an edge-centred 1-D array */
  int s0mac_interp_type;    /* 1 via cell centres then average, 2 linear to faces, 3 quadratic to faces */
  int w0mac_interp_type;    /* 1, 2, 3 as above, 4 via nodes */
} mgpu_geom;

/* put_1d_array_on_cart_3d_sphr (Source/fill_3d_data.f90:269): valid cells lo:hi of s0_cart (1 comp, or 3 comps
 * = value * unit radial vector if is_output_a_vector); ghost cells are the caller's mgpu_fill_boundary. */
int mgpu_put_1d_array_on_cart(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0,
                              mgpu_fab* s0_cart, int is_input_edge_centered, int is_output_a_vector);
/* make_w0mac_3d_sphr (fill_3d_data.f90:621): normal component of w0 on faces lo-1:hi+1(+1), ng_w0 = 1.
 * w0_cart (3 comps, ng >= 2, ghosts filled) is read only when w0mac_interp_type == 1. */
int mgpu_make_w0mac(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* w0,
                    mgpu_fab* const* w0mac, const mgpu_fab* w0_cart);
/* make_s0mac_3d_sphr (fill_3d_data.f90:1017): s0_cart (1 comp, ng >= 2) read only when s0mac_interp_type == 1. */
int mgpu_make_s0mac(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0,
                    mgpu_fab* const* s0mac, const mgpu_fab* s0_cart);
/* addw0_3d_sphr (Source/addw0.f90:171): umac_d += mult * w0mac_d on the valid faces of every direction. */
int mgpu_addw0_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                    double mult);
/* mk_rhoX_flux_3d_sphr (Source/mkflux.f90:509) and mk_rhoh_flux_3d_sphr (:1289). */
int mgpu_mk_rhoX_flux_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                           const mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                           const mgpu_fab* const* rho0mac_old, const mgpu_fab* const* rho0mac_new, int startcomp,
                           int endcomp);
int mgpu_mk_rhoh_flux_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                           const mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                           const mgpu_fab* const* rho0mac_old, const mgpu_fab* const* rho0mac_new,
                           const mgpu_fab* const* h0mac_old, const mgpu_fab* const* h0mac_new);
/* update_velocity_3d, spherical branch (Source/update_vel.f90:227, :317-360). */
int mgpu_update_velocity_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                              const mgpu_fab* const* umac, const mgpu_fab* const* uedge, const mgpu_fab* force,
                              const mgpu_fab* sponge, const mgpu_fab* const* w0mac);
/* mkutrans_3d / velpred_3d with spherical == 1: every Riemann problem carries w0mac of its direction
 * (mkutrans.f90:601,709,817; velpred.f90:1588,1687,1786). */
int mgpu_mkutrans_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                       mgpu_fab* const* utrans, const mgpu_fab* const* w0mac, const int* adv_bc,
                       const int* phys_bc);
int mgpu_velpred_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                      mgpu_fab* const* umac, const mgpu_fab* const* utrans, const mgpu_fab* force,
                      const mgpu_fab* const* w0mac, const int* adv_bc, const int* phys_bc);
/* modify_scal_force_3d_sphr (Source/modify_scal_force.f90:256) and pert_form_3d_sphr
 * (Source/put_in_pert_form.f90:185).  s0_cart: 1 comp, ng >= 1 with ghosts filled. */
int mgpu_modify_scal_force_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* force,
                                const mgpu_fab* s, const mgpu_fab* const* umac, const mgpu_fab* s0_cart,
                                const double* w0, int comp, int fullform);
int mgpu_put_in_pert_form_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* s,
                               const double* s0, int comp, int flag);

/* density_advance with spherical == 1 as one device-resident episode (Source/density_advance.f90:20): rho0_old_cart and
 * the rho0mac arrays are built on the device (put_1d_array_on_cart + ghost fill, make_s0mac); w0mac is the caller's
 * (make_w0mac); no etarhoflux in spherical geometry.  rho0_old / rho0_new: radial arrays (0:nr_fine-1). */
int mgpu_density_advance_sphr(const mgpu_params* p, const mgpu_geom* g, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                              mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                              mgpu_fab* const* umac, const double* w0, const mgpu_fab* const* w0mac,
                              const double* rho0_old, const double* rho0_new, const int* adv_bc, const int* pmask);

/* make_normal (fill_3d_data.f90:1280): the unit radial vector on the cell centres, ghost cells included. */
int mgpu_make_normal(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* normal);

/* mk_vel_force with spherical == 1 (mkforce.f90:22 -> mk_vel_force_3d_sphr :484), valid cells; the caller fills the
 * ghost cells (mkforce.f90:209: bc comps 1..dm).  w0 is the radial array (0:nr_fine); w0_cart and grad w0 on the cell
 * centres are built inside, as the reference's wrapper does (:92-127).  s(index_rho) is the density, normal and
 * w0_force_cart the cell-centred fabs of make_normal / put_1d_array_on_cart(w0_force, vector). */
int mgpu_mk_vel_force_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* vel_force, int is_final_update,
                           const mgpu_fab* uold, const mgpu_fab* const* uedge, const double* w0,
                           const mgpu_fab* const* w0mac, const mgpu_fab* gpi, const mgpu_fab* s, int index_rho,
                           const mgpu_fab* normal, const double* rho0, const double* grav, const mgpu_fab* w0_force_cart,
                           int do_add_utilde_force);

/* advance_premac (advance_premac.f90:21) and velocity_advance (velocity_advance.f90:16) with spherical == 1 as
 * device-resident episodes; argument lists of the Fortran drivers (normal, w0mac, w0_force_cart included). */
int mgpu_advance_premac_sphr(const mgpu_params* p, const mgpu_geom* g, const mgpu_fab* uold, const mgpu_fab* sold,
                             mgpu_fab* const* umac, const mgpu_fab* gpi, const mgpu_fab* normal, const double* w0,
                             const mgpu_fab* const* w0mac, const mgpu_fab* w0_force_cart, const double* rho0_old,
                             const double* grav_cell_old, const int* adv_bc, const int* phys_bc, const int* pmask);
int mgpu_velocity_advance_sphr(const mgpu_params* p, const mgpu_geom* g, const mgpu_fab* uold, mgpu_fab* unew,
                               const mgpu_fab* sold, const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi,
                               const mgpu_fab* normal, const double* w0, const mgpu_fab* const* w0mac,
                               const mgpu_fab* w0_force_cart, const double* rho0_old, const double* rho0_nph,
                               const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                               const int* adv_bc, const int* pmask);

/* mkrhohforce with spherical == 1 (mkscalforce.f90:31 -> mkrhohforce_3d_sphr :388): p0 = (p0_1 + p0_2)/2 is put on the
 * cell centres and on the faces inside (put_1d_array_on_cart, make_s0mac), psi on the cell centres; writes the rhoh
 * component of scal_force, ghost cells included (:177-181). */
int mgpu_mkrhohforce_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* scal_force, int is_prediction,
                          const mgpu_fab* thermal, const mgpu_fab* const* umac, const double* p0_1, const double* p0_2,
                          const double* psi, int add_thermal, const int* adv_bc, const int* pmask);
/* enthalpy_advance (enthalpy_advance.f90:16) with spherical == 1 as a device-resident episode: rhoh0_old_cart, the
 * rho0mac / h0mac arrays and p0 on cells and faces are built on the device; radial arrays are (0:nr_fine-1). */
int mgpu_enthalpy_advance_sphr(const mgpu_params* p, const mgpu_geom* g, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                               mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                               const mgpu_fab* thermal, mgpu_fab* const* umac, const double* w0,
                               const mgpu_fab* const* w0mac, const double* rho0_old, const double* rhoh0_old,
                               const double* rho0_new, const double* rhoh0_new, const double* p0_old, const double* p0_new,
                               const double* tempbar, const double* psi, const int* adv_bc, const int* pmask);

/* ---- L4 episode: density_advance (Source/density_advance.f90:20), planar, one level ------
 * Signature mirrors the Fortran argument list; sold is modified in place exactly as the reference
 * does (rhoX->X->rhoX, rho->rho'->rho round trips), umac is (umac+w0)-w0 on return, sedge, sflux,
 * etarhoflux, scal_force and snew (valid + ghost cells) are fully populated on return.
 * Exception (FAST build, 3-D, fully periodic domain: the "lean+" episode): the input transforms
 * happen inside the edge kernel, so sold and umac come back UNTOUCHED -- one rounding away from the reference's round
 * trips (rhoX/rho)*rho and (umac+w0)-w0; option exact = 1 or leanplus = 0 restores the literal sequence.
 * pmask[d]=1 => periodic in d (FBoxLib layout pmask); single box per rank (slab).
 * spherical == 1: call mgpu_density_advance_sphr (this entry raises). */
int mgpu_density_advance(const mgpu_params* p, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                         mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                         mgpu_fab* const* umac, const double* w0, mgpu_fab* etarhoflux,
                         const double* rho0_old, const double* rho0_new, const double* p0_dummy,
                         const double* rho0_predicted_edge, const int* adv_bc, const int* pmask);
/* density_advance over a multifab of nfabs boxes on this rank (the reference's test_advect lays 64^3 out as 8 x 32^3,
 * Exec/UNIT_TESTS/test_advect/gr0_3d): arrays of nfabs fabs, sedge / sflux / umac dm arrays of nfabs fabs, adv_bc the
 * domain's table.  General path of density_advance.f90:20, every ghost fill through the multifab fill above. */
int mgpu_density_advance_mf(const mgpu_params* p, int which_step, int nfabs, mgpu_fab* sold, mgpu_fab* snew,
                            mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force, mgpu_fab* const* umac,
                            const double* w0, mgpu_fab* etarhoflux, const double* rho0_old, const double* rho0_new,
                            const double* p0_dummy, const double* rho0_predicted_edge, const int* adv_bc,
                            const int* pmask);

/* The other three planar episodes over the nfabs boxes of this rank (arrays of nfabs fabs; umac / sedge / sflux: dm arrays
 * of nfabs fabs; adv_bc / phys_bc: the DOMAIN's tables, the per-box ones are derived): the single-box episodes stage by
 * stage, every ghost fill through the multifab fill.  enthalpy_advance: predict_rhoh / predict_rhohprime / predict_h. */
int mgpu_velocity_advance_mf(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew, const mgpu_fab* sold,
                             const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0,
                             const double* w0_force, const double* rho0_old, const double* rho0_nph,
                             const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                             const int* adv_bc, const int* pmask);
int mgpu_advance_premac_mf(const mgpu_params* p, int nfabs, const mgpu_fab* uold, const mgpu_fab* sold,
                           mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0, const double* w0_force,
                           const double* rho0_old, const double* grav_cell_old, const int* adv_bc, const int* phys_bc,
                           const int* pmask);
int mgpu_enthalpy_advance_mf(const mgpu_params* p, int which_step, int nfabs, mgpu_fab* sold, mgpu_fab* snew,
                             mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                             const mgpu_fab* thermal, mgpu_fab* const* umac, const double* w0, const double* rho0_old,
                             const double* rhoh0_old, const double* rho0_new, const double* rhoh0_new,
                             const double* p0_old, const double* p0_new, const double* psi, const double* grav_old,
                             const double* grav_nph, const int* adv_bc, const int* pmask);

/* ---- force builders inside the L4 drivers ----------------------------------------
 * mkrhohforce (Source/mkscalforce.f90:31; _2d :249, _3d :310), planar: writes comp rhoh_comp of scal_force on the
 * valid cells (the caller follows with mgpu_fill_boundary, mkscalforce.f90:177).  grav is the 1-D array
 * make_grav_cell gives for rho0 = (rho0_1+rho0_2)/2 (host base-state work, stays with the Fortran). */
int mgpu_mkrhohforce(const mgpu_params* p, int nfabs, mgpu_fab* scal_force, int is_prediction,
                     const mgpu_fab* thermal, const mgpu_fab* const* umac, const double* p0_1, const double* p0_2,
                     const double* rho0_1, const double* rho0_2, const double* grav, const double* psi,
                     int add_thermal);
/* mk_vel_force (Source/mkforce.f90:22; _2d :283, _3d_cart :342), planar: all dm comps of vel_force (zeroed incl.
 * ghosts, then valid cells); rho = comp index_rho (1-based) of s; the caller follows with mgpu_fill_boundary. */
int mgpu_mk_vel_force(const mgpu_params* p, int nfabs, mgpu_fab* vel_force, int is_final_update,
                      const mgpu_fab* uold, const mgpu_fab* const* uedge, const double* w0, const mgpu_fab* gpi,
                      const mgpu_fab* s, int index_rho, const double* rho0, const double* grav,
                      const double* w0_force, int do_add_utilde_force);

/* ---- L4 episodes, planar, one level, one box (slab) per rank; signatures mirror the Fortran argument lists
 * (multifab -> mgpu_fab, base-state (n,0:nr) arrays -> the level's 1-D slice).  Temporaries the reference
 * builds and destroys inside the driver (ufull, utrans, force, uedge) are device-only.
 * advance_premac (Source/advance_premac.f90:21): umac valid faces are written (no ghost fill, as the reference). */
int mgpu_advance_premac(const mgpu_params* p, const mgpu_fab* uold, const mgpu_fab* sold, mgpu_fab* const* umac,
                        const mgpu_fab* gpi, const double* w0, const double* w0_force, const double* rho0_old,
                        const double* grav_cell_old, const int* adv_bc, const int* phys_bc, const int* pmask);
/* velocity_advance (Source/velocity_advance.f90:16): unew valid + ghost cells; umac is (umac+w0)-w0 on return. */
int mgpu_velocity_advance(const mgpu_params* p, const mgpu_fab* uold, mgpu_fab* unew, const mgpu_fab* sold,
                          const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0,
                          const double* w0_force, const double* rho0_old, const double* rho0_nph,
                          const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                          const int* adv_bc, const int* pmask);
/* enthalpy_advance (Source/enthalpy_advance.f90:16), every enthalpy_pred_type the reference codes (predict_hprime
 * raises, mkflux.f90:1167).  The temperature-based types (predict_T_then_rhohprime, predict_T_then_h,
 * predict_Tprime_then_h: mktempforce, T or T' edge states, makeHfromRhoT_edge) and the reset of rhoh below the cutoff
 * (update_scal.f90:421-447, do_eos_h_above_cutoff) call the EOS: they need mgpu_set_eos and fail without it.  tempbar:
 * the level's 1-D array (may be NULL for the other types).  grav_old / grav_nph: make_grav_cell of rho0_old /
 * (rho0_old+rho0_new)/2.  sedge(rho_comp) and the species edge states must hold what density_advance left
 * (mkflux.f90:1126, rhoh_vs_t.f90:437-461). */
int mgpu_enthalpy_advance(const mgpu_params* p, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                          mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                          const mgpu_fab* thermal, mgpu_fab* const* umac, const double* w0, const double* rho0_old,
                          const double* rhoh0_old, const double* rho0_new, const double* rhoh0_new,
                          const double* p0_old, const double* p0_new, const double* tempbar, const double* psi,
                          const double* grav_old, const double* grav_nph, const int* adv_bc, const int* pmask);

/* ---- consumers / producers next to the path, planar ---------------------------
 * estdt (Source/estdt.f90:29; per box _2d :348, _3d_cart :467): the time-step limits of one level from the advective
 * speeds (u, w0), the velocity force (the caller builds it with mgpu_mk_vel_force, estdt.f90:117-120), the divU
 * and the dS/dt constraints.  All reductions are max / min, so the result is bit-identical to the reference's
 * whatever the order.  The boxes of this rank are reduced like the reference's loop over fabs (:148-200), then
 * across the ranks of the slab run (NCCL min/max, the reference's parallel_reduce :202-203), then the
 * "protect against huge time steps" rule (:209-217).  On return *dt = min(*dt, dt_lev), *umax = max(*umax,
 * umax_lev); the caller sets rel_eps = 1.d-8*umax (:229).  rho_min = 1.d-20 (:80) is passed explicitly. */
int mgpu_estdt(const mgpu_params* p, int nfabs, const mgpu_fab* u, const mgpu_fab* s, const mgpu_fab* force,
               const mgpu_fab* divU, const mgpu_fab* dSdt, const double* w0, const double* p0,
               const double* gamma1bar, double rho_min, double cflfac, double* dt, double* umax);
/* multifab_min_c / multifab_max_c (FBoxLib multifab_f.f90) over the valid zones of component `comp` (1-based) of
 * one level, reduced over the ranks: the `verbose >= 1` report of density_advance.f90:374-402 (species as X = rhoX / rho:
 * pass div_comp = rho_comp; 0: no division; the state is left untouched where the reference divides and multiplies
 * back), enthalpy_advance.f90:440-449 and velocity_advance.f90:142-160.  The caller prints the lines (formats 2000-2003,
 * 1001-1003 of those files). */
int mgpu_minmax(const mgpu_params* p, int nfabs, const mgpu_fab* s, int comp, int div_comp, double* smin, double* smax);

/* estdt for spherical geometry (estdt_3d_sphr, Source/estdt.f90:620): w0mac = make_w0mac's face fabs (zero when
 * evolve_base_state is off, :100-110); gp0 (:734-739) and its put_1d_array_on_cart (:741) happen inside. */
int mgpu_estdt_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* u, const mgpu_fab* s,
                    const mgpu_fab* force, const mgpu_fab* divU, const mgpu_fab* dSdt,
                    const mgpu_fab* const* w0mac, const double* w0, const double* p0, const double* gamma1bar,
                    double rho_min, double cflfac, double* dt, double* umax);
/* make_etarho_planar (Source/make_eta.f90:36; sum_etarho_2d :176, _3d :213): plane averages of etarhoflux,
 * etarho_ec(0:nr) on edges and etarho_cc(0:nr-1) = their two-point means.  The sums of one rank's planes are
 * combined over the ranks (NCCL sum, the reference's parallel_reduce :101); ncell = cells per plane of the whole
 * domain (p->domlo/domhi).  Floating-point sums: the order differs from the reference's loop, parity is 1e-12
 * relative (the reference's own MPI reduction order is not fixed either). */
int mgpu_make_etarho_planar(const mgpu_params* p, int nfabs, const mgpu_fab* etarhoflux, double* etarho_ec,
                            double* etarho_cc);

/* average (Source/average.f90:24) of component incomp (1-based) of one level's fabs into phibar(0:nr-1).  Planar: the
 * mean of every plane (sums over the ranks, NCCL).  Spherical (g != NULL): the reference's binning by the exact set of
 * radii a cell centre can have (sum_phi_3d_sphr :564), then quadratic interpolation onto the base-state radii (:312-352),
 * for one level (max_levs = 1); nr_irreg (geometry, initialize.f90:1272) and drdxfac (probin) are passed explicitly.
 * Floating-point sums in a different order than the reference's loops: parity is 1e-12 relative. */
int mgpu_average(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* phi, int incomp, int nr_irreg,
                 int drdxfac, double* phibar);
/* make_etarho_spherical (Source/make_eta.f90:256): eta_cart = [rho' (U . e_r)] at the half time (construct_eta_cart
 * :345), its average -> etarho_cc(0:nr_fine-1), and etarho_ec(0:nr_fine) on the edges (:337-343). */
int mgpu_make_etarho_spherical(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* sold,
                               const mgpu_fab* snew, const mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                               const double* rho0_old, const double* rho0_new, const mgpu_fab* normal, int nr_irreg,
                               int drdxfac, double* etarho_ec, double* etarho_cc);

/* ---- EOS -----------------------------------------------------------------------------
 * The generic front end of Microphysics/EOS/eos.F90:99 (composition eos_type.f90:157, reset_inputs eos.F90:129,
 * eos_reset :291) over gamma_law_general (Microphysics/EOS/gamma_law_general/gamma_law_general.f90:60).  Like the
 * reference's eos_init (eos.F90:26) the EOS is process state, set once; mgpu_set_eos(NULL) unsets it.  The network's
 * aion / zion are not part of MAESTRO's tree (they come with the problem's network), so the caller passes them; k_B
 * and n_A are the values of Source/constants_cgs.f90:15,24.  The entropy of gamma_law_general (Sackur-Tetrode, :177)
 * is not evaluated -- nothing on the advective path reads it -- so eos_input_ps raises, like eos_input_ph / _th do in
 * the reference (:147-160).  A tabulated EOS (helmeos) is out of scope. */
#define MGPU_EOS_MAXSPEC 32
enum { MGPU_EOS_NONE = 0, MGPU_EOS_GAMMA_LAW = 1 };
enum { /* eos_type.f90:8-15 */
  MGPU_EOS_INPUT_RT = 1, MGPU_EOS_INPUT_RH = 2, MGPU_EOS_INPUT_TP = 3, MGPU_EOS_INPUT_RP = 4, MGPU_EOS_INPUT_RE = 5,
  MGPU_EOS_INPUT_PS = 6, MGPU_EOS_INPUT_PH = 7, MGPU_EOS_INPUT_TH = 8
};
typedef struct {
  int kind;           /* MGPU_EOS_GAMMA_LAW */
  int assume_neutral; /* eos_assume_neutral (gamma_law_general/_parameters) */
  int nspec;
  int pad_;
  double gamma;       /* eos_gamma */
  double k_B, n_A;    /* constants_cgs.f90:15,24: 1.3806488e-16, 6.02214129e23 */
  double mintemp, maxtemp, mindens, maxdens, mine, maxe, minp, maxp, minh, maxh; /* eos_type.f90:42-57 */
  double small_temp;  /* probin small_temp (rhoh_vs_t.f90:431) */
  double aion[MGPU_EOS_MAXSPEC], zion[MGPU_EOS_MAXSPEC]; /* network */
} mgpu_eos;
int mgpu_set_eos(const mgpu_eos* e);
/* eos(input, state) at n points, host arrays.  state is (n, MGPU_EOS_NQ) point-fastest: state[q*n + i]; xn is
 * (n, nspec) point-fastest.  Inputs are read from the columns the input mode names, every column is written. */
enum {
  MGPU_EOS_Q_RHO = 0, MGPU_EOS_Q_T, MGPU_EOS_Q_P, MGPU_EOS_Q_E, MGPU_EOS_Q_H, MGPU_EOS_Q_CV, MGPU_EOS_Q_CP,
  MGPU_EOS_Q_CS, MGPU_EOS_Q_DPDT, MGPU_EOS_Q_DPDR, MGPU_EOS_Q_DEDT, MGPU_EOS_Q_DEDR, MGPU_EOS_Q_DHDT, MGPU_EOS_Q_MU,
  MGPU_EOS_Q_ABAR, MGPU_EOS_Q_ZBAR, MGPU_EOS_NQ
};
int mgpu_eos_eval(int input, long n, double* state, const double* xn);

/* makeHfromRhoT_edge (Source/rhoh_vs_t.f90:20; _2d :252, _3d_cart :392): the rhoh component of the edge states from
 * the predicted T (or T') and the density / species edge states, per enthalpy_pred_type and species_pred_type.  The
 * base-state arrays are the level's 1-D slices (cell (0:nr-1), edge (0:nr)).  QUIRK kept: the 3-D Cartesian y and z
 * faces read T' from the x-face array (:487, :543). */
int mgpu_make_h_from_rhot_edge(const mgpu_params* p, int nfabs, mgpu_fab* const* sedge, const double* rho0_old,
                               const double* rhoh0_old, const double* t0_old, const double* rho0_edge_old,
                               const double* rhoh0_edge_old, const double* t0_edge_old, const double* rho0_new,
                               const double* rhoh0_new, const double* t0_new, const double* rho0_edge_new,
                               const double* rhoh0_edge_new, const double* t0_edge_new);
/* the same with spherical == 1 (wrapper :84-105 + _3d_sphr :596): the half-time base state is put on the cell centres
 * (2 ghost cells, put_1d_array_on_cart with its boundary fill) inside. */
int mgpu_make_h_from_rhot_edge_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* const* sedge,
                                    const double* rho0_old, const double* rhoh0_old, const double* t0_old,
                                    const double* rho0_new, const double* rhoh0_new, const double* t0_new,
                                    const int* adv_bc, const int* pmask);
/* mktempforce (Source/mkscalforce.f90:719; _2d :896, _3d :954, _3d_sphr :1026): comp temp_comp of temp_force on the
 * valid cells, then its ghost cells (foextrap_comp, :833-837).  g may be NULL for planar geometry. */
int mgpu_mktempforce(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* temp_force,
                     const mgpu_fab* const* umac, const mgpu_fab* s, const mgpu_fab* thermal, const double* p0_old,
                     const double* psi, const int* adv_bc, const int* pmask);
/* firstdt (Source/firstdt.f90:25; _2d :330, _3d :460, _3d_sphr :599) for one level: the velocity force is built inside
 * like the reference does (:96-100, mk_vel_force with zero w0 and no utilde force); on return *dt = min(*dt,
 * dt_lev*init_shrink) and *umax = max(*umax, umax_lev) (:171-190); the small_dt / max_dt / fixed_dt rules (:199-216)
 * stay with the caller.  g may be NULL for planar geometry. */
int mgpu_firstdt(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* u, const mgpu_fab* gpi,
                 const mgpu_fab* s, const mgpu_fab* divU, const double* rho0, const double* p0, const double* grav,
                 const double* gamma1bar, double cflfac, double init_shrink, int use_soundspeed_firstdt,
                 int use_divu_firstdt, double* dt, double* umax);
/* makeTfromRhoH / makeTfromRhoP (Source/rhoh_vs_t.f90:800, :1165), planar and spherical: the temperature component
 * of the state on the valid cells from (rho, h, X) or (rho, p0, X), then its ghost cells (:846-852, :1224-1230).
 * use_eos_e_instead_of_h / use_pprime_in_tfromp: the probin switches of the same names.  p0: the level's 1-D array;
 * with spherical == 1 it is put on the cell centres inside (:1097, :1410).  update_rhoh (makeTfromRhoP only): also reset
 * rhoh = rho * h(rho, p0, X). */
int mgpu_make_t_from_rhoh(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0,
                          int use_eos_e_instead_of_h, const int* adv_bc, const int* pmask);
int mgpu_make_t_from_rhop(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0,
                          int update_rhoh, int use_pprime_in_tfromp, const int* adv_bc, const int* pmask);

#ifdef __cplusplus
}
#endif
#endif /* MAESTRO_B200_H */
