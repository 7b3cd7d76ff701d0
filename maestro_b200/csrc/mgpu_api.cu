// C ABI of the library (include/maestro_b200.h): context, host staging, and the L3/L4 entry points.
//
// Host-pointer calls (mem_space == MGPU_HOST) are what the Fortran shim issues: inputs are copied to
// the device, the kernels run, outputs are copied back and the call returns synchronised, so the
// unchanged Fortran sees exactly what the reference routine would have left in its multifabs.
// Device-pointer calls (MGPU_DEVICE) skip the copies and stay asynchronous on the library stream.
// There is no CPU fallback anywhere: without a CUDA device every entry point fails.
#include <algorithm>
#include <cstring>
#include <map>

#include "mgpu_bds.cuh"
#include "mgpu_edge.cuh"
#include "mgpu_fused.cuh"
#include "mgpu_halo.cuh"
#include "mgpu_stream.cuh"
#include "mgpu_velpred.cuh"

namespace mgpu {

static Context g_ctx;
static std::string g_err;
static int g_opt_fused = 1;    // use the fused 3-D edge kernel when it covers the case
static int g_opt_kchunk = 32;  // z planes per CTA of the fused kernel
static int g_opt_exact = 0;    // 1: bit-identical arithmetic everywhere (fused kernel built with -fmad=false)
Context& ctx() { return g_ctx; }

void require_init() {
  if (!g_ctx.initialised) throw Error("mgpu: library not initialised (call mgpu_init; a CUDA device is required)");
}

// ---- optional kernel-class profiling --------------------------------------------------------------
struct Prof {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[TAG_COUNT];
  size_t used[TAG_COUNT] = {0};
  cudaEvent_t cur = nullptr;
};
static Prof g_prof;
void prof_begin(int tag) {
  if (!g_prof.on) return;
  auto& v = g_prof.ev[tag];
  if (g_prof.used[tag] == v.size()) {
    cudaEvent_t a, b;
    MGPU_CUDA(cudaEventCreate(&a));
    MGPU_CUDA(cudaEventCreate(&b));
    v.push_back({a, b});
  }
  MGPU_CUDA(cudaEventRecord(v[g_prof.used[tag]].first, g_ctx.stream));
}
void prof_end(int tag) {
  if (!g_prof.on) return;
  MGPU_CUDA(cudaEventRecord(g_prof.ev[tag][g_prof.used[tag]].second, g_ctx.stream));
  g_prof.used[tag]++;
}

// ---- arena -------------------------------------------------------------------------------------
void arena_reserve(size_t bytes) {
  Context& c = g_ctx;
  if (bytes <= c.arena_bytes) return;
  MGPU_CUDA(cudaStreamSynchronize(c.stream));
  if (c.arena) MGPU_CUDA(cudaFree(c.arena));
  c.arena = nullptr;
  c.arena_bytes = 0;
  size_t want = bytes + bytes / 8 + (1u << 20);
  MGPU_CUDA(cudaMalloc((void**)&c.arena, want));
  c.arena_bytes = want;
  c.arena_used = 0;
}
void arena_reset() { g_ctx.arena_used = 0; }
size_t arena_mark() { return g_ctx.arena_used; }
void arena_release(size_t mark) { g_ctx.arena_used = mark; }
double* arena_alloc(size_t nd) {
  Context& c = g_ctx;
  size_t bytes = (nd * sizeof(double) + 255) & ~(size_t)255;
  if (c.arena_used + bytes > c.arena_bytes) throw Error("mgpu: scratch arena overflow (internal sizing error)");
  double* p = (double*)(c.arena + c.arena_used);
  c.arena_used += bytes;
  return p;
}
template <class T>
T* upload_small(const T* h, size_t n) {
  if (!h) return nullptr;
  T* d = (T*)arena_alloc((n * sizeof(T) + 7) / 8);
  MGPU_CUDA(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, g_ctx.stream));
  return d;
}
template double* upload_small<double>(const double*, size_t);
template int* upload_small<int>(const int*, size_t);

// ---- staging pool for host-pointer calls -------------------------------------------------------
struct Pool {
  std::multimap<size_t, double*> free_;
  std::map<double*, size_t> live_;
  double* get(size_t n) {
    auto it = free_.lower_bound(n);
    if (it != free_.end() && it->first <= n + n / 4) {
      double* p = it->second;
      live_[p] = it->first;
      free_.erase(it);
      return p;
    }
    double* p = nullptr;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(double));
    if (e != cudaSuccess) {  // drop the cache and retry once
      cudaGetLastError();
      for (auto& kv : free_) cudaFree(kv.second);
      free_.clear();
      MGPU_CUDA(cudaMalloc((void**)&p, n * sizeof(double)));
    }
    live_[p] = n;
    return p;
  }
  void put(double* p) {
    auto it = live_.find(p);
    if (it == live_.end()) return;
    free_.insert({it->second, p});
    live_.erase(it);
  }
  void clear() {
    for (auto& kv : free_) cudaFree(kv.second);
    for (auto& kv : live_) cudaFree(kv.first);
    free_.clear();
    live_.clear();
  }
};
static Pool g_pool;

// One entry-point invocation: maps fabs to device views, copies in/out for host-pointer calls.
struct Call {
  const mgpu_params& P;
  bool host;
  struct Item { double* h; double* d; size_t n; bool out; };
  std::vector<Item> items;
  explicit Call(const mgpu_params* p, size_t scratch_bytes) : P(*p), host(p->mem_space == MGPU_HOST) {
    require_init();
    if (P.dm != 2 && P.dm != 3) throw Error("mgpu: dm must be 2 or 3 (1-D stays with the Fortran)");
    arena_reserve(scratch_bytes + (1u << 20));
    arena_reset();
  }
  ~Call() {
    for (auto& it : items) g_pool.put(it.d);
  }
  DV view(const mgpu_fab& f, bool copy_in, bool copy_out) {
    if (!f.ptr) throw Error("mgpu: null fab pointer");
    if (!host) return make_view(f, P.dm);
    for (auto& it : items)  // the same host fab passed twice maps to one device buffer
      if (it.h == f.ptr) {
        it.out = it.out || copy_out;
        return make_view(f, P.dm, it.d);
      }
    DV v = make_view(f, P.dm);
    size_t n = (size_t)v.size();
    double* d = g_pool.get(n);
    if (copy_in) MGPU_CUDA(cudaMemcpyAsync(d, f.ptr, n * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
    items.push_back({f.ptr, d, n, copy_out});
    return make_view(f, P.dm, d);
  }
  void views(const mgpu_fab* const* f, int i, bool in, bool out, DV* v) {
    for (int d = 0; d < P.dm; ++d) v[d] = view(f[d][i], in, out);
  }
  void finish() {
    if (!host) return;
    for (auto& it : items)
      if (it.out) MGPU_CUDA(cudaMemcpyAsync(it.h, it.d, it.n * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  }
};

// cell_to_edge (Source/cell_to_edge.f90:14) for one level with a single chunk: O(nr) host work on the
// base state, exactly as the reference does it on the host.
static void cell_to_edge_host(const double* c, double* e, int nr) {
  for (int r = 0; r <= nr; ++r) {
    if (r == 0) e[r] = c[r];
    else if (r == 1) e[r] = 0.5 * (c[r - 1] + c[r]);
    else if (r == nr) e[r] = c[r - 1];
    else if (r == nr - 1) e[r] = 0.5 * (c[r] + c[r - 1]);
    else {
      double tmp = 7.0 / 12.0 * (c[r] + c[r - 1]) - 1.0 / 12.0 * (c[r + 1] + c[r - 2]);
      double mn = c[r] < c[r - 1] ? c[r] : c[r - 1], mx = c[r] > c[r - 1] ? c[r] : c[r - 1];
      double t2 = tmp > mn ? tmp : mn;
      e[r] = t2 < mx ? t2 : mx;
    }
  }
}

static void fill_flux_args(const mgpu_params& P, FluxArgs& a, const int* lo, const int* hi) {
  memset(&a, 0, sizeof(a));
  a.dm = P.dm;
  a.species_pred_type = P.species_pred_type;
  a.evolve_base_state = P.evolve_base_state != 0;
  a.rho = P.rho_comp - 1;
  a.rhoh = P.rhoh_comp - 1;
  a.spec0 = P.spec_comp - 1;
  a.nspec = P.nspec;
  a.vb = grown(lo, hi, P.dm, 0);
}

// make_edge_scal for one component: fused single-launch kernel when it covers the case, else the staged path
static void edge_one_comp(const mgpu_params& P, const DV& s, DV* sedge, const DV* umac, const DV& force,
                          const int* lo, const int* hi, const int* adv_bc, int comp, int bccomp, bool is_vel,
                          bool is_cons, int ng_s, int ng_f) {
  for (int d = 0; d < P.dm; ++d) {  // validate BC codes up front (make_edge_scal.f90:853)
    for (int side = 0; side < 2; ++side) {
      const int bc = adv_bc[d + P.dm * (side + 2 * (bccomp - 1))];
      if (bc != MGPU_BC_EXT_DIR && bc != MGPU_BC_FOEXTRAP && bc != MGPU_BC_HOEXTRAP && bc != MGPU_BC_REFLECT_EVEN &&
          bc != MGPU_BC_REFLECT_ODD && bc != MGPU_BC_INTERIOR)
        throw Error("make_edge_scal: invalid boundary type adv_bc");
    }
  }
  if (g_opt_fused && fused_edge_supported(P, is_cons)) {
    fused_edge_dev(P, s, sedge, umac, force, lo, hi, adv_bc, comp, bccomp, is_vel, ng_s, ng_f, g_opt_kchunk,
                   g_opt_exact != 0);
  } else {
    size_t mark = arena_mark();
    make_edge_scal_dev(P, s, sedge, umac, force, lo, hi, adv_bc, comp, bccomp, is_vel, is_cons, ng_s, ng_f);
    arena_release(mark);
  }
}

// ---- density_advance on the device (general path) ----------------------------------------------
static void density_advance_dev(const mgpu_params& P, int which_step, DV& sold, DV& snew, DV* sedge, DV* sflux,
                                DV& scal_force, DV* umac, const double* w0_h, DV& eta, const double* rho0_old_h,
                                const double* rho0_new_h, const double* rho0_pe_h, const int* lo, const int* hi,
                                int ng_s, int ng_f, const int* adv_bc, const int* pmask) {
  const int dm = P.dm, nr = P.nr;
  const int spt = P.species_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;  // variables.f90:119-121
  std::vector<double> e_old(nr + 1), e_new(nr + 1);
  cell_to_edge_host(rho0_old_h, e_old.data(), nr);  // density_advance.f90:90-91
  cell_to_edge_host(rho0_new_h, e_new.data(), nr);
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* rho0_new = upload_small(rho0_new_h, nr);
  const double* rho0_pe = upload_small(rho0_pe_h, nr + 1);
  const double* rho0_edge_old = upload_small(e_old.data(), nr + 1);
  const double* rho0_edge_new = upload_small(e_new.data(), nr + 1);
  int nodal_d[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  auto fill_umac = [&]() {  // addw0.f90:85-93
    for (int d = 0; d < dm; ++d) fill_boundary_dev(P, umac[d], lo, hi, 1, nodal_d[d], 1, 1, 1, adv_bc, pmask, false);
  };

  set_dev(scal_force.p, 0.0, scal_force.size());  // :101-103
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :119-128
    modify_scal_force_dev(P, scal_force, sold, umac, rho0_old, rho0_edge_old, w0, P.rho_comp,
                          spt == MGPU_PREDICT_RHO_AND_X, lo, hi);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
  }
  addw0_dev(P, umac, w0, 1.0, lo, hi);  // :148
  fill_umac();
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :160
    convert_rhoX_to_X_dev(P, sold, true, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, foextrap_comp, P.nspec, adv_bc, pmask, true);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {  // :166
    put_in_pert_form_dev(P, sold, rho0_old, P.rho_comp, true, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
  }
  auto edge = [&](int scomp, int ncomp, bool cons) {
    for (int n = 0; n < ncomp; ++n) {
      if (P.bds_type != 0) {  // density_advance.f90:183-185 etc.
        size_t mark = arena_mark();
        bds_dev(P, sold, sedge, umac, scal_force, lo, hi, scomp - 1 + n, cons, ng_s, ng_f);
        arena_release(mark);
        continue;
      }
      edge_one_comp(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, scomp - 1 + n, dm + scomp + n, false, cons, ng_s,
                    ng_f);
    }
  };
  if (spt == MGPU_PREDICT_RHOX) edge(P.spec_comp, P.nspec, true);  // :190-198
  else edge(P.spec_comp, P.nspec, false);                            // :178-186
  if (spt == MGPU_PREDICT_RHOX) {  // :204-213
    for (int d = 0; d < dm; ++d) sum_comps_dev(sedge[d], P.rho_comp - 1, P.spec_comp - 1, P.nspec);
  } else {
    edge(P.rho_comp, 1, false);  // :216-224
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {  // :229
    put_in_pert_form_dev(P, sold, rho0_old, P.rho_comp, false, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :235
    convert_rhoX_to_X_dev(P, sold, false, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
  }
  if (P.ntrac >= 1) edge(P.trac_comp, P.ntrac, false);  // :242-252
  addw0_dev(P, umac, w0, -1.0, lo, hi);                 // :258
  fill_umac();

  FluxArgs fa;
  fill_flux_args(P, fa, lo, hi);
  for (int d = 0; d < dm; ++d) { fa.sflux[d] = sflux[d]; fa.sedge[d] = sedge[d]; fa.umac[d] = umac[d]; }
  fa.eta = eta;
  fa.w0 = w0;
  fa.rho0_old = rho0_old;
  fa.rho0_edge_old = rho0_edge_old;
  fa.rho0_new = (which_step == 1) ? rho0_old : rho0_new;  // :280-329
  fa.rho0_edge_new = (which_step == 1) ? rho0_edge_old : rho0_edge_new;
  fa.rho0_predicted_edge = rho0_pe;
  set_dev(scal_force.p, 0.0, scal_force.size());  // :349-351
  UpdArgs ua;
  ua.dm = dm;
  ua.dt = P.dt;
  for (int d = 0; d < 3; ++d) ua.dx[d] = P.dx[d];
  ua.vb = grown(lo, hi, dm, 0);
  ua.sold = sold;
  ua.snew = snew;
  ua.force = scal_force;
  for (int d = 0; d < dm; ++d) ua.sflux[d] = sflux[d];
  if (g_opt_fused) {
    // :280-366 in one launch: species + tracer fluxes, etarhoflux, update, density, floors
    flux_update_all_dev(P, fa, ua);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
    if (P.ntrac >= 1)
      fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask, false);
    return;
  }
  mk_rhoX_flux_dev(P, fa, P.spec_comp, P.spec_comp + P.nspec - 1);
  if (P.ntrac >= 1) mk_rhoX_flux_dev(P, fa, P.trac_comp, P.trac_comp + P.ntrac - 1);
  update_scal_dev(P, ua, P.spec_comp, P.spec_comp + P.nspec - 1);  // :360
  fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
  fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
  if (P.ntrac >= 1) {
    update_scal_dev(P, ua, P.trac_comp, P.trac_comp + P.ntrac - 1);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask, false);
  }
}

}  // namespace mgpu

using namespace mgpu;

#define MGPU_TRY try {
#define MGPU_CATCH                    \
  }                                   \
  catch (const std::exception& e) {   \
    g_err = e.what();                 \
    return 1;                         \
  }                                   \
  return 0;

extern "C" {

const char* mgpu_last_error(void) { return g_err.c_str(); }
const char* mgpu_version(void) { return "maestro_b200 0.1 (sm_100a)"; }

int mgpu_init(int device) {
  MGPU_TRY
  if (g_ctx.initialised) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) throw Error("mgpu_init: no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= n) throw Error("mgpu_init: invalid device index");
  MGPU_CUDA(cudaSetDevice(device));
  g_ctx.device = device;
  MGPU_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  g_ctx.own_stream = true;
  g_ctx.initialised = true;
  MGPU_CATCH
}

int mgpu_set_stream(void* stream) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  if (g_ctx.own_stream) cudaStreamDestroy(g_ctx.stream);
  g_ctx.stream = (cudaStream_t)stream;
  g_ctx.own_stream = false;
  MGPU_CATCH
}

int mgpu_finalize(void) {
  MGPU_TRY
  if (!g_ctx.initialised) return 0;
  cudaStreamSynchronize(g_ctx.stream);
  comm_finalize();
  g_pool.clear();
  if (g_ctx.arena) cudaFree(g_ctx.arena);
  g_ctx.arena = nullptr;
  g_ctx.arena_bytes = 0;
  if (g_ctx.own_stream) cudaStreamDestroy(g_ctx.stream);
  g_ctx.stream = nullptr;
  g_ctx.initialised = false;
  MGPU_CATCH
}

int mgpu_set_option(const char* key, int value) {
  MGPU_TRY
  std::string k(key ? key : "");
  if (k == "fused") g_opt_fused = value;
  else if (k == "kchunk") g_opt_kchunk = value;
  else if (k == "exact") g_opt_exact = value;
  else throw Error("mgpu_set_option: unknown key " + k);
  MGPU_CATCH
}

int mgpu_profile(int on) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  g_prof.on = on != 0;
  for (int t = 0; t < TAG_COUNT; ++t) g_prof.used[t] = 0;
  MGPU_CATCH
}
/* accumulated device time (ms) and launch count of kernel class `tag` since mgpu_profile(1) */
int mgpu_profile_get(int tag, double* ms, long* launches) {
  MGPU_TRY
  require_init();
  if (tag < 0 || tag >= TAG_COUNT) throw Error("mgpu_profile_get: bad tag");
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  double tot = 0.0;
  for (size_t i = 0; i < g_prof.used[tag]; ++i) {
    float f = 0.f;
    MGPU_CUDA(cudaEventElapsedTime(&f, g_prof.ev[tag][i].first, g_prof.ev[tag][i].second));
    tot += f;
  }
  *ms = tot;
  *launches = (long)g_prof.used[tag];
  MGPU_CATCH
}

int mgpu_synchronize(void) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  MGPU_CATCH
}

long mgpu_launch_count(int reset) {
  long n = g_ctx.launches;
  if (reset) g_ctx.launches = 0;
  return n;
}
void* mgpu_stream(void) { return (void*)g_ctx.stream; }

int mgpu_malloc(double** dptr, long n) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaMalloc((void**)dptr, (size_t)n * sizeof(double)));
  MGPU_CATCH
}
int mgpu_free(double* dptr) {
  MGPU_TRY
  MGPU_CUDA(cudaFree(dptr));
  MGPU_CATCH
}
int mgpu_memcpy_h2d(double* dst, const double* src, long n) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  MGPU_CATCH
}
int mgpu_memcpy_d2h(double* dst, const double* src, long n) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  MGPU_CATCH
}
int mgpu_host_register(double* hptr, long n) {
  MGPU_TRY
  require_init();
  MGPU_CUDA(cudaHostRegister(hptr, (size_t)n * sizeof(double), cudaHostRegisterDefault));
  MGPU_CATCH
}
int mgpu_host_unregister(double* hptr) {
  MGPU_TRY
  MGPU_CUDA(cudaHostUnregister(hptr));
  MGPU_CATCH
}

int mgpu_halo_plan_make(const mgpu_params* p, const mgpu_fab* f, const int* pmask, int rank, int nranks,
                        mgpu_halo_plan* out) {
  MGPU_TRY
  if (p->dm != 2 && p->dm != 3) throw Error("mgpu_halo_plan_make: dm must be 2 or 3");
  long ext[3];
  for (int d = 0; d < 3; ++d) ext[d] = d < p->dm ? f->hi[d] - f->lo[d] + 1 + 2 * f->ng + f->nodal[d] : 1;
  halo_plan_make(p->dm, p->domlo, p->domhi, f->lo, f->hi, f->ng, f->nodal, ext, pmask, rank, nranks, out);
  MGPU_CATCH
}
int mgpu_comm_unique_id(void* out128) {
  MGPU_TRY
  comm_unique_id(out128);
  MGPU_CATCH
}
int mgpu_comm_init(int rank, int nranks, const void* unique_id128) {
  MGPU_TRY
  comm_init(rank, nranks, unique_id128);
  MGPU_CATCH
}
int mgpu_comm_finalize(void) {
  MGPU_TRY
  comm_finalize();
  MGPU_CATCH
}

int mgpu_fill_boundary(const mgpu_params* p, mgpu_fab* s, int scomp, int bccomp, int ncomp, const int* adv_bc,
                       const int* pmask) {
  MGPU_TRY
  Call c(p, 0);
  DV v = c.view(*s, true, true);
  fill_boundary_dev(*p, v, s->lo, s->hi, s->ng, s->nodal, scomp, bccomp, ncomp, adv_bc, pmask, false);
  c.finish();
  MGPU_CATCH
}

int mgpu_make_edge_scal(const mgpu_params* p, int nfabs, const mgpu_fab* s, mgpu_fab* const* sedge,
                        const mgpu_fab* const* umac, const mgpu_fab* force, const int* adv_bc, int is_vel,
                        int start_scomp, int start_bccomp, int num_comp, int is_conservative) {
  MGPU_TRY
  if (p->bds_type != 0) throw Error("mgpu_make_edge_scal called with bds_type /= 0; call mgpu_bds");
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) {
    size_t n = make_edge_scal_scratch(*p, s[i].lo, s[i].hi);
    if (n > need) need = n;
  }
  Call c(p, need);
  for (int i = 0; i < nfabs; ++i) {
    DV sv = c.view(s[i], true, false), fv = c.view(force[i], true, false);
    DV se[3], um[3];
    // sedge holds other components the caller may already have filled: copy in as well as out
    c.views((const mgpu_fab* const*)sedge, i, true, true, se);
    c.views(umac, i, true, false, um);
    for (int scomp = start_scomp; scomp < start_scomp + num_comp; ++scomp)
      edge_one_comp(*p, sv, se, um, fv, s[i].lo, s[i].hi, adv_bc, scomp - 1, start_bccomp + scomp - start_scomp,
                    is_vel != 0, is_conservative != 0, s[i].ng, force[i].ng);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_bds(const mgpu_params* p, int nfabs, const mgpu_fab* s, mgpu_fab* const* sedge,
             const mgpu_fab* const* umac, const mgpu_fab* force, const int* adv_bc, int is_vel, int start_scomp,
             int start_bccomp, int num_comp, int is_conservative) {
  MGPU_TRY
  (void)adv_bc; (void)is_vel; (void)start_bccomp;  // unused by the reference too (bds.f90:16-131)
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) need = std::max(need, bds_scratch(*p, s[i].lo, s[i].hi));
  Call c(p, need);
  for (int i = 0; i < nfabs; ++i) {
    DV sv = c.view(s[i], true, false), fv = c.view(force[i], true, false);
    DV se[3], um[3];
    c.views((const mgpu_fab* const*)sedge, i, true, true, se);
    c.views(umac, i, true, false, um);
    for (int scomp = start_scomp; scomp < start_scomp + num_comp; ++scomp) {
      size_t mark = arena_mark();
      bds_dev(*p, sv, se, um, fv, s[i].lo, s[i].hi, scomp - 1, is_conservative != 0, s[i].ng, force[i].ng);
      arena_release(mark);
    }
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mk_rhoX_flux(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, mgpu_fab* etarhoflux,
                      const mgpu_fab* const* sedge, const mgpu_fab* const* umac, const double* w0,
                      const double* rho0_old, const double* rho0_edge_old, const double* rho0_new,
                      const double* rho0_edge_new, const double* rho0_predicted_edge, int startcomp, int endcomp) {
  MGPU_TRY
  Call c(p, (size_t)(6 * (p->nr + 2)) * sizeof(double) + 4096);
  const int nr = p->nr;
  for (int i = 0; i < nfabs; ++i) {
    FluxArgs a;
    const int* lo = umac[0][i].lo;
    const int* hi = umac[0][i].hi;
    fill_flux_args(*p, a, lo, hi);
    c.views((const mgpu_fab* const*)sflux, i, true, true, a.sflux);
    c.views(sedge, i, true, false, a.sedge);
    c.views(umac, i, true, false, a.umac);
    a.eta = c.view(etarhoflux[i], true, true);
    a.w0 = upload_small(w0, nr + 1);
    a.rho0_old = upload_small(rho0_old, nr);
    a.rho0_new = upload_small(rho0_new, nr);
    a.rho0_edge_old = upload_small(rho0_edge_old, nr + 1);
    a.rho0_edge_new = upload_small(rho0_edge_new, nr + 1);
    a.rho0_predicted_edge = upload_small(rho0_predicted_edge, nr + 1);
    mk_rhoX_flux_dev(*p, a, startcomp, endcomp);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mk_rhoh_flux(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                      const mgpu_fab* const* umac, const double* w0, const double* rho0_old,
                      const double* rho0_edge_old, const double* rho0_new, const double* rho0_edge_new,
                      const double* rhoh0_old, const double* rhoh0_edge_old, const double* rhoh0_new,
                      const double* rhoh0_edge_new) {
  MGPU_TRY
  Call c(p, (size_t)(9 * (p->nr + 2)) * sizeof(double) + 8192);
  const int nr = p->nr;
  for (int i = 0; i < nfabs; ++i) {
    FluxArgs a;
    fill_flux_args(*p, a, umac[0][i].lo, umac[0][i].hi);
    c.views((const mgpu_fab* const*)sflux, i, true, true, a.sflux);
    c.views(sedge, i, true, false, a.sedge);
    c.views(umac, i, true, false, a.umac);
    a.w0 = upload_small(w0, nr + 1);
    a.rho0_old = upload_small(rho0_old, nr);
    a.rho0_new = upload_small(rho0_new, nr);
    a.rho0_edge_old = upload_small(rho0_edge_old, nr + 1);
    a.rho0_edge_new = upload_small(rho0_edge_new, nr + 1);
    a.rhoh0_old = upload_small(rhoh0_old, nr);
    a.rhoh0_new = upload_small(rhoh0_new, nr);
    a.rhoh0_edge_old = upload_small(rhoh0_edge_old, nr + 1);
    a.rhoh0_edge_new = upload_small(rhoh0_edge_new, nr + 1);
    mk_rhoh_flux_dev(*p, a);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_update_scal(const mgpu_params* p, int nfabs, int nstart, int nstop, const mgpu_fab* sold, mgpu_fab* snew,
                     const mgpu_fab* const* sflux, const mgpu_fab* force) {
  MGPU_TRY
  Call c(p, 0);
  for (int i = 0; i < nfabs; ++i) {
    UpdArgs a;
    a.dm = p->dm;
    a.dt = p->dt;
    for (int d = 0; d < 3; ++d) a.dx[d] = p->dx[d];
    a.vb = grown(sold[i].lo, sold[i].hi, p->dm, 0);
    a.sold = c.view(sold[i], true, false);
    a.snew = c.view(snew[i], true, true);
    a.force = c.view(force[i], true, false);
    c.views(sflux, i, true, false, a.sflux);
    update_scal_dev(*p, a, nstart, nstop);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_update_velocity(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                         const mgpu_fab* const* umac, const mgpu_fab* const* uedge, const mgpu_fab* force,
                         const mgpu_fab* sponge, const double* w0) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_update_velocity: spherical geometry not available on the device yet");
  Call c(p, (size_t)(p->nr + 2) * sizeof(double) + 4096);
  for (int i = 0; i < nfabs; ++i) {
    VelArgs a;
    a.dm = p->dm;
    a.do_sponge = p->do_sponge != 0;
    a.dt = p->dt;
    for (int d = 0; d < 3; ++d) a.dx[d] = p->dx[d];
    a.vb = grown(uold[i].lo, uold[i].hi, p->dm, 0);
    a.uold = c.view(uold[i], true, false);
    a.unew = c.view(unew[i], true, true);
    a.force = c.view(force[i], true, false);
    a.sponge = c.view(sponge[i], true, false);
    c.views(umac, i, true, false, a.umac);
    c.views(uedge, i, true, false, a.uedge);
    a.w0 = upload_small(w0, p->nr + 1);
    update_velocity_dev(a);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_addw0(const mgpu_params* p, int nfabs, mgpu_fab* const* umac, const double* w0, double mult) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_addw0: spherical geometry not available on the device yet");
  Call c(p, (size_t)(p->nr + 2) * sizeof(double) + 4096);
  const double* w0d = upload_small(w0, p->nr + 1);
  for (int i = 0; i < nfabs; ++i) {
    DV um[3];
    c.views((const mgpu_fab* const*)umac, i, true, true, um);
    addw0_dev(*p, um, w0d, mult, umac[0][i].lo, umac[0][i].hi);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mkutrans(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                  mgpu_fab* const* utrans, const double* w0, const int* adv_bc, const int* phys_bc) {
  MGPU_TRY
  Call c(p, (size_t)(p->nr + 2) * sizeof(double) + 4096);
  const double* w0d = upload_small(w0, p->nr + 1);
  for (int i = 0; i < nfabs; ++i) {
    DV ut = c.view(utilde[i], true, false), uf = c.view(ufull[i], true, false);
    DV tr[3];
    c.views((const mgpu_fab* const*)utrans, i, true, true, tr);
    mkutrans_dev(*p, ut, uf, tr, w0d, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_velpred(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                 mgpu_fab* const* umac, const mgpu_fab* const* utrans, const mgpu_fab* force, const double* w0,
                 const int* adv_bc, const int* phys_bc) {
  MGPU_TRY
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) need = std::max(need, velpred_scratch(*p, utilde[i].lo, utilde[i].hi));
  Call c(p, need + (size_t)(p->nr + 2) * sizeof(double) + 4096);
  const double* w0d = upload_small(w0, p->nr + 1);
  for (int i = 0; i < nfabs; ++i) {
    size_t mark = arena_mark();
    DV ut = c.view(utilde[i], true, false), uf = c.view(ufull[i], true, false), fv = c.view(force[i], true, false);
    DV um[3], tr[3];
    c.views((const mgpu_fab* const*)umac, i, true, true, um);
    c.views(utrans, i, true, false, tr);
    velpred_dev(*p, ut, uf, um, tr, fv, w0d, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng, force[i].ng);
    arena_release(mark);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_modify_scal_force(const mgpu_params* p, int nfabs, mgpu_fab* force, const mgpu_fab* s,
                           const mgpu_fab* const* umac, const double* s0, const double* s0_edge, const double* w0,
                           int comp, int fullform) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_modify_scal_force: spherical geometry not available on the device yet");
  Call c(p, (size_t)(3 * (p->nr + 2)) * sizeof(double) + 4096);
  const double* s0d = upload_small(s0, p->nr);
  const double* s0ed = upload_small(s0_edge, p->nr + 1);
  const double* w0d = upload_small(w0, p->nr + 1);
  for (int i = 0; i < nfabs; ++i) {
    DV fv = c.view(force[i], true, true), sv = c.view(s[i], true, false);
    DV um[3];
    c.views(umac, i, true, false, um);
    modify_scal_force_dev(*p, fv, sv, um, s0d, s0ed, w0d, comp, fullform != 0, s[i].lo, s[i].hi);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_convert_rhoX_to_X(const mgpu_params* p, int nfabs, mgpu_fab* s, int flag) {
  MGPU_TRY
  Call c(p, 0);
  for (int i = 0; i < nfabs; ++i) {
    DV sv = c.view(s[i], true, true);
    convert_rhoX_to_X_dev(*p, sv, flag != 0, s[i].lo, s[i].hi);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_put_in_pert_form(const mgpu_params* p, int nfabs, mgpu_fab* s, const double* base, int comp, int flag) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_put_in_pert_form: spherical geometry not available on the device yet");
  Call c(p, (size_t)(p->nr + 2) * sizeof(double) + 4096);
  const double* bd = upload_small(base, p->nr);
  for (int i = 0; i < nfabs; ++i) {
    DV sv = c.view(s[i], true, true);
    put_in_pert_form_dev(*p, sv, bd, comp, flag != 0, s[i].lo, s[i].hi);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_density_advance(const mgpu_params* p, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                         mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                         mgpu_fab* const* umac, const double* w0, mgpu_fab* etarhoflux, const double* rho0_old,
                         const double* rho0_new, const double* p0_dummy, const double* rho0_predicted_edge,
                         const int* adv_bc, const int* pmask) {
  MGPU_TRY
  (void)p0_dummy;
  if (p->spherical) throw Error("mgpu_density_advance: spherical geometry not available on the device yet");
  Call c(p, std::max(make_edge_scal_scratch(*p, sold->lo, sold->hi), bds_scratch(*p, sold->lo, sold->hi)) +
             (size_t)(8 * (p->nr + 2)) * sizeof(double) + 8192);
  DV so = c.view(*sold, true, true), sn = c.view(*snew, true, true), fv = c.view(*scal_force, true, true);
  DV eta = c.view(*etarhoflux, true, true);
  DV se[3], sf[3], um[3];
  c.views((const mgpu_fab* const*)sedge, 0, true, true, se);
  c.views((const mgpu_fab* const*)sflux, 0, true, true, sf);
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  density_advance_dev(*p, which_step, so, sn, se, sf, fv, um, w0, eta, rho0_old, rho0_new, rho0_predicted_edge,
                      sold->lo, sold->hi, sold->ng, scal_force->ng, adv_bc, pmask);
  c.finish();
  MGPU_CATCH
}

}  // extern "C"
