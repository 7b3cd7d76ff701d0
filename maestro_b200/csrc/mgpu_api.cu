// C ABI of the library (include/maestro_b200.h): context, host staging, and the L3/L4 entry points.
//
// Host-pointer calls (mem_space == MGPU_HOST) are what the Fortran shim issues: inputs are copied to
// the device, the kernels run, outputs are copied back and the call returns synchronised, so the
// unchanged Fortran sees exactly what the reference routine would have left in its multifabs.
// Device-pointer calls (MGPU_DEVICE) skip the copies and stay asynchronous on the library stream.
// There is no CPU fallback anywhere: without a CUDA device every entry point fails.
#include <algorithm>
#include <cstring>
#include <limits>
#include <map>

#include "mgpu_bds.cuh"
#include "mgpu_edge.cuh"
#include "mgpu_eos.cuh"
#include "mgpu_fused.cuh"
#include "mgpu_halo.cuh"
#include "mgpu_reduce.cuh"
#include "mgpu_sphr.cuh"
#include "mgpu_stream.cuh"
#include "mgpu_velpred.cuh"

namespace mgpu {

static Context g_ctx;
static std::string g_err;
static int g_opt_fused = 1;    // use the fused 3-D edge kernel when it covers the case
static int g_opt_kchunk = -1;  // z planes per CTA of the fused kernels; < 0: chosen per launch (fused2_auto_kchunk)
static int g_opt_overlap = 1;   // slab runs: exchange the updated boundary planes while the interior is updated
static int g_opt_leanplus = 1;  // density_advance: on-the-fly input transforms where the upwind-first kernel applies
static int g_opt_premac_fuse = 1;  // advance_premac: utrans formed inside the face kernel of velpred where no physical boundary is near
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
// Copies are per component (bit c of the masks = component c): an L4 episode moves only the components it reads
// across PCIe on the way in and only those it writes on the way out.
static long g_h2d_bytes = 0, g_d2h_bytes = 0;  // PCIe traffic of host-pointer calls (bench e2e accounting)
typedef unsigned long long cmask_t;
static const cmask_t ALLC = ~0ull;
static inline cmask_t crange(int c0, int nc) {  // components c0 .. c0+nc-1 (0-based)
  cmask_t m = 0;
  for (int c = c0; c < c0 + nc; ++c) m |= (1ull << c);
  return m;
}
// ---- asynchronous uploads ---------------------------------------------------------------------------------------------
// A host-pointer call normally issues its host-to-device copies on the compute stream, so the first kernel starts after
// the last byte has arrived.  An episode that knows the order in which it consumes its inputs (density_advance in its
// lean+ form) switches uploads to a copy stream instead: every component is one copy + one event, and the episode makes
// the compute stream wait for a component right before the first kernel that reads it (uploads_wait_range).  The PCIe
// transfer of component n+1 then runs under the edge kernel of component n.
struct PendingUpload {
  const double *b, *e;  // device range
  cudaEvent_t ev;
};
static bool g_async_uploads = false;       // set by the entry point for the duration of its view() calls
static int g_opt_async_upload = 1;         // option "async_upload"
static cudaStream_t g_copy_stream = nullptr;
static std::vector<PendingUpload> g_pending;
static std::vector<cudaEvent_t> g_ev_pool;
static cudaEvent_t upload_event() {
  if (g_ev_pool.empty()) {
    cudaEvent_t e;
    MGPU_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return e;
  }
  cudaEvent_t e = g_ev_pool.back();
  g_ev_pool.pop_back();
  return e;
}
// the compute stream waits for every pending upload that overlaps [b, e)
static void uploads_wait_range(const double* b, const double* e) {
  for (size_t i = 0; i < g_pending.size();) {
    if (g_pending[i].b < e && b < g_pending[i].e) {
      MGPU_CUDA(cudaStreamWaitEvent(g_ctx.stream, g_pending[i].ev, 0));
      g_ev_pool.push_back(g_pending[i].ev);
      g_pending[i] = g_pending.back();
      g_pending.pop_back();
    } else {
      ++i;
    }
  }
}
static void uploads_wait_all() {
  for (auto& p : g_pending) {
    MGPU_CUDA(cudaStreamWaitEvent(g_ctx.stream, p.ev, 0));
    g_ev_pool.push_back(p.ev);
  }
  g_pending.clear();
}
static void uploads_wait_comp(const DV& f, int c) { uploads_wait_range(f.p + f.cs * c, f.p + f.cs * (c + 1)); }

// ---- residency registry (lifecycle) ---------------------------------------------------------
// A host fab registered with mgpu_register keeps ONE device mirror for as long as it is registered.  Per component
// the registry knows which side holds the truth:
//   host_dirty  the host copy is newer: the next call that reads the component uploads it first
//   dev_dirty   the device copy is newer (a call wrote it): the host sees it after mgpu_download
// Calls with host pointers then move only what is stale, and outputs stay on the device until someone asks for them:
// the episodes of one MAESTRO step hand sedge / sflux / scal_force / umac to each other without crossing PCIe.
struct Resident {
  double* d = nullptr;
  size_t n = 0;  // doubles
  cmask_t host_dirty = ALLC, dev_dirty = 0;
  size_t cs = 0;  // component stride, known from the first call that views the fab
  int nc = 0;
  bool pinned = false;
};
static std::map<const double*, Resident> g_res;
static const int NODAL_OF[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};

struct Call {
  const mgpu_params& P;
  bool host;
  struct Item { double* h; double* d; size_t cs; int nc; cmask_t out; cmask_t zero; cmask_t in_done; Resident* res; };
  std::vector<Item> items;
  explicit Call(const mgpu_params* p, size_t scratch_bytes) : P(*p), host(p->mem_space == MGPU_HOST) {
    require_init();
    if (P.dm != 2 && P.dm != 3) throw Error("mgpu: dm must be 2 or 3 (1-D stays with the Fortran)");
    arena_reserve(scratch_bytes + (1u << 20));
    arena_reset();
  }
  ~Call() {
    if (!g_pending.empty()) {  // an exception cut the call short: let the copies finish before their buffers are reused
      cudaStreamSynchronize(g_copy_stream);
      for (auto& p : g_pending) g_ev_pool.push_back(p.ev);
      g_pending.clear();
    }
    g_async_uploads = false;
    for (auto& it : items)
      if (!it.res) g_pool.put(it.d);
  }
  // runs of set bits of m (restricted to nc components) -> one cudaMemcpyAsync each
  template <class F>
  static void for_runs(cmask_t m, int nc, F f) {
    int c = 0;
    while (c < nc) {
      if (!((m >> c) & 1ull)) { ++c; continue; }
      int c1 = c;
      while (c1 < nc && ((m >> c1) & 1ull)) ++c1;
      f(c, c1 - c);
      c = c1;
    }
  }
  DV view(const mgpu_fab& f, cmask_t in, cmask_t out) {
    if (!f.ptr) throw Error("mgpu: null fab pointer");
    if (!host) return make_view(f, P.dm);
    if (f.nc > 64) in = out = (in || out) ? ALLC : 0;
    DV v = make_view(f, P.dm);
    auto upload = [&](double* d, cmask_t comps) {
      for_runs(comps, f.nc, [&](int c0, int n) {
        g_h2d_bytes += (long)v.cs * n * (long)sizeof(double);
        if (g_async_uploads) {  // one copy + one event per component, on the copy stream (see uploads_wait_range)
          for (int c = c0; c < c0 + n; ++c) {
            MGPU_CUDA(cudaMemcpyAsync(d + v.cs * c, f.ptr + v.cs * c, (size_t)v.cs * sizeof(double), cudaMemcpyHostToDevice,
                                      g_copy_stream));
            cudaEvent_t ev = upload_event();
            MGPU_CUDA(cudaEventRecord(ev, g_copy_stream));
            g_pending.push_back({d + v.cs * c, d + v.cs * (c + 1), ev});
          }
          return;
        }
        MGPU_CUDA(cudaMemcpyAsync(d + v.cs * c0, f.ptr + v.cs * c0, (size_t)v.cs * n * sizeof(double),
                                  cudaMemcpyHostToDevice, g_ctx.stream));
      });
    };
    const cmask_t all = f.nc >= 64 ? ALLC : ((cmask_t(1) << f.nc) - 1);
    for (auto& it : items)  // the same host fab passed twice maps to one device buffer
      if (it.h == f.ptr) {
        cmask_t more = in & all & ~it.in_done;  // components this view reads that the first one did not bring in
        if (it.res) {
          more &= it.res->host_dirty;
          it.res->host_dirty &= ~more;
        }
        upload(it.d, more);
        it.in_done |= in & all;
        it.out |= out;
        return make_view(f, P.dm, it.d);
      }
    auto rit = g_res.find(f.ptr);
    if (rit != g_res.end()) {  // registered: the persistent mirror, only the stale components travel
      Resident& r = rit->second;
      if ((size_t)v.size() > r.n) throw Error("mgpu: the fab is larger than the region registered for it");
      if (r.cs != 0 && (r.cs != (size_t)v.cs || r.nc != f.nc)) throw Error("mgpu: a registered fab changed its shape (unregister it first)");
      r.cs = (size_t)v.cs;
      r.nc = f.nc;
      const cmask_t need = in & all & r.host_dirty;
      upload(r.d, need);
      r.host_dirty &= ~need;
      items.push_back({f.ptr, r.d, (size_t)v.cs, f.nc, out, (cmask_t)0, in & all, &r});
      return make_view(f, P.dm, r.d);
    }
    double* d = g_pool.get((size_t)v.size());
    upload(d, in & all);
    items.push_back({f.ptr, d, (size_t)v.cs, f.nc, out, (cmask_t)0, in & all, nullptr});
    return make_view(f, P.dm, d);
  }
  DV view(const mgpu_fab& f, bool copy_in, bool copy_out) { return view(f, copy_in ? ALLC : 0, copy_out ? ALLC : 0); }
  // the caller-visible result of this fab is all zeros (setval(...,ZERO,all=.true.) was the last thing the
  // reference did to it): produce it on the host instead of copying zeros back over PCIe
  void zero_on_host(const mgpu_fab& f, cmask_t comps = ALLC) {
    for (auto& it : items)
      if (it.h == f.ptr) { it.out &= ~comps; it.zero |= comps; }
  }
  void views(const mgpu_fab* const* f, int i, bool in, bool out, DV* v) {
    for (int d = 0; d < P.dm; ++d) v[d] = view(f[d][i], in, out);
  }
  void views(const mgpu_fab* const* f, int i, cmask_t in, cmask_t out, DV* v) {
    for (int d = 0; d < P.dm; ++d) v[d] = view(f[d][i], in, out);
  }
  void finish() {
    uploads_wait_all();  // nothing may still be in flight when the results travel back (or the call returns)
    g_async_uploads = false;
    if (!host) return;
    for (auto& it : items) {
      if (it.res) {  // registered: the results stay on the device until mgpu_download
        const cmask_t all = it.nc >= 64 ? ALLC : ((cmask_t(1) << it.nc) - 1);
        it.res->dev_dirty |= it.out & all;
        it.res->host_dirty &= ~(it.out & all);
        continue;
      }
      for_runs(it.out, it.nc, [&](int c0, int n) {
        MGPU_CUDA(cudaMemcpyAsync(it.h + it.cs * c0, it.d + it.cs * c0, it.cs * n * sizeof(double),
                                  cudaMemcpyDeviceToHost, g_ctx.stream));
        g_d2h_bytes += (long)(it.cs * n * sizeof(double));
      });
    }
    for (auto& it : items)  // overlaps the asynchronous device-to-host copies
      for_runs(it.zero, it.nc, [&](int c0, int n) {
        if (it.res) {  // registered: the zeros are produced in the mirror and reach the host with mgpu_download
          MGPU_CUDA(cudaMemsetAsync(it.d + it.cs * c0, 0, it.cs * n * sizeof(double), g_ctx.stream));
          const cmask_t m = ((n >= 64 ? ALLC : ((cmask_t(1) << n) - 1)) << c0);
          it.res->dev_dirty |= m;
          it.res->host_dirty &= ~m;
          return;
        }
        memset(it.h + it.cs * c0, 0, it.cs * n * sizeof(double));
      });
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
                          bool is_cons, int ng_s, int ng_f, bool force_zero = false, const double* sdiv = nullptr,
                          const double* ssub = nullptr, const double* wadd = nullptr) {
  for (int d = 0; d < P.dm; ++d) {  // validate BC codes up front (make_edge_scal.f90:853)
    for (int side = 0; side < 2; ++side) {
      const int bc = adv_bc[d + P.dm * (side + 2 * (bccomp - 1))];
      if (bc != MGPU_BC_EXT_DIR && bc != MGPU_BC_FOEXTRAP && bc != MGPU_BC_HOEXTRAP && bc != MGPU_BC_REFLECT_EVEN &&
          bc != MGPU_BC_REFLECT_ODD && bc != MGPU_BC_INTERIOR)
        throw Error("make_edge_scal: invalid boundary type adv_bc");
    }
  }
  if (g_opt_fused && !sdiv && !ssub && !wadd && fused_edge2d_supported(P, is_cons, adv_bc, bccomp, g_opt_exact != 0)) {
    fused_edge2d_dev(P, s, sedge, umac, force, lo, hi, adv_bc, comp, bccomp, is_vel, ng_s, ng_f, force_zero);
  } else if (g_opt_fused && fused_edge_supported(P, is_cons)) {
    fused_edge_dev(P, s, sedge, umac, force, lo, hi, adv_bc, comp, bccomp, is_vel, ng_s, ng_f, g_opt_kchunk,
                   g_opt_exact != 0, force_zero, sdiv, ssub, wadd);
  } else {
    size_t mark = arena_mark();
    make_edge_scal_dev(P, s, sedge, umac, force, lo, hi, adv_bc, comp, bccomp, is_vel, is_cons, ng_s, ng_f);
    arena_release(mark);
  }
}

// Which device-resident form of density_advance a call takes ("The density_advance episode").  Only global
// properties enter (options, parameters, the BC table, pmask): every rank of a slab run takes the same form.
static bool density_advance_is_lean(const mgpu_params& P) {
  const int spt = P.species_pred_type;
  return g_opt_exact == 0 && g_opt_fused && P.bds_type == 0 && fused_edge_supported(P, false) &&
         (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X);
}
static bool density_advance_is_leanplus(const mgpu_params& P, const int* adv_bc, const int* pmask) {
  if (!density_advance_is_lean(P) || P.dm != 3 || g_opt_leanplus == 0) return false;
  if (!(pmask[0] && pmask[1] && pmask[2])) return false;
  const int dm = P.dm;
  for (int c : {P.rho_comp, P.trac_comp})
    if (!fused_edge_is_upwind_first(P, adv_bc, dm + c, false)) return false;
  for (int n = 0; n < P.nspec; ++n)
    if (!fused_edge_is_upwind_first(P, adv_bc, dm + P.spec_comp + n, false)) return false;
  return true;
}

// ---- density_advance, "lean+" device-resident episode (see density_advance_dev) --------------------------------
static void fill_flux_args(const mgpu_params& P, FluxArgs& a, const int* lo, const int* hi);
static void density_advance_leanplus(const mgpu_params& P, int which_step, DV& sold, DV& snew, DV* sedge, DV* sflux,
                                     DV& scal_force, DV* umac, const double* w0_h, const double* w0,
                                     DV& eta, const double* rho0_old_h, const double* rho0_old,
                                     const double* rho0_edge_old, const double* rho0_new_eff,
                                     const double* rho0_edge_new_eff, const double* rho0_pe, const int* lo,
                                     const int* hi, int ng_s, int ng_f, const int* adv_bc, const int* pmask) {
  (void)which_step;
  const int dm = 3, spt = P.species_pred_type, nr = P.nr;
  const int foextrap_comp = dm + P.nscal + 2;
  const int r = 2;
  // base-state values by fab plane: the ghost planes take the value of the plane they are a copy of (periodic wrap
  // of the global domain; planes of a neighbouring slab keep their own global index)
  auto wrap_cell = [&](int k) {
    const int n = P.domhi[r] - P.domlo[r] + 1;
    while (k < P.domlo[r]) k += n;
    while (k > P.domhi[r]) k -= n;
    return k;
  };
  std::vector<double> sub(sold.n[r]), wadd(umac[r].n[r]);
  for (int q = 0; q < sold.n[r]; ++q) sub[q] = rho0_old_h[wrap_cell(sold.lo[r] + q) - 0];
  for (int q = 0; q < umac[r].n[r]; ++q) {
    int k = umac[r].lo[r] + q;  // z-face index; faces domlo..domhi+1 are the domain's own, ghost faces wrap
    const int n = P.domhi[r] - P.domlo[r] + 1;
    if (k < P.domlo[r]) k += n;
    else if (k > P.domhi[r] + 1) k -= n;
    wadd[q] = w0_h[k];
  }
  (void)nr;
  const double* sub_d = upload_small(sub.data(), sub.size());
  const double* wadd_d = upload_small(wadd.data(), wadd.size());
  int nodal_d[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};

  // uploads still in flight (host-pointer call on one rank, see uploads_wait_range): the species are then waited for
  // and ghost-filled one by one, right before the edge kernel that reads them
  const bool piped = !g_pending.empty();
  if (piped) {
    for (int d = 0; d < dm; ++d) uploads_wait_comp(umac[d], 0);
    uploads_wait_comp(sold, P.rho_comp - 1);
  }
  {  // one batch: ghost cells of umac, of the raw rho / rhoX inputs and of the density force.  The inputs travel
     // (communication stream) while the force is built from the valid cells (modify_scal_force, :119-128)
    FillBatch fb;
    for (int d = 0; d < dm; ++d) fill_boundary_dev(P, umac[d], lo, hi, 1, nodal_d[d], 1, 1, 1, adv_bc, pmask, false);
    if (!piped)
      fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
    if (g_opt_overlap) fb.start_exchange();
    set_dev(scal_force.p + scal_force.cs * (P.rho_comp - 1), 0.0, scal_force.cs);
    modify_scal_force_dev(P, scal_force, sold, umac, rho0_old, rho0_edge_old, w0, P.rho_comp,
                          spt == MGPU_PREDICT_RHO_AND_X, lo, hi, true, true);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
    fb.start_exchange();
    fb.finish();
  }
  // 1/rho once (whole fab, ghost cells included) so that the edge kernels form X = rhoX * (1/rho) with a multiply
  double* rinv = arena_alloc((size_t)sold.cs);
  recip_dev(rinv, sold.p + sold.cs * (P.rho_comp - 1), sold.cs);
  const double* rho_p = rinv;
  for (int n = 0; n < P.nspec; ++n) {  // X = rhoX / rho, zero force (:178-186)
    if (piped) {
      uploads_wait_comp(sold, P.spec_comp - 1 + n);
      FillBatch fb;
      fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp + n, dm + P.spec_comp + n, 1, adv_bc, pmask, false);
      fb.run();
    }
    edge_one_comp(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, P.spec_comp - 1 + n, dm + P.spec_comp + n, false,
                  false, ng_s, ng_f, true, rho_p, nullptr, wadd_d);
  }
  // rho' = rho - rho0 (or rho itself for predict_rho_and_X), with its force (:216-224)
  edge_one_comp(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, P.rho_comp - 1, dm + P.rho_comp, false, false, ng_s,
                ng_f, false, nullptr, spt == MGPU_PREDICT_RHOPRIME_AND_X ? sub_d : nullptr, wadd_d);
  for (int n = 0; n < P.ntrac; ++n) {  // tracers (:242-252)
    if (piped) uploads_wait_comp(sold, P.trac_comp - 1 + n);
    edge_one_comp(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, P.trac_comp - 1 + n, dm + P.trac_comp + n, false,
                  false, ng_s, ng_f, true, nullptr, nullptr, wadd_d);
  }
  uploads_wait_all();

  FluxArgs fa;
  fill_flux_args(P, fa, lo, hi);
  for (int d = 0; d < dm; ++d) { fa.sflux[d] = sflux[d]; fa.sedge[d] = sedge[d]; fa.umac[d] = umac[d]; }
  fa.eta = eta;
  fa.w0 = w0;
  fa.rho0_old = rho0_old;
  fa.rho0_edge_old = rho0_edge_old;
  fa.rho0_new = rho0_new_eff;
  fa.rho0_edge_new = rho0_edge_new_eff;
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
  // scal_force is zero by construction here and the periodic / slab fills below rewrite every ghost cell of snew
  auto fill_snew = [&]() {
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
    if (P.ntrac >= 1)
      fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask, false);
  };
  const int nz = hi[r] - lo[r] + 1;
  if (comm_size() > 1 && g_opt_overlap && nz >= 4 * ng_s && flux_update_range_supported(P, fa, ua)) {
    // slab run: update the ng_s planes next to each slab face first, send them to the neighbours on the
    // communication stream, and update the interior planes while they travel
    flux_update_range_dev(P, fa, ua, true, lo[r], lo[r] + ng_s - 1);
    flux_update_range_dev(P, fa, ua, true, hi[r] - ng_s + 1, hi[r]);
    FillBatch fb;
    fill_snew();
    fb.start_exchange();
    flux_update_range_dev(P, fa, ua, true, lo[r] + ng_s, hi[r] - ng_s);
    fb.finish();
    return;
  }
  flux_update_all_dev(P, fa, ua, false, true, true);
  FillBatch fb;
  fill_snew();
  fb.run();
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
  auto fill_umac = [&]() {  // addw0.f90:85-93; the dm exchanges travel as one NCCL group
    FillBatch fb;
    for (int d = 0; d < dm; ++d) fill_boundary_dev(P, umac[d], lo, hi, 1, nodal_d[d], 1, 1, 1, adv_bc, pmask, false);
    fb.run();
  };

  // "lean" episode (FAST arithmetic, fused edge kernel for every component): the forces of the species and tracers
  // are identically zero (:99-103), so the edge kernels do not read them and only the density component of
  // scal_force has to be zero before modify_scal_force writes it; scal_force is zeroed once, at the end (:349-351).
  const bool lean = density_advance_is_lean(P);
  const bool leanp = density_advance_is_leanplus(P, adv_bc, pmask);
  if (leanp) {
    density_advance_leanplus(P, which_step, sold, snew, sedge, sflux, scal_force, umac, w0_h, w0, eta, rho0_old_h,
                             rho0_old, rho0_edge_old, (which_step == 1) ? rho0_old : rho0_new,
                             (which_step == 1) ? rho0_edge_old : rho0_edge_new, rho0_pe, lo, hi, ng_s, ng_f, adv_bc, pmask);
    return;
  }
  uploads_wait_all();  // (only the lean+ episode consumes its inputs component by component)
  if (lean) set_dev(scal_force.p + scal_force.cs * (P.rho_comp - 1), 0.0, scal_force.cs);
  else set_dev(scal_force.p, 0.0, scal_force.size());  // :101-103
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {  // :119-128
    modify_scal_force_dev(P, scal_force, sold, umac, rho0_old, rho0_edge_old, w0, P.rho_comp,
                          spt == MGPU_PREDICT_RHO_AND_X, lo, hi, g_opt_exact == 0, lean);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
  }
  addw0_dev(P, umac, w0, 1.0, lo, hi);  // :148
  fill_umac();
  // :160-171.  rhoX -> X and rho -> rho' act on the valid cells of different components, so both run before the two
  // ghost fills, which then share one NCCL group (same values as the reference's convert, fill, perturb, fill)
  species_form_dev(P, sold, rho0_old, spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X,
                   spt == MGPU_PREDICT_RHOPRIME_AND_X, true, lo, hi);
  {
    FillBatch fb;
    if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X)
      fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, foextrap_comp, P.nspec, adv_bc, pmask, true);
    if (spt == MGPU_PREDICT_RHOPRIME_AND_X)
      fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
    fb.run();
  }
  auto edge = [&](int scomp, int ncomp, bool cons, bool force_zero = false) {
    for (int n = 0; n < ncomp; ++n) {
      if (P.bds_type != 0) {  // density_advance.f90:183-185 etc.
        size_t mark = arena_mark();
        bds_dev(P, sold, sedge, umac, scal_force, lo, hi, scomp - 1 + n, cons, ng_s, ng_f);
        arena_release(mark);
        continue;
      }
      edge_one_comp(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, scomp - 1 + n, dm + scomp + n, false, cons, ng_s,
                    ng_f, force_zero);
    }
  };
  if (spt == MGPU_PREDICT_RHOX) edge(P.spec_comp, P.nspec, true);  // :190-198
  else edge(P.spec_comp, P.nspec, false, lean);                      // :178-186
  if (spt == MGPU_PREDICT_RHOX) {  // :204-213
    for (int d = 0; d < dm; ++d) sum_comps_dev(sedge[d], P.rho_comp - 1, P.spec_comp - 1, P.nspec);
  } else {
    edge(P.rho_comp, 1, false);  // :216-224
  }
  // :229-240: rho' -> rho, then X -> rhoX with the restored rho (valid cells), then both ghost fills as one group
  species_form_dev(P, sold, rho0_old, spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X,
                   spt == MGPU_PREDICT_RHOPRIME_AND_X, false, lo, hi);
  {
    FillBatch fb;
    if (spt == MGPU_PREDICT_RHOPRIME_AND_X)
      fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
    if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X)
      fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
    fb.run();
  }
  if (P.ntrac >= 1) edge(P.trac_comp, P.ntrac, false, lean);  // :242-252
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
    flux_update_all_dev(P, fa, ua, g_opt_exact != 0);
    FillBatch fb;
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
    if (P.ntrac >= 1)
      fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask, false);
    fb.run();
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


// ---- density_advance over several boxes of one rank (general path, statement by statement) ------------------------
// What density_advance.f90:20 does with a multifab of nfabs boxes: every stage runs box by box, every ghost fill is
// the multifab fill (box-to-box copies with the periodic images, then the physical BCs of each box).  The reference's
// unit test lays 64^3 out as 8 boxes of 32^3 (Exec/UNIT_TESTS/test_advect/gr0_3d).
struct BoxSet {
  int n;
  std::vector<const int*> lo, hi, bc;  // per box: valid range and its own adv_bc table
};
static void fill_mf(const mgpu_params& P, const BoxSet& B, std::vector<DV>& f, int ng, const int* nodal, int scomp, int bccomp,
                    int ncomp, const int* pmask, bool same_boundary = false) {
  fill_boundary_mf_dev(P, B.n, f.data(), B.lo.data(), B.hi.data(), ng, nodal, scomp, bccomp, ncomp, B.bc.data(), pmask,
                       same_boundary);
}
static void density_advance_mf_dev(const mgpu_params& P, int which_step, const BoxSet& B, std::vector<DV>& sold,
                                   std::vector<DV>& snew, std::vector<DV>* sedge, std::vector<DV>* sflux,
                                   std::vector<DV>& scal_force, std::vector<DV>* umac, const double* w0_h,
                                   std::vector<DV>& eta, const double* rho0_old_h, const double* rho0_new_h,
                                   const double* rho0_pe_h, int ng_s, int ng_f, const int* pmask) {
  const int dm = P.dm, nr = P.nr, nf = B.n;
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
  const bool rx = spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X;
  auto box_umac = [&](int i, DV* u) { for (int d = 0; d < dm; ++d) u[d] = umac[d][i]; };
  auto box_faces = [&](std::vector<DV>* f, int i, DV* out) { for (int d = 0; d < dm; ++d) out[d] = f[d][i]; };
  auto fill_umac = [&]() {  // addw0.f90:85-93
    for (int d = 0; d < dm; ++d) fill_mf(P, B, umac[d], 1, NODAL_OF[d], 1, 1, 1, pmask);
  };
  for (int i = 0; i < nf; ++i) set_dev(scal_force[i].p, 0.0, scal_force[i].size());  // :101-103
  if (rx) {  // :119-128
    for (int i = 0; i < nf; ++i) {
      DV u[3];
      box_umac(i, u);
      modify_scal_force_dev(P, scal_force[i], sold[i], u, rho0_old, rho0_edge_old, w0, P.rho_comp,
                            spt == MGPU_PREDICT_RHO_AND_X, B.lo[i], B.hi[i], g_opt_exact == 0, false);
    }
    fill_mf(P, B, scal_force, ng_f, nullptr, P.rho_comp, foextrap_comp, 1, pmask);
  }
  for (int i = 0; i < nf; ++i) {  // :148
    DV u[3];
    box_umac(i, u);
    addw0_dev(P, u, w0, 1.0, B.lo[i], B.hi[i]);
  }
  fill_umac();
  for (int i = 0; i < nf; ++i)  // :160-171
    species_form_dev(P, sold[i], rho0_old, rx, spt == MGPU_PREDICT_RHOPRIME_AND_X, true, B.lo[i], B.hi[i]);
  if (rx) fill_mf(P, B, sold, ng_s, nullptr, P.spec_comp, foextrap_comp, P.nspec, pmask, true);
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) fill_mf(P, B, sold, ng_s, nullptr, P.rho_comp, foextrap_comp, 1, pmask);
  auto edge = [&](int scomp, int ncomp, bool cons) {
    for (int i = 0; i < nf; ++i) {
      DV u[3], se[3];
      box_umac(i, u);
      box_faces(sedge, i, se);
      for (int n = 0; n < ncomp; ++n) {
        if (P.bds_type != 0) {  // density_advance.f90:183-185 etc.
          size_t mark = arena_mark();
          bds_dev(P, sold[i], se, u, scal_force[i], B.lo[i], B.hi[i], scomp - 1 + n, cons, ng_s, ng_f);
          arena_release(mark);
          continue;
        }
        edge_one_comp(P, sold[i], se, u, scal_force[i], B.lo[i], B.hi[i], B.bc[i], scomp - 1 + n, dm + scomp + n, false,
                      cons, ng_s, ng_f);
      }
    }
  };
  if (spt == MGPU_PREDICT_RHOX) edge(P.spec_comp, P.nspec, true);  // :190-198
  else edge(P.spec_comp, P.nspec, false);                            // :178-186
  if (spt == MGPU_PREDICT_RHOX) {  // :204-213
    for (int i = 0; i < nf; ++i)
      for (int d = 0; d < dm; ++d) sum_comps_dev(sedge[d][i], P.rho_comp - 1, P.spec_comp - 1, P.nspec);
  } else {
    edge(P.rho_comp, 1, false);  // :216-224
  }
  for (int i = 0; i < nf; ++i)  // :229-240
    species_form_dev(P, sold[i], rho0_old, rx, spt == MGPU_PREDICT_RHOPRIME_AND_X, false, B.lo[i], B.hi[i]);
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) fill_mf(P, B, sold, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, pmask);
  if (rx) fill_mf(P, B, sold, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, pmask);
  if (P.ntrac >= 1) edge(P.trac_comp, P.ntrac, false);  // :242-252
  for (int i = 0; i < nf; ++i) {                        // :258
    DV u[3];
    box_umac(i, u);
    addw0_dev(P, u, w0, -1.0, B.lo[i], B.hi[i]);
  }
  fill_umac();
  for (int i = 0; i < nf; ++i) {  // :280-366
    FluxArgs fa;
    fill_flux_args(P, fa, B.lo[i], B.hi[i]);
    for (int d = 0; d < dm; ++d) { fa.sflux[d] = sflux[d][i]; fa.sedge[d] = sedge[d][i]; fa.umac[d] = umac[d][i]; }
    fa.eta = eta[i];
    fa.w0 = w0;
    fa.rho0_old = rho0_old;
    fa.rho0_edge_old = rho0_edge_old;
    fa.rho0_new = (which_step == 1) ? rho0_old : rho0_new;
    fa.rho0_edge_new = (which_step == 1) ? rho0_edge_old : rho0_edge_new;
    fa.rho0_predicted_edge = rho0_pe;
    set_dev(scal_force[i].p, 0.0, scal_force[i].size());  // :349-351
    UpdArgs ua;
    ua.dm = dm;
    ua.dt = P.dt;
    for (int d = 0; d < 3; ++d) ua.dx[d] = P.dx[d];
    ua.vb = grown(B.lo[i], B.hi[i], dm, 0);
    ua.sold = sold[i];
    ua.snew = snew[i];
    ua.force = scal_force[i];
    for (int d = 0; d < dm; ++d) ua.sflux[d] = sflux[d][i];
    flux_update_all_dev(P, fa, ua, g_opt_exact != 0);
  }
  fill_mf(P, B, snew, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, pmask);
  fill_mf(P, B, snew, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, pmask);
  if (P.ntrac >= 1) fill_mf(P, B, snew, ng_s, nullptr, P.trac_comp, dm + P.trac_comp, P.ntrac, pmask);
}

// ---- shared pieces of the other L4 episodes ---------------------------------------------------------
static const int (&NODAL_D)[3][3] = NODAL_OF;
static void fill_faces_dev(const mgpu_params& P, DV* u, const int* lo, const int* hi, const int* adv_bc,
                           const int* pmask) {  // addw0.f90:85-93 / mkutrans.f90:105-115
  FillBatch fb;
  for (int d = 0; d < P.dm; ++d) fill_boundary_dev(P, u[d], lo, hi, 1, NODAL_D[d], 1, 1, 1, adv_bc, pmask, false);
  fb.run();
}

// umac + w0mac for the predictions of a spherical episode (addw0 with mult = 1, its ghost fill, and after the
// predictions addw0 with mult = -1 and its ghost fill: density_advance.f90:148/258, enthalpy_advance.f90:201/293,
// velocity_advance.f90:90/115).  Exact build: in place, as the reference.  FAST build: the sum goes into a scratch copy in
// one pass and umac is left alone -- half the traffic, and umac comes back untouched where the reference returns
// (umac + w0mac) - w0mac, one rounding away.
struct UmacPlusW0 {
  const mgpu_params& P;
  DV* umac;
  const DV* w0mac;
  const int *lo, *hi, *adv_bc, *pmask;
  DV utot[3];
  bool copy;
  UmacPlusW0(const mgpu_params& P_, DV* umac_, const DV* w0mac_, const int* lo_, const int* hi_, const int* adv_bc_,
             const int* pmask_)
      : P(P_), umac(umac_), w0mac(w0mac_), lo(lo_), hi(hi_), adv_bc(adv_bc_), pmask(pmask_), copy(g_opt_exact == 0) {
    DV wm[3] = {w0mac[0], w0mac[1], w0mac[2]};
    if (copy) {
      for (int d = 0; d < 3; ++d) {
        utot[d] = umac[d];
        utot[d].p = arena_alloc((size_t)umac[d].size());
      }
      sum_faces_sphr_dev(utot, umac, wm, lo, hi);
    } else {
      addw0_sphr_dev(umac, wm, 1.0, lo, hi);
    }
    fill(copy ? utot : umac);
  }
  DV* faces() { return copy ? utot : umac; }
  void restore() {
    if (copy) return;
    DV wm[3] = {w0mac[0], w0mac[1], w0mac[2]};
    addw0_sphr_dev(umac, wm, -1.0, lo, hi);
    fill(umac);
  }
  void fill(DV* u) {
    FillBatch fb;
    for (int d = 0; d < 3; ++d) fill_boundary_dev(P, u[d], lo, hi, 1, NODAL_OF[d], 1, 1, 1, adv_bc, pmask, false);
    fb.run();
  }
};
static size_t umac_copy_bytes(const int* lo, const int* hi) {
  size_t n = 1;
  for (int d = 0; d < 3; ++d) n *= (size_t)(hi[d] - lo[d] + 1 + 3);
  return 3 * (n * sizeof(double) + 256);
}

// ---- density_advance, spherical (density_advance.f90:20 with spherical == 1) ---------------------------------------
static void density_advance_sphr_dev(const mgpu_params& P, const mgpu_geom& g, int which_step, DV& sold, DV& snew,
                                     DV* sedge, DV* sflux, DV& scal_force, DV* umac, const double* w0_h, const DV* w0mac,
                                     const double* rho0_old_h, const double* rho0_new_h, const int* lo, const int* hi,
                                     int ng_s, int ng_f, const int* adv_bc, const int* pmask) {
  const int dm = 3, spt = P.species_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  Geom gd = make_geom(P, g);
  const double* rho0_old = upload_small(rho0_old_h, (size_t)g.nr_fine);
  const double* rho0_new = upload_small(rho0_new_h, (size_t)g.nr_fine);
  const int zero3[3] = {0, 0, 0};
  auto cart_of = [&](const double* s0_dev, int ng) {  // put_1d_array_on_cart incl. its ghost fill (fill_3d_data.f90:21)
    DV c = make_view(nullptr, lo, hi, dm, ng, zero3, 1);
    c.p = arena_alloc((size_t)c.size());
    put_1d_array_on_cart_dev(P, g, gd, s0_dev, c, false, false, lo, hi);
    fill_boundary_dev(P, c, lo, hi, ng, nullptr, 1, dm + P.rho_comp, 1, adv_bc, pmask, false);
    return c;
  };
  set_dev(scal_force.p, 0.0, scal_force.size());
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {
    DV rho0_old_cart = cart_of(rho0_old, 1);
    modify_scal_force_sphr_dev(P, g, gd, scal_force, sold, umac, rho0_old_cart, w0_h, P.rho_comp,
                               spt == MGPU_PREDICT_RHO_AND_X, lo, hi);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
  }
  UmacPlusW0 uw(P, umac, w0mac, lo, hi, adv_bc, pmask);  // :148
  DV* const ue = uw.faces();
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {
    convert_rhoX_to_X_dev(P, sold, true, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, foextrap_comp, P.nspec, adv_bc, pmask, true);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
    pert_form_sphr_dev(g, gd, sold, rho0_old, P.rho_comp, true, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, foextrap_comp, 1, adv_bc, pmask, false);
  }
  auto edge = [&](int scomp, int ncomp, bool cons) {
    for (int n = 0; n < ncomp; ++n) {
      if (P.bds_type != 0) {
        size_t mark = arena_mark();
        bds_dev(P, sold, sedge, ue, scal_force, lo, hi, scomp - 1 + n, cons, ng_s, ng_f);
        arena_release(mark);
        continue;
      }
      edge_one_comp(P, sold, sedge, ue, scal_force, lo, hi, adv_bc, scomp - 1 + n, dm + scomp + n, false, cons, ng_s,
                    ng_f);
    }
  };
  edge(P.spec_comp, P.nspec, spt == MGPU_PREDICT_RHOX);
  if (spt == MGPU_PREDICT_RHOX) {
    for (int d = 0; d < dm; ++d) sum_comps_dev(sedge[d], P.rho_comp - 1, P.spec_comp - 1, P.nspec);
  } else {
    edge(P.rho_comp, 1, false);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X) {
    pert_form_sphr_dev(g, gd, sold, rho0_old, P.rho_comp, false, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
  }
  if (spt == MGPU_PREDICT_RHOPRIME_AND_X || spt == MGPU_PREDICT_RHO_AND_X) {
    convert_rhoX_to_X_dev(P, sold, false, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
  }
  if (P.ntrac >= 1) edge(P.trac_comp, P.ntrac, false);
  uw.restore();  // :258
  // rho0mac_old / rho0mac_new (make_s0mac, fill_3d_data.f90:942)
  auto mac_of = [&](const double* s0_dev, DV* mac) {
    for (int d = 0; d < 3; ++d) {
      mac[d] = make_view(nullptr, lo, hi, dm, 1, NODAL_D[d], 1);
      mac[d].p = arena_alloc((size_t)mac[d].size());
    }
    if (g.s0mac_interp_type == 1) {
      DV c = cart_of(s0_dev, 2);
      make_mac_dev(g, gd, s0_dev, mac, &c, 1, lo, hi);
    } else {
      make_mac_dev(g, gd, s0_dev, mac, nullptr, 1, lo, hi);
    }
  };
  SphrFluxArgs fa;
  fa.spt = spt;
  fa.rho = P.rho_comp - 1;
  fa.rhoh = P.rhoh_comp - 1;
  fa.vb = grown(lo, hi, dm, 0);
  mac_of(rho0_old, fa.r0o);
  if (which_step == 2) mac_of(rho0_new, fa.r0n);
  else for (int d = 0; d < 3; ++d) fa.r0n[d] = fa.r0o[d];
  for (int d = 0; d < 3; ++d) {
    fa.sflux[d] = sflux[d];
    fa.sedge[d] = sedge[d];
    fa.umac[d] = umac[d];
    fa.w0mac[d] = w0mac[d];
    fa.h0o[d] = fa.h0n[d] = fa.r0o[d];
  }
  mk_rhoX_flux_sphr_dev(fa, P.spec_comp, P.spec_comp + P.nspec - 1);
  if (P.ntrac >= 1) mk_rhoX_flux_sphr_dev(fa, P.trac_comp, P.trac_comp + P.ntrac - 1);
  set_dev(scal_force.p, 0.0, scal_force.size());
  UpdArgs ua;
  ua.dm = dm;
  ua.dt = P.dt;
  for (int d = 0; d < 3; ++d) ua.dx[d] = P.dx[d];
  ua.vb = grown(lo, hi, dm, 0);
  ua.sold = sold;
  ua.snew = snew;
  ua.force = scal_force;
  for (int d = 0; d < dm; ++d) ua.sflux[d] = sflux[d];
  update_scal_dev(P, ua, P.spec_comp, P.spec_comp + P.nspec - 1);
  {
    FillBatch fb;
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.spec_comp, dm + P.spec_comp, P.nspec, adv_bc, pmask, false);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rho_comp, dm + P.rho_comp, 1, adv_bc, pmask, false);
    fb.run();
  }
  if (P.ntrac >= 1) {
    update_scal_dev(P, ua, P.trac_comp, P.trac_comp + P.ntrac - 1);
    fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.trac_comp, dm + P.trac_comp, P.ntrac, adv_bc, pmask, false);
  }
}

static DV arena_fab(const int* lo, const int* hi, int dm, int ng, const int* nodal, int nc) {
  DV v = make_view(nullptr, lo, hi, dm, ng, nodal, nc);
  v.p = arena_alloc((size_t)v.size());
  return v;
}
static size_t fab_bytes(const int* lo, const int* hi, int dm, int ng, int nodal_dirs, int nc) {
  size_t n = 1;
  for (int d = 0; d < dm; ++d) n *= (size_t)(hi[d] - lo[d] + 1 + 2 * ng + (nodal_dirs ? 1 : 0));
  return n * nc * sizeof(double) + 256;
}
static void vel_force_dev(const mgpu_params& P, DV& force, bool is_final, const DV& uold, const DV* uedge,
                          const double* w0, const DV& gpi, const DV& rho1, const double* rho0, const double* grav,
                          const double* w0_force, const int* lo, const int* hi, int ng_f, const int* adv_bc,
                          const int* pmask, bool add_utilde = true) {
  VelForceArgs a;
  a.dm = P.dm;
  a.nr = P.nr;
  a.is_final_update = is_final;
  a.add_utilde = add_utilde;
  a.dr = P.dx[P.dm - 1];
  a.rho_cut = P.buoyancy_cutoff_factor * P.base_cutoff_density;
  a.omega = P.omega; a.sin_theta = P.sin_theta; a.cos_theta = P.cos_theta; a.rotation_radius = P.rotation_radius;
  a.vb = grown(lo, hi, P.dm, 0);
  a.force = force; a.uold = uold; a.gpi = gpi; a.rho = rho1;
  for (int d = 0; d < P.dm; ++d) a.uedge[d] = uedge[d];
  a.w0 = w0; a.rho0 = rho0; a.grav = grav; a.w0_force = w0_force;
  mk_vel_force_dev(a);
  fill_boundary_dev(P, force, lo, hi, ng_f, nullptr, 1, 1, P.dm, adv_bc, pmask, false);  // mkforce.f90:209
}

// advance_premac (Source/advance_premac.f90:21)
static size_t advance_premac_scratch(const mgpu_params& P, const int* lo, const int* hi, int ng_u) {
  const int ng_f = P.ppm_trace_forces == 1 ? ng_u : 1;
  return fab_bytes(lo, hi, P.dm, ng_u, 0, P.dm) + fab_bytes(lo, hi, P.dm, ng_f, 0, P.dm) +
         (P.dm + 1) * fab_bytes(lo, hi, P.dm, 1, 1, 1) + velpred_scratch(P, lo, hi) +
         (size_t)(6 * (P.nr + 2)) * sizeof(double) + 8192;
}
static void advance_premac_dev(const mgpu_params& P, const DV& uold, const DV& sold, DV* umac, const DV& gpi,
                               const double* w0_h, const double* w0_force_h, const double* rho0_old_h,
                               const double* grav_h, const int* lo, const int* hi, int ng_u, const int* adv_bc,
                               const int* phys_bc, const int* pmask) {
  const int dm = P.dm, nr = P.nr;
  const int ng_f = P.ppm_trace_forces == 1 ? ng_u : 1;  // advance_premac.f90:62-66
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* w0_force = upload_small(w0_force_h, nr);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* grav = upload_small(grav_h, nr);
  int z3[3] = {0, 0, 0};
  // The fused path skips both ghost fills of utrans, i.e. their halo exchanges: every rank of a slab run must take the
  // same branch, so the decision rests on a global property -- the domain is periodic in every direction (pmask) -- and
  // not on this rank's own boundaries (a middle slab of a box with walls in the slab direction has none either, while
  // its neighbours do and would wait for its sends).
  bool all_periodic = true;
  for (int d = 0; d < dm; ++d) all_periodic = all_periodic && pmask[d] != 0;
  const bool fuse = g_opt_premac_fuse && all_periodic && velpred_premac_fusable(P, phys_bc);
  // fused path: the face kernel forms ufull = w0 on the cells + uold (:75-78) per cell; only traced forces read the
  // array again afterwards
  const bool need_ufull = !fuse || P.ppm_trace_forces == 1;
  DV ufull = need_ufull ? arena_fab(lo, hi, dm, ng_u, z3, dm) : make_view(nullptr, lo, hi, dm, ng_u, z3, dm);
  DV force = arena_fab(lo, hi, dm, ng_f, z3, dm);
  DV utrans[3];
  for (int d = 0; d < dm; ++d) utrans[d] = arena_fab(lo, hi, dm, 1, NODAL_D[d], 1);
  if (need_ufull) {
    radial_cell_avg_dev(P, ufull, w0, lo, hi);  // :75 put_1d_array_on_cart(w0, ufull, 1, .true., .true.)
    fill_boundary_dev(P, ufull, lo, hi, ng_u, nullptr, 1, 1, dm, adv_bc, pmask, false);
    if (ufull.size() != uold.size()) throw Error("advance_premac: internal size mismatch");
    add_dev(ufull.p, uold.p, ufull.size());  // :76-78
  }
  if (fuse) {
    // no physical boundary: utrans (:90), utrans + w0 (:109) and the face states of velpred (:116) come from one
    // reconstruction per direction, ghost rows included, so neither ghost fill of utrans is needed
    DV utfull[3];
    for (int d = 0; d < dm; ++d) utfull[d] = (d == dm - 1) ? arena_fab(lo, hi, dm, 1, NODAL_D[d], 1) : utrans[d];
    VpArgs st;
    velpred_premac_begin(P, uold, ufull, utrans, utfull, w0, lo, hi, adv_bc, phys_bc, ng_u, &st);
    vel_force_dev(P, force, false, uold, utrans, w0, gpi, sold.comp(P.rho_comp - 1), rho0_old, grav, w0_force, lo, hi,
                  ng_f, adv_bc, pmask);  // :98
    velpred_premac_finish(P, &st, umac, force, ng_u, ng_f);
    return;
  }
  mkutrans_dev(P, uold, ufull, utrans, w0, lo, hi, adv_bc, phys_bc, ng_u);  // :90
  fill_faces_dev(P, utrans, lo, hi, adv_bc, pmask);
  vel_force_dev(P, force, false, uold, utrans, w0, gpi, sold.comp(P.rho_comp - 1), rho0_old, grav, w0_force, lo, hi, ng_f,
                adv_bc, pmask);  // :98
  addw0_dev(P, utrans, w0, 1.0, lo, hi);  // :109
  fill_faces_dev(P, utrans, lo, hi, adv_bc, pmask);
  velpred_dev(P, uold, ufull, umac, utrans, force, w0, lo, hi, adv_bc, phys_bc, ng_u, ng_f);  // :116
}

// velocity_advance (Source/velocity_advance.f90:16)
static size_t velocity_advance_scratch(const mgpu_params& P, const int* lo, const int* hi, int ng_u) {
  const int ng_f = P.ppm_trace_forces == 0 ? 1 : ng_u;
  return fab_bytes(lo, hi, P.dm, ng_f, 0, P.dm) + P.dm * fab_bytes(lo, hi, P.dm, 0, 1, P.dm) +
         std::max(make_edge_scal_scratch(P, lo, hi), bds_scratch(P, lo, hi)) + (size_t)(8 * (P.nr + 2)) * sizeof(double) +
         8192;
}
static void velocity_advance_dev(const mgpu_params& P, const DV& uold, DV& unew, const DV& sold, const DV& rhohalf,
                                 DV* umac, const DV& gpi, const double* w0_h, const double* w0_force_h,
                                 const double* rho0_old_h, const double* rho0_nph_h, const double* grav_old_h,
                                 const double* grav_nph_h, const DV& sponge, const int* lo, const int* hi, int ng_u,
                                 const int* adv_bc, const int* pmask) {
  const int dm = P.dm, nr = P.nr;
  const int ng_f = P.ppm_trace_forces == 0 ? 1 : ng_u;  // velocity_advance.f90:69-75
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* w0_force = upload_small(w0_force_h, nr);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* rho0_nph = upload_small(rho0_nph_h, nr);
  const double* grav_old = upload_small(grav_old_h, nr);
  const double* grav_nph = upload_small(grav_nph_h, nr);
  int z3[3] = {0, 0, 0};
  DV force = arena_fab(lo, hi, dm, ng_f, z3, dm);
  DV uedge[3];
  for (int d = 0; d < dm; ++d) uedge[d] = arena_fab(lo, hi, dm, 0, NODAL_D[d], dm);
  vel_force_dev(P, force, false, uold, umac, w0, gpi, sold.comp(P.rho_comp - 1), rho0_old, grav_old, w0_force, lo, hi,
                ng_f, adv_bc, pmask);  // :80
  addw0_dev(P, umac, w0, 1.0, lo, hi);  // :90
  fill_faces_dev(P, umac, lo, hi, adv_bc, pmask);
  for (int c = 0; c < dm; ++c) {  // :102-109: is_vel, comps 1..dm, bc comps 1..dm, advective form
    size_t mark = arena_mark();
    if (P.bds_type != 0) bds_dev(P, uold, uedge, umac, force, lo, hi, c, false, ng_u, ng_f);
    else edge_one_comp(P, uold, uedge, umac, force, lo, hi, adv_bc, c, 1 + c, true, false, ng_u, ng_f);
    arena_release(mark);
  }
  addw0_dev(P, umac, w0, -1.0, lo, hi);  // :115
  fill_faces_dev(P, umac, lo, hi, adv_bc, pmask);
  vel_force_dev(P, force, true, uold, umac, w0, gpi, rhohalf.comp(0), rho0_nph, grav_nph, w0_force, lo, hi, ng_f, adv_bc,
                pmask);  // :122
  VelArgs a;
  a.dm = dm;
  a.do_sponge = P.do_sponge != 0;
  a.dt = P.dt;
  for (int d = 0; d < 3; ++d) a.dx[d] = P.dx[d];
  a.vb = grown(lo, hi, dm, 0);
  a.uold = uold; a.unew = unew; a.force = force; a.sponge = sponge;
  for (int d = 0; d < dm; ++d) { a.umac[d] = umac[d]; a.uedge[d] = uedge[d]; }
  a.w0 = w0;
  update_velocity_dev(a, g_opt_exact == 0);  // :132
  fill_boundary_dev(P, unew, lo, hi, ng_u, nullptr, 1, 1, dm, adv_bc, pmask, false);  // update_vel.f90:121
}

// ---- advance_premac and velocity_advance with spherical == 1 ------------------------------------------------------
// The planar episodes above with the spherical operators: w0 reaches the cells and faces through w0_cart / w0mac, the
// force is mk_vel_force_3d_sphr (mkforce.f90:484), the update is the spherical branch of update_velocity_3d.
static void vel_force_sphr_full(const mgpu_params& P, const mgpu_geom& g, const Geom& gd, DV& force, bool is_final,
                                const DV& uold, const DV* uedge, const double* w0_h, const DV* w0mac, const DV& gpi,
                                const DV& rho1, const DV& normal, const double* rho0_h, const double* grav_h,
                                const DV& w0fc, const int* lo, const int* hi, int ng_f, const int* adv_bc,
                                const int* pmask) {
  size_t mark = arena_mark();
  mk_vel_force_sphr_dev(P, g, gd, force, is_final, uold, uedge, w0_h, w0mac, gpi, rho1, normal, rho0_h, grav_h, w0fc, lo,
                        hi, true);
  arena_release(mark);
  fill_boundary_dev(P, force, lo, hi, ng_f, nullptr, 1, 1, 3, adv_bc, pmask, false);  // mkforce.f90:209
}
static size_t sphr_force_scratch(const int* lo, const int* hi) { return 4 * fab_bytes(lo, hi, 3, 0, 0, 3); }

static void advance_premac_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const DV& uold, const DV& sold, DV* umac,
                                    const DV& gpi, const DV& normal, const double* w0_h, const DV* w0mac, const DV& w0fc,
                                    const double* rho0_old_h, const double* grav_h, const int* lo, const int* hi, int ng_u,
                                    const int* adv_bc, const int* phys_bc, const int* pmask) {
  const int dm = 3;
  const int ng_f = P.ppm_trace_forces == 1 ? ng_u : 1;  // advance_premac.f90:62-66
  Geom gd = make_geom(P, g);
  int z3[3] = {0, 0, 0};
  DV ufull = arena_fab(lo, hi, dm, ng_u, z3, dm), force = arena_fab(lo, hi, dm, ng_f, z3, dm);
  DV utrans[3];
  for (int d = 0; d < dm; ++d) utrans[d] = arena_fab(lo, hi, dm, 1, NODAL_D[d], 1);
  set_dev(ufull.p, 0.0, ufull.size());
  put_1d_array_on_cart_dev(P, g, gd, upload_small(w0_h, (size_t)g.nr_fine + 1), ufull, true, true, lo, hi);  // :75
  fill_boundary_dev(P, ufull, lo, hi, ng_u, nullptr, 1, 1, dm, adv_bc, pmask, false);
  if (ufull.size() != uold.size()) throw Error("advance_premac: internal size mismatch");
  add_dev(ufull.p, uold.p, ufull.size());  // :76-78
  mkutrans_dev(P, uold, ufull, utrans, nullptr, lo, hi, adv_bc, phys_bc, ng_u, w0mac);  // :90
  fill_faces_dev(P, utrans, lo, hi, adv_bc, pmask);
  vel_force_sphr_full(P, g, gd, force, false, uold, utrans, w0_h, w0mac, gpi, sold.comp(P.rho_comp - 1), normal, rho0_old_h,
                      grav_h, w0fc, lo, hi, ng_f, adv_bc, pmask);  // :98
  addw0_sphr_dev(utrans, w0mac, 1.0, lo, hi);  // :109
  fill_faces_dev(P, utrans, lo, hi, adv_bc, pmask);
  velpred_dev(P, uold, ufull, umac, utrans, force, nullptr, lo, hi, adv_bc, phys_bc, ng_u, ng_f, w0mac);  // :116
}

static void velocity_advance_sphr_dev(const mgpu_params& P, const mgpu_geom& g, const DV& uold, DV& unew, const DV& sold,
                                      const DV& rhohalf, DV* umac, const DV& gpi, const DV& normal, const double* w0_h,
                                      const DV* w0mac, const DV& w0fc, const double* rho0_old_h, const double* rho0_nph_h,
                                      const double* grav_old_h, const double* grav_nph_h, const DV& sponge, const int* lo,
                                      const int* hi, int ng_u, const int* adv_bc, const int* pmask) {
  const int dm = 3;
  const int ng_f = P.ppm_trace_forces == 0 ? 1 : ng_u;  // velocity_advance.f90:69-75
  Geom gd = make_geom(P, g);
  int z3[3] = {0, 0, 0};
  DV force = arena_fab(lo, hi, dm, ng_f, z3, dm);
  DV uedge[3];
  for (int d = 0; d < dm; ++d) uedge[d] = arena_fab(lo, hi, dm, 0, NODAL_D[d], dm);
  vel_force_sphr_full(P, g, gd, force, false, uold, umac, w0_h, w0mac, gpi, sold.comp(P.rho_comp - 1), normal, rho0_old_h,
                      grav_old_h, w0fc, lo, hi, ng_f, adv_bc, pmask);  // :80
  UmacPlusW0 uw(P, umac, w0mac, lo, hi, adv_bc, pmask);  // :90
  for (int c = 0; c < dm; ++c) {  // :102-109
    size_t mark = arena_mark();
    if (P.bds_type != 0) bds_dev(P, uold, uedge, uw.faces(), force, lo, hi, c, false, ng_u, ng_f);
    else edge_one_comp(P, uold, uedge, uw.faces(), force, lo, hi, adv_bc, c, 1 + c, true, false, ng_u, ng_f);
    arena_release(mark);
  }
  uw.restore();  // :115
  vel_force_sphr_full(P, g, gd, force, true, uold, umac, w0_h, w0mac, gpi, rhohalf.comp(0), normal, rho0_nph_h, grav_nph_h,
                      w0fc, lo, hi, ng_f, adv_bc, pmask);  // :122
  VelArgs a;
  a.dm = dm;
  a.do_sponge = P.do_sponge != 0;
  a.dt = P.dt;
  for (int d = 0; d < 3; ++d) a.dx[d] = P.dx[d];
  a.vb = grown(lo, hi, dm, 0);
  a.uold = uold; a.unew = unew; a.force = force; a.sponge = sponge;
  for (int d = 0; d < dm; ++d) { a.umac[d] = umac[d]; a.uedge[d] = uedge[d]; }
  a.w0 = nullptr;
  update_velocity_sphr_dev(a, w0mac);  // :132
  fill_boundary_dev(P, unew, lo, hi, ng_u, nullptr, 1, 1, dm, adv_bc, pmask, false);  // update_vel.f90:121
}

// ---- enthalpy_advance with spherical == 1 --------------------------------------------------------------------------
struct SphrCtx {  // what the spherical helpers of one episode share
  const mgpu_params& P;
  const mgpu_geom& g;
  Geom gd;
  const int *lo, *hi, *adv_bc, *pmask;
};
// put_1d_array_on_cart of a bin-centred array incl. its ghost fill with the BCs of component bccomp (fill_3d_data.f90:21)
static DV sphr_cart(const SphrCtx& X, const double* s0_dev, int ng, int bccomp) {
  const int z3[3] = {0, 0, 0};
  DV c = make_view(nullptr, X.lo, X.hi, 3, ng, z3, 1);
  c.p = arena_alloc((size_t)c.size());
  put_1d_array_on_cart_dev(X.P, X.g, X.gd, s0_dev, c, false, false, X.lo, X.hi);
  fill_boundary_dev(X.P, c, X.lo, X.hi, ng, nullptr, 1, bccomp, 1, X.adv_bc, X.pmask, false);
  return c;
}
// make_s0mac (fill_3d_data.f90:942): through the cell centres (two ghost cells) when s0mac_interp_type = 1
static void sphr_s0mac(const SphrCtx& X, const double* s0_dev, DV* mac, int bccomp) {
  for (int d = 0; d < 3; ++d) {
    mac[d] = make_view(nullptr, X.lo, X.hi, 3, 1, NODAL_D[d], 1);
    mac[d].p = arena_alloc((size_t)mac[d].size());
  }
  if (X.g.s0mac_interp_type == 1) {
    DV c = sphr_cart(X, s0_dev, 2, bccomp);
    make_mac_dev(X.g, X.gd, s0_dev, mac, &c, 1, X.lo, X.hi);
  } else {
    make_mac_dev(X.g, X.gd, s0_dev, mac, nullptr, 1, X.lo, X.hi);
  }
}
// the spherical branch of mkrhohforce (mkscalforce.f90:31): p0 = (p0_1 + p0_2)/2 on cells and faces, then the kernel,
// then the ghost fill of the rhoh component (:177-181)
static void rhoh_force_sphr_full(const SphrCtx& X, DV& scal_force, bool is_prediction, const DV& thermal, const DV* umac,
                                 const double* p0_1, const double* p0_2, const double* psi_h, bool add_thermal, int ng_f) {
  const mgpu_params& P = X.P;
  const int nr = X.g.nr_fine, foextrap_comp = 3 + P.nscal + 2;
  std::vector<double> p0(nr);
  for (int r = 0; r < nr; ++r) p0[r] = 0.5 * (p0_1[r] + p0_2[r]);
  size_t mark = arena_mark();
  const double* p0d = upload_small(p0.data(), (size_t)nr);
  DV p0c = sphr_cart(X, p0d, 1, foextrap_comp);
  DV p0m[3];
  sphr_s0mac(X, p0d, p0m, foextrap_comp);
  mkrhohforce_sphr_dev(P, X.g, X.gd, scal_force.comp(P.rhoh_comp - 1), is_prediction, thermal, umac, p0c, p0m, psi_h,
                       add_thermal, X.lo, X.hi);
  arena_release(mark);
  fill_boundary_dev(P, scal_force, X.lo, X.hi, ng_f, nullptr, P.rhoh_comp, foextrap_comp, 1, X.adv_bc, X.pmask, false);
}

static void enthalpy_advance_sphr_dev(const mgpu_params& P, const mgpu_geom& g, int which_step, DV& sold, DV& snew,
                                      DV* sedge, DV* sflux, DV& scal_force, const DV& thermal, DV* umac, const double* w0_h,
                                      const DV* w0mac, const double* rho0_old_h, const double* rhoh0_old_h,
                                      const double* rho0_new_h, const double* rhoh0_new_h, const double* p0_old_h,
                                      const double* p0_new_h, const double* tempbar_h, const double* psi_h, const int* lo,
                                      const int* hi, int ng_s, int ng_f, const int* adv_bc, const int* pmask) {
  const int dm = 3, nr = g.nr_fine;
  const int ept = P.enthalpy_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  const int rhoh = P.rhoh_comp - 1, rho = P.rho_comp - 1;
  if (ept == MGPU_PREDICT_HPRIME) throw Error("mk_rhoh_flux : predict_hprime not coded yet");  // mkflux.f90:1167
  const bool pred_T =
      (ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  if (pred_T) {
    the_eos("enthalpy_advance (temperature-based prediction: mktempforce, makeHfromRhoT_edge)");
    if (!tempbar_h) throw Error("enthalpy_advance: the temperature-based predictions need tempbar");
  }
  const int temp = P.temp_comp - 1;
  SphrCtx X{P, g, make_geom(P, g), lo, hi, adv_bc, pmask};
  const double* tempbar = pred_T ? upload_small(tempbar_h, (size_t)nr) : nullptr;
  const double* rho0_old = upload_small(rho0_old_h, (size_t)nr);
  const double* rho0_new = upload_small(rho0_new_h, (size_t)nr);
  const double* rhoh0_old = upload_small(rhoh0_old_h, (size_t)nr);
  std::vector<double> h0o(nr), h0n(nr);
  for (int r = 0; r < nr; ++r) {
    h0o[r] = rhoh0_old_h[r] / rho0_old_h[r];
    h0n[r] = rhoh0_new_h[r] / rho0_new_h[r];
  }
  const double* h0_old = upload_small(h0o.data(), (size_t)nr);
  const double* h0_new = upload_small(h0n.data(), (size_t)nr);
  auto rhoh_to_h = [&](bool flag) {  // convert_rhoh_to_h, convert_rhoX_to_X.f90:80
    comp_muldiv_dev(P, sold, rhoh, sold, rho, flag ? 0 : 1, 0, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc,
                      pmask, false);
  };
  auto pert = [&](bool flag) {
    pert_form_sphr_dev(g, X.gd, sold, rhoh0_old, P.rhoh_comp, flag, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc,
                      pmask, false);
  };
  auto pert_T = [&](bool flag) {  // :214-217, :268-272
    pert_form_sphr_dev(g, X.gd, sold, tempbar, P.temp_comp, flag, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.temp_comp, flag ? foextrap_comp : dm + P.temp_comp, 1, adv_bc,
                      pmask, false);
  };
  if (ept == MGPU_PREDICT_H) rhoh_to_h(true);     // :122-126
  set_dev(scal_force.p, 0.0, scal_force.size());  // :132-134
  if (pred_T) {  // :190-195: mktempforce with spherical == 1 (mkscalforce.f90:770-776, _3d_sphr :1026)
    size_t mark = arena_mark();
    TempForceArgs ta;
    ta.dm = dm; ta.nr = nr; ta.rho = rho; ta.temp = temp; ta.spec0 = P.spec_comp - 1;
    ta.sphr = true;
    ta.dr = g.dr;
    for (int d = 0; d < 3; ++d) ta.dx[d] = P.dx[d];
    ta.vb = grown(lo, hi, dm, 0);
    ta.f = scal_force.comp(temp); ta.s = sold; ta.thermal = thermal;
    for (int d = 0; d < dm; ++d) ta.umac[d] = umac[d];
    ta.p0_old = ta.psi = nullptr;
    ta.p0_cart = sphr_cart(X, upload_small(p0_old_h, (size_t)nr), 1, foextrap_comp);
    ta.psi_cart = arena_fab(lo, hi, 3, 0, nullptr, 1);
    put_1d_array_on_cart_dev(P, g, X.gd, upload_small(psi_h, (size_t)nr), ta.psi_cart, false, false, lo, hi);
    mktempforce_dev(ta);
    arena_release(mark);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.temp_comp, foextrap_comp, 1, adv_bc, pmask, false);
  } else {
    rhoh_force_sphr_full(X, scal_force, true, thermal, umac, p0_old_h, p0_old_h, psi_h, true, ng_f);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) {  // :141-156
    size_t mark = arena_mark();
    DV rhoh0_old_cart = sphr_cart(X, rhoh0_old, 1, dm + P.rhoh_comp);
    modify_scal_force_sphr_dev(P, g, X.gd, scal_force, sold, umac, rhoh0_old_cart, w0_h, P.rhoh_comp, false, lo, hi);
    arena_release(mark);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask, false);
  } else if (ept == MGPU_PREDICT_H) {  // :173-178
    comp_muldiv_dev(P, scal_force, rhoh, sold, rho, 0, 1, lo, hi);
  }
  UmacPlusW0 uw(P, umac, w0mac, lo, hi, adv_bc, pmask);  // :201
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(true);  // :204-208
  if (ept == MGPU_PREDICT_TPRIME_THEN_H) pert_T(true);
  {
    const bool cons = (ept == MGPU_PREDICT_RHOH);  // :232-254
    const int pc = pred_T ? temp : rhoh;           // :220-226
    size_t mark = arena_mark();
    if (P.bds_type != 0) bds_dev(P, sold, sedge, uw.faces(), scal_force, lo, hi, pc, cons, ng_s, ng_f);
    else edge_one_comp(P, sold, sedge, uw.faces(), scal_force, lo, hi, adv_bc, pc, dm + pc + 1, false, cons, ng_s, ng_f);
    arena_release(mark);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(false);  // :257-261
  if (ept == MGPU_PREDICT_TPRIME_THEN_H) pert_T(false);
  if (ept == MGPU_PREDICT_H) rhoh_to_h(false);     // :272-276
  if (pred_T) {  // :280-286: makeHfromRhoT_edge with spherical == 1 (rhoh_vs_t.f90:84-105, _3d_sphr :596)
    size_t mark = arena_mark();
    std::vector<double> r0h(nr), rh0h(nr), t0h(nr);
    for (int r = 0; r < nr; ++r) {
      r0h[r] = 0.5 * (rho0_old_h[r] + rho0_new_h[r]);
      rh0h[r] = 0.5 * (rhoh0_old_h[r] + rhoh0_new_h[r]);
      t0h[r] = 0.5 * (tempbar_h[r] + tempbar_h[r]);
    }
    HEdgeArgs ha{};
    ha.dm = dm; ha.ept = ept; ha.spt = P.species_pred_type;
    ha.rho = rho; ha.rhoh = rhoh; ha.temp = temp; ha.spec0 = P.spec_comp - 1;
    ha.sphr = true;
    ha.vb = grown(lo, hi, dm, 0);
    for (int d = 0; d < dm; ++d) ha.sedge[d] = sedge[d];
    ha.rho0_cart = sphr_cart(X, upload_small(r0h.data(), (size_t)nr), 2, dm + P.rho_comp);
    ha.rhoh0_cart = sphr_cart(X, upload_small(rh0h.data(), (size_t)nr), 2, dm + P.rhoh_comp);
    ha.t0_cart = sphr_cart(X, upload_small(t0h.data(), (size_t)nr), 2, dm + P.temp_comp);
    h_from_rhot_edge_dev(ha);
    arena_release(mark);
  }
  uw.restore();  // :293
  const bool s1 = (which_step == 1);
  {
    size_t mark = arena_mark();
    SphrFluxArgs fa;
    fa.spt = P.species_pred_type;
    fa.rho = rho;
    fa.rhoh = rhoh;
    fa.vb = grown(lo, hi, dm, 0);
    sphr_s0mac(X, rho0_old, fa.r0o, dm + P.rho_comp);  // :301-399
    sphr_s0mac(X, h0_old, fa.h0o, foextrap_comp);
    if (s1) {
      for (int d = 0; d < 3; ++d) { fa.r0n[d] = fa.r0o[d]; fa.h0n[d] = fa.h0o[d]; }
    } else {
      sphr_s0mac(X, rho0_new, fa.r0n, dm + P.rho_comp);
      sphr_s0mac(X, h0_new, fa.h0n, foextrap_comp);
    }
    for (int d = 0; d < 3; ++d) { fa.sflux[d] = sflux[d]; fa.sedge[d] = sedge[d]; fa.umac[d] = umac[d]; fa.w0mac[d] = w0mac[d]; }
    mk_rhoh_flux_sphr_dev(P, fa);
    arena_release(mark);
  }
  // :401-403 setval(scal_force, ZERO, all): every component but the predicted one is still zero from :132-134
  set_dev(scal_force.p + scal_force.cs * (pred_T ? temp : rhoh), 0.0, scal_force.cs);
  rhoh_force_sphr_full(X, scal_force, false, thermal, umac, p0_old_h, s1 ? p0_old_h : p0_new_h, psi_h, false, ng_f);  // :405-416
  UpdArgs ua;
  ua.dm = dm;
  ua.dt = P.dt;
  for (int d = 0; d < 3; ++d) ua.dx[d] = P.dx[d];
  ua.vb = grown(lo, hi, dm, 0);
  ua.sold = sold; ua.snew = snew; ua.force = scal_force;
  for (int d = 0; d < dm; ++d) ua.sflux[d] = sflux[d];
  size_t mark_p0 = arena_mark();
  if (P.do_eos_h_above_cutoff && have_eos()) {  // :418-424: p0_new on the cell centres (only the EOS reset reads it)
    ua.p0_new_cart = sphr_cart(X, upload_small(p0_new_h, (size_t)nr), 1, foextrap_comp);
    ua.have_p0_new_cart = true;
  }
  update_scal_dev(P, ua, P.rhoh_comp, P.rhoh_comp);  // :431
  arena_release(mark_p0);
  fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rhoh_comp, dm + P.rhoh_comp, 1, adv_bc, pmask, false);
}

// enthalpy_advance (Source/enthalpy_advance.f90:16)
static void enthalpy_advance_dev(const mgpu_params& P, int which_step, DV& sold, DV& snew, DV* sedge, DV* sflux,
                                 DV& scal_force, const DV& thermal, DV* umac, const double* w0_h,
                                 const double* rho0_old_h, const double* rhoh0_old_h, const double* rho0_new_h,
                                 const double* rhoh0_new_h, const double* p0_old_h, const double* p0_new_h,
                                 const double* tempbar_h, const double* psi_h, const double* grav_old_h,
                                 const double* grav_nph_h, const int* lo, const int* hi, int ng_s, int ng_f,
                                 const int* adv_bc, const int* pmask) {
  const int dm = P.dm, nr = P.nr;
  const int ept = P.enthalpy_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  if (ept == MGPU_PREDICT_HPRIME) throw Error("mk_rhoh_flux : predict_hprime not coded yet");  // mkflux.f90:1167
  const bool pred_T =
      (ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  if (pred_T) {
    the_eos("enthalpy_advance (temperature-based prediction: mktempforce, makeHfromRhoT_edge)");
    if (!tempbar_h) throw Error("enthalpy_advance: the temperature-based predictions need tempbar");
  }
  std::vector<double> e[4] = {std::vector<double>(nr + 1), std::vector<double>(nr + 1), std::vector<double>(nr + 1),
                              std::vector<double>(nr + 1)};
  std::vector<double> t0e(nr + 1);
  if (pred_T) cell_to_edge_host(tempbar_h, t0e.data(), nr);  // :118-119 (old and new are both tempbar)
  const double* tempbar = pred_T ? upload_small(tempbar_h, nr) : nullptr;
  const double* t0_edge = pred_T ? upload_small(t0e.data(), nr + 1) : nullptr;
  const int temp = P.temp_comp - 1;
  cell_to_edge_host(rho0_old_h, e[0].data(), nr);  // enthalpy_advance.f90:114-117
  cell_to_edge_host(rho0_new_h, e[1].data(), nr);
  cell_to_edge_host(rhoh0_old_h, e[2].data(), nr);
  cell_to_edge_host(rhoh0_new_h, e[3].data(), nr);
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* rho0_new = upload_small(rho0_new_h, nr);
  const double* rhoh0_old = upload_small(rhoh0_old_h, nr);
  const double* rhoh0_new = upload_small(rhoh0_new_h, nr);
  const double* p0_old = upload_small(p0_old_h, nr);
  const double* p0_new = upload_small(p0_new_h, nr);
  const double* psi = upload_small(psi_h, nr);
  const double* grav_old = upload_small(grav_old_h, nr);
  const double* grav_nph = upload_small(grav_nph_h, nr);
  const double* r0e_old = upload_small(e[0].data(), nr + 1);
  const double* r0e_new = upload_small(e[1].data(), nr + 1);
  const double* rh0e_old = upload_small(e[2].data(), nr + 1);
  const double* rh0e_new = upload_small(e[3].data(), nr + 1);
  const int rhoh = P.rhoh_comp - 1, rho = P.rho_comp - 1;
  auto rhoh_force = [&](bool is_pred, const double* p02, const double* r02, const double* grav, bool add_thermal) {
    if (is_pred && !(ept == MGPU_PREDICT_RHOHPRIME || ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH))
      throw Error("ERROR: should only call mkrhohforce when predicting rhoh', h, or rhoh");
    RhohForceArgs a;
    a.dm = dm; a.nr = nr; a.cutoff_coord = P.base_cutoff_density_coord;
    a.with_psi = (is_pred && (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH)) || !is_pred;
    a.add_thermal = add_thermal;
    a.dr = P.dx[dm - 1];
    a.vb = grown(lo, hi, dm, 0);
    a.f = scal_force.comp(rhoh); a.thermal = thermal; a.wm = umac[dm - 1];
    a.p0_1 = p0_old; a.p0_2 = p02; a.rho0_1 = rho0_old; a.rho0_2 = r02; a.grav = grav; a.psi = psi;
    mkrhohforce_dev(a);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask, false);
  };
  auto rhoh_to_h = [&](bool flag) {  // convert_rhoh_to_h
    comp_muldiv_dev(P, sold, rhoh, sold, rho, flag ? 0 : 1, 0, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc,
                      pmask, false);
  };
  auto pert = [&](bool flag) {
    put_in_pert_form_dev(P, sold, rhoh0_old, P.rhoh_comp, flag, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, adv_bc,
                      pmask, false);
  };
  auto pert_T = [&](bool flag) {  // put_in_pert_form on the temperature with tempbar (:214-217, :268-272)
    put_in_pert_form_dev(P, sold, tempbar, P.temp_comp, flag, lo, hi);
    fill_boundary_dev(P, sold, lo, hi, ng_s, nullptr, P.temp_comp, flag ? foextrap_comp : dm + P.temp_comp, 1, adv_bc,
                      pmask, false);
  };
  if (ept == MGPU_PREDICT_H) rhoh_to_h(true);  // :122-126
  set_dev(scal_force.p, 0.0, scal_force.size());  // :132-134
  if (pred_T) {  // :190-195: mktempforce (mkscalforce.f90:719) and its ghost fill (:833-837)
    TempForceArgs ta;
    ta.dm = dm; ta.nr = nr; ta.rho = rho; ta.temp = temp; ta.spec0 = P.spec_comp - 1;
    ta.sphr = false;
    ta.dr = P.dx[dm - 1];
    for (int d = 0; d < 3; ++d) ta.dx[d] = P.dx[d];
    ta.vb = grown(lo, hi, dm, 0);
    ta.f = scal_force.comp(temp); ta.s = sold; ta.thermal = thermal;
    for (int d = 0; d < dm; ++d) ta.umac[d] = umac[d];
    ta.p0_old = p0_old; ta.psi = psi;
    mktempforce_dev(ta);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.temp_comp, foextrap_comp, 1, adv_bc, pmask, false);
  } else {
    rhoh_force(true, p0_old, rho0_old, grav_old, true);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) {  // :153-156
    modify_scal_force_dev(P, scal_force, sold, umac, rhoh0_old, rh0e_old, w0, P.rhoh_comp, false, lo, hi, g_opt_exact == 0);
    fill_boundary_dev(P, scal_force, lo, hi, ng_f, nullptr, P.rhoh_comp, foextrap_comp, 1, adv_bc, pmask, false);
  } else if (ept == MGPU_PREDICT_H) {  // :173-178
    comp_muldiv_dev(P, scal_force, rhoh, sold, rho, 0, 1, lo, hi);
  }
  addw0_dev(P, umac, w0, 1.0, lo, hi);  // :201
  fill_faces_dev(P, umac, lo, hi, adv_bc, pmask);
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(true);  // :204-208
  if (ept == MGPU_PREDICT_TPRIME_THEN_H) pert_T(true);
  {
    const bool cons = (ept == MGPU_PREDICT_RHOH);  // :232-254
    const int pc = pred_T ? temp : rhoh;           // :220-226
    size_t mark = arena_mark();
    if (P.bds_type != 0) bds_dev(P, sold, sedge, umac, scal_force, lo, hi, pc, cons, ng_s, ng_f);
    else edge_one_comp(P, sold, sedge, umac, scal_force, lo, hi, adv_bc, pc, dm + pc + 1, false, cons, ng_s, ng_f);
    arena_release(mark);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(false);  // :257-261
  if (ept == MGPU_PREDICT_TPRIME_THEN_H) pert_T(false);
  if (ept == MGPU_PREDICT_H) rhoh_to_h(false);     // :272-276
  if (pred_T) {                                    // :280-286: makeHfromRhoT_edge (rhoh_vs_t.f90:20)
    HEdgeArgs ha;
    ha.dm = dm; ha.ept = ept; ha.spt = P.species_pred_type;
    ha.rho = rho; ha.rhoh = rhoh; ha.temp = temp; ha.spec0 = P.spec_comp - 1;
    ha.sphr = false;
    ha.vb = grown(lo, hi, dm, 0);
    for (int d = 0; d < dm; ++d) ha.sedge[d] = sedge[d];
    ha.rho0_old = rho0_old; ha.rhoh0_old = rhoh0_old; ha.t0_old = tempbar;
    ha.rho0_edge_old = r0e_old; ha.rhoh0_edge_old = rh0e_old; ha.t0_edge_old = t0_edge;
    ha.rho0_new = rho0_new; ha.rhoh0_new = rhoh0_new; ha.t0_new = tempbar;
    ha.rho0_edge_new = r0e_new; ha.rhoh0_edge_new = rh0e_new; ha.t0_edge_new = t0_edge;
    ha.rho0_cart = ha.rhoh0_cart = ha.t0_cart = sold;
    h_from_rhot_edge_dev(ha);
  }
  addw0_dev(P, umac, w0, -1.0, lo, hi);             // :293
  fill_faces_dev(P, umac, lo, hi, adv_bc, pmask);
  const bool s1 = (which_step == 1);
  FluxArgs fa;
  fill_flux_args(P, fa, lo, hi);
  for (int d = 0; d < dm; ++d) { fa.sflux[d] = sflux[d]; fa.sedge[d] = sedge[d]; fa.umac[d] = umac[d]; }
  fa.w0 = w0;
  fa.rho0_old = rho0_old; fa.rho0_edge_old = r0e_old;
  fa.rho0_new = s1 ? rho0_old : rho0_new; fa.rho0_edge_new = s1 ? r0e_old : r0e_new;
  fa.rhoh0_old = rhoh0_old; fa.rhoh0_edge_old = rh0e_old;
  fa.rhoh0_new = s1 ? rhoh0_old : rhoh0_new; fa.rhoh0_edge_new = s1 ? rh0e_old : rh0e_new;
  mk_rhoh_flux_dev(P, fa);  // :326 / :375
  // :401-403 setval(scal_force, ZERO, all): every component but the predicted one is still zero from :132-134
  set_dev(scal_force.p + scal_force.cs * (pred_T ? temp : rhoh), 0.0, scal_force.cs);
  rhoh_force(false, s1 ? p0_old : p0_new, s1 ? rho0_old : rho0_new, s1 ? grav_old : grav_nph, false);  // :405-416
  UpdArgs ua;
  ua.dm = dm;
  ua.dt = P.dt;
  for (int d = 0; d < 3; ++d) ua.dx[d] = P.dx[d];
  ua.vb = grown(lo, hi, dm, 0);
  ua.sold = sold; ua.snew = snew; ua.force = scal_force;
  for (int d = 0; d < dm; ++d) ua.sflux[d] = sflux[d];
  ua.p0_new = p0_new;
  update_scal_dev(P, ua, P.rhoh_comp, P.rhoh_comp);  // :431
  fill_boundary_dev(P, snew, lo, hi, ng_s, nullptr, P.rhoh_comp, dm + P.rhoh_comp, 1, adv_bc, pmask, false);
}

// ---- the other three L4 episodes over several boxes of one rank (planar) -----------------------------------------------
// Same statements as the single-box episodes above, stage by stage over the boxes, every ghost fill through the
// multifab fill (box-to-box copies, periodic images, per-box physical BCs).  B.bc[i] is box i's adv_bc table.
struct MfBase {  // base-state arrays of one level on the device
  const double *w0, *w0_force, *rho0_a, *rho0_b, *grav_a, *grav_b;
};
static void vel_force_box(const mgpu_params& P, DV& force, bool is_final, const DV& uold, const DV* uedge, const double* w0,
                          const DV& gpi, const DV& rho1, const double* rho0, const double* grav, const double* w0_force,
                          const int* lo, const int* hi) {
  VelForceArgs a;
  a.dm = P.dm;
  a.nr = P.nr;
  a.is_final_update = is_final;
  a.add_utilde = true;
  a.dr = P.dx[P.dm - 1];
  a.rho_cut = P.buoyancy_cutoff_factor * P.base_cutoff_density;
  a.omega = P.omega; a.sin_theta = P.sin_theta; a.cos_theta = P.cos_theta; a.rotation_radius = P.rotation_radius;
  a.vb = grown(lo, hi, P.dm, 0);
  a.force = force; a.uold = uold; a.gpi = gpi; a.rho = rho1;
  for (int d = 0; d < P.dm; ++d) a.uedge[d] = uedge[d];
  a.w0 = w0; a.rho0 = rho0; a.grav = grav; a.w0_force = w0_force;
  mk_vel_force_dev(a);
}
static void fill_faces_mf(const mgpu_params& P, const BoxSet& B, std::vector<DV>* u, const int* pmask) {
  for (int d = 0; d < P.dm; ++d) fill_mf(P, B, u[d], 1, NODAL_OF[d], 1, 1, 1, pmask);
}
static void faces_of(std::vector<DV>* f, int i, int dm, DV* out) {
  for (int d = 0; d < dm; ++d) out[d] = f[d][i];
}

// velocity_advance (velocity_advance.f90:16) over the boxes
static void velocity_advance_mf_dev(const mgpu_params& P, const BoxSet& B, std::vector<DV>& uold, std::vector<DV>& unew,
                                    std::vector<DV>& sold, std::vector<DV>& rhohalf, std::vector<DV>* umac,
                                    std::vector<DV>& gpi, std::vector<DV>& sponge, const double* w0_h,
                                    const double* w0_force_h, const double* rho0_old_h, const double* rho0_nph_h,
                                    const double* grav_old_h, const double* grav_nph_h, int ng_u, const int* pmask) {
  const int dm = P.dm, nr = P.nr, nf = B.n;
  const int ng_f = P.ppm_trace_forces == 0 ? 1 : ng_u;  // :69-75
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* w0_force = upload_small(w0_force_h, nr);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* rho0_nph = upload_small(rho0_nph_h, nr);
  const double* grav_old = upload_small(grav_old_h, nr);
  const double* grav_nph = upload_small(grav_nph_h, nr);
  int z3[3] = {0, 0, 0};
  std::vector<DV> force(nf), uedge[3];
  for (int i = 0; i < nf; ++i) force[i] = arena_fab(B.lo[i], B.hi[i], dm, ng_f, z3, dm);
  for (int d = 0; d < dm; ++d) {
    uedge[d].resize(nf);
    for (int i = 0; i < nf; ++i) uedge[d][i] = arena_fab(B.lo[i], B.hi[i], dm, 0, NODAL_D[d], dm);
  }
  auto force_stage = [&](bool is_final, std::vector<DV>& rho_src, int rho_c, const double* rho0, const double* grav) {
    for (int i = 0; i < nf; ++i) {
      DV u[3];
      faces_of(umac, i, dm, u);
      vel_force_box(P, force[i], is_final, uold[i], u, w0, gpi[i], rho_src[i].comp(rho_c), rho0, grav, w0_force, B.lo[i],
                    B.hi[i]);
    }
    fill_mf(P, B, force, ng_f, nullptr, 1, 1, dm, pmask);  // mkforce.f90:209
  };
  auto addw0_stage = [&](double mult) {
    for (int i = 0; i < nf; ++i) {
      DV u[3];
      faces_of(umac, i, dm, u);
      addw0_dev(P, u, w0, mult, B.lo[i], B.hi[i]);
    }
    fill_faces_mf(P, B, umac, pmask);
  };
  force_stage(false, sold, P.rho_comp - 1, rho0_old, grav_old);  // :80
  addw0_stage(1.0);                                              // :90
  for (int i = 0; i < nf; ++i) {                                 // :102-109
    DV u[3], ue[3];
    faces_of(umac, i, dm, u);
    faces_of(uedge, i, dm, ue);
    for (int c = 0; c < dm; ++c) {
      size_t mark = arena_mark();
      if (P.bds_type != 0) bds_dev(P, uold[i], ue, u, force[i], B.lo[i], B.hi[i], c, false, ng_u, ng_f);
      else edge_one_comp(P, uold[i], ue, u, force[i], B.lo[i], B.hi[i], B.bc[i], c, 1 + c, true, false, ng_u, ng_f);
      arena_release(mark);
    }
  }
  addw0_stage(-1.0);                                            // :115
  force_stage(true, rhohalf, 0, rho0_nph, grav_nph);            // :122
  for (int i = 0; i < nf; ++i) {                                // :132
    VelArgs a;
    a.dm = dm;
    a.do_sponge = P.do_sponge != 0;
    a.dt = P.dt;
    for (int d = 0; d < 3; ++d) a.dx[d] = P.dx[d];
    a.vb = grown(B.lo[i], B.hi[i], dm, 0);
    a.uold = uold[i]; a.unew = unew[i]; a.force = force[i]; a.sponge = sponge[i];
    for (int d = 0; d < dm; ++d) { a.umac[d] = umac[d][i]; a.uedge[d] = uedge[d][i]; }
    a.w0 = w0;
    update_velocity_dev(a, g_opt_exact == 0);
  }
  fill_mf(P, B, unew, ng_u, nullptr, 1, 1, dm, pmask);  // update_vel.f90:121
}

// advance_premac (advance_premac.f90:21) over the boxes; phys[i]: box i's phys_bc table
static void advance_premac_mf_dev(const mgpu_params& P, const BoxSet& B, const std::vector<const int*>& phys,
                                  std::vector<DV>& uold, std::vector<DV>& sold, std::vector<DV>* umac, std::vector<DV>& gpi,
                                  const double* w0_h, const double* w0_force_h, const double* rho0_old_h,
                                  const double* grav_h, int ng_u, const int* pmask) {
  const int dm = P.dm, nr = P.nr, nf = B.n;
  const int ng_f = P.ppm_trace_forces == 1 ? ng_u : 1;  // :62-66
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* w0_force = upload_small(w0_force_h, nr);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* grav = upload_small(grav_h, nr);
  int z3[3] = {0, 0, 0};
  std::vector<DV> ufull(nf), force(nf), utrans[3];
  for (int i = 0; i < nf; ++i) {
    ufull[i] = arena_fab(B.lo[i], B.hi[i], dm, ng_u, z3, dm);
    force[i] = arena_fab(B.lo[i], B.hi[i], dm, ng_f, z3, dm);
  }
  for (int d = 0; d < dm; ++d) {
    utrans[d].resize(nf);
    for (int i = 0; i < nf; ++i) utrans[d][i] = arena_fab(B.lo[i], B.hi[i], dm, 1, NODAL_D[d], 1);
  }
  for (int i = 0; i < nf; ++i) radial_cell_avg_dev(P, ufull[i], w0, B.lo[i], B.hi[i]);  // :75
  fill_mf(P, B, ufull, ng_u, nullptr, 1, 1, dm, pmask);
  for (int i = 0; i < nf; ++i) {  // :76-78
    if (ufull[i].size() != uold[i].size()) throw Error("advance_premac: uold must carry the ghost cells of the state");
    add_dev(ufull[i].p, uold[i].p, ufull[i].size());
  }
  for (int i = 0; i < nf; ++i) {  // :90
    DV ut[3];
    faces_of(utrans, i, dm, ut);
    mkutrans_dev(P, uold[i], ufull[i], ut, w0, B.lo[i], B.hi[i], B.bc[i], phys[i], ng_u);
  }
  fill_faces_mf(P, B, utrans, pmask);
  for (int i = 0; i < nf; ++i) {  // :98
    DV ut[3];
    faces_of(utrans, i, dm, ut);
    vel_force_box(P, force[i], false, uold[i], ut, w0, gpi[i], sold[i].comp(P.rho_comp - 1), rho0_old, grav, w0_force,
                  B.lo[i], B.hi[i]);
  }
  fill_mf(P, B, force, ng_f, nullptr, 1, 1, dm, pmask);
  for (int i = 0; i < nf; ++i) {  // :109
    DV ut[3];
    faces_of(utrans, i, dm, ut);
    addw0_dev(P, ut, w0, 1.0, B.lo[i], B.hi[i]);
  }
  fill_faces_mf(P, B, utrans, pmask);
  for (int i = 0; i < nf; ++i) {  // :116
    DV ut[3], um[3];
    faces_of(utrans, i, dm, ut);
    faces_of(umac, i, dm, um);
    size_t mark = arena_mark();
    velpred_dev(P, uold[i], ufull[i], um, ut, force[i], w0, B.lo[i], B.hi[i], B.bc[i], phys[i], ng_u, ng_f);
    arena_release(mark);
  }
}

// enthalpy_advance (enthalpy_advance.f90:16) over the boxes: predict_rhoh / predict_rhohprime / predict_h (the
// temperature-based predictions take one box per rank)
static void enthalpy_advance_mf_dev(const mgpu_params& P, int which_step, const BoxSet& B, std::vector<DV>& sold,
                                    std::vector<DV>& snew, std::vector<DV>* sedge, std::vector<DV>* sflux,
                                    std::vector<DV>& scal_force, std::vector<DV>& thermal, std::vector<DV>* umac,
                                    const double* w0_h, const double* rho0_old_h, const double* rhoh0_old_h,
                                    const double* rho0_new_h, const double* rhoh0_new_h, const double* p0_old_h,
                                    const double* p0_new_h, const double* psi_h, const double* grav_old_h,
                                    const double* grav_nph_h, int ng_s, int ng_f, const int* pmask) {
  const int dm = P.dm, nr = P.nr, nf = B.n;
  const int ept = P.enthalpy_pred_type;
  const int foextrap_comp = dm + P.nscal + 2;
  if (ept == MGPU_PREDICT_HPRIME) throw Error("mk_rhoh_flux : predict_hprime not coded yet");  // mkflux.f90:1167
  if (!(ept == MGPU_PREDICT_RHOH || ept == MGPU_PREDICT_RHOHPRIME || ept == MGPU_PREDICT_H))
    throw Error("enthalpy_advance over several boxes: the temperature-based predictions take one box per rank");
  std::vector<double> e[4] = {std::vector<double>(nr + 1), std::vector<double>(nr + 1), std::vector<double>(nr + 1),
                              std::vector<double>(nr + 1)};
  cell_to_edge_host(rho0_old_h, e[0].data(), nr);  // :114-117
  cell_to_edge_host(rho0_new_h, e[1].data(), nr);
  cell_to_edge_host(rhoh0_old_h, e[2].data(), nr);
  cell_to_edge_host(rhoh0_new_h, e[3].data(), nr);
  const double* w0 = upload_small(w0_h, nr + 1);
  const double* rho0_old = upload_small(rho0_old_h, nr);
  const double* rho0_new = upload_small(rho0_new_h, nr);
  const double* rhoh0_old = upload_small(rhoh0_old_h, nr);
  const double* rhoh0_new = upload_small(rhoh0_new_h, nr);
  const double* p0_old = upload_small(p0_old_h, nr);
  const double* p0_new = upload_small(p0_new_h, nr);
  const double* psi = upload_small(psi_h, nr);
  const double* grav_old = upload_small(grav_old_h, nr);
  const double* grav_nph = upload_small(grav_nph_h, nr);
  const double* r0e_old = upload_small(e[0].data(), nr + 1);
  const double* r0e_new = upload_small(e[1].data(), nr + 1);
  const double* rh0e_old = upload_small(e[2].data(), nr + 1);
  const double* rh0e_new = upload_small(e[3].data(), nr + 1);
  const int rhoh = P.rhoh_comp - 1, rho = P.rho_comp - 1;
  auto rhoh_force = [&](bool is_pred, const double* p02, const double* r02, const double* grav, bool add_thermal) {
    for (int i = 0; i < nf; ++i) {
      RhohForceArgs a;
      a.dm = dm; a.nr = nr; a.cutoff_coord = P.base_cutoff_density_coord;
      a.with_psi = (is_pred && (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH)) || !is_pred;
      a.add_thermal = add_thermal;
      a.dr = P.dx[dm - 1];
      a.vb = grown(B.lo[i], B.hi[i], dm, 0);
      a.f = scal_force[i].comp(rhoh); a.thermal = thermal[i]; a.wm = umac[dm - 1][i];
      a.p0_1 = p0_old; a.p0_2 = p02; a.rho0_1 = rho0_old; a.rho0_2 = r02; a.grav = grav; a.psi = psi;
      mkrhohforce_dev(a);
    }
    fill_mf(P, B, scal_force, ng_f, nullptr, P.rhoh_comp, foextrap_comp, 1, pmask);
  };
  auto rhoh_to_h = [&](bool flag) {  // convert_rhoh_to_h
    for (int i = 0; i < nf; ++i) comp_muldiv_dev(P, sold[i], rhoh, sold[i], rho, flag ? 0 : 1, 0, B.lo[i], B.hi[i]);
    fill_mf(P, B, sold, ng_s, nullptr, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, pmask);
  };
  auto pert = [&](bool flag) {
    for (int i = 0; i < nf; ++i) put_in_pert_form_dev(P, sold[i], rhoh0_old, P.rhoh_comp, flag, B.lo[i], B.hi[i]);
    fill_mf(P, B, sold, ng_s, nullptr, P.rhoh_comp, flag ? foextrap_comp : dm + P.rhoh_comp, 1, pmask);
  };
  auto addw0_stage = [&](double mult) {
    for (int i = 0; i < nf; ++i) {
      DV u[3];
      faces_of(umac, i, dm, u);
      addw0_dev(P, u, w0, mult, B.lo[i], B.hi[i]);
    }
    fill_faces_mf(P, B, umac, pmask);
  };
  if (ept == MGPU_PREDICT_H) rhoh_to_h(true);  // :122-126
  for (int i = 0; i < nf; ++i) set_dev(scal_force[i].p, 0.0, scal_force[i].size());  // :132-134
  rhoh_force(true, p0_old, rho0_old, grav_old, true);
  if (ept == MGPU_PREDICT_RHOHPRIME) {  // :153-156
    for (int i = 0; i < nf; ++i) {
      DV u[3];
      faces_of(umac, i, dm, u);
      modify_scal_force_dev(P, scal_force[i], sold[i], u, rhoh0_old, rh0e_old, w0, P.rhoh_comp, false, B.lo[i], B.hi[i],
                            g_opt_exact == 0);
    }
    fill_mf(P, B, scal_force, ng_f, nullptr, P.rhoh_comp, foextrap_comp, 1, pmask);
  } else if (ept == MGPU_PREDICT_H) {  // :173-178
    for (int i = 0; i < nf; ++i) comp_muldiv_dev(P, scal_force[i], rhoh, sold[i], rho, 0, 1, B.lo[i], B.hi[i]);
  }
  addw0_stage(1.0);                               // :201
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(true);  // :204-208
  for (int i = 0; i < nf; ++i) {                  // :232-254
    const bool cons = (ept == MGPU_PREDICT_RHOH);
    DV u[3], se[3];
    faces_of(umac, i, dm, u);
    faces_of(sedge, i, dm, se);
    size_t mark = arena_mark();
    if (P.bds_type != 0) bds_dev(P, sold[i], se, u, scal_force[i], B.lo[i], B.hi[i], rhoh, cons, ng_s, ng_f);
    else edge_one_comp(P, sold[i], se, u, scal_force[i], B.lo[i], B.hi[i], B.bc[i], rhoh, dm + P.rhoh_comp, false, cons, ng_s,
                       ng_f);
    arena_release(mark);
  }
  if (ept == MGPU_PREDICT_RHOHPRIME) pert(false);  // :257-261
  if (ept == MGPU_PREDICT_H) rhoh_to_h(false);     // :272-276
  addw0_stage(-1.0);                               // :293
  const bool s1 = (which_step == 1);
  for (int i = 0; i < nf; ++i) {  // :326 / :375
    FluxArgs fa;
    fill_flux_args(P, fa, B.lo[i], B.hi[i]);
    for (int d = 0; d < dm; ++d) { fa.sflux[d] = sflux[d][i]; fa.sedge[d] = sedge[d][i]; fa.umac[d] = umac[d][i]; }
    fa.w0 = w0;
    fa.rho0_old = rho0_old; fa.rho0_edge_old = r0e_old;
    fa.rho0_new = s1 ? rho0_old : rho0_new; fa.rho0_edge_new = s1 ? r0e_old : r0e_new;
    fa.rhoh0_old = rhoh0_old; fa.rhoh0_edge_old = rh0e_old;
    fa.rhoh0_new = s1 ? rhoh0_old : rhoh0_new; fa.rhoh0_edge_new = s1 ? rh0e_old : rh0e_new;
    mk_rhoh_flux_dev(P, fa);
  }
  for (int i = 0; i < nf; ++i) set_dev(scal_force[i].p + scal_force[i].cs * rhoh, 0.0, scal_force[i].cs);  // :401-403
  rhoh_force(false, s1 ? p0_old : p0_new, s1 ? rho0_old : rho0_new, s1 ? grav_old : grav_nph, false);  // :405-416
  for (int i = 0; i < nf; ++i) {  // :431
    UpdArgs ua;
    ua.dm = dm;
    ua.dt = P.dt;
    for (int d = 0; d < 3; ++d) ua.dx[d] = P.dx[d];
    ua.vb = grown(B.lo[i], B.hi[i], dm, 0);
    ua.sold = sold[i]; ua.snew = snew[i]; ua.force = scal_force[i];
    for (int d = 0; d < dm; ++d) ua.sflux[d] = sflux[d][i];
    ua.p0_new = p0_new;
    update_scal_dev(P, ua, P.rhoh_comp, P.rhoh_comp);
  }
  fill_mf(P, B, snew, ng_s, nullptr, P.rhoh_comp, dm + P.rhoh_comp, 1, pmask);
}

// which components of the caller's fabs an enthalpy_advance episode reads / writes (host-pointer calls copy only these)
struct EnthalpyMasks {
  cmask_t mrhoh, sold_in, sold_out, snew_in, sedge_in, sedge_out, force_out;
};
static EnthalpyMasks enthalpy_masks(const mgpu_params& P) {
  const int ept = P.enthalpy_pred_type;
  const bool pred_T =
      (ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H);
  const cmask_t mrho = crange(P.rho_comp - 1, 1), mrhoh = crange(P.rhoh_comp - 1, 1), mtemp = crange(P.temp_comp - 1, 1);
  const cmask_t mspec = crange(P.spec_comp - 1, P.nspec);
  const bool eos_reset = P.do_eos_h_above_cutoff != 0;  // reads rho, X of snew and the temperature of sold
  EnthalpyMasks m;
  m.mrhoh = mrhoh;
  m.sold_in = mrho | mrhoh | ((pred_T || eos_reset) ? (mtemp | mspec) : 0);
  m.sold_out = mrhoh | (pred_T ? mtemp : 0);
  m.snew_in = mrhoh | (eos_reset ? (mrho | mspec) : 0);
  m.sedge_in = mrho | (pred_T ? (mspec | mtemp) : 0);  // the x-face T' is read by the other faces (QUIRK) after it is written
  m.sedge_out = mrhoh | (pred_T ? mtemp : 0);
  m.force_out = mrhoh;  // zeroed again at :401-403: only the force of the final update survives the episode
  return m;
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
  if (!g_copy_stream) MGPU_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
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
  for (auto& kv : g_res) {
    if (kv.second.pinned) cudaHostUnregister(const_cast<double*>(kv.first));
    cudaFree(kv.second.d);
  }
  g_res.clear();
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
  if (k == "defaults") {  // every switch back to its initial value (tests call this between cases)
    g_opt_fused = 1; g_opt_kchunk = -1; g_opt_async_upload = 1; g_opt_exact = 0; g_opt_leanplus = 1;
    g_opt_premac_fuse = 1; g_opt_overlap = 1;
    bds_set_fast(1); velpred_set_fast(1); sphr_set_fast(1); fused_edge3_set_split(1); fused_edge3_set_thin_edge(1); fused_edge3_set_pair_streams(1); fused_edge_set_variant(1);
    fused_edge2_set_by(MGPU_FUSED2_BY); fused_edge2d_set_tile(2);
  }
  else if (k == "fused") g_opt_fused = value;
  else if (k == "kchunk") g_opt_kchunk = value;
  else if (k == "async_upload") g_opt_async_upload = value;
  else if (k == "exact") { g_opt_exact = value; bds_set_fast(value == 0); velpred_set_fast(value == 0); sphr_set_fast(value == 0); }
  else if (k == "leanplus") g_opt_leanplus = value;
  else if (k == "premac_fuse") g_opt_premac_fuse = value;
  else if (k == "split_tiles") fused_edge3_set_split(value);
  else if (k == "thin_edge") fused_edge3_set_thin_edge(value);
  else if (k == "pair_streams") fused_edge3_set_pair_streams(value);
  else if (k == "overlap") g_opt_overlap = value;
  else if (k == "fused_variant") fused_edge_set_variant(value);
  else if (k == "fused_by") fused_edge2_set_by(value);
  else if (k == "tile2d") fused_edge2d_set_tile(value);
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
int mgpu_copy_bytes(long* h2d, long* d2h, int reset) {
  if (h2d) *h2d = g_h2d_bytes;
  if (d2h) *d2h = g_d2h_bytes;
  if (reset) g_h2d_bytes = g_d2h_bytes = 0;
  return 0;
}

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

/* ---- residency registry -------------------------------------------------------------------------------------- */
int mgpu_register(double* hptr, long n, int pin) {
  MGPU_TRY
  require_init();
  if (!hptr || n <= 0) throw Error("mgpu_register: null pointer or empty region");
  if (g_res.count(hptr)) throw Error("mgpu_register: the fab is already registered");
  Resident r;
  r.n = (size_t)n;
  MGPU_CUDA(cudaMalloc((void**)&r.d, r.n * sizeof(double)));
  if (pin) {
    cudaError_t e = cudaHostRegister(hptr, r.n * sizeof(double), cudaHostRegisterDefault);
    if (e == cudaSuccess) r.pinned = true;
    else if (e == cudaErrorHostMemoryAlreadyRegistered) cudaGetLastError();
    else { cudaFree(r.d); MGPU_CUDA(e); }
  }
  g_res[hptr] = r;
  MGPU_CATCH
}
static Resident& resident_of(const double* hptr, const char* who) {
  auto it = g_res.find(hptr);
  if (it == g_res.end()) throw Error(std::string(who) + ": the fab is not registered");
  return it->second;
}
static cmask_t comp_mask(const Resident& r, int comp0, int ncomp, const char* who) {
  if (ncomp < 0) return ALLC;  // every component
  if (comp0 < 0 || comp0 + ncomp > 64) throw Error(std::string(who) + ": component range outside 0..63");
  (void)r;
  return ncomp == 0 ? 0 : ((ncomp >= 64 ? ALLC : ((cmask_t(1) << ncomp) - 1)) << comp0);
}
/* the host changed components comp0 .. comp0+ncomp-1 (0-based; ncomp < 0: all): the mirror is stale */
int mgpu_invalidate(double* hptr, int comp0, int ncomp) {
  MGPU_TRY
  Resident& r = resident_of(hptr, "mgpu_invalidate");
  const cmask_t m = comp_mask(r, comp0, ncomp, "mgpu_invalidate");
  r.host_dirty |= m;
  r.dev_dirty &= ~m;
  MGPU_CATCH
}
/* bring the host copy of the components up to date (copies only what a call wrote since the last download) */
int mgpu_download(double* hptr, int comp0, int ncomp) {
  MGPU_TRY
  require_init();
  Resident& r = resident_of(hptr, "mgpu_download");
  const cmask_t m = comp_mask(r, comp0, ncomp, "mgpu_download") & r.dev_dirty;
  if (m && r.cs == 0) throw Error("mgpu_download: no call has used this fab yet");
  Call::for_runs(m, r.nc, [&](int c0, int n) {
    MGPU_CUDA(cudaMemcpyAsync(hptr + r.cs * c0, r.d + r.cs * c0, r.cs * n * sizeof(double), cudaMemcpyDeviceToHost,
                              g_ctx.stream));
    g_d2h_bytes += (long)(r.cs * n * sizeof(double));
  });
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  r.dev_dirty &= ~m;
  MGPU_CATCH
}
/* eager upload of the stale components (optional: calls upload what they read on their own) */
int mgpu_upload(double* hptr, int comp0, int ncomp) {
  MGPU_TRY
  require_init();
  Resident& r = resident_of(hptr, "mgpu_upload");
  if (r.cs == 0) throw Error("mgpu_upload: the shape of the fab is known after the first call that uses it");
  const cmask_t m = comp_mask(r, comp0, ncomp, "mgpu_upload") & r.host_dirty;
  Call::for_runs(m, r.nc, [&](int c0, int n) {
    MGPU_CUDA(cudaMemcpyAsync(r.d + r.cs * c0, hptr + r.cs * c0, r.cs * n * sizeof(double), cudaMemcpyHostToDevice,
                              g_ctx.stream));
    g_h2d_bytes += (long)(r.cs * n * sizeof(double));
  });
  r.host_dirty &= ~m;
  MGPU_CATCH
}
/* downloads nothing: call mgpu_download first if the host needs the device's results */
int mgpu_unregister(double* hptr) {
  MGPU_TRY
  auto it = g_res.find(hptr);
  if (it == g_res.end()) throw Error("mgpu_unregister: the fab is not registered");
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  if (it->second.pinned) cudaHostUnregister(hptr);
  cudaFree(it->second.d);
  g_res.erase(it);
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
                     const mgpu_fab* const* sflux, const mgpu_fab* force, const double* p0_new,
                     const mgpu_fab* p0_new_cart) {
  MGPU_TRY
  Call c(p, (size_t)(p->nr + 8) * sizeof(double) + 4096);
  const bool eos_reset = p->do_eos_h_above_cutoff && nstart == p->rhoh_comp && have_eos();
  const double* p0_d = (eos_reset && !p->spherical && p0_new) ? upload_small(p0_new, (size_t)p->nr) : nullptr;
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
    a.p0_new = p0_d;
    if (eos_reset && p->spherical && p0_new_cart) {
      a.p0_new_cart = c.view(p0_new_cart[i], true, false);
      a.have_p0_new_cart = true;
    }
    update_scal_dev(*p, a, nstart, nstop);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_update_velocity(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                         const mgpu_fab* const* umac, const mgpu_fab* const* uedge, const mgpu_fab* force,
                         const mgpu_fab* sponge, const double* w0) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_update_velocity: spherical == 1 is served by mgpu_update_velocity_sphr");
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
  if (p->spherical) throw Error("mgpu_addw0: spherical == 1 is served by mgpu_addw0_sphr");
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
  if (p->spherical) throw Error("mgpu_modify_scal_force: spherical == 1 is served by mgpu_modify_scal_force_sphr");
  Call c(p, (size_t)(3 * (p->nr + 2)) * sizeof(double) + 4096);
  const double* s0d = upload_small(s0, p->nr);
  const double* s0ed = upload_small(s0_edge, p->nr + 1);
  const double* w0d = upload_small(w0, p->nr + 1);
  for (int i = 0; i < nfabs; ++i) {
    DV fv = c.view(force[i], true, true), sv = c.view(s[i], true, false);
    DV um[3];
    c.views(umac, i, true, false, um);
    modify_scal_force_dev(*p, fv, sv, um, s0d, s0ed, w0d, comp, fullform != 0, s[i].lo, s[i].hi, g_opt_exact == 0);
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
  if (p->spherical) throw Error("mgpu_put_in_pert_form: spherical == 1 is served by mgpu_put_in_pert_form_sphr");
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
  if (p->spherical) throw Error("mgpu_density_advance: spherical == 1 is served by mgpu_density_advance_sphr");
  size_t one_comp = sizeof(double) + 256;  // the lean+ path keeps 1/rho of the whole fab in the arena
  for (int d = 0; d < p->dm; ++d) one_comp *= (size_t)(sold->hi[d] - sold->lo[d] + 1 + 2 * sold->ng);
  Call c(p, std::max(make_edge_scal_scratch(*p, sold->lo, sold->hi), bds_scratch(*p, sold->lo, sold->hi)) + one_comp +
             (size_t)(12 * (p->nr + 16)) * sizeof(double) + 8192);
  // components the episode reads / writes (density_advance.f90:101-366): rho, the species and the tracers
  const cmask_t adv = crange(p->rho_comp - 1, 1) | crange(p->spec_comp - 1, p->nspec) |
                      (p->ntrac >= 1 ? crange(p->trac_comp - 1, p->ntrac) : 0);
  const cmask_t flx = crange(p->spec_comp - 1, p->nspec) | (p->ntrac >= 1 ? crange(p->trac_comp - 1, p->ntrac) : 0);
  const cmask_t edg = p->species_pred_type == MGPU_PREDICT_RHOX ? adv : adv;  // rho edge = sum of rhoX edges or predicted
  // lean+ form: sold and umac are not rewritten and every ghost cell of the advanced components of snew is refilled,
  // so neither needs to travel back / in over PCIe
  const bool leanp = density_advance_is_leanplus(*p, adv_bc, pmask);
  // lean+ on one rank consumes its inputs in a known order (umac, rho, species one by one, tracers): uploads go to the
  // copy stream in that order and the episode waits for each component where it first reads it
  g_async_uploads = leanp && c.host && g_opt_async_upload && comm_size() == 1 && p->dm == 3;
  DV se[3], sf[3], um[3];
  c.views((const mgpu_fab* const*)umac, 0, true, !leanp, um);
  // sold: rho and species are transformed in place and restored (round trips); tracers are read only
  DV so = c.view(*sold, adv, leanp ? (cmask_t)0 : (crange(p->rho_comp - 1, 1) | crange(p->spec_comp - 1, p->nspec)));
  // snew: only the advanced components are written; their ghost corners next to physical walls keep the caller's
  // values (multifab_physbc.f90:165-175), hence copy-in of exactly those components
  DV sn = c.view(*snew, leanp ? (cmask_t)0 : adv, adv);
  DV fv = c.view(*scal_force, (cmask_t)0, (cmask_t)0);  // zeroed on entry (:101) and again before the update (:349)
  DV eta = c.view(*etarhoflux, true, true);
  c.views((const mgpu_fab* const*)sedge, 0, (cmask_t)0, edg, se);  // every face of the predicted components is written
  c.views((const mgpu_fab* const*)sflux, 0, (cmask_t)0, flx, sf);
  g_async_uploads = false;  // (pending uploads stay pending until the episode or finish() waits for them)
  c.zero_on_host(*scal_force);
  density_advance_dev(*p, which_step, so, sn, se, sf, fv, um, w0, eta, rho0_old, rho0_new, rho0_predicted_edge,
                      sold->lo, sold->hi, sold->ng, scal_force->ng, adv_bc, pmask);
  c.finish();
  MGPU_CATCH
}

// per-box adv_bc tables from the domain's (define_bc_tower.f90:150-197, 199-294: a face that does not lie on the
// domain boundary is INTERIOR for every component)
static void box_bc_tables(const mgpu_params& P, int nfabs, const mgpu_fab* f, const int* adv_bc, std::vector<std::vector<int>>& out) {
  const int dm = P.dm, nbc = dm + P.nscal + 3;
  out.assign(nfabs, std::vector<int>(adv_bc, adv_bc + dm * 2 * nbc));
  for (int i = 0; i < nfabs; ++i)
    for (int d = 0; d < dm; ++d) {
      const bool at[2] = {f[i].lo[d] == P.domlo[d], f[i].hi[d] == P.domhi[d]};
      for (int side = 0; side < 2; ++side)
        if (!at[side])
          for (int c = 0; c < nbc; ++c) out[i][d + dm * (side + 2 * c)] = MGPU_BC_INTERIOR;
    }
}

int mgpu_fill_boundary_mf(const mgpu_params* p, int nfabs, mgpu_fab* s, int scomp, int bccomp, int ncomp,
                          const int* adv_bc, const int* pmask) {
  MGPU_TRY
  if (nfabs < 1) throw Error("mgpu_fill_boundary_mf: nfabs must be at least 1");
  Call c(p, (size_t)nfabs * nfabs * 27 * 256 + 65536);
  std::vector<DV> v(nfabs);
  std::vector<std::vector<int>> bcs;
  box_bc_tables(*p, nfabs, s, adv_bc, bcs);
  BoxSet B;
  B.n = nfabs;
  for (int i = 0; i < nfabs; ++i) {
    v[i] = c.view(s[i], true, true);
    B.lo.push_back(s[i].lo); B.hi.push_back(s[i].hi); B.bc.push_back(bcs[i].data());
    if (s[i].ng != s[0].ng || s[i].nc != s[0].nc) throw Error("mgpu_fill_boundary_mf: the boxes of a multifab share ng and nc");
  }
  const bool nod = s[0].nodal[0] || s[0].nodal[1] || s[0].nodal[2];
  fill_mf(*p, B, v, s[0].ng, nod ? s[0].nodal : nullptr, scomp, bccomp, ncomp, pmask);
  c.finish();
  MGPU_CATCH
}

/* density_advance for a multifab of nfabs boxes on this rank (one level); arguments as mgpu_density_advance with
 * arrays of nfabs fabs (sedge / sflux / umac: dm arrays of nfabs fabs).  adv_bc is the DOMAIN's table. */
int mgpu_density_advance_mf(const mgpu_params* p, int which_step, int nfabs, mgpu_fab* sold, mgpu_fab* snew,
                            mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force, mgpu_fab* const* umac,
                            const double* w0, mgpu_fab* etarhoflux, const double* rho0_old, const double* rho0_new,
                            const double* p0_dummy, const double* rho0_predicted_edge, const int* adv_bc,
                            const int* pmask) {
  MGPU_TRY
  (void)p0_dummy;
  if (p->spherical) throw Error("mgpu_density_advance_mf: the spherical episode takes one box per rank (mgpu_density_advance_sphr)");
  if (nfabs < 1) throw Error("mgpu_density_advance_mf: nfabs must be at least 1");
  size_t scratch = 0;
  for (int i = 0; i < nfabs; ++i)
    scratch = std::max(scratch, std::max(make_edge_scal_scratch(*p, sold[i].lo, sold[i].hi), bds_scratch(*p, sold[i].lo, sold[i].hi)));
  Call c(p, scratch + (size_t)(12 * (p->nr + 16)) * sizeof(double) + (size_t)nfabs * nfabs * 27 * 256 * 16 + 65536);
  const int dm = p->dm;
  std::vector<std::vector<int>> bcs;
  box_bc_tables(*p, nfabs, sold, adv_bc, bcs);
  BoxSet B;
  B.n = nfabs;
  std::vector<DV> so(nfabs), sn(nfabs), fv(nfabs), et(nfabs), se[3], sf[3], um[3];
  for (int d = 0; d < dm; ++d) { se[d].resize(nfabs); sf[d].resize(nfabs); um[d].resize(nfabs); }
  for (int i = 0; i < nfabs; ++i) {
    B.lo.push_back(sold[i].lo); B.hi.push_back(sold[i].hi); B.bc.push_back(bcs[i].data());
    so[i] = c.view(sold[i], true, true);
    sn[i] = c.view(snew[i], true, true);
    fv[i] = c.view(scal_force[i], false, true);
    et[i] = c.view(etarhoflux[i], true, true);
    for (int d = 0; d < dm; ++d) {
      se[d][i] = c.view(sedge[d][i], true, true);
      sf[d][i] = c.view(sflux[d][i], true, true);
      um[d][i] = c.view(umac[d][i], true, true);
    }
  }
  density_advance_mf_dev(*p, which_step, B, so, sn, se, sf, fv, um, w0, et, rho0_old, rho0_new, rho0_predicted_edge,
                         sold[0].ng, scal_force[0].ng, pmask);
  c.finish();
  MGPU_CATCH
}

// per-box phys_bc tables: a side that is not on the domain boundary is INTERIOR
static void box_phys_tables(const mgpu_params& P, int nfabs, const mgpu_fab* f, const int* phys_bc, std::vector<std::vector<int>>& out) {
  const int dm = P.dm;
  out.assign(nfabs, std::vector<int>(phys_bc, phys_bc + dm * 2));
  for (int i = 0; i < nfabs; ++i)
    for (int d = 0; d < dm; ++d) {
      if (f[i].lo[d] != P.domlo[d]) out[i][d] = MGPU_BC_INTERIOR;
      if (f[i].hi[d] != P.domhi[d]) out[i][d + dm] = MGPU_BC_INTERIOR;
    }
}
static size_t boxset_scratch(int nfabs) { return (size_t)nfabs * nfabs * 27 * 256 * 16 + 65536; }
static void mf_check(const mgpu_params* p, int nfabs, const char* who) {
  if (p->spherical) throw Error(std::string(who) + ": spherical geometry takes one box per rank");
  if (nfabs < 1) throw Error(std::string(who) + ": nfabs must be at least 1");
}

int mgpu_velocity_advance_mf(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew, const mgpu_fab* sold,
                             const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0,
                             const double* w0_force, const double* rho0_old, const double* rho0_nph,
                             const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                             const int* adv_bc, const int* pmask) {
  MGPU_TRY
  mf_check(p, nfabs, "mgpu_velocity_advance_mf");
  const int dm = p->dm;
  size_t scratch = 0, edge = 0;
  for (int i = 0; i < nfabs; ++i) {
    const int ng_f = p->ppm_trace_forces == 0 ? 1 : uold[i].ng;
    scratch += fab_bytes(uold[i].lo, uold[i].hi, dm, ng_f, 0, dm) + dm * fab_bytes(uold[i].lo, uold[i].hi, dm, 0, 1, dm);
    edge = std::max(edge, std::max(make_edge_scal_scratch(*p, uold[i].lo, uold[i].hi), bds_scratch(*p, uold[i].lo, uold[i].hi)));
  }
  Call c(p, scratch + edge + (size_t)(8 * (p->nr + 2)) * sizeof(double) + boxset_scratch(nfabs));
  std::vector<std::vector<int>> bcs;
  box_bc_tables(*p, nfabs, uold, adv_bc, bcs);
  BoxSet B;
  B.n = nfabs;
  std::vector<DV> uo(nfabs), un(nfabs), so(nfabs), rh(nfabs), gp(nfabs), sp(nfabs), um[3];
  for (int d = 0; d < dm; ++d) um[d].resize(nfabs);
  for (int i = 0; i < nfabs; ++i) {
    B.lo.push_back(uold[i].lo); B.hi.push_back(uold[i].hi); B.bc.push_back(bcs[i].data());
    uo[i] = c.view(uold[i], true, false);
    un[i] = c.view(unew[i], true, true);
    so[i] = c.view(sold[i], crange(p->rho_comp - 1, 1), (cmask_t)0);
    rh[i] = c.view(rhohalf[i], true, false);
    gp[i] = c.view(gpi[i], true, false);
    sp[i] = c.view(sponge[i], true, false);
    for (int d = 0; d < dm; ++d) um[d][i] = c.view(umac[d][i], true, true);
  }
  velocity_advance_mf_dev(*p, B, uo, un, so, rh, um, gp, sp, w0, w0_force, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph,
                          uold[0].ng, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_advance_premac_mf(const mgpu_params* p, int nfabs, const mgpu_fab* uold, const mgpu_fab* sold,
                           mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0, const double* w0_force,
                           const double* rho0_old, const double* grav_cell_old, const int* adv_bc, const int* phys_bc,
                           const int* pmask) {
  MGPU_TRY
  mf_check(p, nfabs, "mgpu_advance_premac_mf");
  const int dm = p->dm;
  size_t scratch = 0, vp = 0;
  for (int i = 0; i < nfabs; ++i) {
    const int ng_u = uold[i].ng, ng_f = p->ppm_trace_forces == 1 ? ng_u : 1;
    scratch += fab_bytes(uold[i].lo, uold[i].hi, dm, ng_u, 0, dm) + fab_bytes(uold[i].lo, uold[i].hi, dm, ng_f, 0, dm) +
               dm * fab_bytes(uold[i].lo, uold[i].hi, dm, 1, 1, 1);
    vp = std::max(vp, velpred_scratch(*p, uold[i].lo, uold[i].hi));
  }
  Call c(p, scratch + vp + (size_t)(8 * (p->nr + 2)) * sizeof(double) + boxset_scratch(nfabs));
  std::vector<std::vector<int>> bcs, phs;
  box_bc_tables(*p, nfabs, uold, adv_bc, bcs);
  box_phys_tables(*p, nfabs, uold, phys_bc, phs);
  BoxSet B;
  B.n = nfabs;
  std::vector<const int*> phys;
  std::vector<DV> uo(nfabs), so(nfabs), gp(nfabs), um[3];
  for (int d = 0; d < dm; ++d) um[d].resize(nfabs);
  for (int i = 0; i < nfabs; ++i) {
    B.lo.push_back(uold[i].lo); B.hi.push_back(uold[i].hi); B.bc.push_back(bcs[i].data());
    phys.push_back(phs[i].data());
    uo[i] = c.view(uold[i], true, false);
    so[i] = c.view(sold[i], crange(p->rho_comp - 1, 1), (cmask_t)0);
    gp[i] = c.view(gpi[i], true, false);
    for (int d = 0; d < dm; ++d) um[d][i] = c.view(umac[d][i], true, true);
  }
  advance_premac_mf_dev(*p, B, phys, uo, so, um, gp, w0, w0_force, rho0_old, grav_cell_old, uold[0].ng, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_enthalpy_advance_mf(const mgpu_params* p, int which_step, int nfabs, mgpu_fab* sold, mgpu_fab* snew,
                             mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                             const mgpu_fab* thermal, mgpu_fab* const* umac, const double* w0, const double* rho0_old,
                             const double* rhoh0_old, const double* rho0_new, const double* rhoh0_new,
                             const double* p0_old, const double* p0_new, const double* psi, const double* grav_old,
                             const double* grav_nph, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  mf_check(p, nfabs, "mgpu_enthalpy_advance_mf");
  const int dm = p->dm;
  size_t edge = 0;
  for (int i = 0; i < nfabs; ++i)
    edge = std::max(edge, std::max(make_edge_scal_scratch(*p, sold[i].lo, sold[i].hi), bds_scratch(*p, sold[i].lo, sold[i].hi)));
  Call c(p, edge + (size_t)(16 * (p->nr + 2)) * sizeof(double) + boxset_scratch(nfabs));
  std::vector<std::vector<int>> bcs;
  box_bc_tables(*p, nfabs, sold, adv_bc, bcs);
  const EnthalpyMasks m = enthalpy_masks(*p);
  BoxSet B;
  B.n = nfabs;
  std::vector<DV> so(nfabs), sn(nfabs), fv(nfabs), th(nfabs), se[3], sf[3], um[3];
  for (int d = 0; d < dm; ++d) { se[d].resize(nfabs); sf[d].resize(nfabs); um[d].resize(nfabs); }
  for (int i = 0; i < nfabs; ++i) {
    B.lo.push_back(sold[i].lo); B.hi.push_back(sold[i].hi); B.bc.push_back(bcs[i].data());
    so[i] = c.view(sold[i], m.sold_in, m.sold_out);
    sn[i] = c.view(snew[i], m.snew_in, m.mrhoh);
    fv[i] = c.view(scal_force[i], (cmask_t)0, m.force_out);
    c.zero_on_host(scal_force[i], ~m.force_out);
    th[i] = c.view(thermal[i], true, false);
    for (int d = 0; d < dm; ++d) {
      se[d][i] = c.view(sedge[d][i], m.sedge_in | (sedge[d][i].ng > 0 ? m.sedge_out : (cmask_t)0), m.sedge_out);
      sf[d][i] = c.view(sflux[d][i], (cmask_t)0, m.mrhoh);
      um[d][i] = c.view(umac[d][i], true, true);
    }
  }
  enthalpy_advance_mf_dev(*p, which_step, B, so, sn, se, sf, fv, th, um, w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old,
                          p0_new, psi, grav_old, grav_nph, sold[0].ng, scal_force[0].ng, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_mkrhohforce(const mgpu_params* p, int nfabs, mgpu_fab* scal_force, int is_prediction, const mgpu_fab* thermal,
                     const mgpu_fab* const* umac, const double* p0_1, const double* p0_2, const double* rho0_1,
                     const double* rho0_2, const double* grav, const double* psi, int add_thermal) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_mkrhohforce: spherical == 1 is served by mgpu_mkrhohforce_sphr");
  const int ept = p->enthalpy_pred_type;
  if (is_prediction && !(ept == MGPU_PREDICT_RHOHPRIME || ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH))
    throw Error("ERROR: should only call mkrhohforce when predicting rhoh', h, or rhoh");  // mkscalforce.f90:87-92
  Call c(p, (size_t)(8 * (p->nr + 2)) * sizeof(double) + 8192);
  const int nr = p->nr;
  RhohForceArgs a;
  a.dm = p->dm; a.nr = nr; a.cutoff_coord = p->base_cutoff_density_coord;
  a.with_psi = (is_prediction && (ept == MGPU_PREDICT_H || ept == MGPU_PREDICT_RHOH)) || !is_prediction;
  a.add_thermal = add_thermal != 0;
  a.dr = p->dx[p->dm - 1];
  a.p0_1 = upload_small(p0_1, nr); a.p0_2 = upload_small(p0_2, nr);
  a.rho0_1 = upload_small(rho0_1, nr); a.rho0_2 = upload_small(rho0_2, nr);
  a.grav = upload_small(grav, nr); a.psi = upload_small(psi, nr);
  for (int i = 0; i < nfabs; ++i) {
    a.vb = grown(scal_force[i].lo, scal_force[i].hi, p->dm, 0);
    const cmask_t m = crange(p->rhoh_comp - 1, 1);
    a.f = c.view(scal_force[i], m, m).comp(p->rhoh_comp - 1);
    a.thermal = c.view(thermal[i], true, false);
    DV um[3];
    c.views(umac, i, true, false, um);
    a.wm = um[p->dm - 1];
    mkrhohforce_dev(a);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mk_vel_force(const mgpu_params* p, int nfabs, mgpu_fab* vel_force, int is_final_update, const mgpu_fab* uold,
                      const mgpu_fab* const* uedge, const double* w0, const mgpu_fab* gpi, const mgpu_fab* s,
                      int index_rho, const double* rho0, const double* grav, const double* w0_force,
                      int do_add_utilde_force) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_mk_vel_force: spherical == 1 is served by mgpu_mk_vel_force_sphr");
  Call c(p, (size_t)(6 * (p->nr + 2)) * sizeof(double) + 8192);
  const int nr = p->nr;
  VelForceArgs a;
  a.dm = p->dm; a.nr = nr;
  a.is_final_update = is_final_update != 0;
  a.add_utilde = do_add_utilde_force != 0;
  a.dr = p->dx[p->dm - 1];
  a.rho_cut = p->buoyancy_cutoff_factor * p->base_cutoff_density;
  a.omega = p->omega; a.sin_theta = p->sin_theta; a.cos_theta = p->cos_theta; a.rotation_radius = p->rotation_radius;
  a.w0 = upload_small(w0, nr + 1); a.rho0 = upload_small(rho0, nr);
  a.grav = upload_small(grav, nr); a.w0_force = upload_small(w0_force, nr);
  for (int i = 0; i < nfabs; ++i) {
    a.vb = grown(vel_force[i].lo, vel_force[i].hi, p->dm, 0);
    a.force = c.view(vel_force[i], false, true);
    a.uold = c.view(uold[i], true, false);
    a.gpi = c.view(gpi[i], true, false);
    const cmask_t m = crange(index_rho - 1, 1);
    a.rho = c.view(s[i], m, (cmask_t)0).comp(index_rho - 1);
    c.views(uedge, i, true, false, a.uedge);
    mk_vel_force_dev(a);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_advance_premac(const mgpu_params* p, const mgpu_fab* uold, const mgpu_fab* sold, mgpu_fab* const* umac,
                        const mgpu_fab* gpi, const double* w0, const double* w0_force, const double* rho0_old,
                        const double* grav_cell_old, const int* adv_bc, const int* phys_bc, const int* pmask) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_advance_premac: spherical == 1 is served by mgpu_advance_premac_sphr");
  Call c(p, advance_premac_scratch(*p, uold->lo, uold->hi, uold->ng));
  DV uo = c.view(*uold, true, false), gp = c.view(*gpi, true, false);
  DV so = c.view(*sold, crange(p->rho_comp - 1, 1), (cmask_t)0);
  DV um[3];
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);  // ghost faces keep the caller's values
  advance_premac_dev(*p, uo, so, um, gp, w0, w0_force, rho0_old, grav_cell_old, uold->lo, uold->hi, uold->ng, adv_bc,
                     phys_bc, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_velocity_advance(const mgpu_params* p, const mgpu_fab* uold, mgpu_fab* unew, const mgpu_fab* sold,
                          const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi, const double* w0,
                          const double* w0_force, const double* rho0_old, const double* rho0_nph,
                          const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                          const int* adv_bc, const int* pmask) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_velocity_advance: spherical == 1 is served by mgpu_velocity_advance_sphr");
  Call c(p, velocity_advance_scratch(*p, uold->lo, uold->hi, uold->ng));
  DV uo = c.view(*uold, true, false), un = c.view(*unew, true, true), gp = c.view(*gpi, true, false);
  DV so = c.view(*sold, crange(p->rho_comp - 1, 1), (cmask_t)0), rh = c.view(*rhohalf, true, false);
  DV sp = c.view(*sponge, true, false);
  DV um[3];
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  velocity_advance_dev(*p, uo, un, so, rh, um, gp, w0, w0_force, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph, sp,
                       uold->lo, uold->hi, uold->ng, adv_bc, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_enthalpy_advance(const mgpu_params* p, int which_step, mgpu_fab* sold, mgpu_fab* snew, mgpu_fab* const* sedge,
                          mgpu_fab* const* sflux, mgpu_fab* scal_force, const mgpu_fab* thermal, mgpu_fab* const* umac,
                          const double* w0, const double* rho0_old, const double* rhoh0_old, const double* rho0_new,
                          const double* rhoh0_new, const double* p0_old, const double* p0_new, const double* tempbar,
                          const double* psi, const double* grav_old, const double* grav_nph, const int* adv_bc,
                          const int* pmask) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_enthalpy_advance: use mgpu_enthalpy_advance_sphr with spherical == 1");
  Call c(p, std::max(make_edge_scal_scratch(*p, sold->lo, sold->hi), bds_scratch(*p, sold->lo, sold->hi)) +
                (size_t)(16 * (p->nr + 2)) * sizeof(double) + 16384);
  // the episode reads rho and rhoh of sold (rhoh is transformed in place and restored), writes rhoh of snew, of the
  // edge states and of the fluxes, and reads the density edge states density_advance left in sedge(rho_comp)
  const EnthalpyMasks m = enthalpy_masks(*p);
  DV so = c.view(*sold, m.sold_in, m.sold_out), sn = c.view(*snew, m.snew_in, m.mrhoh);
  DV fv = c.view(*scal_force, (cmask_t)0, m.force_out);  // zeroed on entry; only the predicted component is non-zero on return
  c.zero_on_host(*scal_force, ~m.force_out);
  DV th = c.view(*thermal, true, false);
  DV se[3], sf[3], um[3];
  // edge-state fabs with ghost cells: the episode writes the valid faces only, so what it writes is read first
  c.views((const mgpu_fab* const*)sedge, 0, m.sedge_in | (sedge[0]->ng > 0 ? m.sedge_out : (cmask_t)0), m.sedge_out, se);
  c.views((const mgpu_fab* const*)sflux, 0, (cmask_t)0, m.mrhoh, sf);
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  enthalpy_advance_dev(*p, which_step, so, sn, se, sf, fv, th, um, w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old,
                       p0_new, tempbar, psi, grav_old, grav_nph, sold->lo, sold->hi, sold->ng, scal_force->ng, adv_bc,
                       pmask);
  c.finish();
  MGPU_CATCH
}

// ---- spherical geometry (mgpu_sphr.cu) ----------------------------------------------------------------------
static size_t geom_scratch(const mgpu_geom* g) { return (size_t)(6 * (g->nr_fine + 4)) * sizeof(double) + 8192; }
static void need_sphr(const mgpu_params* p, const mgpu_geom* g) {
  if (p->dm != 3) throw Error("spherical geometry is 3-D only");
  if (g && g->nr_fine < 3) throw Error("spherical geometry: nr_fine must be at least 3");
}

int mgpu_put_1d_array_on_cart(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0, mgpu_fab* s0_cart,
                              int is_input_edge_centered, int is_output_a_vector) {
  MGPU_TRY
  need_sphr(p, g);
  Call c(p, geom_scratch(g));
  Geom gd = make_geom(*p, *g);
  const double* s0d = upload_small(s0, (size_t)g->nr_fine + (is_input_edge_centered ? 1 : 0));
  for (int i = 0; i < nfabs; ++i) {
    DV cv = c.view(s0_cart[i], true, true);
    put_1d_array_on_cart_dev(*p, *g, gd, s0d, cv, is_input_edge_centered != 0, is_output_a_vector != 0, s0_cart[i].lo,
                             s0_cart[i].hi);
  }
  c.finish();
  MGPU_CATCH
}

static int make_mac_api(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0, mgpu_fab* const* mac,
                        const mgpu_fab* cart, int kind) {
  MGPU_TRY
  need_sphr(p, g);
  Call c(p, geom_scratch(g));
  Geom gd = make_geom(*p, *g);
  const double* s0d = upload_small(s0, (size_t)g->nr_fine + (kind == 0 ? 1 : 0));
  for (int i = 0; i < nfabs; ++i) {
    if (mac[0][i].ng != 1)  // fill_3d_data.f90:588, :992
      throw Error(kind == 0 ? "Error: make_w0mac_3d_sphr assumes one ghost cell"
                            : "Error: make_s0mac assumes one ghost cell in s0mac");
    DV m[3], cv;
    c.views((const mgpu_fab* const*)mac, i, true, true, m);
    if (cart) cv = c.view(cart[i], true, false);
    make_mac_dev(*g, gd, s0d, m, cart ? &cv : nullptr, kind, mac[0][i].lo, mac[0][i].hi);
  }
  c.finish();
  MGPU_CATCH
}
int mgpu_make_w0mac(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* w0, mgpu_fab* const* w0mac,
                    const mgpu_fab* w0_cart) {
  return make_mac_api(p, g, nfabs, w0, w0mac, w0_cart, 0);
}
int mgpu_make_s0mac(const mgpu_params* p, const mgpu_geom* g, int nfabs, const double* s0, mgpu_fab* const* s0mac,
                    const mgpu_fab* s0_cart) {
  return make_mac_api(p, g, nfabs, s0, s0mac, s0_cart, 1);
}

int mgpu_addw0_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* umac, const mgpu_fab* const* w0mac, double mult) {
  MGPU_TRY
  need_sphr(p, nullptr);
  Call c(p, 0);
  for (int i = 0; i < nfabs; ++i) {
    DV um[3], wm[3];
    c.views((const mgpu_fab* const*)umac, i, true, true, um);
    c.views(w0mac, i, true, false, wm);
    addw0_sphr_dev(um, wm, mult, umac[0][i].lo, umac[0][i].hi);
  }
  c.finish();
  MGPU_CATCH
}

static void fill_sphr_flux(Call& c, const mgpu_params* p, int i, SphrFluxArgs& a, mgpu_fab* const* sflux,
                           const mgpu_fab* const* sedge, const mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                           const mgpu_fab* const* r0o, const mgpu_fab* const* r0n, const mgpu_fab* const* h0o,
                           const mgpu_fab* const* h0n) {
  a.spt = p->species_pred_type;
  a.rho = p->rho_comp - 1;
  a.rhoh = p->rhoh_comp - 1;
  a.vb = grown(sflux[0][i].lo, sflux[0][i].hi, 3, 0);
  c.views((const mgpu_fab* const*)sflux, i, true, true, a.sflux);
  c.views(sedge, i, true, false, a.sedge);
  c.views(umac, i, true, false, a.umac);
  c.views(w0mac, i, true, false, a.w0mac);
  c.views(r0o, i, true, false, a.r0o);
  c.views(r0n, i, true, false, a.r0n);
  for (int d = 0; d < 3; ++d) a.h0o[d] = a.h0n[d] = a.r0o[d];
  if (h0o) c.views(h0o, i, true, false, a.h0o);
  if (h0n) c.views(h0n, i, true, false, a.h0n);
}

int mgpu_mk_rhoX_flux_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                           const mgpu_fab* const* umac, const mgpu_fab* const* w0mac, const mgpu_fab* const* rho0mac_old,
                           const mgpu_fab* const* rho0mac_new, int startcomp, int endcomp) {
  MGPU_TRY
  need_sphr(p, nullptr);
  Call c(p, 0);
  for (int i = 0; i < nfabs; ++i) {
    SphrFluxArgs a;
    fill_sphr_flux(c, p, i, a, sflux, sedge, umac, w0mac, rho0mac_old, rho0mac_new, nullptr, nullptr);
    mk_rhoX_flux_sphr_dev(a, startcomp, endcomp);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mk_rhoh_flux_sphr(const mgpu_params* p, int nfabs, mgpu_fab* const* sflux, const mgpu_fab* const* sedge,
                           const mgpu_fab* const* umac, const mgpu_fab* const* w0mac, const mgpu_fab* const* rho0mac_old,
                           const mgpu_fab* const* rho0mac_new, const mgpu_fab* const* h0mac_old,
                           const mgpu_fab* const* h0mac_new) {
  MGPU_TRY
  need_sphr(p, nullptr);
  Call c(p, 0);
  for (int i = 0; i < nfabs; ++i) {
    SphrFluxArgs a;
    fill_sphr_flux(c, p, i, a, sflux, sedge, umac, w0mac, rho0mac_old, rho0mac_new, h0mac_old, h0mac_new);
    mk_rhoh_flux_sphr_dev(*p, a);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_update_velocity_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* uold, mgpu_fab* unew,
                              const mgpu_fab* const* umac, const mgpu_fab* const* uedge, const mgpu_fab* force,
                              const mgpu_fab* sponge, const mgpu_fab* const* w0mac) {
  MGPU_TRY
  need_sphr(p, nullptr);
  Call c(p, 0);
  for (int i = 0; i < nfabs; ++i) {
    VelArgs a;
    a.dm = 3;
    a.do_sponge = p->do_sponge != 0;
    a.dt = p->dt;
    for (int d = 0; d < 3; ++d) a.dx[d] = p->dx[d];
    a.vb = grown(uold[i].lo, uold[i].hi, 3, 0);
    a.uold = c.view(uold[i], true, false);
    a.unew = c.view(unew[i], true, true);
    a.force = c.view(force[i], true, false);
    a.sponge = c.view(sponge[i], true, false);
    c.views(umac, i, true, false, a.umac);
    c.views(uedge, i, true, false, a.uedge);
    a.w0 = nullptr;
    DV wm[3];
    c.views(w0mac, i, true, false, wm);
    update_velocity_sphr_dev(a, wm);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mkutrans_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                       mgpu_fab* const* utrans, const mgpu_fab* const* w0mac, const int* adv_bc, const int* phys_bc) {
  MGPU_TRY
  need_sphr(p, nullptr);
  if (!p->spherical) throw Error("mkutrans_sphr: params.spherical must be 1");
  Call c(p, 4096);
  for (int i = 0; i < nfabs; ++i) {
    DV ut = c.view(utilde[i], true, false), uf = c.view(ufull[i], true, false);
    DV tr[3], wm[3];
    c.views((const mgpu_fab* const*)utrans, i, true, true, tr);
    c.views(w0mac, i, true, false, wm);
    mkutrans_dev(*p, ut, uf, tr, nullptr, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng, wm);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_velpred_sphr(const mgpu_params* p, int nfabs, const mgpu_fab* utilde, const mgpu_fab* ufull,
                      mgpu_fab* const* umac, const mgpu_fab* const* utrans, const mgpu_fab* force,
                      const mgpu_fab* const* w0mac, const int* adv_bc, const int* phys_bc) {
  MGPU_TRY
  need_sphr(p, nullptr);
  if (!p->spherical) throw Error("velpred_sphr: params.spherical must be 1");
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) need = std::max(need, velpred_scratch(*p, utilde[i].lo, utilde[i].hi));
  Call c(p, need + 4096);
  for (int i = 0; i < nfabs; ++i) {
    size_t mark = arena_mark();
    DV ut = c.view(utilde[i], true, false), uf = c.view(ufull[i], true, false), fv = c.view(force[i], true, false);
    DV um[3], tr[3], wm[3];
    c.views((const mgpu_fab* const*)umac, i, true, true, um);
    c.views(utrans, i, true, false, tr);
    c.views(w0mac, i, true, false, wm);
    velpred_dev(*p, ut, uf, um, tr, fv, nullptr, utilde[i].lo, utilde[i].hi, adv_bc, phys_bc, utilde[i].ng, force[i].ng, wm);
    arena_release(mark);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_modify_scal_force_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* force, const mgpu_fab* s,
                                const mgpu_fab* const* umac, const mgpu_fab* s0_cart, const double* w0, int comp,
                                int fullform) {
  MGPU_TRY
  need_sphr(p, g);
  Call c(p, geom_scratch(g));
  Geom gd = make_geom(*p, *g);
  for (int i = 0; i < nfabs; ++i) {
    DV fv = c.view(force[i], true, true), sv = c.view(s[i], true, false), sc = c.view(s0_cart[i], true, false);
    DV um[3];
    c.views(umac, i, true, false, um);
    modify_scal_force_sphr_dev(*p, *g, gd, fv, sv, um, sc, w0, comp, fullform != 0, s[i].lo, s[i].hi);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_put_in_pert_form_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* s, const double* s0,
                               int comp, int flag) {
  MGPU_TRY
  need_sphr(p, g);
  Call c(p, geom_scratch(g));
  Geom gd = make_geom(*p, *g);
  const double* s0d = upload_small(s0, (size_t)g->nr_fine);
  for (int i = 0; i < nfabs; ++i) {
    DV sv = c.view(s[i], true, true);
    pert_form_sphr_dev(*g, gd, sv, s0d, comp, flag != 0, s[i].lo, s[i].hi);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_density_advance_sphr(const mgpu_params* p, const mgpu_geom* g, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                              mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                              mgpu_fab* const* umac, const double* w0, const mgpu_fab* const* w0mac,
                              const double* rho0_old, const double* rho0_new, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  need_sphr(p, g);
  size_t cell = 1;
  for (int d = 0; d < 3; ++d) cell *= (size_t)(sold->hi[d] - sold->lo[d] + 1 + 6);
  Call c(p, std::max(make_edge_scal_scratch(*p, sold->lo, sold->hi), bds_scratch(*p, sold->lo, sold->hi)) +
                10 * (cell * sizeof(double) + 256) + umac_copy_bytes(sold->lo, sold->hi) + geom_scratch(g) + 8192);
  DV so = c.view(*sold, true, true), sn = c.view(*snew, true, true), fv = c.view(*scal_force, false, true);
  DV se[3], sf[3], um[3], wm[3];
  c.views((const mgpu_fab* const*)sedge, 0, true, true, se);
  c.views((const mgpu_fab* const*)sflux, 0, true, true, sf);
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  c.views(w0mac, 0, true, false, wm);
  density_advance_sphr_dev(*p, *g, which_step, so, sn, se, sf, fv, um, w0, wm, rho0_old, rho0_new, sold->lo, sold->hi,
                           sold->ng, scal_force->ng, adv_bc, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_make_normal(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* normal) {
  MGPU_TRY
  need_sphr(p, g);
  Call c(p, geom_scratch(g));
  Geom gd = make_geom(*p, *g);
  for (int i = 0; i < nfabs; ++i) {
    if (normal[i].nc < 3) throw Error("make_normal: normal needs three components");
    DV nv = c.view(normal[i], false, true);
    make_normal_dev(gd, nv, normal[i].lo, normal[i].hi, normal[i].ng);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mk_vel_force_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* vel_force, int is_final_update,
                           const mgpu_fab* uold, const mgpu_fab* const* uedge, const double* w0,
                           const mgpu_fab* const* w0mac, const mgpu_fab* gpi, const mgpu_fab* s, int index_rho,
                           const mgpu_fab* normal, const double* rho0, const double* grav, const mgpu_fab* w0_force_cart,
                           int do_add_utilde_force) {
  MGPU_TRY
  need_sphr(p, g);
  if (!p->spherical) throw Error("mk_vel_force_sphr: params.spherical must be 1");
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) need = std::max(need, sphr_force_scratch(vel_force[i].lo, vel_force[i].hi));
  Call c(p, need + geom_scratch(g) + (size_t)(4 * (g->nr_fine + 4)) * sizeof(double));
  Geom gd = make_geom(*p, *g);
  for (int i = 0; i < nfabs; ++i) {
    DV fv = c.view(vel_force[i], false, true), uo = c.view(uold[i], true, false), gp = c.view(gpi[i], true, false);
    DV sv = c.view(s[i], crange(index_rho - 1, 1), (cmask_t)0), nm = c.view(normal[i], true, false);
    DV wf = c.view(w0_force_cart[i], true, false);
    DV ue[3], wm[3];
    c.views(uedge, i, true, false, ue);
    c.views(w0mac, i, true, false, wm);
    size_t mark = arena_mark();
    mk_vel_force_sphr_dev(*p, *g, gd, fv, is_final_update != 0, uo, ue, w0, wm, gp, sv.comp(index_rho - 1), nm, rho0, grav, wf,
                          vel_force[i].lo, vel_force[i].hi, do_add_utilde_force != 0);
    arena_release(mark);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_advance_premac_sphr(const mgpu_params* p, const mgpu_geom* g, const mgpu_fab* uold, const mgpu_fab* sold,
                             mgpu_fab* const* umac, const mgpu_fab* gpi, const mgpu_fab* normal, const double* w0,
                             const mgpu_fab* const* w0mac, const mgpu_fab* w0_force_cart, const double* rho0_old,
                             const double* grav_cell_old, const int* adv_bc, const int* phys_bc, const int* pmask) {
  MGPU_TRY
  need_sphr(p, g);
  if (!p->spherical) throw Error("advance_premac_sphr: params.spherical must be 1");
  Call c(p, advance_premac_scratch(*p, uold->lo, uold->hi, uold->ng) + sphr_force_scratch(uold->lo, uold->hi) +
                geom_scratch(g) + (size_t)(6 * (g->nr_fine + 4)) * sizeof(double));
  DV uo = c.view(*uold, true, false), so = c.view(*sold, crange(p->rho_comp - 1, 1), (cmask_t)0);
  DV gp = c.view(*gpi, true, false), nm = c.view(*normal, true, false), wf = c.view(*w0_force_cart, true, false);
  DV um[3], wm[3];
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  c.views(w0mac, 0, true, false, wm);
  advance_premac_sphr_dev(*p, *g, uo, so, um, gp, nm, w0, wm, wf, rho0_old, grav_cell_old, uold->lo, uold->hi, uold->ng,
                          adv_bc, phys_bc, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_velocity_advance_sphr(const mgpu_params* p, const mgpu_geom* g, const mgpu_fab* uold, mgpu_fab* unew,
                               const mgpu_fab* sold, const mgpu_fab* rhohalf, mgpu_fab* const* umac, const mgpu_fab* gpi,
                               const mgpu_fab* normal, const double* w0, const mgpu_fab* const* w0mac,
                               const mgpu_fab* w0_force_cart, const double* rho0_old, const double* rho0_nph,
                               const double* grav_cell_old, const double* grav_cell_nph, const mgpu_fab* sponge,
                               const int* adv_bc, const int* pmask) {
  MGPU_TRY
  need_sphr(p, g);
  if (!p->spherical) throw Error("velocity_advance_sphr: params.spherical must be 1");
  Call c(p, velocity_advance_scratch(*p, uold->lo, uold->hi, uold->ng) + sphr_force_scratch(uold->lo, uold->hi) +
                umac_copy_bytes(uold->lo, uold->hi) + geom_scratch(g) + (size_t)(6 * (g->nr_fine + 4)) * sizeof(double));
  DV uo = c.view(*uold, true, false), un = c.view(*unew, true, true), gp = c.view(*gpi, true, false);
  DV so = c.view(*sold, crange(p->rho_comp - 1, 1), (cmask_t)0), rh = c.view(*rhohalf, true, false);
  DV sp = c.view(*sponge, true, false), nm = c.view(*normal, true, false), wf = c.view(*w0_force_cart, true, false);
  DV um[3], wm[3];
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  c.views(w0mac, 0, true, false, wm);
  velocity_advance_sphr_dev(*p, *g, uo, un, so, rh, um, gp, nm, w0, wm, wf, rho0_old, rho0_nph, grav_cell_old,
                            grav_cell_nph, sp, uold->lo, uold->hi, uold->ng, adv_bc, pmask);
  c.finish();
  MGPU_CATCH
}

int mgpu_mkrhohforce_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* scal_force, int is_prediction,
                          const mgpu_fab* thermal, const mgpu_fab* const* umac, const double* p0_1, const double* p0_2,
                          const double* psi, int add_thermal, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  need_sphr(p, g);
  if (!p->spherical) throw Error("mkrhohforce_sphr: params.spherical must be 1");
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i)
    need = std::max(need, 6 * fab_bytes(scal_force[i].lo, scal_force[i].hi, 3, 2, 1, 1));
  Call c(p, need + geom_scratch(g) + (size_t)(4 * (g->nr_fine + 4)) * sizeof(double));
  const cmask_t mrhoh = crange(p->rhoh_comp - 1, 1);
  for (int i = 0; i < nfabs; ++i) {
    DV fv = c.view(scal_force[i], mrhoh, mrhoh), th = c.view(thermal[i], true, false);
    DV um[3];
    c.views(umac, i, true, false, um);
    SphrCtx X{*p, *g, make_geom(*p, *g), scal_force[i].lo, scal_force[i].hi, adv_bc, pmask};
    rhoh_force_sphr_full(X, fv, is_prediction != 0, th, um, p0_1, p0_2, psi, add_thermal != 0, scal_force[i].ng);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_enthalpy_advance_sphr(const mgpu_params* p, const mgpu_geom* g, int which_step, mgpu_fab* sold, mgpu_fab* snew,
                               mgpu_fab* const* sedge, mgpu_fab* const* sflux, mgpu_fab* scal_force,
                               const mgpu_fab* thermal, mgpu_fab* const* umac, const double* w0,
                               const mgpu_fab* const* w0mac, const double* rho0_old, const double* rhoh0_old,
                               const double* rho0_new, const double* rhoh0_new, const double* p0_old, const double* p0_new,
                               const double* tempbar, const double* psi, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  need_sphr(p, g);
  if (!p->spherical) throw Error("enthalpy_advance_sphr: params.spherical must be 1");
  Call c(p, std::max(make_edge_scal_scratch(*p, sold->lo, sold->hi), bds_scratch(*p, sold->lo, sold->hi)) +
                14 * fab_bytes(sold->lo, sold->hi, 3, 2, 1, 1) + umac_copy_bytes(sold->lo, sold->hi) + geom_scratch(g) +
                (size_t)(16 * (g->nr_fine + 4)) * sizeof(double) + 16384);
  const EnthalpyMasks m = enthalpy_masks(*p);
  DV so = c.view(*sold, m.sold_in, m.sold_out), sn = c.view(*snew, m.snew_in, m.mrhoh);
  DV fv = c.view(*scal_force, (cmask_t)0, m.force_out);
  c.zero_on_host(*scal_force, ~m.force_out);
  DV th = c.view(*thermal, true, false);
  DV se[3], sf[3], um[3], wm[3];
  // edge-state fabs with ghost cells: the episode writes the valid faces only, so what it writes is read first
  c.views((const mgpu_fab* const*)sedge, 0, m.sedge_in | (sedge[0]->ng > 0 ? m.sedge_out : (cmask_t)0), m.sedge_out, se);
  c.views((const mgpu_fab* const*)sflux, 0, (cmask_t)0, m.mrhoh, sf);
  c.views((const mgpu_fab* const*)umac, 0, true, true, um);
  c.views(w0mac, 0, true, false, wm);
  enthalpy_advance_sphr_dev(*p, *g, which_step, so, sn, se, sf, fv, th, um, w0, wm, rho0_old, rhoh0_old, rho0_new, rhoh0_new,
                            p0_old, p0_new, tempbar, psi, sold->lo, sold->hi, sold->ng, scal_force->ng, adv_bc, pmask);
  c.finish();
  MGPU_CATCH
}

// ---- estdt and make_etarho_planar ----------------------------------------------------------
int mgpu_estdt(const mgpu_params* p, int nfabs, const mgpu_fab* u, const mgpu_fab* s, const mgpu_fab* force,
               const mgpu_fab* divU, const mgpu_fab* dSdt, const double* w0, const double* p0,
               const double* gamma1bar, double rho_min, double cflfac, double* dt, double* umax) {
  MGPU_TRY
  if (p->spherical) throw Error("mgpu_estdt: spherical == 1 is served by mgpu_estdt_sphr");
  const int nr = p->nr, dm = p->dm;
  Call c(p, (size_t)(4 * (nr + 2) + 148 * 8 * 8 + 64) * sizeof(double) + 8192);
  const double* w0_d = upload_small(w0, nr + 1);
  const double* p0_d = upload_small(p0, nr);
  const double* g1_d = upload_small(gamma1bar, nr);
  const double dt_start = 1.e99;  // estdt.f90:144-146
  double dt_proc = 1.e99, umax_proc = 0.0;
  for (int i = 0; i < nfabs; ++i) {
    DV uv = c.view(u[i], true, false);
    DV sv = c.view(s[i], crange(p->rho_comp - 1, 1), (cmask_t)0);
    DV fv = c.view(force[i], true, false);
    DV dUv = c.view(divU[i], true, false), dSv = c.view(dSdt[i], true, false);
    double dt_grid = std::numeric_limits<double>::max(), umax_grid = 0.0;  // HUGE(dt_grid), :158-159
    estdt_box_dev(*p, uv, sv, fv, dUv, dSv, w0_d, w0, p0_d, g1_d, u[i].lo, u[i].hi, rho_min, cflfac, &dt_grid,
                  &umax_grid);
    dt_proc = std::min(dt_proc, dt_grid);
    umax_proc = std::max(umax_proc, umax_grid);
  }
  double dt_lev = dt_proc, umax_lev = umax_proc;
  if (comm_size() > 1) {  // parallel_reduce MPI_MIN / MPI_MAX, :202-203: one MIN over (dt, -umax)
    double h[2] = {dt_proc, -umax_proc};
    double* d = arena_alloc(2);
    MGPU_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, g_ctx.stream));
    allreduce_dev(d, 2, 1);
    MGPU_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
    dt_lev = h[0];
    umax_lev = -h[1];
  }
  *umax = std::max(*umax, umax_lev);  // :206
  if (dt_lev == dt_start) {           // protect against huge time steps, :209-217
    dt_lev = p->dx[0];
    for (int d = 1; d < dm; ++d) dt_lev = std::min(dt_lev, p->dx[d]);
  }
  *dt = std::min(*dt, dt_lev);  // :220
  c.finish();
  MGPU_CATCH
}

int mgpu_minmax(const mgpu_params* p, int nfabs, const mgpu_fab* s, int comp, int div_comp, double* smin, double* smax) {
  MGPU_TRY
  if (nfabs < 1) throw Error("mgpu_minmax: needs at least one box");
  Call c(p, (size_t)(2 * 148 * 8 + 64) * sizeof(double) * (size_t)nfabs + 8192);
  double mn = std::numeric_limits<double>::max(), mx = -std::numeric_limits<double>::max();  // fab.f90: Huge(r) / -Huge(r)
  for (int i = 0; i < nfabs; ++i) {
    if (comp < 1 || comp > s[i].nc || div_comp > s[i].nc) throw Error("mgpu_minmax: component out of range");
    cmask_t in = crange(comp - 1, 1);
    if (div_comp >= 1) in |= crange(div_comp - 1, 1);
    DV sv = c.view(s[i], in, (cmask_t)0);
    minmax_box_dev(*p, sv, s[i].lo, s[i].hi, comp - 1, div_comp >= 1 ? div_comp - 1 : -1, &mn, &mx);
  }
  if (comm_size() > 1) {  // parallel_reduce MPI_MIN / MPI_MAX (multifab_f.f90 multifab_min_c / _max_c): one MIN over (min, -max)
    double h[2] = {mn, -mx};
    double* d = arena_alloc(2);
    MGPU_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, g_ctx.stream));
    allreduce_dev(d, 2, 1);
    MGPU_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
    mn = h[0];
    mx = -h[1];
  }
  *smin = mn;
  *smax = mx;
  c.finish();
  MGPU_CATCH
}

int mgpu_make_etarho_planar(const mgpu_params* p, int nfabs, const mgpu_fab* etarhoflux, double* etarho_ec,
                            double* etarho_cc) {
  MGPU_TRY
  if (p->spherical) throw Error("ERROR: make_eta should not be called for spherical");  // make_eta.f90:73-75
  const int nr = p->nr, dm = p->dm, r = dm - 1;
  Call c(p, (size_t)(4 * (nr + 2)) * sizeof(double) + 8192);
  std::vector<double> sum(nr + 1, 0.0);
  for (int i = 0; i < nfabs; ++i) {
    const mgpu_fab& f = etarhoflux[i];
    DV ev = c.view(f, true, false);
    // the top edge only where the box touches the top of the domain (no double counting), :230-247
    const int k1 = f.hi[r] + (f.hi[r] == nr - 1 ? 1 : 0);
    if (f.lo[r] < 0 || k1 > nr) throw Error("make_etarho_planar: box outside the base-state range 0:nr");
    std::vector<double> part(k1 - f.lo[r] + 1);
    plane_sums_dev(*p, ev.comp(0), f.lo, f.hi, f.lo[r], k1, part.data());
    for (int k = f.lo[r]; k <= k1; ++k) sum[k] = sum[k] + part[k - f.lo[r]];
  }
  if (comm_size() > 1) {  // parallel_reduce MPI_SUM, :101
    double* d = arena_alloc((size_t)nr + 1);
    MGPU_CUDA(cudaMemcpyAsync(d, sum.data(), (nr + 1) * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
    allreduce_dev(d, nr + 1, 0);
    MGPU_CUDA(cudaMemcpyAsync(sum.data(), d, (nr + 1) * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  }
  double ncell = 1.0;  // :80-86
  for (int d = 0; d < r; ++d) ncell *= (double)(p->domhi[d] - p->domlo[d] + 1);
  for (int k = 0; k <= nr; ++k) etarho_ec[k] = sum[k] / ncell;                            // :103-107
  for (int k = 0; k < nr; ++k) etarho_cc[k] = 0.5 * (etarho_ec[k] + etarho_ec[k + 1]);  // :117-123
  c.finish();
  MGPU_CATCH
}


int mgpu_estdt_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* u, const mgpu_fab* s,
                    const mgpu_fab* force, const mgpu_fab* divU, const mgpu_fab* dSdt, const mgpu_fab* const* w0mac,
                    const double* w0, const double* p0, const double* gamma1bar, double rho_min, double cflfac,
                    double* dt, double* umax) {
  MGPU_TRY
  need_sphr(p, g);
  const int nr = g->nr_fine;
  size_t cells = 0;
  for (int i = 0; i < nfabs; ++i) {
    size_t c1 = 1;
    for (int d = 0; d < 3; ++d) c1 *= (size_t)(u[i].hi[d] - u[i].lo[d] + 1);
    cells = std::max(cells, c1);
  }
  Call c(p, geom_scratch(g) + (3 * cells + 148 * 8 * 8 + 64) * sizeof(double) + 8192);
  Geom gd = make_geom(*p, *g);
  // gp0 on the radial edges (estdt.f90:734-739)
  std::vector<double> gp0(nr + 1);
  for (int r = 1; r <= nr - 1; ++r) {
    const double gamma1bar_p_avg = 0.5 * (gamma1bar[r] * p0[r] + gamma1bar[r - 1] * p0[r - 1]);
    gp0[r] = ((p0[r] - p0[r - 1]) / g->dr) / gamma1bar_p_avg;
  }
  gp0[nr] = gp0[nr - 1];
  gp0[0] = gp0[1];
  const double* gp0_d = upload_small(gp0.data(), (size_t)nr + 1);
  const double* w0_d = upload_small(w0, (size_t)nr + 1);
  const double dt_start = 1.e99;
  double dt_proc = 1.e99, umax_proc = 0.0;
  for (int i = 0; i < nfabs; ++i) {
    DV uv = c.view(u[i], true, false);
    DV sv = c.view(s[i], crange(p->rho_comp - 1, 1), (cmask_t)0);
    DV fv = c.view(force[i], true, false);
    DV dUv = c.view(divU[i], true, false), dSv = c.view(dSdt[i], true, false);
    DV wm[3];
    c.views(w0mac, i, true, false, wm);
    const size_t mark = arena_mark();
    size_t c1 = 1;
    for (int d = 0; d < 3; ++d) c1 *= (size_t)(u[i].hi[d] - u[i].lo[d] + 1);
    DV gc = make_view(arena_alloc(3 * c1), u[i].lo, u[i].hi, 3, 0, nullptr, 3);
    put_1d_array_on_cart_dev(*p, *g, gd, gp0_d, gc, true, true, u[i].lo, u[i].hi);  // :741
    double dt_grid = std::numeric_limits<double>::max(), umax_grid = 0.0;
    estdt_box_dev(*p, uv, sv, fv, dUv, dSv, w0_d, w0, nullptr, nullptr, u[i].lo, u[i].hi, rho_min, cflfac, &dt_grid,
                  &umax_grid, wm, &gc, g->dr, nr);
    arena_release(mark);
    dt_proc = std::min(dt_proc, dt_grid);
    umax_proc = std::max(umax_proc, umax_grid);
  }
  double dt_lev = dt_proc, umax_lev = umax_proc;
  if (comm_size() > 1) {
    double h[2] = {dt_proc, -umax_proc};
    double* d = arena_alloc(2);
    MGPU_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, g_ctx.stream));
    allreduce_dev(d, 2, 1);
    MGPU_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
    dt_lev = h[0];
    umax_lev = -h[1];
  }
  *umax = std::max(*umax, umax_lev);
  if (dt_lev == dt_start) dt_lev = std::min(p->dx[0], std::min(p->dx[1], p->dx[2]));
  *dt = std::min(*dt, dt_lev);
  c.finish();
  MGPU_CATCH
}

// ---- the EOS and the pieces of the path that call it (mgpu_eos.cu) ---------------------------
int mgpu_set_eos(const mgpu_eos* e) {
  MGPU_TRY
  set_eos(e);
  MGPU_CATCH
}

int mgpu_eos_eval(int input, long n, double* state, const double* xn) {
  MGPU_TRY
  require_init();
  const EosDev& E = the_eos("eos");
  if (n < 0) throw Error("mgpu_eos_eval: n < 0");
  arena_reserve((size_t)n * (MGPU_EOS_NQ + E.nspec) * sizeof(double) + (1u << 20));
  arena_reset();
  double* sd = arena_alloc((size_t)n * MGPU_EOS_NQ);
  double* xd = arena_alloc((size_t)n * E.nspec);
  MGPU_CUDA(cudaMemcpyAsync(sd, state, (size_t)n * MGPU_EOS_NQ * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
  MGPU_CUDA(cudaMemcpyAsync(xd, xn, (size_t)n * E.nspec * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
  eos_points_dev(input, n, sd, xd);
  MGPU_CUDA(cudaMemcpyAsync(state, sd, (size_t)n * MGPU_EOS_NQ * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  MGPU_CATCH
}

static void h_edge_common(const mgpu_params* p, HEdgeArgs& a, const mgpu_fab& f0) {
  a.dm = p->dm;
  a.ept = p->enthalpy_pred_type;
  a.spt = p->species_pred_type;
  a.rho = p->rho_comp - 1; a.rhoh = p->rhoh_comp - 1; a.temp = p->temp_comp - 1; a.spec0 = p->spec_comp - 1;
  a.vb = grown(f0.lo, f0.hi, p->dm, 0);
  if (!(a.ept == MGPU_PREDICT_T_THEN_RHOHPRIME || a.ept == MGPU_PREDICT_T_THEN_H || a.ept == MGPU_PREDICT_TPRIME_THEN_H))
    throw Error("makeHfromRhoT_edge: enthalpy_pred_type must be one of the temperature-based predictions");
}
static cmask_t h_edge_in(const mgpu_params* p) {
  return crange(p->rho_comp - 1, 1) | crange(p->temp_comp - 1, 1) | crange(p->spec_comp - 1, p->nspec);
}

int mgpu_make_h_from_rhot_edge(const mgpu_params* p, int nfabs, mgpu_fab* const* sedge, const double* rho0_old,
                               const double* rhoh0_old, const double* t0_old, const double* rho0_edge_old,
                               const double* rhoh0_edge_old, const double* t0_edge_old, const double* rho0_new,
                               const double* rhoh0_new, const double* t0_new, const double* rho0_edge_new,
                               const double* rhoh0_edge_new, const double* t0_edge_new) {
  MGPU_TRY
  if (p->spherical) throw Error("make_h_from_rhot_edge: use the _sphr entry point with spherical == 1");
  const int nr = p->nr;
  Call c(p, (size_t)(12 * (nr + 4)) * sizeof(double) + 16384);
  HEdgeArgs a{};
  a.sphr = false;
  a.rho0_old = upload_small(rho0_old, (size_t)nr); a.rhoh0_old = upload_small(rhoh0_old, (size_t)nr);
  a.t0_old = upload_small(t0_old, (size_t)nr);
  a.rho0_new = upload_small(rho0_new, (size_t)nr); a.rhoh0_new = upload_small(rhoh0_new, (size_t)nr);
  a.t0_new = upload_small(t0_new, (size_t)nr);
  a.rho0_edge_old = upload_small(rho0_edge_old, (size_t)nr + 1); a.rhoh0_edge_old = upload_small(rhoh0_edge_old, (size_t)nr + 1);
  a.t0_edge_old = upload_small(t0_edge_old, (size_t)nr + 1);
  a.rho0_edge_new = upload_small(rho0_edge_new, (size_t)nr + 1); a.rhoh0_edge_new = upload_small(rhoh0_edge_new, (size_t)nr + 1);
  a.t0_edge_new = upload_small(t0_edge_new, (size_t)nr + 1);
  for (int i = 0; i < nfabs; ++i) {
    h_edge_common(p, a, sedge[0][i]);
    c.views((const mgpu_fab* const*)sedge, i, h_edge_in(p) | (sedge[0][i].ng > 0 ? crange(p->rhoh_comp - 1, 1) : (cmask_t)0),
            crange(p->rhoh_comp - 1, 1), a.sedge);
    a.rho0_cart = a.rhoh0_cart = a.t0_cart = a.sedge[0];
    h_from_rhot_edge_dev(a);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_make_h_from_rhot_edge_sphr(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* const* sedge,
                                    const double* rho0_old, const double* rhoh0_old, const double* t0_old,
                                    const double* rho0_new, const double* rhoh0_new, const double* t0_new,
                                    const int* adv_bc, const int* pmask) {
  MGPU_TRY
  need_sphr(p, g);
  if (!p->spherical) throw Error("make_h_from_rhot_edge_sphr: params.spherical must be 1");
  const int nr = g->nr_fine, dm = 3;
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) need = std::max(need, 3 * fab_bytes(sedge[0][i].lo, sedge[0][i].hi, 3, 2, 0, 1));
  Call c(p, need + geom_scratch(g) + (size_t)(4 * (nr + 4)) * sizeof(double) + 16384);
  std::vector<double> r0h(nr), rh0h(nr), t0h(nr);
  for (int r = 0; r < nr; ++r) {  // rhoh_vs_t.f90:91-95
    r0h[r] = 0.5 * (rho0_old[r] + rho0_new[r]);
    rh0h[r] = 0.5 * (rhoh0_old[r] + rhoh0_new[r]);
    t0h[r] = 0.5 * (t0_old[r] + t0_new[r]);
  }
  const double* r0d = upload_small(r0h.data(), (size_t)nr);
  const double* rh0d = upload_small(rh0h.data(), (size_t)nr);
  const double* t0d = upload_small(t0h.data(), (size_t)nr);
  for (int i = 0; i < nfabs; ++i) {
    HEdgeArgs a{};
    a.sphr = true;
    h_edge_common(p, a, sedge[0][i]);
    c.views((const mgpu_fab* const*)sedge, i, h_edge_in(p) | (sedge[0][i].ng > 0 ? crange(p->rhoh_comp - 1, 1) : (cmask_t)0),
            crange(p->rhoh_comp - 1, 1), a.sedge);
    SphrCtx X{*p, *g, make_geom(*p, *g), sedge[0][i].lo, sedge[0][i].hi, adv_bc, pmask};
    const size_t mark = arena_mark();
    a.rho0_cart = sphr_cart(X, r0d, 2, dm + p->rho_comp);
    a.rhoh0_cart = sphr_cart(X, rh0d, 2, dm + p->rhoh_comp);
    a.t0_cart = sphr_cart(X, t0d, 2, dm + p->temp_comp);
    h_from_rhot_edge_dev(a);
    arena_release(mark);
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_mktempforce(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* temp_force,
                     const mgpu_fab* const* umac, const mgpu_fab* s, const mgpu_fab* thermal, const double* p0_old,
                     const double* psi, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  if (p->spherical) need_sphr(p, g);
  const int ept = p->enthalpy_pred_type;
  if (!(ept == MGPU_PREDICT_T_THEN_RHOHPRIME || ept == MGPU_PREDICT_T_THEN_H || ept == MGPU_PREDICT_TPRIME_THEN_H))
    throw Error("ERROR: should only call mkrhohforce when predicting T or T'");  // mkscalforce.f90:756 (the reference's wording)
  const int dm = p->dm, foextrap_comp = dm + p->nscal + 2;
  const int nr = p->spherical ? g->nr_fine : p->nr;
  size_t need = 0;
  if (p->spherical)
    for (int i = 0; i < nfabs; ++i) need = std::max(need, 2 * fab_bytes(s[i].lo, s[i].hi, 3, 1, 0, 1));
  Call c(p, need + (p->spherical ? geom_scratch(g) : 0) + (size_t)(4 * (nr + 4)) * sizeof(double) + 16384);
  const double* p0d = upload_small(p0_old, (size_t)nr);
  const double* psid = upload_small(psi, (size_t)nr);
  const cmask_t mtemp = crange(p->temp_comp - 1, 1);
  const cmask_t sin = crange(p->rho_comp - 1, 1) | mtemp | crange(p->spec_comp - 1, p->nspec);
  for (int i = 0; i < nfabs; ++i) {
    const int* lo = s[i].lo;
    const int* hi = s[i].hi;
    DV fv = c.view(temp_force[i], mtemp, mtemp), sv = c.view(s[i], sin, (cmask_t)0), th = c.view(thermal[i], true, false);
    TempForceArgs a;
    a.dm = dm; a.nr = nr; a.rho = p->rho_comp - 1; a.temp = p->temp_comp - 1; a.spec0 = p->spec_comp - 1;
    a.sphr = p->spherical != 0;
    a.dr = p->spherical ? g->dr : p->dx[dm - 1];
    for (int d = 0; d < 3; ++d) a.dx[d] = p->dx[d];
    a.vb = grown(lo, hi, dm, 0);
    a.f = fv.comp(a.temp); a.s = sv; a.thermal = th;
    c.views(umac, i, true, false, a.umac);
    a.p0_old = p0d; a.psi = psid;
    a.p0_cart = a.psi_cart = sv;
    const size_t mark = arena_mark();
    if (p->spherical) {  // mkscalforce.f90:770-776, :1059-1061
      SphrCtx X{*p, *g, make_geom(*p, *g), lo, hi, adv_bc, pmask};
      a.p0_cart = sphr_cart(X, p0d, 1, foextrap_comp);
      a.psi_cart = arena_fab(lo, hi, 3, 0, nullptr, 1);
      put_1d_array_on_cart_dev(*p, *g, X.gd, psid, a.psi_cart, false, false, lo, hi);
    }
    mktempforce_dev(a);
    arena_release(mark);
    fill_boundary_dev(*p, fv, lo, hi, temp_force[i].ng, nullptr, p->temp_comp, foextrap_comp, 1, adv_bc, pmask, false);  // :833-837
  }
  c.finish();
  MGPU_CATCH
}

int mgpu_firstdt(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* u, const mgpu_fab* gpi,
                 const mgpu_fab* s, const mgpu_fab* divU, const double* rho0, const double* p0, const double* grav,
                 const double* gamma1bar, double cflfac, double init_shrink, int use_soundspeed_firstdt,
                 int use_divu_firstdt, double* dt, double* umax) {
  MGPU_TRY
  if (p->spherical) need_sphr(p, g);
  the_eos("firstdt");
  const int dm = p->dm;
  const bool sphr = p->spherical != 0;
  const int nr = sphr ? g->nr_fine : p->nr;
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i)
    need = std::max(need, (size_t)(2 * dm + 2 * dm + 3) * fab_bytes(u[i].lo, u[i].hi, dm, 1, 1, 1) +
                              (sphr ? 8 * fab_bytes(u[i].lo, u[i].hi, 3, 1, 0, 3) : 0));
  Call c(p, need + (sphr ? geom_scratch(g) : 0) + (size_t)(8 * (nr + 4) + 148 * 8 * 8 + 64) * sizeof(double) + 16384);
  std::vector<double> zeros((size_t)nr + 1, 0.0);
  const double* w0_dummy = upload_small(zeros.data(), (size_t)nr + 1);        // firstdt.f90:74-77
  const double* w0_force_dummy = upload_small(zeros.data(), (size_t)nr);
  const double* rho0_d = upload_small(rho0, (size_t)nr);
  const double* grav_d = upload_small(grav, (size_t)nr);
  const double* p0_d = upload_small(p0, (size_t)nr);
  const double* g1_d = upload_small(gamma1bar, (size_t)nr);
  const double* gp0_d = nullptr;
  if (sphr) {  // gp0 on the radial edges (:708-719)
    std::vector<double> gp0(nr + 1);
    for (int r = 1; r <= nr - 1; ++r) {
      const double gamma1bar_p_avg = 0.5 * (gamma1bar[r] * p0[r] + gamma1bar[r - 1] * p0[r - 1]);
      gp0[r] = ((p0[r] - p0[r - 1]) / g->dr) / gamma1bar_p_avg;
    }
    gp0[nr] = gp0[nr - 1];
    gp0[0] = gp0[1];
    gp0_d = upload_small(gp0.data(), (size_t)nr + 1);
  }
  // the force of firstdt.f90:96-100 needs no boundary data: adv_bc / pmask of an all-interior box keep the ghost fill
  // of mk_vel_force away from anything the reductions read (they run over the valid cells only)
  double dt_proc = 1.e99, umax_proc = 0.0;  // :134-137
  const cmask_t sin = crange(p->rho_comp - 1, 1) | crange(p->temp_comp - 1, 1) | crange(p->spec_comp - 1, p->nspec);
  for (int i = 0; i < nfabs; ++i) {
    const int* lo = u[i].lo;
    const int* hi = u[i].hi;
    DV uv = c.view(u[i], true, false), gp = c.view(gpi[i], true, false), sv = c.view(s[i], sin, (cmask_t)0);
    DV dUv = c.view(divU[i], true, false);
    const size_t mark = arena_mark();
    DV force = arena_fab(lo, hi, dm, 1, nullptr, dm);
    DV umd[3];
    for (int d = 0; d < dm; ++d) {  // umac_dummy: zero, one ghost layer (:84-88)
      umd[d] = arena_fab(lo, hi, dm, 1, NODAL_D[d], 1);
      set_dev(umd[d].p, 0.0, umd[d].size());
    }
    double dt_grid, umax_grid;
    if (sphr) {
      Geom gd = make_geom(*p, *g);
      DV w0md[3];
      for (int d = 0; d < 3; ++d) {
        w0md[d] = arena_fab(lo, hi, 3, 1, NODAL_D[d], 1);
        set_dev(w0md[d].p, 0.0, w0md[d].size());
      }
      DV normal_dummy = arena_fab(lo, hi, 3, 1, nullptr, 3), w0fc = arena_fab(lo, hi, 3, 1, nullptr, 3);
      set_dev(normal_dummy.p, 0.0, normal_dummy.size());
      set_dev(w0fc.p, 0.0, w0fc.size());
      mk_vel_force_sphr_dev(*p, *g, gd, force, false, uv, umd, zeros.data(), w0md, gp, sv.comp(p->rho_comp - 1),
                            normal_dummy, rho0, grav, w0fc, lo, hi, false);
      DV gc = arena_fab(lo, hi, 3, 0, nullptr, 3);
      put_1d_array_on_cart_dev(*p, *g, gd, gp0_d, gc, true, true, lo, hi);  // :721
      firstdt_box_dev(*p, uv, sv, force, dUv, p0_d, g1_d, &gc, lo, hi, cflfac, use_soundspeed_firstdt != 0,
                      use_divu_firstdt != 0, &dt_grid, &umax_grid);
    } else {
      VelForceArgs a;
      a.dm = dm;
      a.nr = nr;
      a.is_final_update = false;
      a.add_utilde = false;
      a.dr = p->dx[dm - 1];
      a.rho_cut = p->buoyancy_cutoff_factor * p->base_cutoff_density;
      a.omega = p->omega; a.sin_theta = p->sin_theta; a.cos_theta = p->cos_theta; a.rotation_radius = p->rotation_radius;
      a.vb = grown(lo, hi, dm, 0);
      a.force = force; a.uold = uv; a.gpi = gp; a.rho = sv.comp(p->rho_comp - 1);
      for (int d = 0; d < dm; ++d) a.uedge[d] = umd[d];
      a.w0 = w0_dummy; a.rho0 = rho0_d; a.grav = grav_d; a.w0_force = w0_force_dummy;
      mk_vel_force_dev(a);
      firstdt_box_dev(*p, uv, sv, force, dUv, p0_d, g1_d, nullptr, lo, hi, cflfac, use_soundspeed_firstdt != 0,
                      use_divu_firstdt != 0, &dt_grid, &umax_grid);
    }
    arena_release(mark);
    dt_proc = std::min(dt_proc, dt_grid);
    umax_proc = std::max(umax_proc, umax_grid);
  }
  double dt_lev = dt_proc, umax_lev = umax_proc;
  if (comm_size() > 1) {  // parallel_reduce MPI_MIN / MPI_MAX, :169-170: one MIN over (dt, -umax)
    double h[2] = {dt_proc, -umax_proc};
    double* d = arena_alloc(2);
    MGPU_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, g_ctx.stream));
    allreduce_dev(d, 2, 1);
    MGPU_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
    dt_lev = h[0];
    umax_lev = -h[1];
  }
  *umax = std::max(*umax, umax_lev);  // :173
  dt_lev = dt_lev * init_shrink;      // :180
  *dt = std::min(*dt, dt_lev);        // :187
  c.finish();
  MGPU_CATCH
}

static void make_t_api(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0, bool from_p,
                       bool flag, bool update_rhoh, const int* adv_bc, const int* pmask) {
  if (p->spherical) need_sphr(p, g);
  the_eos(from_p ? "makeTfromRhoP" : "makeTfromRhoH");
  const int dm = p->dm;
  const bool sphr = p->spherical != 0;
  const int nr = sphr ? g->nr_fine : p->nr;
  size_t need = 0;
  if (sphr)
    for (int i = 0; i < nfabs; ++i) need = std::max(need, fab_bytes(state[i].lo, state[i].hi, 3, 0, 0, 1));
  Call c(p, need + (sphr ? geom_scratch(g) : 0) + (size_t)(2 * (nr + 4)) * sizeof(double) + 16384);
  const double* p0d = upload_small(p0, (size_t)nr);
  const cmask_t mtemp = crange(p->temp_comp - 1, 1), mrhoh = crange(p->rhoh_comp - 1, 1);
  cmask_t in = crange(p->rho_comp - 1, 1) | mtemp | crange(p->spec_comp - 1, p->nspec);
  if (!from_p) in |= mrhoh;
  if (from_p && flag) in |= crange(p->pi_comp - 1, 1);
  const cmask_t out = mtemp | ((from_p && update_rhoh) ? mrhoh : 0);
  for (int i = 0; i < nfabs; ++i) {
    const int* lo = state[i].lo;
    const int* hi = state[i].hi;
    DV sv = c.view(state[i], in | out, out);
    const size_t mark = arena_mark();
    DV p0c;
    if (sphr) {  // rhoh_vs_t.f90:1097, :1410
      Geom gd = make_geom(*p, *g);
      p0c = arena_fab(lo, hi, 3, 0, nullptr, 1);
      put_1d_array_on_cart_dev(*p, *g, gd, p0d, p0c, false, false, lo, hi);
    }
    make_t_dev(*p, sv, p0d, sphr ? &p0c : nullptr, from_p, flag, update_rhoh, grown(lo, hi, dm, 0));
    arena_release(mark);
    fill_boundary_dev(*p, sv, lo, hi, state[i].ng, nullptr, p->temp_comp, dm + p->temp_comp, 1, adv_bc, pmask, false);
    if (from_p && update_rhoh)
      fill_boundary_dev(*p, sv, lo, hi, state[i].ng, nullptr, p->rhoh_comp, dm + p->rhoh_comp, 1, adv_bc, pmask, false);
  }
  c.finish();
}

int mgpu_make_t_from_rhoh(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0,
                          int use_eos_e_instead_of_h, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  make_t_api(p, g, nfabs, state, p0, false, use_eos_e_instead_of_h != 0, false, adv_bc, pmask);
  MGPU_CATCH
}

int mgpu_make_t_from_rhop(const mgpu_params* p, const mgpu_geom* g, int nfabs, mgpu_fab* state, const double* p0,
                          int update_rhoh, int use_pprime_in_tfromp, const int* adv_bc, const int* pmask) {
  MGPU_TRY
  make_t_api(p, g, nfabs, state, p0, true, use_pprime_in_tfromp != 0, update_rhoh != 0, adv_bc, pmask);
  MGPU_CATCH
}

}  // extern "C"

// ---- average and make_etarho_spherical ------------------------------------------------------------------
// average (average.f90:24) of component incomp of one level's fabs (device views) into phibar(0:nr-1)
static void average_views(const mgpu_params* p, const mgpu_geom* g, int nfabs, const DV* phi1, const mgpu_fab* boxes,
                          int nr_irreg, int drdxfac, double* phibar) {
  const int dm = p->dm;
  if (!p->spherical) {  // :114-163: plane sums over the ranks / cells of the domain's plane
    const int nr = p->nr, r = dm - 1;
    std::vector<double> sum(nr, 0.0);
    for (int i = 0; i < nfabs; ++i) {
      const mgpu_fab& f = boxes[i];
      if (f.lo[r] < 0 || f.hi[r] > nr - 1) throw Error("average: box outside the base-state range 0:nr-1");
      std::vector<double> part(f.hi[r] - f.lo[r] + 1);
      plane_sums_dev(*p, phi1[i], f.lo, f.hi, f.lo[r], f.hi[r], part.data());
      for (int k = f.lo[r]; k <= f.hi[r]; ++k) sum[k] = sum[k] + part[k - f.lo[r]];
    }
    if (comm_size() > 1) {  // parallel_reduce MPI_SUM, :152
      double* d = arena_alloc((size_t)nr);
      MGPU_CUDA(cudaMemcpyAsync(d, sum.data(), nr * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
      allreduce_dev(d, nr, 0);
      MGPU_CUDA(cudaMemcpyAsync(sum.data(), d, nr * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
      MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
    }
    double ncell = 1.0;  // :127-135
    for (int d = 0; d < r; ++d) ncell *= (double)(p->domhi[d] - p->domlo[d] + 1);
    for (int k = 0; k < nr; ++k) phibar[k] = sum[k] / ncell;  // :155-159
    return;
  }
  if (nr_irreg < 2) throw Error("average: nr_irreg (geometry) must be passed for spherical geometry");
  for (int i = 0; i < nfabs; ++i) {  // the reference would write past phisum(nr_irreg) for such a box: refuse instead
    double far2 = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double a = g->prob_lo[d] + ((double)boxes[i].lo[d] + 0.5) * p->dx[d] - g->center[d];
      const double b = g->prob_lo[d] + ((double)boxes[i].hi[d] + 0.5) * p->dx[d] - g->center[d];
      far2 += std::max(a * a, b * b);
    }
    if ((int)((far2 / (p->dx[0] * p->dx[0]) - 0.75) / 2.0) > nr_irreg) throw Error("average: a cell maps beyond nr_irreg");
  }
  const int nb = nr_irreg + 1;
  std::vector<double> radii(nr_irreg + 3), phisum(nr_irreg + 2, 0.0);
  std::vector<long> ncell(nr_irreg + 2, 0);
  for (int r = 0; r <= nr_irreg; ++r) radii[r + 1] = std::sqrt(0.75 + 2.0 * r) * p->dx[0];  // :92
  radii[nr_irreg + 2] = 1.e99;                                                               // :98
  const double* radii_d = upload_small(radii.data() + 1, (size_t)nr_irreg + 2);
  double* ps_d = arena_alloc((size_t)2 * nb);  // sums, then counts (as doubles for the reduction over the ranks)
  unsigned long long* nc_d = reinterpret_cast<unsigned long long*>(arena_alloc((size_t)nb));
  MGPU_CUDA(cudaMemsetAsync(ps_d, 0, (size_t)2 * nb * sizeof(double), g_ctx.stream));
  MGPU_CUDA(cudaMemsetAsync(nc_d, 0, (size_t)nb * sizeof(unsigned long long), g_ctx.stream));
  for (int i = 0; i < nfabs; ++i) sum_phi_sphr_dev(*p, *g, phi1[i], boxes[i].lo, boxes[i].hi, radii_d, nr_irreg, ps_d, nc_d);
  std::vector<unsigned long long> nc_h(nb);
  MGPU_CUDA(cudaMemcpyAsync(nc_h.data(), nc_d, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_ctx.stream));
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  if (comm_size() > 1) {  // parallel_reduce of ncell and phisum, :197-198 (counts are exact in fp64)
    std::vector<double> cnt(nb);
    for (int r = 0; r < nb; ++r) cnt[r] = (double)nc_h[r];
    MGPU_CUDA(cudaMemcpyAsync(ps_d + nb, cnt.data(), nb * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream));
    allreduce_dev(ps_d, 2 * nb, 0);
    MGPU_CUDA(cudaMemcpyAsync(cnt.data(), ps_d + nb, nb * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
    MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
    for (int r = 0; r < nb; ++r) nc_h[r] = (unsigned long long)cnt[r];
  }
  MGPU_CUDA(cudaMemcpyAsync(phisum.data() + 1, ps_d, nb * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
  MGPU_CUDA(cudaStreamSynchronize(g_ctx.stream));
  for (int r = 0; r < nb; ++r) ncell[r + 1] = (long)nc_h[r];
  average_sphr_tail(g->dr, g->nr_fine, nr_irreg, drdxfac, phisum, ncell, radii, phibar);
}

extern "C" {

int mgpu_average(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* phi, int incomp, int nr_irreg,
                 int drdxfac, double* phibar) {
  MGPU_TRY
  if (p->spherical) need_sphr(p, g);
  const int nr = p->spherical ? g->nr_fine : p->nr;
  Call c(p, (size_t)(4 * (nr + 4) + 4 * (std::max(nr_irreg, 0) + 8)) * sizeof(double) + 16384);
  std::vector<DV> v(nfabs);
  for (int i = 0; i < nfabs; ++i) {
    if (incomp < 1 || incomp > phi[i].nc) throw Error("average: incomp out of range");
    v[i] = c.view(phi[i], crange(incomp - 1, 1), (cmask_t)0).comp(incomp - 1);
  }
  average_views(p, g, nfabs, v.data(), phi, nr_irreg, drdxfac, phibar);
  c.finish();
  MGPU_CATCH
}

int mgpu_make_etarho_spherical(const mgpu_params* p, const mgpu_geom* g, int nfabs, const mgpu_fab* sold,
                               const mgpu_fab* snew, const mgpu_fab* const* umac, const mgpu_fab* const* w0mac,
                               const double* rho0_old, const double* rho0_new, const mgpu_fab* normal, int nr_irreg,
                               int drdxfac, double* etarho_ec, double* etarho_cc) {
  MGPU_TRY
  if (!p->spherical) throw Error("ERROR: make_eta_spherical should not be called for plane-parallel");  // make_eta.f90:289
  need_sphr(p, g);
  const int nr = g->nr_fine;
  size_t need = 0;
  for (int i = 0; i < nfabs; ++i) need += 2 * fab_bytes(sold[i].lo, sold[i].hi, 3, 0, 0, 1);
  Call c(p, need + geom_scratch(g) + (size_t)(4 * (nr + 4) + 4 * (std::max(nr_irreg, 0) + 8)) * sizeof(double) + 16384);
  Geom gd = make_geom(*p, *g);
  std::vector<double> nph(nr);
  for (int r = 0; r < nr; ++r) nph[r] = 0.5 * (rho0_old[r] + rho0_new[r]);  // :376-378
  const double* nph_d = upload_small(nph.data(), (size_t)nr);
  const cmask_t mrho = crange(p->rho_comp - 1, 1);
  std::vector<DV> eta(nfabs);
  for (int i = 0; i < nfabs; ++i) {  // construct_eta_cart, :345-408 (its ghost fill :322 feeds nothing: average reads valid cells)
    const int* lo = sold[i].lo;
    const int* hi = sold[i].hi;
    DV so = c.view(sold[i], mrho, (cmask_t)0), sn = c.view(snew[i], mrho, (cmask_t)0), nm = c.view(normal[i], true, false);
    DV um[3], wm[3];
    c.views(umac, i, true, false, um);
    c.views(w0mac, i, true, false, wm);
    eta[i] = arena_fab(lo, hi, 3, 0, nullptr, 1);
    const size_t mark = arena_mark();
    DV nc = arena_fab(lo, hi, 3, 0, nullptr, 1);
    put_1d_array_on_cart_dev(*p, *g, gd, nph_d, nc, false, false, lo, hi);
    eta_cart_dev(eta[i], so.comp(p->rho_comp - 1), sn.comp(p->rho_comp - 1), um, wm, nm, nc, lo, hi);
    arena_release(mark);
  }
  average_views(p, g, nfabs, eta.data(), sold, nr_irreg, drdxfac, etarho_cc);  // :327
  etarho_ec[0] = 0.0;                                                          // :337-343
  for (int r = 1; r < nr; ++r) etarho_ec[r] = 0.5 * (etarho_cc[r] + etarho_cc[r - 1]);
  etarho_ec[nr] = etarho_cc[nr - 1];
  c.finish();
  MGPU_CATCH
}

}  // extern "C"
