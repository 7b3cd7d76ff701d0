// Device-resident halo exchange between slab neighbours: the replacement of FBoxLib's multifab_fill_boundary
// (MPI point-to-point inside FBoxLib; call sites listed in SURVEY.md section 2d) for a domain partitioned into
// one slab per GPU along the slowest index (z in 3-D, y in 2-D).
//
// A slab's ghost planes in the slab direction are whole contiguous (x,y) planes of the fab (x fastest, slab index
// slowest inside a component), so there is no pack/unpack: ncclSend reads the boundary planes in place and
// ncclRecv writes the neighbour's ghost planes in place, one chunk per component, all in one NCCL group
// (NVLink 5 / NVSwitch peer traffic).  The plan (who sends which planes to whom) is pure host arithmetic and is
// exported (mgpu_halo_plan_make) so the CPU tests can execute the same plan over gloo.
//
// NCCL is resolved at run time (dlopen libnccl.so.2: the copy PyTorch has already loaded when running under
// torch.distributed, the system one under a Fortran/MPI launcher), so the library has no link-time dependency.
#include <dlfcn.h>
#include <nccl.h>

#include "mgpu_halo.cuh"

namespace mgpu {

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

template <class F>
void sym(F& f, const char* name) {
  f = (F)dlsym(g_nccl.h, name);
  if (!f) throw Error(std::string("mgpu: NCCL symbol not found: ") + name);
}

void nccl_load() {
  if (g_nccl.h) return;
  g_nccl.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.h) throw Error(std::string("mgpu: cannot load libnccl.so.2: ") + dlerror());
  sym(g_nccl.GetUniqueId, "ncclGetUniqueId");
  sym(g_nccl.CommInitRank, "ncclCommInitRank");
  sym(g_nccl.CommDestroy, "ncclCommDestroy");
  sym(g_nccl.GroupStart, "ncclGroupStart");
  sym(g_nccl.GroupEnd, "ncclGroupEnd");
  sym(g_nccl.Send, "ncclSend");
  sym(g_nccl.Recv, "ncclRecv");
  sym(g_nccl.AllReduce, "ncclAllReduce");
  sym(g_nccl.GetErrorString, "ncclGetErrorString");
}

#define MGPU_NCCL(call)                                                                                    \
  do {                                                                                                     \
    ncclResult_t r__ = (call);                                                                             \
    if (r__ != ncclSuccess)                                                                                \
      throw mgpu::Error(std::string("NCCL error: ") + g_nccl.GetErrorString(r__) + " at " + __FILE__ + ":" + \
                        std::to_string(__LINE__));                                                         \
  } while (0)

struct Comm {
  bool on = false;
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;  // dedicated communication stream (overlap with interior compute)
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};
Comm g_comm;

}  // namespace

int comm_rank() { return g_comm.rank; }
int comm_size() { return g_comm.on ? g_comm.nranks : 1; }

void halo_plan_make(int dm, const int* domlo, const int* domhi, const int* lo, const int* hi, int ng, const int* nodal,
                    const long* ext, const int* pmask, int rank, int nranks, mgpu_halo_plan* pl) {
  const int r = dm - 1;
  const int nod = nodal ? nodal[r] : 0;
  pl->dir = r;
  pl->nplanes = ng;
  pl->plane_doubles = 1;
  for (int d = 0; d < r; ++d) pl->plane_doubles *= ext[d];
  const bool at_lo = (lo[r] == domlo[r]), at_hi = (hi[r] == domhi[r]);
  const bool per = pmask[r] != 0;
  pl->dn_rank = at_lo ? (per ? nranks - 1 : -1) : rank - 1;
  pl->up_rank = at_hi ? (per ? 0 : -1) : rank + 1;
  if (nranks == 1) pl->dn_rank = pl->up_rank = -1;  // a single slab wraps inside the box (k_wrap_all)
  // planes are addressed by their index in the slab direction (the reference's global k or j)
  pl->send_up_k0 = hi[r] + 1 - ng;       // -> up neighbour's low ghosts   lo_up-ng .. lo_up-1
  pl->send_dn_k0 = lo[r] + nod;          // -> down neighbour's high ghosts hi_dn+nod+1 .. hi_dn+nod+ng
  pl->recv_lo_k0 = lo[r] - ng;           // <- from the down neighbour
  pl->recv_hi_k0 = hi[r] + nod + 1;      // <- from the up neighbour
}

void comm_init(int rank, int nranks, const void* unique_id) {
  require_init();
  if (g_comm.on) return;
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error("mgpu_comm_init: bad rank/nranks");
  g_comm.rank = rank;
  g_comm.nranks = nranks;
  if (nranks > 1) {
    nccl_load();
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    MGPU_NCCL(g_nccl.CommInitRank(&g_comm.comm, nranks, id, rank));
    MGPU_CUDA(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
    MGPU_CUDA(cudaEventCreateWithFlags(&g_comm.ev_ready, cudaEventDisableTiming));
    MGPU_CUDA(cudaEventCreateWithFlags(&g_comm.ev_done, cudaEventDisableTiming));
  }
  g_comm.on = true;
}

void comm_unique_id(void* out128) {
  nccl_load();
  ncclUniqueId id;
  MGPU_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

void comm_finalize() {
  if (!g_comm.on) return;
  if (g_comm.comm) g_nccl.CommDestroy(g_comm.comm);
  if (g_comm.stream) cudaStreamDestroy(g_comm.stream);
  if (g_comm.ev_ready) cudaEventDestroy(g_comm.ev_ready);
  if (g_comm.ev_done) cudaEventDestroy(g_comm.ev_done);
  g_comm = Comm();
}

static int g_group_depth = 0;  // > 0: inside halo_group_begin/end (one NCCL launch for the whole batch)
// ghost planes of comps [c0, c0+nc) of s in the slab direction; returns false when there is nothing to exchange
// (single rank: the caller wraps locally).  Enqueued on `stream`.
bool halo_exchange_dev(const mgpu_params& P, const DV& s, const int* lo, const int* hi, int ng, const int* nodal,
                       int c0, int nc, const int* pmask, cudaStream_t stream) {
  if (!g_comm.on || g_comm.nranks == 1 || ng == 0) return false;
  mgpu_halo_plan pl;
  long ext[3] = {s.n[0], s.n[1], s.n[2]};
  halo_plan_make(P.dm, P.domlo, P.domhi, lo, hi, ng, nodal, ext, pmask, g_comm.rank, g_comm.nranks, &pl);
  const int r = pl.dir;
  // a slab thinner than the ghost width would send planes that are its own (stale) ghost planes: FBoxLib copies from
  // the box that owns the zones, two slabs away -- not built; fail instead of exchanging wrong planes
  if (hi[r] - lo[r] + 1 < ng && (pl.up_rank >= 0 || pl.dn_rank >= 0))
    throw Error("halo exchange: a slab of " + std::to_string(hi[r] - lo[r] + 1) + " plane(s) is thinner than the ghost width " +
                std::to_string(ng) + " (use fewer ranks or a taller domain)");
  const size_t cnt = (size_t)pl.plane_doubles * pl.nplanes;
  auto plane = [&](int c, int k) { return s.p + s.cs * (long)c + (long)(k - s.lo[r]) * pl.plane_doubles; };
  const bool outer = g_group_depth > 0;
  if (!outer) prof_begin(TAG_HALO);
  MGPU_NCCL(g_nccl.GroupStart());
  struct CloseGroup {  // an exception below must not leave the NCCL group open
    bool armed = true;
    ~CloseGroup() { if (armed) g_nccl.GroupEnd(); }
  } closer;
  for (int c = c0; c < c0 + nc; ++c) {
    // order per peer must be the same on both sides: (send up, recv from down, send down, recv from up)
    if (pl.up_rank >= 0) MGPU_NCCL(g_nccl.Send(plane(c, pl.send_up_k0), cnt, ncclDouble, pl.up_rank, g_comm.comm, stream));
    if (pl.dn_rank >= 0) MGPU_NCCL(g_nccl.Recv(plane(c, pl.recv_lo_k0), cnt, ncclDouble, pl.dn_rank, g_comm.comm, stream));
    if (pl.dn_rank >= 0) MGPU_NCCL(g_nccl.Send(plane(c, pl.send_dn_k0), cnt, ncclDouble, pl.dn_rank, g_comm.comm, stream));
    if (pl.up_rank >= 0) MGPU_NCCL(g_nccl.Recv(plane(c, pl.recv_hi_k0), cnt, ncclDouble, pl.up_rank, g_comm.comm, stream));
  }
  closer.armed = false;
  MGPU_NCCL(g_nccl.GroupEnd());
  if (!outer) {
    count_launch();  // one grouped NCCL kernel
    prof_end(TAG_HALO);
  }
  return true;
}

void halo_group_begin() {
  if (!(g_comm.on && g_comm.nranks > 1)) return;
  if (g_group_depth++ == 0) prof_begin(TAG_HALO);
  MGPU_NCCL(g_nccl.GroupStart());
}
// closes a batch that an exception interrupted: the NCCL group ends and the depth returns to zero (FillBatch's destructor)
void halo_group_abort() {
  if (!(g_comm.on && g_comm.nranks > 1)) return;
  while (g_group_depth > 0) {
    g_nccl.GroupEnd();
    --g_group_depth;
  }
}
void halo_group_end() {
  if (!(g_comm.on && g_comm.nranks > 1)) return;
  MGPU_NCCL(g_nccl.GroupEnd());
  if (--g_group_depth == 0) {
    count_launch();  // one grouped NCCL kernel for every field of the batch
    prof_end(TAG_HALO);
  }
}

cudaStream_t comm_stream() { return g_comm.stream; }
cudaEvent_t comm_event(int which) { return which == 0 ? g_comm.ev_ready : g_comm.ev_done; }

void allreduce_minmax_dev(double* d_minmax2) {
  // d_minmax2[0] = local min, [1] = local max: reduce both with one MIN on (min, -max)
  if (!g_comm.on || g_comm.nranks == 1) return;
  MGPU_NCCL(g_nccl.AllReduce(d_minmax2, d_minmax2, 2, ncclDouble, ncclMin, g_comm.comm, ctx().stream));
}

// in-place all-reduce of n doubles over the ranks of the slab run (op: 0 sum, 1 min, 2 max); no-op on one rank
void allreduce_dev(double* d, int n, int op) {
  if (!g_comm.on || g_comm.nranks == 1) return;
  const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
  MGPU_NCCL(g_nccl.AllReduce(d, d, (size_t)n, ncclDouble, o, g_comm.comm, ctx().stream));
}

}  // namespace mgpu
