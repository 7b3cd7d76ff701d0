// Fused 3-D edge-state kernel, third design: the upwind-first algorithm of mgpu_fused2.cu (make_edge_scal_3d,
// Source/make_edge_scal.f90:677, + ppm_3d / slope, Source/ppm.f90:1629, Source/slope.f90; one component per launch,
// FAST arithmetic) with the staging and the bookkeeping rebuilt for sm_100a:
//
//   * TMA.  The s tile of every plane -- (BX+2H+2) x (BY+2H) doubles, stencil halo included -- is fetched by ONE
//     cp.async.bulk.tensor.3d issued by one thread into a ring of H+3 shared-memory slots, each guarded by an mbarrier
//     (expect_tx / complete_tx); planes are requested two steps before their first use.  The per-thread halo staging of
//     the second design (index tables, register prefetch, publish pass) is gone, and so is the second global read of s
//     for the z stencil: the z window takes its new plane from the centre of the tile.  Cells outside the fab are
//     zero-filled by the TMA unit (they only feed columns / planes whose results are never stored).  On-the-fly input
//     transforms (X = rhoX * (1/rho): a second tensor map on the reciprocal density; rho' = rho - rho0(k)) are applied
//     to the landed tile by the whole CTA, one step before the tile is read.
//   * No register shuffling.  The second design aged its carried state (z window, parabolas, simhz, ...) with ~70
//     IMAD.MOV per thread and plane (profiles/r01k_final.md).  Here the plane loop is unrolled by 4 -- the period of
//     every ring, in registers and in shared memory -- and a value of age a lives in slot (R - a) & 3 of a 4-slot
//     array indexed by compile-time constants: ages advance by renaming.
//   * ppm_type 1 edge values are not clipped: with van Leer slopes the 4th-order edge value always lies between its
//     two cells (mgpu_upwind.cuh, edge_fast).
//
// Schedule of one step (plane t), two barriers as before:
//   cell phase   C1(t) parabolas x,y (-> smem) and z (registers); Z(t) simhz; C2(t-1) T_x,T_y,T_z; C3(t-2) G, sedgez
//   face phase   F1(t) simhx, simhy; F2(t-1) transverse states; F3(t-2) sedgex, sedgey -> HBM;
//                wait for the tile of plane t+H+1 and transform it; thread 0 (after barrier A) requests plane t+H+2
#include <cuda.h>

#include <cmath>
#include <cstdint>
#include <map>
#include <type_traits>

#include "mgpu_fused.cuh"
#include "mgpu_recon.cuh"
#include "mgpu_upwind.cuh"

namespace mgpu {
namespace {

template <int H, int BX, int BY, bool DIV>
struct Smem3 {
  // the x origin of a TMA box must be 16-byte aligned (an even fp64 coordinate; an odd one raises an illegal
  // instruction, scripts/probes/tma_probe.cu): the box is two columns wider and starts at the even coordinate at or
  // below the tile's first column
  static constexpr int TXW = BX + 2 * H + 2, TYW = BY + 2 * H;
  static constexpr int SN = TXW * TYW;                 // doubles per tile (the TMA box)
  static constexpr int SLOT = (SN + 15) / 16 * 16;     // slot pitch: a multiple of 128 B
  static constexpr int NS = H + 3;                     // tile slots
  static constexpr int KCAP = 128;                     // capacity of the per-plane constant arrays (kchunk + 16 <= KCAP)
  static constexpr int P = BX;                         // pitch of every exchange plane
  static constexpr int PL = (BY + 1) * P;              // doubles per plane (one spare row: reads at row+1 stay inside)
  enum { AX0 = 0, AX1, AY0, AY1, TZ, GX, GY, XY, YX, XZ, YZ, TXA, TXB, TYA, TYB, NPL };
  // ring of four slots (planes t .. t-3): simhx, simhy and this thread's own u(i+1)+u(i), v(j+1)+v(j)
  enum { R_SHX = 0, R_SHY, R_US, R_VS, NRING };
  static constexpr int RS = NRING * PL;
  static constexpr int NTILE = NS * SLOT * (DIV ? 2 : 1);
  // thread (0, 0) reads plane AX0 one element below its start when its x face takes the low cell (a value no face
  // uses): the pad keeps that read off the last tile slot, which the TMA unit and the tile transform write
  static constexpr int PAD = 8;
  static constexpr int NPLANES = NPL + 4 * NRING;
  // ppm_type 2 (H == 3): two tables of the limited edge values of a plane's tile, (BX+3) x BY x-faces and BX x (BY+3)
  // y-faces (planes t and t+1), see the E phase of the kernel
  static constexpr int NEX = (BX + 3) * BY, NEY = BX * (BY + 3), NE = NEX + NEY;
  // ppm_type 1 (H == 2): one table of the van Leer slopes of a plane's tile and of the cells just outside it,
  // BY x (BX+2) in x and (BY+2) x BX in y (see the S phase of the kernel); same place as the edge tables
  static constexpr int NSX = BY * (BX + 2), NSY = (BY + 2) * BX;
  static constexpr int NEDGE = (H == 3) ? 2 * NE : NSX + NSY;
  static constexpr int BYTES = (NTILE + PAD + NPLANES * PL) * 8 + 64 + 2 * KCAP * 8 + NEDGE * 8;  // + mbarriers + per-plane constants + edge tables
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}

#define PLN3(A, dy, dx) pl[SM::A * SM::PL + (dy) * SM::P + (dx)]
// plane A of the ring slot that holds the plane of age `age` (compile-time)
#define RNG3(age, A, dy, dx) rg[((R - (age)) & 3) * SM::RS + SM::A * SM::PL + (dy) * SM::P + (dx)]
#define AGE(age) ((R - (age)) & 3)

// XF: on-the-fly input transform of s (0 none, 1 multiply by sdiv (X = rhoX * (1/rho)), 2 subtract ssub(k));
// WADD: add wadd(k) to the z velocity.  BC: the box has physical boundaries (wall stencils of the reconstruction and
// the boundary-face rules of make_edge_scal.f90:900-1560 in upwind-first form, mgpu_upwind.cuh).
template <int PPM, int BX, int BY, int XF, bool WADD, bool BC>
__global__ void __launch_bounds__(BX* BY, 2)
    k_fused_edge3(const FusedArgs a, const __grid_constant__ CUtensorMap tm_s, const __grid_constant__ CUtensorMap tm_d) {
  constexpr int H = (PPM == 2) ? 3 : 2;
  constexpr bool ZROT = (PPM == 1) && !BC;  // z reconstruction carried incrementally over a 4-deep rotating window
  using SM = Smem3<H, BX, BY, XF == 1>;
  constexpr int SP = SM::TXW, P = SM::P, NS = SM::NS, NT = BX * BY;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* const tiles = reinterpret_cast<double*>(smem_raw);
  double* const dtiles = tiles + NS * SM::SLOT;
  double* const planes = tiles + SM::NTILE + SM::PAD;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(planes + SM::NPLANES * SM::PL);
  // per-plane constants of the on-the-fly transforms, staged once per CTA (KCAP doubles each, kchunk + 16 <= KCAP):
  // sWadd[m] = wadd of z-face t0 + m, sSub[m] = ssub of plane t0 + m
  double* const sWadd = reinterpret_cast<double*>(bars + 8);
  double* const sSub = sWadd + SM::KCAP;
  constexpr bool ESH = (PPM == 2) && !BC;  // x / y edge values shared through the tables below
  constexpr bool SSH = (PPM == 1) && !BC;  // x / y van Leer slopes shared through the table below
  double* const etab = sSub + SM::KCAP;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * BX + tx;
  double* const pl = planes + ty * P + tx;       // this thread's cell in plane 0
  double* const rg = pl + SM::NPL * SM::PL;      // this thread's cell in ring slot 0, plane 0

  const int ibase = a.lo[0] - 1 + blockIdx.x * (BX - 2);
  const int jbase = a.lo[1] - 1 + blockIdx.y * (BY - 2);
  const int i = ibase + tx, j = jbase + ty;
  int kz0 = a.lo[2] + blockIdx.z * a.kchunk;
  int kz1 = min(kz0 + a.kchunk - 1, a.hi[2]);
  if (a.kedge > 0) {  // thin first and last chunk (launch_fused3_xf: boxes with a physical boundary in z)
    if (blockIdx.z == 0) {
      kz0 = a.lo[2];
      kz1 = a.lo[2] + a.kedge - 1;
    } else if (blockIdx.z == gridDim.z - 1) {
      kz0 = a.hi[2] - a.kedge + 1;
      kz1 = a.hi[2];
    } else {
      kz0 = a.lo[2] + a.kedge + ((int)blockIdx.z - 1) * a.kchunk;
      kz1 = min(kz0 + a.kchunk - 1, a.hi[2] - a.kedge);
    }
  }
  const bool top = (kz1 == a.hi[2]);
  const int ic = min(i, a.hi[0] + 1), jc = min(j, a.hi[1] + 1);
  if (a.tile_mode != 0) {
    // A tile is "interior" when none of the cells it reconstructs (columns ibase .. ibase+BX-1, rows likewise, planes
    // kz0-1 .. kz1+3) lies within two cells of a physical boundary and none of its faces is a boundary face: the wall
    // stencils (ppm.f90:1758-1856, :1983-2216, slope.f90:245-285) and the boundary-face rules cannot reach it, and the
    // plain kernel gives what the boundary kernel would.  CTA-uniform.
    const int c0[3] = {ibase, jbase, kz0 - 1}, c1[3] = {ibase + BX - 1, jbase + BY - 1, kz1 + 3};
    bool interior = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (a.bclo[d] != MGPU_BC_INTERIOR && c0[d] <= a.lo[d] + 2) interior = false;
      if (a.bchi[d] != MGPU_BC_INTERIOR && c1[d] >= a.hi[d] - 2) interior = false;
    }
    if ((a.tile_mode == 1) != interior) return;
  }
  const LineBC nb = no_wall2();
  const LineBC lbx = BC ? make_linebc(3, 0, a.lo[0], a.hi[0], a.bclo[0], a.bchi[0]) : nb;
  const LineBC lby = BC ? make_linebc(3, 1, a.lo[1], a.hi[1], a.bclo[1], a.bchi[1]) : nb;
  const LineBC lbz = BC ? make_linebc(3, 2, a.lo[2], a.hi[2], a.bclo[2], a.bchi[2]) : nb;
  FaceRule frx, fry;  // rules of this thread's x- and y-face (the same on every plane)
  frx.kind = fry.kind = FB_NONE;
  frx.clamp = fry.clamp = 0;
  frx.low = fry.low = false;
  if constexpr (BC) {
    frx = face_rule(i, a.lo[0], a.hi[0], a.bclo[0], a.bchi[0], a.velnorm[0]);
    fry = face_rule(j, a.lo[1], a.hi[1], a.bclo[1], a.bchi[1], a.velnorm[1]);
  }

  // loop-invariant factors come precomputed in the kernel arguments (constant bank operands)
#define rel_eps a.rel_eps
#define tdx a.td[0]
#define tdy a.td[1]
#define tdz a.td[2]
#define c6x a.c6[0]
#define c6y a.c6[1]
#define c6z a.c6[2]
#define c4x a.c4[0]
#define c4y a.c4[1]
#define c4z a.c4[2]
#define dt2 a.dt2

  // ---- element offsets of this thread inside each fab (32-bit; advanced by one plane per step) -----------------
  const int v_row = a.umac[1].n[0];
  const int s_sz = (int)a.s.stride(2), f_sz = (int)a.force.stride(2);
  const int u_sz = (int)a.umac[0].stride(2), v_sz = (int)a.umac[1].stride(2), w_sz = (int)a.umac[2].stride(2);
  const int s_k0 = a.s.lo[2], s_k1 = a.s.lo[2] + a.s.n[2] - 1;
  const int f_k0 = a.force.lo[2], f_k1 = a.force.lo[2] + a.force.n[2] - 1;
  const int u_k0 = a.umac[0].lo[2], u_k1 = a.umac[0].lo[2] + a.umac[0].n[2] - 1;
  const int w_k0 = a.umac[2].lo[2], w_k1 = a.umac[2].lo[2] + a.umac[2].n[2] - 1;
  auto clampk = [](int k, int k0, int k1) { return max(k0, min(k, k1)) - k0; };
  auto adv = [](int& q, int k, int k0, int k1, int sz) {
    if (k >= k0 && k < k1) q += sz;
  };
  auto adv_hi = [](int& q, int k, int k1, int sz) {
    if (k < k1) q += sz;
  };
  const double* __restrict__ gs = a.s.p;
  const double* __restrict__ gmul = a.sdiv;
  const double* __restrict__ gsub = a.ssub;
  const double* __restrict__ gwadd = a.wadd;
  const double* __restrict__ gf = a.force.p;
  const double* __restrict__ gu = a.umac[0].p;
  const double* __restrict__ gv = a.umac[1].p;
  const double* __restrict__ gw = a.umac[2].p;

  // ---- TMA ring ------------------------------------------------------------------------------------------------
  const uint32_t bars_a = smem_u32(bars), tiles_a = smem_u32(tiles), dtiles_a = smem_u32(dtiles);
  const int xsh = (ibase - H - a.s.lo[0]) & 1;  // the same in every CTA (BX - 2 is even)
  const int tmx = ibase - H - a.s.lo[0] - xsh, tmy = jbase - H - a.s.lo[1];
  const int sc_idx = (ty + H) * SP + tx + H + xsh;  // this thread's cell in an s tile
  constexpr uint32_t TILE_BYTES = SM::SN * 8u * (XF == 1 ? 2u : 1u);
  auto issue = [&](int plane, int slot) {  // one thread
    const uint32_t bar = bars_a + slot * 8;
    mbar_expect_tx(bar, TILE_BYTES);
    tma_load_3d(tiles_a + slot * (SM::SLOT * 8), &tm_s, tmx, tmy, plane - s_k0, bar);
    if constexpr (XF == 1) tma_load_3d(dtiles_a + slot * (SM::SLOT * 8), &tm_d, tmx, tmy, plane - s_k0, bar);
  };
  auto transform = [&](int slot, double sub) {  // whole CTA, after the tile has landed; sub = ssub(plane) for XF == 2
    if constexpr (XF != 0) {
      double* T = tiles + slot * SM::SLOT;
      const double* D = dtiles + slot * SM::SLOT;
#pragma unroll
      for (int e = tid; e < SM::SN; e += NT) {
        if constexpr (XF == 1) T[e] = T[e] * D[e];
        else T[e] = T[e] - sub;
      }
      __syncwarp();  // the last pass runs in a partly active warp: reconverge before the shared-memory reads that follow
    }
  };

  // E phase (ppm_type 2, no boundary rules): the limiter of a cell reads the limited edge values on the four faces
  // i-1 .. i+2 of each direction (ppm.f90:1905-1974); left to itself every thread evaluates eight of them per plane
  // and every edge is evaluated four times.  Here the faces of a plane's tile are shared out over the CTA (608 edges
  // for 256 threads) a step before the plane is reconstructed; T = the (transformed) tile of that plane.
  auto edge_tables = [&](const double* T, double* E) {
    for (int e = tid; e < SM::NE; e += NT) {
      if (e < SM::NEX) {
        const int row = e / (BX + 3), fx = e - row * (BX + 3);
        E[e] = sedge2_of(T + (row + H) * SP + fx - 1 + H + xsh, 1);
      } else {
        const int e2 = e - SM::NEX, fy = e2 / BX, col = e2 - fy * BX;
        E[e] = sedge2_of(T + (fy - 1 + H) * SP + col + H + xsh, SP);
      }
    }
  };

  // S phase (ppm_type 1, no boundary rules): the parabola of a cell needs the van Leer slopes of the cells i-1, i, i+1
  // of each direction (ppm.f90:1697-1752); left to itself every thread evaluates six slopes per plane and every slope
  // is evaluated three times.  Here every thread evaluates the two slopes of its own cell a step before the plane is
  // reconstructed and two warps add the 4 x 16 cells just outside the tile; T = the (transformed) tile of that plane.
  auto slope_tables = [&](const double* T) {
    const double* q = T + sc_idx;
    etab[ty * (BX + 2) + tx + 1] = dsvl_fast(q[-1], q[0], q[1]);
    etab[SM::NSX + (ty + 1) * BX + tx] = dsvl_fast(q[-SP], q[0], q[SP]);
    if (tid < 64) {  // (spread over eight lanes of every warp instead: measured slower, 3.34 against 3.25 ms per episode)
      const int r = tid & 15, w = tid >> 4;  // w: 0 the column left of the tile, 1 right of it, 2 the row below, 3 above
      if (w < 2) {
        const double* c = T + (r + H) * SP + (w == 0 ? -1 : BX) + H + xsh;
        etab[r * (BX + 2) + (w == 0 ? 0 : BX + 1)] = dsvl_fast(c[-1], c[0], c[1]);
      } else {
        const double* c = T + ((w == 2 ? -1 : BY) + H) * SP + r + H + xsh;
        etab[SM::NSX + (w == 2 ? 0 : BY + 1) * BX + r] = dsvl_fast(c[-SP], c[0], c[SP]);
      }
    }
  };

  const int t0 = kz0 - 1;
  const int t1 = kz1 + 2 + (top ? 1 : 0);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(bars_a + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int n = 0; n <= H + 1; ++n) issue(t0 + n, n);
  }
  if constexpr (WADD || XF == 2) {
    for (int m = tid; m < SM::KCAP; m += NT) {
      if constexpr (WADD) sWadd[m] = gwadd[clampk(t0 + m, w_k0, w_k1)];
      if constexpr (XF == 2) sSub[m] = gsub[clampk(t0 + m, s_k0, s_k1)];
    }
    __syncthreads();
  }

  // ---- prologue: z window, velocities of plane t0 ------------------------------------------------------------
  const int o_s = (int)a.s.off(min(i, a.s.lo[0] + a.s.n[0] - 1), min(j, a.s.lo[1] + a.s.n[1] - 1), a.s.lo[2]);
  auto gload_s = [&](int plane) {  // transformed s(i,j,plane) straight from global memory (prologue only)
    const int pk = clampk(plane, s_k0, s_k1);
    double v = gs[o_s + pk * s_sz];
    if constexpr (XF == 1) v = v * gmul[o_s + pk * s_sz];
    if constexpr (XF == 2) v = v - gsub[pk];
    return v;
  };
  double zs[4] = {0.0, 0.0, 0.0, 0.0};      // ZROT: age a = s(t+2-a)
  double dzc[4] = {0.0, 0.0, 0.0, 0.0};     // ZROT: age 0 = van Leer slope of cell t+1, age 1 = of cell t
  double ezc[4] = {0.0, 0.0, 0.0, 0.0};     // ZROT: age 0 = edge value on z-face t+1, age 1 = on z-face t
  double sw[2 * H + 1];                     // !ZROT: sw[m] = s(i,j,t-H+m) after the shift at the top of step t
  double ez2[4] = {0.0, 0.0, 0.0, 0.0};     // ppm_type 2, no BC: limited edge value on z-face t+2-age
#pragma unroll
  for (int m = 0; m <= 2 * H; ++m) sw[m] = 0.0;
  if constexpr (ZROT) {
    // the first step runs with rotation 0: age a lives in slot (0 - a) & 3
    const double sm2 = gload_s(t0 - 2), sm1 = gload_s(t0 - 1), sc0 = gload_s(t0), sp1 = gload_s(t0 + 1);
    zs[3] = sp1;  // age 1 = s(t0+1)
    zs[2] = sc0;  // age 2 = s(t0)
    zs[1] = sm1;  // age 3 = s(t0-1)
    const double dz_m = dsvl_fast(sm2, sm1, sc0);
    dzc[3] = dsvl_fast(sm1, sc0, sp1);
    ezc[3] = edge_fast(sm1, sc0, dz_m, dzc[3]);
  } else {
#pragma unroll
    for (int m = 1; m <= 2 * H; ++m) sw[m] = gload_s(t0 - 1 - H + m);
    if constexpr (PPM == 2 && !BC) {  // edge values on z-faces t0+1, t0, t0-1 (ages 1, 2, 3 of the first step)
      ez2[3] = sedge2_of(&sw[H + 2], 1);  // s(t0+1) sits at sw[H+2] before the first shift
      ez2[2] = sedge2_of(&sw[H + 1], 1);
      ez2[1] = sedge2_of(&sw[H], 1);
    }
  }
  // running offsets for the step t about to start: q_u/q_v -> plane t+1, q_w -> z-face t+1, q_f -> plane t-2
  int q_u = (int)a.umac[0].off(ic, jc, a.umac[0].lo[2]) + clampk(t0 + 1, u_k0, u_k1) * u_sz;
  int q_v = (int)a.umac[1].off(ic, jc, a.umac[1].lo[2]) + clampk(t0 + 1, u_k0, u_k1) * v_sz;
  int q_w = (int)a.umac[2].off(ic, jc, a.umac[2].lo[2]) + clampk(t0 + 1, w_k0, w_k1) * w_sz;
  int q_f = (int)a.force.off(ic, jc, a.force.lo[2]) + clampk(t0 - 2, f_k0, f_k1) * f_sz;
  // u, v on this thread's faces and on the next ones (i+1 / j+1): age 0 = plane t+1 (loaded in the face phase of
  // step t), age 1 = plane t
  double lu[4] = {0, 0, 0, 0}, lv[4] = {0, 0, 0, 0}, lu1[4] = {0, 0, 0, 0}, lv1[4] = {0, 0, 0, 0};
  double wf[4] = {0, 0, 0, 0};  // w on z-faces: age 0 = t+1 (loaded this step), 1 = t, 2 = t-1, 3 = t-2
  {
    const int ou = (int)a.umac[0].off(ic, jc, a.umac[0].lo[2]) + clampk(t0, u_k0, u_k1) * u_sz;
    const int ov = (int)a.umac[1].off(ic, jc, a.umac[1].lo[2]) + clampk(t0, u_k0, u_k1) * v_sz;
    const int ow = (int)a.umac[2].off(ic, jc, a.umac[2].lo[2]);
    lu[3] = gu[ou];
    lv[3] = gv[ov];
    lu1[3] = gu[ou + 1];
    lv1[3] = gv[ov + v_row];
    wf[3] = gw[ow + clampk(t0, w_k0, w_k1) * w_sz];
    // (WADD: wadd of z-face t0 is added at the top of the first step, like every other face)
  }
  // carried state: arrays of four, age a in slot (R - a) & 3
  double pz0[4] = {0, 0, 0, 0}, pz1[4] = {0, 0, 0, 0};  // z parabola: age 0 = cell t, 1 = cell t-1
  double shz[4] = {0, 0, 0, 0};                         // simhz: age 0 = z-face t, 1 = t-1, 2 = t-2
  double zx[4] = {0, 0, 0, 0}, zy[4] = {0, 0, 0, 0};    // simhzx, simhzy: age 0 = z-face t-1, 1 = z-face t-2
  double gz[4] = {0, 0, 0, 0};                          // G_z: age 0 = cell t-2, 1 = cell t-3
  double s_m3 = 0.0;                                    // s(i,j,t-3) (BC only: EXT_DIR value of the final z state)
  unsigned selx = 0, sely = 0;  // 2 bits per plane (age 0,1,2): bit0 = upwind is the low cell, bit1 = |u| <= rel_eps

  const bool st_x = (tx >= 1) && (tx <= BX - 2 || i == a.hi[0] + 1) && (i <= a.hi[0] + 1) && (ty >= 1) &&
                    (ty <= BY - 2) && (j <= a.hi[1]);
  const bool st_y = (ty >= 1) && (ty <= BY - 2 || j == a.hi[1] + 1) && (j <= a.hi[1] + 1) && (tx >= 1) &&
                    (tx <= BX - 2) && (i <= a.hi[0]);
  const bool st_z = (tx >= 1) && (tx <= BX - 2) && (i <= a.hi[0]) && (ty >= 1) && (ty <= BY - 2) && (j <= a.hi[1]);
  const int ex_sz = (int)a.sedge[0].stride(2), ey_sz = (int)a.sedge[1].stride(2), ez_sz = (int)a.sedge[2].stride(2);
  int q_ex = (st_x ? (int)a.sedge[0].off(i, j, a.sedge[0].lo[2]) : 0) + (t0 - 2 - a.sedge[0].lo[2]) * ex_sz;
  int q_ey = (st_y ? (int)a.sedge[1].off(i, j, a.sedge[1].lo[2]) : 0) + (t0 - 2 - a.sedge[1].lo[2]) * ey_sz;
  int q_ez = (st_z ? (int)a.sedge[2].off(i, j, a.sedge[2].lo[2]) : 0) + (t0 - 2 - a.sedge[2].lo[2]) * ez_sz;
  double* __restrict__ gex = a.sedge[0].p;
  double* __restrict__ gey = a.sedge[1].p;
  double* __restrict__ gez = a.sedge[2].p;

  // tiles of planes t0 .. t0+H are needed by the first step
#pragma unroll
  for (int n = 0; n <= H; ++n) {
    mbar_wait(bars_a + n * 8, 0);
    transform(n, XF == 2 ? sSub[n] : 0.0);
  }
  if constexpr (ESH) {  // the tables of plane t0 (slot 0) for the first step; barrier A of that step publishes them
    if constexpr (XF != 0) __syncthreads();
    edge_tables(tiles, etab);
  }
  if constexpr (SSH) {
    static_assert(BX == 16 && BY == 16, "slope_tables shares the outer cells out over 4 x 16 threads");
    if constexpr (XF != 0) __syncthreads();
    slope_tables(tiles);
  }
  int c0 = 0;                // slot of plane t
  int slot_w = H + 1;        // slot (and parity) of the next tile to wait for: plane t+H+1
  unsigned par_w = 0;
  auto wrap = [](int s) { return s >= NS ? s - NS : s; };

  // ---- one step --------------------------------------------------------------------------------------------------
  auto step = [&](auto rtag, const int t) {
    constexpr int R = decltype(rtag)::value;
    // the force of plane t-2 (consumed at the end of this cell phase)
    const double f2 = a.force_zero ? 0.0 : gf[q_f];
    adv(q_f, t - 2, f_k0, f_k1, f_sz);
    const double u0 = lu[AGE(1)], v0 = lv[AGE(1)];  // face velocities of plane t
    if constexpr (WADD) wf[AGE(1)] += sWadd[t - t0];  // z-face t (the face loaded a step ago)
    const double w0c = wf[AGE(1)], w1 = wf[AGE(2)], w2 = wf[AGE(3)];

    __syncthreads();  // A: tile(t+H) transformed, simhx/simhy(t-1), simhxy..simhyz(t-2) are visible
    if (tid == 0) {   // the slot of plane t-1 is free: request plane t+H+2
      if constexpr (XF != 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the transform wrote the slot
      issue(t + H + 2, wrap(c0 + NS - 1));
    }
    const double* S = tiles + c0 * SM::SLOT + sc_idx;
    const double znew = tiles[wrap(c0 + H) * SM::SLOT + sc_idx];  // s(i,j,t+H)
    double s0, s1;
    // ==== cell phase ============================================================================================
    // C1(t): limited parabolas
    {
      double a0, a1;
      if constexpr (ESH) {  // edge values of plane t from the table written a step ago; the table of plane t+1 follows
        const double* const ex = etab + (R & 1) * SM::NE + ty * (BX + 3) + tx + 1;
        const double* const ey = etab + (R & 1) * SM::NE + SM::NEX + (ty + 1) * BX + tx;
        cs_limit_fast(S, 1, [&](int o) { return ex[o]; }, a0, a1);
        PLN3(AX0, 0, 0) = a0;
        PLN3(AX1, 0, 0) = a1;
        cs_limit_fast(S, SP, [&](int o) { return ey[o * BX]; }, a0, a1);
        PLN3(AY0, 0, 0) = a0;
        PLN3(AY1, 0, 0) = a1;
        edge_tables(tiles + wrap(c0 + 1) * SM::SLOT, etab + ((R + 1) & 1) * SM::NE);
      } else if constexpr (SSH) {  // slopes of plane t from the table written in the face phase of the last step
        const double* const sx = etab + ty * (BX + 2) + tx + 1;
        const double* const sy = etab + SM::NSX + (ty + 1) * BX + tx;
        const double c0v = S[0];
        {
          const double m1 = S[-1], p1 = S[1], d0 = sx[0];
          a0 = edge_fast(m1, c0v, sx[-1], d0);
          a1 = edge_fast(c0v, p1, d0, sx[1]);
          cw_limit(c0v, a0, a1);
          PLN3(AX0, 0, 0) = a0;
          PLN3(AX1, 0, 0) = a1;
        }
        {
          const double m1 = S[-SP], p1 = S[SP], d0 = sy[0];
          a0 = edge_fast(m1, c0v, sy[-BX], d0);
          a1 = edge_fast(c0v, p1, d0, sy[BX]);
          cw_limit(c0v, a0, a1);
          PLN3(AY0, 0, 0) = a0;
          PLN3(AY1, 0, 0) = a1;
        }
      } else {
        if constexpr (BC) cell_par_bc<PPM>(S, 1, i, a.slope_order, lbx, a0, a1);
        else cell_par<PPM>(S, 1, a.slope_order, nb, a0, a1);
        PLN3(AX0, 0, 0) = a0;
        if (PPM != 0) PLN3(AX1, 0, 0) = a1;
        if constexpr (BC) cell_par_bc<PPM>(S, SP, j, a.slope_order, lby, a0, a1);
        else cell_par<PPM>(S, SP, a.slope_order, nb, a0, a1);
        PLN3(AY0, 0, 0) = a0;
        if (PPM != 0) PLN3(AY1, 0, 0) = a1;
      }
      if constexpr (ZROT) {
        zs[AGE(0)] = znew;
        s0 = zs[AGE(2)];
        s1 = zs[AGE(3)];
        const double dz_n = dsvl_fast(s0, zs[AGE(1)], znew);
        const double e = edge_fast(s0, zs[AGE(1)], dzc[AGE(1)], dz_n);
        double p0 = ezc[AGE(1)], p1 = e;
        cw_limit(s0, p0, p1);
        pz0[AGE(0)] = p0;
        pz1[AGE(0)] = p1;
        dzc[AGE(0)] = dz_n;
        ezc[AGE(0)] = e;
      } else {
#pragma unroll
        for (int m = 0; m < 2 * H; ++m) sw[m] = sw[m + 1];
        sw[2 * H] = znew;
        s0 = sw[H];
        s1 = sw[H - 1];
        double p0, p1;
        if constexpr (BC) {
          cell_par_bc<PPM>(&sw[H], 1, t, a.slope_order, lbz, p0, p1);
        } else if constexpr (PPM == 2) {
          // the limiter of cell t reads the edge values on z-faces t-1 .. t+2: one new edge per step, three carried
          ez2[AGE(0)] = sedge2_of(&sw[H + 2], 1);
          cs_limit_fast(&sw[H], 1, [&](int o) { return o == -1 ? ez2[AGE(3)] : (o == 0 ? ez2[AGE(2)] : (o == 1 ? ez2[AGE(1)] : ez2[AGE(0)])); },
                        p0, p1);
        } else {
          cell_par<PPM>(&sw[H], 1, a.slope_order, nb, p0, p1);
        }
        pz0[AGE(0)] = p0;
        pz1[AGE(0)] = p1;
      }
    }
    const double pz0_0 = pz0[AGE(0)], pz1_0 = pz1[AGE(0)], pz0_1 = pz0[AGE(1)], pz1_1 = pz1[AGE(1)];
    // Z(t): simhz on z-face t (between cells t-1 and t)
    bool upz0 = w0c > 0.0, slz0 = !(fabs(w0c) > rel_eps);
    bool upz1 = w1 > 0.0, slz1 = !(fabs(w1) > rel_eps);
    bool upz2 = w2 > 0.0, slz2 = !(fabs(w2) > rel_eps);
    // faces with |u| <= rel_eps are rare (or fill whole planes): their averaging code sits behind warp-uniform branches
    const bool anyz = __any_sync(0xffffffffu, slz0 || slz1 || slz2);
    FaceRule fz0, fz1, fz2;  // rules of z-faces t, t-1, t-2 (uniform over the CTA)
    fz0.kind = fz1.kind = fz2.kind = FB_NONE;
    fz0.clamp = fz1.clamp = fz2.clamp = 0;
    fz0.low = fz1.low = fz2.low = false;
    if constexpr (BC) {
      fz0 = face_rule(t, a.lo[2], a.hi[2], a.bclo[2], a.bchi[2], a.velnorm[2]);
      fz1 = face_rule(t - 1, a.lo[2], a.hi[2], a.bclo[2], a.bchi[2], a.velnorm[2]);
      fz2 = face_rule(t - 2, a.lo[2], a.hi[2], a.bclo[2], a.bchi[2], a.velnorm[2]);
    }
    double shz0;
    if (BC && fz0.kind != FB_NONE) {
      if (fz0.kind == FB_GHOST) {
        shz0 = s0;  // QUIRK make_edge_scal.f90:1010-1011: the z-lo EXT_DIR state of this stage is s(lo), not s(lo-1)
      } else if (fz0.kind == FB_ZERO) {
        shz0 = 0.0;
      } else {
        const bool left = fz0.kind == FB_LEFT;
        shz0 = clamp_rule(forced_state<PPM>(left, left ? pz0_1 : pz0_0, left ? pz1_1 : pz1_0, left ? s1 : s0, w0c, tdz,
                                            rel_eps), fz0.clamp);
      }
    } else {
      shz0 = trace1<PPM>(upz0 ? pz0_1 : pz0_0, upz0 ? pz1_1 : pz1_0, upz0 ? s1 : s0, w0c * tdz, upz0);
    }
    shz[AGE(0)] = shz0;
    const double shz1 = shz[AGE(1)], shz2 = shz[AGE(2)];
    // C2(t-1): cell-centred transverse terms of plane t-1
    const double ws1 = w0c + w1;
    constexpr int tq0 = (R & 1) ? SM::PL : 0, tq1 = SM::PL - tq0;  // TX/TY planes of t-1 (written now) and t-2
    const double tx1 = RNG3(1, R_US, 0, 0) * (RNG3(1, R_SHX, 0, 1) - RNG3(1, R_SHX, 0, 0));
    const double ty1 = RNG3(1, R_VS, 0, 0) * (RNG3(1, R_SHY, 1, 0) - RNG3(1, R_SHY, 0, 0));
    const double tx2 = pl[SM::TXA * SM::PL + tq1], ty2 = pl[SM::TYA * SM::PL + tq1];
    const double tz1 = ws1 * (shz0 - shz1);
    pl[SM::TXA * SM::PL + tq0] = tx1;
    pl[SM::TYA * SM::PL + tq0] = ty1;
    PLN3(TZ, 0, 0) = tz1;
    if (BC && fz1.kind != FB_NONE) {
      upz1 = fz1.kind == FB_LEFT;
      slz1 = false;
    }
    const double txs = upz1 ? tx2 : tx1, tys = upz1 ? ty2 : ty1;
    double zx1 = fma(-c6x, txs, shz1);  // simhzx on z-face t-1
    double zy1 = fma(-c6y, tys, shz1);  // simhzy
    if (BC && fz1.kind != FB_NONE) {
      if (fz1.kind == FB_GHOST) zx1 = zy1 = fz1.low ? sw[H - 2] : s1;  // s(lo-1) / s(hi+1) of this stage
      else if (fz1.kind == FB_ZERO) zx1 = zy1 = 0.0;
      else {
        zx1 = clamp_rule(zx1, fz1.clamp);
        zy1 = clamp_rule(zy1, fz1.clamp);
      }
    }
    zx[AGE(0)] = zx1;
    zy[AGE(0)] = zy1;
    const double zx2 = zx[AGE(1)], zy2 = zy[AGE(1)];
    // C3(t-2): cell-centred final corrections of plane t-2
    const double ws2 = w1 + w2;
    const double hf = dt2 * f2;
    const double dzy = c4z * ws2 * (zy1 - zy2), dzx = c4z * ws2 * (zx1 - zx2);
    const double us2 = RNG3(2, R_US, 0, 0), vs2 = RNG3(2, R_VS, 0, 0);
    const double gx2 = fma(c4y * vs2, PLN3(YZ, 1, 0) - PLN3(YZ, 0, 0), dzy) - hf;
    const double gy2 = fma(c4x * us2, PLN3(XZ, 0, 1) - PLN3(XZ, 0, 0), dzx) - hf;
    const double gz2 =
        fma(c4x * us2, PLN3(XY, 0, 1) - PLN3(XY, 0, 0), c4y * vs2 * (PLN3(YX, 1, 0) - PLN3(YX, 0, 0))) - hf;
    PLN3(GX, 0, 0) = gx2;
    PLN3(GY, 0, 0) = gy2;
    gz[AGE(0)] = gz2;
    {
      const double gz3 = gz[AGE(1)];
      if (BC && fz2.kind != FB_NONE) {
        upz2 = fz2.kind == FB_LEFT;
        slz2 = false;
      }
      const double g = upz2 ? gz3 : gz2;
      double e = shz2 - g;
      if (BC && fz2.kind != FB_NONE) {
        if (fz2.kind == FB_GHOST) e = fz2.low ? s_m3 : sw[H - 2];  // s(lo-1) / s(hi+1)
        else if (fz2.kind == FB_ZERO) e = 0.0;
        else e = clamp_rule(e, fz2.clamp);
      }
      const int f = t - 2;  // z-face index
      if (st_z && f >= kz0 && (f <= kz1 || (top && f == kz1 + 1))) gez[q_ez] = e;
    }
    // Faces with |w| <= rel_eps (the reference averages the left and the right state) are rare, or fill whole planes:
    // everything above took the upwind formulas; one warp-uniform branch redoes what such a face changes, from values
    // re-read out of shared memory.  (Stores inside: the block stays a branch instead of being predicated into
    // every warp.)  Boundary faces never take it (slz1 / slz2 were cleared above, fz0 is tested here).
    if (anyz) {
      if (slz0 && !(BC && fz0.kind != FB_NONE)) {
        const double sh = trace_slow<PPM>(pz0_1, s1, pz0_0, s0, w0c * tdz);
        shz[AGE(0)] = sh;
        PLN3(TZ, 0, 0) = ws1 * (sh - shz1);
      }
      if (slz1) {
        const double zxs = fma(-c6x, 0.5 * (tx2 + tx1), shz1), zys = fma(-c6y, 0.5 * (ty2 + ty1), shz1);
        zx[AGE(0)] = zxs;
        zy[AGE(0)] = zys;
        const double dzys = c4z * ws2 * (zys - zy2), dzxs = c4z * ws2 * (zxs - zx2);
        PLN3(GX, 0, 0) = fma(c4y * vs2, PLN3(YZ, 1, 0) - PLN3(YZ, 0, 0), dzys) - hf;
        PLN3(GY, 0, 0) = fma(c4x * us2, PLN3(XZ, 0, 1) - PLN3(XZ, 0, 0), dzxs) - hf;
      }
      if (slz2) {
        const int f = t - 2;
        if (st_z && f >= kz0 && (f <= kz1 || (top && f == kz1 + 1))) gez[q_ez] = shz2 - 0.5 * (gz[AGE(1)] + gz2);
      }
    }
    // velocity sums of plane t (loaded a step ago) into the ring: own cell, read back by this thread only at ages 1, 2
    RNG3(0, R_US, 0, 0) = lu1[AGE(1)] + u0;
    RNG3(0, R_VS, 0, 0) = lv1[AGE(1)] + v0;

    __syncthreads();  // B: parabolas(t), T(t-1), G(t-2) are visible
    // loads for the next step: velocities of plane t+1, w on z-face t+1 (in flight across this face phase only)
    lu[AGE(0)] = gu[q_u];
    lu1[AGE(0)] = gu[q_u + 1];
    lv[AGE(0)] = gv[q_v];
    lv1[AGE(0)] = gv[q_v + v_row];
    wf[AGE(0)] = gw[q_w];
    adv_hi(q_u, t + 1, u_k1, u_sz);
    adv_hi(q_v, t + 1, u_k1, v_sz);
    adv_hi(q_w, t + 1, w_k1, w_sz);
    // ==== face phase ============================================================================================
    // the tile of plane t+H+1 (requested a step and a half ago) has landed: transform it now, so that the shared-memory
    // round trip overlaps the face work below; it is first read after the next barrier A
    mbar_wait(bars_a + slot_w * 8, par_w);
    transform(slot_w, XF == 2 ? sSub[t - t0 + H + 1] : 0.0);
    if (++slot_w == NS) {
      slot_w = 0;
      par_w ^= 1u;
    }
    // S(t+1): the slopes of plane t+1 (its tile was transformed before barrier A); every read of the table by C1(t)
    // lies before barrier B, the next ones behind the next barrier A
    if constexpr (SSH) slope_tables(tiles + wrap(c0 + 1) * SM::SLOT);
    // F1(t): simhx, simhy
    bool anyf;
    {
      bool upx = u0 > 0.0, slowx = !(fabs(u0) > rel_eps);
      bool upy = v0 > 0.0, slowy = !(fabs(v0) > rel_eps);
      if (BC && frx.kind != FB_NONE) {
        upx = frx.kind == FB_LEFT;
        slowx = false;
      }
      if (BC && fry.kind != FB_NONE) {
        upy = fry.kind == FB_LEFT;
        slowy = false;
      }
      selx = (selx << 2) | (upx ? 1u : 0u) | (slowx ? 2u : 0u);
      sely = (sely << 2) | (upy ? 1u : 0u) | (slowy ? 2u : 0u);
      anyf = __any_sync(0xffffffffu, ((selx | sely) & 0x2au) != 0u);
    }
    {
      double shx0;
      bool up = (selx & 1u) != 0u;
      if (BC && frx.kind != FB_NONE) {
        up = frx.kind == FB_LEFT;
        const int off = up ? -1 : 0;
        if (frx.kind == FB_GHOST) shx0 = frx.low ? S[-1] : S[0];
        else if (frx.kind == FB_ZERO) shx0 = 0.0;
        else
          shx0 = clamp_rule(forced_state<PPM>(up, (pl + off)[SM::AX0 * SM::PL], (pl + off)[SM::AX1 * SM::PL], S[off], u0, tdx,
                                              rel_eps), frx.clamp);
      } else {
        const int off = up ? -1 : 0;
        shx0 = trace1<PPM>((pl + off)[SM::AX0 * SM::PL], (pl + off)[SM::AX1 * SM::PL], S[off], u0 * tdx, up);
      }
      RNG3(0, R_SHX, 0, 0) = shx0;
    }
    {
      double shy0;
      bool up = (sely & 1u) != 0u;
      if (BC && fry.kind != FB_NONE) {
        up = fry.kind == FB_LEFT;
        const int off = up ? -P : 0;
        if (fry.kind == FB_GHOST) shy0 = fry.low ? S[-SP] : S[0];
        else if (fry.kind == FB_ZERO) shy0 = 0.0;
        else
          shy0 = clamp_rule(forced_state<PPM>(up, (pl + off)[SM::AY0 * SM::PL], (pl + off)[SM::AY1 * SM::PL], S[up ? -SP : 0], v0,
                                              tdy, rel_eps), fry.clamp);
      } else {
        const int off = up ? -P : 0;
        shy0 = trace1<PPM>((pl + off)[SM::AY0 * SM::PL], (pl + off)[SM::AY1 * SM::PL], S[up ? -SP : 0], v0 * tdy, up);
      }
      RNG3(0, R_SHY, 0, 0) = shy0;
    }
    // F2(t-1): transverse face states of plane t-1
    {
      const int off = (selx & 4u) ? -1 : 0;
      const double tys2 = (pl + off)[SM::TYA * SM::PL + tq0], tzs = (pl + off)[SM::TZ * SM::PL];
      const double shx1 = RNG3(1, R_SHX, 0, 0);
      double xy = fma(-c6y, tys2, shx1), xz = fma(-c6z, tzs, shx1);
      if (BC && frx.kind != FB_NONE) {
        if (frx.kind >= FB_GHOST) xy = xz = shx1;  // EXT_DIR / REFLECT_ODD: the boundary value at every stage
        else {
          xy = clamp_rule(xy, frx.clamp);
          xz = clamp_rule(xz, frx.clamp);
        }
      }
      PLN3(XY, 0, 0) = xy;
      PLN3(XZ, 0, 0) = xz;
    }
    {
      const int off = (sely & 4u) ? -P : 0;
      const double txs2 = (pl + off)[SM::TXA * SM::PL + tq0], tzs = (pl + off)[SM::TZ * SM::PL];
      const double shy1 = RNG3(1, R_SHY, 0, 0);
      double yx = fma(-c6x, txs2, shy1), yz = fma(-c6z, tzs, shy1);
      if (BC && fry.kind != FB_NONE) {
        if (fry.kind >= FB_GHOST) yx = yz = shy1;
        else {
          yx = clamp_rule(yx, fry.clamp);
          yz = clamp_rule(yz, fry.clamp);
        }
      }
      PLN3(YX, 0, 0) = yx;
      PLN3(YZ, 0, 0) = yz;
    }
    // F3(t-2): final edge states of plane t-2
    {
      const int k = t - 2;
      const bool kin = (k >= kz0) && (k <= kz1);
      if (kin && st_x) {
        const double g = (pl + ((selx & 16u) ? -1 : 0))[SM::GX * SM::PL];
        double e = RNG3(2, R_SHX, 0, 0) - g;
        if (BC && frx.kind != FB_NONE) e = (frx.kind >= FB_GHOST) ? RNG3(2, R_SHX, 0, 0) : clamp_rule(e, frx.clamp);
        gex[q_ex] = e;
      }
      if (kin && st_y) {
        const double g = (pl + ((sely & 16u) ? -P : 0))[SM::GY * SM::PL];
        double e = RNG3(2, R_SHY, 0, 0) - g;
        if (BC && fry.kind != FB_NONE) e = (fry.kind >= FB_GHOST) ? RNG3(2, R_SHY, 0, 0) : clamp_rule(e, fry.clamp);
        gey[q_ey] = e;
      }
    }
    // faces with |u| <= rel_eps or |v| <= rel_eps: as in the cell phase, one warp-uniform branch redoes their values
    // (boundary faces never set the bits)
    if (anyf) {
      if (selx & 2u) RNG3(0, R_SHX, 0, 0) = trace_slow<PPM>(PLN3(AX0, 0, -1), S[-1], PLN3(AX0, 0, 0), S[0], u0 * tdx);
      if (sely & 2u) RNG3(0, R_SHY, 0, 0) = trace_slow<PPM>(PLN3(AY0, -1, 0), S[-SP], PLN3(AY0, 0, 0), S[0], v0 * tdy);
      if (selx & 8u) {
        const double tya = 0.5 * (pl[SM::TYA * SM::PL + tq0 - 1] + pl[SM::TYA * SM::PL + tq0]);
        const double tza = 0.5 * (PLN3(TZ, 0, -1) + PLN3(TZ, 0, 0));
        const double shx1 = RNG3(1, R_SHX, 0, 0);
        PLN3(XY, 0, 0) = fma(-c6y, tya, shx1);
        PLN3(XZ, 0, 0) = fma(-c6z, tza, shx1);
      }
      if (sely & 8u) {
        const double txa = 0.5 * (pl[SM::TXA * SM::PL + tq0 - P] + pl[SM::TXA * SM::PL + tq0]);
        const double tza = 0.5 * (PLN3(TZ, -1, 0) + PLN3(TZ, 0, 0));
        const double shy1 = RNG3(1, R_SHY, 0, 0);
        PLN3(YX, 0, 0) = fma(-c6x, txa, shy1);
        PLN3(YZ, 0, 0) = fma(-c6z, tza, shy1);
      }
      const bool kin = (t - 2 >= kz0) && (t - 2 <= kz1);
      if ((selx & 32u) && kin && st_x) gex[q_ex] = RNG3(2, R_SHX, 0, 0) - 0.5 * (PLN3(GX, 0, -1) + PLN3(GX, 0, 0));
      if ((sely & 32u) && kin && st_y) gey[q_ey] = RNG3(2, R_SHY, 0, 0) - 0.5 * (PLN3(GY, -1, 0) + PLN3(GY, 0, 0));
    }
    q_ex += ex_sz;
    q_ey += ey_sz;
    q_ez += ez_sz;
    if constexpr (BC) s_m3 = sw[H - 2];
    c0 = wrap(c0 + 1);
  };

  for (int t = t0; t <= t1; t += 4) {
    step(std::integral_constant<int, 0>{}, t);
    if (t + 1 > t1) break;
    step(std::integral_constant<int, 1>{}, t + 1);
    if (t + 2 > t1) break;
    step(std::integral_constant<int, 2>{}, t + 2);
    if (t + 3 > t1) break;
    step(std::integral_constant<int, 3>{}, t + 3);
  }
  // every requested tile has been waited for except the last request (plane t1+H+2): drain it so that no bulk copy
  // is in flight when the CTA exits
  mbar_wait(bars_a + slot_w * 8, par_w);
}

#undef rel_eps
#undef tdx
#undef tdy
#undef tdz
#undef c6x
#undef c6y
#undef c6z
#undef c4x
#undef c4y
#undef c4z
#undef dt2

// z planes per CTA (same rule as the second design): the chunk count whose last wave is fullest for the least
// repeated work (every chunk repeats 4 pipeline steps)
int fused3_auto_kchunk(int ncols, int nz, int slots) {
  double best = 0.0;
  int bk = nz;
  for (int kz = 1; kz <= (nz + 7) / 8; ++kz) {
    const int c = (nz + kz - 1) / kz;
    const int nkz = (nz + c - 1) / c;
    const double waves = double(ncols) * nkz / slots;
    const double eff = waves / std::ceil(waves) * nz / (double(nkz) * (c + 4));
    if (eff > best * 1.0001) {
      best = eff;
      bk = c;
    }
  }
  return bk;
}

bool g_thin_edge = true;    // boxes with a boundary in z: 8-plane end chunks (mgpu_set_option "thin_edge")
bool g_split_tiles = true;  // boundary boxes: interior tiles through the plain kernel (mgpu_set_option "split_tiles")

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled encode_fn() {
  static EncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    MGPU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) throw Error("make_edge_scal: cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiled)p;
  }
  return fn;
}
// tensor map of one fp64 fab component (x fastest), box = one (bx x by) tile of one plane
CUtensorMap make_tmap(const double* base, const DV& v, int bx, int by) {
  CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)v.n[0], (cuuint64_t)v.n[1], (cuuint64_t)v.n[2]};
  cuuint64_t gstr[2] = {(cuuint64_t)v.n[0] * 8u, (cuuint64_t)v.n[0] * (cuuint64_t)v.n[1] * 8u};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("make_edge_scal: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return tm;
}

// stream: nullptr = the library's stream; timed: false when the caller brackets several launches as one profile entry
template <int PPM, int BX, int BY, int XF, bool WADD, bool BC>
void launch_fused3(const FusedArgs& a0, int nx, int ny, int nz, cudaStream_t stream = nullptr, bool timed = true) {
  constexpr int H = (PPM == 2) ? 3 : 2;
  using SM = Smem3<H, BX, BY, XF == 1>;
  Context& c = ctx();
  static bool configured = false;
  static int slots = 0;
  auto kern = k_fused_edge3<PPM, BX, BY, XF, WADD, BC>;
  constexpr int bytes = SM::BYTES;
  constexpr int KMAX = SM::KCAP - 16;  // z planes per CTA at most (the per-plane constants are staged in shared memory)
  if (!configured) {
    MGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    int dev = 0, sms = 0, per_sm = 0;
    MGPU_CUDA(cudaGetDevice(&dev));
    MGPU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BX * BY, bytes));
    slots = sms * (per_sm > 0 ? per_sm : 1);
    configured = true;
  }
  FusedArgs a = a0;
  for (int d = 0; d < 3; ++d) {
    a.td[d] = a.dt / a.dx[d];
    a.c6[d] = a.td[d] * (1.0 / 6.0);
    a.c4[d] = a.td[d] * 0.25;
  }
  a.dt2 = 0.5 * a.dt;
  const int gx = (nx + BX - 3) / (BX - 2), gy = (ny + BY - 3) / (BY - 2);
  if (a.kchunk <= 0) a.kchunk = fused3_auto_kchunk(gx * gy, nz - 2 * a.kedge, slots);
  if (a.kchunk > KMAX) a.kchunk = KMAX;
  const int smem = bytes;
  const CUtensorMap tm_s = make_tmap(a.s.p, a.s, SM::TXW, SM::TYW);
  const CUtensorMap tm_d = (XF == 1) ? make_tmap(a.sdiv, a.s, SM::TXW, SM::TYW) : tm_s;
  dim3 block(BX, BY, 1);
  dim3 grid(gx, gy, a.kedge > 0 ? 2 + (nz - 2 * a.kedge + a.kchunk - 1) / a.kchunk : (nz + a.kchunk - 1) / a.kchunk);
  cudaStream_t st = stream ? stream : c.stream;
  if (timed) {
    MGPU_TIMED(TAG_FUSED_EDGE, (kern<<<grid, block, smem, st>>>(a, tm_s, tm_d)));
  } else {
    kern<<<grid, block, smem, st>>>(a, tm_s, tm_d);
    MGPU_LAUNCH_CHECK();
  }
}

// the second stream of the tile split (boundary kernel next to the plain kernel) and its fork / join events
struct PairStream {
  cudaStream_t aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
PairStream& pair_stream() {
  static PairStream ps;
  if (!ps.aux) {
    MGPU_CUDA(cudaStreamCreateWithFlags(&ps.aux, cudaStreamNonBlocking));
    MGPU_CUDA(cudaEventCreateWithFlags(&ps.fork, cudaEventDisableTiming));
    MGPU_CUDA(cudaEventCreateWithFlags(&ps.join, cudaEventDisableTiming));
  }
  return ps;
}
bool g_pair_streams = true;  // mgpu_set_option "pair_streams"

template <int PPM>
void launch_fused3_xf(const FusedArgs& a0, int nx, int ny, int nz, bool bc) {
  const int xf = a0.sdiv ? 1 : (a0.ssub ? 2 : 0);
  if (bc) {
    // Most tiles of a box with physical boundaries are out of reach of every boundary rule: those run the plain
    // kernel (incremental z reconstruction, no per-face rule tests), the shell next to the boundaries runs the boundary
    // kernel.  Two launches over the same grid; a CTA of the wrong kind exits at once.  Shorter z chunks keep the
    // boundary shell thin when the box has a boundary in z.
    FusedArgs a = a0;
    const bool zbc = a.bclo[2] != MGPU_BC_INTERIOR || a.bchi[2] != MGPU_BC_INTERIOR;
    const bool split = g_split_tiles && nx >= 3 * 14 && ny >= 3 * 14 && nz >= 24;
    if (split && zbc && a.kchunk <= 0 && nz >= 64 && g_thin_edge) {
      // a boundary in z: only the chunks at the two ends are boundary tiles, and the tile test (kernel prologue) lets a
      // chunk be interior from the fifth plane off the low wall and up to the seventh below the high one -- so the end
      // chunks are 8 planes thick and the plain kernel takes everything between them in chunks balanced over its slots
      a.kedge = 8;
      int slots = 296;
      {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
          slots = 2 * sms;
      }
      const int gx = (nx + 16 - 3) / (16 - 2), gy = (ny + 16 - 3) / (16 - 2);
      a.kchunk = fused3_auto_kchunk(gx * gy, nz - 2 * a.kedge, slots);
      if (a.kchunk > 128 - 16) a.kchunk = 128 - 16;
    }
    if (a.kchunk <= 0 && zbc && nz >= 128) a.kchunk = 32;
    if (split) {
      if (a.kchunk <= 0) a.kchunk = nz >= 128 ? 64 : nz;  // both launches must cut the same chunks
      if (g_pair_streams) {
        // the two launches write disjoint tiles and read the same inputs: the boundary kernel goes to a second stream
        // so that its CTAs fill the slots the plain kernel's partial last wave leaves idle (and the other way round).
        // One profile entry for the pair = one component, the unit the roofline line counts.
        Context& c = ctx();
        PairStream& ps = pair_stream();
        prof_begin(TAG_FUSED_EDGE);
        MGPU_CUDA(cudaEventRecord(ps.fork, c.stream));
        MGPU_CUDA(cudaStreamWaitEvent(ps.aux, ps.fork, 0));
        a.tile_mode = 2;
        launch_fused3<PPM, 16, 16, 0, false, true>(a, nx, ny, nz, ps.aux, false);
        a.tile_mode = 1;
        launch_fused3<PPM, 16, 16, 0, false, false>(a, nx, ny, nz, nullptr, false);
        MGPU_CUDA(cudaEventRecord(ps.join, ps.aux));
        MGPU_CUDA(cudaStreamWaitEvent(c.stream, ps.join, 0));
        prof_end(TAG_FUSED_EDGE);
        return;
      }
      a.tile_mode = 1;
      launch_fused3<PPM, 16, 16, 0, false, false>(a, nx, ny, nz);
      a.tile_mode = 2;
    }
    launch_fused3<PPM, 16, 16, 0, false, true>(a, nx, ny, nz);
    return;
  }
  const FusedArgs& a = a0;
  if (a.wadd) {
    if (xf == 0) launch_fused3<PPM, 16, 16, 0, true, false>(a, nx, ny, nz);
    else if (xf == 1) launch_fused3<PPM, 16, 16, 1, true, false>(a, nx, ny, nz);
    else launch_fused3<PPM, 16, 16, 2, true, false>(a, nx, ny, nz);
    return;
  }
  launch_fused3<PPM, 16, 16, 0, false, false>(a, nx, ny, nz);
}

}  // namespace

// The TMA tile needs 16-byte global strides: an even number of doubles per row of s (and of the reciprocal density,
// which has the layout of s) and 16-byte aligned component bases.
bool fused_edge3_supported(const FusedArgs& a, bool bc) {
  const int xf = a.sdiv ? 1 : (a.ssub ? 2 : 0);
  if (a.sdiv && a.ssub) return false;
  if (bc && (xf != 0 || a.wadd)) return false;
  if (!bc && !a.wadd && xf != 0) return false;
  if ((a.s.n[0] & 1) || (reinterpret_cast<uintptr_t>(a.s.p) & 15u)) return false;
  if (a.sdiv && (reinterpret_cast<uintptr_t>(a.sdiv) & 15u)) return false;
  return true;
}

void fused_edge3_set_split(int on) { g_split_tiles = on != 0; }
void fused_edge3_set_thin_edge(int on) { g_thin_edge = on != 0; }
void fused_edge3_set_pair_streams(int on) { g_pair_streams = on != 0; }

void fused_edge3_launch(const FusedArgs& a, int ppm_type, int nx, int ny, int nz, bool bc) {
  switch (ppm_type) {
    case 0: launch_fused3_xf<0>(a, nx, ny, nz, bc); break;
    case 1: launch_fused3_xf<1>(a, nx, ny, nz, bc); break;
    case 2: launch_fused3_xf<2>(a, nx, ny, nz, bc); break;
    default: throw Error("make_edge_scal: unknown ppm_type");
  }
}

}  // namespace mgpu
