// Probe: TMA (cp.async.bulk.tensor.3d) load of an fp64 tile with halo from a Fortran-order fab; prints pass/fail.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu      run: ./tma_probe <dtype 0=f64 1=u64 2=f32x2> <boxx> <boxy>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Pad { double v[64]; };
__global__ void k2(const Pad pad, const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2, double* out, int x, int y, int z, int n, uint32_t bytes);
__global__ void k(const __grid_constant__ CUtensorMap tm, double* out, int x, int y, int z, int n, uint32_t bytes) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  uint64_t* bar = (uint64_t*)(raw + ((n * 8 + 127) / 128) * 128);
  const uint32_t b = s32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)),
                 "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b), "r"(0) : "memory");
  for (int e = threadIdx.x; e < n; e += blockDim.x) out[e] = tile[e];
}
__global__ void k2(const Pad pad, const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2, double* out, int x, int y, int z, int n, uint32_t bytes) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  uint64_t* bar = (uint64_t*)(raw + ((n * 8 + 127) / 128) * 128);
  const uint32_t b = s32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)),
                 "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b), "r"(0) : "memory");
  for (int e = threadIdx.x; e < n; e += blockDim.x) out[e] = tile[e] + pad.v[e & 63];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int dt = argc > 1 ? atoi(argv[1]) : 0, bx = argc > 2 ? atoi(argv[2]) : 20, by = argc > 3 ? atoi(argv[3]) : 20;
  const int nn = argc > 4 ? atoi(argv[4]) : 72; const int usek2 = argc > 5 ? atoi(argv[5]) : 0; const int nx = nn, ny = nn, nz = nn;
  std::vector<double> h((size_t)nx * ny * nz);
  for (size_t q = 0; q < h.size(); ++q) h[q] = (double)q + 0.25;
  double *d, *o;
  CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  const int n = bx * by; CK(cudaMalloc(&o, n * 8));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap tm;
  const int f = dt == 2 ? 2 : 1;  // f32x2: the inner dimension counts floats
  cuuint64_t gd[3] = {(cuuint64_t)nx * f, (cuuint64_t)ny, (cuuint64_t)nz};
  cuuint64_t gs[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  cuuint32_t box[3] = {(cuuint32_t)bx * f, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
  CUtensorMapDataType t = dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : (dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = ((Enc)p)(&tm, t, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("dtype %d box %dx%d encode rc=%d\n", dt, bx, by, (int)r);
  if (r) return 1;
  const int x0 = argc > 6 ? atoi(argv[6]) : -2, y0 = argc > 7 ? atoi(argv[7]) : 5, z0 = argc > 8 ? atoi(argv[8]) : 7;
  const int smem = ((n * 8 + 127) / 128) * 128 + 64;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Pad pad; for (int q = 0; q < 64; ++q) pad.v[q] = 0.0;
  CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (usek2) k2<<<1, 128, smem>>>(pad, tm, tm, o, x0 * f, y0, z0, n, (uint32_t)n * 8);
  else k<<<1, 128, smem>>>(tm, o, x0 * f, y0, z0, n, (uint32_t)n * 8);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<double> g(n); CK(cudaMemcpy(g.data(), o, n * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int yy = 0; yy < by; ++yy) for (int xx = 0; xx < bx; ++xx) {
    const int X = x0 + xx, Y = y0 + yy;
    const double want = (X < 0 || X >= nx || Y < 0 || Y >= ny || z0 < 0 || z0 >= nz) ? 0.0 : h[(size_t)X + nx * ((size_t)Y + ny * (size_t)z0)];
    if (g[yy * bx + xx] != want) ++bad;
  }
  printf("dtype %d box %dx%d: %s (%d mismatches)\n", dt, bx, by, bad ? "FAIL" : "ok", bad);
  return bad != 0;
}
