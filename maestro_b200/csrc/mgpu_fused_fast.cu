// FAST build of the fused edge-state kernel: same source as mgpu_fused.cu with MGPU_FAST=1, compiled
// with -fmad=true (see Makefile).  dt/dx folded, FMA contraction allowed; parity <= 1e-12 relative.
#define MGPU_FAST 1
#include "mgpu_fused.cu"
