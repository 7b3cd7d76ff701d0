// FAST build of the BDS kernels (see the header of mgpu_bds.cu): compiled with FMA contraction on.
#define MGPU_BDS_FAST 1
#include "mgpu_bds.cu"
