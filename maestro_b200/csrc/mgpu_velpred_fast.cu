// FAST build of the mkutrans / velpred kernels (see the header of mgpu_velpred.cu): compiled with FMA contraction on.
#define MGPU_VP_FAST 1
#include "mgpu_velpred.cu"
