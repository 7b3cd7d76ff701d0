// Fused single-launch kernels for the headline configurations (see mgpu_fused.cu).
#pragma once
#include "mgpu_common.cuh"

namespace mgpu {}  // namespace mgpu
