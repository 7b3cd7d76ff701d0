// Slab-neighbour halo exchange over NCCL (see mgpu_halo.cu).
#pragma once
#include <cstring>

#include "mgpu_common.cuh"

namespace mgpu {

void halo_plan_make(int dm, const int* domlo, const int* domhi, const int* lo, const int* hi, int ng, const int* nodal,
                    const long* ext, const int* pmask, int rank, int nranks, mgpu_halo_plan* pl);
void comm_init(int rank, int nranks, const void* unique_id);
void comm_unique_id(void* out128);
void comm_finalize();
int comm_rank();
int comm_size();
cudaStream_t comm_stream();
cudaEvent_t comm_event(int which);
// exchange the ghost planes (slab direction) of comps [c0, c0+nc) of s; false if single rank (caller wraps locally)
bool halo_exchange_dev(const mgpu_params& P, const DV& s, const int* lo, const int* hi, int ng, const int* nodal,
                       int c0, int nc, const int* pmask, cudaStream_t stream);
// one NCCL group around several halo_exchange_dev calls (NCCL groups nest): a single launch for all fields
void halo_group_begin();
void halo_group_end();
void halo_group_abort();  // after an exception inside a group: end the NCCL group(s), depth back to zero
void allreduce_minmax_dev(double* d_minmax2);
void allreduce_dev(double* d, int n, int op);  // op: 0 sum, 1 min, 2 max

}  // namespace mgpu
