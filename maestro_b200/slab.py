"""Slab partition of the domain across the GPUs of one node and the NCCL bootstrap.

One process per GPU; rank r owns the r-th slab along the slowest index (z in 3-D, y in 2-D).  torch.distributed
is plumbing only: it carries the 128-byte NCCL unique id from rank 0 to the other ranks; the halo exchange itself
is ncclSend/ncclRecv issued by the CUDA library on device-resident fabs (maestro_b200/csrc/mgpu_halo.cu).
"""
import ctypes as C

import numpy as np

from . import abi


def slab_bounds(n, rank, nranks):
    """[lo, hi] (inclusive) of rank's slab of n zones; the first n % nranks slabs get one extra zone."""
    base, extra = divmod(n, nranks)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0) - 1


def slab_phys_bc(phys_bc, dm, rank, nranks):
    """Per-rank physical BCs: the sides shared with a neighbouring slab are INTERIOR, exactly what FBoxLib puts in
    phys_bc_level_array(i,:,:) for a box that does not touch the domain boundary (define_bc_tower.f90:150-197)."""
    out = [list(x) for x in phys_bc]
    r = dm - 1
    if nranks > 1:
        if rank > 0:
            out[r][0] = abi.INTERIOR
        if rank < nranks - 1:
            out[r][1] = abi.INTERIOR
        if phys_bc[r][0] == abi.PERIODIC:
            out[r] = [abi.INTERIOR, abi.INTERIOR]
    return out


def halo_plan(lib, p, fab, pmask, rank, nranks):
    pl = abi.mgpu_halo_plan()
    pm = np.ascontiguousarray(pmask, dtype=np.int32)
    f = (abi.mgpu_fab * 1)(fab.cfab())
    rc = lib.mgpu_halo_plan_make(C.byref(p), f, pm.ctypes.data_as(abi.c_int_p), rank, nranks, C.byref(pl))
    if rc != 0:
        raise RuntimeError(lib.mgpu_last_error().decode())
    return pl


def comm_init_from_torch(lib, device):
    """Create the library's NCCL communicator over the ranks of the default torch.distributed group."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_char * 128)()
    if rank == 0:
        if lib.mgpu_comm_unique_id(C.cast(buf, C.c_void_p)) != 0:
            raise RuntimeError(lib.mgpu_last_error().decode())
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    idb = C.create_string_buffer(raw, 128)
    if lib.mgpu_comm_init(rank, world, C.cast(idb, C.c_void_p)) != 0:
        raise RuntimeError(lib.mgpu_last_error().decode())
    return rank, world
