"""Host-side logic of the multi-GPU path on CPU: slab partition, per-rank BC tables and the halo-exchange plan
(mgpu_halo_plan_make, pure host arithmetic in the CUDA library) executed over torch.distributed/gloo with
world_size 2 and 3, checked against the single-box ghost fill of the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dm, nglob, ng, nodal_r, periodic, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist

    import oracle_lib
    from maestro_b200 import Fab, abi, lib, make_adv_bc, make_params, slab

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = lib.load()
        r = dm - 1
        n = [8] * dm
        n[r] = nglob
        p = make_params(dm, n=n + [1] * (3 - dm))
        pb = abi.PERIODIC if periodic else abi.OUTLET
        phys = [[abi.PERIODIC, abi.PERIODIC]] * (dm - 1) + [[pb, pb]]
        pmask = [1] * (dm - 1) + [1 if periodic else 0] + [0] * (3 - dm)
        nod = [0, 0, 0]
        nod[r] = nodal_r
        # global reference: one box, ghost cells filled by the oracle (cell-centred) / plain periodic copy (nodal)
        glo, ghi = [0, 0, 0], [n[d] - 1 if d < dm else 0 for d in range(3)]
        G = Fab(glo, ghi, ng, 2, nodal=nod, dm=dm)
        rng = np.random.default_rng(11)
        G.a[...] = rng.uniform(-1, 1, size=G.shape)
        if nodal_r:  # the duplicated periodic face hi+1 == lo
            sl = [slice(None)] * 4
            sl[3 - r] = slice(ng + nglob, ng + nglob + 1)
            s0 = [slice(None)] * 4
            s0[3 - r] = slice(ng, ng + 1)
            if periodic:
                G.a[tuple(sl)] = G.a[tuple(s0)]
            from synth import fill_face_ghosts
            fill_face_ghosts([G], pmask, dm)
        else:
            oracle_lib.load().fill_boundary(p, G, 1, dm + 1, 2, make_adv_bc(p, phys), pmask)
        # my slab, valid data only (ghosts poisoned)
        klo, khi = slab.slab_bounds(nglob, rank, world)
        lo, hi = list(glo), list(ghi)
        lo[r], hi[r] = klo, khi
        F = Fab(lo, hi, ng, 2, nodal=nod, dm=dm, fill=np.nan)
        ax = 3 - r

        def planes(fab, k0, nk):  # view of planes [k0, k0+nk) (global index) of a fab
            o = k0 - (fab.lo[r] - fab.ng)
            sl = [slice(None)] * 4
            sl[ax] = slice(o, o + nk)
            return fab.a[tuple(sl)]

        nv = khi - klo + 1 + nodal_r
        planes(F, klo, nv)[...] = planes(G, klo, nv)  # includes x/y ghosts: those are local wraps
        pl = slab.halo_plan(L, p, F, pmask, rank, world)
        assert pl.dir == r and pl.nplanes == ng
        reqs, bufs = [], []
        for (dst, k0) in ((pl.up_rank, pl.send_up_k0), (pl.dn_rank, pl.send_dn_k0)):
            if dst >= 0:
                t = torch.from_numpy(np.ascontiguousarray(planes(F, k0, ng)))
                reqs.append(dist.isend(t, dst=dst, tag=0 if dst == pl.up_rank and k0 == pl.send_up_k0 else 1))
        for (src, k0, tag) in ((pl.dn_rank, pl.recv_lo_k0, 0), (pl.up_rank, pl.recv_hi_k0, 1)):
            if src >= 0:
                t = torch.empty(planes(F, k0, ng).shape, dtype=torch.float64)
                reqs.append(dist.irecv(t, src=src, tag=tag))
                bufs.append((k0, t))
        for rq in reqs:
            rq.wait()
        for k0, t in bufs:
            planes(F, k0, ng)[...] = t.numpy()
        # every ghost plane that has a source must now equal the global reference (periodic images included)
        ok = True
        for (src, k0) in ((pl.dn_rank, pl.recv_lo_k0), (pl.up_rank, pl.recv_hi_k0)):
            if src < 0:
                continue
            kk = np.arange(k0, k0 + ng)
            for i, k in enumerate(kk):
                kg = k
                if periodic:  # periodic image inside the global box's own ghost range
                    while kg < glo[r] - ng:
                        kg += nglob
                    while kg > ghi[r] + nodal_r + ng:
                        kg -= nglob
                got = planes(F, k, 1)
                want = planes(G, kg, 1)
                ok = ok and np.array_equal(got, want)
        phys_r = slab.slab_phys_bc(phys, dm, rank, world)
        if world > 1 and 0 < rank < world - 1:
            ok = ok and phys_r[r] == [abi.INTERIOR, abi.INTERIOR]
        q.put((rank, ok, pl.up_rank, pl.dn_rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dm,nodal_r,ng,periodic", [
    (2, 3, 0, 4, True), (2, 3, 1, 1, True), (2, 2, 0, 3, False), (3, 3, 0, 4, True), (3, 2, 1, 1, False)])
def test_halo_plan_over_gloo(world, dm, nodal_r, ng, periodic):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dm, 13, ng, nodal_r, periodic, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    res.sort()
    for rank, ok, up, dn in res:
        assert ok, "rank %d ghost planes differ from the single-box fill" % rank
        if periodic:
            assert up == (rank + 1) % world and dn == (rank - 1) % world
        else:
            assert up == (rank + 1 if rank < world - 1 else -1) and dn == rank - 1


def test_slab_bounds_cover_the_domain():
    from maestro_b200 import slab

    for n in (13, 256, 7):
        for w in (1, 2, 3, 8):
            if w > n:
                continue
            b = [slab.slab_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n - 1
            assert all(b[i][1] + 1 == b[i + 1][0] for i in range(w - 1))
