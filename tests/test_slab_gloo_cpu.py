"""world_size-2 (and 3) gloo test of the slab plumbing: after the halo exchange every rank's local array equals the
periodic window of the global state -- the N>1 path of bench.py minus the kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, nplanes, halo, plane_dofs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
    from pressiodemoapps.halo import post_halo_exchange, slab_range, wait_all
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        glob = torch.arange(nplanes * plane_dofs, dtype=torch.float64) * 0.5 + 1.0
        k0, k1 = slab_range(nplanes, rank, world)
        local = torch.full(((k1 - k0 + 2 * halo) * plane_dofs,), -7.0, dtype=torch.float64)
        local[halo * plane_dofs:(halo + k1 - k0) * plane_dofs] = glob[k0 * plane_dofs:k1 * plane_dofs]
        wait_all(post_halo_exchange(local, halo, plane_dofs, rank, world))
        planes = [(k % nplanes) for k in range(k0 - halo, k1 + halo)]
        expect = torch.cat([glob[p * plane_dofs:(p + 1) * plane_dofs] for p in planes])
        q.put((rank, bool(torch.equal(local, expect))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nplanes,halo", [(2, 8, 3), (2, 6, 3), (3, 12, 2), (1, 5, 1)])
def test_halo_exchange_gloo(world, nplanes, halo):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + world * 7 + nplanes
    procs = [ctx.Process(target=_worker, args=(r, world, port, nplanes, halo, 10, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_slab_problem_extents():
    import pressiodemoapps as pda
    mesh = pda.create_full_mesh([8, 8, 12], [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    R = pda.InviscidFluxReconstruction
    full = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5).initialCondition()
    got = []
    for r in range(3):
        p = pda.create_problem_slab(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5, r, 3)
        k0, k1, h, pd = p.slabExtent()
        assert (k0, k1, h, pd) == (4 * r, 4 * r + 4, 3, 8 * 8 * 5)
        got.append(p.slabInitialCondition())
    assert np.array_equal(np.concatenate(got), full)
    with pytest.raises(pda.PdaError):
        pda.create_problem_slab(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5, 0, 5)   # 12 planes / 5 ranks
