"""The multi-process halo exchange of pressiodemoapps.sharded on gloo (world size 2 and 3, CPU tensors): every rank
ends up with exactly the full state's values in its halo planes -- periodic and non-periodic slab axes, 2D and 3D,
state vectors and row-major operands.  No compute call, no GPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, "pressio-demoapps_b200"))
import pressiodemoapps as pda
from pressiodemoapps.sharded import Shard, exchange_halos
R = pda.InviscidFluxReconstruction
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cases = [([14, 18], [0, 1, 0, 1], (), 7, pda.Euler2d.Riemann, R.Weno5),
         ([12, 16], [-1, 1, -1, 1], ("x", "y"), 5, pda.Euler2d.PeriodicSmooth, R.Weno3),
         ([6, 5, 12], [-1, 1] * 3, ("x", "y", "z"), 7, pda.Euler3d.PeriodicSmooth, R.Weno5),
         ([60, 1], [-0.5, 0.5], (), 7, pda.Euler1d.Sod, R.Weno5)]
for n, b, per, st, enum, rec in cases:
    full = pda.create_full_mesh(n, b, st, per)
    pf = pda.create_problem(full, enum, rec)
    Uf = pf.initialCondition() * (1 + 1e-3 * np.random.default_rng(3).uniform(-1, 1, pf.totalDofStencilMesh()))
    s = Shard(full, enum, rec, rank=rank, nranks=world)
    for ncols in (0, 3):
        src = Uf if ncols == 0 else np.stack([Uf * (c + 1) for c in range(ncols)], axis=1)
        u = torch.full((s.local_size(),) + ((ncols,) if ncols else ()), float("nan"), dtype=torch.float64)
        u[s.owned()] = torch.from_numpy(np.ascontiguousarray(src[s.global_rows()]))
        exchange_halos(s, u, dist)
        want = src[s.global_columns()]
        assert np.array_equal(u.numpy(), want), (n, ncols, rank)
dist.barrier()
if rank == 0:
    print("sharded gloo exchange ok", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_halo_exchange_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29650 + world), str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "sharded gloo exchange ok" in r.stdout, r.stdout[-3000:]
