"""Multi-process check of the general sharded path (run under torchrun, one rank per GPU): every rank owns one slab
window of a NON-periodic 2D lattice and of a periodic 3D lattice, refreshes its halos with NCCL send/recv
(pressiodemoapps.sharded.exchange_halos), evaluates velocity + Jacobian rows and advances a few RK4 steps with
ShardedStepper (state resident in HBM, one halo refresh per stage); results are compared with the single-GPU problem
every rank also computes on its own device: bit for bit in reference-order mode.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded_multi.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import numpy as np
import torch
import torch.distributed as dist
import pressiodemoapps as pda
from pressiodemoapps.sharded import Shard, ShardedStepper, exchange_halos

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
R = pda.InviscidFluxReconstruction
st = torch.cuda.current_stream().cuda_stream
bad = 0
cases = [([96, 128], [0, 1, 0, 1], (), 7, pda.Euler2d.Riemann, R.Weno5),
         ([64, 96], [-5, 5, -5, 5], (), 5, pda.Swe2d.SlipWall, R.Weno3),
         ([24, 20, 64], [-1, 1] * 3, ("x", "y", "z"), 7, pda.Euler3d.PeriodicSmooth, R.Weno5)]
for n, b, per, sten, enum, rec in cases:
    full = pda.create_full_mesh(n, b, sten, per)
    for order in ("reference", "fast"):
        pf = pda.create_problem(full, enum, rec, device=lr)
        pf.setOption("order", order)
        Uf = pf.initialCondition() * (1 + 1e-3 * np.random.default_rng(11).uniform(-1, 1, pf.totalDofStencilMesh()))
        s = Shard(full, enum, rec, rank=rank, nranks=world, device=lr)
        s.problem.setOption("order", order)
        U = torch.full((s.local_size(),), float("nan"), dtype=torch.float64, device="cuda")
        U[s.owned()] = torch.from_numpy(Uf[s.global_rows()]).cuda()
        exchange_halos(s, U, dist)
        bad += 0 if np.array_equal(U.cpu().numpy(), s.scatter_from_full(Uf)) else 1
        V = torch.empty(s.owned_size(), dtype=torch.float64, device="cuda")
        s.problem.rightHandSideDevice(U.data_ptr(), 0.02, V.data_ptr(), st)
        Vf = pf.createRightHandSide()
        pf.rightHandSide(Uf, 0.02, Vf)
        torch.cuda.synchronize()
        got = V.cpu().numpy()
        if order == "reference":
            bad += 0 if np.array_equal(got, Vf[s.global_rows()]) else 1
        else:
            bad += 0 if np.allclose(got, Vf[s.global_rows()], rtol=1e-12, atol=1e-10) else 1
        if order == "reference":
            # device-resident RK4 over the shards vs pda_problem_advance_host on the full mesh
            dt, nsteps = 5e-4, 4
            Uref = Uf.copy()
            pf.advance("rk4", Uref, dt, nsteps)
            stepper = ShardedStepper(s, torch, lambda X: exchange_halos(s, X, dist))
            stepper.advance("rk4", U, dt, nsteps)
            torch.cuda.synchronize()
            got = U.cpu().numpy()
            want = s.scatter_from_full(Uref)
            bad += 0 if np.allclose(got, want, rtol=1e-11, atol=1e-11) else 1
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print("sharded multi-process check: world %d, %d cases x 2 orders, mismatches: %d" % (world, len(cases), int(t.item())), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
