"""Multi-process check of the peer-mode halo exchange (run under torchrun, one rank per GPU):
every rank evaluates its z-slab with pda_slab_velocity_peer_dev (neighbours' halos arrive by copy-engine pushes into
IPC-mapped buffers) for several changing states and compares with the full-mesh velocity it computes on its own GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_peer_multi.py --cells 128
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import numpy as np
import torch
import torch.distributed as dist
import pressiodemoapps as pda
from pressiodemoapps.halo import connect_peer_halo

ap = argparse.ArgumentParser()
ap.add_argument("--cells", dest="n", type=int, default=128)
ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
R = pda.InviscidFluxReconstruction
n = a.n
mesh = pda.create_full_mesh([n, n, n], [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
full = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5, device=lr)
p = pda.create_problem_slab(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5, rank, world, device=lr)
k0, k1, h, pd = p.slabExtent()
connect_peer_halo(p, rank, world)
st = torch.cuda.current_stream().cuda_stream
U0 = full.initialCondition()
bad = 0
dV = torch.zeros((k1 - k0) * pd, dtype=torch.float64, device="cuda")
Vf = torch.empty(U0.size, dtype=torch.float64, device="cuda")
for it in range(a.iters):
    rng = np.random.default_rng(1000 + it)          # same state on every rank
    U = torch.from_numpy(U0 * (1 + 1e-3 * rng.uniform(-1, 1, U0.size))).cuda()
    Uo = U[k0 * pd:k1 * pd].clone()
    full.rightHandSideDevice(U.data_ptr(), 0.0, Vf.data_ptr(), st)
    # back-to-back evaluations without host synchronisation between ranks: the epoch/parity protocol is on its own
    p.slabVelocityPeerDevice(Uo.data_ptr(), 0.0, dV.data_ptr(), st)
    torch.cuda.synchronize()
    ok = torch.equal(dV, Vf[k0 * pd:k1 * pd])
    bad += 0 if ok else 1
    # host-pointer flavour (chunked H2D -> kernel -> D2H pipeline, pushes after the boundary chunks)
    hU = Uo.cpu().pin_memory()
    hV = torch.zeros_like(hU).pin_memory()
    p.slabVelocityPeer(hU.numpy(), 0.0, hV.numpy())
    ok = torch.equal(hV, Vf[k0 * pd:k1 * pd].cpu())
    bad += 0 if ok else 1
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print("peer-mode multi-process check: world %d, n %d, %d iterations, mismatching evaluations: %d" % (world, n, a.iters, int(t.item())), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
