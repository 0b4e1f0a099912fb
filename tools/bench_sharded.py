"""Sharded evaluation of the non-periodic BASELINE configs over the GPUs of one box (second half):
cfg 2 (2D Euler Riemann WENO5 2048^2: velocity + Jacobian row blocks) and cfg 3 (SWE slip wall WENO3 4096^2: velocity),
one process per GPU, halo refresh by NCCL send/recv (pressiodemoapps.sharded.exchange_halos) before every evaluation.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_sharded.py [--steps K]
    python tools/bench_sharded.py            # N = 1 (no exchange)

Rank 0 prints one JSON object: per config the whole-job velocity cells/s and Jacobian nnz/s (strong scaling: the mesh
is fixed, each rank owns n/N rows), max over ranks of CUDA-event times, the share of the halo exchange, and a
correctness check of the exchange (the halo planes equal the neighbour's owned planes)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))


def main():
    import torch
    import pressiodemoapps as pda
    from pressiodemoapps.sharded import Shard, exchange_halos
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)     # NCCL_DEBUG lines go to stderr; the JSON object goes to the real stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        saved = os.dup(1)
    R = pda.InviscidFluxReconstruction
    st = torch.cuda.current_stream().cuda_stream

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(fn, K, W):
        for _ in range(W):
            fn()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        return maxr(e0.elapsed_time(e1)) / K

    def case(name, n, bounds, sten, enum, rec, jac):
        full = pda.create_full_mesh(n, bounds, sten)
        s = Shard(full, enum, rec, rank=rank, nranks=world, device=local)
        p = s.problem
        U = torch.from_numpy(p.initialCondition()).cuda()
        V = torch.empty(s.owned_size(), dtype=torch.float64, device="cuda")
        # poison the halos, exchange, compare with the initial condition (which filled them from coordinates)
        ref = U.clone()
        if s.h_lo:
            U[s.recv_lower()] = float("nan")
        if s.h_hi:
            U[s.recv_upper()] = float("nan")
        if dist is not None:
            exchange_halos(s, U, dist)
        ok = bool(torch.equal(U, ref))
        ncells = int(np.prod(n))

        def vel():
            if dist is not None:
                exchange_halos(s, U, dist)
            p.rightHandSideDevice(U.data_ptr(), 0.0, V.data_ptr(), st)

        def vel_only():
            p.rightHandSideDevice(U.data_ptr(), 0.0, V.data_ptr(), st)
        l0 = p.launchCount()
        ms_v = timed(vel, a.steps, a.warmup)
        launches_v = (p.launchCount() - l0) // (a.steps + a.warmup)
        ms_vk = timed(vel_only, a.steps, 1)
        out = {"workload": name, "mesh": n, "n_gpus": world, "halo_exchange_ok": ok,
               "velocity": {"ms": ms_v, "cells_per_s": ncells / (ms_v * 1e-3), "kernel_only_ms": ms_vk,
                            "exchange_share": max(0.0, 1.0 - ms_vk / ms_v), "launches_per_eval": int(launches_v)},
               "window_fast_path": True}
        if jac:
            nnz_local = int(p.jacobianNnz())
            nnz = int(sumr(float(nnz_local)))
            Jv = torch.empty(nnz_local, dtype=torch.float64, device="cuda")

            def jf():
                if dist is not None:
                    exchange_halos(s, U, dist)
                p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), Jv.data_ptr(), st)
            ms_j = timed(jf, max(2, a.steps // 2), 2)
            out["jacobian"] = {"ms": ms_j, "nnz": nnz, "nnz_per_s": nnz / (ms_j * 1e-3), "rows": "owned dofs per rank, local column ids"}
            del Jv
        del U, V
        torch.cuda.empty_cache()
        return out

    res = {}
    n2 = [512, 512] if a.small else [2048, 2048]
    n3 = [1024, 1024] if a.small else [4096, 4096]
    res["cfg2_euler2d_riemann_weno5"] = case("2D Euler Riemann WENO5 %dx%d, y-slabs" % tuple(n2), n2, [0, 1, 0, 1], 7,
                                             pda.Euler2d.Riemann, R.Weno5, True)
    res["cfg3_swe_slipwall_weno3"] = case("2D SWE slip wall WENO3 %dx%d, y-slabs" % tuple(n3), n3, [-5, 5, -5, 5], 5,
                                          pda.Swe2d.SlipWall, R.Weno3, True)
    res["cfg3_swe_slipwall_firstorder"] = case("2D SWE slip wall first order %dx%d, y-slabs" % tuple(n3), n3, [-5, 5, -5, 5], 3,
                                               pda.Swe2d.SlipWall, R.FirstOrder, False)
    if rank == 0:
        os.write(saved, (json.dumps({"tool": "bench_sharded", "n_gpus": world, "steps": a.steps, "results": res}) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
