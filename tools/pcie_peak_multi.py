"""Host<->device copy bandwidth with ALL ranks copying AT ONCE (one process per GPU, torchrun): the bound of bench.py's
`e2e` at N > 1, where every rank moves 1/N of the 5.37 GB state each way over its own PCIe link.  Per-rank and aggregate
GB/s for H2D alone, D2H alone and both at once, with and without binding the process to its GPU's NUMA node before
the pinned buffers are allocated (first touch).  If the aggregate stops growing with N, the host side (memory
bandwidth / root complexes of the -- virtualised -- box) is the limit, not the GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_peak_multi.py [GiB per rank]"""
import json
import os
import sys

import torch
import torch.distributed as dist


def numa_cpus(dev):
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        spec = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        return cpus & os.sched_getaffinity(0), node, bdf
    except Exception:
        return set(), "?", "?"


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    saved = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n = int(gib * (1 << 30)) // 8
    res = {"n_ranks": world, "gb_per_rank_per_direction": n * 8 / 1e9}
    all_cpus = os.sched_getaffinity(0)
    for bind in (False, True):
        cpus, node, bdf = numa_cpus(local)
        if bind and cpus:
            os.sched_setaffinity(0, cpus)
        hin = torch.empty(n, dtype=torch.float64).pin_memory()
        hout = torch.empty(n, dtype=torch.float64).pin_memory()
        hin.fill_(1.0)
        hout.zero_()
        os.sched_setaffinity(0, all_cpus)
        din = torch.empty(n, dtype=torch.float64, device="cuda")
        dout = torch.ones(n, dtype=torch.float64, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def h2d():
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)

        def both():
            h2d()
            d2h()
        key = "numa_bound" if bind else "unbound"
        res[key] = {"rank0_gpu": {"pci": bdf, "numa_node": node, "local_cpus": len(cpus)}}
        for name, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("both_at_once", both)):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                s1.synchronize(); s2.synchronize()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b)
                if world > 1:
                    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                best = min(best, ms)
            gbs = n * 8 / 1e9 / (best * 1e-3)
            res[key][name] = {"ms_max_over_ranks": best, "GB/s_per_rank_per_direction": gbs, "GB/s_aggregate_per_direction": gbs * world}
        del hin, hout, din, dout
    if rank == 0:
        os.write(saved, (json.dumps(res) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
