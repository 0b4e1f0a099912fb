"""Host<->device copy bandwidth of this box with pinned buffers: H2D alone, D2H alone, both at once (two streams).
The e2e leg of bench.py (pda_problem_velocity_host: 5.37 GB in, 5.37 GB out per step, chunked and overlapped) is bound
by the both-at-once figure.   python tools/pcie_peak.py [GiB per direction]"""
import json
import sys
import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30)) // 8
hin = torch.empty(n, dtype=torch.float64).pin_memory()
hout = torch.empty(n, dtype=torch.float64).pin_memory()
hin.fill_(1.0)
din = torch.empty(n, dtype=torch.float64, device="cuda")
dout = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        a.record()
        fn()
        torch.cuda.synchronize()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def h2d():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hout.copy_(dout, non_blocking=True)


def both():
    h2d(); d2h()


gb = n * 8 / 1e9
res = {"gb_per_direction": gb}
for name, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("both_at_once", both)):
    ms = timed(fn)
    res[name] = {"ms": ms, "GB/s_per_direction": gb / (ms * 1e-3)}
print(json.dumps(res))
