"""J*v of the headline problem (3D Euler WENO5 periodic) without a stored Jacobian: python tools/time_apply3d.py 256 512"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import torch, pressiodemoapps as pda
R = pda.InviscidFluxReconstruction
for n in [int(a) for a in sys.argv[1:]] or [256]:
    mesh = pda.create_full_mesh([n] * 3, [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5)
    U = torch.from_numpy(p.initialCondition()).cuda()
    b = torch.rand_like(U)
    r = torch.empty_like(U)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): p.applyJacobianDevice(U.data_ptr(), b.data_ptr(), 1, 1, 0.0, r.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): p.applyJacobianDevice(U.data_ptr(), b.data_ptr(), 1, 1, 0.0, r.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("n=%d J*v %.3f ms  %.2f Gcells/s (a stored Jacobian would hold %.2e entries)" % (n, ms, n ** 3 / ms * 1e-6, n ** 3 * 475.0), flush=True)
    del p, mesh, U, b, r
