"""times velocity+Jacobian of 2D Euler Riemann WENO5 (cfg 2) on the device: python tools/time_jacobian.py 1024 2048"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import torch, pressiodemoapps as pda
R = pda.InviscidFluxReconstruction
for n in [int(a) for a in sys.argv[1:]] or [1024]:
    mesh = pda.create_full_mesh([n, n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    t0 = time.time(); nnz = p.jacobianPattern()[1].size; tp = time.time() - t0
    U = torch.from_numpy(p.initialCondition()).cuda()
    V = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    J = torch.empty(nnz, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), J.data_ptr(), st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), J.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("n=%d pattern %.2fs  %.3f ms  %.1f Gnnz/s  %.0f GB/s" % (n, tp, ms, nnz / ms * 1e-6, nnz * 8 / ms * 1e-6), flush=True)
    del p, mesh
