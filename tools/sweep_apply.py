import sys, os
sys.path.insert(0, "pressio-demoapps_b200")
import torch, pressiodemoapps as pda
R = pda.InviscidFluxReconstruction
mesh = pda.create_full_mesh([2048, 2048], [0, 1, 0, 1], 7)
p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
U = torch.from_numpy(p.initialCondition()).cuda()
st = torch.cuda.current_stream().cuda_stream
for nc in [int(a) for a in sys.argv[1:]]:
    B = torch.rand(p.totalDofStencilMesh(), nc, dtype=torch.float64, device="cuda")
    Rm = torch.empty(p.totalDofSampleMesh(), nc, dtype=torch.float64, device="cuda")
    for _ in range(2): p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), nc, 1, 0.0, Rm.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), nc, 1, 0.0, Rm.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("PDA_FUSED_APPLY_MAX_COLS", "8"), nc, round(e0.elapsed_time(e1) / 3, 3), flush=True)
