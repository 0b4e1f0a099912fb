"""Minimal driver for ncu captures: a few evaluations of one workload through the device-pointer C-ABI.
   python tools/profile_kernel.py --workload euler3d_weno5 --n 256 --reps 3"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import torch
import pressiodemoapps as pda

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="euler3d_weno5")
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
R = pda.InviscidFluxReconstruction
st = torch.cuda.current_stream().cuda_stream
if a.workload.startswith("euler3d"):
    rec = {"weno5": R.Weno5, "weno3": R.Weno3, "fo": R.FirstOrder}[a.workload.split("_")[1]]
    mesh = pda.create_full_mesh([a.n] * 3, [-1, 1, -1, 1, -1, 1], 3 + 2 * int(rec), ("x", "y", "z"))
    p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, rec)
    U = torch.from_numpy(p.initialCondition()).cuda()
    V = torch.empty_like(U)
    for _ in range(a.reps):
        p.rightHandSideDevice(U.data_ptr(), 0.0, V.data_ptr(), st)
elif a.workload == "euler2d_jac":
    mesh = pda.create_full_mesh([a.n, a.n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    nnz = p.jacobianPattern()[1].size
    U = torch.from_numpy(p.initialCondition()).cuda()
    V = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    J = torch.empty(nnz, dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), J.data_ptr(), st)
elif a.workload == "euler2d_applyvec":   # matrix-free J*v (k_applyjac_lattice2d)
    mesh = pda.create_full_mesh([a.n, a.n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    U = torch.from_numpy(p.initialCondition()).cuda()
    B = torch.rand(p.totalDofStencilMesh(), dtype=torch.float64, device="cuda")
    Rm = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), 1, 1, 0.0, Rm.data_ptr(), st)
elif a.workload == "euler2d_apply":
    mesh = pda.create_full_mesh([a.n, a.n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    U = torch.from_numpy(p.initialCondition()).cuda()
    B = torch.rand(p.totalDofStencilMesh(), 25, dtype=torch.float64, device="cuda")
    Rm = torch.empty(p.totalDofSampleMesh(), 25, dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), 25, 1, 0.0, Rm.data_ptr(), st)
elif a.workload == "euler2d_applyvec":
    mesh = pda.create_full_mesh([a.n, a.n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    U = torch.from_numpy(p.initialCondition()).cuda()
    B = torch.rand(p.totalDofStencilMesh(), dtype=torch.float64, device="cuda")
    Rm = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), 1, 1, 0.0, Rm.data_ptr(), st)
elif a.workload == "euler2d_vel":
    mesh = pda.create_full_mesh([a.n, a.n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    U = torch.from_numpy(p.initialCondition()).cuda()
    V = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.rightHandSideDevice(U.data_ptr(), 0.0, V.data_ptr(), st)
elif a.workload.startswith("swe_"):      # swe_{fo,weno3,weno5}_{vel,jac}: SWE slip wall on an n x n lattice (cfg 3)
    _, sch, what = a.workload.split("_")
    rec = {"weno5": R.Weno5, "weno3": R.Weno3, "fo": R.FirstOrder}[sch]
    mesh = pda.create_full_mesh([a.n, a.n], [-5, 5, -5, 5], 3 + 2 * int(rec))
    p = pda.create_problem(mesh, pda.Swe2d.SlipWall, rec)
    U = torch.from_numpy(p.initialCondition()).cuda()
    V = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    if what == "jac":
        J = torch.empty(p.jacobianPattern()[1].size, dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        if what == "jac":
            p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), J.data_ptr(), st)
        else:
            p.rightHandSideDevice(U.data_ptr(), 0.0, V.data_ptr(), st)
elif a.workload.startswith("jac3d_"):    # jac3d_{weno3,weno5}: 3D Euler velocity + Jacobian (graph-driven staged kernel)
    rec = {"weno5": R.Weno5, "weno3": R.Weno3, "fo": R.FirstOrder}[a.workload.split("_")[1]]
    mesh = pda.create_full_mesh([a.n] * 3, [-1, 1, -1, 1, -1, 1], 3 + 2 * int(rec), ("x", "y", "z"))
    p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, rec)
    U = torch.from_numpy(p.initialCondition()).cuda()
    V = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    J = torch.empty(p.jacobianPattern()[1].size, dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), J.data_ptr(), st)
elif a.workload == "euler2d_apply_f":    # 25-column COLUMN-major operand (tests_perf/main.py:37 order='F')
    mesh = pda.create_full_mesh([a.n, a.n], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    U = torch.from_numpy(p.initialCondition()).cuda()
    B = torch.rand(25, p.totalDofStencilMesh(), dtype=torch.float64, device="cuda")
    Rm = torch.empty(25, p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), 25, 0, 0.0, Rm.data_ptr(), st)
torch.cuda.synchronize()
print("done", p.launchCount())
