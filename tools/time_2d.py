"""times velocity and velocity+Jacobian of a 2D lattice problem on the device (A/B runs under environment switches):
   python tools/time_2d.py swe fo 4096      families: swe | euler | burgers ; schemes: fo | weno3 | weno5"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import torch, pressiodemoapps as pda
R = pda.InviscidFluxReconstruction
fam, sch, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
rec = {"fo": R.FirstOrder, "weno3": R.Weno3, "weno5": R.Weno5}[sch]
sten = 3 + 2 * int(rec)
if fam == "swe":
    p = pda.create_problem(pda.create_full_mesh([n, n], [-5, 5, -5, 5], sten), pda.Swe2d.SlipWall, rec)
elif fam == "euler":
    p = pda.create_problem(pda.create_full_mesh([n, n], [0, 1, 0, 1], sten), pda.Euler2d.Riemann, rec)
else:
    p = pda.create_problem(pda.create_full_mesh([n, n], [-1, 1, -1, 1], sten, ("x", "y")), pda.AdvectionDiffusion2d.BurgersPeriodic,
                           rec, pda.ViscousFluxReconstruction.FirstOrder)
nnz = int(p.jacobianNnz())
U = torch.from_numpy(p.initialCondition()).cuda()
V = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
J = torch.empty(nnz, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


mv = timed(lambda: p.rightHandSideDevice(U.data_ptr(), 0.0, V.data_ptr(), st))
mj = timed(lambda: p.rightHandSideAndJacobianDevice(U.data_ptr(), 0.0, V.data_ptr(), J.data_ptr(), st))
nd = p.numDofPerCell()
print("%s %s %d: vel %.3f ms (%.0f GB/s) | vel+jac %.3f ms (%.0f GB/s, %.1f Gnnz/s) env %s" % (
    fam, sch, n, mv, 2 * nd * 8 * n * n / mv * 1e-6, mj, (nnz * 8 + 2 * nd * 8 * n * n) / mj * 1e-6, nnz / mj * 1e-6,
    {k: v for k, v in os.environ.items() if k.startswith("PDA_")}), flush=True)
