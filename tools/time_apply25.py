"""applyJacobian with a 25-column row-major operand on cfg 2 (2D Euler Riemann WENO5 2048^2), the workload of the
reference's tests_perf/main.py:37-48; PDA_SPMM_DMMA=0 selects the FMA form of the per-cell J*B kernel.
   python tools/time_apply25.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
import torch, pressiodemoapps as pda
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
R = pda.InviscidFluxReconstruction
mesh = pda.create_full_mesh([n, n], [0, 1, 0, 1], 7)
p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
U = torch.from_numpy(p.initialCondition()).cuda()
B = torch.rand(p.totalDofStencilMesh(), 25, dtype=torch.float64, device="cuda")
Rm = torch.empty(p.totalDofSampleMesh(), 25, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), 25, 1, 0.0, Rm.data_ptr(), st)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), 25, 1, 0.0, Rm.data_ptr(), st)
e.record(); torch.cuda.synchronize()
print("PDA_SPMM_DMMA=%s  n=%d  applyJacobian x 25 columns: %.3f ms  checksum %.12e" % (os.environ.get("PDA_SPMM_DMMA", "1"), n, a.elapsed_time(e) / 5, float(Rm.abs().sum())), flush=True)
