#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_sharded_gpu.py -q -m gpu -k "apply or matrix_free" 2>&1 | tail -3
python - <<'PY' 2>&1 | grep -v Warning | tee gpurun_out/s28_apply3d_rowmajor.txt
import sys
sys.path.insert(0, "pressio-demoapps_b200")
import torch, pressiodemoapps as pda
R = pda.InviscidFluxReconstruction
n = 256
p = pda.create_problem(pda.create_full_mesh([n] * 3, [-1, 1] * 3, 7, ("x", "y", "z")), pda.Euler3d.PeriodicSmooth, R.Weno5)
U = torch.from_numpy(p.initialCondition()).cuda()
st = torch.cuda.current_stream().cuda_stream
for nc, layout in ((2, 1), (2, 0)):
    B = torch.rand(U.numel() * nc, dtype=torch.float64, device="cuda"); Rm = torch.empty_like(B)
    for _ in range(2): p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), nc, layout, 0.0, Rm.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): p.applyJacobianDevice(U.data_ptr(), B.data_ptr(), nc, layout, 0.0, Rm.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    print("256^3 WENO5 applyJacobian %d column(s) %s: %.3f ms" % (nc, "row-major" if layout == 1 else "col-major", e0.elapsed_time(e1) / 3))
PY
