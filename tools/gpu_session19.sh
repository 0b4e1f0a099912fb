#!/bin/bash
# session 19 (1 GPU): generic 2D marching J*v kernel: matrix-free tests for every family, J*v timings per family
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_refgold_implicit_gpu.py tests/test_sharded_gpu.py -q -m gpu -k "apply or matrix_free or finite or sharded" 2>&1 | tail -4
python - <<'PY' 2>&1 | grep -v Warning | tee gpurun_out/s19_jv_families.txt
import os, sys
sys.path.insert(0, "pressio-demoapps_b200")
import torch, pressiodemoapps as pda
R = pda.InviscidFluxReconstruction
st = torch.cuda.current_stream().cuda_stream
def run(name, p):
    U = torch.from_numpy(p.initialCondition()).cuda()
    b = torch.rand_like(U); r = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    for _ in range(2): p.applyJacobianDevice(U.data_ptr(), b.data_ptr(), 1, 1, 0.0, r.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): p.applyJacobianDevice(U.data_ptr(), b.data_ptr(), 1, 1, 0.0, r.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    print("%s J*v %.3f ms (PDA_APPLY2D_MARCH=%s)" % (name, e0.elapsed_time(e1) / 5, os.environ.get("PDA_APPLY2D_MARCH", "1")), flush=True)
n = 4096
run("swe weno3 4096^2", pda.create_problem(pda.create_full_mesh([n, n], [-5, 5, -5, 5], 5), pda.Swe2d.SlipWall, R.Weno3))
run("swe fo 4096^2", pda.create_problem(pda.create_full_mesh([n, n], [-5, 5, -5, 5], 3), pda.Swe2d.SlipWall, R.FirstOrder))
run("burgers weno5 4096^2", pda.create_problem(pda.create_full_mesh([n, n], [-1, 1, -1, 1], 7, ("x", "y")), pda.AdvectionDiffusion2d.BurgersPeriodic, R.Weno5, pda.ViscousFluxReconstruction.FirstOrder))
PY
