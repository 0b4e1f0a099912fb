#!/bin/bash
# session 17 (1 GPU): 2D marching J*v kernel: tests, A/B, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_refgold_implicit_gpu.py -q -m gpu -k "apply or finite or riemann2d" 2>&1 | tail -3
( PDA_APPLY2D_MARCH=0 python tools/sweep_apply.py 1; python tools/sweep_apply.py 1 ) 2>&1 | grep -v Warning | tee gpurun_out/s17_jv2d.txt
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:k_applyjac_march2d -s 1 -c 1 -o /tmp/jv2d python tools/profile_kernel.py --workload euler2d_applyvec --n 2048 --reps 3 > /tmp/jv2d.log 2>&1
python tools/ncu_summary.py /tmp/jv2d.ncu-rep > gpurun_out/ncu_apply2d_march_r02.txt 2>&1; cat gpurun_out/ncu_apply2d_march_r02.txt | cut -c1-220
