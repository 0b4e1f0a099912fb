#!/bin/bash
# session 16 (1 GPU): full suite with the (value, tangent) J*v kernels; 2D J*v A/B; ncu of the tiled 3D J*v kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s16_tests.log; tail -8 gpurun_out/s16_tests.log
echo "== 2D J*v A/B (cols 1 4 8; col-major 4)"
( PDA_APPLY2D_MARCH=0 python tools/sweep_apply.py 1 4 8; python tools/sweep_apply.py 1 4 8 ) 2>&1 | grep -v Warning | tee gpurun_out/s16_jv2d.txt
echo "== ncu tiled J*v 3D"
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:k_applyjac_tiled3d -s 1 -c 1 -o /tmp/jv3dt python tools/time_apply3d.py 256 > /tmp/jv3dt.log 2>&1
python tools/ncu_summary.py /tmp/jv3dt.ncu-rep > gpurun_out/ncu_apply3d_tiled_r02.txt 2>&1; cat gpurun_out/ncu_apply3d_tiled_r02.txt | cut -c1-220
echo "== ncu marching J*v 2D"
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:k_applyjac_march2d -s 1 -c 1 -o /tmp/jv2d python tools/profile_kernel.py --workload euler2d_applyvec --n 2048 --reps 3 > /tmp/jv2d.log 2>&1
python tools/ncu_summary.py /tmp/jv2d.ncu-rep > gpurun_out/ncu_apply2d_march_r02.txt 2>&1; cat gpurun_out/ncu_apply2d_march_r02.txt | cut -c1-220
