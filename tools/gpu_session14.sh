#!/bin/bash
# session 14 (1 GPU): 3D J*v with / without the L1 prefetch of the next line; applyJacobian tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( PDA_APPLY3D_PREFETCH=0 python tools/time_apply3d.py 256 512; PDA_APPLY3D_PREFETCH=1 python tools/time_apply3d.py 256 512 ) 2>&1 | grep -v Warning | tee gpurun_out/s14_jv.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "apply or finite" 2>&1 | tail -3
