#!/bin/bash
# session 15 (1 GPU): tiled (value, tangent) J*v kernel: tests + A/B against the line kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sharded_gpu.py -q -m gpu -k "apply or finite or sharded" -x 2>&1 | tail -15
( PDA_APPLY3D_TILED=0 python tools/time_apply3d.py 256 512; python tools/time_apply3d.py 256 512 ) 2>&1 | grep -v Warning | tee gpurun_out/s15_jv.txt
