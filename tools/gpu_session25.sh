#!/bin/bash
# session 25 (1 GPU): compute-sanitizer on the round-2 kernels (memcheck + racecheck over the tests that reach them)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
K="lattice_and_graph or matrix_free or slab_peer_mode_equals or structured or medium or 3d_weno5 or sharded or jacobian_vs_finite"
echo "memcheck: pytest tests/test_parity_gpu.py tests/test_sharded_gpu.py -k '$K'" > gpurun_out/sanitizer_r02.txt
timeout 230 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_sharded_gpu.py -q -m gpu -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> gpurun_out/sanitizer_r02.txt
K2="lattice_and_graph or matrix_free or structured or 3d_weno5"
echo "racecheck: pytest tests/test_parity_gpu.py -k '$K2'" >> gpurun_out/sanitizer_r02.txt
timeout 230 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -q -m gpu -k "$K2" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | tail -8 >> gpurun_out/sanitizer_r02.txt
cat gpurun_out/sanitizer_r02.txt
