#!/bin/bash
# session 26 (1 GPU): compute-sanitizer, wider: memcheck over the whole parity / reference-order / gradient suites, racecheck over
# the golden fixtures and the slab / many-column tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "memcheck: pytest tests/test_parity_gpu.py tests/test_reforder_gpu.py tests/test_gradient_gpu.py (all)" > gpurun_out/sanitizer_r02b.txt
timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_reforder_gpu.py tests/test_gradient_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> gpurun_out/sanitizer_r02b.txt
K2="golden or apply_jacobian_many or slab_decomposition or two_devices"
echo "racecheck: pytest tests/test_parity_gpu.py -k '$K2'" >> gpurun_out/sanitizer_r02b.txt
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -q -m gpu -k "$K2" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | tail -8 >> gpurun_out/sanitizer_r02b.txt
cat gpurun_out/sanitizer_r02b.txt
