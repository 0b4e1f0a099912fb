#!/bin/bash
# round-2 GPU session 1: parity (new reference-order mode, full-size cases), the whole GPU suite, bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/s1_gpu.txt 2>&1
nproc >> gpurun_out/s1_gpu.txt; free -g >> gpurun_out/s1_gpu.txt
echo "== reforder tests" ; timeout 900 python -m pytest tests/test_reforder_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/s1_reforder_tests.log; tail -5 gpurun_out/s1_reforder_tests.log
echo "== parity report"; timeout 1500 python tools/parity_report_r02.py > gpurun_out/parity_report_r02.txt 2> gpurun_out/parity_report_r02.err; tail -30 gpurun_out/parity_report_r02.txt; tail -5 gpurun_out/parity_report_r02.err
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_s1.json 2> gpurun_out/bench_r02_s1.err; head -c 3000 gpurun_out/bench_r02_s1.json; tail -3 gpurun_out/bench_r02_s1.err
echo "== bench LZ sweep"; for lz in 128 256 512; do PDA_TILED_LZ=$lz timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-jacobian --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('LZ',$lz,'value',d['value'],'ms',d['ms_per_step'])"; done > gpurun_out/s1_lz_sweep.txt 2>&1; cat gpurun_out/s1_lz_sweep.txt
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02_s1_ref.json 2> gpurun_out/bench_r02_s1_ref.err; cat gpurun_out/bench_r02_s1_ref.json | head -c 1500
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/s1_gpu_tests.log; tail -15 gpurun_out/s1_gpu_tests.log
