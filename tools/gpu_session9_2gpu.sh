#!/bin/bash
# session 9 (2 GPUs): does the N=2 step equal the kernel time now that no rank is pinned to core 0?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-jacobian --no-configs --no-e2e > gpurun_out/bench_r02_s9_n1.json 2>gpurun_out/bench_r02_s9_n1.err
timeout 300 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_r02_s9_n2.json 2>gpurun_out/bench_r02_s9_n2.err
timeout 300 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_r02_s9_n2b.json 2>gpurun_out/bench_r02_s9_n2b.err
python - <<'PY'
import json
base=None
for tag in ("n1","n2","n2b"):
    try:
        d=json.load(open('gpurun_out/bench_r02_s9_%s.json'%tag))
        if tag=="n1": base=d['value']
        print('%s value %.4g ms %.3f kernel_ms %.3f eff %.3f cpus %s'%(tag,d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['value']/(d['n_gpus']*base),d['execution']['host_numa_binding']))
    except Exception as e: print(tag,'failed',e)
PY
timeout 300 python -m pytest tests/test_reference_tests_py_gpu.py -q -m gpu -k "1d_linear_adv or test_version or ccu_mesh" 2>&1 | tail -3
