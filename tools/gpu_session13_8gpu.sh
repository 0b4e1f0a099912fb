#!/bin/bash
# session 13 (8 GPUs): the headline at N = 8 and 4 after the launching-thread fix; one-chunk peer slabs as an experiment
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_s13_n8.json 2>gpurun_out/bench_r02_s13_n8.err
PDA_PEER_MINCHUNKS=1 timeout 300 $TR --nproc-per-node 8 --master-port 29582 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_r02_s13_n8_1chunk.json 2>gpurun_out/bench_r02_s13_n8_1chunk.err
timeout 300 $TR --nproc-per-node 4 --master-port 29583 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_r02_s13_n4.json 2>gpurun_out/bench_r02_s13_n4.err
python - <<'PY'
import json
for tag in ("n8","n8_1chunk","n4"):
    try:
        d=json.load(open('gpurun_out/bench_r02_s13_%s.json'%tag))
        print('%s value %.4g ms %.3f kernel_ms %.3f eff vs 14.80 ms: %.3f e2e %s cpus %s'%(tag,d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],14.80/(d['n_gpus']*d['ms_per_step']),d['e2e']['ms_per_step'],d['execution']['host_numa_binding']))
    except Exception as e: print(tag,'failed',e)
PY
