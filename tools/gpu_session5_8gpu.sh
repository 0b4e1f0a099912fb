#!/bin/bash
# 8-GPU session: scaling of the headline (peer mode, v2 kernel) at N = 1,2,4,8, sharded configs at N = 2,4,8, multi-process
# checks, concurrent PCIe bandwidth, the 2-device single-process test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s5_gpus.txt; nvidia-smi topo -m >> gpurun_out/s5_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== two devices in one process"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "two_devices or peer" 2>&1 | tail -4
echo "== multi-process checks"
for n in 8; do timeout 600 $TR --nproc-per-node $n --master-port 2951$n tools/check_peer_multi.py --cells 128 2>&1 | grep "check:" | tee gpurun_out/check_peer_n${n}_r02.log; done
for n in 4 8; do timeout 600 $TR --nproc-per-node $n --master-port 2952$n tools/check_sharded_multi.py 2>&1 | grep "check:" | tee gpurun_out/check_sharded_n${n}_r02.log; done
echo "== headline scaling"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-jacobian --no-configs --no-e2e > gpurun_out/bench_r02_scale_n1.json 2>gpurun_out/bench_r02_scale_n1.err
for n in 2 4; do timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_r02_scale_n$n.json 2>gpurun_out/bench_r02_scale_n$n.err; done
for n in 8; do timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_r02_scale_n$n.json 2>gpurun_out/bench_r02_scale_n$n.err; done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.load(open('gpurun_out/bench_r02_scale_n%d.json'%n))
        if n==1: base=d['value']
        print('N=%d value %.4g ms %.3f eff %.3f | e2e %s ms'%(n,d['value'],d['ms_per_step'],d['value']/(n*base),d['e2e']['ms_per_step']))
    except Exception as e: print('N=%d failed'%n, e)
PY
echo "== N=8 with LZ=16"; PDA_TILED_LZ=16 timeout 600 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=8 LZ16 value %.4g ms %.3f'%(d['value'],d['ms_per_step']))"
echo "== sharded scaling"
for n in 4 8; do timeout 600 $TR --nproc-per-node $n --master-port 2955$n tools/bench_sharded.py --steps 10 > gpurun_out/bench_sharded_n$n.json 2> gpurun_out/bench_sharded_n$n.err; done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open('gpurun_out/bench_sharded_n%d.json'%n))['results']
        print('N=%d'%n, ' | '.join('%s: vel %.3f ms%s'%(k.split('_',1)[1][:22], v['velocity']['ms'], (' J %.3f ms %.0f Gnnz/s'%(v['jacobian']['ms'], v['jacobian']['nnz_per_s']/1e9)) if 'jacobian' in v else '') for k,v in d.items()))
    except Exception as e: print('N=%d failed'%n, e)
PY
echo "== concurrent PCIe"
timeout 300 python tools/pcie_peak_multi.py 2 > gpurun_out/pcie_multi_n1.json 2>/dev/null
for n in 2 4 8; do timeout 300 $TR --nproc-per-node $n --master-port 2956$n tools/pcie_peak_multi.py 2 > gpurun_out/pcie_multi_n$n.json 2>/dev/null; done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open('gpurun_out/pcie_multi_n%d.json'%n))
        print('N=%d'%n, {k:{m:round(d[k][m]['GB/s_aggregate_per_direction'],1) for m in ('h2d_alone','d2h_alone','both_at_once')} for k in ('unbound','numa_bound')})
    except Exception as e: print('N=%d failed'%n, e)
PY
