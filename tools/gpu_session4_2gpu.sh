#!/bin/bash
# 2-GPU session: window fast path, sharded bench N=1/2, peer-mode check + bench N=2 with the v2 kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s4_gpus.txt
echo "== sharded tests (window fast path)"
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/s4_sharded.log; tail -12 gpurun_out/s4_sharded.log
echo "== slab/peer tests incl. the 2-GPU ones"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "slab or peer" 2>&1 | tail -8 > gpurun_out/s4_peer_tests.log; tail -5 gpurun_out/s4_peer_tests.log
echo "== check_peer_multi N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_peer_multi.py --cells 128 2>&1 | tail -3 | tee gpurun_out/s4_check_peer_n2.log
echo "== sharded bench N=1"
timeout 600 python tools/bench_sharded.py --steps 10 > gpurun_out/bench_sharded_n1.json 2> gpurun_out/bench_sharded_n1.err; cat gpurun_out/bench_sharded_n1.json; tail -3 gpurun_out/bench_sharded_n1.err
echo "== sharded bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded.py --steps 10 > gpurun_out/bench_sharded_n2.json 2> gpurun_out/bench_sharded_n2.err; cat gpurun_out/bench_sharded_n2.json; tail -3 gpurun_out/bench_sharded_n2.err
echo "== bench N=2 peer"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; head -c 1200 gpurun_out/bench_r02_n2.json; echo; tail -3 gpurun_out/bench_r02_n2.err
