#!/bin/bash
# session 11 (1 GPU): GPU suite (new march cases, flux JVP in the 3D J*v kernel, fused value+gradient reconstruction), A/B:
# TY=15 / LZ=128 of the headline kernel, register cap of the staged graph Jacobian kernel, J*v 3D, march Jacobians
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s11_tests.log; tail -8 gpurun_out/s11_tests.log
echo "== headline A/B"
for env in "X=1" "PDA_TILED_TY=15" "PDA_TILED_LZ=128" "PDA_TILED_LZ=32"; do
  env $env timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-jacobian --no-configs --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); w=d.get('weno3_reference_pinned',{}); a=d.get('apply_jacobian_matrix_free',{}); print('$env value %.4g ms %.3f frac %.4f | weno3 ms %.3f | Jv3d %.2f ms' % (d['value'],d['ms_per_step'],d['roofline']['frac'],w.get('ms_per_step',0),a.get('ms',0)))"
done 2>&1 | tee gpurun_out/s11_headline_ab.txt
echo "== staged graph Jacobian register cap A/B"
for occ in 0 1; do
  PDA_JAC_STAGED_OCC=$occ PDA_BENCH_ONLY=cfg4,cfg5,cfg1 timeout 600 python tools/bench_configs.py 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin)
print('OCC=$occ', ' | '.join('%s vel %.3f jac %.3f ms'%(k[:22],v['velocity']['ms'],v['jacobian']['ms']) for k,v in d.items() if 'jacobian' in v))"
done 2>&1 | tee gpurun_out/s11_staged_ab.txt
echo "== march Jacobians (fused value+gradient reconstruction)"
( python tools/time_2d.py swe weno3 4096; python tools/time_2d.py swe weno5 4096; python tools/time_2d.py burgers weno5 4096 ) 2>&1 | grep -v Warning | tee gpurun_out/s11_march.txt
python tools/time_jacobian.py 2048 2>&1 | tail -1
python tools/sweep_apply.py 1 4 8 2>/dev/null | tail -5
du -sh gpurun_out
