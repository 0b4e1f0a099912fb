#!/bin/bash
# session 20 (1 GPU): full suite, smoke, SWE first-order velocity timing, full default bench line + reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s20_tests.log; tail -6 gpurun_out/s20_tests.log
echo "== swe fo"
python tools/time_2d.py swe fo 4096 2>&1 | grep -v Warning | tee gpurun_out/s20_swe_fo.txt
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02_s20_ref.json 2> gpurun_out/bench_r02_s20_ref.err; cut -c1-300 gpurun_out/bench_r02_s20_ref.json
echo "== bench N=1 (defaults)"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_s20.json 2> gpurun_out/bench_r02_s20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_s20.json'))
print('value %.4g ms %.3f frac %.4f e2e %.1f ms launches %d clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['gpu_launches'],d['clocks']))
print('cpu', d.get('cpu_baseline'))
print('weno3', {k:v for k,v in d.get('weno3_reference_pinned',{}).items() if k in ('ms_per_step','value','fp64_frac')})
print('jv3d', d.get('apply_jacobian_matrix_free'))
print('jac', {k:v for k,v in d.get('jacobian',{}).items() if k in ('value','ms_per_eval')})
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    line=k+': vel %.3f ms (hbm %.2f)'%(v['velocity']['ms'],v['velocity']['hbm_frac'])
    if 'jacobian' in v: line+=' | jac %.3f ms (hbm %.2f)'%(v['jacobian']['ms'],v['jacobian']['hbm_frac'])
    if 'apply_jacobian' in v: line+=' | apply25 F %.2f ms C %.2f ms vec %.2f ms'%(v['apply_jacobian']['ms'],v['apply_jacobian']['row_major_ms'],v['apply_jacobian_vector']['ms'])
    print(line)
PY
