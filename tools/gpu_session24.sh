#!/bin/bash
# session 24 (1 GPU): final state: smoke, full suite, full default bench line, launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s24_tests.log; tail -6 gpurun_out/s24_tests.log
echo "== bench N=1 (defaults)"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_s24.json 2> gpurun_out/bench_r02_s24.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_s24.json'))
print('value %.4g ms %.3f frac %.4f e2e %.1f ms launches %d clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['gpu_launches'],d['clocks']))
print('cpu', d.get('cpu_baseline',{}).get('value'))
print('weno3', {k:v for k,v in d.get('weno3_reference_pinned',{}).items() if k in ('ms_per_step','value','fp64_frac')})
print('jv3d', d.get('apply_jacobian_matrix_free',{}).get('ms'))
print('jac', {k:v for k,v in d.get('jacobian',{}).items() if k in ('value','ms_per_eval')})
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    line=k+': vel %.3f ms (hbm %.2f)'%(v['velocity']['ms'],v['velocity']['hbm_frac'])
    if 'jacobian' in v: line+=' | jac %.3f ms (hbm %.2f)'%(v['jacobian']['ms'],v['jacobian']['hbm_frac'])
    if 'apply_jacobian' in v: line+=' | apply25 F %.2f ms C %.2f ms vec %.2f ms'%(v['apply_jacobian']['ms'],v['apply_jacobian']['row_major_ms'],v['apply_jacobian_vector']['ms'])
    print(line)
PY
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
wc -l gpurun_out/launches_r02_bench.csv
