#!/bin/bash
# session 12 (1 GPU): full default bench line (all legs) + reference arm, ncu of the 3D J*v kernel and the launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02_s12_ref.json 2> gpurun_out/bench_r02_s12_ref.err; cut -c1-400 gpurun_out/bench_r02_s12_ref.json
echo "== bench N=1 (defaults)"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_s12.json 2> gpurun_out/bench_r02_s12.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_s12.json'))
print('value %.4g ms %.3f frac %.4f e2e %.1f ms launches %d clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['gpu_launches'],d['clocks']))
print('cpu', d.get('cpu_baseline'))
print('weno3', {k:v for k,v in d.get('weno3_reference_pinned',{}).items() if k in ('ms_per_step','value','fp64_frac')})
print('jv3d', d.get('apply_jacobian_matrix_free'))
print('jac', {k:v for k,v in d.get('jacobian',{}).items() if k in ('value','ms_per_eval','cpu_baseline')})
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    line=k+': vel %.3f ms (hbm %.2f)'%(v['velocity']['ms'],v['velocity']['hbm_frac'])
    if 'jacobian' in v: line+=' | jac %.3f ms (hbm %.2f)'%(v['jacobian']['ms'],v['jacobian']['hbm_frac'])
    if 'apply_jacobian' in v: line+=' | apply25 F %.2f ms C %.2f ms vec %.2f ms'%(v['apply_jacobian']['ms'],v['apply_jacobian']['row_major_ms'],v['apply_jacobian_vector']['ms'])
    print(line)
PY
echo "== ncu J*v 3D"
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:k_applyjac_lattice3d -s 1 -c 1 -o /tmp/jv3d python tools/time_apply3d.py 256 > /tmp/jv3d.log 2>&1
python tools/ncu_summary.py /tmp/jv3d.ncu-rep > gpurun_out/ncu_apply3d_r02.txt 2>&1; cat gpurun_out/ncu_apply3d_r02.txt | cut -c1-220
ncu -i /tmp/jv3d.ncu-rep --page raw --csv > gpurun_out/ncu_apply3d_r02_raw.csv 2>/dev/null
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
wc -l gpurun_out/launches_r02_bench.csv
du -sh gpurun_out
