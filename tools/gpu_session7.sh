#!/bin/bash
# session 7 (1 GPU): full GPU suite, cfg legs with the single-copy staged Jacobian kernel, ncu captures exported to text on
# the box (the .ncu-rep files stay in /tmp: gpurun_out/ is limited to 64 MiB per call)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s7_tests.log; tail -8 gpurun_out/s7_tests.log
echo "== configs"
timeout 900 python tools/bench_configs.py > gpurun_out/s7_configs.json 2> gpurun_out/s7_configs.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s7_configs.json'))
for k,v in d.items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    line=k+': vel %.3f ms (hbm %.2f)'%(v['velocity']['ms'],v['velocity']['hbm_frac'])
    if 'jacobian' in v: line+=' | jac %.3f ms (hbm %.2f)'%(v['jacobian']['ms'],v['jacobian']['hbm_frac'])
    if 'apply_jacobian' in v: line+=' | apply25 F %.2f ms C %.2f ms vec %.2f ms'%(v['apply_jacobian']['ms'],v['apply_jacobian']['row_major_ms'],v['apply_jacobian_vector']['ms'])
    print(line)
PY
echo "== ncu"
cap() {  # name kernel-regex skip workload n reps
  timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/profile_kernel.py --workload $4 --n $5 --reps $6 > /tmp/$1.log 2>&1
  if [ -f /tmp/$1.ncu-rep ]; then
    python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
    ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
    ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
    head -12 gpurun_out/ncu_$1.txt | cut -c1-160
  else tail -3 /tmp/$1.log; fi
}
cap swe_fo_vel_r02 k_velocity_march2d 2 swe_fo_vel 4096 3
cap swe_fo_jac_r02 k_jacobian_lattice2d 1 swe_fo_jac 2048 2
cap swe_weno3_jac_r02 k_jacobian_lattice2d 1 swe_weno3_jac 2048 2
cap vel3d_weno3_r02 k_euler3d_velocity_tiled2 2 euler3d_weno3 256 3
cap jac3d_weno5_r02 k_jacobian_inner_staged 1 jac3d_weno5 64 2
cap spmm_f_r02 k_spmm 1 euler2d_apply_f 1024 2
du -sh gpurun_out
