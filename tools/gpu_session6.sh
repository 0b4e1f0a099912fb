#!/bin/bash
# session 6 (1 GPU): full GPU suite (incl. the implicit regression tests), cfg legs with the deeper-prefetch 2D march,
# N=1 bench (CPU legs in their own process), ncu captures of the kernels furthest below their roofline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/s6_tests.log; tail -6 gpurun_out/s6_tests.log
echo "== configs"
timeout 900 python tools/bench_configs.py > gpurun_out/s6_configs.json 2> gpurun_out/s6_configs.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s6_configs.json'))
for k,v in d.items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    line=k+': vel %.3f ms (hbm %.2f)'%(v['velocity']['ms'],v['velocity']['hbm_frac'])
    if 'jacobian' in v: line+=' | jac %.3f ms (hbm %.2f)'%(v['jacobian']['ms'],v['jacobian']['hbm_frac'])
    if 'apply_jacobian' in v: line+=' | apply25 F %.2f ms C %.2f ms vec %.2f ms'%(v['apply_jacobian']['ms'],v['apply_jacobian']['row_major_ms'],v['apply_jacobian_vector']['ms'])
    print(line)
PY
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/bench_r02_s6.json 2> gpurun_out/bench_r02_s6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_s6.json'))
print('value %.4g ms %.3f frac %.3f e2e %.1f ms binding %s cpu %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['execution'],d.get('cpu_baseline')))
print('weno3', d.get('weno3_reference_pinned')); print('jac', {k:v for k,v in d.get('jacobian',{}).items() if k!='roofline'})"
echo "== ncu"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_velocity_march2d -s 2 -c 1 -o gpurun_out/swe_fo_vel_r02 python tools/profile_kernel.py --workload swe_fo_vel --n 4096 --reps 3 2>&1 | tail -1
timeout 300 $NCU -k regex:k_jacobian_lattice2d -s 1 -c 1 -o gpurun_out/swe_fo_jac_r02 python tools/profile_kernel.py --workload swe_fo_jac --n 2048 --reps 2 2>&1 | tail -1
timeout 300 $NCU -k regex:k_jacobian_lattice2d -s 1 -c 1 -o gpurun_out/swe_weno3_jac_r02 python tools/profile_kernel.py --workload swe_weno3_jac --n 2048 --reps 2 2>&1 | tail -1
timeout 300 $NCU -k regex:k_euler3d_velocity_tiled2 -s 2 -c 1 -o gpurun_out/vel3d_weno3_r02 python tools/profile_kernel.py --workload euler3d_weno3 --n 256 --reps 3 2>&1 | tail -1
timeout 400 $NCU -k regex:k_jacobian_inner_staged -s 1 -c 1 -o gpurun_out/jac3d_weno5_r02 python tools/profile_kernel.py --workload jac3d_weno5 --n 64 --reps 2 2>&1 | tail -1
timeout 400 $NCU -k regex:k_spmm -s 1 -c 1 -o gpurun_out/spmm_f_r02 python tools/profile_kernel.py --workload euler2d_apply_f --n 1024 --reps 2 2>&1 | tail -1
ls -la gpurun_out/*_r02.ncu-rep
