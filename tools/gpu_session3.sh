#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== sharded + 3D tests"
timeout 1200 python -m pytest tests/test_sharded_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/s3_sharded.log; tail -25 gpurun_out/s3_sharded.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "3d or slab or structured or golden" 2>&1 | tail -8 > gpurun_out/s3_3d.log; tail -5 gpurun_out/s3_3d.log
echo "== A/B"
for v in 0 1; do
  PDA_TILED_V2=$v PDA_TILED_LZ=64 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-jacobian --no-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); w=d.get('weno3_reference_pinned',{}); print('V2=$v value %.4g ms %.3f frac %.4f | weno3 ms %.3f frac %.3f | probe %.2f TF rate %.1f mhz %.0f' % (d['value'],d['ms_per_step'],d['roofline']['frac'],w.get('ms_per_step',0),w.get('fp64_frac',0),d['roofline']['probe']['peak'],d['roofline']['probe']['dfma_per_sm_clk'],d['roofline']['probe']['sm_mhz_under_probe']))"
done > gpurun_out/s3_ab.txt 2>&1; cat gpurun_out/s3_ab.txt
echo "== ncu v2b"
PDA_TILED_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_euler3d_velocity_tiled2 -s 2 -c 1 -o gpurun_out/vel3d_v2b_r02 -f python tools/profile_kernel.py --workload euler3d_weno5 --n 512 --reps 3 > gpurun_out/s3_ncu.log 2>&1; tail -2 gpurun_out/s3_ncu.log
echo "== configs"
timeout 900 python tools/bench_configs.py > gpurun_out/s3_configs.json 2> gpurun_out/s3_configs.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s3_configs.json'))
for k,v in d.items():
    if 'error' in v: print(k,'ERROR',v['error']); continue
    line=k+': vel %.3f ms (hbm %.2f)'%(v['velocity']['ms'],v['velocity']['hbm_frac'])
    if 'jacobian' in v: line+=' | jac %.3f ms (hbm %.2f)'%(v['jacobian']['ms'],v['jacobian']['hbm_frac'])
    if 'apply_jacobian' in v: line+=' | apply25 F %.2f ms C %.2f ms vec %.2f ms'%(v['apply_jacobian']['ms'],v['apply_jacobian']['row_major_ms'],v['apply_jacobian_vector']['ms'])
    print(line)
PY
