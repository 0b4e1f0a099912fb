#!/bin/bash
# round-2 GPU session 2: second-generation headline kernel (A/B + ncu), reference-order twins, full-size parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== 3D + reforder + fullsize tests (v2 kernel on)"
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_reforder_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/s2_tests.log; tail -12 gpurun_out/s2_tests.log
echo "== A/B"
for v in 0 1; do for lz in 64 128; do
  PDA_TILED_V2=$v PDA_TILED_LZ=$lz timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-jacobian --no-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); w=d.get('weno3_reference_pinned',{}); print('V2=$v LZ=$lz value %.4g ms %.3f frac %.4f | weno3 ms %.3f frac %.3f | probe %.2f TF rate %.1f mhz %.0f | e2e %.1f ms' % (d['value'],d['ms_per_step'],d['roofline']['frac'],w.get('ms_per_step',0),w.get('fp64_frac',0),d['roofline']['probe']['peak'],d['roofline']['probe']['dfma_per_sm_clk'],d['roofline']['probe']['sm_mhz_under_probe'],d['e2e']['ms_per_step']))"
done; done > gpurun_out/s2_ab.txt 2>&1; cat gpurun_out/s2_ab.txt
echo "== ncu v2"
PDA_TILED_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_euler3d_velocity_tiled2 -s 2 -c 1 -o gpurun_out/vel3d_v2_r02 -f python tools/profile_kernel.py --workload euler3d_weno5 --n 512 --reps 3 > gpurun_out/s2_ncu.log 2>&1; tail -3 gpurun_out/s2_ncu.log
echo "== parity report"
timeout 1500 python tools/parity_report_r02.py > gpurun_out/parity_report_r02.txt 2> gpurun_out/parity_report_r02.err; grep -E "WORST|cfg" gpurun_out/parity_report_r02.txt; tail -3 gpurun_out/parity_report_r02.err
