#!/bin/bash
# session 18 (1 GPU): full suite + smoke with the lane-per-face graph velocity kernel; A/B on cfg 4 / cfg 1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s18_tests.log; tail -6 gpurun_out/s18_tests.log
echo "== graph velocity A/B"
for f in 0 1; do
  PDA_VEL_ROWS_FACES=$f PDA_BENCH_ONLY=cfg4,cfg1 timeout 600 python tools/bench_configs.py 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin)
print('FACES=$f', ' | '.join('%s vel %.4f ms'%(k[:24],v['velocity']['ms']) for k,v in d.items()))"
done 2>&1 | tee gpurun_out/s18_vel_rows_ab.txt
