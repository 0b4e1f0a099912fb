#!/bin/bash
# session 10 (1 GPU): GPU suite with the WENO march Jacobian kernel, A/B against the tile kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s10_tests.log; tail -8 gpurun_out/s10_tests.log
echo "== A/B"
( for c in "swe weno3 4096" "swe weno5 4096" "burgers weno3 4096" "burgers weno5 4096"; do
    PDA_JAC_WENO_MARCH=0 python tools/time_2d.py $c
    python tools/time_2d.py $c
  done ) 2>&1 | grep -v Warning | tee gpurun_out/s10_ab.txt
echo "== ncu"
cap() {  # name kernel-regex skip workload n reps
  timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/profile_kernel.py --workload $4 --n $5 --reps $6 > /tmp/$1.log 2>&1
  if [ -f /tmp/$1.ncu-rep ]; then
    python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
    ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
    cat gpurun_out/ncu_$1.txt | cut -c1-200
  else tail -3 /tmp/$1.log; fi
}
cap swe_weno3_jac_march_r02 k_jacobian_march2d_weno 1 swe_weno3_jac 4096 2
du -sh gpurun_out
