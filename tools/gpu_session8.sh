#!/bin/bash
# session 8 (1 GPU): GPU suite, A/B of the first-order Jacobian march kernel and the 2D march occupancy, ncu of the new kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s8_tests.log; tail -8 gpurun_out/s8_tests.log
echo "== A/B"
( PDA_JAC_FO_MARCH=0 python tools/time_2d.py swe fo 4096
  python tools/time_2d.py swe fo 4096
  PDA_JAC_FO_CTAS=3 python tools/time_2d.py swe fo 4096
  python tools/time_2d.py swe weno3 4096
  PDA_JAC_FO_MARCH=0 python tools/time_2d.py burgers fo 4096
  python tools/time_2d.py burgers fo 4096
  PDA_JAC_FO_MARCH=0 python tools/time_2d.py euler fo 2048
  python tools/time_2d.py euler fo 2048
  python tools/time_2d.py euler weno5 2048 ) 2>&1 | grep -v Warning | tee gpurun_out/s8_ab.txt
echo "== ncu"
cap() {  # name kernel-regex skip workload n reps
  timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/profile_kernel.py --workload $4 --n $5 --reps $6 > /tmp/$1.log 2>&1
  if [ -f /tmp/$1.ncu-rep ]; then
    python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/ncu_$1.txt 2>&1
    ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
    ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
    cat gpurun_out/ncu_$1.txt | cut -c1-200
  else tail -3 /tmp/$1.log; fi
}
cap swe_fo_jac_march_r02 k_jacobian_march2d_fo 1 swe_fo_jac 4096 2
cap swe_fo_vel_occ_r02 k_velocity_march2d 2 swe_fo_vel 4096 3
du -sh gpurun_out
