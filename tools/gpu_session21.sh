#!/bin/bash
# session 21 (1 GPU): cell-based WENO edges in the 2D velocity march: full suite + timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s21_tests.log; tail -6 gpurun_out/s21_tests.log
( python tools/time_2d.py euler weno5 2048; python tools/time_2d.py euler weno3 2048; python tools/time_2d.py swe weno3 4096; python tools/time_2d.py swe weno5 4096; python tools/time_2d.py burgers weno5 4096 ) 2>&1 | grep -v Warning | tee gpurun_out/s21_vel2d.txt
