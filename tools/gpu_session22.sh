#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_reforder_gpu.py -q -m gpu -k "dmr or riemann or normalshock or crossshock or kh or sedov" 2>&1 | tail -8
python - <<'PY' 2>&1 | grep -v Warning
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, "pressio-demoapps_b200")
import numpy as np, glob, os
import pressiodemoapps as pda
from conftest import scaled_err
from parity_cases import *  # noqa
PY
python tools/time_2d.py euler weno5 2048 2>&1 | grep -v Warning
