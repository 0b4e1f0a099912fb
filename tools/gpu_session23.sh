#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for m in 0 1 2 3; do
  echo "== PDA_MARCH2D_CELL=$m"
  PDA_MARCH2D_CELL=$m timeout 300 python - <<'PY' 2>&1 | grep -v Warning
import sys, glob, os
sys.path.insert(0, "tests"); sys.path.insert(0, "pressio-demoapps_b200")
import numpy as np
import pressiodemoapps as pda
from conftest import scaled_err
import conftest
from test_host_cpu import make_mesh, make_problem
worst = []
for f in sorted(glob.glob("tests/golden/euler2d_*.npz")):
    name = os.path.basename(f)[:-4]
    g = conftest.Golden(name)
    mesh, _ = make_mesh(g); p = make_problem(g, mesh)
    V = p.createRightHandSide(); p.rightHandSide(g["U"], g.meta["t"], V)
    worst.append((scaled_err(V, g["V"]), name))
worst.sort(reverse=True)
print(worst[:4])
PY
  PDA_MARCH2D_CELL=$m python tools/time_2d.py euler weno5 2048 2>&1 | grep -v Warning
done
