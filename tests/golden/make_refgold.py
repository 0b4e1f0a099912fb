"""Converts the gold files of the reference's own explicit-run regression tests (tests_cpp/*/[scheme/]*gold*.txt)
into ONE compressed fixture, tests/golden/refgold/refgold.npz, keyed "<case>/<scheme>/<check>" (tests/refgold_cases.py).
Run in the build container (needs /root/reference); the GPU box only sees the fixture.
    python tests/golden/make_refgold.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refgold_cases import CASES, gold_key   # noqa: E402

REF = "/root/reference/tests_cpp"
out = {}
for name, c in CASES.items():
    for scheme in c["schemes"]:
        for check, fname in c["checks"].items():
            if check == "rho_linf":
                continue   # a constant asserted by compare.py; lives in refgold_cases.py
            if c["subdirs"] is False:
                path = os.path.join(REF, c["ref_dir"], fname)
            elif c["subdirs"]:
                path = os.path.join(REF, c["ref_dir"], c["subdirs"].format(scheme=scheme), fname)
            else:
                path = os.path.join(REF, c["ref_dir"], scheme, fname)
            out[gold_key(name, scheme, check)] = np.loadtxt(path)
os.makedirs(os.path.join(HERE, "refgold"), exist_ok=True)
np.savez_compressed(os.path.join(HERE, "refgold", "refgold.npz"), **out)
print("wrote %d gold vectors, %d values" % (len(out), sum(v.size for v in out.values())))

# tests_cpp/eigen_2d_euler_riemann_explicit_with_gradients/{firstorder,weno3,weno5}/grad_gold_{init,final}.txt:
# one row per boundary face (x, y, 4 normal gradients, normal direction), compared by its compare.py at 1e-8
gout = {}
for scheme in ("firstorder", "weno3", "weno5"):
    for which in ("init", "final"):
        gout["%s/%s" % (scheme, which)] = np.loadtxt(os.path.join(
            REF, "eigen_2d_euler_riemann_explicit_with_gradients", scheme, "grad_gold_%s.txt" % which))
np.savez_compressed(os.path.join(HERE, "refgold", "gradients_riemann2d.npz"), **gout)
print("wrote %d gradient gold tables" % len(gout))

# the IMPLICIT-run tests (tests_cpp/*_implicit), tests/refgold_implicit.py
from refgold_implicit import CASES as ICASES, gold_key as igold_key   # noqa: E402
iout = {}
for name, c in ICASES.items():
    for scheme in c["schemes"]:
        for check, fname in c["checks"].items():
            if check == "rho_linf":
                continue
            path = os.path.join(REF, c["ref_dir"], fname) if c["subdirs"] is False else os.path.join(REF, c["ref_dir"], scheme, fname)
            iout[igold_key(name, scheme, check)] = np.loadtxt(path)
np.savez_compressed(os.path.join(HERE, "refgold", "refgold_implicit.npz"), **iout)
print("wrote %d implicit gold vectors, %d values" % (len(iout), sum(v.size for v in iout.values())))
