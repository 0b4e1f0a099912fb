"""Pins the C oracle (oracle/pda_oracle.c) against the gold files of the reference's own explicit-run regression tests
(tests/golden/refgold/refgold.npz <- tests_cpp/*/gold*.txt): the oracle's velocity drives numpy restatements of
pressio's RK4 / SSPRK3 stage arithmetic and the result must pass the reference's compare.py criterion.  CPU only."""
import os

import numpy as np
import pytest

import pressiodemoapps as pda
from refdrv import OracleProblem
from refgold_cases import CASES, SCHEMES, gold_key

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgold", "refgold.npz"))
FAMILY = {"Euler1d": "euler1d", "Euler2d": "euler2d", "Euler3d": "euler3d", "Swe2d": "swe2d",
          "DiffusionReaction2d": "diffreac2d", "AdvectionDiffusion2d": "advdiff2d",
          "AdvectionDiffusionReaction2d": "advdiffreac2d", "Advection1d": "advection1d", "DiffusionReaction1d": "diffreac1d"}
# bounded CPU time: the long 3D / 2000-step runs are covered on the GPU (tests/test_refgold_gpu.py)
CPU_CASES = ["sod1d", "lax1d", "shuosher1d", "advection1d", "riemann2d", "sedov2d", "normalshock2d", "dmr2d", "kh2d",
             "swe2d", "burgers2d", "smooth3d"]


def _step(o, stepper, U, t, dt):
    f = o.velocity
    if stepper == "rk4":   # ode_explicit_stepper_without_mass_matrix.hpp:284-340
        k1 = f(U, t)
        k2 = f(U + (dt / 2) * k1, t + dt / 2)
        k3 = f(U + (dt / 2) * k2, t + dt / 2)
        k4 = f(U + dt * k3, t + dt)
        return U + (dt / 6) * k1 + (dt / 3) * k2 + (dt / 3) * k3 + (dt / 6) * k4
    k = f(U, t)            # SSPRK3 :230-281
    u1 = U + dt * k
    k = f(u1, t + dt)
    u2 = 0.25 * u1 + 0.75 * U + (0.25 * dt) * k
    k = f(u2, t + dt / 2)
    return (1.0 / 3.0) * U + (2.0 / 3.0) * u2 + ((2.0 / 3.0) * dt) * k


@pytest.mark.parametrize("name,scheme", [(n, s) for n in CPU_CASES for s in CASES[n]["schemes"]])
def test_oracle_reproduces_reference_gold(name, scheme):
    c = CASES[name]
    recon_name, stencil = SCHEMES[scheme]
    recon = int(getattr(pda.InviscidFluxReconstruction, recon_name))
    mesh = pda.create_full_mesh(c["n"], c["bounds"], stencil, c["periodic"])
    x, y, z = mesh._coords()
    arrays = dict(dim=mesh.dimensionality(), stencil=stencil, d=mesh._deltas()[0], graph=mesh.graph(), x=x, y=y, z=z)
    prob = int(getattr(getattr(pda, c["enum"][0]), c["enum"][1]))
    o = OracleProblem(None, FAMILY[c["enum"][0]], prob, recon, icFlag=c["ic"], arrays=arrays)
    U = o.initialCondition()
    t, snaps = 0.0, {}
    for s in range(c["nsteps"]):
        U = _step(o, c["stepper"], U, t, c["dt"])
        t += c["dt"]
        if s + 1 == 100:
            snaps[100] = U.copy()
    ndpc = o.ndpc
    cells = U.reshape(-1, ndpc)
    for check, what in c["checks"].items():
        if check == "rho_linf":
            dim = len(c["n"])
            ssum = x + y + (z if dim == 3 else 0.0)
            err = float(np.max(np.abs(cells[:, 0] - (1.0 + 0.2 * np.sin(np.pi * (ssum - dim * c["dt"] * c["nsteps"]))))))
            assert abs(err - what[scheme]) <= 1e-9 * max(abs(err), abs(what[scheme]))
            continue
        gold = GOLD[gold_key(name, scheme, check)]
        if check == "state":
            got = U
        elif check == "state@100+150":
            got = np.concatenate([snaps[100], U])
        elif check in ("rho", "h"):
            got = cells[:, 0]
        else:
            rho = cells[:, 0]
            vel2 = sum((cells[:, 1 + m] / rho) ** 2 for m in range(ndpc - 2))
            got = (1.4 - 1.0) * (cells[:, ndpc - 1] - rho * vel2 * 0.5)
        assert np.allclose(got, gold, rtol=c["rtol"], atol=c["atol"])
