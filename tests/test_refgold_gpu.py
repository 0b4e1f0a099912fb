"""The reference's OWN end-to-end regression tests, run on the GPU engine.

Every tests_cpp/*_explicit test of the reference integrates a problem in time (pressio RK4 / SSPRK3) and compares the
final state (or density / pressure / depth, or the L-inf error against an analytic solution) with a gold file, using
np.allclose with the tolerance written in its compare.py.  Here the same runs are done with the state resident in HBM
(pda_problem_advance_dev/_host: every evaluation is the CUDA velocity path, the stage arithmetic that of pressio's
steppers) and the SAME acceptance criterion is applied against the same gold values (tests/golden/refgold/refgold.npz,
made by tests/golden/make_refgold.py from the reference's files).  tests/refgold_cases.py cites each test."""
import os

import numpy as np
import pytest

import pressiodemoapps as pda
from refgold_cases import CASES, SCHEMES, gold_key

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgold", "refgold.npz"))
GAMMA = 1.4


def _params():
    return [(name, scheme) for name, c in CASES.items() for scheme in c["schemes"]]


def _make_problem(c, mesh, recon):
    enum = getattr(getattr(pda, c["enum"][0]), c["enum"][1])
    fam = c["enum"][0]
    if c["factory"] == "cross_shock":   # create_cross_shock_problem_eigen(meshObj, recEnum) -- defaults 0.1, 10, 1
        return pda.create_cross_shock_problem(mesh, recon, 0.1, 10.0, 1.0)
    if c["factory"] == "diffreac1d_A":  # create_diffusion_reaction_1d_problem_A_eigen(meshObj, 0.01, 0.005)
        return pda.create_diffusion_reaction_1d_problem_A(mesh, 0.01, 0.005)
    if fam in ("DiffusionReaction2d",):
        return pda.create_problem(mesh, enum, pda.ViscousFluxReconstruction.FirstOrder)
    if fam == "AdvectionDiffusion2d":
        return pda.create_problem(mesh, enum, recon, pda.ViscousFluxReconstruction.FirstOrder)
    if fam in ("Euler2d", "Swe2d"):
        return pda.create_problem(mesh, enum, recon, c["ic"])
    return pda.create_problem(mesh, enum, recon)


@pytest.mark.parametrize("name,scheme", _params())
def test_reference_regression_on_gpu(name, scheme):
    c = CASES[name]
    recon_name, stencil = SCHEMES[scheme]
    recon = getattr(pda.InviscidFluxReconstruction, recon_name)
    mesh = pda.create_full_mesh(c["n"], c["bounds"], stencil, c["periodic"])
    p = _make_problem(c, mesh, recon)
    U = p.initialCondition()
    if "state@100+150" in c["checks"]:
        p.advance(c["stepper"], U, c["dt"], 100, 0.0)
        U100 = U.copy()
        p.advance(c["stepper"], U, c["dt"], 50, 100 * c["dt"])
    else:
        p.advance(c["stepper"], U, c["dt"], c["nsteps"], 0.0)
    assert not np.isnan(U).any()
    ndpc = p.numDofPerCell()
    cells = U.reshape(-1, ndpc)
    for check, what in c["checks"].items():
        if check == "rho_linf":
            x, y, z = mesh._coords()
            t_end = c["dt"] * c["nsteps"]
            dim = len(c["n"])
            s = x + y + (z if dim == 3 else 0.0)
            exact = 1.0 + 0.2 * np.sin(np.pi * (s - dim * t_end))
            err = float(np.max(np.abs(cells[:, 0] - exact)))
            ref = what[scheme]
            assert abs(err - ref) <= 1e-9 * max(abs(err), abs(ref)), (err, ref)   # math.isclose(err, ref)
            continue
        gold = GOLD[gold_key(name, scheme, check)]
        if check == "state":
            got = U
        elif check == "state@100+150":
            got = np.concatenate([U100, U])
        elif check in ("rho", "h"):
            got = cells[:, 0]
        else:   # pressure, as the compare.py scripts compute it
            rho = cells[:, 0]
            vel2 = sum((cells[:, 1 + m] / rho) ** 2 for m in range(ndpc - 2))
            got = (GAMMA - 1.0) * (cells[:, ndpc - 1] - rho * vel2 * 0.5)
        assert got.shape == gold.shape
        assert np.allclose(got, gold, rtol=c["rtol"], atol=c["atol"]), \
            "%s/%s/%s: max scaled err %.3g" % (name, scheme, check,
                                               float(np.max(np.abs(got - gold) / (c["atol"] + c["rtol"] * np.abs(gold)))))
