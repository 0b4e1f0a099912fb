"""The reference's implicit-run regression tests (tests_cpp/*_implicit) reproduced with the ORACLE's velocity and
Jacobian inside the restated BDF1 / Crank-Nicolson + Newton loop (tests/refgold_implicit.py): pins the oracle's Jacobian
end to end against gold files the reference holds itself, and validates the stepper restatement the GPU twin
(tests/test_refgold_implicit_gpu.py) uses.  CPU only."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import pressiodemoapps as pda
from refdrv import OracleProblem
from refgold_implicit import CASES, SCHEMES, advance_implicit, check_against_gold, params

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgold", "refgold_implicit.npz"))
FAMILY = {"Euler1d": "euler1d", "Euler2d": "euler2d", "Swe2d": "swe2d", "DiffusionReaction2d": "diffreac2d",
          "AdvectionDiffusion2d": "advdiff2d", "AdvectionDiffusionReaction2d": "advdiffreac2d", "Advection1d": "advection1d",
          "DiffusionReaction1d": "diffreac1d"}


@pytest.mark.parametrize("name,scheme", params())
def test_oracle_reproduces_reference_implicit_gold(name, scheme):
    c = CASES[name]
    recon_name, stencil = SCHEMES[scheme]
    recon = int(getattr(pda.InviscidFluxReconstruction, recon_name))
    if c["enum"][0].startswith("DiffusionReaction"):
        stencil, recon = 3, 0
    mesh = pda.create_full_mesh(c["n"], c["bounds"], stencil, c["periodic"])
    x, y, z = mesh._coords()
    arrays = dict(dim=mesh.dimensionality(), stencil=stencil, d=mesh._deltas()[0], graph=mesh.graph(), x=x, y=y, z=z)
    prob = int(getattr(getattr(pda, c["enum"][0]), c["enum"][1]))
    prm = {"diffusion": 0.01, "reaction": 0.005} if c["factory"] == "diffreac1d_A" else None
    o = OracleProblem(None, FAMILY[c["enum"][0]], prob, recon, icFlag=c["ic"], params=prm, arrays=arrays)
    rp, ci = o.pattern()
    n = o.nDofStencil

    def rhs_and_jac(U, t):
        V, vals = o.velocityAndJacobian(U, t)
        return V, sp.csr_matrix((vals, ci, rp), shape=(n, n))
    U, iters = advance_implicit(rhs_and_jac, o.velocity, o.initialCondition(), c["ode"], c["dt"], c["nsteps"], c["tol"])
    assert max(iters) < 100 and not np.isnan(U).any()
    check_against_gold(c, scheme, U, (x, y), GOLD, name)
