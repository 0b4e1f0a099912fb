"""The reference's IMPLICIT end-to-end regression tests (tests_cpp/*_implicit, 11 directories / 17 scheme variants) on
the GPU engine: BDF1 / Crank-Nicolson + Newton as their main.cc drive them (restated in tests/refgold_implicit.py from
the vendored pressio sources), every residual AND every Jacobian evaluated by the CUDA path through the C-ABI
(rightHandSideAndJacobian), the linear systems solved on the host (scipy sparse LU instead of Eigen's BiCGSTAB), and
the reference's own compare.py criterion applied against its own gold files.  These are the only reference-held goldens
that exercise the Jacobian end to end.  Both Jacobian modes are run: the default fast kernels and the reference-order
mode."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import pressiodemoapps as pda
from refgold_implicit import CASES, SCHEMES, advance_implicit, check_against_gold, params

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgold", "refgold_implicit.npz"))


def _make_problem(c, mesh, recon):
    enum = getattr(getattr(pda, c["enum"][0]), c["enum"][1])
    fam = c["enum"][0]
    if c["factory"] == "diffreac1d_A":   # create_diffusion_reaction_1d_problem_A_eigen(meshObj, 0.01, 0.005)
        return pda.create_diffusion_reaction_1d_problem_A(mesh, 0.01, 0.005)
    if fam == "AdvectionDiffusion2d":
        return pda.create_problem(mesh, enum, recon, pda.ViscousFluxReconstruction.FirstOrder)
    if fam in ("Euler2d", "Swe2d"):
        return pda.create_problem(mesh, enum, recon, c["ic"])
    return pda.create_problem(mesh, enum, recon)


@pytest.mark.parametrize("order", ["fast", "reference"])
@pytest.mark.parametrize("name,scheme", params())
def test_reference_implicit_regression_on_gpu(name, scheme, order):
    c = CASES[name]
    recon_name, stencil = SCHEMES[scheme]
    recon = getattr(pda.InviscidFluxReconstruction, recon_name)
    if c["enum"][0].startswith("DiffusionReaction"):
        stencil = 3
    mesh = pda.create_full_mesh(c["n"], c["bounds"], stencil, c["periodic"])
    p = _make_problem(c, mesh, recon)
    p.setOption("order", order)
    J = p.createJacobian()
    n = p.totalDofStencilMesh()

    def rhs(U, t):
        V = p.createRightHandSide()
        p.rightHandSide(U, t, V)
        return V

    def rhs_and_jac(U, t):
        V = p.createRightHandSide()
        p.rightHandSideAndJacobian(U, t, V, J)
        return V, sp.csr_matrix((J.data.copy(), J.indices, J.indptr), shape=(n, n))
    U, iters = advance_implicit(rhs_and_jac, rhs, p.initialCondition(), c["ode"], c["dt"], c["nsteps"], c["tol"])
    assert max(iters) < 100 and not np.isnan(U).any()
    x, y, _ = mesh._coords()
    check_against_gold(c, scheme, U, (x, y), GOLD, name)
