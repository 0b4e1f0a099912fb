"""The reference-order mode (pda_problem_set_option "jacobian_order"/"velocity_order" = "reference",
csrc/kernels_reforder.cu): velocity and Jacobian with the reference's formulas, operation order and accumulation
order, every operation individually rounded, std::pow reproduced bit for bit.

Acceptance here is the north-star tolerance against the UNMODIFIED reference's golden values, entry by entry, on
EVERY fixture -- no fallback rule, no field scaling.  (The default "fast" kernels are accepted by
tests/test_parity_gpu.py under assert_jacobian_parity, whose relaxed branch is a documented property of the reference's
own WENO-gradient rounding noise; this file is what makes "same results as the reference" checkable without it.)"""
import math

import numpy as np
import pytest

import pressiodemoapps as pda
from conftest import golden_names, oracle_arrays, scaled_err
from parity_cases import err_stats
from refdrv import OracleProblem
from test_host_cpu import make_mesh, make_problem

pytestmark = pytest.mark.gpu
R = pda.InviscidFluxReconstruction


def test_device_glibc_pow_is_bit_exact():
    """csrc/glibc_pow.h on the device == the host's libm pow, bit for bit, on the exponents the reference uses
    (std::pow(x, 3.) in the WENO gradients, std::pow(h, 0.5) in the shallow-water flux Jacobian) and two more"""
    rng = np.random.default_rng(5)
    x = np.exp(rng.uniform(math.log(1e-12), math.log(1e12), 200000))
    x[:1000] = 1.0 + (rng.uniform(-0.5, 0.5, 1000)) * 1e-3
    x[1000] = 1.0
    for y in (3.0, 0.5, 2.5, -1.75):
        got = pda._device_glibc_pow(x, y)
        ref = np.array([math.pow(v, y) for v in x])
        bad = np.count_nonzero(got != ref)
        assert bad == 0, "pow(x, %g): %d of %d results differ from libm" % (y, bad, x.size)


@pytest.mark.parametrize("name", golden_names())
def test_reference_order_matches_reference_golden(name, load_golden):
    """all fixtures of the unmodified reference: V, V from the Jacobian entry, J -- strict tolerance, no fallback"""
    g = load_golden(name)
    m = g.meta
    mesh, _ = make_mesh(g)
    p = make_problem(g, mesh)
    p.setOption("order", "reference")
    assert p.getOption("jacobian_order") == "reference" and p.getOption("velocity_order") == "reference"
    U, t = g["U"], m["t"]
    V = p.createRightHandSide()
    p.rightHandSide(U, t, V)
    assert scaled_err(V, g["V"]) <= 1.0
    J = p.createJacobian()
    assert np.array_equal(J.indptr, g["rowptr"]) and np.array_equal(J.indices, g["colidx"])
    V2 = p.createRightHandSide()
    p.rightHandSideAndJacobian(U, t, V2, J)
    assert scaled_err(V2, g["V2"]) <= 1.0
    s = scaled_err(J.data, g["Jv"])
    assert s <= 1.0, "reference-order Jacobian: scaled error %.3f vs the reference" % s
    # jacobian() alone, and a second evaluation, give the same bits
    J2 = p.createJacobian()
    p.jacobian(U, t, J2)
    assert np.array_equal(np.nan_to_num(J2.data), np.nan_to_num(J.data))
    # switching back restores the fast kernels (and they still agree to their own documented accuracy)
    p.setOption("order", "fast")
    V3 = p.createRightHandSide()
    p.rightHandSide(U, t, V3)
    assert scaled_err(V3, g["V"]) <= 1.0


@pytest.mark.parametrize("name", golden_names())
def test_reference_order_is_bitwise_the_reference(name, load_golden):
    """stronger than the tolerance: how many entries differ from the reference in ANY bit.  The reference's values
    come from x86-64 SSE2 arithmetic + glibc's pow; the kernels reproduce both, so the count is expected to be zero.
    -0.0 vs +0.0 is not counted (numpy compares them equal; the reference accumulates into +0.0 where the kernel may
    start from a -0.0 product)."""
    g = load_golden(name)
    m = g.meta
    mesh, _ = make_mesh(g)
    p = make_problem(g, mesh)
    p.setOption("order", "reference")
    U, t = g["U"], m["t"]
    V = p.createRightHandSide()
    p.rightHandSide(U, t, V)
    J = p.createJacobian()
    V2 = p.createRightHandSide()
    p.rightHandSideAndJacobian(U, t, V2, J)
    sv, sv2, sj = err_stats(V, g["V"]), err_stats(V2, g["V2"]), err_stats(J.data, g["Jv"])
    assert sv["nan_mismatch"] == 0 and sv2["nan_mismatch"] == 0 and sj["nan_mismatch"] == 0
    assert (sv["bits"], sv2["bits"], sj["bits"]) == (0, 0, 0), \
        "entries differing in any bit: V %d/%d, V2 %d/%d, J %d/%d (max scaled %.3g)" % (
            sv["bits"], sv["n"], sv2["bits"], sv2["n"], sj["bits"], sj["n"], sj["strict"])


@pytest.mark.parametrize("n", [(12, 10, 9), (7, 7, 7)])
def test_reference_order_3d_weno5_extension(n):
    """3D WENO5 (no reference exists): reference-order kernels vs the oracle's restatement, strict"""
    mesh = pda.create_full_mesh(list(n), [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5)
    p.setOption("order", "reference")
    x, y, z = mesh._coords()
    o = OracleProblem(None, "euler3d", 0, 2, arrays=dict(dim=3, stencil=7, d=mesh._deltas()[0], graph=mesh.graph(), x=x, y=y, z=z))
    rng = np.random.default_rng(20261017)
    U = p.initialCondition()
    U *= 1 + 1e-3 * rng.uniform(-1, 1, U.size)
    V = p.createRightHandSide()
    p.rightHandSide(U, 0.0, V)
    assert scaled_err(V, o.velocity(U, 0.0)) <= 1.0
    J = p.createJacobian()
    V2 = p.createRightHandSide()
    p.rightHandSideAndJacobian(U, 0.0, V2, J)
    Vo, Jo = o.velocityAndJacobian(U, 0.0)
    assert scaled_err(V2, Vo) <= 1.0 and scaled_err(J.data, Jo) <= 1.0


def test_reference_order_apply_jacobian_and_options():
    mesh = pda.create_full_mesh([30, 24], [0, 1, 0, 1], 7)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    with pytest.raises(pda.PdaError):
        p.setOption("jacobian_order", "sloppy")
    with pytest.raises(pda.PdaError):
        p.setOption("no_such_option", "1")
    p.setOption("jacobian_order", "reference")
    rng = np.random.default_rng(3)
    U = p.initialCondition() * (1 + 1e-3 * rng.uniform(-1, 1, p.totalDofStencilMesh()))
    J = p.createJacobian()
    p.jacobian(U, 0.0, J)
    B = rng.uniform(-1, 1, (U.size, 3))
    Rm = p.createApplyJacobianResult(B)
    p.applyJacobian(U, B, 0.0, Rm)   # multiplies the reference-order Jacobian (no matrix-free inner rows in this mode)
    assert scaled_err(Rm, J @ B, 1e-11, 1e-9) <= 1.0
