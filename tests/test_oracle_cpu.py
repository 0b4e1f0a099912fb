"""Pins the oracle (oracle/pda_oracle.c, the plain-C restatement of the reference's hot path) against
  (a) the known-answer vectors in the reference's own unit tests, and
  (b) the committed golden fixtures tests/golden/*.npz, which were produced by the UNMODIFIED reference
      (tests/golden/make_golden.py: reference mesh scripts + oracle/_ref/libpda_ref.so).
No GPU involved."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_names, oracle_arrays, scaled_err
from refdrv import OracleProblem, oracle_leaf, have_ref, RefProblem


def test_weno5_known_answer():
    # /root/reference/tests_cpp/weno5/main.cc:7-33: stencil {8,5,0,1,2,1,4}, tol 1e-13
    L = oracle_leaf()
    q = [8., 5., 0., 1., 2., 1., 4.]
    a, b = C.c_double(), C.c_double()
    L.or_weno5(C.byref(a), C.byref(b), *q[0:6])
    uMinusHalfNeg, uMinusHalfPos = a.value, b.value
    L.or_weno5(C.byref(a), C.byref(b), *q[1:7])
    uPlusHalfNeg, uPlusHalfPos = a.value, b.value
    gold = [0.498169371091894, 0.498239224509689, 1.50249211669906, 1.53308488724067]
    got = [uMinusHalfNeg, uMinusHalfPos, uPlusHalfNeg, uPlusHalfPos]
    assert np.max(np.abs(np.array(got) - np.array(gold))) < 1e-13


def test_weno3_known_answer():
    # /root/reference/tests_cpp/weno3/main.cc: stencil {8,5,0,1,2}: Jiang-Shu weights recomputed independently
    L = oracle_leaf()

    def ref_weno3(qm1, q, qp1, qp2):
        eps = 1e-6
        b0, b1 = (q - qm1) ** 2, (qp1 - q) ** 2
        a0, a1 = (1. / 3) / (eps + b0) ** 2, (2. / 3) / (eps + b1) ** 2
        neg = (a0 * (-0.5 * qm1 + 1.5 * q) + a1 * 0.5 * (q + qp1)) / (a0 + a1)
        b0, b1 = (qp1 - q) ** 2, (qp2 - qp1) ** 2
        a0, a1 = (2. / 3) / (eps + b0) ** 2, (1. / 3) / (eps + b1) ** 2
        pos = (a0 * 0.5 * (q + qp1) + a1 * (1.5 * qp1 - 0.5 * qp2)) / (a0 + a1)
        return neg, pos
    a, b = C.c_double(), C.c_double()
    for st in ([8., 5., 0., 1.], [5., 0., 1., 2.], [1.0, 1.1, 1.3, 1.2]):
        L.or_weno3(C.byref(a), C.byref(b), *st)
        n, p = ref_weno3(*st)
        assert abs(a.value - n) < 1e-13 and abs(b.value - p) < 1e-13


def _complex_step_weno5(q):
    """weno5 with complex arithmetic (the check of /root/reference/tests_cpp/weno5/main_jacobians.cc)"""
    eps = 1e-6
    a, b, c, d, e, f = q

    def side(v0, v1, v2, v3, v4, cw, polys):
        B0 = 13. / 12 * (v0 - 2 * v1 + v2) ** 2 + 0.25 * (v0 - 4 * v1 + 3 * v2) ** 2
        B1 = 13. / 12 * (v1 - 2 * v2 + v3) ** 2 + 0.25 * (v1 - v3) ** 2
        B2 = 13. / 12 * (v2 - 2 * v3 + v4) ** 2 + 0.25 * (3 * v2 - 4 * v3 + v4) ** 2
        al = [cw[0] / (eps + B0) ** 2, cw[1] / (eps + B1) ** 2, cw[2] / (eps + B2) ** 2]
        s = sum(al)
        return sum(al[k] / s * polys[k] for k in range(3))
    neg = side(a, b, c, d, e, (0.1, 0.6, 0.3),
               ((2 * a - 7 * b + 11 * c) / 6, (-b + 5 * c + 2 * d) / 6, (2 * c + 5 * d - e) / 6))
    pos = side(b, c, d, e, f, (0.3, 0.6, 0.1),
               ((-b + 5 * c + 2 * d) / 6, (2 * c + 5 * d - e) / 6, (11 * d - 7 * e + 2 * f) / 6))
    return neg, pos


def test_weno5_gradients_complex_step():
    # tests_cpp/weno5/main_jacobians.cc: analytic gradients vs complex step, tol 1e-12
    L = oracle_leaf()
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rng.uniform(0.5, 2.0, 6)
        gN = (C.c_double * 6)()
        gP = (C.c_double * 6)()
        a, b = C.c_double(), C.c_double()
        L.or_weno5_grad(C.byref(a), C.byref(b), gN, gP, *q)
        for m in range(6):
            qc = q.astype(complex)
            qc[m] += 1e-30j
            n, p = _complex_step_weno5(qc)
            assert abs(gN[m] - n.imag / 1e-30) < 1e-11
            assert abs(gP[m] - p.imag / 1e-30) < 1e-11


def test_euler_flux_jacobian_fd():
    # tests_cpp/eigen_rusanov_flux_jacobians_euler/main{1d,2d,3d}.cc: flux Jacobians vs finite differences, tol 1e-4
    L = oracle_leaf()
    rng = np.random.default_rng(2)
    for ndpc in (3, 4, 5):
        dim = ndpc - 2
        for ax in range(dim):
            n = np.zeros(3)
            n[ax] = 1.0

            def rand_state():
                rho = rng.uniform(0.5, 2.0)
                vel = rng.uniform(-1, 1, dim)
                p = rng.uniform(0.5, 2.0)
                return np.concatenate([[rho], rho * vel, [p / 0.4 + 0.5 * rho * vel @ vel]])
            qL, qR = rand_state(), rand_state()
            JL, JR = np.zeros((ndpc, ndpc)), np.zeros((ndpc, ndpc))
            L.or_euler_flux_jac(ndpc, JL.ctypes.data, JR.ctypes.data, qL.ctypes.data, qR.ctypes.data, n.ctypes.data, 1.4)

            def F(l, r):
                out = np.zeros(ndpc)
                L.or_euler_flux(ndpc, out.ctypes.data, l.ctypes.data, r.ctypes.data, n.ctypes.data, 1.4)
                return out
            h = 1e-7
            for j in range(ndpc):
                e = np.zeros(ndpc)
                e[j] = h
                assert np.max(np.abs((F(qL + e, qR) - F(qL - e, qR)) / (2 * h) - JL[:, j])) < 1e-5
                assert np.max(np.abs((F(qL, qR + e) - F(qL, qR - e)) / (2 * h) - JR[:, j])) < 1e-5


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name, load_golden):
    """mesh classification, initial condition, velocity, CSR pattern, Jacobian values and ghosts of the C restatement
    equal the reference's (fixtures).  Pattern/indices bit-exact; floating point to 1e-12/1e-10 (observed: bit-exact
    or last-ulp, the restatement keeps the reference's operation order)."""
    g = load_golden(name)
    m = g.meta
    prm = {k: v for k, v in (m["params"] or {}).items() if k != "testSource"}
    o = OracleProblem(None, m["family"], m["prob"], m["recon"], m["ic"], prm, arrays=oracle_arrays(g))
    if "src" in g:
        o.setSource(g["src"])
    assert (o.nSample, o.nStencil, o.ndpc, o.nnz) == (g["graph"].shape[0], g["x"].size, m["ndpc"], m["nnz"])
    ma = o.mesh_arrays()
    assert np.array_equal(ma["rowsInner"], g["rowsInner"])
    assert np.array_equal(ma["rowsNearBd"], g["rowsNearBd"])
    assert np.array_equal(o.initialCondition(), g["IC"])
    rp, ci = o.pattern()
    assert np.array_equal(rp, g["rowptr"]) and np.array_equal(ci, g["colidx"])
    U, t = g["U"], m["t"]
    assert scaled_err(o.velocity(U, t), g["V"]) <= 1.0
    V2, Jv = o.velocityAndJacobian(U, t)
    assert scaled_err(V2, g["V2"]) <= 1.0
    assert scaled_err(Jv, g["Jv"]) <= 1.0
    for s in range(4):
        if "ghost%d" % s in g:
            got = o.ghosts(s)
            ref = g["ghost%d" % s]
            # rows the reference never writes keep its numeric_limits<double>::min() initialiser
            w = ref != np.finfo(np.float64).tiny
            assert np.array_equal(got.reshape(ref.shape)[w], ref[w])


@pytest.mark.skipif(not have_ref(), reason="compiled reference not present")
def test_golden_regenerates_from_live_reference(load_golden, tmp_path):
    """spot check that the fixtures are what the compiled reference produces now (guards stale fixtures)"""
    import pressiodemoapps as pda
    g = load_golden("swe_slipwall_weno3_25")
    m = g.meta
    mesh = pda.create_full_mesh(m["n"], m["bounds"], m["stencil"], m["periodic"])
    mesh.write(str(tmp_path))
    r = RefProblem(str(tmp_path), m["family"], m["prob"], m["recon"], m["ic"], m["params"])
    assert np.array_equal(r.velocity(g["U"], m["t"]), g["V"])


@pytest.mark.parametrize("fam,prob,recon,n,bounds,per,sten", [
    ("euler3d", 0, 2, [9, 8, 7], [-1, 1, -1, 1, -1, 1], ("x", "y", "z"), 7),
    ("euler3d", 1, 1, [10, 9, 8], [0, 1, 0, 1, 0, 1], (), 5),
    ("euler2d", 6, 1, [40, 12], [0, 4, 0, 1], (), 5),
    ("euler2d", 4, 2, [20, 18], [0, 1, 0, 1], (), 7),
    ("swe2d", 0, 1, [17, 19], [-5, 5, -5, 5], (), 5),
    ("euler1d", 1, 2, [50, 1], [-0.5, 0.5], (), 7),
    ("diffreac2d", 1, 0, [16, 12], [-1.25, 1.25, -1.25, 1.25], ("x", "y"), 3),
    ("advdiff2d", 1, 1, [14, 15], [-1, 1, -1, 1], (), 5),
])
def test_lattice_mode_equals_stored_graph_mode(fam, prob, recon, n, bounds, per, sten):
    """or_create_lattice (connectivity by index arithmetic, what makes 512^3 / 4096^2 checkable) gives the SAME bits as
    the stored-graph oracle that the golden fixtures pin: classification, initial condition, velocity at two times,
    Jacobian values and pattern."""
    import pressiodemoapps as pda
    from refdrv import lattice_spec
    mesh = pda.create_full_mesh(n, bounds, sten, per)
    x, y, z = mesh._coords()
    oa = OracleProblem(None, fam, prob, recon, arrays=dict(dim=mesh.dimensionality(), stencil=sten, d=mesh._deltas()[0],
                                                           graph=mesh.graph(), x=x, y=y, z=z))
    spec = lattice_spec(n, bounds, sten, per)
    assert np.array_equal(np.array(spec["d"]), mesh._deltas()[0][:spec["dim"]])
    for a, c in zip(("cx", "cy", "cz"), (x, y, z)):
        if spec[a] is not None:
            assert set(spec[a]) == set(np.unique(c))
    ol = OracleProblem(None, fam, prob, recon, lattice=spec)
    assert (oa.nSample, oa.nInner, oa.nNearBd, oa.periodic) == (ol.nSample, ol.nInner, ol.nNearBd, ol.periodic)
    U = oa.initialCondition()
    assert np.array_equal(U, ol.initialCondition())
    rng = np.random.default_rng(1)
    U = U * (1 + 1e-3 * rng.uniform(-1, 1, U.size)) if np.any(U) else 0.1 * rng.uniform(-1, 1, U.size)
    for t in (0.0, 0.07):
        assert np.array_equal(oa.velocity(U, t), ol.velocity(U, t), equal_nan=True)
    Va, Ja = oa.velocityAndJacobian(U, 0.0)
    Vl, Jl = ol.velocityAndJacobian(U, 0.0)
    assert np.array_equal(Va, Vl, equal_nan=True) and np.array_equal(Ja, Jl, equal_nan=True)
    for a, b in zip(oa.pattern(), ol.pattern()):
        assert np.array_equal(a, b)
