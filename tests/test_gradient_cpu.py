"""CPU side of the boundary-face gradient row (GradientEvaluator, gradient.hpp:61-121): the C restatement against the
fixtures generated from the unmodified reference (tests/golden/make_gradient_golden.py), the product's host logic
(graphRowsOfCellsStrictlyOnBd, the face records) through the C-ABI, its error behaviour and its refusal to compute
without a GPU.  No compute call runs here."""
import glob
import os

import numpy as np
import pytest

import pressiodemoapps as pda
from conftest import GOLDEN
from refdrv import have_ref_grad, oracle_gradient, ref_gradient, ref_rows_strictly_on_bd

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gradients", "*.npz")))
# the reference's own gold table (tests_cpp/gradients/main.cc:246-253), tolerance 1e-6 (main.cc:216-217)
REF_TEST_RMSE = {"fullmesh_s3": (0.118044, 0.082173), "samplemesh_s3": (0.134392, 0.0759626),
                 "fullmesh_s5": (0.0512737, 0.0298508), "samplemesh_s5": (0.0682185, 0.0307574),
                 "fullmesh_s7": (0.0512737, 0.0298508), "samplemesh_s7": (0.0682185, 0.0307574)}


def load(name):
    return np.load(os.path.join(GOLDEN, "gradients", name + ".npz"))


def mesh_of(g):
    return pda.mesh_from_arrays(2, int(g["stencil"]), g["d"], g["x"], g["y"], g["z"], g["graph"])


def test_fixture_inventory():
    assert len(CASES) == 8 and set(REF_TEST_RMSE) <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = load(name)
    for f, nd, key in ((g["f1"], 1, "grad1"), (g["f3"].ravel(), 3, "grad3")):
        o = oracle_gradient(int(g["stencil"]), g["graph"], g["rowsNearBd"], g["x"], g["y"], g["z"], g["d"][0], g["d"][1], f, nd)
        for k in ("cellGid", "position", "parentRow", "normalDir"):
            assert np.array_equal(o[k], g[k]), k
        assert np.array_equal(o["centers"], g["centers"])      # bit-exact
        assert np.array_equal(o["grad"], g[key])               # bit-exact


@pytest.mark.parametrize("name", sorted(REF_TEST_RMSE))
def test_reference_test_rmse_table(name):
    """the reference's own pass criterion (tests_cpp/gradients/main.cc:48-100, 216-217) applied to the oracle"""
    g = load(name)
    o = oracle_gradient(int(g["stencil"]), g["graph"], g["rowsNearBd"], g["x"], g["y"], g["z"], g["d"][0], g["d"][1], g["f1"], 1)
    cx, cy = o["centers"][:, 0], o["centers"][:, 1]
    isx = o["normalDir"] == 1
    gold = np.where(isx, cy, cx) * np.pi * np.cos(np.pi * cx * cy)
    e2 = (o["grad"][:, 0] - gold) ** 2
    n = o["grad"].shape[0]
    assert abs(np.sqrt(e2[isx].sum() / n) - REF_TEST_RMSE[name][0]) <= 1e-6
    assert abs(np.sqrt(e2[~isx].sum() / n) - REF_TEST_RMSE[name][1]) <= 1e-6


@pytest.mark.parametrize("name", CASES)
def test_host_face_records_match_reference(name):
    g = load(name)
    mesh = mesh_of(g)
    assert np.array_equal(mesh.graphRowsOfCellsStrictlyOnBd(), g["rowsStrictlyOnBd"])
    ev = pda.GradientEvaluator(mesh, 3)
    assert ev.numFaces() == g["cellGid"].size
    assert np.array_equal(ev.cellGIDs, g["cellGid"])
    assert np.array_equal(ev.positions, g["position"])
    assert np.array_equal(ev.parentRows, g["parentRow"])
    assert np.array_equal(ev.normalDirections, g["normalDir"])
    assert np.array_equal(ev.centers, g["centers"])           # bit-exact
    k = ev.numFaces() // 2
    f = ev.queryFace(int(g["cellGid"][k]), pda.FacePosition(int(g["position"][k])))
    assert np.array_equal(f.centerCoordinates, g["centers"][k]) and f.normalDirection == g["normalDir"][k]
    assert f.normalGradient.shape == (3,) and not f.normalGradient.any()   # zero until evaluated, like the reference


def test_native_lattice_matches_loaded_mesh():
    """the lattice descriptor (no stored graph) lists the same rows and faces as the reference's mesh files"""
    g = load("full_20x16_s5")
    mesh = pda.create_full_mesh([20, 16], [-1, 1, 0, 2], 5)
    assert np.array_equal(mesh.graphRowsOfCellsStrictlyOnBd(), g["rowsStrictlyOnBd"])
    ev = pda.GradientEvaluator(mesh)
    assert np.array_equal(ev.cellGIDs, g["cellGid"]) and np.array_equal(ev.positions, g["position"])
    assert np.array_equal(ev.centers, g["centers"])
    g = load("full_perx_14x9_s7")
    mesh = pda.create_full_mesh([14, 9], [0, 1, 0, 1], 7, ("x",))
    ev = pda.GradientEvaluator(mesh)
    assert np.array_equal(ev.cellGIDs, g["cellGid"]) and np.array_equal(ev.positions, g["position"])
    assert set(ev.positions.tolist()) == {int(pda.FacePosition.Front), int(pda.FacePosition.Back)}
    # fully periodic: no boundary face at all
    assert pda.GradientEvaluator(pda.create_full_mesh([8, 8], [0, 1, 0, 1], 3, ("x", "y"))).numFaces() == 0


def test_errors_mirror_reference():
    m1 = pda.create_full_mesh([20, 1], [0, 1], 3)
    m3 = pda.create_full_mesh([4, 4, 4], [0, 1, 0, 1, 0, 1], 3)
    for m in (m1, m3):
        with pytest.raises(pda.PdaError, match="gradients currently only supported for 2D"):   # gradient.hpp:71-73
            pda.GradientEvaluator(m)
        assert m.graphRowsOfCellsStrictlyOnBd().size == 0      # filled for 2D only, mesh_ccu.hpp:441-447
    m2 = pda.create_full_mesh([6, 5], [0, 1, 0, 1], 3)
    ev = pda.GradientEvaluator(m2, 2)
    with pytest.raises(pda.PdaError, match="numDofPerCell > MaxNumDofPerCell: 3 > 2"):   # gradient.hpp:87-91
        ev(np.zeros(m2.stencilMeshSize() * 3), 3)
    with pytest.raises(pda.PdaError, match="no such boundary face"):
        ev.queryFace(7, pda.FacePosition.Left)     # an inner cell
    with pytest.raises(ValueError):
        ev(np.zeros(5), 1)


def test_no_cpu_fallback():
    if pda.device_count() > 0:
        pytest.skip("a CUDA device is present")
    m = pda.create_full_mesh([6, 5], [0, 1, 0, 1], 3)
    ev = pda.GradientEvaluator(m)
    with pytest.raises(pda.PdaError) as e:
        ev(np.zeros(m.stencilMeshSize()))
    assert e.value.code == 3   # PDA_ERR_NO_DEVICE


@pytest.mark.skipif(not (have_ref_grad() and os.path.isdir("/root/reference")), reason="needs the compiled reference and its mesh files")
def test_fixtures_are_current():
    """the committed fixtures equal what the unmodified reference produces now (build container only)"""
    for name in REF_TEST_RMSE:
        g = load(name)
        d = os.path.join("/root/reference/tests_cpp/gradients", name)
        r = ref_gradient(d, g["f3"].ravel(), 3)
        assert np.array_equal(r["grad"], g["grad3"]) and np.array_equal(r["centers"], g["centers"])
        assert np.array_equal(ref_rows_strictly_on_bd(d), g["rowsStrictlyOnBd"])


@pytest.mark.parametrize("scheme,stencil", [("firstorder", 3), ("weno3", 5), ("weno5", 7)])
def test_oracle_reference_regression_initial_gradients(scheme, stencil):
    """tests_cpp/eigen_2d_euler_riemann_explicit_with_gradients: the t = 0 table (grad_gold_init.txt) from the oracle
    on the problem's initial condition, the reference's own criterion (compare.py: rtol = atol = 1e-8)"""
    gold = np.load(os.path.join(GOLDEN, "refgold", "gradients_riemann2d.npz"))["%s/init" % scheme]
    mesh = pda.create_full_mesh([22, 22], [0.0, 1.0, 0.0, 1.0], stencil)
    recon = {"firstorder": 0, "weno3": 1, "weno5": 2}[scheme]
    U = pda.create_problem(mesh, pda.Euler2d.Riemann, pda.InviscidFluxReconstruction(recon), 2).initialCondition()
    o = oracle_gradient(stencil, mesh.graph(), mesh.graphRowsOfCellsNearBd(), mesh.viewX(), mesh.viewY(), mesh.viewZ(),
                        mesh.dx(), mesh.dy(), U, 4)
    t = np.column_stack([o["centers"][:, 0], o["centers"][:, 1], o["grad"], o["normalDir"].astype(float)])
    t[0, 0], t[0, 1] = float("%.6g" % t[0, 0]), float("%.6g" % t[0, 1])
    assert t.shape == gold.shape and np.allclose(gold, t, rtol=1e-8, atol=1e-8)
