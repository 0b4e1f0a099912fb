"""GPU parity of the boundary-face gradient kernel (k_boundary_face_normal_gradient, csrc/gradient.cu) through the
C-ABI: bit-exact against the fixtures generated from the unmodified reference's GradientEvaluator and against the C
restatement on larger lattices; host- and device-pointer entry points."""
import glob
import os

import numpy as np
import pytest

import pressiodemoapps as pda
from conftest import GOLDEN
from refdrv import oracle_gradient

pytestmark = pytest.mark.gpu
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gradients", "*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN, "gradients", name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_golden(name):
    g = load(name)
    mesh = pda.mesh_from_arrays(2, int(g["stencil"]), g["d"], g["x"], g["y"], g["z"], g["graph"])
    ev1 = pda.GradientEvaluator(mesh)                  # one dof per cell: scalar normalGradient
    out = ev1(g["f1"])
    assert np.array_equal(out, g["grad1"])             # bit-exact
    k = out.shape[0] - 1
    face = ev1.queryFace(int(g["cellGid"][k]), int(g["position"][k]))
    assert isinstance(face.normalGradient, float) and face.normalGradient == g["grad1"][k, 0]
    ev3 = pda.GradientEvaluator(mesh, 5)               # MaxNumDofPerCell 5, called with 3 (and then with 1)
    assert np.array_equal(ev3(g["f3"].ravel(), 3), g["grad3"])
    face = ev3.queryFace(int(g["cellGid"][0]), int(g["position"][0]))
    assert np.array_equal(face.normalGradient[:3], g["grad3"][0]) and not face.normalGradient[3:].any()
    assert np.array_equal(ev3(g["f1"], 1), g["grad1"])
    assert ev1.launchCount() == 1 and ev3.launchCount() == 2


@pytest.mark.parametrize("n,bounds,periodic,stencil,nd", [([300, 200], [0, 1, -1, 1], (), 3, 1),
                                                          ([257, 129], [-2, 2, 0, 1], (), 5, 4),
                                                          ([128, 96], [0, 1, 0, 1], ("y",), 7, 2),
                                                          ([2048, 2048], [0, 1, 0, 1], (), 7, 4)])
def test_lattice_matches_oracle(n, bounds, periodic, stencil, nd):
    mesh = pda.create_full_mesh(n, bounds, stencil, periodic)
    rng = np.random.default_rng(20261018)
    x, y, z = mesh.viewX(), mesh.viewY(), mesh.viewZ()
    field = np.sin(np.pi * x * y)[:, None] * (1.0 + 0.5 * rng.uniform(-1, 1, (x.size, nd)))
    ev = pda.GradientEvaluator(mesh, nd)
    out = ev(field.ravel(), nd)
    o = oracle_gradient(stencil, mesh.graph(), mesh.graphRowsOfCellsNearBd(), x, y, z, mesh.dx(), mesh.dy(), field.ravel(), nd)
    assert np.array_equal(ev.cellGIDs, o["cellGid"]) and np.array_equal(ev.positions, o["position"])
    assert np.array_equal(ev.centers, o["centers"])
    assert np.array_equal(out, o["grad"])              # bit-exact
    nwalls = (0 if "x" in periodic else 2 * n[1]) + (0 if "y" in periodic else 2 * n[0])
    assert ev.numFaces() == nwalls
    # linearity (size-independent property): grad(a f + b g) = a grad f + b grad g up to rounding
    f2 = np.cos(x + 2 * y)[:, None] * np.ones((1, nd))
    lin = ev((2.0 * field + 0.5 * f2).ravel(), nd)
    assert np.allclose(lin, 2.0 * out + 0.5 * ev(f2.ravel(), nd), rtol=1e-12, atol=1e-11 * np.abs(out).max())


def test_device_pointer_entry():
    torch = pytest.importorskip("torch")
    g = load("fullmesh_s5")
    mesh = pda.mesh_from_arrays(2, int(g["stencil"]), g["d"], g["x"], g["y"], g["z"], g["graph"])
    ev = pda.GradientEvaluator(mesh, 3)
    dF = torch.from_numpy(np.ascontiguousarray(g["f3"])).cuda()
    dG = torch.empty((ev.numFaces(), 3), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ev.computeDevice(dF.data_ptr(), 3, dG.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(dG.cpu().numpy(), g["grad3"])


@pytest.mark.parametrize("scheme", ["firstorder", "weno3", "weno5"])
def test_reference_regression_with_gradients(scheme):
    """tests_cpp/eigen_2d_euler_riemann_explicit_with_gradients (main.cc:66-103, test.cmake, compare.py): Riemann
    icFlag 2 on 22x22, normal gradients of the 4-dof state at t = 0 and after 50 SSPRK3 steps of dt = 0.01, one row
    per boundary face (x, y, 4 gradients, normal direction) against grad_gold_{init,final}.txt with the reference's own
    criterion np.allclose(rtol=1e-8, atol=1e-8).  Everything (time stepping, gradients) runs on the GPU."""
    from refgold_cases import SCHEMES
    gold = np.load(os.path.join(GOLDEN, "refgold", "gradients_riemann2d.npz"))
    recon_name, stencil = SCHEMES[scheme]
    mesh = pda.create_full_mesh([22, 22], [0.0, 1.0, 0.0, 1.0], stencil)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, getattr(pda.InviscidFluxReconstruction, recon_name), 2)
    ev = pda.GradientEvaluator(mesh, 4)

    def table(state):
        g = ev(state, 4)
        t = np.column_stack([ev.centers[:, 0], ev.centers[:, 1], g, ev.normalDirections.astype(float)])
        # the reference's writer prints the first row's coordinates with the stream's default 6 significant digits
        # (std::setprecision(14) is applied after them and then sticks: main.cc:18-27)
        t[0, 0], t[0, 1] = float("%.6g" % t[0, 0]), float("%.6g" % t[0, 1])
        return t

    U = p.initialCondition()
    t0 = table(U)
    p.advance("ssprk3", U, 0.01, 50, 0.0)
    t1 = table(U)
    for got, key in ((t0, "init"), (t1, "final")):
        ref = gold["%s/%s" % (scheme, key)]
        assert got.shape == ref.shape and not np.isnan(got).any()
        assert np.allclose(ref, got, rtol=1e-8, atol=1e-8), (scheme, key, float(np.abs(ref - got).max()))
