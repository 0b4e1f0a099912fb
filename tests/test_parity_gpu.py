"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes -> libpda_b200.so), against
  * the golden fixtures produced by the unmodified reference (tests/golden/*.npz),
  * the C oracle run live on the same seeded inputs (incl. the 3D WENO5 extension, which has no reference),
  * size-independent properties at larger sizes (sample-mesh consistency, slab == full, lattice == graph kernel,
    J*a vs finite differences of the velocity, applyJacobian == J @ B).
Tolerance (BASELINE.json north_star): 1e-12 relative, 1e-10 absolute; pattern / indexing bit-exact."""
import numpy as np
import pytest

import pressiodemoapps as pda
from conftest import assert_jacobian_parity, golden_names, oracle_arrays, scaled_err
from refdrv import OracleProblem, exact_velocity_and_jacobian
from test_host_cpu import make_mesh, make_problem

pytestmark = pytest.mark.gpu
R = pda.InviscidFluxReconstruction


def mesh_arrays(mesh):
    x, y, z = mesh._coords()
    return dict(dim=mesh.dimensionality(), stencil=mesh.stencilSize(), d=mesh._deltas()[0], graph=mesh.graph(),
                x=x, y=y, z=z)


def perturbed(p, seed=20261017, amp=1e-3):
    U = p.initialCondition()
    rng = np.random.default_rng(seed)
    if not np.any(U):   # families with an identically zero initial condition
        return 0.1 * rng.uniform(-1, 1, U.size)
    return U * (1.0 + amp * rng.uniform(-1, 1, U.size))


@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_reference_golden(name, load_golden, tmp_path):
    g = load_golden(name)
    m = g.meta
    mesh, _ = make_mesh(g)
    p = make_problem(g, mesh)
    U, t = g["U"], m["t"]
    V = p.createRightHandSide()
    p.rightHandSide(U, t, V)
    assert scaled_err(V, g["V"]) <= 1.0
    J = p.createJacobian()
    assert np.array_equal(J.indptr, g["rowptr"]) and np.array_equal(J.indices, g["colidx"])
    V2 = p.createRightHandSide()
    p.rightHandSideAndJacobian(U, t, V2, J)
    assert scaled_err(V2, g["V2"]) <= 1.0

    def exact():
        mesh.write(str(tmp_path))
        prm = {k: v for k, v in (m["params"] or {}).items() if k != "testSource"}   # J does not depend on the source
        return exact_velocity_and_jacobian(str(tmp_path), m["family"], m["prob"], m["recon"], m["ic"], prm, U, t)[1]
    assert_jacobian_parity(J.data, g["Jv"], exact)
    # jacobian(U,t,J) alone (adapter_cpp.hpp:215-221) gives the same values; evaluation is repeatable
    J2 = p.createJacobian()
    p.jacobian(U, t, J2)
    assert np.array_equal(np.nan_to_num(J2.data), np.nan_to_num(J.data))
    for s in range(4):
        if "ghost%d" % s in g:
            ref = g["ghost%d" % s]
            got = p.viewGhost(s)
            w = ref != np.finfo(np.float64).tiny
            assert np.array_equal(got.reshape(ref.shape)[w], ref[w])


@pytest.mark.parametrize("n,per", [((16, 16, 16), ("x", "y", "z")), ((20, 9, 7), ("x", "y", "z")),
                                   ((7, 7, 7), ("x", "y", "z"))])
def test_3d_weno5_extension_matches_oracle(n, per, tmp_path):
    """3D WENO5 has no reference implementation: the CUDA kernels are checked against the oracle's
    restatement of the natural extension (the oracle is pinned to the reference for 3D WENO3/first order and 2D WENO5)."""
    mesh = pda.create_full_mesh(list(n), [-1, 1, -1, 1, -1, 1], 7, per)
    p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5)
    o = OracleProblem(None, "euler3d", 0, 2, arrays=mesh_arrays(mesh))
    U = perturbed(p)
    V = p.createRightHandSide()
    p.rightHandSide(U, 0.0, V)
    assert scaled_err(V, o.velocity(U, 0.0)) <= 1.0
    if np.prod(n) <= 2000:
        J = p.createJacobian()
        rp, ci = o.pattern()
        assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices)
        V2 = p.createRightHandSide()
        p.rightHandSideAndJacobian(U, 0.0, V2, J)
        Vo, Jo = o.velocityAndJacobian(U, 0.0)
        assert scaled_err(V2, Vo) <= 1.0
        assert_jacobian_parity(J.data, Jo, lambda: _exact(mesh, tmp_path, "euler3d", 0, 2, U, 0.0))


def _exact(mesh, tmp_path, fam, prob, recon, U, t):
    mesh.write(str(tmp_path))
    return exact_velocity_and_jacobian(str(tmp_path), fam, int(prob), int(recon), 1, None, U, t)[1]


@pytest.mark.parametrize("fam,prob,recon,n,bounds,per,sten", [
    ("euler3d", pda.Euler3d.PeriodicSmooth, R.Weno3, [40, 36, 32], [-1, 1, -1, 1, -1, 1], ("x", "y", "z"), 5),
    ("euler3d", pda.Euler3d.SedovSymmetry, R.Weno3, [24, 24, 24], [0, 1, 0, 1, 0, 1], (), 5),
    ("euler2d", pda.Euler2d.Riemann, R.Weno5, [200, 160], [0, 1, 0, 1], (), 7),
    ("euler2d", pda.Euler2d.DoubleMachReflection, R.Weno3, [240, 60], [0, 4, 0, 1], (), 5),
    ("swe2d", pda.Swe2d.SlipWall, R.Weno5, [150, 170], [-5, 5, -5, 5], (), 7),
    ("euler1d", pda.Euler1d.Sod, R.Weno5, [1000, 1], [-0.5, 0.5], (), 7),
    ("diffreac2d", pda.DiffusionReaction2d.GrayScott, 0, [128, 96], [-1.25, 1.25, -1.25, 1.25], ("x", "y"), 3),
    ("diffreac2d", pda.DiffusionReaction2d.ProblemA, 0, [120, 90], [0, 1, 0, 1], (), 3),
    ("diffreac1d", pda.DiffusionReaction1d.ProblemA, 0, [5000, 1], [0, 1], (), 3),
    ("advdiff2d", pda.AdvectionDiffusion2d.BurgersPeriodic, R.Weno5, [160, 130], [-1, 1, -1, 1], ("x", "y"), 7),
    ("advdiff2d", pda.AdvectionDiffusion2d.BurgersOutflow, R.Weno3, [150, 140], [-1, 1, -1, 1], (), 5),
    ("advdiffreac2d", pda.AdvectionDiffusionReaction2d.ProblemA, R.Weno5, [140, 150], [0, 1, 0, 1], (), 7),
    ("advection1d", pda.Advection1d.PeriodicLinear, R.Weno5, [4000, 1], [-1, 1], ("x",), 7),
])
def test_medium_sizes_match_oracle(fam, prob, recon, n, bounds, per, sten, tmp_path):
    """sizes between the reference's test meshes and the BASELINE configs, against the (pinned) oracle run live"""
    mesh = pda.create_full_mesh(n, bounds, sten, per)
    if fam in ("diffreac2d", "diffreac1d"):
        p = pda.create_problem(mesh, prob)
    elif fam == "advdiff2d":
        p = pda.create_problem(mesh, prob, recon, pda.ViscousFluxReconstruction.FirstOrder)
    else:
        p = pda.create_problem(mesh, prob, recon)
    o = OracleProblem(None, fam, int(prob), int(recon), arrays=mesh_arrays(mesh), omp=True)
    U = perturbed(p)
    for t in (0.0, 0.05):
        V = p.createRightHandSide()
        p.rightHandSide(U, t, V)
        assert scaled_err(V, o.velocity(U, t), field=True) <= 1.0
    if np.prod(n) * p.numDofPerCell() ** 2 * 13 < 4e7:
        J = p.createJacobian()
        V2 = p.createRightHandSide()
        p.rightHandSideAndJacobian(U, 0.0, V2, J)
        Vo, Jo = o.velocityAndJacobian(U, 0.0)
        assert scaled_err(V2, Vo, field=True) <= 1.0
        assert_jacobian_parity(J.data, Jo, lambda: _exact(mesh, tmp_path, fam, prob, recon, U, 0.0))


@pytest.mark.parametrize("prob,recon,n,bounds,sten,frac", [
    (pda.Euler2d.DoubleMachReflection, R.Weno3, [512, 128], [0, 4, 0, 1], 5, 0.05),
    (pda.Euler2d.DoubleMachReflection, R.Weno5, [512, 128], [0, 4, 0, 1], 7, 0.05),
    (pda.Euler2d.Riemann, R.Weno5, [300, 300], [0, 1, 0, 1], 7, 0.05),
    (pda.Swe2d.SlipWall, R.Weno3, [300, 280], [-5, 5, -5, 5], 5, 0.05),
])
def test_sample_mesh_consistency(prob, recon, n, bounds, sten, frac, tmp_path):
    """the rule of /root/reference/tests_cpp/sample_mesh_compare.py:36-101: V_sample == V_full[sample rows] and
    J_sample == J_full[sample rows][:, stencil cols] (here to rounding, not 1e-8) -- cfg 4 at a larger size"""
    full = pda.create_full_mesh(n, bounds, sten)
    ncell = n[0] * n[1]
    rng = np.random.default_rng(20261017)
    gids = np.sort(rng.choice(ncell, int(frac * ncell), replace=False)).astype(np.int32)
    smesh = pda.create_sample_mesh(full, gids)
    pf = pda.create_problem(full, prob, recon)
    ps = pda.create_problem(smesh, prob, recon)
    ndpc = pf.numDofPerCell()
    Uf = perturbed(pf)
    sg = smesh.stencilMeshGids()
    Us = Uf.reshape(-1, ndpc)[sg].ravel().copy()
    t = 0.03
    Vf, Vs = pf.createRightHandSide(), ps.createRightHandSide()
    pf.rightHandSide(Uf, t, Vf)
    ps.rightHandSide(Us, t, Vs)
    # a sample cell next to a cell that is absent from the stencil mesh is a near-boundary cell of the SAMPLE mesh only
    # when the absent neighbour is outside the domain; the stencil mesh contains every stencil neighbour, so rows agree
    assert scaled_err(Vs, Vf.reshape(-1, ndpc)[gids].ravel(), field=True) <= 1.0
    Jf, Js = pf.createJacobian(), ps.createJacobian()
    pf.jacobian(Uf, t, Jf)
    ps.jacobian(Us, t, Js)
    rows = (gids[:, None] * ndpc + np.arange(ndpc)[None, :]).ravel()
    cols = (sg[:, None] * ndpc + np.arange(ndpc)[None, :]).ravel()
    sub = Jf[rows][:, cols]
    diff = abs(sub - Js)
    ref = abs(sub)
    assert Js.nnz == sub.nnz or Js.nnz >= sub.nnz
    if diff.max() <= 1e-10 + 1e-12 * ref.max():
        return
    # The two rows come from two kernels (lattice: one-reciprocal arithmetic; sample mesh: graph-driven, IEEE
    # divisions) evaluating the WENO gradient, whose own rounding noise at a shock exceeds 1e-12 of the largest entry
    # (DESIGN.md 'Jacobian tolerance').  Arbiter = the exact (80-bit) Jacobian of the sample mesh: each kernel must be
    # at least as close to it as the reference's formula in double precision (the oracle) is.
    fam = "euler2d" if isinstance(prob, pda.Euler2d) else "swe2d"
    Jx = _exact(smesh, tmp_path, fam, prob, recon, Us, t)
    x, y, z = smesh._coords()
    o = OracleProblem(None, fam, int(prob), int(recon), arrays=dict(dim=2, stencil=sten, d=smesh._deltas()[0],
                                                                   graph=smesh.graph(), x=x, y=y, z=z))
    Jo = o.velocityAndJacobian(Us, t)[1]
    scale = 1e-10 + 1e-12 * np.abs(Jx).max()
    e_ref = np.abs(Jo - Jx).max() / scale
    ri = np.repeat(np.arange(Js.shape[0]), np.diff(Js.indptr))
    sub_on_pattern = np.asarray(sub.tocsr()[ri, Js.indices]).ravel()   # the full-mesh rows on the sample mesh's pattern
    assert np.abs(Js.data - Jx).max() / scale <= max(1.0, e_ref)
    assert np.abs(sub_on_pattern - Jx).max() / scale <= max(1.0, e_ref)


def test_jacobian_vs_finite_differences_3d():
    """/root/reference/tests_cpp/eigen_3d_euler_jacobian_fd_xyz_periodic/main.cc:54-126: J*a vs FD of the velocity
    (eps 1e-8, tol 1e-4 there); evaluated twice to check re-entrancy.  Also for the WENO5 extension."""
    for recon, sten in ((R.Weno3, 5), (R.Weno5, 7)):
        mesh = pda.create_full_mesh([10, 9, 8], [-1, 1, -1, 1, -1, 1], sten, ("x", "y", "z"))
        p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, recon)
        U = perturbed(p, amp=1e-2)
        rng = np.random.default_rng(7)
        a = rng.uniform(-1, 1, U.size)
        for _ in range(2):
            J = p.createJacobian()
            V = p.createRightHandSide()
            p.rightHandSideAndJacobian(U, 0.0, V, J)
            Ja = J @ a
            eps = 1e-6
            Vp, Vm = p.createRightHandSide(), p.createRightHandSide()
            p.rightHandSide(U + eps * a, 0.0, Vp)
            p.rightHandSide(U - eps * a, 0.0, Vm)
            fd = (Vp - Vm) / (2 * eps)
            assert np.max(np.abs(fd - Ja)) < 1e-5 * max(1.0, np.abs(Ja).max())


def test_apply_jacobian_layouts():
    """adapter_cpp.hpp:231-259: R = J(U,t) B for a vector, a col-major and a row-major operand"""
    mesh = pda.create_full_mesh([40, 30], [0, 1, 0, 1], 5)
    p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno3)
    U = perturbed(p)
    J = p.createJacobian()
    p.jacobian(U, 0.0, J)
    rng = np.random.default_rng(3)
    b = rng.uniform(-1, 1, U.size)
    r = p.createApplyJacobianResult(b)
    p.applyJacobian(U, b, 0.0, r)
    assert scaled_err(r, J @ b, 1e-11, 1e-9) <= 1.0
    for order in ("C", "F"):
        B = np.asarray(rng.uniform(-1, 1, (U.size, 5)), order=order)
        Rm = p.createApplyJacobianResult(B)
        assert Rm.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
        p.applyJacobian(U, B, 0.0, Rm)
        assert scaled_err(Rm, J @ B, 1e-11, 1e-9) <= 1.0


def test_structured_and_graph_kernels_agree_large():
    """full-size property: the structured lattice kernel (no graph in HBM) and the graph-driven kernel (same mesh
    handed over as arrays => not recognised as a lattice) give the same velocity on a 96^3 WENO5 mesh"""
    n = [96, 96, 96]
    lat = pda.create_full_mesh(n, [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    p1 = pda.create_problem(lat, pda.Euler3d.PeriodicSmooth, R.Weno5)
    U = perturbed(p1)
    V1 = p1.createRightHandSide()
    p1.rightHandSide(U, 0.0, V1)
    ma = mesh_arrays(lat)
    generic = pda.mesh_from_arrays(3, 7, ma["d"], ma["x"], ma["y"], ma["z"], ma["graph"], detect_lattice=False)
    p2 = pda.create_problem(generic, pda.Euler3d.PeriodicSmooth, R.Weno5)
    V2 = p2.createRightHandSide()
    p2.rightHandSide(U, 0.0, V2)
    assert scaled_err(V1, V2) <= 1.0
    # translation invariance on the periodic lattice: shifting the state by (3,5,7) cells shifts the velocity
    Ug = U.reshape(n[2], n[1], n[0], 5)
    Us = np.roll(Ug, (7, 5, 3), axis=(0, 1, 2)).ravel().copy()
    V3 = p1.createRightHandSide()
    p1.rightHandSide(Us, 0.0, V3)
    assert np.array_equal(np.roll(V1.reshape(n[2], n[1], n[0], 5), (7, 5, 3), axis=(0, 1, 2)).ravel(), V3)


def test_custom_bcs_swe():
    """Swe2d::CustomBCs with device-expressible rules: Reflective on all sides == SlipWall;
    Dirichlet / HomogNeumann as in /root/reference/tests_cpp/eigen_2d_swe_custom_bcs/main.cc:6-58 vs a numpy restatement
    of those functors (ghost = fixed state / ghost = own cell) through the oracle's first-order velocity."""
    mesh = pda.create_full_mesh([30, 26], [-5, 5, -5, 5], 5)
    slip = pda.create_problem(mesh, pda.Swe2d.SlipWall, R.Weno3)
    cust = pda.create_problem(mesh, pda.Swe2d.CustomBCs, R.Weno3)
    for s in range(4):
        cust.setBC(s, pda.BC.Reflective)
    U = perturbed(slip)
    U[1::3] = 0.1 * np.sin(np.arange(U.size // 3))
    U[2::3] = 0.1 * np.cos(np.arange(U.size // 3))
    Va, Vb = slip.createRightHandSide(), cust.createRightHandSide()
    slip.rightHandSide(U, 0.0, Va)
    cust.rightHandSide(U, 0.0, Vb)
    assert np.array_equal(Va, Vb)
    Ja, Jb = slip.createJacobian(), cust.createJacobian()
    slip.jacobian(U, 0.0, Ja)
    cust.jacobian(U, 0.0, Jb)
    assert np.array_equal(Ja.data, Jb.data)
    # Dirichlet left / HomogNeumann elsewhere: ghost rows hold exactly the prescribed values
    d = pda.create_problem(mesh, pda.Swe2d.CustomBCs, R.Weno3)
    d.setBC(0, pda.BC.Dirichlet, [1.5, 0.2, -0.1])
    for s in (1, 2, 3):
        d.setBC(s, pda.BC.HomogNeumann)
    V = d.createRightHandSide()
    d.rightHandSide(U, 0.0, V)
    gl = d.viewGhost(0)
    g = mesh.graph()
    nb = mesh.graphRowsOfCellsNearBd()
    for r, row in enumerate(nb):
        if g[row, 1] == -1:
            assert np.array_equal(gl[r, :3], [1.5, 0.2, -0.1])
        if g[row, 3] == -1:
            assert np.array_equal(d.viewGhost(2)[r, :3], U[3 * g[row, 0]: 3 * g[row, 0] + 3])
    assert np.isfinite(V).all()


@pytest.mark.parametrize("nranks,n", [(2, (40, 24, 16)), (4, (33, 20, 24))])
def test_slab_decomposition_equals_full(nranks, n):
    """multi-GPU layout on ONE device: every rank's slab (owned planes + 3 halo planes per side, filled the way the
    halo exchange fills them) evaluated with the interior + boundary launches reproduces the full-mesh velocity
    bit for bit (same kernel, same arithmetic, different storage)."""
    import torch
    mesh = pda.create_full_mesh(list(n), [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    full = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5)
    U = perturbed(full)
    Vfull = full.createRightHandSide()
    full.rightHandSide(U, 0.0, Vfull)
    nz = n[2]
    pd = n[0] * n[1] * 5
    Ug = torch.from_numpy(U).cuda().reshape(nz, pd)
    st = torch.cuda.current_stream().cuda_stream
    for r in range(nranks):
        p = pda.create_problem_slab(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5, r, nranks)
        k0, k1, h, pdofs = p.slabExtent()
        assert pdofs == pd and h == 3
        planes = [(k % nz) for k in range(k0 - h, k1 + h)]
        Ul = Ug[planes].contiguous().reshape(-1)
        Vl = torch.zeros((k1 - k0) * pd, dtype=torch.float64, device="cuda")
        p.slabVelocityInteriorDevice(Ul.data_ptr(), 0.0, Vl.data_ptr(), st)
        p.slabVelocityBoundaryDevice(Ul.data_ptr(), 0.0, Vl.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(Vl.cpu().numpy(), Vfull[k0 * pd:k1 * pd])


@pytest.mark.parametrize("nranks,n,recon", [(1, (20, 18, 16), "Weno5"), (2, (20, 18, 16), "Weno5"),
                                            (4, (36, 20, 24), "Weno5"), (4, (36, 20, 48), "Weno3"),
                                            (2, (64, 48, 160), "Weno5")])
def test_slab_peer_mode_equals_full(nranks, n, recon):
    """peer mode (halo exchange fused with the evaluation, include/pda_b200.h) on ONE device: the ranks are slab
    problems of one process, wired to each other with pda_slab_peer_connect_local (distinct streams, as the header
    demands).  Each rank's owned planes only are passed in; the neighbours' copy-engine pushes + flags deliver the
    halos.  Several evaluations with changing states exercise the epoch/parity protocol; every result must equal the
    full-mesh velocity bit for bit."""
    import torch
    scheme = getattr(R, recon)
    mesh = pda.create_full_mesh(list(n), [-1, 1, -1, 1, -1, 1], 7 if recon == "Weno5" else 5, ("x", "y", "z"))
    full = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, scheme)
    nz, pd = n[2], n[0] * n[1] * 5
    probs = [pda.create_problem_slab(mesh, pda.Euler3d.PeriodicSmooth, scheme, r, nranks) for r in range(nranks)]
    if nranks == 1:
        probs[0].peerConnect([probs[0].peerHandle()])
    else:
        for r, p in enumerate(probs):
            p.peerConnectLocal(probs[(r - 1) % nranks], probs[(r + 1) % nranks])
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    ext = [p.slabExtent() for p in probs]
    for it in range(5):
        U = perturbed(full, seed=100 + it)
        Vfull = full.createRightHandSide()
        full.rightHandSide(U, 0.0, Vfull)
        Ug = torch.from_numpy(U).cuda()
        Us = [Ug[k0 * pd:k1 * pd].clone() for (k0, k1, _, _) in ext]
        Vs = [torch.zeros((k1 - k0) * pd, dtype=torch.float64, device="cuda") for (k0, k1, _, _) in ext]
        torch.cuda.synchronize()
        for r, p in enumerate(probs):
            p.slabVelocityPeerDevice(Us[r].data_ptr(), 0.0, Vs[r].data_ptr(), streams[r].cuda_stream)
        torch.cuda.synchronize()
        for r, (k0, k1, _, _) in enumerate(ext):
            assert np.array_equal(Vs[r].cpu().numpy(), Vfull[k0 * pd:k1 * pd]), (it, r)


@pytest.mark.parametrize("case", ["euler_weno5", "euler_weno3_per", "swe_weno5", "euler_fo", "burgers_weno5_per",
                                  "swe_fo", "swe_weno3", "burgers_fo_per", "burgers_weno3_out", "swe_weno3_on_s7_mesh"])
def test_lattice_and_graph_jacobian_kernels_agree(case):
    """medium-size property: the face-sharing lattice Jacobian kernel (kernels_jaclattice.cuh: full 2D lattices) and
    the graph-driven staged kernel (same mesh handed over as arrays => not recognised as a lattice) give the same
    velocity and Jacobian; ragged 31x31 tiles, periodic wrap and scheme stencil < mesh stencil included."""
    if case == "euler_weno5":
        n, bounds, st, per = [100, 71], [0, 1, 0, 1], 7, ()
        mk = lambda m: pda.create_problem(m, pda.Euler2d.Riemann, R.Weno5)
    elif case == "euler_weno3_per":
        n, bounds, st, per = [64, 45], [-1, 1, -1, 1], 7, ("x", "y")
        mk = lambda m: pda.create_problem(m, pda.Euler2d.PeriodicSmooth, R.Weno3)
    elif case == "swe_weno5":
        n, bounds, st, per = [70, 66], [-5, 5, -5, 5], 7, ()
        mk = lambda m: pda.create_problem(m, pda.Swe2d.SlipWall, R.Weno5)
    elif case == "euler_fo":
        n, bounds, st, per = [40, 37], [0, 1, 0, 1], 3, ()
        mk = lambda m: pda.create_problem(m, pda.Euler2d.Riemann, R.FirstOrder)
    # the y-marching Jacobian kernels (k_jacobian_march2d_fo / _weno): several strips (30 / 28 / 26 output columns per
    # warp), several y chunks, ragged last strip, periodic wrap, a scheme stencil narrower than the mesh stencil
    elif case == "swe_fo":
        n, bounds, st, per = [95, 41], [-5, 5, -5, 5], 3, ()
        mk = lambda m: pda.create_problem(m, pda.Swe2d.SlipWall, R.FirstOrder)
    elif case == "swe_weno3":
        n, bounds, st, per = [90, 44], [-5, 5, -5, 5], 5, ()
        mk = lambda m: pda.create_problem(m, pda.Swe2d.SlipWall, R.Weno3)
    elif case == "swe_weno3_on_s7_mesh":
        n, bounds, st, per = [66, 37], [-5, 5, -5, 5], 7, ()
        mk = lambda m: pda.create_problem(m, pda.Swe2d.SlipWall, R.Weno3)
    elif case == "burgers_fo_per":
        n, bounds, st, per = [64, 33], [-1, 1, -1, 1], 3, ("x", "y")
        mk = lambda m: pda.create_problem(m, pda.AdvectionDiffusion2d.BurgersPeriodic, R.FirstOrder,
                                          pda.ViscousFluxReconstruction.FirstOrder)
    elif case == "burgers_weno3_out":
        n, bounds, st, per = [61, 35], [-1, 1, -1, 1], 5, ()
        mk = lambda m: pda.create_problem(m, pda.AdvectionDiffusion2d.BurgersOutflow, R.Weno3,
                                          pda.ViscousFluxReconstruction.FirstOrder)
    else:
        n, bounds, st, per = [48, 40], [-1, 1, -1, 1], 7, ("x", "y")
        mk = lambda m: pda.create_problem(m, pda.AdvectionDiffusion2d.BurgersPeriodic, R.Weno5,
                                          pda.ViscousFluxReconstruction.FirstOrder)
    lat = pda.create_full_mesh(n, bounds, st, per)
    p1 = mk(lat)
    ma = mesh_arrays(lat)
    generic = pda.mesh_from_arrays(2, st, ma["d"], ma["x"], ma["y"], ma["z"], ma["graph"], detect_lattice=False)
    p2 = mk(generic)
    U = perturbed(p1)
    V1, V2 = p1.createRightHandSide(), p2.createRightHandSide()
    J1, J2 = p1.createJacobian(), p2.createJacobian()
    p1.rightHandSideAndJacobian(U, 0.0, V1, J1)
    p2.rightHandSideAndJacobian(U, 0.0, V2, J2)
    assert np.array_equal(J1.indptr, J2.indptr) and np.array_equal(J1.indices, J2.indices)
    assert scaled_err(V1, V2, field=True) <= 1.0
    # same formulas, different association of the two faces' products: rounding-level differences only
    assert scaled_err(J1.data, J2.data, 1e-11, 1e-9) <= 1.0


def test_slab_peer_mode_host_pipeline_single_rank():
    """pda_slab_velocity_peer_host with one rank (its own ring neighbour): chunked H2D -> kernel -> D2H pipeline, pushes
    issued after the two boundary chunks; equals the full-mesh velocity bit for bit over several calls.  (More ranks
    need one thread or process each -- the call is synchronous -- see the multi-process test below.)"""
    n = (24, 20, 96)
    mesh = pda.create_full_mesh(list(n), [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    full = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5)
    p = pda.create_problem_slab(mesh, pda.Euler3d.PeriodicSmooth, R.Weno5, 0, 1)
    p.peerConnect([p.peerHandle()])
    for it in range(3):
        U = perturbed(full, seed=300 + it)
        Vfull = full.createRightHandSide()
        full.rightHandSide(U, 0.0, Vfull)
        V = np.zeros_like(U)
        p.slabVelocityPeer(U, 0.0, V)
        assert np.array_equal(V, Vfull), it


def test_slab_peer_mode_multi_process():
    """peer mode across PROCESSES (one rank per GPU, IPC-mapped halo buffers, cross-GPU flags): tools/check_peer_multi.py
    under torchrun on two GPUs; every rank's slab must equal the full-mesh velocity bit for bit.  Skipped on a
    single-GPU box (the single-process protocol test above still runs there)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(root, "tools", "check_peer_multi.py"), "--cells", "64"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatching evaluations: 0" in r.stdout


@pytest.mark.parametrize("case", ["euler2d_lattice", "swe_sample", "euler3d", "burgers"])
@pytest.mark.parametrize("ncols", [8, 25, 40])
def test_apply_jacobian_many_columns(case, ncols):
    """operands with >= 8 columns take the per-cell J*B kernel (lanes across operand columns; column-major operands
    transposed around it; lattices visited in tile-major order): every layout must equal J @ B of the assembled
    Jacobian (tests_perf/main.py:37-48 uses 25 columns)"""
    if case == "euler2d_lattice":
        mesh = pda.create_full_mesh([37, 29], [0, 1, 0, 1], 7)
        p = pda.create_problem(mesh, pda.Euler2d.Riemann, R.Weno5)
    elif case == "swe_sample":
        full = pda.create_full_mesh([30, 30], [-5, 5, -5, 5], 5)
        gids = np.sort(np.random.default_rng(5).choice(900, 200, replace=False)).astype(np.int32)
        p = pda.create_problem(pda.create_sample_mesh(full, gids), pda.Swe2d.SlipWall, R.Weno3)
    elif case == "euler3d":
        mesh = pda.create_full_mesh([10, 9, 8], [-1, 1, -1, 1, -1, 1], 5, ("x", "y", "z"))
        p = pda.create_problem(mesh, pda.Euler3d.PeriodicSmooth, R.Weno3)
    else:
        mesh = pda.create_full_mesh([20, 18], [-1, 1, -1, 1], 5, ("x", "y"))
        p = pda.create_problem(mesh, pda.AdvectionDiffusion2d.BurgersPeriodic, R.Weno3,
                               pda.ViscousFluxReconstruction.FirstOrder)
    U = perturbed(p)
    J = p.createJacobian()
    p.jacobian(U, 0.0, J)
    rng = np.random.default_rng(11)
    for order in ("C", "F"):
        B = np.asarray(rng.uniform(-1, 1, (U.size, ncols)), order=order)
        Rm = p.createApplyJacobianResult(B)
        p.applyJacobian(U, B, 0.0, Rm)
        assert scaled_err(Rm, J @ B, 1e-11, 1e-9) <= 1.0


def test_c_abi_demo_on_gpu(tmp_path):
    """the plain-C client of include/pda_b200.h (examples/c_abi_demo.c): velocity + Jacobian + 10 SSPRK3 steps"""
    import subprocess
    from test_host_cpu import _build_c_demo
    exe = _build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max|V|" in r.stdout and "c_abi_demo ok" in r.stdout


def test_custom_bcs_host_functors_equal_device_rules():
    """PDA_BC_HOST_CALLBACK: the reference's functor contract (tests_cpp/eigen_2d_swe_custom_bcs/main.cc:6-58: Dirichlet
    on the left, homogeneous Neumann elsewhere, with their Jacobian-factor overloads) through host callbacks gives the
    same ghosts, velocity and Jacobian as the device-expressible rules -- bit for bit."""
    mesh = pda.create_full_mesh([30, 26], [-5, 5, -5, 5], 3)
    dirich = np.array([0.00001, 0.004, 0.001])

    def make(kind):
        p = pda.create_problem(mesh, pda.Swe2d.CustomBCs, R.FirstOrder)
        if kind == "device":
            p.setBC(0, pda.BC.Dirichlet, dirich)
            for s in (1, 2, 3):
                p.setBC(s, pda.BC.HomogNeumann)
            return p

        def g_dirichlet(row_id, grow, x, y, U, ndpc, width, out):
            out[:3] = dirich

        def f_dirichlet(grow, x, y, ndpc, fac):
            fac[:] = 0.0

        def g_neumann(row_id, grow, x, y, U, ndpc, width, out):
            c = grow[0] * ndpc
            out[:3] = U[c:c + 3]

        def f_neumann(grow, x, y, ndpc, fac):
            fac[:] = 1.0
        p.setBCFunctor(0, g_dirichlet, f_dirichlet)
        for s in (1, 2, 3):
            p.setBCFunctor(s, g_neumann, f_neumann)
        return p
    pd_, ph = make("device"), make("host")
    U = perturbed(pd_)
    Vd, Vh = pd_.createRightHandSide(), ph.createRightHandSide()
    Jd, Jh = pd_.createJacobian(), ph.createJacobian()
    pd_.rightHandSideAndJacobian(U, 0.0, Vd, Jd)
    ph.rightHandSideAndJacobian(U, 0.0, Vh, Jh)
    for s in range(4):
        assert np.array_equal(pd_.viewGhost(s), ph.viewGhost(s)), s
    assert np.array_equal(Vd, Vh)
    assert np.array_equal(Jd.data, Jh.data)
    V2, V3 = ph.createRightHandSide(), pd_.createRightHandSide()
    ph.rightHandSide(U, 0.0, V2)      # velocity-only path (other inner-cell kernel than the Jacobian path)
    pd_.rightHandSide(U, 0.0, V3)
    assert np.array_equal(V2, V3)


@pytest.mark.parametrize("nranks,n,recon,sten", [(2, (70, 48), "Weno5", 7), (3, (33, 60), "Weno3", 5), (4, (40, 32), "FirstOrder", 3)])
def test_slab_decomposition_2d_equals_full(nranks, n, recon, sten):
    """2D y-slabs (send/recv layout: [halo rows | owned rows | halo rows]) through the y-marching kernel reproduce the
    full-mesh velocity bit for bit"""
    import torch
    scheme = getattr(R, recon)
    mesh = pda.create_full_mesh(list(n), [-1, 1, -1, 1], sten, ("x", "y"))
    full = pda.create_problem(mesh, pda.Euler2d.PeriodicSmooth, scheme)
    U = perturbed(full)
    Vfull = full.createRightHandSide()
    full.rightHandSide(U, 0.0, Vfull)
    ny, pd = n[1], n[0] * 4
    Ug = torch.from_numpy(U).cuda().reshape(ny, pd)
    st = torch.cuda.current_stream().cuda_stream
    for r in range(nranks):
        p = pda.create_problem_slab(mesh, pda.Euler2d.PeriodicSmooth, scheme, r, nranks)
        k0, k1, h, pdofs = p.slabExtent()
        assert pdofs == pd and h == (sten - 1) // 2
        rows = [(k % ny) for k in range(k0 - h, k1 + h)]
        Ul = Ug[rows].contiguous().reshape(-1)
        Vl = torch.zeros((k1 - k0) * pd, dtype=torch.float64, device="cuda")
        p.slabVelocityInteriorDevice(Ul.data_ptr(), 0.0, Vl.data_ptr(), st)
        p.slabVelocityBoundaryDevice(Ul.data_ptr(), 0.0, Vl.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(Vl.cpu().numpy(), Vfull[k0 * pd:k1 * pd]), r


@pytest.mark.parametrize("case", ["euler_riemann_weno5", "euler_per_weno3", "swe_weno5", "burgers_per_weno5", "adr_weno3",
                                  "euler_tiny_per", "swe_fo", "burgers_out_weno3", "euler_fo_wide"])
def test_matrix_free_apply_jacobian_equals_assembled(case):
    """operands with <= 12 columns on 2D lattices take the matrix-free inner-row kernels -- vectors the y-marching
    (value, tangent) kernel (kernels_applymarch2d.cuh), several columns the tile kernel (kernels_applylattice.cuh): the
    directional derivative of every face flux, no CSR values stored; result must equal J @ B of the assembled Jacobian
    for vectors and both matrix layouts, incl. point / diffusion terms, periodic wrap, several strips and tiny meshes"""
    V = pda.ViscousFluxReconstruction.FirstOrder
    if case == "swe_fo":
        p = pda.create_problem(pda.create_full_mesh([70, 21], [-5, 5, -5, 5], 3), pda.Swe2d.SlipWall, R.FirstOrder)
    elif case == "burgers_out_weno3":
        p = pda.create_problem(pda.create_full_mesh([63, 19], [-1, 1, -1, 1], 5), pda.AdvectionDiffusion2d.BurgersOutflow, R.Weno3, V)
    elif case == "euler_fo_wide":
        p = pda.create_problem(pda.create_full_mesh([100, 17], [0, 1, 0, 1], 3), pda.Euler2d.Riemann, R.FirstOrder)
    elif case == "euler_riemann_weno5":
        p = pda.create_problem(pda.create_full_mesh([47, 33], [0, 1, 0, 1], 7), pda.Euler2d.Riemann, R.Weno5)
    elif case == "euler_per_weno3":
        p = pda.create_problem(pda.create_full_mesh([40, 31], [-1, 1, -1, 1], 7, ("x", "y")), pda.Euler2d.PeriodicSmooth, R.Weno3)
    elif case == "swe_weno5":
        p = pda.create_problem(pda.create_full_mesh([36, 40], [-5, 5, -5, 5], 7), pda.Swe2d.SlipWall, R.Weno5)
    elif case == "burgers_per_weno5":
        p = pda.create_problem(pda.create_full_mesh([30, 32], [-1, 1, -1, 1], 7, ("x", "y")),
                               pda.AdvectionDiffusion2d.BurgersPeriodic, R.Weno5, V)
    elif case == "adr_weno3":
        p = pda.create_problem(pda.create_full_mesh([26, 22], [0, 1, 0, 1], 5), pda.AdvectionDiffusionReaction2d.ProblemA, R.Weno3)
    else:
        p = pda.create_problem(pda.create_full_mesh([6, 5], [-1, 1, -1, 1], 7, ("x", "y")), pda.Euler2d.PeriodicSmooth, R.Weno5)
    U = perturbed(p)
    J = p.createJacobian()
    p.jacobian(U, 0.0, J)
    rng = np.random.default_rng(21)
    b = rng.uniform(-1, 1, U.size)
    r = p.createApplyJacobianResult(b)
    p.applyJacobian(U, b, 0.0, r)
    assert scaled_err(r, J @ b, 1e-11, 1e-9) <= 1.0
    for ncols in (3, 8, 12):
        for order in ("C", "F"):
            B = np.asarray(rng.uniform(-1, 1, (U.size, ncols)), order=order)
            Rm = p.createApplyJacobianResult(B)
            p.applyJacobian(U, B, 0.0, Rm)
            assert scaled_err(Rm, J @ B, 1e-11, 1e-9) <= 1.0


@pytest.mark.parametrize("case", ["per_weno5", "per_weno3", "sedov_weno3", "per_fo_ragged", "per_weno5_tma", "sedov_weno3_tma",
                                  "sedov_weno5_tma"])
def test_matrix_free_apply_jacobian_3d(case):
    """3D lattices: applyJacobian never assembles the inner rows: vectors and column-major operands go column by column
    through the tiled (value, tangent) kernel k_applyjac_tiled3d (TMA-staged tiles once nx >= 40 and even: the *_tma
    cases; 8-byte cp.async copies otherwise), row-major multi-column operands through the line kernel
    k_applyjac_lattice3d (7^3 tiles, x/y/z phases); both equal J @ B of the assembled Jacobian, incl. near-boundary rows
    (Sedov symmetry walls) and ragged tiles"""
    if case == "per_weno5_tma":
        p = pda.create_problem(pda.create_full_mesh([48, 9, 8], [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z")), pda.Euler3d.PeriodicSmooth, R.Weno5)
    elif case == "sedov_weno3_tma":
        p = pda.create_problem(pda.create_full_mesh([44, 16, 7], [0, 1, 0, 1, 0, 1], 5), pda.Euler3d.SedovSymmetry, R.Weno3)
    elif case == "sedov_weno5_tma":
        p = pda.create_problem(pda.create_full_mesh([70, 8, 9], [0, 1, 0, 1, 0, 1], 7), pda.Euler3d.SedovSymmetry, R.Weno5)
    elif case == "per_weno5":
        p = pda.create_problem(pda.create_full_mesh([10, 9, 8], [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z")), pda.Euler3d.PeriodicSmooth, R.Weno5)
    elif case == "per_weno3":
        p = pda.create_problem(pda.create_full_mesh([16, 15, 9], [-1, 1, -1, 1, -1, 1], 5, ("x", "y", "z")), pda.Euler3d.PeriodicSmooth, R.Weno3)
    elif case == "sedov_weno3":
        p = pda.create_problem(pda.create_full_mesh([12, 11, 10], [0, 1, 0, 1, 0, 1], 5), pda.Euler3d.SedovSymmetry, R.Weno3)
    else:
        p = pda.create_problem(pda.create_full_mesh([8, 15, 6], [-1, 1, -1, 1, -1, 1], 3, ("x", "y", "z")), pda.Euler3d.PeriodicSmooth, R.FirstOrder)
    U = perturbed(p, amp=1e-2)
    J = p.createJacobian()
    p.jacobian(U, 0.0, J)
    rng = np.random.default_rng(31)
    b = rng.uniform(-1, 1, U.size)
    r = p.createApplyJacobianResult(b)
    p.applyJacobian(U, b, 0.0, r)
    assert scaled_err(r, J @ b, 1e-11, 1e-9) <= 1.0
    for order in ("C", "F"):
        B = np.asarray(rng.uniform(-1, 1, (U.size, 3)), order=order)
        Rm = p.createApplyJacobianResult(B)
        p.applyJacobian(U, B, 0.0, Rm)
        assert scaled_err(Rm, J @ B, 1e-11, 1e-9) <= 1.0


def test_set_bc_pointer_repoints_host_functor_state():
    """setBCPointer(loc, ptr) (euler_2d_prob_class.hpp:213-216 -> custom_bc_holder.hpp:89-103): the state handed to a
    side's host functor can be swapped between evaluations (pda_problem_set_bc_pointer) -- a Dirichlet functor reading
    its ghost state through that pointer equals the device Dirichlet rule with the respective values, bit for bit."""
    import ctypes as C
    mesh = pda.create_full_mesh([20, 18], [-5, 5, -5, 5], 3)
    valsA = np.array([0.00001, 0.004, 0.001])
    valsB = np.array([0.5, -0.25, 0.125])
    p = pda.create_problem(mesh, pda.Swe2d.CustomBCs, R.FirstOrder)

    def ghost(user, row_id, grow, x, y, U, nd, width, out):
        src = C.cast(user, C.POINTER(C.c_double))
        for d in range(nd):
            out[d] = src[d]

    def neumann(user, row_id, grow, x, y, U, nd, width, out):
        for d in range(nd):
            out[d] = U[grow[0] * nd + d]
    cg, cn, cf = pda._GHOST_FN(ghost), pda._GHOST_FN(neumann), pda._FACTOR_FN()
    pda._check(pda._lib.pda_problem_set_bc_callback(p._h, 0, cg, cf, valsA.ctypes.data))
    for s in (1, 2, 3):
        pda._check(pda._lib.pda_problem_set_bc_callback(p._h, s, cn, cf, None))
    U = perturbed(p)
    for vals in (valsA, valsB, valsA):
        p.setBCPointer(0, vals.ctypes.data)
        V = p.createRightHandSide()
        p.rightHandSide(U, 0.0, V)
        q = pda.create_problem(mesh, pda.Swe2d.CustomBCs, R.FirstOrder)
        q.setBC(0, pda.BC.Dirichlet, vals)
        for s in (1, 2, 3):
            q.setBC(s, pda.BC.HomogNeumann)
        Vq = q.createRightHandSide()
        q.rightHandSide(U, 0.0, Vq)
        assert np.array_equal(V, Vq)
        assert np.array_equal(p.viewGhost(0), q.viewGhost(0))


@pytest.mark.parametrize("n,stencil,recon", [([176, 160, 168], 7, "Weno5"), ([200, 168, 150], 5, "Weno3"),
                                             ([2100, 2050], 7, "Weno5")])
def test_host_pipeline_equals_device_path(n, stencil, recon):
    """pda_problem_velocity_host on large periodic lattices (>= 4.2 M cells) runs the chunked H2D -> kernel -> D2H
    pipeline (chunks of 5-6 planes, evaluatePlanes per chunk): same bits as ONE evaluation of the whole mesh through
    the device-pointer entry, also on a second call (buffers reused) and from pinned memory"""
    torch = pytest.importorskip("torch")
    dim = len(n)
    bounds = [-1, 1] * dim
    mesh = pda.create_full_mesh(n, bounds, stencil, ("x", "y", "z")[:dim])
    fam = pda.Euler3d if dim == 3 else pda.Euler2d
    p = pda.create_problem(mesh, fam.PeriodicSmooth, getattr(R, recon))
    U = perturbed(p)
    dU = torch.from_numpy(U).cuda()
    dV = torch.empty(p.totalDofSampleMesh(), dtype=torch.float64, device="cuda")
    p.rightHandSideDevice(dU.data_ptr(), 0.0, dV.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Vdev = dV.cpu().numpy()
    V = p.createRightHandSide()
    p.rightHandSide(U, 0.0, V)
    assert np.array_equal(V, Vdev)
    Up = torch.from_numpy(U * 1.25).pin_memory()
    Vp = torch.empty_like(Up).pin_memory()
    p.rightHandSide(Up.numpy(), 0.0, Vp.numpy())
    dU.mul_(1.25)
    p.rightHandSideDevice(dU.data_ptr(), 0.0, dV.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(Vp.numpy(), dV.cpu().numpy())


def test_cpp_eigen_shim_on_device(tmp_path):
    """include/pda_b200_eigen.hpp on a GPU: examples/cpp_shim_demo (prebuilt by __graft_entry__.build() where Eigen is
    available; travels with the snapshot) creates problems like the reference's tests_cpp mains and checks, on the
    device, applyJacobian vs J*B, the boundary-face gradients, and custom-BC host FUNCTORS with the reference's two call
    operators (tests_cpp/eigen_2d_swe_custom_bcs/main.cc) against the device BC tables (bit-equal V and J, FD check)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "_bin", "cpp_shim_demo")
    if not os.path.exists(exe):
        pytest.skip("examples/_bin/cpp_shim_demo not built (needs Eigen at build time)")
    mdir = os.path.join(str(tmp_path), "mesh")
    pda.create_full_mesh([20, 20], [0, 1, 0, 1], 7).write(mdir)
    r = subprocess.run([exe, mdir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp_shim_demo ok" in r.stdout and "custom BC functors vs device rules: |dV|max 0.0e+00 |dJ|max 0.0e+00" in r.stdout
    assert "left-wall faces 20" in r.stdout


def test_two_devices_in_one_process():
    """kernel attributes (dynamic shared memory > 48 KB) are per DEVICE: a second problem on another GPU of the same
    process must get its own set-up (func_attrs.hpp).  Needs two GPUs; identical results on both."""
    if pda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mesh3 = pda.create_full_mesh([40, 21, 16], [-1, 1, -1, 1, -1, 1], 7, ("x", "y", "z"))
    mesh2 = pda.create_full_mesh([64, 48], [0, 1, 0, 1], 7)
    out = []
    for dev in (0, 1):
        p3 = pda.create_problem(mesh3, pda.Euler3d.PeriodicSmooth, R.Weno5, device=dev)
        U3 = perturbed(p3)
        V3 = p3.createRightHandSide()
        p3.rightHandSide(U3, 0.0, V3)                      # tiled 3D kernel: ~105 KB dynamic shared memory
        p2 = pda.create_problem(mesh2, pda.Euler2d.Riemann, R.Weno5, device=dev)
        U2 = perturbed(p2)
        J = p2.createJacobian()
        V2 = p2.createRightHandSide()
        p2.rightHandSideAndJacobian(U2, 0.0, V2, J)        # lattice Jacobian kernel: > 48 KB too
        B = np.random.default_rng(2).uniform(-1, 1, (U2.size, 3))
        Rm = p2.createApplyJacobianResult(B)
        p2.applyJacobian(U2, B, 0.0, Rm)
        out.append((V3, V2, J.data.copy(), Rm))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
