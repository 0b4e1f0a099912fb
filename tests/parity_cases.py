"""Shared bodies of the parity checks that run at BASELINE.json's real sizes and of the reference-order mode checks:
used by tests/test_fullsize_gpu.py, tests/test_reforder_gpu.py (which assert) and tools/parity_report.py (which prints
the table committed as profiles/parity_report_r02.txt).  Every function returns a dict of numbers; 'strict' is
max |a-b| / (1e-10 + 1e-12 |b|) per entry (<= 1 <=> the north-star tolerance), 'field' measures the relative part
against max |b| (reported, never silently substituted), 'bits' is the number of entries that differ in any bit.
Oracle = oracle/pda_oracle.c run live (OpenMP build) -- TEST INFRASTRUCTURE."""
import time

import numpy as np

import pressiodemoapps as pda
from refdrv import OracleProblem, lattice_spec

R = pda.InviscidFluxReconstruction
RTOL, ATOL = 1e-12, 1e-10
SEED = 20261017


def err_stats(a, b, chunk=1 << 24):
    """strict / field-scaled error and bitwise differences of two (possibly huge) arrays; NaNs must coincide"""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    assert a.shape == b.shape
    strict = 0.0
    maxabs = 0.0
    bmax = 0.0
    bits = 0
    nan_mismatch = 0
    nans = 0
    for s in range(0, a.size, chunk):
        x, y = a[s:s + chunk], b[s:s + chunk]
        nx, ny = np.isnan(x), np.isnan(y)
        nan_mismatch += int(np.count_nonzero(nx != ny))
        nans += int(np.count_nonzero(ny))
        ok = ~(nx | ny)
        if not ok.any():
            continue
        d = np.abs(x[ok] - y[ok])
        ay = np.abs(y[ok])
        strict = max(strict, float(np.max(d / (ATOL + RTOL * ay))))
        maxabs = max(maxabs, float(d.max()))
        bmax = max(bmax, float(ay.max()))
        bits += int(np.count_nonzero(x[ok] != y[ok]))
    return dict(strict=strict, field=maxabs / (ATOL + RTOL * bmax), maxabs=maxabs, maxref=bmax, bits=bits, n=int(a.size),
                nans=nans, nan_mismatch=nan_mismatch)


def perturb_inplace(U, seed=SEED, amp=1e-3, chunk=1 << 24):
    """U *= 1 + amp*xi, xi ~ U(-1,1), chunked so that a 5 GB state needs no second copy"""
    rng = np.random.default_rng(seed)
    if not np.any(U[:min(U.size, 1 << 20)]) and not np.any(U):
        U[:] = 0.1 * rng.uniform(-1, 1, U.size)
        return U
    for s in range(0, U.size, chunk):
        U[s:s + chunk] *= 1.0 + amp * rng.uniform(-1, 1, min(chunk, U.size - s))
    return U


def make_problem(mesh, fam, prob, recon):
    if fam == "diffreac2d":
        return pda.create_problem(mesh, prob)
    if fam == "advdiff2d":
        return pda.create_problem(mesh, prob, recon, pda.ViscousFluxReconstruction.FirstOrder)
    return pda.create_problem(mesh, prob, recon)


def full_lattice_velocity(fam, prob, recon, n, bounds, sten, per, order="fast", t=0.0):
    """velocity of a FULL lattice at its BASELINE size: CUDA (host-pointer C-ABI entry) vs the oracle in lattice mode"""
    mesh = pda.create_full_mesh(n, bounds, sten, per)
    p = make_problem(mesh, fam, prob, recon)
    p.setOption("velocity_order", order)
    U = p.initialCondition()
    perturb_inplace(U)
    V = p.createRightHandSide()
    t0 = time.time()
    p.rightHandSide(U, t, V)
    tg = time.time() - t0
    o = OracleProblem(None, fam, int(prob), int(recon), lattice=lattice_spec(n, bounds, sten, per), omp=True)
    Vo = np.empty_like(V)
    t0 = time.time()
    o.velocity(U, t, out=Vo)
    to = time.time() - t0
    st = err_stats(V, Vo)
    st.update(cells=int(np.prod(n)), gpu_s=tg, oracle_s=to, oracle_threads=o.num_threads(), order=order)
    return st


def sample_arrays(smesh):
    x, y, z = smesh._coords()
    return dict(dim=smesh.dimensionality(), stencil=smesh.stencilSize(), d=smesh._deltas()[0], graph=smesh.graph(),
                x=x, y=y, z=z)


def sample_mesh_case(fam, prob, recon, n, bounds, sten, gids, t, order="fast", full_state=None):
    """velocity + Jacobian of a SAMPLE mesh (the cfg 4 shape): CUDA vs the oracle on the same sample mesh arrays.
    Returns the stats of V (velocity entry), V2 (velocity from the Jacobian entry) and J."""
    full = pda.create_full_mesh(n, bounds, sten)
    smesh = pda.create_sample_mesh(full, gids)
    ps = make_problem(smesh, fam, prob, recon)
    ps.setOption("order", order)
    ndpc = ps.numDofPerCell()
    if full_state is None:
        pf = make_problem(full, fam, prob, recon)
        full_state = perturb_inplace(pf.initialCondition())
    sg = smesh.stencilMeshGids()
    Us = full_state.reshape(-1, ndpc)[sg].ravel().copy()
    o = OracleProblem(None, fam, int(prob), int(recon), arrays=sample_arrays(smesh), omp=True)
    V = ps.createRightHandSide()
    ps.rightHandSide(Us, t, V)
    J = ps.createJacobian()
    V2 = ps.createRightHandSide()
    ps.rightHandSideAndJacobian(Us, t, V2, J)
    rp, ci = o.pattern()
    assert np.array_equal(rp, J.indptr) and np.array_equal(ci, J.indices), "CSR pattern differs from the oracle"
    Vo = o.velocity(Us, t)
    Vo2, Jo = o.velocityAndJacobian(Us, t)
    return dict(V=err_stats(V, Vo), V2=err_stats(V2, Vo2), J=err_stats(J.data, Jo), cells=int(gids.size),
                nnz=int(Jo.size), order=order, Us=Us, smesh=smesh, Vs=V, Js=J)


def full_rows_vs_sample_oracle(fam, prob, recon, n, bounds, sten, frac, t=0.0, order="fast", seed=SEED):
    """cfg 2 shape: Jacobian (and velocity) of the FULL lattice at its BASELINE size, checked on a random `frac` of the
    rows PLUS every near-boundary row against the oracle evaluated on the sample mesh made of exactly those cells
    (the rule of /root/reference/tests_cpp/sample_mesh_compare.py:36-101: J_full[rows][:, stencil cols] == J_sample)."""
    full = pda.create_full_mesh(n, bounds, sten)
    pf = make_problem(full, fam, prob, recon)
    pf.setOption("order", order)
    ndpc = pf.numDofPerCell()
    ncell = int(np.prod(n))
    rng = np.random.default_rng(seed)
    pick = rng.choice(ncell, int(frac * ncell), replace=False)
    gids = np.unique(np.concatenate([pick, full.graphRowsOfCellsNearBd()])).astype(np.int32)
    smesh = pda.create_sample_mesh(full, gids)
    Uf = perturb_inplace(pf.initialCondition())
    sg = smesh.stencilMeshGids()
    Us = Uf.reshape(-1, ndpc)[sg].ravel().copy()
    o = OracleProblem(None, fam, int(prob), int(recon), arrays=sample_arrays(smesh), omp=True)
    Vo, Jo = o.velocityAndJacobian(Us, t)
    rp, ci = o.pattern()
    Vf = pf.createRightHandSide()
    Jf = pf.createJacobian()
    t0 = time.time()
    pf.rightHandSideAndJacobian(Uf, t, Vf, Jf)
    tg = time.time() - t0
    # the sampled rows of the full Jacobian, columns renumbered full gid -> stencil-mesh id, compared ENTRY BY ENTRY on
    # the sample mesh's pattern (both patterns hold the same cells: the stencil mesh contains every stencil neighbour)
    rows = (gids.astype(np.int64)[:, None] * ndpc + np.arange(ndpc)[None, :]).ravel()
    starts, ends = Jf.indptr[rows], Jf.indptr[rows + 1]
    assert np.array_equal(ends - starts, np.diff(rp)), "row lengths differ between the full and the sample pattern"
    idx = np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)]) if rows.size < 200000 else _ranges(starts, ends)
    cols_full = Jf.indices[idx].astype(np.int64)
    inv = np.full(ncell, -1, dtype=np.int64)
    inv[sg] = np.arange(sg.size)
    cols_s = inv[cols_full // ndpc] * ndpc + cols_full % ndpc
    assert np.array_equal(cols_s, ci), "columns differ between the full rows and the sample pattern"
    Jsub = Jf.data[idx]
    return dict(V=err_stats(Vf.reshape(-1, ndpc)[gids].ravel(), Vo), J=err_stats(Jsub, Jo), rows=int(gids.size),
                near_bd_rows=int(full.numCellsNearBd()), nnz_checked=int(Jo.size), nnz_full=int(Jf.data.size), gpu_s=tg,
                order=order, Jsub=Jsub, Jo=Jo, smesh=smesh, Us=Us)


def _ranges(starts, ends):
    """concatenated aranges without a Python loop"""
    lens = (ends - starts).astype(np.int64)
    tot = int(lens.sum())
    out = np.ones(tot, dtype=np.int64)
    pos = np.cumsum(lens)[:-1]
    out[0] = starts[0]
    out[pos] = starts[1:].astype(np.int64) - (ends[:-1].astype(np.int64) - 1)
    return np.cumsum(out)
