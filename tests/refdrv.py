"""ctypes front-end of the checkers under oracle/_ref (TEST INFRASTRUCTURE ONLY).

`RefProblem`   -> oracle/_ref/libpda_ref.so : the UNMODIFIED reference compiled from /root/reference (oracle/ref_driver.cc)
`OracleProblem`-> oracle/_ref/libpda_oracle.so : the plain-C restatement oracle/pda_oracle.c
Both expose the same methods so a parity test can be written once and pointed at either checker.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

FAM = {"euler1d": 1, "euler2d": 2, "euler3d": 3, "swe2d": 4, "diffreac2d": 5, "advdiff2d": 6, "advdiffreac2d": 7,
       "advection1d": 8, "diffreac1d": 9}


def ref_lib_path(omp=False):
    return os.path.join(REFDIR, "libpda_ref_omp.so" if omp else "libpda_ref.so")


def have_ref(omp=False):
    return os.path.exists(ref_lib_path(omp))


_libs = {}


def _ref(omp=False):
    key = ("ref", omp)
    if key not in _libs:
        L = C.CDLL(ref_lib_path(omp))
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.pdaref_last_error.restype = C.c_char_p
        L.pdaref_num_threads.restype = ci
        L.pdaref_create.restype = vp
        L.pdaref_create.argtypes = [C.c_char_p, ci, ci, ci, ci, ci, vp, vp]
        L.pdaref_destroy.argtypes = [vp]
        L.pdaref_query.restype = C.c_longlong
        L.pdaref_query.argtypes = [vp, ci]
        L.pdaref_mesh_arrays.argtypes = [vp] * 8
        L.pdaref_ic.argtypes = [vp, vp]
        L.pdaref_velocity.restype = ci
        L.pdaref_velocity.argtypes = [vp, vp, cd, vp]
        L.pdaref_velocity_and_jacobian.restype = ci
        L.pdaref_velocity_and_jacobian.argtypes = [vp, vp, cd, vp, vp]
        L.pdaref_pattern.argtypes = [vp, vp, vp]
        L.pdaref_apply_jacobian.restype = ci
        L.pdaref_apply_jacobian.argtypes = [vp, vp, vp, cd, vp]
        L.pdaref_ghosts.restype = ci
        L.pdaref_ghosts.argtypes = [vp, ci, vp]
        L.pdaref_time_velocity.restype = cd
        L.pdaref_time_velocity.argtypes = [vp, vp, cd, ci, ci]
        L.pdaref_time_jacobian.restype = cd
        L.pdaref_time_jacobian.argtypes = [vp, vp, cd, ci, ci]
        _libs[key] = L
    return _libs[key]


class RefProblem:
    """The reference's own problem object (mesh read from the text files in `meshDir`)."""

    def __init__(self, meshDir, family, probEnum, recon, icFlag=1, params=None, omp=False):
        self.L = _ref(omp)
        params = params or {}
        names = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
        vals = (C.c_double * max(1, len(params)))(*[float(v) for v in params.values()])
        fam = FAM[family] if isinstance(family, str) else int(family)
        self.h = self.L.pdaref_create(str(meshDir).encode(), fam, int(probEnum), int(recon), int(icFlag),
                                      len(params), names, vals)
        if not self.h:
            raise RuntimeError("reference: " + self.L.pdaref_last_error().decode())
        q = lambda i: int(self.L.pdaref_query(self.h, i))
        (self.dim, self.stencil, self.nSample, self.nStencil, self.ncols, self.nInner, self.nNearBd,
         self.periodic, self.ndpc, self.nDofStencil, self.nDofSample, self.nnz) = [q(i) for i in range(12)]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pdaref_destroy(self.h)
            self.h = None

    def num_threads(self):
        return int(self.L.pdaref_num_threads())

    def mesh_arrays(self):
        g = np.zeros((self.nSample, self.ncols), dtype=np.int32)
        x, y, z = (np.zeros(self.nStencil) for _ in range(3))
        ri = np.zeros(self.nInner, dtype=np.int32)
        rb = np.zeros(self.nNearBd, dtype=np.int32)
        d = np.zeros(6)
        self.L.pdaref_mesh_arrays(self.h, g.ctypes.data, x.ctypes.data, y.ctypes.data, z.ctypes.data,
                                  ri.ctypes.data if ri.size else None, rb.ctypes.data if rb.size else None,
                                  d.ctypes.data)
        return dict(graph=g, x=x, y=y, z=z, rowsInner=ri, rowsNearBd=rb, d=d[:3], dInv=d[3:])

    def initialCondition(self):
        U = np.zeros(self.nDofStencil)
        self.L.pdaref_ic(self.h, U.ctypes.data)
        return U

    def velocity(self, U, t=0.0):
        V = np.zeros(self.nDofSample)
        if self.L.pdaref_velocity(self.h, U.ctypes.data, float(t), V.ctypes.data):
            raise RuntimeError("reference: " + self.L.pdaref_last_error().decode())
        return V

    def velocityAndJacobian(self, U, t=0.0):
        V = np.zeros(self.nDofSample)
        vals = np.zeros(self.nnz)
        if self.L.pdaref_velocity_and_jacobian(self.h, U.ctypes.data, float(t), V.ctypes.data, vals.ctypes.data):
            raise RuntimeError("reference: " + self.L.pdaref_last_error().decode())
        return V, vals

    def pattern(self):
        rowptr = np.zeros(self.nDofSample + 1, dtype=np.int32)
        colidx = np.zeros(self.nnz, dtype=np.int32)
        self.L.pdaref_pattern(self.h, rowptr.ctypes.data, colidx.ctypes.data)
        return rowptr, colidx

    def applyJacobian(self, U, B, t=0.0):
        R = np.zeros(self.nDofSample)
        if self.L.pdaref_apply_jacobian(self.h, U.ctypes.data, B.ctypes.data, float(t), R.ctypes.data):
            raise RuntimeError("reference: " + self.L.pdaref_last_error().decode())
        return R

    def ghosts(self, side):
        n = self.L.pdaref_ghosts(self.h, side, None)
        if n < 0:
            return None
        out = np.zeros(max(n, 1))
        self.L.pdaref_ghosts(self.h, side, out.ctypes.data)
        return out[:n].reshape(self.nNearBd, -1) if self.nNearBd else out[:0]

    def time_velocity(self, U, t=0.0, warmup=1, reps=5):
        return float(self.L.pdaref_time_velocity(self.h, U.ctypes.data, float(t), warmup, reps))

    def time_jacobian(self, U, t=0.0, warmup=1, reps=3):
        return float(self.L.pdaref_time_jacobian(self.h, U.ctypes.data, float(t), warmup, reps))


# ------------------------------------------------------------------------------------------------ C restatement
def oracle_lib_path(omp=False):
    return os.path.join(REFDIR, "libpda_oracle_omp.so" if omp else "libpda_oracle.so")


def have_oracle(omp=False):
    return os.path.exists(oracle_lib_path(omp))


def _oracle(omp=False):
    key = ("oracle", omp)
    if key not in _libs:
        L = C.CDLL(oracle_lib_path(omp))
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.or_last_error.restype = C.c_char_p
        L.or_num_threads.restype = ci
        L.or_create.restype = vp
        L.or_create.argtypes = [C.c_char_p, ci, ci, ci, ci, ci, vp, vp]
        L.or_create_from_arrays.restype = vp
        L.or_create_from_arrays.argtypes = [ci, ci, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp]
        L.or_create_lattice.restype = vp
        L.or_create_lattice.argtypes = [ci, ci, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp]
        L.or_destroy.argtypes = [vp]
        L.or_set_source.argtypes = [vp, vp]
        L.or_query.restype = C.c_longlong
        L.or_query.argtypes = [vp, ci]
        L.or_mesh_arrays.argtypes = [vp] * 8
        L.or_ic.argtypes = [vp, vp]
        L.or_velocity.restype = ci
        L.or_velocity.argtypes = [vp, vp, cd, vp]
        L.or_velocity_and_jacobian.restype = ci
        L.or_velocity_and_jacobian.argtypes = [vp, vp, cd, vp, vp]
        L.or_pattern.argtypes = [vp, vp, vp]
        L.or_ghosts.restype = ci
        L.or_ghosts.argtypes = [vp, ci, vp]
        L.or_time_velocity.restype = cd
        L.or_time_velocity.argtypes = [vp, vp, cd, ci, ci]
        L.or_time_velocity_inner_range.restype = cd
        L.or_time_velocity_inner_range.argtypes = [vp, vp, cd, vp, C.c_int32, C.c_int32, ci]
        L.or_weno5.argtypes = [vp, vp] + [cd] * 6
        L.or_weno3.argtypes = [vp, vp] + [cd] * 4
        L.or_weno5_grad.argtypes = [vp] * 4 + [cd] * 6
        L.or_weno3_grad.argtypes = [vp] * 4 + [cd] * 4
        L.or_euler_flux.argtypes = [ci, vp, vp, vp, vp, cd]
        L.or_euler_flux_jac.argtypes = [ci, vp, vp, vp, vp, vp, cd]
        L.or_swe_flux.argtypes = [vp, vp, vp, vp, cd]
        L.or_swe_flux_jac.argtypes = [vp, vp, vp, vp, vp, cd]
        _libs[key] = L
    return _libs[key]


class OracleProblem:
    """Plain-C restatement (oracle/pda_oracle.c); same interface as RefProblem."""

    def __init__(self, meshDir, family, probEnum, recon, icFlag=1, params=None, omp=False, arrays=None, lattice=None):
        """mesh: a directory in the reference's text format, `arrays` (graph + coordinates), or `lattice` = dict(dim,
        stencil, n, d, periodic[, cx, cy, cz]): a full mesh in natural ordering WITHOUT a stored graph -- the way the
        BASELINE sizes (512^3, 4096^2) are evaluated (see lattice_spec())."""
        self.L = _oracle(omp)
        params = params or {}
        names = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
        vals = (C.c_double * max(1, len(params)))(*[float(v) for v in params.values()])
        fam = FAM[family] if isinstance(family, str) else int(family)
        self._lazy_nnz = lattice is not None
        if lattice is not None:
            a = lattice
            n3 = (C.c_int32 * 3)(*(list(a["n"]) + [1, 1, 1])[:3])
            d3 = (C.c_double * 3)(*(list(a["d"]) + [0.0, 0.0, 0.0])[:3])
            per = (C.c_int32 * 3)(*(list(a["periodic"]) + [1, 1, 1])[:3])
            ax = [None if a.get(k) is None else np.ascontiguousarray(a[k], dtype=np.float64) for k in ("cx", "cy", "cz")]
            self.h = self.L.or_create_lattice(int(a["dim"]), int(a["stencil"]), n3, d3, per,
                                              *[None if v is None else v.ctypes.data for v in ax], fam,
                                              int(probEnum), int(recon), int(icFlag), len(params), names, vals)
        elif arrays is not None:
            a = arrays
            d = np.ascontiguousarray(a["d"], dtype=np.float64)
            g = np.ascontiguousarray(a["graph"], dtype=np.int32)
            x, y, z = (np.ascontiguousarray(a[k], dtype=np.float64) for k in ("x", "y", "z"))
            self.h = self.L.or_create_from_arrays(a["dim"], a["stencil"], g.shape[0], x.size, d.ctypes.data,
                                                  x.ctypes.data, y.ctypes.data, z.ctypes.data, g.ctypes.data, fam,
                                                  int(probEnum), int(recon), int(icFlag), len(params), names, vals)
        else:
            self.h = self.L.or_create(str(meshDir).encode(), fam, int(probEnum), int(recon), int(icFlag),
                                      len(params), names, vals)
        if not self.h:
            raise RuntimeError("oracle: " + self.L.or_last_error().decode())
        q = lambda i: int(self.L.or_query(self.h, i))
        (self.dim, self.stencil, self.nSample, self.nStencil, self.ncols, self.nInner, self.nNearBd,
         self.periodic, self.ndpc, self.nDofStencil, self.nDofSample) = [q(i) for i in range(11)]
        self._nnz = None if self._lazy_nnz else q(11)   # lattice mode builds the pattern on demand only

    @property
    def nnz(self):
        if self._nnz is None:
            self._nnz = int(self.L.or_query(self.h, 11))
        return self._nnz

    def __del__(self):
        if getattr(self, "h", None):
            self.L.or_destroy(self.h)
            self.h = None

    def num_threads(self):
        return int(self.L.or_num_threads())

    def mesh_arrays(self):
        g = np.zeros((self.nSample, self.ncols), dtype=np.int32)
        x, y, z = (np.zeros(self.nStencil) for _ in range(3))
        ri = np.zeros(max(self.nInner, 1), dtype=np.int32)
        rb = np.zeros(max(self.nNearBd, 1), dtype=np.int32)
        d = np.zeros(6)
        self.L.or_mesh_arrays(self.h, g.ctypes.data, x.ctypes.data, y.ctypes.data, z.ctypes.data, ri.ctypes.data,
                              rb.ctypes.data, d.ctypes.data)
        return dict(graph=g, x=x, y=y, z=z, rowsInner=ri[:self.nInner], rowsNearBd=rb[:self.nNearBd], d=d[:3],
                    dInv=d[3:])

    def setSource(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.size == self.nSample
        self.L.or_set_source(self.h, v.ctypes.data)

    def initialCondition(self):
        U = np.zeros(self.nDofStencil)
        self.L.or_ic(self.h, U.ctypes.data)
        return U

    def velocity(self, U, t=0.0, out=None):
        V = np.zeros(self.nDofSample) if out is None else out
        self.L.or_velocity(self.h, U.ctypes.data, float(t), V.ctypes.data)
        return V

    def velocityAndJacobian(self, U, t=0.0):
        V = np.zeros(self.nDofSample)
        vals = np.zeros(self.nnz)
        self.L.or_velocity_and_jacobian(self.h, U.ctypes.data, float(t), V.ctypes.data, vals.ctypes.data)
        return V, vals

    def pattern(self):
        rowptr = np.zeros(self.nDofSample + 1, dtype=np.int32)
        colidx = np.zeros(self.nnz, dtype=np.int32)
        self.L.or_pattern(self.h, rowptr.ctypes.data, colidx.ctypes.data)
        return rowptr, colidx

    def ghosts(self, side):
        n = self.L.or_ghosts(self.h, side, None)
        out = np.zeros(max(n, 1))
        self.L.or_ghosts(self.h, side, out.ctypes.data)
        return out[:n].reshape(self.nNearBd, -1) if self.nNearBd else out[:0]

    def time_velocity(self, U, t=0.0, warmup=1, reps=5):
        return float(self.L.or_time_velocity(self.h, U.ctypes.data, float(t), warmup, reps))

    def time_velocity_inner_range(self, U, V, it0, it1, t=0.0, reps=1):
        """seconds for `reps` evaluations of the inner rows [it0, it1) (a bounded sample of a big workload)"""
        return float(self.L.or_time_velocity_inner_range(self.h, U.ctypes.data, float(t), V.ctypes.data, int(it0), int(it1), int(reps)))


def oracle_leaf():
    return _oracle(False)


# ------------------------------------------------------------------------------------------------ 80-bit restatement
def have_oracle_ld():
    return os.path.exists(os.path.join(REFDIR, "libpda_oracle_ld.so"))


def exact_velocity_and_jacobian(meshDir, family, probEnum, recon, icFlag, params, U, t):
    """The C restatement compiled in 80-bit long double (oracle/Makefile: libpda_oracle_ld.so): V and J accurate to
    ~1e-19, returned rounded to double.  Used ONLY to bound the reference's own rounding noise in the WENO Jacobians
    (tools/jacobian_noise.py, DESIGN.md 'Jacobian tolerance')."""
    LD = np.longdouble
    L = C.CDLL(os.path.join(REFDIR, "libpda_oracle_ld.so"))
    L.or_create.restype = C.c_void_p
    L.or_create.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_void_p] * 2
    L.or_query.restype = C.c_longlong
    L.or_query.argtypes = [C.c_void_p, C.c_int]
    L.or_destroy.argtypes = [C.c_void_p]
    params = params or {}
    names = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
    vals = np.array([float(v) for v in params.values()] or [0.0], dtype=LD)
    fam = FAM[family] if isinstance(family, str) else int(family)
    h = L.or_create(str(meshDir).encode(), fam, int(probEnum), int(recon), int(icFlag), len(params), names,
                    vals.ctypes.data)
    if not h:
        raise RuntimeError("oracle_ld: create failed")
    nnz, nv = L.or_query(h, 11), L.or_query(h, 10)
    Ul = np.ascontiguousarray(U, dtype=LD)
    V = np.zeros(nv, dtype=LD)
    J = np.zeros(nnz, dtype=LD)
    L.or_velocity_and_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_longdouble, C.c_void_p, C.c_void_p]
    L.or_velocity_and_jacobian(h, Ul.ctypes.data, C.c_longdouble(t), V.ctypes.data, J.ctypes.data)
    L.or_destroy(h)
    return V.astype(np.float64), J.astype(np.float64)


# ------------------------------------------------------------------------------------ boundary-face gradients
def ref_grad_lib_path():
    return os.path.join(REFDIR, "libpda_ref_grad.so")


def have_ref_grad():
    return os.path.exists(ref_grad_lib_path())


def _ref_grad():
    key = ("refgrad", False)
    if key not in _libs:
        L = C.CDLL(ref_grad_lib_path())
        vp, ci = C.c_void_p, C.c_int
        L.pdaref_grad_last_error.restype = C.c_char_p
        L.pdaref_rows_strictly_on_bd.restype = C.c_int64
        L.pdaref_rows_strictly_on_bd.argtypes = [C.c_char_p, vp]
        L.pdaref_grad_faces.restype = C.c_int64
        L.pdaref_grad_faces.argtypes = [C.c_char_p, vp, vp, vp]
        L.pdaref_grad_eval.restype = ci
        L.pdaref_grad_eval.argtypes = [C.c_char_p, ci, ci, vp, vp, vp, vp]
        _libs[key] = L
    return _libs[key]


def ref_rows_strictly_on_bd(meshDir):
    """graphRowsOfCellsStrictlyOnBd() of the unmodified reference's mesh"""
    L = _ref_grad()
    n = L.pdaref_rows_strictly_on_bd(str(meshDir).encode(), None)
    if n < 0:
        raise RuntimeError("reference: " + L.pdaref_grad_last_error().decode())
    r = np.zeros(max(n, 1), dtype=np.int32)
    L.pdaref_rows_strictly_on_bd(str(meshDir).encode(), r.ctypes.data)
    return r[:n]


def ref_gradient(meshDir, field, ndpc=1, scalar_api=False):
    """The unmodified reference's GradientEvaluator on the mesh in `meshDir`: dict(cellGid, position, parentRow,
    normalDir, centers [n][3], grad [n][ndpc]); faces in the order of tests_cpp/gradients/main.cc:48-97."""
    L = _ref_grad()
    d = str(meshDir).encode()
    n = L.pdaref_grad_faces(d, None, None, None)
    if n < 0:
        raise RuntimeError("reference: " + L.pdaref_grad_last_error().decode())
    gid, pos, row, nd = (np.zeros(max(n, 1), dtype=np.int32) for _ in range(4))
    L.pdaref_grad_faces(d, gid.ctypes.data, pos.ctypes.data, row.ctypes.data)
    grad = np.zeros((max(n, 1), ndpc))
    cen = np.zeros((max(n, 1), 3))
    f = np.ascontiguousarray(field, dtype=np.float64)
    if L.pdaref_grad_eval(d, int(ndpc), int(bool(scalar_api)), f.ctypes.data, grad.ctypes.data, cen.ctypes.data,
                          nd.ctypes.data):
        raise RuntimeError("reference: " + L.pdaref_grad_last_error().decode())
    return dict(cellGid=gid[:n], position=pos[:n], parentRow=row[:n], normalDir=nd[:n], centers=cen[:n], grad=grad[:n])


def oracle_gradient(stencil, graph, rowsNearBd, x, y, z, dx, dy, field, ndpc=1):
    """The C restatement (oracle/pda_oracle.c: or_gradient_faces / or_gradient_eval) on mesh arrays; same dict"""
    L = _oracle()
    L.or_gradient_faces.restype = C.c_int64
    L.or_gradient_faces.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_double, C.c_double] + [C.c_void_p] * 5
    L.or_gradient_eval.restype = None
    L.or_gradient_eval.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                   C.c_void_p, C.c_int, C.c_void_p]
    graph = np.ascontiguousarray(graph, dtype=np.int32)
    rows = np.ascontiguousarray(rowsNearBd, dtype=np.int32)
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    head = (int(stencil), graph.ctypes.data, rows.ctypes.data if rows.size else None, int(rows.size), x.ctypes.data,
            y.ctypes.data, z.ctypes.data, float(dx), float(dy))
    n = L.or_gradient_faces(*head, None, None, None, None, None)
    gid, pos, row, nd = (np.zeros(max(n, 1), dtype=np.int32) for _ in range(4))
    cen = np.zeros((max(n, 1), 3))
    L.or_gradient_faces(*head, gid.ctypes.data, pos.ctypes.data, row.ctypes.data, nd.ctypes.data, cen.ctypes.data)
    grad = np.zeros((max(n, 1), ndpc))
    f = np.ascontiguousarray(field, dtype=np.float64)
    L.or_gradient_eval(int(stencil), graph.ctypes.data, n, pos.ctypes.data, row.ctypes.data, float(dx), float(dy),
                       f.ctypes.data, int(ndpc), grad.ctypes.data)
    return dict(cellGid=gid[:n], position=pos[:n], parentRow=row[:n], normalDir=nd[:n], centers=cen[:n], grad=grad[:n])


def lattice_spec(n, bounds, stencil, periodic=()):
    """dict for OracleProblem(lattice=...): cells per axis, dx and per-axis centre coordinates exactly as the reference's
    mesh files carry them -- create_full_mesh.py writes dx and every coordinate with "%.14f" and the C++ loader reads
    those roundings back (meshing_scripts/create_full_mesh.py:151-218, impl/mesh_read_info.hpp:84-99)."""
    n = list(n)
    dim = len(n)
    if dim == 2 and n[1] == 1:
        dim, n = 1, n[:1]
    d, axes = [], []
    for a in range(dim):
        lo, hi = float(bounds[2 * a]), float(bounds[2 * a + 1])
        dx = (hi - lo) / n[a]
        ox = lo + 0.5 * dx   # natural_order_mesh_{2,3}d.py: ox = lo + 0.5*dx, x = ox + gi*dx (unrounded dx)
        axes.append(np.array([float("%.14f" % (ox + i * dx)) for i in range(n[a])]))
        d.append(float("%.14f" % dx))
    while len(axes) < 3:
        axes.append(None)
    return dict(dim=dim, stencil=int(stencil), n=n, d=d, periodic=[1 if a in periodic else 0 for a in ("x", "y", "z")[:dim]],
                cx=axes[0], cy=axes[1], cz=axes[2])
