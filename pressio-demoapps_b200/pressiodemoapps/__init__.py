"""pressiodemoapps (B200 engine) -- Python mirror of the reference's `pressiodemoapps` module.

Same names, argument meaning and error behaviour as the reference's pybind11 module
(`/root/reference/src_py/main_binder.cc`, `pressiodemoapps/__init__.py`), but every evaluation runs as CUDA kernels
behind the C-ABI in `include/pda_b200.h` (loaded with ctypes from `../lib/libpda_b200.so`).  There is no CPU
fallback: `rightHandSide` & friends raise if the library or a CUDA device is missing.

Extras beyond the reference's Python surface (which only exposes `applyJacobian`):
`createJacobian()/jacobian()/rightHandSideAndJacobian()` (the C++ API, adapter_cpp.hpp:105-229), native mesh tools
(`create_full_mesh`, `create_sample_mesh`, `mesh.write`) and device-pointer entry points for callers that keep the
state in HBM (`rightHandSideDevice`).
"""
import ctypes as _C
import os as _os

import numpy as _np

from enum import IntEnum as _IntEnum

__all__ = [
    "CellCenteredUniformMesh", "load_cellcentered_uniform_mesh", "create_full_mesh", "create_sample_mesh", "mesh_from_arrays",
    "create_slab_window_mesh",
    "InviscidFluxReconstruction", "InviscidFluxScheme", "ViscousFluxReconstruction", "ViscousFluxScheme",
    "Euler1d", "Euler2d", "Euler3d", "Swe2d", "DiffusionReaction1d", "DiffusionReaction2d", "AdvectionDiffusion2d",
    "AdvectionDiffusionReaction2d", "Advection1d",
    "create_problem", "create_gray_scott_2d_problem", "create_slip_wall_swe_2d_problem", "create_cross_shock_problem",
    "create_linear_advection_1d_problem", "create_diffusion_reaction_1d_problem_A", "create_burgers_2d_problem",
    "create_diffusion_reaction_2d_problem_A", "create_adv_diff_reac_2d_problem_A",
    "advanceRK2", "advanceRK4", "advanceSSP3", "PdaError", "device_count", "BC",
    "FacePosition", "GradientEvaluator",
]

__version__ = "0.16.0"   # the reference release this module mirrors (/root/reference/version.txt)

_HERE = _os.path.dirname(_os.path.abspath(__file__))
_LIBPATH = _os.path.join(_os.path.dirname(_HERE), "lib", "libpda_b200.so")


_STEPPERS = {"euler": 0, "rk2": 1, "rk4": 2, "ssprk3": 3}
_GHOST_FN = _C.CFUNCTYPE(None, _C.c_void_p, _C.c_int32, _C.POINTER(_C.c_int32), _C.c_double, _C.c_double,
                         _C.POINTER(_C.c_double), _C.c_int, _C.c_double, _C.POINTER(_C.c_double))
_FACTOR_FN = _C.CFUNCTYPE(None, _C.c_void_p, _C.POINTER(_C.c_int32), _C.c_double, _C.c_double, _C.c_int,
                          _C.POINTER(_C.c_double))


class PdaError(RuntimeError):
    """Raised where the reference throws std::runtime_error (or calls exit())."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _load():
    if not _os.path.exists(_LIBPATH):
        raise ImportError(
            "libpda_b200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
            "at the repo root -- there is no Python/CPU fallback" % _LIBPATH)
    return _C.CDLL(_LIBPATH)


_lib = _load()
_vp, _i32, _i64, _dbl, _cp = _C.c_void_p, _C.c_int32, _C.c_int64, _C.c_double, _C.c_char_p
_pi32 = _C.POINTER(_C.c_int32)


def _sig(name, restype, *argtypes):
    f = getattr(_lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("pda_last_error", _cp)
_sig("pda_version", _cp)
_sig("pda_device_count", _C.c_int)
_sig("pda_measure_fp64_peak", _C.c_int, _C.c_int, _C.POINTER(_dbl), _vp)
_sig("pda_measure_fp64_peak_ex", _C.c_int, _C.c_int, _C.POINTER(_dbl), _C.POINTER(_dbl), _C.POINTER(_dbl))
_sig("pda_test_glibc_pow", _C.c_int, _C.c_int, _vp, _dbl, _vp, _i64)
_sig("pda_mesh_load", _C.c_int, _cp, _C.POINTER(_vp))
_sig("pda_mesh_make_lattice", _C.c_int, _C.c_int, _vp, _vp, _vp, _C.c_int, _C.POINTER(_vp))
_sig("pda_mesh_make_sample", _C.c_int, _vp, _vp, _i64, _C.POINTER(_vp))
_sig("pda_mesh_make_slab_window", _C.c_int, _vp, _C.c_int, _C.c_int, _C.POINTER(_vp))
_sig("pda_mesh_slab_window_info", _C.c_int, _vp, _vp)
_sig("pda_mesh_from_arrays", _C.c_int, _C.c_int, _C.c_int, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _C.POINTER(_vp))
_sig("pda_mesh_write", _C.c_int, _vp, _cp)
_sig("pda_mesh_free", _C.c_int, _vp)
for _n in ("dimensionality", "stencil_size", "graph_cols", "is_fully_periodic", "is_lattice"):
    _sig("pda_mesh_" + _n, _C.c_int, _vp)
for _n in ("stencil_mesh_size", "sample_mesh_size", "num_cells_inner", "num_cells_near_bd"):
    _sig("pda_mesh_" + _n, _i32, _vp)
_sig("pda_mesh_deltas", _C.c_int, _vp, _vp, _vp)
_sig("pda_mesh_coordinates", _C.c_int, _vp, _vp, _vp, _vp)
_sig("pda_mesh_graph", _C.c_int, _vp, _vp)
_sig("pda_mesh_rows_inner", _C.c_int, _vp, _vp)
_sig("pda_mesh_rows_near_bd", _C.c_int, _vp, _vp)
_sig("pda_mesh_stencil_gids", _C.c_int, _vp, _vp)
_sig("pda_mesh_num_cells_strictly_on_bd", _i32, _vp)
_sig("pda_mesh_rows_strictly_on_bd", _C.c_int, _vp, _vp)
_sig("pda_gradient_create", _C.c_int, _vp, _C.c_int, _C.POINTER(_vp))
_sig("pda_gradient_free", _C.c_int, _vp)
_sig("pda_gradient_num_faces", _i32, _vp)
_sig("pda_gradient_launch_count", _i64, _vp)
_sig("pda_gradient_faces", _C.c_int, _vp, _vp, _vp, _vp, _vp, _vp)
_sig("pda_gradient_query_face", _C.c_int, _vp, _i32, _C.c_int, _C.POINTER(_i32))
_sig("pda_gradient_compute_host", _C.c_int, _vp, _vp, _C.c_int, _vp)
_sig("pda_gradient_compute_dev", _C.c_int, _vp, _vp, _C.c_int, _vp, _vp)
_sig("pda_problem_create", _C.c_int, _vp, _C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_int, _vp, _vp, _C.c_int,
     _C.POINTER(_vp))
_sig("pda_problem_set_bc", _C.c_int, _vp, _C.c_int, _C.c_int, _vp)
_sig("pda_problem_set_source", _C.c_int, _vp, _vp)
_sig("pda_problem_set_option", _C.c_int, _vp, _cp, _cp)
_sig("pda_problem_get_option", _C.c_int, _vp, _cp, _vp, _C.c_int)
_sig("pda_problem_free", _C.c_int, _vp)
_sig("pda_problem_num_dof_per_cell", _C.c_int, _vp)
_sig("pda_problem_total_dof_sample_mesh", _i32, _vp)
_sig("pda_problem_total_dof_stencil_mesh", _i32, _vp)
_sig("pda_problem_query_parameter", _C.c_int, _vp, _cp, _C.POINTER(_dbl))
_sig("pda_problem_initial_condition", _C.c_int, _vp, _vp)
_sig("pda_problem_jacobian_nnz", _C.c_int, _vp, _C.POINTER(_i64))
_sig("pda_problem_jacobian_pattern", _C.c_int, _vp, _vp, _vp)
_sig("pda_problem_velocity_host", _C.c_int, _vp, _vp, _dbl, _vp)
_sig("pda_problem_velocity_and_jacobian_host", _C.c_int, _vp, _vp, _dbl, _vp, _vp)
_sig("pda_problem_apply_jacobian_host", _C.c_int, _vp, _vp, _vp, _C.c_int, _C.c_int, _dbl, _vp)
_sig("pda_problem_velocity_dev", _C.c_int, _vp, _vp, _dbl, _vp, _vp)
_sig("pda_problem_velocity_and_jacobian_dev", _C.c_int, _vp, _vp, _dbl, _vp, _vp, _vp)
_sig("pda_problem_apply_jacobian_dev", _C.c_int, _vp, _vp, _vp, _C.c_int, _C.c_int, _dbl, _vp, _vp)
_sig("pda_problem_ghosts", _C.c_int, _vp, _C.c_int, _vp)
_sig("pda_problem_launch_count", _i64, _vp)
_sig("pda_problem_create_slab", _C.c_int, _vp, _C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_int,
     _C.POINTER(_vp))
_sig("pda_slab_extent", _C.c_int, _vp, _C.POINTER(_i32), _C.POINTER(_i32), _C.POINTER(_i32), _C.POINTER(_i64))
_sig("pda_slab_initial_condition", _C.c_int, _vp, _vp)
_sig("pda_slab_velocity_interior_dev", _C.c_int, _vp, _vp, _dbl, _vp, _vp)
_sig("pda_slab_velocity_boundary_dev", _C.c_int, _vp, _vp, _dbl, _vp, _vp)
_sig("pda_problem_set_bc_callback", _C.c_int, _vp, _C.c_int, _GHOST_FN, _FACTOR_FN, _vp)
_sig("pda_problem_set_bc_pointer", _C.c_int, _vp, _C.c_int, _vp)
_sig("pda_problem_advance_dev", _C.c_int, _vp, _C.c_int, _vp, _dbl, _dbl, _C.c_int32, _vp)
_sig("pda_problem_advance_host", _C.c_int, _vp, _C.c_int, _vp, _dbl, _dbl, _C.c_int32)
_sig("pda_slab_peer_handle", _C.c_int, _vp, _vp)
_sig("pda_slab_peer_connect", _C.c_int, _vp, _vp)
_sig("pda_slab_peer_connect_local", _C.c_int, _vp, _vp, _vp)
_sig("pda_slab_velocity_peer_dev", _C.c_int, _vp, _vp, _dbl, _vp, _vp)
_sig("pda_slab_velocity_peer_host", _C.c_int, _vp, _vp, _dbl, _vp)


def _check(status):
    if status != 0:
        raise PdaError(status, _lib.pda_last_error().decode())


def device_count():
    return int(_lib.pda_device_count())


def measure_fp64_peak(device=0):
    """peak FP64 FMA throughput (TFLOP/s) of `device` from a pure DFMA loop: the FP64 roofline denominator"""
    v = _dbl()
    _check(_lib.pda_measure_fp64_peak(int(device), _C.byref(v), None))
    return v.value


def measure_fp64_peak_ex(device=0):
    """(TFLOP/s, SM MHz the probe ran at, DFMA lanes issued per SM and cycle): the probe with its own clock evidence"""
    v, mhz, rate = _dbl(), _dbl(), _dbl()
    _check(_lib.pda_measure_fp64_peak_ex(int(device), _C.byref(v), _C.byref(mhz), _C.byref(rate)))
    return v.value, mhz.value, rate.value


def _device_glibc_pow(x, y, device=0):
    """test hook: the device restatement of glibc's pow (csrc/glibc_pow.h) applied to a host array"""
    x = _np.ascontiguousarray(x, dtype=_np.float64)
    out = _np.zeros_like(x)
    _check(_lib.pda_test_glibc_pow(int(device), x.ctypes.data, float(y), out.ctypes.data, x.size))
    return out


def _f64(a, n=None, name="array"):
    """float64, contiguous (C order, or F order for a 2-D operand) numpy array -- what the reference's pybind/Eigen::Ref
    binding accepts (adapter_py.hpp); a strided view is refused instead of reading or writing the wrong memory."""
    if not isinstance(a, _np.ndarray) or a.dtype != _np.float64:
        raise TypeError("%s must be a float64 numpy array" % name)
    if not (a.flags["C_CONTIGUOUS"] or (a.ndim == 2 and a.flags["F_CONTIGUOUS"])):
        raise TypeError("%s must be contiguous (got a strided view)" % name)
    if n is not None and a.size != n:
        raise ValueError("%s has %d entries, expected %d" % (name, a.size, n))
    return a


def _f64_in(a, n=None, name="array"):
    """INPUT arguments (state, applyJacobian operand): the reference binds them as `const Eigen::Ref<const ...>&`
    (adapter_py.hpp:153-185), for which pybind11 converts -- an integer array (tests_py/1d_linear_adv passes
    np.arange(n)) or a strided view is copied into a contiguous float64 temporary.  Outputs stay strict (_f64): a copy
    there would silently drop the result."""
    if isinstance(a, _np.ndarray) and a.dtype == _np.float64 and \
            (a.flags["C_CONTIGUOUS"] or (a.ndim == 2 and a.flags["F_CONTIGUOUS"])):
        b = a
    else:
        try:
            b = _np.array(a, dtype=_np.float64, order="A")
        except (TypeError, ValueError):
            raise TypeError("%s must be convertible to a float64 numpy array" % name)
        if not (b.flags["C_CONTIGUOUS"] or (b.ndim == 2 and b.flags["F_CONTIGUOUS"])):
            b = _np.ascontiguousarray(b)
    if n is not None and b.size != n:
        raise ValueError("%s has %d entries, expected %d" % (name, b.size, n))
    return b


# ------------------------------------------------------------------------------------------------ enums
class InviscidFluxReconstruction(_IntEnum):
    FirstOrder = 0
    Weno3 = 1
    Weno5 = 2


class InviscidFluxScheme(_IntEnum):
    Rusanov = 0


class ViscousFluxReconstruction(_IntEnum):
    FirstOrder = 0


class ViscousFluxScheme(_IntEnum):
    Central = 0


class Euler1d(_IntEnum):
    PeriodicSmooth = 0
    Sod = 1
    Lax = 2
    ShuOsher = 3


class Euler2d(_IntEnum):
    PeriodicSmooth = 0
    KelvinHelmholtz = 1
    SedovFull = 2
    SedovSymmetry = 3
    Riemann = 4
    NormalShock = 5
    DoubleMachReflection = 6
    CrossShock = 7
    testingonlyneumann = 8


class Euler3d(_IntEnum):
    PeriodicSmooth = 0
    SedovSymmetry = 1


class Swe2d(_IntEnum):
    SlipWall = 0
    CustomBCs = 1


class DiffusionReaction2d(_IntEnum):
    ProblemA = 0
    GrayScott = 1


class AdvectionDiffusion2d(_IntEnum):
    BurgersPeriodic = 0
    BurgersOutflow = 1


class AdvectionDiffusionReaction2d(_IntEnum):
    ProblemA = 0


class Advection1d(_IntEnum):
    PeriodicLinear = 0


class DiffusionReaction1d(_IntEnum):
    ProblemA = 0


class BC(_IntEnum):
    """Device-expressible custom boundary rules (custom_bcs_functions.hpp)."""
    Dirichlet = 0
    HomogNeumann = 1
    Reflective = 2


_FAMILY = {Euler1d: 1, Euler2d: 2, Euler3d: 3, Swe2d: 4, DiffusionReaction2d: 5, AdvectionDiffusion2d: 6,
           AdvectionDiffusionReaction2d: 7, Advection1d: 8, DiffusionReaction1d: 9}


# ------------------------------------------------------------------------------------------------ mesh
class CellCenteredUniformMesh:
    """impl/mesh_ccu.hpp:67-473; python methods of src_py/main_binder.cc:230-243 (+ the C++-only row lists)."""

    def __init__(self, meshDir=None, _handle=None):
        if _handle is None:
            h = _vp()
            _check(_lib.pda_mesh_load(str(meshDir).encode(), _C.byref(h)))
            _handle = h
        self._h = _handle
        self._keep = []

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:   # module globals are gone at interpreter shutdown
            _lib.pda_mesh_free(h)
            self._h = None

    def dimensionality(self): return _lib.pda_mesh_dimensionality(self._h)
    def stencilMeshSize(self): return _lib.pda_mesh_stencil_mesh_size(self._h)
    def sampleMeshSize(self): return _lib.pda_mesh_sample_mesh_size(self._h)
    def stencilSize(self): return _lib.pda_mesh_stencil_size(self._h)
    def isFullyPeriodic(self): return bool(_lib.pda_mesh_is_fully_periodic(self._h))
    def isLattice(self): return bool(_lib.pda_mesh_is_lattice(self._h))
    def numCellsInner(self): return _lib.pda_mesh_num_cells_inner(self._h)
    def numCellsNearBd(self): return _lib.pda_mesh_num_cells_near_bd(self._h)

    def _deltas(self):
        d = _np.zeros(3)
        di = _np.zeros(3)
        _check(_lib.pda_mesh_deltas(self._h, d.ctypes.data, di.ctypes.data))
        return d, di

    def dx(self): return float(self._deltas()[0][0])
    def dy(self): return float(self._deltas()[0][1])
    def dz(self): return float(self._deltas()[0][2])
    def dxInv(self): return float(self._deltas()[1][0])
    def dyInv(self): return float(self._deltas()[1][1])
    def dzInv(self): return float(self._deltas()[1][2])

    def _coords(self):
        n = self.stencilMeshSize()
        x, y, z = _np.zeros(n), _np.zeros(n), _np.zeros(n)
        _check(_lib.pda_mesh_coordinates(self._h, x.ctypes.data, y.ctypes.data, z.ctypes.data))
        return x, y, z

    def viewX(self): return self._coords()[0]
    def viewY(self): return self._coords()[1]
    def viewZ(self): return self._coords()[2]

    def graph(self):
        g = _np.zeros((self.sampleMeshSize(), _lib.pda_mesh_graph_cols(self._h)), dtype=_np.int32)
        _check(_lib.pda_mesh_graph(self._h, g.ctypes.data))
        return g

    def graphRowsOfCellsAwayFromBd(self):
        r = _np.zeros(self.numCellsInner(), dtype=_np.int32)
        _check(_lib.pda_mesh_rows_inner(self._h, r.ctypes.data))
        return r

    def graphRowsOfCellsNearBd(self):
        r = _np.zeros(self.numCellsNearBd(), dtype=_np.int32)
        _check(_lib.pda_mesh_rows_near_bd(self._h, r.ctypes.data))
        return r

    def graphRowsOfCellsStrictlyOnBd(self):
        """impl/mesh_ccu.hpp:153-155: rows of the cells with at least one face on the domain boundary (2D meshes)."""
        r = _np.zeros(max(0, _lib.pda_mesh_num_cells_strictly_on_bd(self._h)), dtype=_np.int32)
        _check(_lib.pda_mesh_rows_strictly_on_bd(self._h, r.ctypes.data))
        return r

    def stencilMeshGids(self):
        r = _np.zeros(self.stencilMeshSize(), dtype=_np.int32)
        _check(_lib.pda_mesh_stencil_gids(self._h, r.ctypes.data))
        return r

    def write(self, outDir):
        """info.dat / connectivity.dat / coordinates.dat in the reference's text format."""
        _os.makedirs(outDir, exist_ok=True)
        _check(_lib.pda_mesh_write(self._h, str(outDir).encode()))


def load_cellcentered_uniform_mesh(meshDir):
    """mesh.hpp:87-91 / src_py/main_binder.cc:246-250."""
    return CellCenteredUniformMesh(meshDir)


def create_full_mesh(numCells, bounds, stencilSize=3, periodic=()):
    """Native `meshing_scripts/create_full_mesh.py -n .. --bounds .. -s .. --periodic ..` (in memory)."""
    numCells = list(numCells)
    dim = len(numCells)
    if dim == 2 and numCells[1] == 1:
        dim = 1
    n = (_C.c_int32 * 3)(*(numCells + [1, 1, 1])[:3])
    b = list(bounds) + [0.0] * 6
    bd = (_C.c_double * 6)(*b[:6])
    per = (_C.c_int32 * 3)(*[1 if a in periodic else 0 for a in ("x", "y", "z")])
    h = _vp()
    _check(_lib.pda_mesh_make_lattice(dim, n, bd, per, int(stencilSize), _C.byref(h)))
    return CellCenteredUniformMesh(_handle=h)


def mesh_from_arrays(dim, stencilSize, dxyz, x, y, z, graph, detect_lattice=False):
    """Mesh from caller arrays (graph row-major [nSample][(stencil-1)*dim+1]); always evaluated by the graph-driven
    kernels (no lattice detection)."""
    g = _np.ascontiguousarray(graph, dtype=_np.int32)
    x, y, z = (_np.ascontiguousarray(a, dtype=_np.float64) for a in (x, y, z))
    d = _np.ascontiguousarray(dxyz, dtype=_np.float64)
    h = _vp()
    _check(_lib.pda_mesh_from_arrays(int(dim), int(stencilSize), g.shape[0], x.size, d.ctypes.data, x.ctypes.data,
                                     y.ctypes.data, z.ctypes.data, g.ctypes.data, _C.byref(h)))
    return CellCenteredUniformMesh(_handle=h)


def create_slab_window_mesh(fullMesh, rank, nranks):
    """Shard `rank` of `nranks` of a full lattice cut into slabs along its slowest axis (pda_mesh_make_slab_window):
    sample cells = owned cells, stencil cells = [lower halo planes | owned planes | upper halo planes].  See
    pressiodemoapps.sharded for the halo exchange and the sharded evaluation helpers."""
    h = _vp()
    _check(_lib.pda_mesh_make_slab_window(fullMesh._h, int(rank), int(nranks), _C.byref(h)))
    m = CellCenteredUniformMesh(_handle=h)
    info = (_C.c_int64 * 8)()
    _check(_lib.pda_mesh_slab_window_info(m._h, info))
    m.window = dict(plane_cells=int(info[0]), k0=int(info[1]), k1=int(info[2]), halo_lo=int(info[3]), halo_hi=int(info[4]),
                    rank=int(info[5]), nranks=int(info[6]), dim=int(info[7]))
    return m


def create_sample_mesh(fullMesh, sampleMeshGids):
    """Native `meshing_scripts/create_sample_mesh.py` (in memory)."""
    g = _np.ascontiguousarray(_np.asarray(sampleMeshGids, dtype=_np.int32))
    h = _vp()
    _check(_lib.pda_mesh_make_sample(fullMesh._h, g.ctypes.data, g.size, _C.byref(h)))
    return CellCenteredUniformMesh(_handle=h)


# ------------------------------------------------------------------------------------------------ problems
class Problem:
    """adapter_py.hpp:57-199 (python names) + adapter_cpp.hpp:59-264 (C++ names)."""

    def __init__(self, mesh, family, problemId, recon, icFlag=1, params=None, device=0, _slab=None):
        self._mesh = mesh   # the mesh must outlive the problem (euler_2d_prob_class.hpp:1277)
        h = _vp()
        if _slab is not None:
            rank, nranks = _slab
            _check(_lib.pda_problem_create_slab(mesh._h, family, int(problemId), int(recon), rank, nranks, device,
                                                _C.byref(h)))
        else:
            params = params or {}
            names = (_cp * max(len(params), 1))(*[k.encode() for k in params])
            vals = (_dbl * max(len(params), 1))(*[float(v) for v in params.values()])
            _check(_lib.pda_problem_create(mesh._h, family, int(problemId), int(recon), int(icFlag), len(params),
                                           names, vals, device, _C.byref(h)))
        self._h = h
        self._recon = int(recon)
        self._pattern = None
        self._source = None      # host functor f(x[,y],t) of the ProblemA families
        self._source_t = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.pda_problem_free(h)
            self._h = None

    # ---- source term of the ProblemA families (the reference takes a Python functor: main_binder.cc:313-320,433-438)
    def setSourceTable(self, values):
        """per-SAMPLE-cell source values (pda_problem_set_source)"""
        v = _np.ascontiguousarray(values, dtype=_np.float64)
        if v.size != self._mesh.sampleMeshSize():
            raise ValueError("source table has %d entries, expected %d" % (v.size, self._mesh.sampleMeshSize()))
        _check(_lib.pda_problem_set_source(self._h, v.ctypes.data))

    def _set_source_functor(self, f):
        self._source = f
        x, y, _ = self._mesh._coords()
        c = self._mesh.graph()[:, 0]
        self._src_xy = (x[c], y[c])

    def _refresh_source(self, time):
        # the functor is evaluated on the host for every sample cell when the evaluation time changes -- exactly what
        # the reference's pybind wrapper does per cell (diffusion_reaction1d.hpp:176-181), then handed over as a table
        if self._source is None or self._source_t == time:
            return
        xs, ys = self._src_xy
        if self._mesh.dimensionality() == 1:
            tab = [float(self._source(float(a), float(time))) for a in xs]
        else:
            tab = [float(self._source(float(a), float(b), float(time))) for a, b in zip(xs, ys)]
        self.setSourceTable(_np.array(tab))
        self._source_t = time

    # ---- engine options (pda_problem_set_option): "jacobian_order" = "fast" | "reference"
    def setOption(self, name, value):
        _check(_lib.pda_problem_set_option(self._h, str(name).encode(), str(value).encode()))

    def getOption(self, name):
        buf = _C.create_string_buffer(64)
        _check(_lib.pda_problem_get_option(self._h, str(name).encode(), buf, 64))
        return buf.value.decode()

    # ---- sizes / parameters
    def numDofPerCell(self): return _lib.pda_problem_num_dof_per_cell(self._h)
    def totalDofSampleMesh(self): return _lib.pda_problem_total_dof_sample_mesh(self._h)
    def totalDofStencilMesh(self): return _lib.pda_problem_total_dof_stencil_mesh(self._h)

    def queryParameter(self, name):
        v = _dbl()
        _check(_lib.pda_problem_query_parameter(self._h, name.encode(), _C.byref(v)))
        return v.value

    def gamma(self): return self.queryParameter("gamma")
    def gravity(self): return self.queryParameter("gravity")
    def coriolis(self): return self.queryParameter("coriolis")

    def setBC(self, side, kind, values=None):
        v = None if values is None else _np.ascontiguousarray(values, dtype=_np.float64)
        _check(_lib.pda_problem_set_bc(self._h, int(side), int(kind), None if v is None else v.ctypes.data))

    def setBCFunctor(self, side, ghost, factors=None):
        """Arbitrary host functors for one side (pda_problem_set_bc_callback; the reference's C++ functor contract,
        custom_bcs_functions.hpp:107-164 / :60-103) -- the slow, fully general path:
            ghost(nearBdRowId, graphRow, cellX, cellY, U, ndpc, cellWidth, ghostValues)   # numpy views; fill ghostValues
            factors(graphRow, cellX, cellY, ndpc, factorsOut)                              # optional"""
        ncols = self._mesh.graph().shape[1]
        nU = self.totalDofStencilMesh()
        ndpc = self.numDofPerCell()
        h = self._recon + 1   # ghost layers = (scheme stencil - 1)/2, scheme stencil = 3 + 2*recon

        def _ghost(user, row_id, grow, x, y, U, nd, width, out):
            g = _np.ctypeslib.as_array(grow, shape=(ncols,))
            u = _np.ctypeslib.as_array(U, shape=(nU,))
            o = _np.ctypeslib.as_array(out, shape=(h * ndpc,))
            ghost(int(row_id), g, float(x), float(y), u, int(nd), float(width), o)

        cg = _GHOST_FN(_ghost)
        cf = _FACTOR_FN()
        if factors is not None:
            def _fac(user, grow, x, y, nd, out):
                g = _np.ctypeslib.as_array(grow, shape=(ncols,))
                o = _np.ctypeslib.as_array(out, shape=(ndpc,))
                factors(g, float(x), float(y), int(nd), o)
            cf = _FACTOR_FN(_fac)
        self._bc_keepalive = getattr(self, "_bc_keepalive", {})
        self._bc_keepalive[int(side)] = (cg, cf)
        _check(_lib.pda_problem_set_bc_callback(self._h, int(side), cg, cf, None))

    def setBCPointer(self, side, ptr):
        """setBCPointer(GhostRelativeLocation, ptr) of the reference's custom-BC problems: `ptr` (an int address or a
        ctypes pointer) becomes the `user` argument of that side's host functors (pda_problem_set_bc_pointer)."""
        _check(_lib.pda_problem_set_bc_pointer(self._h, int(side), ptr))

    # ---- creators
    def initialCondition(self):
        U = _np.zeros(self.totalDofStencilMesh())
        _check(_lib.pda_problem_initial_condition(self._h, U.ctypes.data))
        return U

    def createState(self): return _np.zeros(self.totalDofStencilMesh())
    def createRightHandSide(self): return _np.zeros(self.totalDofSampleMesh())
    createRhs = createRightHandSide
    createVelocity = createRightHandSide

    def jacobianNnz(self):
        nnz = _i64()
        _check(_lib.pda_problem_jacobian_nnz(self._h, _C.byref(nnz)))
        return int(nnz.value)

    def jacobianPattern(self):
        if self._pattern is None:
            nnz = _i64()
            _check(_lib.pda_problem_jacobian_nnz(self._h, _C.byref(nnz)))
            rowptr = _np.zeros(self.totalDofSampleMesh() + 1, dtype=_np.int32)
            colidx = _np.zeros(nnz.value, dtype=_np.int32)
            _check(_lib.pda_problem_jacobian_pattern(self._h, rowptr.ctypes.data, colidx.ctypes.data))
            self._pattern = (rowptr, colidx)
        return self._pattern

    def createJacobian(self):
        """scipy.sparse.csr_matrix with the fixed pattern (explicit zeros kept), like Eigen's RowMajor SparseMatrix."""
        import scipy.sparse as sp
        rowptr, colidx = self.jacobianPattern()
        return sp.csr_matrix((_np.zeros(colidx.size), colidx.copy(), rowptr.copy()),
                             shape=(self.totalDofSampleMesh(), self.totalDofStencilMesh()))

    def createApplyJacobianResult(self, operand):
        operand = _np.asarray(operand)
        if operand.ndim == 1:
            return _np.zeros(self.totalDofSampleMesh())
        order = "F" if operand.flags["F_CONTIGUOUS"] and not operand.flags["C_CONTIGUOUS"] else "C"
        return _np.zeros((self.totalDofSampleMesh(), operand.shape[1]), order=order)

    createResultOfJacobianActionOn = createApplyJacobianResult

    # ---- evaluation (host buffers)
    def rightHandSide(self, state, time, V):
        state = _f64_in(state, self.totalDofStencilMesh(), "state")
        _f64(V, self.totalDofSampleMesh(), "rhs")
        self._refresh_source(time)
        _check(_lib.pda_problem_velocity_host(self._h, state.ctypes.data, float(time), V.ctypes.data))

    rhs = rightHandSide
    velocity = rightHandSide

    def __call__(self, state, time, V, J=None, computeJacobian=False):
        if J is not None and computeJacobian:
            self.rightHandSideAndJacobian(state, time, V, J)
        else:
            self.rightHandSide(state, time, V)

    def rightHandSideAndJacobian(self, state, time, V, J):
        state = _f64_in(state, self.totalDofStencilMesh(), "state")
        vals = J.data if hasattr(J, "data") and not isinstance(J, _np.ndarray) else J
        _f64(vals, self.jacobianPattern()[1].size, "jacobian values")
        vp = None if V is None else _f64(V, self.totalDofSampleMesh(), "rhs").ctypes.data
        self._refresh_source(time)
        _check(_lib.pda_problem_velocity_and_jacobian_host(self._h, state.ctypes.data, float(time), vp,
                                                           vals.ctypes.data))

    def jacobian(self, state, time, J):
        self.rightHandSideAndJacobian(state, time, None, J)

    def applyJacobian(self, state, operand, time, result):
        state = _f64_in(state, self.totalDofStencilMesh(), "state")
        operand = _f64_in(operand, None, "operand")
        _f64(result, None, "result")
        ncols = 1 if operand.ndim == 1 else operand.shape[1]
        if operand.shape[0] != self.totalDofStencilMesh():
            raise ValueError("operand has %d rows, expected %d" % (operand.shape[0], self.totalDofStencilMesh()))
        want = (self.totalDofSampleMesh(),) if operand.ndim == 1 else (self.totalDofSampleMesh(), ncols)
        if tuple(result.shape) != want:
            raise ValueError("result has shape %s, expected %s" % (tuple(result.shape), want))
        if operand.ndim == 1 or (operand.flags["F_CONTIGUOUS"] and not operand.flags["C_CONTIGUOUS"]):
            layout = 0
            if operand.ndim == 2 and not result.flags["F_CONTIGUOUS"]:
                raise ValueError("col-major operand needs a col-major result")
        else:
            if not operand.flags["C_CONTIGUOUS"] or not result.flags["C_CONTIGUOUS"]:
                raise ValueError("operand/result must be contiguous")
            layout = 1
        self._refresh_source(time)
        _check(_lib.pda_problem_apply_jacobian_host(self._h, state.ctypes.data, operand.ctypes.data, ncols, layout,
                                                    float(time), result.ctypes.data))

    # ---- evaluation (device pointers: ints from tensor.data_ptr(); stream = cudaStream_t as int)
    def rightHandSideDevice(self, dU, time, dV, stream=0):
        self._refresh_source(time)
        _check(_lib.pda_problem_velocity_dev(self._h, dU, float(time), dV, stream))

    def rightHandSideAndJacobianDevice(self, dU, time, dV, dJvalues, stream=0):
        self._refresh_source(time)
        _check(_lib.pda_problem_velocity_and_jacobian_dev(self._h, dU, float(time), dV, dJvalues, stream))

    def applyJacobianDevice(self, dU, dB, ncols, layout, time, dR, stream=0):
        self._refresh_source(time)
        _check(_lib.pda_problem_apply_jacobian_dev(self._h, dU, dB, int(ncols), int(layout), float(time), dR, stream))

    # ---- test hooks / instrumentation
    def viewGhost(self, side):
        """viewGhostLeft/Front/Right/Back (euler_2d_prob_class.hpp:205-210): rows after the last evaluation."""
        nb = self._mesh.numCellsNearBd()
        out = _np.zeros((nb, max(1, self._ghost_stride())))
        _check(_lib.pda_problem_ghosts(self._h, int(side), out.ctypes.data))
        return out

    def _ghost_stride(self):
        return self.numDofPerCell() * ((self._stencil - 1) // 2)

    def launchCount(self):
        return int(_lib.pda_problem_launch_count(self._h))

    # ---- device-resident explicit time stepping (include/pda_b200.h: pda_problem_advance_*)
    def advance(self, stepper, state, dt, nsteps, startTime=0.0):
        """`nsteps` steps of `stepper` ("euler", "rk2", "rk4", "ssprk3") on the GPU; `state` (numpy, updated in place)
        crosses PCIe once each way"""
        _f64(state, self.totalDofStencilMesh(), "state")
        self._check_source_for_advance(startTime)
        _check(_lib.pda_problem_advance_host(self._h, _STEPPERS[stepper], state.ctypes.data, float(startTime), float(dt), int(nsteps)))

    def _check_source_for_advance(self, startTime):
        # the device-resident steppers evaluate every stage on the GPU from ONE source table: a host functor f(x[,y],t)
        # is tabulated at the start time; a time-DEPENDENT one cannot be honoured across stages and steps
        if self._source is None:
            return
        xs, ys = self._src_xy
        probe = [(float(xs[0]), float(ys[0])), (float(xs[-1]), float(ys[-1]))]
        one_d = self._mesh.dimensionality() == 1
        f = (lambda x, y, t: self._source(x, t)) if one_d else self._source
        if any(f(x, y, float(startTime)) != f(x, y, float(startTime) + 1.0) for x, y in probe):
            raise PdaError(5, "advance: the source functor depends on time; the device-resident steppers use one source "
                              "table for all stages -- step on the host with advanceRK2/RK4/SSP3 instead")
        self._refresh_source(startTime)

    def advanceDevice(self, stepper, dU, dt, nsteps, startTime=0.0, stream=0):
        self._check_source_for_advance(startTime)
        _check(_lib.pda_problem_advance_dev(self._h, _STEPPERS[stepper], dU, float(startTime), float(dt), int(nsteps), stream))

    # ---- slab decomposition (one process per GPU)
    def slabExtent(self):
        k0, k1, h, pd = _i32(), _i32(), _i32(), _i64()
        _check(_lib.pda_slab_extent(self._h, _C.byref(k0), _C.byref(k1), _C.byref(h), _C.byref(pd)))
        return k0.value, k1.value, h.value, pd.value

    def slabInitialCondition(self):
        k0, k1, _, pd = self.slabExtent()
        U = _np.zeros((k1 - k0) * pd)
        _check(_lib.pda_slab_initial_condition(self._h, U.ctypes.data))
        return U

    def slabVelocityInteriorDevice(self, dU, time, dV, stream=0):
        _check(_lib.pda_slab_velocity_interior_dev(self._h, dU, float(time), dV, stream))

    def slabVelocityBoundaryDevice(self, dU, time, dV, stream=0):
        _check(_lib.pda_slab_velocity_boundary_dev(self._h, dU, float(time), dV, stream))

    # peer mode: halo exchange over NVLink peer memory fused with the evaluation (include/pda_b200.h)
    def peerHandle(self):
        buf = _C.create_string_buffer(64)
        _check(_lib.pda_slab_peer_handle(self._h, buf))
        return buf.raw

    def peerConnect(self, handles):
        """handles: the 64-byte handles of ALL ranks, in rank order (list of bytes or one bytes object)"""
        blob = handles if isinstance(handles, (bytes, bytearray)) else b"".join(handles)
        if len(blob) % 64:
            raise ValueError("peerConnect: handles must be 64 bytes each")
        _check(_lib.pda_slab_peer_connect(self._h, _C.c_char_p(bytes(blob))))

    def peerConnectLocal(self, lower, upper):
        _check(_lib.pda_slab_peer_connect_local(self._h, lower._h, upper._h))

    def slabVelocityPeer(self, U_owned, time, V_owned):
        """host-pointer flavour: numpy (ideally pinned) arrays of the owned planes"""
        _check(_lib.pda_slab_velocity_peer_host(self._h, U_owned.ctypes.data, float(time), V_owned.ctypes.data))

    def slabVelocityPeerDevice(self, dU_owned, time, dV_owned, stream=0):
        _check(_lib.pda_slab_velocity_peer_dev(self._h, dU_owned, float(time), dV_owned, stream))


def _make(mesh, family, probEnum, recon, icFlag=1, params=None, device=0):
    p = Problem(mesh, family, probEnum, recon, icFlag, params, device)
    p._stencil = 3 + 2 * int(recon)
    return p


def create_problem(mesh, probEnum, *args, device=0):
    """Overload set of `create_problem` (src_py/main_binder.cc:255-546, C++ create_problem_eigen):
       (mesh, Euler1d|Euler3d|Advection1d|AdvectionDiffusionReaction2d, recon) ;
       (mesh, Euler2d|Swe2d, recon[, icFlag][, {name: value}]) ; (mesh, DiffusionReaction1d) ;
       (mesh, DiffusionReaction2d[, ViscousFluxReconstruction]) ;
       (mesh, AdvectionDiffusion2d, recon, ViscousFluxReconstruction[, {name: value}])."""
    fam = _FAMILY.get(type(probEnum))
    if fam is None:
        raise TypeError("create_problem: unknown problem enum %r" % (probEnum,))
    if fam in (5, 9):
        return _make(mesh, fam, probEnum, 0, 1, None, device)
    if not args:
        raise TypeError("create_problem: missing reconstruction enum")
    recon = args[0]
    icFlag, params = 1, None
    for a in args[1:]:
        if a is None or isinstance(a, (ViscousFluxReconstruction,)):
            continue
        if isinstance(a, dict):
            params = a
        else:
            icFlag = int(a)
    return _make(mesh, fam, probEnum, recon, icFlag, params, device)


def create_linear_advection_1d_problem(mesh, recon, *args, ic=1, device=0):
    """advection1d.hpp:107-152: (mesh, recon, InviscidFluxScheme, velocity) or (mesh, recon, velocity, ic=1)."""
    if args and isinstance(args[0], InviscidFluxScheme):
        velocity = float(args[1])
    else:
        velocity = float(args[0])
        if len(args) > 1:
            ic = int(args[1])
    return _make(mesh, 8, Advection1d.PeriodicLinear, recon, ic, {"velocity": velocity}, device)


def _problem_a(mesh, fam, enum, args, device):
    # (mesh, diffusion, reaction) or (mesh, sourceFunctor, diffusion, reaction)
    src = None
    if callable(args[0]):
        src, args = args[0], args[1:]
    p = _make(mesh, fam, enum, 0, 1, {"diffusion": float(args[0]), "reaction": float(args[1])}, device)
    if src is not None:
        p._set_source_functor(src)
    return p


def create_diffusion_reaction_1d_problem_A(mesh, *args, device=0):
    """diffusion_reaction1d.hpp:128-212: (mesh, diffusion, reaction) or (mesh, source(x, t), diffusion, reaction)."""
    return _problem_a(mesh, 9, DiffusionReaction1d.ProblemA, args, device)


def create_diffusion_reaction_2d_problem_A(mesh, viscRecon, *args, device=0):
    """diffusion_reaction2d.hpp:186-257: (mesh, visc, diffusion, reaction) or (mesh, visc, source(x, y, t), D, k)."""
    return _problem_a(mesh, 5, DiffusionReaction2d.ProblemA, args, device)


def create_burgers_2d_problem(mesh, probEnum, recon, viscRecon, params, device=0):
    """advection_diffusion2d.hpp:115-152 (custom coefficients through the {name: value} map)."""
    return _make(mesh, 6, probEnum, recon, 1, dict(params), device)


def create_adv_diff_reac_2d_problem_A(mesh, recon, ux, uy, diffusion, sigma, device=0):
    """advection_diffusion_reaction2d.hpp:140-160."""
    return _make(mesh, 7, AdvectionDiffusionReaction2d.ProblemA, recon, 1,
                 {"ux": ux, "uy": uy, "diffusion": diffusion, "sigma": sigma}, device)


def create_problem_slab(mesh, probEnum, recon, rank, nranks, device=0):
    """One z-slab (3D) / y-slab (2D) of a periodic full lattice per rank."""
    fam = _FAMILY[type(probEnum)]
    p = Problem(mesh, fam, probEnum, recon, device=device, _slab=(rank, nranks))
    p._stencil = 3 + 2 * int(recon)
    return p


def create_gray_scott_2d_problem(mesh, viscRecon, Du, Dv, F, k, device=0):
    """diffusion_reaction2d.hpp:259-285."""
    return _make(mesh, 5, DiffusionReaction2d.GrayScott, 0, 1, {"Du": Du, "Dv": Dv, "F": F, "k": k}, device)


def create_slip_wall_swe_2d_problem(mesh, recon, gravity, coriolis, pulseMagnitude, device=0):
    """swe2d.hpp:283-309 (legacy overload)."""
    return _make(mesh, 4, Swe2d.SlipWall, recon, 1,
                 {"gravity": gravity, "coriolis": coriolis, "pulseMagnitude": pulseMagnitude}, device)


def create_cross_shock_problem(mesh, recon, density, inletXVel, bottomYVel, device=0):
    """euler2d.hpp:252-283."""
    return _make(mesh, 2, Euler2d.CrossShock, recon, 1,
                 {"crossShockDensity": density, "crossShockInletXVel": inletXVel,
                  "crossShockBottomYVel": bottomYVel}, device)


# ------------------------------------------------------------------------------------- boundary-face gradients
class FacePosition(_IntEnum):
    """schemes_info.hpp:112-114."""
    Left = 0
    Front = 1
    Right = 2
    Back = 3
    Bottom = 4
    Top = 5


class _Face:
    """What queryFace returns (gradient.hpp:95-113): centerCoordinates, normalGradient (a float when the evaluator
    was made for one dof per cell, an array of MaxNumDofPerCell entries otherwise), normalDirection (1 = x, 2 = y)."""
    __slots__ = ("centerCoordinates", "normalGradient", "normalDirection", "parentCellGraphRow")

    def __init__(self, c, g, d, r):
        self.centerCoordinates, self.normalGradient, self.normalDirection, self.parentCellGraphRow = c, g, d, r


class GradientEvaluator:
    """GradientEvaluator<MeshType, MaxNumDofPerCell>(mesh) of the reference (gradient.hpp:61-121): normal gradients of
    a cell-centred field at the faces on the domain boundary (2D), computed on the GPU (pda_gradient_*)."""

    def __init__(self, mesh, maxNumDofPerCell=1):
        h = _vp()
        _check(_lib.pda_gradient_create(mesh._h, int(maxNumDofPerCell), _C.byref(h)))
        self._h = h
        self._max = int(maxNumDofPerCell)
        self._nStencil = mesh.stencilMeshSize()
        n = _lib.pda_gradient_num_faces(self._h)
        self.cellGIDs = _np.zeros(n, dtype=_np.int32)
        self.positions = _np.zeros(n, dtype=_np.int32)
        self.parentRows = _np.zeros(n, dtype=_np.int32)
        self.normalDirections = _np.zeros(n, dtype=_np.int32)
        self.centers = _np.zeros((n, 3))
        _check(_lib.pda_gradient_faces(self._h, self.cellGIDs.ctypes.data, self.positions.ctypes.data,
                                       self.parentRows.ctypes.data, self.normalDirections.ctypes.data,
                                       self.centers.ctypes.data))
        self.normalGradients = _np.zeros((n, self._max))   # like the reference's faces: zero until evaluated

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.pda_gradient_free(h)
            self._h = None

    def numFaces(self):
        return int(self.cellGIDs.size)

    def __call__(self, field, numDofPerCell=1):
        """operator()(field) / operator()(field, numDofPerCell), gradient.hpp:77-93."""
        nd = int(numDofPerCell)
        if nd > self._max or nd < 1:   # let the library produce the reference's message
            _check(_lib.pda_gradient_compute_host(self._h, None, nd, None))
        if not isinstance(field, _np.ndarray) or field.dtype != _np.float64:
            raise TypeError("field must be a float64 numpy array")
        if field.size != self._nStencil * nd:
            raise ValueError("field has %d entries, expected %d" % (field.size, self._nStencil * nd))
        fld = _np.ascontiguousarray(field)   # kept alive in a local for the duration of the C call
        out = _np.zeros((self.numFaces(), nd))
        _check(_lib.pda_gradient_compute_host(self._h, fld.ctypes.data, nd, out.ctypes.data))
        self.normalGradients[:, :nd] = out
        return out

    def computeDevice(self, dField, numDofPerCell, dNormalGrad, stream=0):
        """device pointers (ints), asynchronous on `stream`; dNormalGrad = [numFaces][numDofPerCell]"""
        _check(_lib.pda_gradient_compute_dev(self._h, dField, int(numDofPerCell), dNormalGrad, stream))

    def queryFace(self, cellGID, facePosition):
        """gradient.hpp:115-117."""
        i = _i32()
        _check(_lib.pda_gradient_query_face(self._h, int(cellGID), int(facePosition), _C.byref(i)))
        k = i.value
        g = float(self.normalGradients[k, 0]) if self._max == 1 else self.normalGradients[k].copy()
        return _Face(self.centers[k].copy(), g, int(self.normalDirections[k]), int(self.parentRows[k]))

    def launchCount(self):
        return int(_lib.pda_gradient_launch_count(self._h))


# ------------------------------------------------------------------------------------------------ time loops
# pressiodemoapps/__init__.py:90-185 of the reference (host-side RK loops around rightHandSide)
def advanceRK2(appObj, state, dt, Nsteps, startTime=0.0, observer=None, showProgress=False):
    v = appObj.createRightHandSide()
    time = startTime
    for step in range(1, Nsteps + 1):
        appObj.rightHandSide(state, time, v)
        if observer is not None:
            observer(step - 1, state, v)
        k1 = dt * v
        tmp = state + k1
        appObj.rightHandSide(tmp, time + dt, v)
        k2 = dt * v
        state[:] = state + k2 * 0.5 + k1 * 0.5
        time += dt


def advanceRK4(appObj, state, dt, Nsteps, startTime=0.0, observer=None, showProgress=False):
    v = appObj.createRightHandSide()
    time = startTime
    for step in range(1, Nsteps + 1):
        appObj.rightHandSide(state, time, v)
        if observer is not None:
            observer(step - 1, state, v)
        k1 = dt * v
        tmp = state + 0.5 * k1
        appObj.rightHandSide(tmp, time + 0.5 * dt, v)
        k2 = dt * v
        tmp = state + 0.5 * k2
        appObj.rightHandSide(tmp, time + 0.5 * dt, v)
        k3 = dt * v
        tmp = state + k3
        appObj.rightHandSide(tmp, time + dt, v)
        k4 = dt * v
        state[:] = state + (k1 + 2. * k2 + 2. * k3 + k4) * (1. / 6.)
        time += dt


def advanceSSP3(appObj, state, dt, Nsteps, startTime=0.0, observer=None, showProgress=False):
    v = appObj.createRightHandSide()
    t1 = state.copy()
    t2 = state.copy()
    time = startTime
    for step in range(1, Nsteps + 1):
        appObj.rightHandSide(state, time, v)
        if observer is not None:
            observer(step - 1, state, v)
        t1[:] = state + dt * v
        appObj.rightHandSide(t1, time + dt, v)
        t2[:] = (3. / 4.) * state + (1. / 4.) * t1 + (1. / 4.) * dt * v
        appObj.rightHandSide(t2, time + dt * 0.5, v)
        state[:] = (1. / 3.) * (state + 2. * t2 + 2. * dt * v)
        time += dt
