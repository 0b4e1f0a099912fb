"""Evidence for the WENO-Jacobian tolerance: how far is the REFERENCE from itself / from the exact Jacobian?

Compares, on identical perturbed states,
  ref      : the reference compiled without FMA contraction (oracle/_ref/libpda_ref.so, the parity oracle)
  ref_fma  : the SAME reference sources compiled with -O3 -march=native -ffp-contract=fast
  exact    : the C restatement evaluated in 80-bit long double (oracle/_ref/libpda_oracle_ld.so)
and, if a GPU is present, the CUDA engine.  Prints max |dJ| and the north-star score max |dJ|/(1e-10+1e-12|J|).
"""
import ctypes as C, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pressiodemoapps as pda
import refdrv
from refdrv import RefProblem

R = pda.InviscidFluxReconstruction
LD = np.longdouble

def exact_jacobian(meshDir, fam, prob, rec, ic, U, t):
    L = C.CDLL(os.path.join(refdrv.REFDIR, "libpda_oracle_ld.so"))
    L.or_create.restype = C.c_void_p
    L.or_create.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_void_p] * 2
    L.or_query.restype = C.c_longlong; L.or_query.argtypes = [C.c_void_p, C.c_int]
    h = L.or_create(meshDir.encode(), refdrv.FAM[fam], int(prob), int(rec), ic, 0, None, None)
    nnz = L.or_query(h, 11); nv = L.or_query(h, 10)
    Ul = U.astype(LD); V = np.zeros(nv, dtype=LD); J = np.zeros(nnz, dtype=LD)
    L.or_velocity_and_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_longdouble, C.c_void_p, C.c_void_p]
    L.or_velocity_and_jacobian(h, Ul.ctypes.data, C.c_longdouble(t), V.ctypes.data, J.ctypes.data)
    return V, J

def score(a, b):
    a = np.asarray(a, dtype=LD); b = np.asarray(b, dtype=LD)
    d = np.abs(a - b)
    return float(np.nanmax(d)), float(np.nanmax(d / (1e-10 + 1e-12 * np.abs(b))))

def case(name, n, b, per, s, fam, prob, rec, ic=1, t=0.0, gpu=False):
    d = tempfile.mkdtemp(); m = pda.create_full_mesh(n, b, s, per); m.write(d)
    ref = RefProblem(d, fam, int(prob), rec, ic)
    U = ref.initialCondition()
    U = U * (1 + 1e-3 * np.random.default_rng(20261017).uniform(-1, 1, U.size))
    _, J = ref.velocityAndJacobian(U, t)
    refdrv._libs.pop(("ref", False), None)
    saved = refdrv.ref_lib_path
    refdrv.ref_lib_path = lambda omp=False: os.path.join(refdrv.REFDIR, "libpda_ref_fma.so")
    try:
        rf = RefProblem(d, fam, int(prob), rec, ic); _, Jf = rf.velocityAndJacobian(U, t)
    finally:
        refdrv.ref_lib_path = saved; refdrv._libs.pop(("ref", False), None)
    _, Jx = exact_jacobian(d, fam, prob, rec, ic, U, t)
    line = "%-28s max|J| %.1e | ref vs ref_fma: %.2e (%.2f) | ref vs exact: %.2e (%.2f)" % (
        (name, np.abs(J).max()) + score(Jf, J) + score(J, Jx))
    if gpu:
        p = pda.create_problem(m, prob, rec, ic)
        Jg = p.createJacobian(); V = p.createRightHandSide(); p.rightHandSideAndJacobian(U, t, V, Jg)
        line += " | gpu vs exact: %.2e (%.2f) | gpu vs ref: %.2e (%.2f)" % (score(Jg.data, Jx) + score(Jg.data, J))
    print(line, flush=True)

if __name__ == "__main__":
    gpu = pda.device_count() > 0
    case("euler1d sod weno5", [100,1],[-0.5,0.5],(),7,'euler1d',pda.Euler1d.Sod,R.Weno5, gpu=gpu)
    case("euler1d sod weno3", [100,1],[-0.5,0.5],(),5,'euler1d',pda.Euler1d.Sod,R.Weno3, gpu=gpu)
    case("euler2d riemann weno5", [20,20],[0,1,0,1],(),7,'euler2d',pda.Euler2d.Riemann,R.Weno5, gpu=gpu)
    case("euler2d smooth weno5 32^2", [32,32],[-1,1,-1,1],('x','y'),7,'euler2d',pda.Euler2d.PeriodicSmooth,R.Weno5, gpu=gpu)
    case("euler2d dmr weno3", [60,15],[0,4,0,1],(),5,'euler2d',pda.Euler2d.DoubleMachReflection,R.Weno3, gpu=gpu)
    case("swe weno5", [25,25],[-5,5,-5,5],(),7,'swe2d',pda.Swe2d.SlipWall,R.Weno5, gpu=gpu)
    case("euler3d smooth weno3 12^3", [12,12,12],[-1,1,-1,1,-1,1],('x','y','z'),5,'euler3d',pda.Euler3d.PeriodicSmooth,R.Weno3, gpu=gpu)
    case("euler2d riemann firstorder", [20,20],[0,1,0,1],(),3,'euler2d',pda.Euler2d.Riemann,R.FirstOrder, gpu=gpu)
