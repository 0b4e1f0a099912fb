"""Parity table: CUDA engine (through the C-ABI) vs the compiled reference (oracle/_ref/libpda_ref.so).
Diagnostic companion of tests/test_parity_gpu.py; prints max abs / max scaled differences per case."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pressiodemoapps as pda
from refdrv import RefProblem

R = pda.InviscidFluxReconstruction

def scaled(a, b, rtol=1e-12, atol=1e-10):
    """max of |a-b| / (atol + rtol*|b|): <= 1 means np.allclose(a, b, rtol, atol)"""
    return float(np.max(np.abs(a - b) / (atol + rtol * np.abs(b)))) if a.size else 0.0

def case(name, n, b, per, s, famname, prob, rec, ic=1, params=None, t=0.0, perturb=True, jac=True):
    d = tempfile.mkdtemp()
    m = pda.create_full_mesh(n, b, s, per); m.write(d)
    ref = RefProblem(d, famname, int(prob), rec, ic, params)
    p = pda.create_problem(m, prob, rec, ic, params) if famname != 'diffreac2d' else pda.create_problem(m, prob)
    U = ref.initialCondition()
    if perturb:
        rng = np.random.default_rng(20261017)
        U = U * (1 + 1e-3 * rng.uniform(-1, 1, U.size))
    V = p.createRightHandSide(); p.rightHandSide(U, t, V)
    Vr = ref.velocity(U, t)
    line = "%-34s V: abs %.2e scaled %.3f" % (name, np.abs(V - Vr).max(), scaled(V, Vr))
    if jac:
        J = p.createJacobian(); V2 = p.createRightHandSide()
        p.rightHandSideAndJacobian(U, t, V2, J)
        Vr2, Jr = ref.velocityAndJacobian(U, t)
        line += " | V(jac path): scaled %.3f | J: abs %.2e scaled %.3f (max|J| %.1e)" % (
            scaled(V2, Vr2), np.abs(J.data - Jr).max(), scaled(J.data, Jr), np.abs(Jr).max())
    print(line, flush=True)

if __name__ == "__main__":
    case("euler1d sod weno5 100", [100,1],[-0.5,0.5],(),7,'euler1d',pda.Euler1d.Sod,R.Weno5)
    case("euler1d sod weno3 100", [100,1],[-0.5,0.5],(),5,'euler1d',pda.Euler1d.Sod,R.Weno3)
    case("euler1d sod fo 100", [100,1],[-0.5,0.5],(),3,'euler1d',pda.Euler1d.Sod,R.FirstOrder)
    case("euler1d smooth weno5 per", [64,1],[-1,1],('x',),7,'euler1d',pda.Euler1d.PeriodicSmooth,R.Weno5)
    case("euler2d riemann weno5 20^2", [20,20],[0,1,0,1],(),7,'euler2d',pda.Euler2d.Riemann,R.Weno5)
    case("euler2d riemann ic2 weno3", [20,20],[0,1,0,1],(),5,'euler2d',pda.Euler2d.Riemann,R.Weno3,2)
    case("euler2d riemann fo (s7 mesh)", [20,20],[0,1,0,1],(),7,'euler2d',pda.Euler2d.Riemann,R.FirstOrder)
    case("euler2d smooth weno5 per 32^2", [32,32],[-1,1,-1,1],('x','y'),7,'euler2d',pda.Euler2d.PeriodicSmooth,R.Weno5)
    case("euler2d smooth weno5 per 5^2", [5,5],[-1,1,-1,1],('x','y'),7,'euler2d',pda.Euler2d.PeriodicSmooth,R.Weno5)
    case("euler2d dmr weno3 t=0", [60,15],[0,4,0,1],(),5,'euler2d',pda.Euler2d.DoubleMachReflection,R.Weno3)
    case("euler2d dmr weno5 t=0.1", [60,15],[0,4,0,1],(),7,'euler2d',pda.Euler2d.DoubleMachReflection,R.Weno5,t=0.1)
    case("euler2d sedovsym weno3", [20,20],[0,1,0,1],(),5,'euler2d',pda.Euler2d.SedovSymmetry,R.Weno3)
    case("euler2d normalshock weno5", [24,12],[0,2,0,1],(),7,'euler2d',pda.Euler2d.NormalShock,R.Weno5)
    case("euler2d crossshock weno3", [24,12],[0,2,0,1],(),5,'euler2d',pda.Euler2d.CrossShock,R.Weno3)
    case("euler2d KH weno5 per", [24,24],[-5,5,-5,5],('x','y'),7,'euler2d',pda.Euler2d.KelvinHelmholtz,R.Weno5)
    case("swe slipwall fo 25^2", [25,25],[-5,5,-5,5],(),3,'swe2d',pda.Swe2d.SlipWall,R.FirstOrder)
    case("swe slipwall weno3 25^2", [25,25],[-5,5,-5,5],(),5,'swe2d',pda.Swe2d.SlipWall,R.Weno3)
    case("swe slipwall weno5 ic2", [25,25],[-5,5,-5,5],(),7,'swe2d',pda.Swe2d.SlipWall,R.Weno5,2)
    case("grayscott 32^2", [32,32],[-1.25,1.25,-1.25,1.25],('x','y'),3,'diffreac2d',pda.DiffusionReaction2d.GrayScott,0)
    case("euler3d smooth fo per 8^3", [8,8,8],[-1,1,-1,1,-1,1],('x','y','z'),3,'euler3d',pda.Euler3d.PeriodicSmooth,R.FirstOrder)
    case("euler3d smooth weno3 per 12^3", [12,12,12],[-1,1,-1,1,-1,1],('x','y','z'),5,'euler3d',pda.Euler3d.PeriodicSmooth,R.Weno3)
    case("euler3d sedovsym weno3 10^3", [10,10,10],[0,1,0,1,0,1],(),5,'euler3d',pda.Euler3d.SedovSymmetry,R.Weno3)
    case("euler3d sedovsym fo 10^3", [10,10,10],[0,1,0,1,0,1],(),3,'euler3d',pda.Euler3d.SedovSymmetry,R.FirstOrder)
