"""Parity table, round 2 (committed as profiles/parity_report_r02.txt): the CUDA engine through the C-ABI against
  (1) every golden fixture of the UNMODIFIED reference (tests/golden/*.npz): velocity and Jacobian in the default
      ("fast") mode and in the reference-order mode, strict tolerance and number of entries differing in any bit;
  (2) the oracle run live at BASELINE.json's actual sizes (tests/parity_cases.py; the same bodies the -m gpu tests
      assert on).
'strict' = max |a-b| / (1e-10 + 1e-12 |b|) per entry (<= 1 <=> north-star tolerance); 'field' = the relative part
measured against max |b| (reported beside it, never substituted silently).   Usage: python tools/parity_report_r02.py [--quick]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pressio-demoapps_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pressiodemoapps as pda  # noqa: E402
from conftest import Golden, golden_names  # noqa: E402
from parity_cases import (R, SEED, err_stats, full_lattice_velocity, full_rows_vs_sample_oracle, make_problem,  # noqa: E402
                          perturb_inplace, sample_mesh_case)
from test_host_cpu import make_mesh, make_problem as make_golden_problem  # noqa: E402


def fmt(s):
    return "strict %9.3g field %8.3g bits %d/%d%s" % (s["strict"], s["field"], s["bits"], s["n"],
                                                      (" nan %d (mismatch %d)" % (s["nans"], s["nan_mismatch"])) if s["nans"] or s["nan_mismatch"] else "")


def fixtures():
    print("== (1) golden fixtures of the unmodified reference: fast mode | reference-order mode ==", flush=True)
    worst = dict(Vf=0, Jf=0, Vr=0, Jr=0, bitsV=0, bitsJ=0)
    for name in golden_names():
        g = Golden(name)
        m = g.meta
        mesh, _ = make_mesh(g)
        out = {}
        for order in ("fast", "reference"):
            p = make_golden_problem(g, mesh)
            p.setOption("order", order)
            U, t = g["U"], m["t"]
            V = p.createRightHandSide()
            p.rightHandSide(U, t, V)
            J = p.createJacobian()
            V2 = p.createRightHandSide()
            p.rightHandSideAndJacobian(U, t, V2, J)
            out[order] = (err_stats(V, g["V"]), err_stats(V2, g["V2"]), err_stats(J.data, g["Jv"]))
        f, r = out["fast"], out["reference"]
        worst["Vf"] = max(worst["Vf"], f[0]["strict"], f[1]["strict"]); worst["Jf"] = max(worst["Jf"], f[2]["strict"])
        worst["Vr"] = max(worst["Vr"], r[0]["strict"], r[1]["strict"]); worst["Jr"] = max(worst["Jr"], r[2]["strict"])
        worst["bitsV"] += r[0]["bits"] + r[1]["bits"]; worst["bitsJ"] += r[2]["bits"]
        print("%-34s fast: V %.3g V2 %.3g J %.3g | reference-order: V %.3g (bits %d) V2 %.3g (bits %d) J %.3g (bits %d/%d)" % (
            name, f[0]["strict"], f[1]["strict"], f[2]["strict"], r[0]["strict"], r[0]["bits"], r[1]["strict"], r[1]["bits"],
            r[2]["strict"], r[2]["bits"], r[2]["n"]), flush=True)
    print("WORST over %d fixtures: fast V %.3g J %.3g | reference-order V %.3g J %.3g, entries differing in any bit: V %d, J %d" % (
        len(golden_names()), worst["Vf"], worst["Jf"], worst["Vr"], worst["Jr"], worst["bitsV"], worst["bitsJ"]), flush=True)


def fullsize(quick):
    print("== (2) BASELINE sizes vs the oracle run live ==", flush=True)
    P3 = ("x", "y", "z")
    n3 = [128] * 3 if quick else [512] * 3
    for recon, sten, tag in ((R.Weno5, 7, "weno5 (extension)"), (R.Weno3, 5, "weno3 (reference-pinned)")):
        s = full_lattice_velocity("euler3d", pda.Euler3d.PeriodicSmooth, recon, n3, [-1, 1] * 3, sten, P3)
        print("cfg5 euler3d smooth %d^3 %-24s V: %s | gpu(host entry) %.1fs oracle %.1fs x%d thr" % (
            n3[0], tag, fmt(s), s["gpu_s"], s["oracle_s"], s["oracle_threads"]), flush=True)
    n2 = [512, 512] if quick else [2048, 2048]
    s = full_lattice_velocity("euler2d", pda.Euler2d.Riemann, R.Weno5, n2, [0, 1, 0, 1], 7, ())
    print("cfg2 euler2d riemann %d^2 weno5                V: %s" % (n2[0], fmt(s)), flush=True)
    for order in ("fast", "reference"):
        if order == "reference" and not quick:
            n2r = [1024, 1024]   # the reference-order mode needs the stored graph + read-modify-write: a quarter-size mesh
        else:
            n2r = n2
        s = full_rows_vs_sample_oracle("euler2d", pda.Euler2d.Riemann, R.Weno5, n2r, [0, 1, 0, 1], 7, 0.01, order=order)
        print("cfg2 euler2d riemann %d^2 weno5 %-9s rows %d (near-bd %d) nnz checked %d of %d | V: %s | J: %s" % (
            n2r[0], order, s["rows"], s["near_bd_rows"], s["nnz_checked"], s["nnz_full"], fmt(s["V"]), fmt(s["J"])), flush=True)
    n4 = [1024, 1024] if quick else [4096, 4096]
    for fam, prob, recon, sten, b, per, tag in (
            ("swe2d", pda.Swe2d.SlipWall, R.FirstOrder, 3, [-5, 5, -5, 5], (), "swe slipwall first order"),
            ("swe2d", pda.Swe2d.SlipWall, R.Weno3, 5, [-5, 5, -5, 5], (), "swe slipwall weno3"),
            ("diffreac2d", pda.DiffusionReaction2d.GrayScott, 0, 3, [-1.25, 1.25, -1.25, 1.25], ("x", "y"), "gray-scott")):
        s = full_lattice_velocity(fam, prob, recon, n4, b, sten, per)
        print("cfg3 %-26s %d^2      V: %s" % (tag, n4[0], fmt(s)), flush=True)
    nd = [1024, 256] if quick else [4096, 1024]
    rng = np.random.default_rng(SEED)
    gids = np.sort(rng.choice(nd[0] * nd[1], int(0.05 * nd[0] * nd[1]), replace=False)).astype(np.int32)
    for recon, sten, tag in ((R.Weno3, 5, "weno3"), (R.Weno5, 7, "weno5")):
        full = pda.create_full_mesh(nd, [0, 4, 0, 1], sten)
        Uf = perturb_inplace(make_problem(full, "euler2d", pda.Euler2d.DoubleMachReflection, recon).initialCondition())
        for t in (0.0, 0.1):
            for order in ("fast", "reference"):
                s = sample_mesh_case("euler2d", pda.Euler2d.DoubleMachReflection, recon, nd, [0, 4, 0, 1], sten, gids, t, order, Uf)
                print("cfg4 dmr %dx%d 5%% sample %-5s t=%.1f %-9s V: %s | V(jac): %s | J: %s" % (
                    nd[0], nd[1], tag, t, order, fmt(s["V"]), fmt(s["V2"]), fmt(s["J"])), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    print("tools/parity_report_r02.py on", "B200" if pda.device_count() else "NO DEVICE", "(quick sizes)" if quick else "", flush=True)
    if "--no-fixtures" not in sys.argv:
        fixtures()
    if "--no-fullsize" not in sys.argv:
        fullsize(quick)
