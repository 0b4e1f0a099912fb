"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference (run in the build container only).

For every case:
  * the mesh is produced by the reference's OWN Python scripts (/root/reference/meshing_scripts/create_full_mesh.py,
    create_sample_mesh.py) -- not by this repo's native mesh tools, so the fixtures pin those tools too;
  * the problem is created and evaluated by the reference's own C++ (oracle/_ref/libpda_ref.so = oracle/ref_driver.cc
    compiled against /root/reference/include, serial build);
  * the state is the reference initial condition perturbed as SURVEY 8(d): U = IC*(1 + 1e-3*xi), xi ~ U(-1,1),
    numpy.random.default_rng(20261017).
Stored per case (npz): mesh arrays as the reference's loader sees them (graph, x, y, z, d, dInv, rowsInner,
rowsNearBd), sha256 of the three mesh text files, IC, U, t, V (velocity-only path), V2 + CSR (rowptr, colidx, values)
from the velocity+Jacobian path, ghost rows where the reference exposes them, and (sample meshes) the sample gids.

Usage:  python tests/golden/make_golden.py            (needs /root/reference and `make -C oracle`)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refdrv import RefProblem  # noqa: E402

REF = "/root/reference"
FO, W3, W5 = 0, 1, 2

# name, n, bounds, periodic, stencil, family, probEnum, recon, icFlag, params, t, sampleFraction
CASES = [
    # cfg 1: 1D Euler Sod (tests_cpp/eigen_1d_euler_sod_explicit uses 100 cells on [-0.5,0.5])
    ("euler1d_sod_weno5_100", [100, 1], [-0.5, 0.5], (), 7, "euler1d", 1, W5, 1, None, 0.0, None),
    ("euler1d_sod_weno3_100", [100, 1], [-0.5, 0.5], (), 5, "euler1d", 1, W3, 1, None, 0.0, None),
    ("euler1d_sod_fo_100", [100, 1], [-0.5, 0.5], (), 3, "euler1d", 1, FO, 1, None, 0.0, None),
    ("euler1d_sod_weno5_1000", [1000, 1], [-0.5, 0.5], (), 7, "euler1d", 1, W5, 1, None, 0.0, None),
    ("euler1d_lax_weno3_60", [60, 1], [-5.0, 5.0], (), 5, "euler1d", 2, W3, 1, None, 0.0, None),
    ("euler1d_shuosher_weno5_80", [80, 1], [-5.0, 5.0], (), 7, "euler1d", 3, W5, 1, None, 0.0, None),
    ("euler1d_smooth_weno5_per64", [64, 1], [-1.0, 1.0], ("x",), 7, "euler1d", 0, W5, 1, None, 0.0, None),
    # cfg 2: 2D Euler Riemann (tests_cpp/eigen_2d_euler_riemann_explicit: 20x20 on [0,1]^2)
    ("euler2d_riemann_weno5_20", [20, 20], [0, 1, 0, 1], (), 7, "euler2d", 4, W5, 1, None, 0.0, None),
    ("euler2d_riemann_ic2_weno3_20", [20, 20], [0, 1, 0, 1], (), 5, "euler2d", 4, W3, 2, None, 0.0, None),
    ("euler2d_riemann_fo_s7mesh_20", [20, 20], [0, 1, 0, 1], (), 7, "euler2d", 4, FO, 1, None, 0.0, None),
    ("euler2d_riemann_weno5_param", [16, 14], [0, 1, 0, 1], (), 7, "euler2d", 4, W5, 1,
     {"riemannTopRightPressure": 0.7, "gamma": 1.5}, 0.0, None),
    ("euler2d_smooth_weno5_per20", [20, 20], [-1, 1, -1, 1], ("x", "y"), 7, "euler2d", 0, W5, 1, None, 0.0, None),
    ("euler2d_smooth_weno3_per_24x20", [24, 20], [-1, 1, -1, 1], ("x", "y"), 5, "euler2d", 0, W3, 1, None, 0.0, None),
    ("euler2d_smooth_weno5_per5", [5, 5], [-1, 1, -1, 1], ("x", "y"), 7, "euler2d", 0, W5, 1, None, 0.0, None),
    ("euler2d_kh_weno5_per16", [16, 16], [-5, 5, -5, 5], ("x", "y"), 7, "euler2d", 1, W5, 1, None, 0.0, None),
    ("euler2d_sedovfull_weno3_20", [20, 20], [-1.2, 1.2, -1.2, 1.2], (), 5, "euler2d", 2, W3, 1, None, 0.0, None),
    ("euler2d_sedovsym_weno3_20", [20, 20], [0, 1.2, 0, 1.2], (), 5, "euler2d", 3, W3, 1, None, 0.0, None),
    ("euler2d_sedovsym_weno5_18", [18, 18], [0, 1.2, 0, 1.2], (), 7, "euler2d", 3, W5, 1, None, 0.0, None),
    ("euler2d_normalshock_weno5_24x12", [24, 12], [0, 2, 0, 1], (), 7, "euler2d", 5, W5, 1, None, 0.0, None),
    ("euler2d_crossshock_weno3_24x12", [24, 12], [0, 2, 0, 1], (), 5, "euler2d", 7, W3, 1, None, 0.0, None),
    # cfg 4: double Mach reflection (tests_cpp/...double_mach_reflection_explicit: 60x15 on [0,4]x[0,1]); the top
    # wall is time dependent -> t = 0 and t = 0.1; plus a ~5% sample mesh (no reference fixture exists: SURVEY 8c)
    ("euler2d_dmr_fo_60x15", [60, 15], [0, 4, 0, 1], (), 3, "euler2d", 6, FO, 1, None, 0.0, None),
    ("euler2d_dmr_weno3_60x15", [60, 15], [0, 4, 0, 1], (), 5, "euler2d", 6, W3, 1, None, 0.0, None),
    ("euler2d_dmr_weno5_60x15_t002", [60, 15], [0, 4, 0, 1], (), 7, "euler2d", 6, W5, 1, None, 0.02, None),
    ("euler2d_dmr_weno3_sample", [80, 20], [0, 4, 0, 1], (), 5, "euler2d", 6, W3, 1, None, 0.1, 0.05),
    ("euler2d_dmr_weno5_sample", [80, 20], [0, 4, 0, 1], (), 7, "euler2d", 6, W5, 1, None, 0.0, 0.05),
    ("euler2d_riemann_weno5_sample", [30, 30], [0, 1, 0, 1], (), 7, "euler2d", 4, W5, 1, None, 0.0, 0.1),
    # cfg 3: SWE slip wall (tests_cpp/eigen_2d_swe_slip_wall_explicit: 25x25 on [-5,5]^2) and Gray-Scott
    ("swe_slipwall_fo_25", [25, 25], [-5, 5, -5, 5], (), 3, "swe2d", 0, FO, 1, None, 0.0, None),
    ("swe_slipwall_weno3_25", [25, 25], [-5, 5, -5, 5], (), 5, "swe2d", 0, W3, 1, None, 0.0, None),
    ("swe_slipwall_weno5_ic2_25", [25, 25], [-5, 5, -5, 5], (), 7, "swe2d", 0, W5, 2, None, 0.0, None),
    ("swe_slipwall_weno3_param", [20, 22], [-5, 5, -5, 5], (), 5, "swe2d", 0, W3, 1,
     {"gravity": 7.5, "coriolis": -1.5, "pulseMagnitude": 0.2}, 0.0, None),
    ("swe_slipwall_weno3_sample", [30, 30], [-5, 5, -5, 5], (), 5, "swe2d", 0, W3, 1, None, 0.0, 0.1),
    ("grayscott_per32", [32, 32], [-1.25, 1.25, -1.25, 1.25], ("x", "y"), 3, "diffreac2d", 1, FO, 1, None, 0.0, None),
    ("grayscott_per_20x24_param", [20, 24], [-1.25, 1.25, -1.25, 1.25], ("x", "y"), 3, "diffreac2d", 1, FO, 1,
     {"Du": 3e-4, "Dv": 6e-5, "F": 0.03, "k": 0.05}, 0.0, None),
    ("grayscott_sample", [30, 30], [-1.25, 1.25, -1.25, 1.25], ("x", "y"), 3, "diffreac2d", 1, FO, 1, None, 0.0, 0.1),
    # cfg 5: 3D Euler (the reference has first order and WENO3 only: SURVEY F1)
    ("euler3d_smooth_fo_per8", [8, 8, 8], [-1, 1, -1, 1, -1, 1], ("x", "y", "z"), 3, "euler3d", 0, FO, 1, None, 0.0, None),
    ("euler3d_smooth_weno3_per8", [8, 8, 8], [-1, 1, -1, 1, -1, 1], ("x", "y", "z"), 5, "euler3d", 0, W3, 1, None, 0.0, None),
    ("euler3d_smooth_weno3_per_10x8x6", [10, 8, 6], [-1, 1, -1, 1, -1, 1], ("x", "y", "z"), 5, "euler3d", 0, W3, 1, None, 0.0, None),
    ("euler3d_sedovsym_weno3_8", [8, 8, 8], [0, 1, 0, 1, 0, 1], (), 5, "euler3d", 1, W3, 1, None, 0.0, None),
    ("euler3d_sedovsym_fo_7x6x8", [7, 6, 8], [0, 1, 0, 1, 0, 1], (), 3, "euler3d", 1, FO, 1, None, 0.0, None),
    # the remaining families the north star names: Burgers (tests_cpp/eigen_2d_burgers_*), advection-diffusion-reaction
    # (eigen_2d_adv_diff_reac_*), 1D linear advection (eigen_1d_linear_adv_*), diffusion-reaction ProblemA
    # (eigen_{1,2}d_diffusion_reaction_*); "testSource" = a time-dependent analytic source functor (oracle/ref_driver.cc)
    ("burgers_per_weno5_20x18", [20, 18], [-1, 1, -1, 1], ("x", "y"), 7, "advdiff2d", 0, W5, 1, None, 0.0, None),
    ("burgers_per_fo_12x10", [12, 10], [-1, 1, -1, 1], ("x", "y"), 3, "advdiff2d", 0, FO, 1, None, 0.0, None),
    ("burgers_out_weno3_param", [20, 18], [-1, 1, -1, 1], (), 5, "advdiff2d", 1, W3, 1,
     {"diffusion": 1e-3, "pulseX": 0.1, "pulseMagnitude": 0.6}, 0.0, None),
    ("burgers_out_weno5_20x18", [20, 18], [-1, 1, -1, 1], (), 7, "advdiff2d", 1, W5, 1, None, 0.0, None),
    ("burgers_out_fo_16", [16, 16], [-1, 1, -1, 1], (), 3, "advdiff2d", 1, FO, 1, None, 0.0, None),
    ("burgers_out_weno3_sample", [30, 30], [-1, 1, -1, 1], (), 5, "advdiff2d", 1, W3, 1, None, 0.0, 0.1),
    ("adr_fo_16x14", [16, 14], [0, 1, 0, 1], (), 3, "advdiffreac2d", 0, FO, 1, None, 0.0, None),
    ("adr_weno3_16x14", [16, 14], [0, 1, 0, 1], (), 5, "advdiffreac2d", 0, W3, 1, None, 0.0, None),
    ("adr_weno5_param", [18, 16], [0, 1, 0, 1], (), 7, "advdiffreac2d", 0, W5, 1,
     {"ux": 0.3, "uy": 0.2, "diffusion": 0.01, "sigma": 2.0}, 0.0, None),
    ("adr_weno5_sample", [30, 30], [0, 1, 0, 1], (), 7, "advdiffreac2d", 0, W5, 1, None, 0.0, 0.1),
    ("adv1d_weno5_per50", [50, 1], [-1.0, 1.0], ("x",), 7, "advection1d", 0, W5, 1, None, 0.0, None),
    ("adv1d_weno3_ic2_vel", [50, 1], [0.0, 4.0], ("x",), 5, "advection1d", 0, W3, 2, {"velocity": 0.5}, 0.0, None),
    ("adv1d_fo_ic4", [40, 1], [0.0, 4.0], ("x",), 3, "advection1d", 0, FO, 4, None, 0.0, None),
    ("diffreac1d_a_40", [40, 1], [0.0, 1.0], (), 3, "diffreac1d", 0, FO, 1, None, 0.0, None),
    ("diffreac1d_a_src_param", [40, 1], [0.0, 1.0], (), 3, "diffreac1d", 0, FO, 1,
     {"diffusion": 0.02, "reaction": 0.03, "testSource": 1}, 0.3, None),
    ("diffreac2d_a_16x14", [16, 14], [0, 1, 0, 1], (), 3, "diffreac2d", 0, FO, 1, None, 0.0, None),
    ("diffreac2d_a_src", [16, 14], [0, 1, 0, 1], (), 3, "diffreac2d", 0, FO, 1, {"testSource": 1}, 0.3, None),
    ("diffreac2d_a_sample", [30, 30], [0, 1, 0, 1], (), 3, "diffreac2d", 0, FO, 1, None, 0.0, 0.1),
]


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def run(cmd):
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)


def make_case(case, outdir):
    name, n, bounds, periodic, stencil, fam, prob, recon, ic, params, t, frac = case
    tmp = tempfile.mkdtemp(prefix="gold_")
    full = os.path.join(tmp, "full")
    cmd = [sys.executable, os.path.join(REF, "meshing_scripts", "create_full_mesh.py"), "-n"] + [str(v) for v in n] + \
          ["--outDir", full, "-s", str(stencil), "--bounds"] + [repr(float(b)) for b in bounds]
    if periodic:
        cmd += ["--periodic"] + list(periodic)
    run(cmd)
    meshdir = full
    extra = {}
    if frac:
        ncell = int(np.prod(n))
        rng = np.random.default_rng(20261017)
        gids = np.sort(rng.choice(ncell, max(4, int(frac * ncell)), replace=False)).astype(np.int64)
        gfile = os.path.join(tmp, "sample_mesh_gids.dat")
        np.savetxt(gfile, gids, fmt="%8d")
        meshdir = os.path.join(tmp, "sample")
        run([sys.executable, os.path.join(REF, "meshing_scripts", "create_sample_mesh.py"), "--fullMeshDir", full,
             "--sampleMeshIndices", gfile, "--outDir", meshdir])
        extra["sampleGids"] = gids.astype(np.int32)
        extra["stencilGids"] = np.loadtxt(os.path.join(meshdir, "stencil_mesh_gids.dat"), dtype=np.int64).astype(np.int32)
    ref = RefProblem(meshdir, fam, prob, recon, ic, params)
    ma = ref.mesh_arrays()
    IC = ref.initialCondition()
    rng = np.random.default_rng(20261017)
    U = IC * (1.0 + 1e-3 * rng.uniform(-1, 1, IC.size))
    if not np.any(IC):   # families whose initial condition is identically zero: a small random state instead
        U = 0.1 * rng.uniform(-1, 1, IC.size)
    if params and "testSource" in params:   # the functor's values at the sample cells (what the binding tabulates)
        c = ma["graph"][:, 0]
        extra["src"] = np.sin(ma["x"][c] + t) if fam == "diffreac1d" else np.cos(ma["x"][c] * ma["y"][c] + t)
    V = ref.velocity(U, t)
    V2, Jv = ref.velocityAndJacobian(U, t)
    rowptr, colidx = ref.pattern()
    for s in range(4):
        g = ref.ghosts(s)
        if g is not None and ref.nNearBd:
            extra["ghost%d" % s] = g
    meta = dict(name=name, n=n, bounds=[float(b) for b in bounds], periodic=list(periodic), stencil=stencil,
                family=fam, prob=prob, recon=recon, ic=ic, params=params or {}, t=t, sample=bool(frac),
                sha={f: sha(os.path.join(meshdir, f)) for f in ("info.dat", "connectivity.dat", "coordinates.dat")},
                dim=ref.dim, ndpc=ref.ndpc, nnz=ref.nnz)
    np.savez_compressed(os.path.join(outdir, name + ".npz"), meta=json.dumps(meta), graph=ma["graph"], x=ma["x"],
                        y=ma["y"], z=ma["z"], d=ma["d"], dInv=ma["dInv"], rowsInner=ma["rowsInner"],
                        rowsNearBd=ma["rowsNearBd"], IC=IC, U=U, V=V, V2=V2, Jv=Jv, rowptr=rowptr, colidx=colidx,
                        **extra)
    shutil.rmtree(tmp)
    print("%-36s cells %6d nnz %8d  max|V| %.3e" % (name, ref.nSample, ref.nnz, np.abs(V).max()), flush=True)


def main():
    only = set(sys.argv[1:])
    for c in CASES:
        if only and c[0] not in only:
            continue
        make_case(c, HERE)


if __name__ == "__main__":
    main()
