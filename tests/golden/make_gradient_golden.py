"""Generates tests/golden/gradients/*.npz from the UNMODIFIED reference's GradientEvaluator (run in the build
container only; needs /root/reference and `make -C oracle`).

Meshes: the six mesh directories the reference's own test uses (tests_cpp/gradients/{full,sample}mesh_s{3,5,7}:
12x15 cells on [0,1]^2 and a 25-cell sample of it) plus two full meshes written by the reference's own
create_full_mesh.py (one periodic in x, so only the y walls carry faces).  Fields: the reference test's
sin(pi x y) (tests_cpp/gradients/main.cc:7-9) through the one-dof API, and a seeded random 3-dof field through the
multi-dof API.  Stored per case: the mesh arrays as the reference's loader sees them, graphRowsOfCellsStrictlyOnBd(),
the face records (cell gid, FacePosition, parent row, normal direction, centre), both gradients, and the RMSE pair the
reference's test compares with its gold table (main.cc:246-253, |diff| <= 1e-6) -- checked here at generation time.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refdrv import RefProblem, ref_gradient, ref_rows_strictly_on_bd  # noqa: E402

REF = "/root/reference"
GOLD_RMSE = {   # tests_cpp/gradients/main.cc:246-253
    "fullmesh_s3": (0.118044, 0.082173), "samplemesh_s3": (0.134392, 0.0759626),
    "fullmesh_s5": (0.0512737, 0.0298508), "samplemesh_s5": (0.0682185, 0.0307574),
    "fullmesh_s7": (0.0512737, 0.0298508), "samplemesh_s7": (0.0682185, 0.0307574),
}


def mesh_arrays(meshDir):
    # any 2D problem class loads the mesh the same way; DiffusionReaction2d ProblemA accepts every stencil size >= 3
    with open(os.path.join(meshDir, "info.dat")) as f:
        info = dict(line.split() for line in f if line.strip())
    stencil = int(info["stencilSize"])
    p = RefProblem(meshDir, "euler2d", 0, {3: 0, 5: 1, 7: 2}[stencil], 1)
    a = p.mesh_arrays()
    a["stencil"] = stencil
    return a


def one_case(name, meshDir, out):
    a = mesh_arrays(meshDir)
    x, y = a["x"], a["y"]
    f1 = np.sin(np.pi * x * y)
    rng = np.random.default_rng(20261018)
    f3 = np.repeat(f1[:, None], 3, axis=1) * (1.0 + 0.25 * rng.uniform(-1, 1, (x.size, 3)))
    r1 = ref_gradient(meshDir, f1, 1, scalar_api=True)
    r1b = ref_gradient(meshDir, f1, 1, scalar_api=False)
    assert np.array_equal(r1["grad"], r1b["grad"])
    r3 = ref_gradient(meshDir, f3.ravel(), 3)
    rows = ref_rows_strictly_on_bd(meshDir)
    # the reference test's RMSE (main.cc:48-100): x faces against y pi cos(pi x y), y faces against x pi cos(pi x y),
    # both sums divided by the total number of faces
    cx, cy = r1["centers"][:, 0], r1["centers"][:, 1]
    isx = r1["normalDir"] == 1
    gold = np.where(isx, cy, cx) * np.pi * np.cos(np.pi * cx * cy)
    err2 = (r1["grad"][:, 0] - gold) ** 2
    n = max(1, r1["grad"].shape[0])
    rmse = (np.sqrt(err2[isx].sum() / n), np.sqrt(err2[~isx].sum() / n))
    if name in GOLD_RMSE:
        assert abs(rmse[0] - GOLD_RMSE[name][0]) <= 1e-6 and abs(rmse[1] - GOLD_RMSE[name][1]) <= 1e-6, (name, rmse)
    np.savez_compressed(os.path.join(out, name + ".npz"), graph=a["graph"], x=a["x"], y=a["y"], z=a["z"], d=a["d"],
                        stencil=a["stencil"], rowsNearBd=a["rowsNearBd"], rowsStrictlyOnBd=rows, f1=f1, f3=f3,
                        cellGid=r1["cellGid"], position=r1["position"], parentRow=r1["parentRow"],
                        normalDir=r1["normalDir"], centers=r1["centers"], grad1=r1["grad"], grad3=r3["grad"],
                        rmse=np.array(rmse))
    print("%-22s faces %4d  rows on bd %4d  rmse %.6f %.6f" % (name, r1["grad"].shape[0], rows.size, *rmse))


def main():
    out = os.path.join(HERE, "gradients")
    os.makedirs(out, exist_ok=True)
    for name in GOLD_RMSE:
        one_case(name, os.path.join(REF, "tests_cpp", "gradients", name), out)
    with tempfile.TemporaryDirectory() as tmp:
        for name, n, bounds, per, s in (("full_20x16_s5", [20, 16], [-1, 1, 0, 2], (), 5),
                                         ("full_perx_14x9_s7", [14, 9], [0, 1, 0, 1], ("x",), 7)):
            d = os.path.join(tmp, name)
            cmd = [sys.executable, os.path.join(REF, "meshing_scripts", "create_full_mesh.py"), "-n"] + \
                  [str(v) for v in n] + ["--outDir", d, "-s", str(s), "--bounds"] + [str(v) for v in bounds]
            if per:
                cmd += ["--periodic"] + list(per)
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
            one_case(name, d, out)


if __name__ == "__main__":
    main()
