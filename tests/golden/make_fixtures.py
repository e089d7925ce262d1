"""Regenerates tests/golden/tominec_fitted.npz from the reference's own test fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
Sources (plain CSV, read verbatim, stored as float64/int64 arrays):
    test/data/x_nodes_fitted.csv, y_nodes_fitted.csv, Y_idx_in.csv, Y_idx_dirichlet.csv, Y_idx_neumann.csv,
    x_normals.csv, y_normals.csv          (consumed by test/poisson_test.jl:19-31, test/hyperviscosity_test.jl:13-14)
The thresholds those tests assert are recorded alongside (poisson_test.jl:132, hyperviscosity_test.jl:32-33).
Two CGNS meshes are copied byte for byte (mesh data, not source): test/data/tominec_Y.cgns (test/mesh_import_test.jl:27)
and examples/rect_0_10.cgns (BASELINE config 1, examples/adv_diff_test.jl:26 family); rb.mesh.Hdf5File reads them.
"""
import shutil
import os

import numpy as np

REF = "/root/reference/test/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tominec_fitted.npz")


def main():
    ld = lambda f, dt=np.float64: np.loadtxt(os.path.join(REF, f), delimiter=",", dtype=dt)
    np.savez_compressed(
        OUT,
        X=ld("x_nodes_fitted.csv"), Y=ld("y_nodes_fitted.csv"),
        Y_idx_in=ld("Y_idx_in.csv").astype(np.int64), Y_idx_dirichlet=ld("Y_idx_dirichlet.csv").astype(np.int64),
        Y_idx_neumann=ld("Y_idx_neumann.csv").astype(np.int64),
        x_normals=ld("x_normals.csv"), y_normals=ld("y_normals.csv"),
        poisson_threshold=np.float64(0.0027), hyperviscosity_rtol=np.float64(np.sqrt(np.finfo(float).eps)),
    )
    here = os.path.dirname(os.path.abspath(__file__))
    shutil.copyfile(os.path.join(REF, "tominec_Y.cgns"), os.path.join(here, "tominec_Y.cgns"))
    shutil.copyfile("/root/reference/examples/rect_0_10.cgns", os.path.join(here, "rect_0_10.cgns"))
    z = np.load(OUT)
    print({k: z[k].shape for k in z.files})


if __name__ == "__main__":
    main()
