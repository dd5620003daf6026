#!/usr/bin/env python
"""examples/linear_static.jl of the reference, line for line, on the B200 path (BASELINE.json configs[0]).

    python examples/linear_static.py [path/to/JuliaFEMSMP18.med]

The mesh is read from the `.med` file the reference ships (examples/linear_static/JuliaFEMSMP18.med, HDF5 parsed by
juliafem.jl_b200/h5lite.py) or, without an argument, from the committed copy tests/golden/linear_static_smp18.npz.
Needs a B200 (the library has no CPU fallback).  Expected output: max |u| = 2.4052929896922337 (examples/linear_static.jl:133)
and the result files model_results.xmf / model_results.h5 (:92-107, :123-124).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from juliafem.jl_b200 import api as A, mesh as M          # noqa: E402
from juliafem.jl_b200.xdmf import Xdmf                      # noqa: E402


def main():
    # mesh = aster_read_mesh(joinpath(datadir, "JuliaFEMSMP18.med"))                     (:23-24)
    if len(sys.argv) > 1:
        mesh = M.read_med(sys.argv[1])
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "linear_static_smp18.npz"))
        mesh = M.Mesh(10, z["coords"], z["conn"])
    # model = Problem(Elasticity, "OTHER", 3); model_elements = create_elements(mesh, "OTHER")   (:26-27; Seg3 / Tri6 filtered, :35-39)
    model = A.Problem(A.Elasticity, "OTHER", 3)
    model_elements = A.create_elements(mesh, "OTHER")
    A.update_(model_elements, "youngs modulus", 208.0e3)                                  # :28
    A.update_(model_elements, "poissons ratio", 0.30)                                     # :29
    A.update_(model_elements, "density", 7.80e-9)                                         # :30
    A.add_elements_(model, model_elements)                                                # :31
    # nodes with |y - 50| <= 6 plus the three nodes nearest to the dot of the "i"           (:46-72)
    A.add_node_to_node_set_(mesh, "mid_fixed", *M.nodes_at_plane(mesh, 1, 50.0))        # :58
    A.add_node_to_node_set_(mesh, "mid_fixed", *M.find_nearest_nodes(mesh, [165.0, 88.0, 10], 3))   # :64-67
    mid_fixed = mesh.node_sets["mid_fixed"]
    fixed = A.Problem(A.Dirichlet, "fixed", 3, "displacement")                            # :76
    fixed_elements = A.create_nodal_elements(mesh, "mid_fixed")                           # :77
    A.add_elements_(fixed, fixed_elements)                                                # :78
    for c in (1, 2, 3):                                                                   # :79-81
        A.update_(fixed_elements, f"displacement {c}", 0.0)
    A.update_(model_elements, "displacement load 1", 1.0)                                 # :85
    analysis = A.Analysis(A.Linear, model, fixed)                                         # :88
    xdmf = Xdmf("model_results", overwrite=True)                                          # :92
    A.add_results_writer_(analysis, xdmf)                                                 # :93
    model.postprocess_fields.append("stress")                                             # :96
    A.run_(analysis, tol=1e-12, relative=True, max_iter=200000)                           # :100 (projected CG instead of LDLt)
    xdmf.close()                                                                          # :107
    assert os.path.isfile("model_results.xmf") and os.path.isfile("model_results.h5")     # :123-124
    u = analysis("displacement", 0.0)                                                     # :130-131
    u_norms = {i: float(np.linalg.norm(v)) for i, v in u.items()}
    umax = max(u_norms.values())
    print(f"nodes {mesh.n_nodes}, Tet10 elements {mesh.n_elems}, fixed nodes {mid_fixed.size}, CG iterations {analysis.cg_iterations}, "
          f"converged {analysis.converged}")
    print(f"max |u| = {umax!r}   (reference: 2.4052929896922337, examples/linear_static.jl:133)")
    assert abs(umax / 2.4052929896922337 - 1.0) < 1.5e-8                                  # isapprox default rtol
    fr = analysis("reaction force", 0.0)
    print("sum of reaction forces:", np.sum(list(fr.values()), axis=0), " (balances the body load)")


if __name__ == "__main__":
    main()
