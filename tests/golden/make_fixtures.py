"""Regenerates the committed golden fixtures.  Run in the build container (needs /root/reference):

    python tests/golden/make_fixtures.py

* tet10_fixture.npz -- the unstructured Tet10 mesh of the reference's own test fixture
  test/test_problems_contact_3d/tet10.inp (607 nodes, 269 C3D10), parsed with juliafem.jl_b200.mesh.read_abaqus_inp
  and stored as flat arrays (coords, 1-based conn, element sets) so that GPU-box tests need no /root/reference.
* linear_static_smp18.npz -- the mesh of the reference's end-to-end example examples/linear_static.jl
  (examples/linear_static/JuliaFEMSMP18.med: 14 078 nodes, 6 131 Tet10 after the edge/surface cells are dropped), read with
  juliafem.jl_b200.mesh.read_med (pure-Python HDF5 subset, h5lite.py) exactly as aster_read_mesh does, incl. the
  Code Aster -> Abaqus node reordering.  The expected answer of that example (max |u| = 2.4052929896922337,
  examples/linear_static.jl:133) is in pins.json.
* pins.json -- the known-answer values that survive in the reference's docs/comments; each
  entry cites its source.  They are literals, not computed from our code.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"


def main():
    from juliafem.jl_b200 import mesh
    m = mesh.read_abaqus_inp(os.path.join(REF, "test/test_problems_contact_3d/tet10.inp"))
    assert m.elem_type == 10 and m.n_nodes == 607 and m.n_elems == 269, (m.n_nodes, m.n_elems)
    np.savez_compressed(os.path.join(HERE, "tet10_fixture.npz"), coords=m.coords, conn=m.conn,
                        **{"elset_" + k: v for k, v in m.elem_sets.items()})
    ms = mesh.read_med(os.path.join(REF, "examples/linear_static/JuliaFEMSMP18.med"))
    assert ms.elem_type == 10 and ms.n_nodes == 14078 and ms.n_elems == 6131, (ms.n_nodes, ms.n_elems)
    np.savez_compressed(os.path.join(HERE, "linear_static_smp18.npz"), coords=ms.coords, conn=ms.conn)
    pins = {
        "linear_static": {"source": "examples/linear_static.jl:23-100,133 on examples/linear_static/JuliaFEMSMP18.med",
                          "E": 208.0e3, "nu": 0.30, "displacement_load_1": 1.0,
                          "fixed": "all three components of the nodes with |y - 50| <= 6 and of the 3 nodes nearest to (165, 88, 10)",
                          "max_u_norm": 2.4052929896922337, "isapprox_rtol": 1.4901161193847656e-08,
                          "n_nodes": 14078, "n_tet10": 6131},
        "le_uniaxial": {"source": "docs/book/linear_elastic_implementation.md:187-200", "E": 200e9, "nu": 0.3,
                        "eps": [1e-3, 0, 0, 0, 0, 0], "sigma11_MPa": 269.2307692, "sigma22_MPa": 115.3846154},
        "le_pure_shear": {"source": "docs/book/linear_elastic_implementation.md:205-214", "E": 200e9, "nu": 0.3,
                          "gamma12": 0.002, "sigma12_MPa": 153.85},
        "le_tangent_identity": {"source": "docs/book/linear_elastic_implementation.md:233-241",
                                "eps": [1e-3, 5e-4, 3e-4, -2e-4, 4e-4, 6e-4]},
        "pp_uniaxial": {"source": "src/materials/perfect_plasticity.jl:297-337 evaluated as in SURVEY.md 8c "
                                  "(the doc's printed 714.08 MPa at docs/book/perfect_plasticity_implementation.md:305-312 is stale)",
                        "E": 200e9, "nu": 0.3, "sigma_y": 250e6, "H": 1e9, "eps11": 3e-3,
                        "sigma11_MPa": 667.275141, "eps_p11": 9.1271e-4, "alpha11_MPa": 0.608474, "kappa": 1.36907e-3},
        "pp_on_surface": {"source": "docs/book/perfect_plasticity_implementation.md:296-301", "sigma_y_MPa": 250.0},
        "tet10_mass_times_2520": {"source": "src/assembly/assembly.jl:139-149", "table": [
            [6, 1, 1, 1, -4, -6, -4, -4, -6, -6], [1, 6, 1, 1, -4, -4, -6, -6, -4, -6], [1, 1, 6, 1, -6, -4, -4, -6, -6, -4],
            [1, 1, 1, 6, -6, -6, -6, -4, -4, -4], [-4, -4, -6, -6, 32, 16, 16, 16, 16, 8], [-6, -4, -4, -6, 16, 32, 16, 8, 16, 16],
            [-4, -6, -4, -6, 16, 16, 32, 16, 8, 16], [-4, -6, -6, -4, 16, 8, 16, 32, 16, 16], [-6, -4, -6, -4, 16, 16, 8, 16, 32, 16],
            [-6, -6, -4, -4, 8, 16, 16, 16, 16, 32]]},
        "quadrature": {"source": "src/quadrature/gltet.jl:7-25; src/quadrature/quaddata.jl:4-5; "
                                 "test/tutorials/01_fundamentals/basis_functions.jl:44-196",
                       "gltet4_weight": 1.0 / 24.0, "gltet1_weight": 1.0 / 6.0, "glhex8_point": 0.5773502691896258,
                       "tet_volume": 1.0 / 6.0, "hex_volume": 8.0},
        "fixture_mesh": {"source": "test/test_problems_contact_3d/tet10.inp", "n_nodes": 607, "n_elems": 269},
    }
    with open(os.path.join(HERE, "pins.json"), "w") as fh:
        json.dump(pins, fh, indent=1)
    print("wrote fixtures")


if __name__ == "__main__":
    main()
