"""CPU tests of the results writer: the HDF5 writer of h5lite.py round-trips through the reader that is
pinned on the reference's own `.med` file, and the XDMF document has the structure `update_xdmf!` writes (src/io.jl:387-518)."""
import os
import types
from xml.etree import ElementTree as ET

import numpy as np
import pytest


def test_h5_writer_roundtrip(tmp_path, jf):
    from juliafem.jl_b200 import h5lite
    rng = np.random.default_rng(0)
    data = {"DataItem_1": rng.random((7, 3)), "DataItem_2": np.arange(24, dtype=np.int64).reshape(2, 3, 4),
            "DataItem_10": np.arange(5, dtype=np.int32), "single": np.ones(3, dtype=np.float32), "empty": np.zeros((0, 3)),
            "bytes": np.arange(9, dtype=np.uint8)}
    data.update({f"many_{i:03d}": np.full(3, float(i)) for i in range(70)})        # more entries than a default symbol-table leaf
    fn = str(tmp_path / "t.h5")
    h5lite.write(fn, data)
    back = h5lite.read(fn)
    assert sorted(back) == sorted(data)
    for k, v in data.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v)
    raw = open(fn, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                         # classic superblock
    assert int.from_bytes(raw[40:48], "little") == len(raw)                        # end-of-file address
    with pytest.raises(ValueError):
        h5lite.write(fn, {"a/b": np.zeros(2)})
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.write(fn, {"c": np.zeros(2, dtype=np.complex128)})


@pytest.mark.parametrize("fmt", ["HDF", "XML"])
def test_xdmf_document_structure_and_readback(tmp_path, jf, fmt):
    from juliafem.jl_b200 import xdmf as X
    m = jf.mesh.tet10_kuhn(2, 1, 1)
    rng = np.random.default_rng(1)
    u = rng.standard_normal(m.n_dofs)
    stress = rng.standard_normal((m.n_nodes, 6))
    name = str(tmp_path / "model_results")
    w = X.Xdmf(name, overwrite=True, format=fmt)
    assert X.update_xdmf_(w, "body", 0.0, m.coords, m.conn, 10, u=u, stress=stress)
    with pytest.warns(UserWarning):
        assert not X.update_xdmf_(w, "body", 0.0, m.coords, m.conn, 10, u=u)       # same grid twice: skipped, as the reference
    assert X.update_xdmf_(w, "body", 1.0, m.coords, m.conn, 10, u=2 * u)
    w.close()
    with pytest.raises(FileExistsError):
        X.Xdmf(name)                                                               # src/io.jl:41-45
    root = ET.parse(name + ".xmf").getroot()
    assert root.tag == "Xdmf" and root.get("Version") == "3.0"
    tc = root.find("Domain").find("Grid")
    assert tc.get("CollectionType") == "Temporal" and tc.get("GridType") == "Collection"
    scs = tc.findall("Grid")
    assert [sc.get("CollectionType") for sc in scs] == ["Spatial", "Spatial"]
    assert [float(sc.find("Time").get("Value")) for sc in scs] == [0.0, 1.0]
    frame = scs[0].find("Grid")
    assert frame.get("Name") == "body" and frame.find("Geometry").get("Type") == "XYZ"
    assert frame.find("Topology").get("TopologyType") == "Mixed"
    assert {(a.get("Name"), a.get("AttributeType"), a.get("Center")) for a in frame.findall("Attribute")} == \
        {("Displacement", "Vector", "Node"), ("Stress", "Tensor6", "Node")}
    assert frame.find("Geometry").find("DataItem").get("Dimensions") == f"{m.n_nodes} 3"
    if fmt == "HDF":
        assert frame.find("Geometry").find("DataItem").text == "model_results.h5:/DataItem_1"
        assert os.path.isfile(name + ".h5")
    frames = X.read_xdmf(name + ".xmf")
    assert len(frames) == 2 and frames[1]["time"] == 1.0
    f0 = frames[0]
    assert np.array_equal(f0["coords"], m.coords)
    assert len(f0["elements"]) == m.n_elems and all(code == 38 for code, _ in f0["elements"])     # Tet10 = 38 (src/io.jl:346)
    assert np.array_equal(np.array([n for _, n in f0["elements"]]), m.conn - 1)                  # 0-based, reference node order
    assert np.array_equal(f0["fields"]["Displacement"], u.reshape(-1, 3)) and np.array_equal(f0["fields"]["Stress"], stress)
    assert np.array_equal(frames[1]["fields"]["Displacement"], 2 * u.reshape(-1, 3))
    assert np.array_equal(w.read("/Domain/Grid/Grid[2]/Grid/Attribute[@Name='Displacement']"), 2 * u.reshape(-1, 3))


def test_write_results_hook_of_the_analysis_mirror(tmp_path, jf):
    """add_results_writer_ / write_results_ (src/analysis.jl:73-105) with a solved-analysis stand-in: node data goes out in
    the dense node order of the device data (sorted original ids, ext:100-108), post-processed fields included."""
    from juliafem.jl_b200 import api as A, xdmf as X
    m = jf.mesh.hex8_lattice(3, 2, 2, 0.5)
    ids = np.arange(m.n_nodes) * 3 + 7                                            # arbitrary original node ids
    data = types.SimpleNamespace(coords=m.coords, conn=m.conn, topology=A.Hex8, node_ids=ids)
    model = A.Problem(A.Elasticity, "block", 3)
    model._data = data
    model.postprocess_fields.append("stress")
    rng = np.random.default_rng(2)
    model.fields["stress"] = {int(n): rng.standard_normal(6) for n in ids}
    an = A.Analysis(A.Linear, model)
    an.u = rng.standard_normal(m.n_dofs)
    w = X.Xdmf(str(tmp_path / "res"), overwrite=True)
    A.add_results_writer_(an, w)
    A.write_results_(an, 0.5)
    fr = X.read_xdmf(w.xmffile)
    assert len(fr) == 1 and fr[0]["name"] == "block" and fr[0]["time"] == 0.5
    assert all(code == 9 for code, _ in fr[0]["elements"])                        # Hex8 = 9 (src/io.jl:339)
    assert np.array_equal(fr[0]["fields"]["Displacement"], an.u.reshape(-1, 3))
    assert np.array_equal(fr[0]["fields"]["Stress"], np.array([model.fields["stress"][int(n)] for n in ids]))
