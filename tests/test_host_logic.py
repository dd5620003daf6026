"""CPU: mesh generators, Abaqus reader, partitioner, C-ABI symbol export.  No GPU, no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_abi_exports_every_declared_symbol(jf):
    from juliafem.jl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "jfem_b200.h")).read()
    declared = set(re.findall(r"\b(jfem_[a-z_0-9]+)\s*\(", hdr))
    declared.discard("jfem_handle")
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/jfem_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert _lib.lib().jfem_abi_version() == 2


def test_no_gpu_means_loud_failure(jf):
    from juliafem.jl_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    m = jf.mesh.tet10_kuhn(1, 1, 1)
    with pytest.raises(_lib.JfemError) as ei:
        _lib.Handle(10, m.coords, m.conn)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_unsupported_element_is_refused(jf):
    """Mirror of test/test_problems_elasticity_assemble_3d_seg3.jl:7-12 (assemble! throws for Seg3 in 3D)."""
    from juliafem.jl_b200 import _lib
    with pytest.raises(_lib.JfemError) as ei:
        _lib.Handle(3, np.zeros((3, 3)), np.array([[1, 2, 3]], dtype=np.int32))
    assert ei.value.code == 1 and "unsupported element type" in str(ei.value)


def test_hex8_lattice_matches_reference_numbering(jf):
    m = jf.mesh.hex8_lattice(4, 3, 3)
    assert m.n_nodes == 36 and m.n_elems == 12
    nx, ny = 4, 3
    # benchmarks/multigpu_mpi_benchmark.jl:86-94 for the first element and one interior element
    assert m.conn[0].tolist() == [1, 2, 2 + nx, 1 + nx, 1 + nx * ny, 2 + nx * ny, 2 + nx + nx * ny, 1 + nx + nx * ny]
    k, j, i = 1, 1, 2   # 0-based cell indices
    n1 = k * nx * ny + j * nx + i + 1
    e = k * (ny - 1) * (nx - 1) + j * (nx - 1) + i
    assert m.conn[e, 0] == n1 and m.conn[e, 6] == n1 + 1 + nx + nx * ny
    assert np.allclose(m.coords[n1 - 1], [i, j, k])


def test_tet10_kuhn_sizes(jf):
    # T1 = 88x22x22 cells -> 358 425 nodes, 255 552 elements (checked on a scaled-down box + formula)
    m = jf.mesh.tet10_kuhn(4, 2, 2)
    assert m.n_nodes == 9 * 5 * 5 and m.n_elems == 6 * 16
    assert (2 * 88 + 1) * (2 * 22 + 1) ** 2 == 358425 and 6 * 88 * 22 * 22 == 255552
    X = m.coords[m.conn - 1]
    det = np.linalg.det(X[:, 1:4] - X[:, :1])
    assert np.all(det > 0) and abs(det.sum() / 6 - 1.0 * 0.5 * 0.5) < 1e-12
    for k, (a, b) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):
        assert np.allclose(X[:, 4 + k], 0.5 * (X[:, a] + X[:, b]))


def test_abaqus_reader_roundtrip(tmp_path, jf):
    m = jf.mesh.tet10_kuhn(1, 1, 1)
    p = tmp_path / "m.inp"
    with open(p, "w") as fh:
        fh.write("*HEADING\n** comment\n*NODE, NSET=ALL\n")
        for i, c in enumerate(m.coords):
            fh.write(f"{10 * (i + 1)}, {c[0]}, {c[1]}, {c[2]}\n")
        fh.write("*ELEMENT, TYPE=C3D10, ELSET=BODY\n")
        for e, c in enumerate(m.conn):
            ids = [str(10 * int(n)) for n in c]
            fh.write(f"{e + 1}, " + ", ".join(ids[:7]) + ",\n" + ", ".join(ids[7:]) + "\n")
        fh.write("*NSET, NSET=LEFT, GENERATE\n10, 30, 10\n*ELSET, ELSET=FIRST\n1, 2\n")
    r = jf.mesh.read_abaqus_inp(str(p))
    assert r.elem_type == 10 and r.n_nodes == m.n_nodes and np.array_equal(r.conn, m.conn)
    assert np.allclose(r.coords, m.coords)
    assert r.node_sets["LEFT"].tolist() == [1, 2, 3] and r.elem_sets["FIRST"].tolist() == [1, 2]
    assert r.elem_sets["BODY"].size == 6


def test_fixture_mesh_loads(jf):
    d = np.load(os.path.join(HERE, "golden", "tet10_fixture.npz"))
    assert d["coords"].shape == (607, 3) and d["conn"].shape == (269, 10)
    assert d["conn"].min() == 1 and d["conn"].max() == 607


@pytest.mark.parametrize("P", [2, 3, 4])
def test_partition_covers_mesh_and_halo_is_consistent(jf, P):
    m = jf.mesh.hex8_lattice(5, 4, 7)
    parts = [jf.mesh.partition_mesh(m, P, r) for r in range(P)]
    owned = np.concatenate([p.local_nodes[:p.n_owned] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(1, m.n_nodes + 1))
    for p in parts:
        lo, hi = p.owned_range
        assert np.array_equal(p.local_nodes[:p.n_owned], np.arange(lo, hi + 1))
        gh = p.local_nodes[p.n_owned:]
        assert np.all(np.diff(gh) > 0) and np.all((gh < lo) | (gh > hi))
        # every element touching an owned node is local
        touches = ((m.conn >= lo) & (m.conn <= hi)).any(axis=1)
        assert np.array_equal(np.nonzero(touches)[0], p.elems)
        assert np.array_equal(p.local_nodes[p.conn_local - 1], m.conn[p.elems])
        for s, ids in p.send.items():
            q = parts[s]
            assert np.array_equal(p.local_nodes[ids - 1], q.local_nodes[q.recv[p.rank] - 1])
        for s, ids in p.recv.items():
            assert np.all(ids > p.n_owned) and np.all(np.diff(ids) == 1)   # contiguous ghost segment per neighbour


def test_partitioned_matvec_equals_global(oracle, jf):
    """Owner-computes with ghost elements reproduces the global K.u on owned rows (oracle arithmetic, CPU)."""
    m = jf.mesh.tet10_kuhn(3, 2, 2)
    u = jf.mesh.test_vector(m.n_dofs)
    y = oracle.matfree(10, m.coords, m.conn, u).reshape(-1, 3)
    for P in (2, 3):
        for r in range(P):
            p = jf.mesh.partition_mesh(m, P, r)
            ul = u.reshape(-1, 3)[p.local_nodes - 1].ravel()
            yl = oracle.matfree(10, m.coords[p.local_nodes - 1], p.conn_local, ul).reshape(-1, 3)
            ref = y[p.local_nodes[:p.n_owned] - 1]
            assert np.abs(yl[:p.n_owned] - ref).max() < 1e-12 * np.abs(y).max()


def test_partition_receive_lists_are_contiguous_ghost_order(jf):
    """Precondition of the halo exchange fused into the patch kernel: concatenated over the neighbours in ascending rank
    order, the receive lists are exactly the ghost nodes n_owned+1 .. n_local in order, so that ghost node g reads landing
    slot g - n_owned (csrc/comm.cu: recv_contiguous)."""
    import numpy as np
    for m, world in ((jf.mesh.tet10_kuhn(4, 3, 9), 3), (jf.mesh.hex8_lattice(5, 4, 17, 0.1), 4), (jf.mesh.tet10_kuhn(6, 2, 2), 2)):
        for rank in range(world):
            p = jf.mesh.partition_mesh(m, world, rank)
            got = np.concatenate([p.recv[s] for s in sorted(p.recv)]) if p.recv else np.zeros(0, dtype=np.int64)
            assert np.array_equal(got, np.arange(p.n_owned + 1, p.local_nodes.size + 1))
            for s, ids in p.send.items():
                assert ids.min() >= 1 and ids.max() <= p.n_owned            # only owned nodes are sent


GMSH_TET10 = """$MeshFormat
4.1 0 8
$EndMeshFormat
$PhysicalNames
2
2 7 "FixedEnd"
3 1 "Body"
$EndPhysicalNames
$Entities
0 0 1 1
7 0 0 0 0 1 1 1 7 0
1 0 0 0 1 1 1 1 1 0
$EndEntities
$Nodes
2 11 1 40
2 7 0 3
1
2
3
0 0 0
1 0 0
0 1 0
3 1 0 8
4
10
11
12
13
14
15
40
0 0 1
0.5 0 0
0.5 0.5 0
0 0.5 0
0 0 0.5
0 0.5 0.5
0.5 0 0.5
9 9 9
$EndNodes
$Elements
2 2 1 2
2 7 2 1
1 1 2 3
3 1 11 1
2 1 2 3 4 10 11 12 13 14 15
$EndElements
"""


def test_gmsh_reader_tet10_order_sets_and_volume(tmp_path, oracle, jf):
    """MSH 4.1: Tet10 edge order converted to the reference's (Gmsh lists edge 2-3 before 1-3), dense renumbering, physical
    names -> element / node sets; the element's stiffness is symmetric with 6 rigid-body modes only if the order is right."""
    import numpy as np
    f = tmp_path / "one.msh"
    f.write_text(GMSH_TET10)
    m = jf.mesh.read_gmsh_msh(str(f))
    assert m.elem_type == 10 and m.n_elems == 1 and m.n_nodes == 10          # node 40 is unused and dropped
    X = m.coords[m.conn[0] - 1]
    for k, (a, b) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):
        assert np.allclose(X[4 + k], 0.5 * (X[a] + X[b]))                    # mid-edge nodes in the reference's order
    assert list(m.elem_sets) == ["Body"] and list(m.elem_sets["Body"]) == [1]
    assert list(m.node_sets["FixedEnd"]) == [1, 2, 3]
    K = oracle.element(10, X)[0]
    assert np.abs(K - K.T).max() < 1e-9 * np.abs(K).max()
    rb = np.tile([1.0, 0.0, 0.0], 10)
    assert np.abs(K @ rb).max() < 1e-9 * np.abs(K).max()


def _split_top_level(s):
    """split a parameter list at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_shim_binds_declared_symbols_with_matching_arity():
    """The Julia shim cannot be executed here (no Julia toolchain), so it is checked statically: every `ccall` names a
    function include/jfem_b200.h declares, with as many argument types as the C prototype has parameters, and pointer /
    integer / floating-point classes in the same positions."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "jfem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(jfem_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = [a for a in _split_top_level(" ".join(m.group(2).split())) if a != "void"]
        protos[m.group(1)] = args
    jl = open(os.path.join(root, "juliafem.jl_b200", "julia", "JuliaFEMB200.jl")).read()
    calls = list(re.finditer(r"ccall\(\(:(jfem_\w+),\s*libjfem\),\s*(\w+),\s*\(", jl))
    assert len(calls) >= 14

    def cls_c(a):
        if "*" in a:
            return "ptr"
        return "flt" if re.search(r"\bdouble\b", a) else "int"

    def cls_j(t):
        if t.startswith(("Ptr", "Ref", "Cstring")):
            return "ptr"
        return "flt" if t in ("Cdouble", "Float64") else "int"

    for m in calls:
        name = m.group(1)
        assert name in protos, f"{name} is not declared in include/jfem_b200.h"
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(jl[i], 0)
            i += 1
        types = _split_top_level(jl[m.end():i - 1])
        types = [t for t in types if t]
        assert len(types) == len(protos[name]), f"{name}: {len(types)} ccall types vs {len(protos[name])} C parameters"
        for t, a in zip(types, protos[name]):
            assert cls_j(t) == cls_c(a), f"{name}: Julia type {t} bound to C parameter '{a}'"


@pytest.mark.parametrize("et,dims,box", [(10, (3, 2, 7), (1.5, 1.0, 3.5)), (8, (4, 3, 11), 0.25), (10, (2, 2, 2), (1.0, 1.0, 1.0))])
@pytest.mark.parametrize("P", [1, 2, 3, 4, 8])
def test_lattice_window_partition_equals_partition_of_the_full_mesh(jf, et, dims, box, P):
    """bench.py / PartitionedProblem.from_lattice build only a WINDOW of the lattice per rank (a 100 M-DOF mesh per process is
    not affordable).  The partition computed on the window must be the partition of the whole mesh: same local nodes (owned
    then ghosts), coordinates, local connectivity (as a set of elements, same order), send / receive lists."""
    M = jf.mesh
    full = M.tet10_kuhn(*dims, *box) if et == 10 else M.hex8_lattice(*dims, box)
    elems_seen = np.zeros(full.n_elems, dtype=int)
    for r in range(P):
        ref = M.partition_mesh(full, P, r)
        w, off, nn, ne = M.lattice_window(et, dims, box, P, r)
        assert nn == full.n_nodes and ne == full.n_elems
        got = M.partition_mesh(w, P, r, off, nn)
        assert got.n_owned == ref.n_owned and got.owned_range == ref.owned_range
        assert np.array_equal(got.local_nodes, ref.local_nodes)
        assert np.allclose(w.coords[got.local_nodes - 1 - off], full.coords[ref.local_nodes - 1], rtol=0, atol=1e-13)
        assert np.array_equal(got.conn_local, ref.conn_local)
        assert sorted(got.send) == sorted(ref.send) and sorted(got.recv) == sorted(ref.recv)
        for s in ref.send:
            assert np.array_equal(got.send[s], ref.send[s])
        for s in ref.recv:
            assert np.array_equal(got.recv[s], ref.recv[s])
        elems_seen[ref.elems] += 1
    assert elems_seen.min() >= 1                                   # every element is computed by at least one rank


def test_preprocess_helpers_of_the_api_mirror(jf):
    """create_elements / create_nodal_elements / add_node_to_node_set! as examples/linear_static.jl:26-27,58-77 uses them."""
    from juliafem.jl_b200 import api as A
    m = jf.mesh.tet10_kuhn(2, 1, 1)
    m.elem_sets["LEFT"] = np.array([1, 2, 3])
    m.elem_sets["RIGHT"] = np.array([3, 7])
    els = A.create_elements(m, "OTHER")                         # no such set: every element (MED family 0)
    assert len(els) == m.n_elems and els[4].id == 5 and els[4].connectivity == tuple(int(c) for c in m.conn[4])
    assert els[0].fields["geometry"].shape == (3, 10) and np.array_equal(els[0].fields["geometry"].T, m.coords[m.conn[0] - 1])
    assert [e.id for e in A.create_elements(m, "LEFT", "RIGHT")] == [1, 2, 3, 7]
    A.add_node_to_node_set_(m, "fixed", 5, 2)
    A.add_node_to_node_set_(m, "fixed", 2, 9)
    assert list(m.node_sets["fixed"]) == [2, 5, 9]
    pois = A.create_nodal_elements(m, "fixed")
    assert [p.topology for p in pois] == [A.Poi1] * 3 and [p.connectivity for p in pois] == [(2,), (5,), (9,)]


def test_argument_validation_at_the_abi_without_a_device(jf):
    """Edge cases the C ABI refuses before it ever touches a device (so they can be checked here): null / empty inputs, a bad
    index base, more nodes than the 27-bit node ids of the patch tables can hold, and calls on a null handle."""
    import ctypes as C
    from juliafem.jl_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    xyz = np.zeros(3)
    ptr = xyz.ctypes.data_as(C.c_void_p)
    def err():
        return L.jfem_last_error().decode()
    assert L.jfem_create(C.byref(h), 0, 10, 0, 0, ptr, None, 1) == 1 and "bad arguments" in err()          # no nodes
    assert L.jfem_create(C.byref(h), 0, 10, 1, 0, None, None, 1) == 1                                        # null coordinates
    assert L.jfem_create(C.byref(h), 0, 10, 1, 1, ptr, None, 1) == 1                                         # elements without connectivity
    assert L.jfem_create(C.byref(h), 0, 10, 1, 0, ptr, None, 2) == 1                                         # index base must be 0 or 1
    assert L.jfem_create(None, 0, 10, 1, 0, ptr, None, 1) == 1 and "null output" in err()
    assert L.jfem_create(C.byref(h), 0, 10, 1 << 27, 0, ptr, None, 1) == 1 and "nodes per handle" in err()   # maximum size
    assert h.value is None
    assert L.jfem_set_option(None, b"patch_elems", 256.0) == 1 and "null handle" in err()
    assert L.jfem_matvec(None, ptr, ptr, 0, 0) == 1
    assert L.jfem_destroy(None) in (0, 1)                                                                     # harmless
