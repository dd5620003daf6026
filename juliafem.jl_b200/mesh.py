"""Deterministic synthetic meshes, minimal Abaqus / Gmsh readers and the node-range partitioner.

Host-side helpers feeding the flat arrays the C ABI takes (include/jfem_b200.h).  All ids
returned here are **1-based**, as at the reference's boundary (`Element.connectivity`,
src/elements/elements.jl:72; dof = 3*(node-1)+c, src/assembly/problems.jl:476).

Mesh definitions follow SURVEY.md §8(d):
  * Hex8 lattice: node id (k-1)*nx*ny + (j-1)*nx + i, element (n1,n1+1,n1+1+nx,n1+nx,+nx*ny...)
    exactly as benchmarks/multigpu_mpi_benchmark.jl:75-102 (`create_hex_mesh`, nx = #nodes in x).
  * Tet10 lattice: P2 lattice of (2cx+1)(2cy+1)(2cz+1) nodes with the same x-fastest numbering,
    every cell split into 6 Kuhn tets around the main diagonal, node order per
    src/basis/lagrange_generated.jl:261-262 (5=(1-2) 6=(2-3) 7=(1-3) 8=(1-4) 9=(2-4) 10=(3-4)).
  * partition: contiguous node ranges ceil(nn/P) + ghost elements, as
    benchmarks/multigpu_mpi_benchmark.jl:120-229 (`partition_mesh_for_rank`).
"""
from __future__ import annotations

import itertools
import re
from dataclasses import dataclass, field

import numpy as np

TET4, HEX8, TET10 = 4, 8, 10


@dataclass
class Mesh:
    elem_type: int            # nodes per element: 4, 8, 10
    coords: np.ndarray        # (n_nodes, 3) float64, row i = node id i+1
    conn: np.ndarray          # (n_elems, nnpe) int32, 1-based node ids
    node_sets: dict = field(default_factory=dict)
    elem_sets: dict = field(default_factory=dict)

    @property
    def n_nodes(self) -> int:
        return int(self.coords.shape[0])

    @property
    def n_elems(self) -> int:
        return int(self.conn.shape[0])

    @property
    def n_dofs(self) -> int:
        return 3 * self.n_nodes


def hex8_lattice(nx: int, ny: int, nz: int, h: float = 1.0) -> Mesh:
    """Structured Hex8 mesh on an nx*ny*nz NODE lattice (benchmarks/multigpu_mpi_benchmark.jl:75-102)."""
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    coords = np.stack([i.ravel() * h, j.ravel() * h, k.ravel() * h], axis=1).astype(np.float64)
    ke, je, ie = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    n1 = (ke * nx * ny + je * nx + ie + 1).ravel().astype(np.int64)
    n2 = n1 + 1
    n3 = n2 + nx
    n4 = n1 + nx
    lay = nx * ny
    conn = np.stack([n1, n2, n3, n4, n1 + lay, n2 + lay, n3 + lay, n4 + lay], axis=1).astype(np.int32)
    return Mesh(HEX8, coords, conn)


def _kuhn_tets():
    """6 tets of the unit cell in doubled-integer corner coordinates, positively oriented."""
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] += 2
            v.append(nxt)
        v = np.array(v)
        d = np.linalg.det((v[1:] - v[0]).astype(float))
        if d < 0:
            v[[2, 3]] = v[[3, 2]]
        tets.append(v)
    return np.array(tets)  # (6, 4, 3) in P2-lattice index units


def tet10_kuhn(cx: int, cy: int, cz: int, lx: float = 1.0, ly: float | None = None, lz: float | None = None) -> Mesh:
    """Tet10 mesh of a box with cx*cy*cz cells, 6 Kuhn tets per cell."""
    ly = lx * cy / cx if ly is None else ly
    lz = lx * cz / cx if lz is None else lz
    px, py, pz = 2 * cx + 1, 2 * cy + 1, 2 * cz + 1
    k, j, i = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
    coords = np.stack([i.ravel() * (lx / (2 * cx)), j.ravel() * (ly / (2 * cy)), k.ravel() * (lz / (2 * cz))], axis=1)
    tets = _kuhn_tets()                                  # (6,4,3)
    edges = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
    local = np.concatenate([tets, np.stack([(tets[:, a] + tets[:, b]) // 2 for a, b in edges], axis=1)], axis=1)  # (6,10,3)
    ck, cj, ci = np.meshgrid(np.arange(cz), np.arange(cy), np.arange(cx), indexing="ij")
    org = np.stack([2 * ci.ravel(), 2 * cj.ravel(), 2 * ck.ravel()], axis=1).astype(np.int64)  # (ncell,3)
    idx = org[:, None, None, :] + local[None, :, :, :]   # (ncell,6,10,3)
    ids = idx[..., 2] * (px * py) + idx[..., 1] * px + idx[..., 0] + 1
    conn = ids.reshape(-1, 10).astype(np.int32)
    return Mesh(TET10, coords.astype(np.float64), conn)


def tet4_kuhn(cx: int, cy: int, cz: int, lx: float = 1.0) -> Mesh:
    """Tet4 companion of tet10_kuhn (vertices only, compact numbering)."""
    m = tet10_kuhn(cx, cy, cz, lx)
    conn4 = m.conn[:, :4]
    used = np.unique(conn4)
    remap = np.zeros(m.n_nodes + 1, dtype=np.int32)
    remap[used] = np.arange(1, used.size + 1, dtype=np.int32)
    return Mesh(TET4, m.coords[used - 1], remap[conn4])


_ABAQUS_TYPES = {"C3D4": TET4, "C3D10": TET10, "C3D8": HEX8}


def read_abaqus_inp(path: str, elem_type: int | None = None) -> Mesh:
    """Minimal Abaqus reader: *NODE, *ELEMENT (C3D4/C3D10/C3D8), *NSET/*ELSET (incl. GENERATE).

    Mirrors what src/readers/parse_mesh.jl:8-60 / src/io/abaqus_reader.jl:13-67 extract for volume
    elements.  Node ids are renumbered densely by sorted order, as the GPU extension does
    (ext/JuliaFEMCUDAExt.jl:100-108)."""
    nodes, elems, nsets, elsets = {}, {}, {}, {}
    mode, et, cur, gen, pending = None, None, None, False, []
    with open(path) as fh:
        for raw in fh:
            line = raw.strip()
            if not line or line.startswith("**"):
                continue
            if line.startswith("*"):
                pending = []
                head = line.upper()
                opts = dict(re.findall(r"(\w+)\s*=\s*([^,\s]+)", line, flags=re.I))
                opts = {k.upper(): v for k, v in opts.items()}
                if head.startswith("*NODE") and not head.startswith("*NODE "):
                    mode = "node"
                    cur = opts.get("NSET")
                    if cur:
                        nsets.setdefault(cur, [])
                elif head.startswith("*ELEMENT"):
                    t = opts.get("TYPE", "").upper()
                    et = _ABAQUS_TYPES.get(t)
                    mode = "elem" if et else None
                    cur = opts.get("ELSET")
                    if cur and et:
                        elsets.setdefault(cur, [])
                elif head.startswith("*NSET"):
                    mode, cur, gen = "nset", opts.get("NSET"), "GENERATE" in head
                    nsets.setdefault(cur, [])
                elif head.startswith("*ELSET"):
                    mode, cur, gen = "elset", opts.get("ELSET"), "GENERATE" in head
                    elsets.setdefault(cur, [])
                else:
                    mode = None
                continue
            vals = [v for v in (s.strip() for s in line.split(",")) if v]
            if mode == "node":
                nid = int(vals[0])
                nodes[nid] = [float(v) for v in vals[1:4]] + [0.0] * (4 - len(vals))
                if cur:
                    nsets[cur].append(nid)
            elif mode == "elem":
                pending += [int(v) for v in vals]
                if len(pending) >= et + 1:
                    elems[pending[0]] = (et, pending[1:et + 1])
                    if cur:
                        elsets[cur].append(pending[0])
                    pending = []
            elif mode in ("nset", "elset"):
                tgt = nsets if mode == "nset" else elsets
                if gen:
                    a, b = int(vals[0]), int(vals[1])
                    st = int(vals[2]) if len(vals) > 2 else 1
                    tgt[cur] += list(range(a, b + 1, st))
                else:
                    tgt[cur] += [int(v) for v in vals if re.fullmatch(r"-?\d+", v)]
    if elem_type is None:
        kinds = {t for t, _ in elems.values()}
        elem_type = max(kinds)
    keep = sorted(eid for eid, (t, _) in elems.items() if t == elem_type)
    used = sorted({n for eid in keep for n in elems[eid][1]})
    remap = {n: i + 1 for i, n in enumerate(used)}
    coords = np.array([nodes[n][:3] for n in used], dtype=np.float64)
    conn = np.array([[remap[n] for n in elems[eid][1]] for eid in keep], dtype=np.int32)
    eremap = {eid: i + 1 for i, eid in enumerate(keep)}
    m = Mesh(elem_type, coords, conn)
    m.node_sets = {k: np.array(sorted(remap[n] for n in v if n in remap), dtype=np.int64) for k, v in nsets.items()}
    m.elem_sets = {k: np.array(sorted(eremap[e] for e in v if e in eremap), dtype=np.int64) for k, v in elsets.items()}
    return m


# Gmsh element type -> (nodes per element, kind); volume kinds are what the C ABI takes, surface kinds only feed node sets
_GMSH_TYPES = {4: (4, "vol"), 5: (8, "vol"), 11: (10, "vol"), 2: (3, "surf"), 3: (4, "surf"), 9: (6, "surf"), 16: (8, "surf"), 10: (9, "surf")}
_GMSH_TET10_TO_ABAQUS = [0, 1, 2, 3, 4, 5, 6, 7, 9, 8]   # Gmsh numbers edge (2-3) before (1-3); lagrange_generated.jl:261-262 the other way


def read_gmsh_msh(path: str, elem_type: int | None = None) -> Mesh:
    """Gmsh MSH 4.1 (ASCII) reader: $PhysicalNames, $Nodes, $Elements.

    Same walk as the reference's reader (src/gmsh_reader.jl:28-155: entity blocks, node ids first then coordinates, the
    entity tag of an element block names its group through $PhysicalNames), extended from Tet4 (type 4) to Hex8 (type 5)
    and Tet10 (type 11, reordered to the reference's / Abaqus edge order).  Surface elements (triangles / quadrangles) are
    not returned as elements; the nodes of every surface block go into ``node_sets`` under the block's physical name, which
    is what Dirichlet / traction set-up needs.  Node ids are renumbered densely by sorted order (ext/JuliaFEMCUDAExt.jl:100-108)."""
    names, nodes, vol, surf_nodes, groups = {}, {}, [], {}, {}
    with open(path) as fh:
        lines = iter(fh.read().splitlines())
    for line in lines:
        line = line.strip()
        if line == "$MeshFormat":
            ver = next(lines).split()
            if not ver[0].startswith("4") or int(ver[1]) != 0:
                raise ValueError(f"{path}: only ASCII MSH 4.x files are supported (found version {ver[0]}, file-type {ver[1]})")
        elif line == "$PhysicalNames":
            for _ in range(int(next(lines))):
                dim, tag, name = next(lines).split(maxsplit=2)
                names[(int(dim), int(tag))] = name.strip().strip('"')
        elif line == "$Nodes":
            nblocks = int(next(lines).split()[0])
            for _ in range(nblocks):
                nb = int(next(lines).split()[3])
                ids = [int(next(lines)) for _ in range(nb)]
                for nid in ids:
                    nodes[nid] = [float(v) for v in next(lines).split()[:3]]
        elif line == "$Elements":
            nblocks = int(next(lines).split()[0])
            for _ in range(nblocks):
                dim, tag, gtype, nb = (int(v) for v in next(lines).split())
                nn, kind = _GMSH_TYPES.get(gtype, (0, None))
                name = names.get((dim, tag), f"Group_{tag}")
                for _ in range(nb):
                    parts = next(lines).split()
                    if kind == "vol":
                        c = [int(v) for v in parts[1:1 + nn]]
                        if nn == 10:
                            c = [c[k] for k in _GMSH_TET10_TO_ABAQUS]
                        vol.append((nn, c))
                        groups.setdefault(name, []).append(len(vol))
                    elif kind == "surf":
                        surf_nodes.setdefault(name, set()).update(int(v) for v in parts[1:1 + nn])
    if not vol:
        raise ValueError(f"{path}: no Tet4 / Hex8 / Tet10 elements found")
    if elem_type is None:
        elem_type = max(nn for nn, _ in vol)
    keep = [i for i, (nn, _) in enumerate(vol) if nn == elem_type]
    used = sorted({n for i in keep for n in vol[i][1]})
    remap = {n: i + 1 for i, n in enumerate(used)}
    coords = np.array([nodes[n] for n in used], dtype=np.float64)
    conn = np.array([[remap[n] for n in vol[i][1]] for i in keep], dtype=np.int32)
    eremap = {i + 1: k + 1 for k, i in enumerate(keep)}
    m = Mesh(elem_type, coords, conn)
    m.elem_sets = {k: np.array(sorted(eremap[e] for e in v if e in eremap), dtype=np.int64) for k, v in groups.items()}
    m.elem_sets = {k: v for k, v in m.elem_sets.items() if v.size}
    m.node_sets = {k: np.array(sorted(remap[n] for n in v if n in remap), dtype=np.int64) for k, v in surf_nodes.items()}
    return m


# Code Aster (.med) element names and node orders -> the reference's (Abaqus) conventions.
# Names: src/io/aster_reader.jl:27-47 (`med_element_names`); permutations: :16-23 (`med_connectivity`), applied as
# new = old[perm] (src/preprocess.jl:301-309).
_MED_TYPES = {"TE4": (TET4, [4, 3, 1, 2]), "T10": (TET10, [4, 3, 1, 2, 10, 7, 8, 9, 6, 5]), "HE8": (HEX8, [4, 8, 7, 3, 1, 5, 6, 2])}


def read_med(path: str, mesh_name: str | None = None, elem_type: int | None = None, reorder_element_connectivity: bool = True) -> Mesh:
    """`aster_read_mesh(filename, mesh_name)` for volume meshes (src/io/aster_reader.jl:56-82 over
    src/readers/read_aster_mesh.jl:118-222): nodes from ENS_MAA/<mesh>/<increment>/NOE/{COO,NUM,FAM} (coordinates are
    stored non-interlaced: all x, then all y, then all z, :127-128), elements from .../MAI/<TE4|T10|HE8>/{NOD,NUM,FAM}
    (connectivity non-interlaced as well, :149-150), node / element sets from the family groups FAS/<mesh>/{NOEUD,ELEME}
    (family 0 = "OTHER", :69, :131-133).  Lower-dimensional cells (SE3, TR6, ...: the edge and surface meshes FreeCAD
    exports, which examples/linear_static.jl:35-39 filters out) are not returned.  The HDF5 container is parsed by
    h5lite (no h5py in this image).  Node ids are renumbered densely in ascending order of the file's ids."""
    from . import h5lite
    tree = h5lite.read(path)
    names = sorted(tree["FAS"].keys()) if "FAS" in tree else sorted(tree["ENS_MAA"].keys())
    if mesh_name is None:
        if len(names) != 1:
            raise ValueError("several meshes found from med, pick one: " + ", ".join(names))
        mesh_name = names[0]
    elif mesh_name not in names:
        raise ValueError(f"Mesh {mesh_name} not found. Available meshes: " + ", ".join(names))
    incs = tree["ENS_MAA"][mesh_name]
    incs = {k: v for k, v in incs.items() if isinstance(v, dict)}
    if len(incs) != 1:
        raise ValueError("exactly one mesh increment expected")
    inc = next(iter(incs.values()))

    def family_names(kind):
        out = {0: ["OTHER"]}
        fas = tree.get("FAS", {}).get(mesh_name, {}).get(kind) or {}
        for i, k in enumerate(sorted(fas.keys())):
            fid = int(k.split("_")[1]) if k.startswith("FAM") else -(i + 1)
            gro = (fas[k] or {}).get("GRO") if isinstance(fas[k], dict) else None
            nom = gro.get("NOM") if gro else None
            if nom is None:
                out[fid] = [""]
            else:
                out[fid] = [bytes(np.asarray(r, dtype=np.int8).astype(np.uint8)).split(b"\0")[0].decode("ascii").strip() for r in np.atleast_2d(nom)]
        return out

    noe = inc["NOE"]
    nn = int(noe["FAM"].size)
    dim = noe["COO"].size // nn
    coords = np.zeros((nn, 3))
    coords[:, :dim] = noe["COO"].reshape(dim, nn).T
    node_ids = np.asarray(noe["NUM"], dtype=np.int64) if noe.get("NUM") is not None else np.arange(1, nn + 1, dtype=np.int64)
    order = np.argsort(node_ids, kind="stable")
    remap = {int(n): i + 1 for i, n in enumerate(node_ids[order])}
    coords = coords[order]
    nsets: dict = {}
    nfam = family_names("NOEUD")
    for fam, nid in zip(np.asarray(noe["FAM"])[order], range(1, nn + 1)):
        for name in nfam.get(int(fam), ["OTHER"]):
            nsets.setdefault(name, []).append(nid)
    mai = inc.get("MAI", {})
    vol = [k for k in mai if k in _MED_TYPES and (elem_type is None or _MED_TYPES[k][0] == elem_type)]
    if not vol:
        raise ValueError("no TE4 / T10 / HE8 cells in " + path)
    key = max(vol, key=lambda k: mai[k]["FAM"].size)
    et, perm = _MED_TYPES[key]
    ne = int(mai[key]["FAM"].size)
    conn = np.asarray(mai[key]["NOD"], dtype=np.int64).reshape(et, ne).T
    if reorder_element_connectivity:
        conn = conn[:, np.array(perm) - 1]
    lut = np.zeros(int(node_ids.max()) + 1, dtype=np.int64)
    lut[node_ids[order]] = np.arange(1, nn + 1)
    conn = lut[conn].astype(np.int32)
    elem_ids = np.asarray(mai[key]["NUM"], dtype=np.int64) if mai[key].get("NUM") is not None else np.arange(1, ne + 1)
    eorder = np.argsort(elem_ids, kind="stable")
    conn = np.ascontiguousarray(conn[eorder])
    esets: dict = {}
    efam = family_names("ELEME")
    for i, fam in enumerate(np.asarray(mai[key]["FAM"])[eorder]):
        for name in efam.get(int(fam), ["OTHER"]):
            esets.setdefault(name, []).append(i + 1)
    m = Mesh(et, coords, conn)
    m.node_sets = {k: np.array(v, dtype=np.int64) for k, v in nsets.items()}
    m.elem_sets = {k: np.array(v, dtype=np.int64) for k, v in esets.items()}
    return m


def find_nearest_nodes(mesh: Mesh, point, npts: int = 1, node_set=None) -> np.ndarray:
    """`find_nearest_nodes(mesh, coords, npts; node_set)` (src/preprocess.jl:270-282): 1-based ids of the npts nodes
    closest to `point`, nearest first."""
    ids = np.arange(1, mesh.n_nodes + 1) if node_set is None else np.asarray(mesh.node_sets[node_set], dtype=np.int64)
    d = np.linalg.norm(mesh.coords[ids - 1] - np.asarray(point, dtype=np.float64)[None, :], axis=1)
    return ids[np.argsort(d, kind="stable")[:npts]]


def nodes_at_plane(mesh: Mesh, axis: int, distance: float, radius: float = 6.0) -> np.ndarray:
    """1-based ids of the nodes with |x_axis - distance| <= radius: the helper examples/linear_static.jl:46-54 defines
    (`isapprox(coords[vector_id], distance, atol=radius)`); axis is 0-based here."""
    return np.nonzero(np.abs(mesh.coords[:, axis] - distance) <= radius)[0].astype(np.int64) + 1


def clamp_dofs(mesh: Mesh, axis: int = 0, value: float = 0.0, tol: float = 1e-12) -> np.ndarray:
    """1-based dof ids of all three components of nodes on the plane x_axis == value
    (demos/cantilever_physics_gpu.jl:88-93 clamps x = 0)."""
    nodes = np.nonzero(np.abs(mesh.coords[:, axis] - value) <= tol)[0].astype(np.int64)
    return (3 * nodes[:, None] + np.arange(1, 4)[None, :]).ravel()


def test_vector(n_dofs: int, fixed_dofs: np.ndarray | None = None, seed: int = 12345) -> np.ndarray:
    """u = 1e-3*(2U-1), U ~ default_rng(12345).random(nDOF), fixed dofs zeroed."""
    u = 1e-3 * (2.0 * np.random.default_rng(seed).random(n_dofs) - 1.0)
    if fixed_dofs is not None and len(fixed_dofs):
        u[np.asarray(fixed_dofs) - 1] = 0.0
    return u


@dataclass
class Partition:
    rank: int
    n_ranks: int
    owned_range: tuple          # (first, last) 1-based global node ids owned (inclusive)
    local_nodes: np.ndarray     # global 1-based ids: owned (ascending) then ghosts (ascending)
    n_owned: int
    elems: np.ndarray           # global 0-based element indices computed on this rank
    conn_local: np.ndarray      # (n_local_elems, nnpe) int32, 1-based LOCAL node ids
    send: dict                  # neighbour rank -> local (1-based) owned node ids to send, ascending global id
    recv: dict                  # neighbour rank -> local (1-based) ghost node ids to receive into, ascending global id


def partition_mesh(mesh: Mesh, n_ranks: int, rank: int, node_offset: int = 0, n_nodes_global: int | None = None) -> Partition:
    """Owner-computes partition with one layer of ghost elements
    (benchmarks/multigpu_mpi_benchmark.jl:120-229): rank r owns the contiguous node-id range
    [r*ceil(nn/P)+1, min((r+1)*ceil(nn/P), nn)]; local elements = those touching >= 1 owned node;
    ghosts = their non-owned nodes, sorted ascending after the owned ones; rank r sends to s the
    owned nodes that appear in s's ghost list (i.e. owned nodes of elements that also touch
    s-owned nodes).

    `mesh` may be a WINDOW of a larger mesh (lattice_window): its node i+1 is global node i+1+node_offset of a mesh with
    n_nodes_global nodes, and it must contain every element that touches a node owned by `rank`."""
    nn = mesh.n_nodes if n_nodes_global is None else int(n_nodes_global)
    per = -(-nn // n_ranks)
    lo, hi = rank * per + 1, min((rank + 1) * per, nn)
    cg = mesh.conn.astype(np.int64) + node_offset             # global ids
    owner = (cg - 1) // per                                   # (ne, nnpe)
    mine = (owner == rank).any(axis=1)
    elems = np.nonzero(mine)[0]
    c = cg[elems]
    touched = np.unique(c)
    ghosts = touched[(touched < lo) | (touched > hi)]
    owned = np.arange(lo, hi + 1, dtype=np.int64)
    local_nodes = np.concatenate([owned, ghosts])

    def g2l(ids):                                             # global id -> 1-based local id (owned first, then ghosts)
        ids = np.asarray(ids, dtype=np.int64)
        out = ids - lo + 1
        g = (ids < lo) | (ids > hi)
        out[g] = owned.size + 1 + np.searchsorted(ghosts, ids[g])
        return out

    conn_local = g2l(c.ravel()).reshape(c.shape).astype(np.int32)
    recv, send = {}, {}
    gown = (ghosts - 1) // per
    for s in np.unique(gown):
        recv[int(s)] = g2l(ghosts[gown == s])
    # what I must send to s: my owned nodes that are ghosts on s <=> owned nodes of elements touching s-owned nodes
    oe = owner[elems]
    for s in np.unique(oe):
        s = int(s)
        if s == rank:
            continue
        has_s = (oe == s).any(axis=1)
        cand = np.unique(c[has_s])
        cand = cand[(cand >= lo) & (cand <= hi)]
        if cand.size:
            send[s] = g2l(cand)
    return Partition(rank, n_ranks, (lo, hi), local_nodes, owned.size, elems, conn_local, send, recv)


def lattice_window(elem_type: int, dims, box, n_ranks: int, rank: int):
    """The part of a lattice mesh (hex8_lattice / tet10_kuhn with the given dims and box) that rank `rank` of a
    contiguous-node-range partition needs, built WITHOUT the global mesh (a 100 M-DOF lattice per process would cost
    gigabytes and minutes): the node layers of every element that touches an owned node.  Returns
    (window mesh, node_offset, n_nodes_global, n_elems_global); window node i+1 = global node i+1+node_offset.
    Feed it to partition_mesh(window, P, r, node_offset, n_nodes_global)."""
    if elem_type == HEX8:
        nx, ny, nz = dims                                      # nodes per direction
        h = box if np.isscalar(box) else box[0] / (nx - 1)
        lay, nn = nx * ny, nx * ny * nz
        per = -(-nn // n_ranks)
        lo, hi = rank * per + 1, min((rank + 1) * per, nn)
        ka, kb = (lo - 1) // lay, (hi - 1) // lay              # owned node layers
        wa, wb = max(ka - 1, 0), min(kb + 1, nz - 1)
        m = hex8_lattice(nx, ny, wb - wa + 1, h)
        m.coords[:, 2] += wa * h
        return m, wa * lay, nn, (nx - 1) * (ny - 1) * (nz - 1)
    cx, cy, cz = dims
    lx, ly, lz = box
    px, py, pz = 2 * cx + 1, 2 * cy + 1, 2 * cz + 1
    lay, nn = px * py, px * py * pz
    per = -(-nn // n_ranks)
    lo, hi = rank * per + 1, min((rank + 1) * per, nn)
    ka, kb = (lo - 1) // lay, (hi - 1) // lay
    ca, cb = max((ka - 1) // 2, 0), min(kb // 2, cz - 1)       # cell layers [2c, 2c+2] meeting [ka, kb]
    m = tet10_kuhn(cx, cy, cb - ca + 1, lx, ly, lz * (cb - ca + 1) / cz)
    m.coords[:, 2] += ca * (lz / cz)
    return m, 2 * ca * lay, nn, 6 * cx * cy * cz


def hashed_vector(dofs0: np.ndarray, fixed_mask: np.ndarray | None = None) -> np.ndarray:
    """Deterministic pseudo-random test vector defined per GLOBAL 0-based dof id, so that every rank can evaluate its own
    part without the global vector: u = 1e-3 * (2 U - 1), U = frac(golden-ratio hash of the id)."""
    g = np.asarray(dofs0, dtype=np.uint64)
    hsh = (g * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)
    u = 1e-3 * (2.0 * (hsh.astype(np.float64) / float(1 << 53)) - 1.0)
    if fixed_mask is not None:
        u[fixed_mask] = 0.0
    return u
