"""XDMF results writer -- the step right after the solve (mirrors `Xdmf` / `update_xdmf!` of the reference,
src/io.jl:7-72, 360-518).

Same document structure as the reference writes:

    <Xdmf xmlns:xi="http://www.w3.org/2001/XInclude" Version="3.0">
      <Domain>
        <Grid CollectionType="Temporal" GridType="Collection" Name="Time">          (src/io.jl:402-408)
          <Grid GridType="Collection" Name="Problems" CollectionType="Spatial">     (one per time, :360-376)
            <Time Value="0.0"/>
            <Grid Name="<problem name>">                                            (:419-421)
              <Geometry Type="XYZ"> <DataItem .../> </Geometry>                     (:424-436)
              <Topology TopologyType="Mixed"> <DataItem .../> </Topology>           (:441-459: element code, then 0-based nodes)
              <Attribute Name="Displacement" Center="Node" AttributeType="Vector"> <DataItem .../> </Attribute>   (:484-514)

Heavy data goes to `<name>.h5` as flat `/DataItem_N` datasets (src/io.jl:298-312) through the HDF5 writer of h5lite.py (no
h5py in this image), or inline with `format="XML"` (:285-290).  `DataItem/@Dimensions` lists the dimensions slowest first, as
the reference does by reversing Julia's column-major size (:271).

Nodal fields come straight from the solver: displacement (n_nodes x 3) and the least-squares nodal stress / strain of
`jfem_nodal_recover` (n_nodes x 6 in the order 11 22 33 12 23 13 = the reference's Voigt order, src/problems_elasticity.jl:189-197;
stored as `Tensor6` like the reference's 6-component fields).
"""
from __future__ import annotations

import os
from xml.etree import ElementTree as ET

import numpy as np

from . import h5lite

# element codes of the "Mixed" topology (src/io.jl:331-352)
XDMF_ELEMENT_CODE = {"Poi1": 1, "Seg2": 2, "Tri3": 4, "Quad4": 5, "Tet4": 6, "Pyr5": 7, "Wedge6": 8, "Hex8": 9,
                     "Seg3": 34, "Quad9": 35, "Tri6": 36, "Quad8": 37, "Tet10": 38, "Wedge15": 40, "Hex20": 48, "Hex27": 50}
_NNPE_NAME = {4: "Tet4", 8: "Hex8", 10: "Tet10"}


class Xdmf:
    """Xdmf(name; version="3.0", overwrite=false)  (src/io.jl:32-61): results go to `<name>.xmf` + `<name>.h5`."""

    def __init__(self, name: str, version: str = "3.0", overwrite: bool = False, format: str = "HDF"):
        if format not in ("HDF", "XML"):
            raise ValueError(f"Unsupported Xdmf big data format {format}")          # src/io.jl:292
        self.name, self.format = name, format
        for fn in (self.h5file, self.xmffile):
            if os.path.isfile(fn):
                if not overwrite:
                    raise FileExistsError(f"Result file {fn} exists, use Xdmf({name!r}, overwrite=True) to rewrite results")
                os.remove(fn)
        self.xml = ET.Element("Xdmf", {"xmlns:xi": "http://www.w3.org/2001/XInclude", "Version": version})
        self.hdf: dict[str, np.ndarray] = {}
        self.hdf_counter = 1

    @property
    def h5file(self) -> str:
        return self.name + ".h5"

    @property
    def xmffile(self) -> str:
        return self.name + ".xmf"

    # ---- DataItem (src/io.jl:268-312)
    def new_dataitem(self, data: np.ndarray) -> ET.Element:
        a = np.ascontiguousarray(data)
        if a.dtype.kind == "f":
            a, dtype = a.astype(np.float64, copy=False), "Float"
        elif a.dtype.kind in "iu":
            a, dtype = a.astype(np.int64, copy=False), "Int"
        else:
            raise TypeError(f"unsupported data type {a.dtype}")
        item = ET.Element("DataItem", {"DataType": dtype, "Dimensions": " ".join(str(s) for s in a.shape), "Format": self.format})
        if self.format == "HDF":
            path = f"DataItem_{self.hdf_counter}"
            while path in self.hdf:
                self.hdf_counter += 1
                path = f"DataItem_{self.hdf_counter}"
            self.hdf[path] = a
            item.text = f"{os.path.basename(self.h5file)}:/{path}"
        else:
            rows = a.reshape(a.shape[0], -1) if a.ndim > 1 else a.reshape(1, -1)
            fmt = "%.17g" if dtype == "Float" else "%d"
            item.text = "\n" + "\n".join(" ".join(fmt % v for v in row) for row in rows) + "\n"
        return item

    # ---- document structure
    def _temporal_collection(self) -> ET.Element:
        domain = self.xml.find("Domain")
        if domain is None:
            domain = ET.SubElement(self.xml, "Domain")
        tc = domain.find("Grid")
        if tc is None:
            tc = ET.SubElement(domain, "Grid", {"GridType": "Collection", "Name": "Time", "CollectionType": "Temporal"})
        assert tc.get("CollectionType") == "Temporal"
        return tc

    def _spatial_collection(self, time: float) -> ET.Element:
        tc = self._temporal_collection()
        for sc in tc.findall("Grid"):
            t = sc.find("Time")
            if t is not None and np.isclose(float(t.get("Value")), time):
                return sc
        sc = ET.SubElement(tc, "Grid", {"GridType": "Collection", "Name": "Problems", "CollectionType": "Spatial"})
        ET.SubElement(sc, "Time", {"Value": repr(float(time))})
        return sc

    def add_frame(self, name: str, time: float, coords: np.ndarray, conn: np.ndarray, elem_type: int | str, fields: dict | None = None) -> bool:
        """One `<Grid Name=name>` under the spatial collection of `time`: geometry, mixed topology, nodal fields.
        conn: 1-based node numbers (n_elems x nnpe, reference node order).  Returns False (with a warning, like the reference,
        src/io.jl:412-417) if a grid of that name already exists for that time."""
        sc = self._spatial_collection(time)
        for frame in sc.findall("Grid"):
            if frame.get("Name") == name:
                import warnings
                warnings.warn(f"Xdmf: Already found Grid with name {name} for time {time}, skipping.")
                return False
        coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
        conn = np.asarray(conn, dtype=np.int64)
        ename = elem_type if isinstance(elem_type, str) else _NNPE_NAME[int(elem_type)]
        code = XDMF_ELEMENT_CODE[ename]
        conn = conn.reshape(-1, conn.shape[-1])
        if conn.size and (conn.min() < 1 or conn.max() > coords.shape[0]):
            raise ValueError("connectivity refers to nodes outside the geometry")
        frame = ET.SubElement(sc, "Grid", {"Name": name})
        geometry = ET.SubElement(frame, "Geometry", {"Type": "XYZ"})
        geometry.append(self.new_dataitem(coords))
        # mixed topology: element code (Seg2 additionally its node count), then the 0-based nodes (src/io.jl:441-456)
        cols = [np.full((conn.shape[0], 1), code, dtype=np.int64)]
        if code == 2:
            cols.append(np.full((conn.shape[0], 1), conn.shape[1], dtype=np.int64))
        cols.append(conn - 1)
        topology = ET.SubElement(frame, "Topology", {"TopologyType": "Mixed", "NumberOfElements": str(conn.shape[0])})
        topology.append(self.new_dataitem(np.hstack(cols).ravel()))
        for fname, values in (fields or {}).items():
            v = np.asarray(values, dtype=np.float64)
            v = v.reshape(coords.shape[0], -1)
            if v.shape[1] == 2:                                   # 2D vectors are extended to 3 components (src/io.jl:497-503)
                v = np.hstack([v, np.zeros((v.shape[0], 1))])
            ftype = {1: "Scalar", 3: "Vector", 6: "Tensor6"}.get(v.shape[1])
            if ftype is None:
                raise ValueError(f"field {fname!r} has {v.shape[1]} components per node (supported: 1, 3, 6)")
            attr = ET.SubElement(frame, "Attribute", {"Name": fname[:1].upper() + fname[1:], "Center": "Node", "AttributeType": ftype})
            attr.append(self.new_dataitem(v))
        self.save()
        return True

    def save(self) -> None:
        """save!(xdmf) (src/io.jl:258-262): (re)write the XML and the HDF5 heavy data."""
        ET.indent(self.xml)
        with open(self.xmffile, "w", encoding="utf-8") as fh:
            fh.write('<?xml version="1.0" encoding="utf-8"?>\n')
            fh.write(ET.tostring(self.xml, encoding="unicode"))
            fh.write("\n")
        if self.format == "HDF":
            h5lite.write(self.h5file, self.hdf)

    close = save

    # ---- reading back (src/io.jl:199-256: read(xdmf, "/Domain/Grid/Grid[2]/Grid/Geometry"))
    def read(self, path: str) -> np.ndarray:
        """Data of the (first) DataItem below an XPath-like location, e.g. "/Domain/Grid/Grid[1]/Grid/Attribute[@Name='Displacement']"."""
        node = self.xml.find("." + path if path.startswith("/") else path)
        if node is None:
            raise KeyError(path)
        item = node if node.tag == "DataItem" else node.find("DataItem")
        return read_dataitem(item, os.path.dirname(os.path.abspath(self.xmffile)), cache=self.hdf)


def read_dataitem(item: ET.Element, directory: str, cache: dict | None = None) -> np.ndarray:
    dims = tuple(int(s) for s in item.get("Dimensions").split())
    dtype = np.float64 if item.get("DataType") == "Float" else np.int64
    if item.get("Format") == "HDF":
        fn, path = item.text.strip().split(":")
        key = path.lstrip("/")
        data = cache[key] if cache and key in cache else h5lite.read(os.path.join(directory, fn))[key]
        return np.asarray(data, dtype=dtype).reshape(dims)
    return np.array(item.text.split(), dtype=dtype).reshape(dims)


def read_xdmf(xmffile: str) -> list[dict]:
    """Frames of a result file written by `Xdmf`: [{"time", "name", "coords", "elements" (list of (code, 0-based nodes)), "fields"}]."""
    root = ET.parse(xmffile).getroot()
    d = os.path.dirname(os.path.abspath(xmffile))
    nn_of = {1: 1, 4: 3, 5: 4, 6: 4, 7: 5, 8: 6, 9: 8, 34: 3, 35: 9, 36: 6, 37: 8, 38: 10, 40: 15, 48: 20, 50: 27}
    out = []
    for sc in root.find("Domain").find("Grid").findall("Grid"):
        time = float(sc.find("Time").get("Value"))
        for frame in sc.findall("Grid"):
            topo = read_dataitem(frame.find("Topology").find("DataItem"), d)
            elems, p = [], 0
            while p < topo.size:
                code = int(topo[p]); p += 1
                n = nn_of.get(code)
                if code == 2:
                    n = int(topo[p]); p += 1
                elems.append((code, topo[p:p + n].copy())); p += n
            fields = {a.get("Name"): read_dataitem(a.find("DataItem"), d) for a in frame.findall("Attribute")}
            out.append({"time": time, "name": frame.get("Name"), "coords": read_dataitem(frame.find("Geometry").find("DataItem"), d),
                        "elements": elems, "fields": fields})
    return out


def update_xdmf_(xdmf: Xdmf, name: str, time: float, coords, conn, elem_type, u=None, stress=None, strain=None, extra: dict | None = None) -> bool:
    """update_xdmf!(xdmf, problem, time, fields) for the flat arrays of this package: displacement (n_dofs or n_nodes x 3) and the
    nodal stress / strain of `Handle.nodal_recover` (n_nodes x 6)."""
    fields = {}
    if u is not None:
        fields["displacement"] = np.asarray(u, dtype=np.float64).reshape(-1, 3)
    if stress is not None:
        fields["stress"] = stress
    if strain is not None:
        fields["strain"] = strain
    fields.update(extra or {})
    return xdmf.add_frame(name, time, coords, conn, elem_type, fields)
