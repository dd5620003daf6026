"""Python host mirror of the reference's user-facing interface for the elasticity hot path.

Julia is not available in the build image, so the host side above the C ABI is written here with the reference's
names and argument meaning (`!` becomes a trailing underscore); the Julia shim with the same calls is
juliafem.jl_b200/julia/JuliaFEMB200.jl.  Two seams are served:

  seam 1 (new API)      Physics / Element / add_elements_ / add_dirichlet_ / add_neumann_ / solve_(physics, backend=GPU())
                        -> initialize_backend / solve_backend_ -> ElasticitySolution
                        (src/physics.jl:128-156,222 ; src/backend/abstract.jl:138-145,181-217,228,241 ;
                         ext/JuliaFEMCUDAExt.jl:86-217,867-916 ; demos/cantilever_physics_gpu.jl:50-136)
  seam 2 (classic API)  Problem(Elasticity|Dirichlet) / update_ / add_elements_ / Analysis(Linear|Nonlinear) / assemble_ / run_
                        (src/assembly/problems.jl:14-40,95-105 ; src/assembly/assembly.jl:31 ; src/analysis.jl:7-13 ;
                         src/solvers.jl:192-216,575-641 ; src/problems_dirichlet.jl:60-90 ; examples/linear_static.jl:23-100)

All heavy arithmetic (element integration, scatter, K.u, CG, Newton-Krylov) runs in libjfem_b200.so on the GPU; there is
no CPU fallback.  What stays on the host is bookkeeping: node renumbering, load vectors, boundary-condition lists.
"""
from __future__ import annotations

import time as _time
from dataclasses import dataclass, field

import numpy as np

from . import _lib

# element topologies (src/topology/tetrahedra.jl:73-105, hexahedra.jl:14-18, surfaces for loads)
Tet4, Hex8, Tet10, Tri3, Tri6, Quad4, Poi1 = "Tet4", "Hex8", "Tet10", "Tri3", "Tri6", "Quad4", "Poi1"
_NNPE = {Tet4: 4, Hex8: 8, Tet10: 10, Tri3: 3, Tri6: 6, Quad4: 4, Poi1: 1}
_VOLUME = (Tet4, Hex8, Tet10)
_SURFACE = (Tri3, Tri6, Quad4)


class Elasticity:
    """Problem properties of src/problems_elasticity.jl:36-46."""

    def __init__(self):
        self.formulation = "continuum"
        self.finite_strain = False
        self.geometric_stiffness = False
        self.store_fields = []


class Dirichlet:
    pass


class Linear:
    pass


class Nonlinear:
    """Solver knobs of src/solvers.jl:539-551 (max_iterations, convergence_tolerance on ||du||)."""
    max_iterations = 10
    convergence_tolerance = 5.0e-5


@dataclass
class LinearElastic:          # src/materials/linear_elastic.jl:51-56
    E: float
    nu: float


@dataclass
class NeoHookean:             # src/materials/neo_hookean.jl:87-100 (E_mod, nu keywords)
    E_mod: float = 3e6
    nu: float = 0.45


@dataclass
class PerfectPlasticity:      # src/materials/perfect_plasticity.jl:166-181
    E: float
    nu: float
    sigma_y: float
    H: float = 0.0


class GPU:                    # src/backend/abstract.jl:61-65
    def __init__(self, device: int = 0):
        self.device = device


class CPU:
    def __init__(self, nthreads: int = 1):
        self.nthreads = nthreads


class Element:
    """Immutable element (src/elements/elements.jl:70-76): topology, connectivity (node ids, 1-based), named fields."""

    def __init__(self, topology, connectivity, fields=None, id=-1):
        if topology not in _NNPE:
            raise ValueError(f"unknown element topology {topology}")
        if len(connectivity) != _NNPE[topology]:
            raise ValueError(f"{topology} needs {_NNPE[topology]} nodes")
        self.topology, self.connectivity, self.id = topology, tuple(int(c) for c in connectivity), id
        self.fields = dict(fields or {})

    def __call__(self, name):
        return self.fields[name]


def update_(elements, name, value):
    """update!(elements, "youngs modulus", 208.0e3)  (time-independent fields only)"""
    for el in (elements if isinstance(elements, (list, tuple)) else [elements]):
        el.fields[name] = value


# ---- mesh -> elements helpers of the reference's preprocessing layer (src/preprocess.jl:83-89,210-262, src/io/abaqus_reader.jl:62-67)

def add_node_to_node_set_(mesh, set_name, *node_ids):
    """add_node_to_node_set!(mesh, :name, nids...)  (src/preprocess.jl:83-89); mesh = juliafem.jl_b200.mesh.Mesh"""
    cur = set(int(n) for n in mesh.node_sets.get(set_name, ()))
    cur.update(int(n) for n in node_ids)
    mesh.node_sets[set_name] = np.array(sorted(cur), dtype=np.int64)


def create_elements(mesh, *element_sets):
    """create_elements(mesh, "OTHER", ...)  (src/preprocess.jl:210-262): volume elements of the named element sets (all
    elements without a name; a name the mesh has no set for, like MED's family-0 "OTHER" of a mesh without groups, means all),
    each with its "geometry" field (3 x nnpe) and its element id."""
    topo = {4: Tet4, 8: Hex8, 10: Tet10}[int(mesh.elem_type)]
    ids = None
    for name in element_sets:
        if name in mesh.elem_sets:
            sel = np.asarray(mesh.elem_sets[name], dtype=np.int64)
            ids = sel if ids is None else np.union1d(ids, sel)
    if ids is None:
        ids = np.arange(1, mesh.n_elems + 1, dtype=np.int64)
    return [Element(topo, mesh.conn[e - 1], fields={"geometry": mesh.coords[mesh.conn[e - 1] - 1].T}, id=int(e)) for e in ids]


def create_nodal_elements(mesh, node_set_name):
    """create_nodal_elements(mesh, "mid_fixed")  (src/io/abaqus_reader.jl:62-67): one Poi1 element per node of the set"""
    return [Element(Poi1, [int(n)], fields={"geometry": mesh.coords[int(n) - 1][:, None]}) for n in mesh.node_sets[node_set_name]]


def _nodes_by_rows(X, nn):
    """The reference stores element geometry as a 3 x nnpe matrix (one column per node, ext/JuliaFEMCUDAExt.jl:117);
    an (nnpe, 3) array is accepted too when unambiguous."""
    X = np.asarray(X, dtype=np.float64)
    if X.shape == (3, nn):
        return X.T
    if X.shape == (nn, 3):
        return X
    raise ValueError(f"geometry must be 3 x {nn}")


def _get(el, *names, default=None):
    for n in names:
        if n in el.fields:
            return el.fields[n]
    return default


# ------------------------------------------------------------------------------------------------ seam 1

@dataclass
class DirichletBC:            # src/physics.jl:42-52
    node_ids: list = field(default_factory=list)
    components: list = field(default_factory=list)
    values: list = field(default_factory=list)


@dataclass
class NeumannBC:              # src/physics.jl:54-64
    surface_elements: list = field(default_factory=list)
    traction: list = field(default_factory=list)


@dataclass
class ElasticitySolution:     # src/backend/abstract.jl:138-145
    u: np.ndarray
    newton_iterations: int
    cg_iterations: int
    residual: float
    solve_time: float
    history: list
    converged: bool = True


class Physics:
    """Physics(Elasticity, "name", 3)  (src/physics.jl:128-156)"""

    def __init__(self, kind=Elasticity, name="physics", dimension=3):
        if dimension != 3:
            raise ValueError("only 3D continuum elasticity is on the accelerated path")
        self.name, self.dimension = name, dimension
        self.properties = kind() if isinstance(kind, type) else kind
        self.body_elements: list = []
        self.bc_dirichlet = DirichletBC()
        self.bc_neumann = NeumannBC()
        self.material = None


def add_elements_(target, elements):
    if isinstance(target, Physics):
        target.body_elements.extend(elements)
    else:
        target.elements.extend(elements)


def add_dirichlet_(physics: Physics, node_ids, components, value=0.0):
    """add_dirichlet!(physics, [node], [1,2,3], 0.0)  (src/physics.jl:222)"""
    physics.bc_dirichlet.node_ids.append(list(node_ids))
    physics.bc_dirichlet.components.append(list(components))
    physics.bc_dirichlet.values.append(float(value))


def add_neumann_(physics: Physics, surface_element: Element, traction):
    physics.bc_neumann.surface_elements.append(surface_element)
    physics.bc_neumann.traction.append(np.asarray(traction, dtype=np.float64))


def _material_of(elements, properties, explicit=None):
    """(kind, params) for jfem_set_material from the element fields / an explicit material struct."""
    if isinstance(explicit, NeoHookean):
        return _lib.MAT_NEO_HOOKEAN, (explicit.E_mod, explicit.nu)
    if isinstance(explicit, PerfectPlasticity):
        return _lib.MAT_PERFECT_PLASTICITY, (explicit.E, explicit.nu, explicit.sigma_y, explicit.H)
    if isinstance(explicit, LinearElastic):
        return _lib.MAT_LINEAR_ELASTIC, (explicit.E, explicit.nu)
    E = {_get(el, "youngs modulus", "youngs_modulus") for el in elements}
    nu = {_get(el, "poissons ratio", "poissons_ratio") for el in elements}
    if None in E or None in nu:
        raise KeyError("elements need \"youngs modulus\" and \"poissons ratio\" fields")   # reference: KeyError from element(...)
    # props.finite_strain of the classic path = Hooke's D on the Green-Lagrange strain (St. Venant-Kirchhoff,
    # src/problems_elasticity.jl:255-332); Neo-Hookean only when a NeoHookean material is given explicitly
    kind = _lib.MAT_STVK if getattr(properties, "finite_strain", False) else _lib.MAT_LINEAR_ELASTIC
    if len(E) != 1 or len(nu) != 1:      # per-element arrays, like E_vec / nu_vec of ext/JuliaFEMCUDAExt.jl:135-141
        return kind, np.array([[_get(el, "youngs modulus", "youngs_modulus"), _get(el, "poissons ratio", "poissons_ratio")] for el in elements])
    return kind, (E.pop(), nu.pop())


class ElasticityDataGPU:
    """Device-side data of one problem (ext/JuliaFEMCUDAExt.jl:34-60): owns the jfem handle."""

    def __init__(self, elements, dirichlet: DirichletBC, neumann: NeumannBC, properties, material=None, device=0, options=None):
        vol = [el for el in elements if el.topology in _VOLUME]
        if not vol:
            raise ValueError("no volume elements")
        bad = [el for el in elements if el.topology not in _VOLUME and el.topology not in _SURFACE]
        if bad:
            # same refusal as assemble! for e.g. Seg3 in a 3D problem (src/problems_elasticity.jl:510-518)
            raise ValueError(f"unsupported element type {bad[0].topology} in a 3D continuum problem")
        self.surface_elements = [el for el in elements if el.topology in _SURFACE]   # Elasticity3DSurfaceElements (pe:454-458)
        kinds = {el.topology for el in vol}
        if len(kinds) != 1:
            raise NotImplementedError("one element type per problem on the accelerated path")
        self.topology = kinds.pop()
        nnpe = _NNPE[self.topology]
        # node renumbering: sorted unique ids -> 1..n (ext:100-108)
        conn = np.array([el.connectivity for el in vol], dtype=np.int64)
        self.node_ids = np.unique(conn)
        remap = {int(n): i + 1 for i, n in enumerate(self.node_ids)}
        self.conn = np.vectorize(remap.__getitem__)(conn).astype(np.int32)
        coords = np.zeros((self.node_ids.size, 3))
        for el, c in zip(vol, self.conn):
            X = np.asarray(_get(el, "geometry"), dtype=np.float64)
            X = _nodes_by_rows(X, nnpe)
            coords[c - 1] = X
        self.coords = coords
        self.n_nodes, self.n_dofs = coords.shape[0], 3 * coords.shape[0]
        self.handle = _lib.Handle(nnpe, coords, self.conn, device=device)
        for k, v in (options or {}).items():
            self.handle.set_option(k, v)
        self.handle.set_material(*_material_of(vol, properties, material))
        if getattr(properties, "geometric_stiffness", False):
            self.handle.set_option("geometric_stiffness", 1)
        # Dirichlet: is_fixed / prescribed per dof (ext:144-158)
        dofs, vals = [], []
        for nodes, comps, val in zip(dirichlet.node_ids, dirichlet.components, dirichlet.values):
            for n in nodes:
                for c in comps:
                    dofs.append(3 * (remap[int(n)] - 1) + int(c))
                    vals.append(val)
        self.fixed_dofs = np.array(dofs, dtype=np.int64)
        self.prescribed = np.zeros(self.n_dofs)
        if dofs:
            self.prescribed[self.fixed_dofs - 1] = vals
        self.handle.set_dirichlet(self.fixed_dofs, np.array(vals))
        # external loads, integrated on the device (jfem_surface_load / jfem_body_load):
        #  * Neumann surfaces of seam 1: consistent traction; for Tri3 (GLTRI1) this IS the lumped area/3 * t of
        #    apply_surface_traction_kernel! (ext:368-416)
        #  * surface elements of the classic problem with "displacement traction force [i]" / "surface pressure" (pe:454-502)
        #  * "displacement load [i]" body loads of the volume elements (pe:412-426)
        self.remap = remap
        self.f_ext = np.zeros(self.n_dofs)
        groups = {}
        for surf, t in zip(neumann.surface_elements, neumann.traction):
            groups.setdefault(surf.topology, []).append((surf, np.asarray(t, dtype=np.float64), None))
        for surf in self.surface_elements:
            t = _get(surf, "displacement traction force")
            t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).copy()
            for i in range(3):
                ti = _get(surf, f"displacement traction force {i + 1}")
                if ti is not None:
                    t[i] += float(ti)
            groups.setdefault(surf.topology, []).append((surf, t, _get(surf, "surface pressure")))
        for topo, items in groups.items():
            faces = np.array([[remap[int(n)] for n in sf.connectivity] for sf, _, _ in items], dtype=np.int32)
            trac = np.array([t for _, t, _ in items])
            pres = np.array([0.0 if p is None else float(p) for _, _, p in items])
            self.f_ext = self.handle.surface_load(_NNPE[topo], faces, traction=trac if np.any(trac) else None,
                                                  pressure=pres if np.any(pres) else None, f=self.f_ext)
        b = np.array([[float(_get(el, f"displacement load {i + 1}", default=0.0)) for i in range(3)] for el in vol])
        for k, el in enumerate(vol):
            bv = _get(el, "displacement load")
            if bv is not None:
                b[k] += np.asarray(bv, dtype=np.float64)
        if np.any(b):
            self.f_ext = self.handle.body_load(b[0] if np.all(b == b[0]) else b, f=self.f_ext)

    def close(self):
        self.handle.close()


def initialize_backend(backend, physics: Physics, time=0.0, options=None):
    """initialize_backend(::GPU, physics, time)  (ext/JuliaFEMCUDAExt.jl:867-869)"""
    if not isinstance(backend, GPU):
        raise NotImplementedError("this package provides the GPU backend only (no CPU fallback)")
    return ElasticityDataGPU(physics.body_elements, physics.bc_dirichlet, physics.bc_neumann, physics.properties,
                             material=physics.material, device=backend.device, options=options)


def solve_backend_(data: ElasticityDataGPU, physics: Physics = None, tol=1e-6, max_iter=1000, newton_tol=1e-6, max_newton=20,
                   max_cg_per_newton=50, forcing_power=0.5, forcing_max=0.9):
    """solve_backend!(data, physics; ...) -> (u, newton_iters, cg_iters, residual, history)   (ext:886-916).
    Linear elastic: one projected CG solve of K u = f - K u_B with the reference's absolute stop sqrt(r.r) < tol
    (ext:531-577).  Nonlinear materials: inexact Newton-Krylov with the Eisenstat-Walker forcing of ext:819-820."""
    h = data.handle
    nonlinear = physics is not None and (getattr(physics.properties, "finite_strain", False)
                                          or isinstance(physics.material, (NeoHookean, PerfectPlasticity)))
    if nonlinear:
        u0 = data.prescribed.copy()
        u, nit, cgit, res, hist = h.newton_krylov(data.f_ext, u0, newton_tol=newton_tol, max_newton=max_newton,
                                                  max_cg_per_newton=max_cg_per_newton, forcing_power=forcing_power, forcing_max=forcing_max)
        data.converged = bool(res < newton_tol)
        if not data.converged:      # the reference prints "Newton did not converge" and returns (ext:850-853)
            import warnings
            warnings.warn(f"Newton-Krylov did not converge in {nit} iterations: ||R|| = {res:.3e} >= {newton_tol:.3e}")
        return u, nit, cgit, res, hist
    b = data.f_ext
    if np.any(data.prescribed):
        b = b - h.matvec(data.prescribed)          # lifting: f_I - K_IB u_B   (src/solvers.jl:205-210)
    x, it, res = h.cg(b, tol=tol, relative=False, max_iter=max_iter)
    data.converged = bool(res < tol)
    if not data.converged:          # ext:630 "CG did not converge"
        import warnings
        warnings.warn(f"CG did not converge in {it} iterations: sqrt(r.r) = {res:.3e} >= {tol:.3e}")
    u = x + data.prescribed
    rn = float(np.sqrt(np.sum(np.delete(b, data.fixed_dofs - 1) ** 2))) if data.fixed_dofs.size else float(np.linalg.norm(b))
    return u, 1, it, res, [(it, rn, min(forcing_max, rn ** forcing_power) if rn > 0 else 0.0)]


def solve_(physics: Physics, backend=None, time=0.0, tol=1e-6, max_iter=1000, newton_tol=1e-6, max_newton=20, max_cg_per_newton=50):
    """solve!(physics; backend=GPU(), time, tol, max_iter, newton_tol, max_newton, max_cg_per_newton) -> ElasticitySolution
    (src/backend/abstract.jl:181-217)"""
    backend = GPU() if backend is None else backend
    t0 = _time.perf_counter()
    data = initialize_backend(backend, physics, time)
    try:
        u, nit, cgit, res, hist = solve_backend_(data, physics, tol=tol, max_iter=max_iter, newton_tol=newton_tol,
                                                 max_newton=max_newton, max_cg_per_newton=max_cg_per_newton)
    finally:
        data.close()
    return ElasticitySolution(u, nit, cgit, res, _time.perf_counter() - t0, hist, getattr(data, "converged", True))


# ------------------------------------------------------------------------------------------------ seam 2

class SparseMatrixCSR:
    """What assemble_ leaves in problem.assembly.K: the reference stores COO triplets and converts with sparse()
    (src/sparse/sparse.jl:53-55); here the already-summed CSR (== CSC of the symmetric pattern) comes back from the device."""

    def __init__(self, rowptr, colind, vals, n):
        self.rowptr, self.colind, self.vals, self.n = rowptr, colind, vals, n

    def to_coo(self):
        rows = np.repeat(np.arange(1, self.n + 1), np.diff(self.rowptr))
        return rows, self.colind.copy(), self.vals.copy()       # I, J, V (1-based like SparseMatrixCOO)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals, self.colind - 1, self.rowptr - 1), shape=(self.n, self.n))


class Assembly:               # src/assembly/problems.jl:14-40 (fields used on this path)
    def __init__(self):
        self.K = None
        self.f = None
        self.u = None
        self.la = None


class Problem:
    """Problem(Elasticity, "name", 3) / Problem(Dirichlet, "fixed", 3, "displacement")  (src/assembly/problems.jl:95-105)"""

    def __init__(self, kind, name, dimension, parent_field_name=None):
        self.name, self.dimension, self.parent_field_name = name, dimension, parent_field_name
        self.properties = kind() if isinstance(kind, type) else kind
        self.elements: list = []
        self.assembly = Assembly()
        self.material = None
        self.postprocess_fields: list = []
        self.fields: dict = {}
        self._data = None

    def postprocess(self, name):
        """postprocess!(problem, time, Val{:strain|:stress})  (src/problems_elasticity.jl:583-594): least-squares nodal fit
        of the Gauss-point field on the device; returns (and keeps in problem.fields[name]) {node id: 6-vector}."""
        if name not in ("stress", "strain"):
            raise KeyError(name)
        if self._data is None or self.assembly.u is None:
            raise RuntimeError("postprocess needs a solved problem")
        d = self._data
        x = d.handle.nodal_recover(self.assembly.u, _lib.FIELD_STRESS if name == "stress" else _lib.FIELD_STRAIN)
        self.fields[name] = {int(n): x[i] for i, n in enumerate(d.node_ids)}
        return self.fields[name]


def apply_dirichlet_bc_(problem: "Problem", rhs=None, scale=1e10):
    """apply_dirichlet_bc!(assembly, fixed_dofs, prescribed) of the CPU backend (src/element_assembly_structures.jl:237-252):
    penalty = scale * max|K| added to the diagonal of every fixed dof of the ASSEMBLED K (on the device),
    rhs[dof] = penalty * prescribed.  Returns (penalty, rhs)."""
    d = problem._data
    if d is None or problem.assembly.K is None:
        raise RuntimeError("apply_dirichlet_bc_ needs an assembled problem (assemble_)")
    pen, r = d.handle.csr_penalty_bc(problem.assembly.f if rhs is None else rhs, scale)
    return pen, r


def _dirichlet_from(problems):
    bc = DirichletBC()
    for p in problems:
        if not isinstance(p.properties, Dirichlet):
            continue
        for el in p.elements:                       # nodal collocation (src/problems_dirichlet.jl:60-90)
            for c in (1, 2, 3):
                key = f"{p.parent_field_name or 'displacement'} {c}"
                if key in el.fields:
                    bc.node_ids.append(list(el.connectivity))
                    bc.components.append([c])
                    bc.values.append(float(el.fields[key]))
    return bc


def _ensure_data(problem: Problem, boundary=(), device=0):
    if problem._data is None:
        problem._data = ElasticityDataGPU(problem.elements, _dirichlet_from(boundary), NeumannBC(), problem.properties,
                                          material=problem.material, device=device)
    return problem._data


def assemble_(problem: Problem, time=0.0, u=None, symmetrise=False, device=0, boundary=()):
    """assemble!(problem, time)  (src/assembly/assembly.jl:31): leaves K (assembled on the GPU: element integration +
    coloured scatter into the reference's sparsity pattern) and f = f_ext - f_int in problem.assembly."""
    if isinstance(problem.properties, Dirichlet):
        return problem                      # boundary problems only contribute their dof lists (handled by the solver)
    d = _ensure_data(problem, boundary, device=device)     # boundary: Dirichlet problems (only needed by apply_dirichlet_bc_)
    rowptr, colind = d.handle.csr_pattern()
    vals, fint = d.handle.assemble_csr(u, symmetrise=symmetrise, want_f=u is not None)
    problem.assembly.K = SparseMatrixCSR(rowptr, colind, vals, d.n_dofs)
    problem.assembly.f = d.f_ext - (fint if fint is not None else 0.0)
    return problem


class ConvergenceError(RuntimeError):
    """Nonlinear iteration did not converge (the reference throws from the Nonlinear solver, src/solvers.jl:617-619)."""


class Analysis:
    """Analysis(Linear, model, fixed)  (src/analysis.jl:7-13).  After run_: u, reactions, converged, residual,
    iterations (Newton), cg_iterations; analysis("displacement", time) returns {node id: vector} like the reference's
    field access (examples/linear_static.jl:131)."""

    def __init__(self, kind, *problems, name="analysis"):
        self.properties = kind() if isinstance(kind, type) else kind
        self.problems = list(problems)
        self.name = name
        self.u = None
        self.reactions = None
        self.iterations = 0
        self.cg_iterations = 0
        self.converged = False
        self.residual = float("nan")
        self.history = []
        self.results_writers = []           # src/analysis.jl:11

    def __call__(self, field_name, time=0.0):
        model = next(p for p in self.problems if isinstance(p.properties, Elasticity))
        if self.u is None:
            raise KeyError(f"{field_name}: analysis has not been run")
        if field_name == "displacement":
            return nodal_displacements(model, self.u)
        if field_name == "reaction force":
            return nodal_displacements(model, self.reactions)
        if field_name in ("stress", "strain"):
            return model.postprocess(field_name)
        raise KeyError(field_name)


def _free_norm(b, fixed_dofs):
    if len(fixed_dofs):
        b = b.copy()
        b[np.asarray(fixed_dofs) - 1] = 0.0
    return float(np.linalg.norm(b))


def run_(analysis: Analysis, tol=1e-8, relative=True, max_iter=100000, device=0, newton_tol=None, strict=True):
    """run!(analysis)  (src/solvers.jl:641 Linear, :575 Nonlinear).  The direct LDLt of solve!(...,Val{1})
    (src/solvers.jl:192-216) is replaced by projected CG on the device with the same elimination semantics:
    u_B = g, K_II u_I = f_I - K_IB u_B, and the reaction forces la = K u - f on the constrained dofs (:211-216 returns
    them as the Lagrange multipliers).  Default stop: ||r|| <= 1e-8 ||b|| (north_star).
    Convergence is checked: a CG solve that stops at max_iter warns (as ext/JuliaFEMCUDAExt.jl:630 does), a Newton
    iteration that does not reach its tolerance raises ConvergenceError like src/solvers.jl:617-619 (strict=False
    downgrades that to a warning).  The Newton tolerance is relative to ||f_ext|| of the free dofs (an absolute residual
    is meaningless across unit systems): newton_tol defaults to max(tol, 1e-10)."""
    import warnings
    field_problems = [p for p in analysis.problems if isinstance(p.properties, Elasticity)]
    boundary = [p for p in analysis.problems if isinstance(p.properties, Dirichlet)]
    if len(field_problems) != 1:
        raise NotImplementedError("exactly one elasticity problem per analysis on the accelerated path")
    model = field_problems[0]
    model._data = None
    d = _ensure_data(model, boundary, device=device)
    h = d.handle
    if isinstance(analysis.properties, Nonlinear) or model.properties.finite_strain or isinstance(model.material, (NeoHookean, PerfectPlasticity)):
        scale = _free_norm(d.f_ext - (h.matvec(d.prescribed) if np.any(d.prescribed) else 0.0), d.fixed_dofs)
        rel = max(tol, 1e-10) if newton_tol is None else newton_tol
        ntol = rel * (scale if scale > 0 else 1.0)
        max_newton = analysis.properties.max_iterations if isinstance(analysis.properties, Nonlinear) else 20   # src/solvers.jl:539-551
        u, nit, cgit, res, hist = h.newton_krylov(d.f_ext, d.prescribed.copy(), newton_tol=ntol, max_newton=max_newton,
                                                  max_cg_per_newton=max_iter, forcing_max=1e-3)
        analysis.iterations, analysis.cg_iterations, analysis.history = nit, cgit, hist
        analysis.converged = bool(res < ntol)
        if not analysis.converged:
            msg = f"nonlinear iteration did not converge in {nit} iterations: ||R|| = {res:.3e} > {ntol:.3e}"
            if strict:
                analysis.u, analysis.residual = u, res
                raise ConvergenceError(msg)
            warnings.warn(msg)
    else:
        b = d.f_ext - (h.matvec(d.prescribed) if np.any(d.prescribed) else 0.0)
        x, it, res = h.cg(b, tol=tol, relative=relative, max_iter=max_iter)
        u = x + d.prescribed
        thr = tol * _free_norm(b, d.fixed_dofs) if relative else tol
        analysis.iterations, analysis.cg_iterations = 1, it
        analysis.converged = bool(res <= thr if relative else res < thr)
        if not analysis.converged:
            warnings.warn(f"CG did not converge in {it} iterations: ||r|| = {res:.3e} > {thr:.3e}")
    analysis.residual = res
    analysis.u = u
    # reaction forces on the constrained dofs: la = f_int(u) - f_ext (src/solvers.jl:211-216), zero elsewhere
    analysis.reactions = h.reactions(u, d.f_ext)
    model.assembly.u = u
    model.assembly.la = analysis.reactions
    for name in model.postprocess_fields:          # push!(model.postprocess_fields, "stress")  (examples/linear_static.jl:96)
        model.postprocess(name)
    if analysis.results_writers:                   # run!(analysis) ends with write_results!(analysis) (src/solvers.jl:641-650)
        write_results_(analysis, 0.0)
    return analysis


def add_results_writer_(analysis: "Analysis", writer):
    """add_results_writer!(analysis, Xdmf("results"))  (src/analysis.jl:73-76)"""
    analysis.results_writers.append(writer)


def write_results_(analysis: "Analysis", time=0.0):
    """write_results!(analysis) (src/analysis.jl:91-105) -> update_xdmf!(xdmf, problem, time, fields) (src/io.jl:387-518): the
    mesh in the caller's node order plus "displacement" and every field named in problem.postprocess_fields, as nodal data."""
    if not analysis.results_writers:
        import warnings
        warnings.warn(f"No result writers attached to the analysis {analysis.name}; use add_results_writer_(analysis, Xdmf(\"results\"))")
        return
    from .xdmf import update_xdmf_
    for model in (p for p in analysis.problems if isinstance(p.properties, Elasticity)):
        d = model._data
        if d is None or analysis.u is None:
            raise RuntimeError("write_results_ needs a solved analysis")
        extra = {}
        for name in model.postprocess_fields:
            vals = model.fields.get(name) or model.postprocess(name)
            extra[name] = np.array([vals[int(n)] for n in d.node_ids])
        for w in analysis.results_writers:
            update_xdmf_(w, model.name, time, d.coords, d.conn, _NNPE[d.topology], u=analysis.u, extra=extra)


def nodal_displacements(problem_or_data, u):
    """{node id: (ux, uy, uz)} in the caller's original node ids."""
    d = problem_or_data._data if isinstance(problem_or_data, Problem) else problem_or_data
    return {int(n): u[3 * i:3 * i + 3] for i, n in enumerate(d.node_ids)}
