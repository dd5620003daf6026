"""Dry runs of bench.py legs that are closures of main() and need a GPU in production: their PYTHON logic (argument plumbing,
JSON-serialisable output, the comparison with the reference-held value) is executed here with an oracle-backed stand-in for the
device handle.  Test infrastructure only -- nothing here is a fallback of the product."""
import json
import os
import sys
import textwrap
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _closure_source(start, end):
    src = open(os.path.join(ROOT, "bench.py")).read()
    a, b = src.index(start), src.index(end)
    assert a < b
    return textwrap.dedent(src[a:b])


def test_linear_static_leg_logic(oracle, jf):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    sys.path.insert(0, ROOT)
    import bench

    class FakeHandle:                      # the device solve replaced by the oracle's assembly + a direct solve
        def __init__(self, et, coords, conn, device=0):
            self.et, self.c, self.conn = et, coords, conn

        def set_material(self, kind, par):
            self.par = par

        def set_dirichlet(self, fixed):
            self.fixed = np.asarray(fixed)

        def set_stream(self, s):
            pass

        def body_load(self, b):
            return oracle.body_load(self.et, self.c, self.conn, b)

        def cg(self, f, tol, relative, max_iter):
            rp, ci, vals, _ = oracle.assemble_csr(self.et, self.c, self.conn, par=self.par, symmetrise=True)
            K = sp.csr_matrix((vals, ci, rp))
            free = np.setdiff1d(np.arange(K.shape[0]), self.fixed - 1)
            u = np.zeros(K.shape[0])
            u[free] = spla.splu(K[free][:, free].tocsc()).solve(f[free])
            return u, 123, 1e-9

        def close(self):
            pass

    fake_torch = types.SimpleNamespace(cuda=types.SimpleNamespace(current_stream=lambda: types.SimpleNamespace(cuda_stream=0),
                                                                  synchronize=lambda: None))
    ns = dict(np=np, os=os, time=time, mesh=jf.mesh, _lib=types.SimpleNamespace(Handle=FakeHandle, MAT_LINEAR_ELASTIC=0), torch=fake_torch,
              ROOT=ROOT, local_rank=0, do_cpu=True, host_threads=bench.host_threads)
    exec(_closure_source("    def linear_static_leg():", "    cg_out = asm_out = hex_out"), ns)
    out = ns["linear_static_leg"]()
    json.dumps(out)
    assert out["ok"] and out["rel_err"] < 1e-9 and out["reference_value"] == 2.4052929896922337
    assert out["gpu_vs_cpu_field_rel"] < 1e-12 and out["cg_iterations"] == 123


def test_cg_leg_logic_with_and_without_block_jacobi():
    class T:                               # minimal tensor stand-in
        def __init__(self, a):
            self.a = a

        def to(self, dev):
            return self

    fake_torch = types.SimpleNamespace(cuda=types.SimpleNamespace(synchronize=lambda: None), from_numpy=lambda a: T(a),
                                       zeros_like=lambda t: T(np.zeros_like(t.a)), empty_like=lambda t: T(np.empty_like(t.a)))
    seen = {}

    class H:
        def matvec(self, x, y, flags=0):
            return y

        def cg(self, b, x0=None, tol=0.0, relative=False, max_iter=0, flags=0):
            seen["flags"], seen["tol"], seen["relative"] = flags, tol, relative
            return x0, 4321, 1e-3

        def info(self):
            return types.SimpleNamespace(setup_seconds=1.5)

    pp = types.SimpleNamespace(handle=H(), local_nodes=np.arange(10), n_nodes_global=10)
    ns = dict(np=np, time=time, torch=fake_torch, dev=None, _lib=types.SimpleNamespace(PROJECT=1, JACOBI=8), barrier=lambda: None,
              allmax=lambda v: float(v))
    exec(_closure_source("    def cg_solve(pp_, label, flags=0):", "    def guarded(fn, pair=False):"), ns)
    plain = ns["cg_solve"](pp, "label")
    assert seen == {"flags": 0, "tol": 1e-8, "relative": True} and plain["preconditioner"].startswith("none") and plain["iterations"] == 4321
    jac = ns["cg_solve"](pp, "label", flags=8)
    assert seen["flags"] == 8 and "block-Jacobi" in jac["preconditioner"]
    json.dumps(plain), json.dumps(jac)
