#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 2-5 (not the driver's bench; run by hand under gpurun).

    python bench_configs.py hex8 [n]         Hex8 cube n^3 nodes: matrix-free K.u (config 3 at n=321 is 99.2 M DOF)
    python bench_configs.py neohooke [cx]    Neo-Hookean cantilever cx x cx/4 x cx/4 cells: residual / tangent K(u).v / Newton-Krylov
    python bench_configs.py plastic [c]      J2-plastic Tet10 block c^3 cells: state update + assembled CSR tangent + SpMV
    python bench_configs.py assemble [cx]    linear-elastic Tet10: CSR pattern + coloured assembly + SpMV

Each prints one JSON line.  Times are CUDA-event times of the kernels on the library stream (torch current stream).
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from juliafem.jl_b200 import _lib, mesh  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def handle(m, kind, par):
    h = _lib.Handle(m.elem_type, m.coords, m.conn)
    h.set_material(kind, par)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    return h


def hex8(n=161):
    t0 = time.perf_counter()
    m = mesh.hex8_lattice(n, n, n, 1.0 / (n - 1))
    t_mesh = time.perf_counter() - t0
    h = handle(m, 0, (210e9, 0.3))
    h.set_dirichlet(mesh.clamp_dofs(m))
    x = torch.from_numpy(mesh.test_vector(m.n_dofs)).cuda()
    y = torch.empty_like(x)
    t0 = time.perf_counter()
    h.matvec(x, y); torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    ms, mn = timed(lambda: h.matvec(x, y, flags=_lib.PROJECT))
    i = h.info()
    # properties: symmetry and translation kernel at full size
    v = torch.from_numpy(mesh.test_vector(m.n_dofs, seed=5)).cuda()
    Kv = torch.empty_like(v); h.matvec(v, Kv); h.matvec(x, y); torch.cuda.synchronize()
    sym = abs(float(v @ y - x @ Kv)) / abs(float(v @ y))
    print(json.dumps({"config": "hex8", "nodes_per_dir": n, "dofs": m.n_dofs, "elements": m.n_elems, "ms": ms, "ms_min": mn,
                      "gdofs": m.n_dofs / ms / 1e6, "frac_hbm_roofline": 35.67 * m.n_dofs / (ms * 1e-3) / 1e9 / 6456.2,
                      "mesh_s": t_mesh, "setup_s": t_setup, "patches": int(i.n_patches), "device_GB": i.device_bytes / 1e9,
                      "symmetry_rel": sym}))


def neohooke(cx=96):
    m = mesh.tet10_kuhn(cx, cx // 4, cx // 4, 4.0, 1.0, 1.0)
    h = handle(m, 1, (3e6, 0.45))
    fixed = mesh.clamp_dofs(m)
    h.set_dirichlet(fixed)
    G = 0.05 * np.random.default_rng(1).standard_normal((3, 3))
    u = torch.from_numpy((m.coords @ G.T).ravel()).cuda()
    f = torch.empty_like(u)
    v = torch.from_numpy(mesh.test_vector(m.n_dofs, fixed)).cuda()
    h.internal_force(u, f); torch.cuda.synchronize()
    ms_r, _ = timed(lambda: h.internal_force(u, f))
    h.set_linearization(u)
    ms_t, _ = timed(lambda: h.matvec(v, f, flags=_lib.TANGENT | _lib.PROJECT))
    # Newton-Krylov on a tip-loaded beam (forces scaled for ~5 % deflection)
    top = np.nonzero(np.abs(m.coords[:, 2] - 1.0) < 1e-12)[0]
    fext = np.zeros(m.n_dofs); fext[3 * top + 2] = -3e3 / top.size
    fd = torch.from_numpy(fext).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    uu, nit, cgit, res, hist = h.newton_krylov(fd, newton_tol=1e-6, max_newton=30, max_cg_per_newton=20000, forcing_max=1e-3)
    torch.cuda.synchronize(); t_nk = time.perf_counter() - t0
    print(json.dumps({"config": "neohooke", "dofs": m.n_dofs, "elements": m.n_elems, "residual_ms": ms_r, "tangent_matvec_ms": ms_t,
                      "residual_gdofs": m.n_dofs / ms_r / 1e6, "tangent_gdofs": m.n_dofs / ms_t / 1e6,
                      "newton_iters": nit, "cg_iters": cgit, "final_residual": res, "newton_krylov_s": t_nk,
                      "max_deflection": float(uu.abs().max())}))


def plastic(c=40):
    m = mesh.tet10_kuhn(c, c, c, 1.0, 1.0, 1.0)
    par = (200e9, 0.3, 250e6, 1e9)
    h = handle(m, 2, par)
    G = np.array([[3e-3, 1e-3, 0], [0, -1e-3, 5e-4], [2e-4, 0, -8e-4]])
    u = torch.from_numpy((m.coords @ G.T).ravel() + 0.3 * mesh.test_vector(m.n_dofs)).cuda()
    f = torch.empty_like(u)
    h.internal_force(u, f); torch.cuda.synchronize()
    ms_state, _ = timed(lambda: h.internal_force(u, f))
    t0 = time.perf_counter()
    rp, ci = h.csr_pattern()
    t_pat = time.perf_counter() - t0
    ms_asm, _ = timed(lambda: h.assemble_csr(u), reps=3, warm=1)
    v = torch.from_numpy(mesh.test_vector(m.n_dofs, seed=2)).cuda()
    y = torch.empty_like(v)
    ms_spmv, _ = timed(lambda: h.spmv(v, y))
    h.set_linearization(u)
    y2 = torch.empty_like(v)
    ms_mf, _ = timed(lambda: h.matvec(v, y2, flags=_lib.TANGENT))
    torch.cuda.synchronize()
    diff = float((y - y2).abs().max() / y.abs().max())
    st = h.get_state(committed=False)
    print(json.dumps({"config": "plastic", "dofs": m.n_dofs, "elements": m.n_elems, "gauss_points": m.n_elems * 4, "nnz": int(ci.size),
                      "state_update_ms": ms_state, "gp_per_s": m.n_elems * 4 / ms_state * 1e3, "pattern_s": t_pat,
                      "assemble_ms": ms_asm, "elements_per_s": m.n_elems / ms_asm * 1e3, "spmv_ms": ms_spmv,
                      "spmv_gdofs": m.n_dofs / ms_spmv / 1e6, "spmv_GBs": 12.0 * ci.size / ms_spmv / 1e6,
                      "matfree_tangent_ms": ms_mf, "spmv_vs_matfree_rel": diff, "yielded_fraction": float(np.mean(st[:, :, 12] > 0))}))


def assemble(cx=64):
    m = mesh.tet10_kuhn(cx, cx // 4, cx // 4, 4.0, 1.0, 1.0)
    h = handle(m, 0, (210e9, 0.3))
    t0 = time.perf_counter()
    rp, ci = h.csr_pattern()
    t_pat = time.perf_counter() - t0
    ms_asm, _ = timed(lambda: h.assemble_csr(torch.zeros(m.n_dofs, dtype=torch.float64, device="cuda")), reps=3, warm=1)
    v = torch.from_numpy(mesh.test_vector(m.n_dofs)).cuda()
    y, y2 = torch.empty_like(v), torch.empty_like(v)
    ms_spmv, _ = timed(lambda: h.spmv(v, y))
    ms_mf, _ = timed(lambda: h.matvec(v, y2))
    torch.cuda.synchronize()
    print(json.dumps({"config": "assemble", "dofs": m.n_dofs, "elements": m.n_elems, "nnz": int(ci.size), "pattern_s": t_pat,
                      "assemble_ms": ms_asm, "elements_per_s": m.n_elems / ms_asm * 1e3, "spmv_ms": ms_spmv,
                      "spmv_gdofs": m.n_dofs / ms_spmv / 1e6, "spmv_GBs": 12.0 * ci.size / ms_spmv / 1e6, "matfree_ms": ms_mf,
                      "spmv_vs_matfree_rel": float((y - y2).abs().max() / y.abs().max())}))


if __name__ == "__main__":
    which = sys.argv[1]
    arg = [int(sys.argv[2])] if len(sys.argv) > 2 else []
    {"hex8": hex8, "neohooke": neohooke, "plastic": plastic, "assemble": assemble}[which](*arg)
