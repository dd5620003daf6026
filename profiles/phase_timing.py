"""Debug aid: clock64 stamps of block 0 of the warp-specialised patch kernel (option "debug_timing"):
compute warp 0 / helper warp 0, first 7 patches."""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from juliafem.jl_b200 import _lib, mesh
dims = [int(v) for v in os.environ.get('JFEM_DIMS', '88,22,22').split(',')]
m = mesh.tet10_kuhn(dims[0], dims[1], dims[2], 4.0, 1.0, 1.0)
h = _lib.Handle(10, m.coords, m.conn); h.set_material(0, (210e9, 0.3))
h.set_option("debug_timing", 1)
import os
h.set_option("debug_skip", int(os.environ.get("JFEM_SKIP", "0")))
print("debug_skip =", os.environ.get("JFEM_SKIP", "0"))
x = torch.from_numpy(mesh.test_vector(m.n_dofs)).cuda(); y = torch.empty_like(x)
h.set_stream(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device='cuda')
fn = _lib.lib().jfem_debug_timing
for k in range(3):
    flush.fill_(k); h.matvec(x, y); torch.cuda.synchronize()
    t = np.zeros(128, dtype=np.int64)
    _lib.check(fn(h._h, t.ctypes.data_as(C.c_void_p)))
    if k < 2: continue
    c, hp = t[:64].reshape(8, 8), t[64:].reshape(8, 8)
    base = c[0, 0]
    for it in range(7):
        r = c[it]
        print(f"compute it {it} start {r[0]-base:6d} wait x tile + part B {r[1]-r[0]:5d} wait stage_empty {r[2]-r[1]:5d} phase 1 {r[3]-r[2]:5d}")
    for it in range(7):
        r = hp[it]
        print(f"helper  it {it} start {r[0]-base:6d} wait {r[1]-r[0]:5d} gather issue {r[5]-r[1]:5d} phase 2 {r[2]-r[5]:5d} (setup {r[3]-r[5]:5d} rows {r[4]-r[3]:5d} [{r[6]} rows] stores {r[2]-r[4]:5d})")
