"""Debug aid: per-phase clock64 stamps of block 0 / thread 0 of the patch kernel (option "debug_timing")."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from juliafem.jl_b200 import _lib, mesh
m = mesh.tet10_kuhn(88, 22, 22, 4.0, 1.0, 1.0)
h = _lib.Handle(10, m.coords, m.conn); h.set_material(0, (210e9, 0.3))
h.set_option("debug_timing", 1)
x = torch.from_numpy(mesh.test_vector(m.n_dofs)).cuda(); y = torch.empty_like(x)
h.set_stream(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device='cuda')
fn = _lib.lib().jfem_debug_timing
for k in range(3):
    flush.fill_(k); h.matvec(x, y); torch.cuda.synchronize()
    t = np.zeros(64, dtype=np.int64)
    _lib.check(fn(h._h, t.ctypes.data_as(C.c_void_p)))
    t = t.reshape(8, 8)
    if k < 2: continue
    i = h.info(); print("max_nodes", i.patch_max_nodes, "smem", i.smem_bytes, "blocks/SM", i.blocks_per_sm)
    for it in range(4):
        r = t[it]
        print(f" it {it} start {r[0]-t[0,0]} (a) {r[1]-r[0]} sync {r[2]-r[1]} ph1 {r[3]-r[2]} sync {r[4]-r[3]} mbar {max(r[7]-r[4],0)} gather-issue {r[5]-max(r[7],r[4])} ph2 {r[6]-r[5]} total {r[6]-r[0]}")
