import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smnngp_b200 as sm
from smnngp_b200.distributed import CudaBackend
be = CudaBackend("cuda")
rng = np.random.default_rng(0)
for w in (32, 64, 96, 128):
    b = rng.standard_normal((w, w + 8)); a0 = torch.from_numpy(b @ b.T / (w + 8) + 1e-3 * np.eye(w)).cuda()
    linv = torch.zeros(128 * 128, dtype=torch.float64, device="cuda"); ld = torch.zeros(1, dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    a = a0.clone()
    for _ in range(20): be.factor_diag(a, linv, ld, info, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): be.factor_diag(a, linv, ld, info, 0)
    e1.record(); e1.synchronize()
    print(f"potf2 w={w}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call", flush=True)

import ctypes as C
lib = sm._lib.load()
clk = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.smnngp_debug_potf2_clocks(C.c_void_p(clk.data_ptr()))
w = 128
b = rng.standard_normal((w, w + 8)); a = torch.from_numpy(b @ b.T / (w + 8) + 1e-3 * np.eye(w)).cuda()
for _ in range(3):
    be.factor_diag(a, linv, ld, info, 0); torch.cuda.synchronize()
t = clk.cpu().numpy()
lib.smnngp_debug_potf2_clocks(C.c_void_p(0))
# marks (chol.cu potf2_trtri_kernel): start, staged, diag0, then per block b: panel_b, update_b(+diag_{b+1}), then
# inverse assembled, written back
nb = 128 // 16
d = np.diff(t[: 3 + 2 * nb + 2])
names = ["stage", "diag0"] + sum([[f"panel{b}", f"update{b}+diag{b + 1}"] for b in range(nb)], []) + ["inv-assembly", "write-back"]
tot = d.sum()
print("potf2 phases (cycles):", {n: int(v) for n, v in zip(names, d)})
print("total cycles", int(tot), "panel sum", int(d[2::2][:nb].sum()), "update+diag sum", int(d[3::2][:nb].sum()))
