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
