import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smnngp_b200 as sm
from smnngp_b200.distributed import CudaBackend
lib = sm._lib.load(); be = CudaBackend("cuda")
torch.manual_seed(0)
def run(M, N, K, lower, variant, reps=3):
    lib.smnngp_set_tile_variant(variant)
    a = torch.randn(M, K, dtype=torch.float64, device="cuda")
    b = torch.randn(N, K, dtype=torch.float64, device="cuda")
    c0 = torch.randn(M, N, dtype=torch.float64, device="cuda")
    ref = c0 - a @ b.T
    if lower:
        mask = torch.arange(N, device="cuda")[None, :] <= torch.arange(M, device="cuda")[:, None]
        ref = torch.where(mask, ref, c0)
    out = []
    for _ in range(reps):
        c = c0.clone()
        be.update(a, b, c, lower, 0, 1, 0)
        torch.cuda.synchronize()
        d = (c - ref).abs()
        bad = (d > 1e-9).nonzero()
        tiles = sorted({(int(r) // 128, int(cc) // 64) for r, cc in bad[:200000].tolist()})
        out.append((float(d.max()), len(bad), tiles[:6], len(tiles)))
    return out
for (M, N, K, lower) in [(2745, 2744, 256, 1), (2245, 2244, 256, 1), (2745, 2752, 256, 0), (2816, 2752, 256, 0), (2816, 2816, 256, 0),
                         (4096, 4096, 128, 0), (4096, 4096, 64, 0), (4096, 4096, 16, 0), (4096, 4160, 16, 0), (8192, 8256, 512, 1)]:
    for v in (0, 2):
        print(M, N, K, lower, "variant", v, run(M, N, K, lower, v), flush=True)
