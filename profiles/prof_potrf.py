"""One potrf of size N (argv[1]) for ncu launch lists / full captures (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smnngp_b200 as sm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.zeros((n, n), dtype=torch.float64, device="cuda")
a.diagonal().fill_(float(n)); a[:, 0] = 1.0; a[0, 0] = float(n)
sm.device.potrf_(a)
torch.cuda.synchronize()
print("info ok, L[5,5] =", a[5, 5].item())
