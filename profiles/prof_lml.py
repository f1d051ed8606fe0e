"""One fused LML of size N, D (argv) for ncu launch lists (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smnngp_b200 as sm
from tests.synth import regression_data, pixel_data, DEFAULT_HP as HP
n = int(sys.argv[1]); d = int(sys.argv[2])
x, y, *_ = pixel_data(n, d) if d > 100 else regression_data(n, d)
out, info = sm.device.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), spec=sm.StackSpec(3, "relu", "mlp"), hp=sm.make_hp(**HP))
torch.cuda.synchronize()
print("loss", out[1].item(), "info", info.item())
