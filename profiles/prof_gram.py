"""One fused Gram launch (N=20000, D=784, 3-layer ReLU, lower triangle) for an ncu capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smnngp_b200 as sm
from tests.synth import pixel_data
n, d = 20000, 784
x = torch.from_numpy(pixel_data(n, d)[0]).cuda()
out = torch.empty((n, n), dtype=torch.float64, device="cuda")
hp = sm.make_hp(1.0, 1e-8, 1.0, 1e-6, 2.0, 2.0)
for _ in range(2):
    sm.device.gram(x, spec=sm.StackSpec(3, "relu", "mlp"), hp=hp, lower_only=True, out=out)
torch.cuda.synchronize()
print("done")
