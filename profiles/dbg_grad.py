"""Diagnostic: gradient entry point vs oracle at small N, timing of value-only vs value+gradient at larger N."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smnngp_b200 as sm
from oracle import nngp_oracle as orc
from tests.synth import regression_data, pixel_data, DEFAULT_HP

NAMES = ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")


def run(x, y, hp, L, act, arch, kind):
    hpv = torch.tensor([hp[k] for k in NAMES], dtype=torch.float64, device="cuda")
    out, grad, info = sm.device.lml_grad(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         spec=sm.StackSpec(L, act, arch), hp=hpv, kind=kind)
    return out.cpu().numpy(), grad.cpu().numpy(), int(info.item())


if "check" in sys.argv:
    for (n, d, L, act, arch, kind, over) in [(10, 5, 3, "relu", "mlp", "student_t", dict(eps=1e-3)),
                                              (300, 8, 3, "relu", "mlp", "student_t", {}),
                                              (300, 13, 3, "relu", "mlp", "gauss", {}),
                                              (300, 8, 2, "erf", "mlp", "student_t", dict(b_std=0.3)),
                                              (300, 8, 2, "relu", "resnet", "student_t", dict(b_std=0.3)),
                                              (1500, 8, 3, "relu", "mlp", "student_t", {}),
                                              (2600, 8, 3, "relu", "mlp", "student_t", {})]:
        x, y, *_ = regression_data(n, d)
        hp = dict(DEFAULT_HP); hp.update(over)
        try:
            out, grad, info = run(x, y, hp, L, act, arch, kind)
            rl, ref = orc.spr_loss_grad(x, y, num_hiddens=L, act=act, arch=arch, w_std=hp["w_std"], b_std=hp["b_std"],
                                        last_w_std=hp["last_w_std"], eps=hp["eps"], kind=kind, a=hp["alpha"], b=hp["beta"])
            print(n, d, L, act, arch, kind, "info", info, "loss err", abs(out[1] - rl) / abs(rl),
                  "grad err", np.abs(grad - ref) / np.abs(ref).max(), flush=True)
            print("   grad", grad, "\n   ref ", ref, flush=True)
        except Exception as e:
            print(n, d, L, act, arch, kind, "EXC", repr(e), flush=True)

if "time" in sys.argv:
    for (n, d) in [(10000, 8), (20000, 784), (40000, 784)]:
        x, y, *_ = (regression_data(n, d) if d < 100 else pixel_data(n, d))
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        hp = dict(DEFAULT_HP)
        hpv = torch.tensor([hp[k] for k in NAMES], dtype=torch.float64, device="cuda")
        spec = sm.StackSpec(3, "relu", "mlp")
        for name, fn in (("lml", lambda: sm.device.lml(xd, yd, spec=spec, hp=hpv)),
                         ("lml_grad", lambda: sm.device.lml_grad(xd, yd, spec=spec, hp=hpv))):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            fl = n * (n + 1) * d + n ** 3 / 3 + n * n
            if name == "lml_grad":
                fl = 2 * n * (n + 1) * d + n ** 3
            print(f"{name} N={n} D={d}: {ms:9.3f} ms  {fl / ms / 1e9:7.2f} TFLOP/s (algorithmic)  out={r[0].cpu().numpy()[:2]}"
                  + (f" grad={r[1].cpu().numpy()}" if name == "lml_grad" else ""), flush=True)
        sm.device.release_workspaces()
