#!/usr/bin/env python
"""Small-N run of every kernel family of the hot path, meant to be executed under compute-sanitizer:

  compute-sanitizer --tool {memcheck,racecheck,synccheck} python profiles/sanitize_case.py

LML (TMA Gram + Cholesky with look-ahead + TMA trailing updates), value+gradient, predictive / test-NLL, the grid-search
pass over a cached base Gram, the draw stage, and the peer-store panel exchange with every peer aliased to the local
buffers (DistributedLML(emulate=...)).  Every result is checked against the oracle so a sanitizer-clean run is also a
correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import smnngp_b200 as sm
from oracle import nngp_oracle as orc
from tests.synth import regression_data, DEFAULT_HP as hp


def main():
    which = set(sys.argv[1:]) or {"lml", "grad", "predict", "exchange"}
    n, d, t = int(os.environ.get("SAN_N", "1100")), 16, 96
    x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
    spec = sm.StackSpec(3, "relu", "mlp")
    hpd = sm.make_hp(**hp)
    kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"])
    xd, yd, xtd, ytd = (torch.from_numpy(v).cuda() for v in (x, y, xt, yt))
    ref = orc.spr_loss(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
    if "lml" in which:
        sm.device.set_panel_width(256)                      # several outer panels -> look-ahead + TMA updates
        out, info = sm.device.lml(xd, yd, spec=spec, hp=hpd)
        assert int(info.item()) == 0 and abs(out[1].item() - ref) <= 1e-8 * abs(ref), (out[1].item(), ref)
        sm.device.set_panel_width(0)
        print("lml ok", out[1].item())
    if "grad" in which:
        out, grad, info = sm.device.lml_grad(xd, yd, spec=spec, hp=hpd)
        _, gref = orc.spr_loss_grad(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
        g = grad.cpu().numpy()
        assert np.all(np.abs(g - gref) <= 1e-6 * np.abs(gref) + 1e-11 * np.abs(gref).max()), (g, gref)
        print("grad ok")
    if "predict" in which:
        nll, mean, var, info = sm.device.test_nll(xd, yd, xtd, ytd, ym, ys, spec=spec, hp=hpd)
        ref_nll = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
        assert abs(float(nll.item()) - ref_nll) <= 1e-8 * abs(ref_nll)
        gs = sm.device.GridSearch(xd, yd, xtd, spec=spec)
        gs.point(hpd)
        print("predict ok")
    if "exchange" in which:
        from smnngp_b200.distributed import DistributedLML
        # one rank's schedule of a 2-rank job with both "peers" aliased to the local buffers: the kernels, stores,
        # fences and flags are the real ones, the numbers are not (the other rank's rows are missing)
        job = DistributedLML(n, d, spec, "cuda", block=256, emulate=(2, 0), exchange="peer")
        job.lml(xd, yd, hpd)
        torch.cuda.synchronize()
        job.close()
        one = DistributedLML(n, d, spec, "cuda", block=256)          # P = 1 stage path (checked)
        out, info = one.lml(xd, yd, hpd)
        assert int(info.item()) == 0 and abs(out[1].item() - ref) <= 1e-8 * abs(ref), (out[1].item(), ref)
        print("exchange ok")
    torch.cuda.synchronize()
    print("sanitize_case done")


if __name__ == "__main__":
    main()
