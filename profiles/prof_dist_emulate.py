"""Single-GPU timing dry-run of ONE rank of a P-rank distributed LML (DistributedLML(emulate=(P, rank))): collectives
replaced by local copies, kernels keep their real shapes.  Prints the per-panel timeline of the main stream
(update_a / update_b / wait for the look-ahead chain) and of the side stream (diag / trsm / gather)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smnngp_b200 as sm
from smnngp_b200.distributed import DistributedLML
from tests.synth import pixel_data, DEFAULT_HP

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
d = 784
x, y, *_ = pixel_data(n, d)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
hp = sm.make_hp(**DEFAULT_HP)
job = DistributedLML(n, d, sm.StackSpec(3, "relu", "mlp"), "cuda", emulate=(P, rank))
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if it == 2:
        job.timeline = []
    e0.record()
    job.lml(xd, yd, hp)
    e1.record()
    torch.cuda.synchronize()
    print(f"emulated P={P} rank={rank} N={n}: {e0.elapsed_time(e1):.2f} ms", flush=True)
tl = {}
for p, label, ev in job.timeline:
    tl.setdefault(p, {})[label] = e0.elapsed_time(ev)
npan = max(tl) + 1
tot_a = tot_b = tot_wait = 0.0
prev_end = tl[0].get("main_start", 0.0)
print("panel  main_start  wait  update_a  update_b | chain(p+1): diag bcast trsm gather reorder (ms, durations)")
for p in range(npan):
    t = tl[p]
    if "update_a" not in t:
        continue
    ms = t["main_start"]
    wait = ms - prev_end
    ua = t["update_a"] - ms
    ub = t.get("update_b", t["update_a"]) - t["update_a"]
    prev_end = t.get("update_b", t["update_a"])
    tot_a += ua; tot_b += ub; tot_wait += max(wait, 0.0)
    c = tl.get(p + 1, {})
    ch = []
    last = t["update_a"]
    for k in ("diag", "bcast", "trsm", "gather", "reorder"):
        if k in c:
            ch.append(c[k] - last); last = c[k]
        else:
            ch.append(float("nan"))
    if p % 6 == 0 or p > npan - 8:
        print(f"{p:4d} {ms:10.2f} {wait:6.2f} {ua:8.3f} {ub:8.3f} | " + " ".join(f"{v:6.3f}" for v in ch), flush=True)
print(f"sum update_a {tot_a:.1f} ms, update_b {tot_b:.1f} ms, main-stream waits {tot_wait:.1f} ms, "
      f"gram+first panel {tl[0]['main_start']:.1f} ms")
