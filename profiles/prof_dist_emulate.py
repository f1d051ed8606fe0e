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

REAL = "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1      # under torchrun: the real job
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
d = 784
if REAL:
    import torch.distributed as dist
    P, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
x, y, *_ = pixel_data(n, d)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
hp = sm.make_hp(device="cuda", **DEFAULT_HP)
job = DistributedLML(n, d, sm.StackSpec(3, "relu", "mlp"), torch.device("cuda", torch.cuda.current_device()),
                     emulate=None if REAL else (P, rank))
if REAL and rank != 0:
    sys.stdout = open(os.devnull, "w")
NIT = int(os.environ.get("PROF_ITERS", "4"))
for it in range(NIT):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    want_tl = it == NIT - 1 and os.environ.get("PROF_TIMELINE", "1") != "0"
    if want_tl and job.mg is not None:
        job.enable_timeline()                                # C driver (smnngp_lml_mg_f64): events recorded inside
    elif want_tl:
        job.timeline = []
        if os.environ.get("PROF_TIMELINE") == "side":      # no timing events between the main-stream kernels
            job._tl_filter = ("main_start", "diag", "bcast", "trsm", "gather", "reorder")
    e0.record()
    job.lml(xd, yd, hp)
    e1.record()
    torch.cuda.synchronize()
    print(f"{'real' if REAL else 'emulated'} {job.exchange} P={P} rank={rank} N={n}: {e0.elapsed_time(e1):.2f} ms", flush=True)
c_marks = job.timeline_read() if job.mg is not None else None
if job.timeline is None and not c_marks:
    if REAL:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
tl = {}
if c_marks:
    # times are relative to the first mark = the start of the evaluation (before the Gram stage)
    for p, label, ms in c_marks:
        tl.setdefault(p, {})[label] = ms
else:
    for p, label, ev in job.timeline:
        tl.setdefault(p, {})[label] = e0.elapsed_time(ev)
npan = max(tl) + 1
tot_a = tot_b = tot_wait = 0.0
prev_end = tl[0].get("main_start", 0.0)
print("panel  main_start  wait  update_a  update_b | chain(p+1): diag bcast trsm gather reorder (ms, durations)")
for p in range(npan):
    t = tl[p]
    if "update_a" not in t:
        if "main_start" in t and p % 6 == 0:
            c = tl.get(p + 1, {})
            print(f"{p:4d} main_start {t['main_start']:9.2f} | chain(p+1) marks at: " +
                  " ".join(f"{k}={c[k]:.2f}" for k in ("diag", "bcast", "trsm", "gather") if k in c), flush=True)
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

if REAL:
    dist.barrier()
    dist.destroy_process_group()
