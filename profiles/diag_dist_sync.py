"""2+-GPU diagnostic: per-step time of DistributedLML.lml() when every step is followed by a host sync (the e2e
pattern) vs enqueued back to back, for both exchange modes."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smnngp_b200 as sm
from smnngp_b200.distributed import DistributedLML
from tests.synth import pixel_data, DEFAULT_HP

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
x, y, *_ = pixel_data(n, 784)
xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
xp, yp = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
hp = sm.make_hp(device=dev, **DEFAULT_HP)
for mode in sys.argv[2:] or ["peer", "nccl"]:
    job = DistributedLML(n, 784, sm.StackSpec(3, "relu", "mlp"), dev, exchange=mode)
    for _ in range(2):
        job.lml(xd, yd, hp)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    # (a) back to back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out, _ = job.lml(xd, yd, hp)
    e1.record(); torch.cuda.synchronize()
    a = e0.elapsed_time(e1) / 3
    dist.barrier(); torch.cuda.synchronize()
    # (b) sync after every step, device inputs
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        out, _ = job.lml(xd, yd, hp)
        v = float(out[1].item())
        ts.append((time.perf_counter() - t0) * 1e3)
    dist.barrier(); torch.cuda.synchronize()
    # (c) sync after every step, host inputs
    tc = []
    for _ in range(3):
        t0 = time.perf_counter()
        xg, yg = xp.to(dev, non_blocking=True), yp.to(dev, non_blocking=True)
        t1 = time.perf_counter()
        out, _ = job.lml(xg, yg, hp)
        t2 = time.perf_counter()
        v = float(out[1].item())
        tc.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t0) * 1e3))
    print(f"rank {rank} {mode}: back-to-back {a:.1f} ms/step | synced {['%.1f' % t for t in ts]} | "
          f"host inputs (copy-enqueue, lml-enqueue, total) {[tuple('%.1f' % u for u in t) for t in tc]} loss {v:.12f}", flush=True)
    dist.barrier()
    if job.px is not None:
        job.px.close()
    del job
    torch.cuda.empty_cache()
dist.destroy_process_group()
