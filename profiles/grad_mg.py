"""Value + gradient of SPR.loss on N GPUs (torchrun): smnngp_lml_grad_mg_f64 through distributed.DistributedGrad, timed
(max over ranks, CUDA events), compared on rank 0 with the single-GPU fused call (smnngp_lml_grad_f64) on the same inputs
when that fits.  One JSON line on rank 0.   torchrun --nproc-per-node P profiles/grad_mg.py [--rows N] [--features D]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import smnngp_b200 as sm
from smnngp_b200.distributed import DistributedGrad
from tests.synth import pixel_data, DEFAULT_HP


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=30000)
    ap.add_argument("--features", type=int, default=784)
    ap.add_argument("--single", type=int, default=1, help="also run the single-GPU call on rank 0 (needs 16 N^2 bytes)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d = args.rows, args.features
    x, y, *_ = pixel_data(n, d, seed=10)
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    spec = sm.StackSpec(3, "relu", "mlp")
    hp = sm.make_hp(device=dev, **DEFAULT_HP)
    job = DistributedGrad(n, d, spec, dev, emulate=(1, 0) if world == 1 else None)
    out, grad, info = job.lml_grad(xd, yd, hp)
    best = float("inf")
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, grad, info = job.lml_grad(xd, yd, hp)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = min(best, float(ms.item()))
    g_mg = grad.cpu().numpy()
    loss = float(out[1].item())
    job.close()
    del job
    torch.cuda.empty_cache()
    if rank == 0:
        flops = float(n) ** 3 + 2.0 * n * (n + 1.0) * d
        line = {"what": "value + gradient (smnngp_lml_grad_mg_f64)", "n_gpus": world, "N": n, "D": d, "ms": best,
                "tflops": flops / (best * 1e-3) * 1e-12, "loss": loss, "info": int(info.item()),
                "dloss_dhp": [float(v) for v in g_mg]}
        if args.single:
            sm.device.lml_grad(xd, yd, spec=spec, hp=hp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o1, g1, _ = sm.device.lml_grad(xd, yd, spec=spec, hp=hp)
            e1.record()
            torch.cuda.synchronize()
            g1 = g1.cpu().numpy()
            line["single_gpu_ms"] = e0.elapsed_time(e1)
            line["max_rel_diff_vs_1gpu"] = float(max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(g_mg, g1)))
            line["loss_rel_diff_vs_1gpu"] = abs(loss - float(o1[1].item())) / abs(float(o1[1].item()))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
