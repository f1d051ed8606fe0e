"""BASELINE config 5 on N GPUs (run under torchrun, one rank per GPU): scale sweep N x depth x activation of the
distributed LML (Gram + block-row-cyclic Cholesky + Student-t LML), one JSON line per point on rank 0:
time (max over ranks, CUDA events), algorithmic TFLOP/s and the fraction of world x the FP64 tensor peak.

  torchrun --nproc-per-node 8 profiles/c5_sweep.py [--rows 20000,60000,100000,150000] [--depths 1,3,10] [--acts relu,erf]

erf uses b_std = 0.3.  A point whose factorisation reports a non-positive pivot at eps = 1e-6 is re-run
with eps = 1e-4 and says so."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import smnngp_b200 as sm
from smnngp_b200.distributed import DistributedLML
from tests.synth import pixel_data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="20000,60000,100000,150000")
    ap.add_argument("--depths", default="1,3,10")
    ap.add_argument("--acts", default="relu,erf")
    ap.add_argument("--features", type=int, default=784)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = sm._lib.load()
    peak = float(lib.smnngp_dmma_peak_tflops())
    d = args.features
    out = open(args.out, "a") if (args.out and rank == 0) else None
    for n in [int(v) for v in args.rows.split(",")]:
        x, y, *_ = pixel_data(n, d, seed=10)
        xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
        del x
        for act in args.acts.split(","):
            for L in [int(v) for v in args.depths.split(",")]:
                spec = sm.StackSpec(L, act, "mlp")
                job = DistributedLML(n, d, spec, dev)
                eps_used = None
                for eps in (1e-6, 1e-4):
                    hp = sm.make_hp(1.0, 0.3 if act == "erf" else 1e-8, 1.0, eps, 2.0, 2.0, device=dev)
                    o, info = job.lml(xd, yd, hp)                       # warm-up + PD check
                    torch.cuda.synchronize()
                    if int(info.item()) == 0:
                        eps_used = eps
                        break
                best = float("inf")
                for _ in range(args.reps):
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    o, info = job.lml(xd, yd, hp)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                    best = min(best, float(ms.item()))
                flops = n * (n + 1.0) * d + n ** 3 / 3.0 + float(n) * n
                line = {"config": "C5", "n_gpus": world, "N": n, "D": d, "L": L, "act": act, "eps": eps_used,
                        "ms": best, "tflops": flops / (best * 1e-3) * 1e-12, "recursion_evals": L * n * (n + 1) / 2,
                        "frac_of_peak": flops / (best * 1e-3) * 1e-12 / (world * peak), "fp64_peak_per_gpu": peak,
                        "loss": float(o[1].item()), "info": int(info.item()), "exchange": job.exchange,
                        "driver": "c" if job.mg is not None else "python"}
                if rank == 0:
                    print(json.dumps(line), flush=True)
                    if out:
                        out.write(json.dumps(line) + "\n")
                        out.flush()
                job.close()
                del job
                torch.cuda.empty_cache()
        del xd, yd
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
