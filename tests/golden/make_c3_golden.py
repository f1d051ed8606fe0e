#!/usr/bin/env python
"""Evaluate the CPU oracle ONCE at the full size of the bench workload (BASELINE config 3: N = 60 000, D = 784, L = 3
ReLU, Student-t, seed 10, reference CLI defaults) and print / store the loss, so the CUDA path can be pinned at full N
(tests/test_gpu_parity.py::test_c3_full_size_matches_oracle_golden, bench.py `parity.full_n_rel_err`).

Same arithmetic as oracle.nngp_oracle.spr_loss (spax/models.py:93-98 -> spax/likelihoods.py:45-50 -> spax/utils.py:
160-183) written memory-lean and in blocks: the Gram matrix is scaled and factored in place (one 28.8 GB array instead
of three), and the Cholesky / forward substitution run LAPACK potrf / trsm block by block (right-looking, 8192-wide
panels) because SciPy's 32-bit LAPACK interface segfaults on a single matrix with more than 2^31 elements
(N >= 46 341).  `--check` compares the blocked path with the one-call path at a size where both work.
About 3-5 minutes on 16-32 host cores.  Usage:  python tests/golden/make_c3_golden.py [--rows N] [--out file.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import scipy.linalg as sla
from scipy.special import gammaln


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=60000)
    ap.add_argument("--features", type=int, default=784)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c3_golden.json"))
    ap.add_argument("--block", type=int, default=8192)
    ap.add_argument("--check", action="store_true", help="also run scipy's one-call path (N < 46 341 only) and compare")
    args = ap.parse_args()
    from oracle import nngp_oracle as orc
    from tests.synth import pixel_data, DEFAULT_HP as hp
    n, d = args.rows, args.features
    x, y, *_ = pixel_data(n, d, seed=10)
    t0 = time.perf_counter()
    cov = orc.nngp_gram(x, num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"],
                        last_w_std=hp["last_w_std"], fast=True)
    t_gram = time.perf_counter() - t0
    cov[np.diag_indices(n)] += hp["eps"]                                   # spax/models.py:96
    a, b = hp["alpha"], hp["beta"]
    cov *= b / a                                                           # spax/likelihoods.py:49
    df = 2 * a
    t = 0.5 * (df + n)
    ref = None
    if args.check:
        Lr = sla.cholesky(cov, lower=True, check_finite=False)
        zr = sla.solve_triangular(Lr, y, lower=True, check_finite=False)
        ref = (float(zr @ zr), float(np.log(np.diag(Lr)).sum()))
        del Lr
    t1 = time.perf_counter()
    L = orc.cholesky_blocked_inplace(cov, args.block)                      # spax/utils.py:179
    t_chol = time.perf_counter() - t1
    z = orc.forward_substitution_blocked(L, y, args.block)                 # spax/utils.py:180
    logp = float(-t * np.log(1.0 + (1.0 / df) * (z @ z)) - n / 2 * np.log(df * np.pi) + gammaln(t)
                 - gammaln(0.5 * df) - np.log(np.diag(L)).sum())          # spax/utils.py:181-183
    loss = -logp / n                                                       # spax/models.py:98
    res = {"n": n, "d": d, "seed": 10, "hp": hp, "oracle_loss": loss, "oracle_logp": logp,
           "oracle_quad": float(z @ z), "oracle_sum_log_diag": float(np.log(np.diag(L)).sum()),
           "seconds": {"gram": t_gram, "cholesky": t_chol, "total": time.perf_counter() - t0},
           "host_cpus": os.cpu_count()}
    if ref is not None:
        res["check_one_call_vs_blocked"] = {"quad_rel": abs(ref[0] - res["oracle_quad"]) / abs(ref[0]),
                                            "sum_log_diag_rel": abs(ref[1] - res["oracle_sum_log_diag"]) / abs(ref[1])}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
