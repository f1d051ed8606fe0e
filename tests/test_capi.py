"""The C-ABI from a pure-C caller (tests/c/capi_lml.c): no Python, no torch in the process that computes.
CPU part: the program compiles as C11 against include/smnngp.h and links against libsmnngp.so.
GPU part: host-buffer, device-stream and (>= 2 GPUs) multi-GPU entry points against the oracle on the same data."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "capi_lml.c")


def _build(tmp):
    import smnngp_b200 as sm
    lib = sm._lib.build()
    exe = os.path.join(tmp, "capi_lml")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-Werror", SRC, "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(cuda, "include"), "-L", os.path.dirname(lib), "-lsmnngp",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe, os.path.dirname(lib), os.path.join(cuda, "lib64")


def test_c_caller_compiles_and_links(tmp_path):
    exe, *_ = _build(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_caller_matches_oracle(tmp_path):
    import torch
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    exe, libdir, cudalib = _build(str(tmp_path))
    n, d = 3000, 16
    x, y, *_ = regression_data(n, d)
    data = os.path.join(str(tmp_path), "data.bin")
    with open(data, "wb") as f:
        x.tofile(f)
        y.tofile(f)
    world = 2 if torch.cuda.device_count() >= 2 else 1
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + cudalib + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, data, str(n), str(d), str(world)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = orc.spr_loss(x, y, num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"],
                       last_w_std=hp["last_w_std"], eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"])
    vals = {}
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] in ("loss_host", "loss_dev"):
            vals[t[0]] = float(t[1])
            assert int(t[3]) == 0
        elif t[0] == "loss_mg" and t[1] != "skipped:":
            vals.setdefault("mg", []).append(float(t[4]))
            assert int(t[6]) == 0
    assert abs(vals["loss_host"] - ref) <= 1e-8 * abs(ref)
    assert vals["loss_dev"] == vals["loss_host"]
    if world > 1:
        assert len(vals["mg"]) == world
        for v in vals["mg"]:
            assert abs(v - ref) <= 1e-8 * abs(ref) and abs(v - vals["loss_dev"]) <= 1e-11 * abs(ref)
        assert len(set(vals["mg"])) == 1            # bit-identical on every rank
