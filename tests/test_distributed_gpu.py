"""Multi-GPU driver on real devices: P = 1 always; P = 2 (NCCL) when the box has two GPUs."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref(n, d, kind):
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    x, y, *_ = regression_data(n, d)
    return orc.spr_loss(x, y, num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"],
                        last_w_std=hp["last_w_std"], eps=hp["eps"], kind=kind, a=hp["alpha"], b=hp["beta"])


@pytest.mark.parametrize("n,db", [(700, 128), (1024, 256), (2500, 512), (130, 128)])
def test_single_rank_stage_path_matches_oracle_and_fused(n, db):
    import torch
    import smnngp_b200 as sm
    from smnngp_b200.distributed import DistributedLML
    from tests.synth import regression_data, DEFAULT_HP as hp
    d = 8
    x, y, *_ = regression_data(n, d)
    spec = sm.StackSpec(3, "relu", "mlp")
    hpd = sm.make_hp(**hp)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    solver = DistributedLML(n, d, spec, "cuda", block=db)
    out, info = solver.lml(xd, yd, hpd, kind="student_t")
    ref = _ref(n, d, "student_t")
    assert int(info.item()) == 0
    assert abs(out[1].item() - ref) <= 1e-8 * abs(ref)
    fused, _ = sm.device.lml(xd, yd, spec=spec, hp=hpd)
    assert abs(out[1].item() - fused[1].item()) <= 1e-12 * abs(ref)


def _worker(rank, world, port, n, d, db, q, exchange="auto"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import smnngp_b200 as sm
        from smnngp_b200.distributed import DistributedLML
        from tests.synth import regression_data, DEFAULT_HP as hp
        x, y, *_ = regression_data(n, d)
        dev = torch.device("cuda", rank)
        solver = DistributedLML(n, d, sm.StackSpec(3, "relu", "mlp"), dev, block=db, exchange=exchange)
        assert solver.exchange == ("peer" if exchange == "auto" else exchange)
        xd, yd, hpd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), sm.make_hp(device=dev, **hp)
        out, info = solver.lml(xd, yd, hpd)
        out2, _ = solver.lml(xd, yd, hpd)                    # second evaluation: buffers / sequence numbers reused
        assert out2.cpu().tolist() == out.cpu().tolist()
        q.put((rank, out.cpu().tolist(), int(info.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("n,db", [(1500, 128), (3000, 256), (2900, 512), (1501, 128)])
def test_two_rank_matches_oracle(n, db, exchange):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + n % 100 + (50 if exchange == "peer" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 8, db, q, exchange)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _ref(n, 8, "student_t")
    for rank, out, info in res:
        assert info == 0 and abs(out[1] - ref) <= 1e-8 * abs(ref)


@pytest.mark.parametrize("layout", ["auto", "snake_end", "snake", "cyclic"])
@pytest.mark.parametrize("n,db", [(2900, 128), (3000, 256)])
def test_two_rank_block_layouts(n, db, layout, monkeypatch):
    """the block -> rank maps of the C driver (csrc/multigpu.cu: plain cyclic, snake, end-aligned snake, auto = default):
    same loss on both ranks, equal to the oracle"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("SMNNGP_MG_LAYOUT", layout)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + n % 100 + {"snake_end": 0, "snake": 7, "cyclic": 14, "auto": 21}[layout]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 8, db, q, "peer")) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _ref(n, 8, "student_t")
    for rank, out, info in res:
        assert info == 0 and abs(out[1] - ref) <= 1e-8 * abs(ref)
    assert res[0][1] == res[1][1]


def test_emulated_rank_schedule_runs_on_one_gpu():
    """DistributedLML(emulate=(P, rank)): the timing dry-run of one rank of a P-rank job (peer-store exchange with
    every peer aliased to the local buffers) must run to completion and record a timeline; its numbers are not
    meaningful, only its launches are."""
    import torch
    import smnngp_b200 as sm
    from smnngp_b200.distributed import DistributedLML
    from tests.synth import regression_data, DEFAULT_HP as hp
    n, d = 3000, 8
    x, y, *_ = regression_data(n, d)
    job = DistributedLML(n, d, sm.StackSpec(3, "relu", "mlp"), "cuda", block=256, emulate=(4, 1))
    assert job.exchange == "peer" and job.world == 4 and job.rank == 1 and job.mg is not None
    job.enable_timeline()
    job.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), sm.make_hp(**hp))
    torch.cuda.synchronize()
    labels = {lab for _, lab, _ in job.timeline_read()}
    assert {"main_start", "update_a", "update_b", "diag", "trsm", "gather"} <= labels
    job.close()


@pytest.mark.parametrize("driver", ["c", "python"])
@pytest.mark.parametrize("n,db", [(700, 128), (2500, 512), (3000, 256)])
def test_single_rank_peer_drivers_match_oracle(n, db, driver, monkeypatch):
    """the peer-store pipeline (owner chain with the assembled inverse, fused solve + scatter, flags, reduce slots) at
    world size 1, once through the C driver (smnngp_lml_mg_f64) and once through the Python panel loop"""
    import torch
    import smnngp_b200 as sm
    from smnngp_b200.distributed import DistributedLML
    from tests.synth import regression_data, DEFAULT_HP as hp
    monkeypatch.setenv("SMNNGP_MG_DRIVER", driver)
    d = 8
    x, y, *_ = regression_data(n, d)
    # emulate=(1, 0) = a one-rank job whose only peer is itself: results are the real ones
    job = DistributedLML(n, d, sm.StackSpec(3, "relu", "mlp"), "cuda", block=db, emulate=(1, 0), exchange="peer")
    assert (job.mg is not None) == (driver == "c")
    xd, yd, hpd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), sm.make_hp(**hp)
    out, info = job.lml(xd, yd, hpd)
    out2, _ = job.lml(xd, yd, hpd)
    ref = _ref(n, d, "student_t")
    assert int(info.item()) == 0 and out2.cpu().tolist() == out.cpu().tolist()
    assert abs(out[1].item() - ref) <= 1e-8 * abs(ref)
    job.close()


@pytest.mark.parametrize("driver", ["c", "python"])
@pytest.mark.parametrize("n,t,c,db", [(1500, 300, 2, 128), (2600, 700, 1, 256), (3000, 200, 3, 512)])
def test_single_rank_predict_drivers_match_oracle(n, t, c, db, driver, monkeypatch):
    """NNGPKernel.predict / SPR.test_nll through the distributed drivers at world size 1 (peer-store pipeline, carried
    right-hand-side and test rows, Z / result exchange through the peer slots): C driver (smnngp_predict_mg_f64,
    smnngp_test_nll_mg_f64) and Python panel loop against the oracle"""
    import torch
    import smnngp_b200 as sm
    from oracle import nngp_oracle as orc
    from smnngp_b200.distributed import DistributedPredict
    from tests.synth import regression_data, DEFAULT_HP as hp
    monkeypatch.setenv("SMNNGP_MG_DRIVER", driver)
    d = 8
    x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
    rng = np.random.default_rng(5)
    Y = np.column_stack([y] + [rng.standard_normal(n) for _ in range(c - 1)])
    spec = sm.StackSpec(3, "relu", "mlp")
    job = DistributedPredict(n, d, t, c, spec, "cuda", block=db, emulate=(1, 0), exchange="peer")
    assert (job.mg is not None) == (driver == "c")
    hpd = sm.make_hp(**hp)
    xd, Yd, xtd = torch.from_numpy(x).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(xt).cuda()
    kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"])
    mean_ref, cov_ref = orc.nt_predict(x, Y, xt, hp["eps"], kernel_kwargs=kw)
    vref, ktt = np.diag(cov_ref), orc.nngp_diag(xt, **kw)
    for _ in range(2):                                        # second call: sequence numbers / buffers reused
        mean, var, info = job.predict(xd, Yd, xtd, hpd)
        mean, var = mean.cpu().numpy(), var.cpu().numpy()
        assert int(info.item()) == 0
        assert np.abs(mean - mean_ref).max() <= 1e-8 * np.abs(mean_ref).max()
        assert np.all(np.abs(var - vref) <= 1e-8 * np.abs(vref) + 1e-13 * ktt)
    if c == 1:
        nll, m1, v1, info = job.test_nll(xd, torch.from_numpy(y).cuda(), xtd, torch.from_numpy(yt).cuda(), ym, ys, hpd)
        ref = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
        assert int(info.item()) == 0 and abs(float(nll.item()) - ref) <= 1e-8 * abs(ref)
    job.close()


@pytest.mark.parametrize("n,db,act,kind", [(1500, 128, "relu", "student_t"), (2600, 256, "erf", "gauss"),
                                           (3000, 512, "relu", "student_t")])
def test_single_rank_distributed_gradient_matches_oracle(n, db, act, kind):
    """smnngp_lml_grad_mg_f64 at world size 1 (identity rows carried through the block-cyclic factorisation with the
    active-row limit, U copied through the peer region, strip SYRK from the diagonal, strip dual Gram pass, partial sums
    through the peer slots): value and gradient against the oracle and against the single-GPU fused call"""
    import torch
    import smnngp_b200 as sm
    from oracle import nngp_oracle as orc
    from smnngp_b200.distributed import DistributedGrad
    from tests.synth import regression_data, DEFAULT_HP as hp0
    d = 8
    hp = dict(hp0, b_std=0.3 if act == "erf" else hp0["b_std"], eps=1e-4)
    x, y, *_ = regression_data(n, d)
    spec = sm.StackSpec(3, act, "mlp")
    job = DistributedGrad(n, d, spec, "cuda", block=db, emulate=(1, 0))
    hpd = sm.make_hp(**hp)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    kw = dict(num_hiddens=3, act=act, arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"],
              eps=hp["eps"], kind=kind, a=hp["alpha"], b=hp["beta"])
    lref, gref = orc.spr_loss_grad(x, y, **kw)
    o1, g1, _ = sm.device.lml_grad(xd, yd, spec=spec, hp=hpd, kind=kind)
    for _ in range(2):
        out, grad, info = job.lml_grad(xd, yd, hpd, kind=kind)
        g = grad.cpu().numpy()
        assert int(info.item()) == 0
        assert abs(out[1].item() - lref) <= 1e-8 * abs(lref)
        assert np.all(np.abs(g - gref) <= 1e-6 * np.abs(gref) + 1e-10 * np.abs(gref).max()), (g, gref)
        assert np.all(np.abs(g - g1.cpu().numpy()) <= 1e-9 * np.abs(gref) + 1e-12 * np.abs(gref).max())
    job.close()


def _grad_worker(rank, world, port, n, d, db, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import smnngp_b200 as sm
        from smnngp_b200.distributed import DistributedGrad
        from tests.synth import regression_data, DEFAULT_HP as hp0
        hp = dict(hp0, eps=1e-4)
        x, y, *_ = regression_data(n, d)
        dev = torch.device("cuda", rank)
        job = DistributedGrad(n, d, sm.StackSpec(3, "relu", "mlp"), dev, block=db)
        xd, yd, hpd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), sm.make_hp(device=dev, **hp)
        out, grad, info = job.lml_grad(xd, yd, hpd)
        out2, grad2, _ = job.lml_grad(xd, yd, hpd)
        assert grad2.cpu().tolist() == grad.cpu().tolist() and out2.cpu().tolist() == out.cpu().tolist()
        q.put((rank, out.cpu().tolist(), grad.cpu().tolist(), int(info.item())))
        job.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,db", [(2900, 128), (3000, 256), (1501, 128)])
def test_two_rank_gradient_matches_oracle(n, db):
    import torch
    import torch.multiprocessing as mp
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp0
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + n % 100
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, n, 8, db, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hp = dict(hp0, eps=1e-4)
    x, y, *_ = regression_data(n, 8)
    lref, gref = orc.spr_loss_grad(x, y, num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"],
                                   last_w_std=hp["last_w_std"], eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"])
    for rank, out, grad, info in res:
        g = np.array(grad)
        assert info == 0 and abs(out[1] - lref) <= 1e-8 * abs(lref)
        assert np.all(np.abs(g - gref) <= 1e-6 * np.abs(gref) + 1e-10 * np.abs(gref).max()), (g, gref)
    assert res[0][2] == res[1][2]


def _predict_worker(rank, world, port, n, t, c, d, db, q, exchange):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import smnngp_b200 as sm
        from smnngp_b200.distributed import DistributedPredict
        from tests.synth import regression_data, DEFAULT_HP as hp
        x, y, xt, *_ = regression_data(n, d, t=t)
        rng = np.random.default_rng(5)
        Y = np.column_stack([y] + [rng.standard_normal(n) for _ in range(c - 1)])
        dev = torch.device("cuda", rank)
        solver = DistributedPredict(n, d, t, c, sm.StackSpec(3, "relu", "mlp"), dev, block=db, exchange=exchange)
        mean, var, info = solver.predict(torch.from_numpy(x).to(dev), torch.from_numpy(Y).to(dev),
                                         torch.from_numpy(xt).to(dev), sm.make_hp(device=dev, **hp))
        q.put((rank, mean.cpu().numpy(), var.cpu().numpy(), int(info.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("n,t,c,db", [(1500, 300, 2, 128), (3000, 700, 1, 256)])
def test_two_rank_predict_matches_oracle(n, t, c, db, exchange):
    import torch
    import torch.multiprocessing as mp
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + n % 100 + (50 if exchange == "peer" else 0)
    procs = [ctx.Process(target=_predict_worker, args=(r, 2, port, n, t, c, 8, db, q, exchange)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, xt, *_ = regression_data(n, 8, t=t)
    rng = np.random.default_rng(5)
    Y = np.column_stack([y] + [rng.standard_normal(n) for _ in range(c - 1)])
    kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"])
    mean_ref, cov_ref = orc.nt_predict(x, Y, xt, hp["eps"], kernel_kwargs=kw)
    vref, ktt = np.diag(cov_ref), orc.nngp_diag(xt, **kw)
    for rank, mean, var, info in res:
        assert info == 0
        assert np.abs(mean - mean_ref).max() <= 1e-8 * np.abs(mean_ref).max()
        assert np.all(np.abs(var - vref) <= 1e-8 * np.abs(vref) + 1e-13 * ktt)
