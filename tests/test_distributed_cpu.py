"""world_size-2 (and 3) gloo runs of the multi-GPU driver's host logic on CPU, NumPy backend as the arithmetic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, d, db, kind, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import smnngp_b200 as sm
        from smnngp_b200.distributed import DistributedLML
        from tests.np_backend import NumpyBackend
        from tests.synth import regression_data, DEFAULT_HP as hp
        x, y, *_ = regression_data(n, d)
        spec = sm.StackSpec(3, "relu", "mlp")
        hpt = torch.tensor([hp[k] for k in ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")], dtype=torch.float64)
        solver = DistributedLML(n, d, spec, "cpu", block=db, backend=NumpyBackend())
        out, info = solver.lml(torch.from_numpy(x), torch.from_numpy(y), hpt, kind=kind)
        q.put((rank, out.tolist(), int(info[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,db,kind", [(2, 700, 128, "student_t"), (2, 512, 256, "gauss"), (3, 900, 128, "student_t"),
                                             (2, 130, 128, "student_t")])
def test_block_row_cyclic_lml_matches_oracle(world, n, db, kind):
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    d = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, db, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, *_ = regression_data(n, d)
    ref = orc.spr_loss(x, y, num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"],
                       last_w_std=hp["last_w_std"], eps=hp["eps"], kind=kind, a=hp["alpha"], b=hp["beta"])
    for rank, out, info in res:
        assert info == 0
        assert abs(out[1] - ref) <= 1e-9 * abs(ref), (rank, out[1], ref)
    assert res[0][1] == res[1][1], "every rank must hold the same result"


def _predict_worker(rank, world, port, n, t, c, d, db, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import smnngp_b200 as sm
        from smnngp_b200.distributed import DistributedPredict
        from tests.np_backend import NumpyBackend
        from tests.synth import regression_data, DEFAULT_HP as hp
        x, y, xt, *_ = regression_data(n, d, t=t)
        rng = np.random.default_rng(5)
        Y = np.column_stack([y] + [rng.standard_normal(n) for _ in range(c - 1)])
        spec = sm.StackSpec(3, "relu", "mlp")
        hpt = torch.tensor([hp[k] for k in ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")], dtype=torch.float64)
        solver = DistributedPredict(n, d, t, c, spec, "cpu", block=db, backend=NumpyBackend())
        mean, var, info = solver.predict(torch.from_numpy(x), torch.from_numpy(Y), torch.from_numpy(xt), hpt)
        q.put((rank, mean.numpy(), var.numpy(), int(info[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,t,c,db", [(2, 700, 130, 2, 128), (3, 600, 300, 1, 128), (2, 512, 40, 3, 256),
                                            (2, 300, 5, 1, 128)])
def test_block_row_cyclic_predict_matches_oracle(world, n, t, c, db):
    """right-hand sides and test-train rows carried through the distributed factorisation (NCCL-style exchange,
    NumPy arithmetic): NNGPKernel.predict with the relative regulariser"""
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    d = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n + t) % 2000
    procs = [ctx.Process(target=_predict_worker, args=(r, world, port, n, t, c, d, db, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, xt, *_ = regression_data(n, d, t=t)
    rng = np.random.default_rng(5)
    Y = np.column_stack([y] + [rng.standard_normal(n) for _ in range(c - 1)])
    kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"])
    mean_ref, cov_ref = orc.nt_predict(x, Y, xt, hp["eps"], kernel_kwargs=kw)
    vref = np.diag(cov_ref)
    ktt = orc.nngp_diag(xt, **kw)
    for rank, mean, var, info in res:
        assert info == 0 and mean.shape == (t, c) and var.shape == (t,)
        assert np.abs(mean - mean_ref).max() <= 1e-8 * np.abs(mean_ref).max(), rank
        assert np.all(np.abs(var - vref) <= 1e-8 * np.abs(vref) + 1e-12 * ktt), rank
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


def _nll_worker(rank, world, port, n, t, d, db, kind, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import smnngp_b200 as sm
        from smnngp_b200.distributed import DistributedPredict
        from tests.np_backend import NumpyBackend
        from tests.synth import regression_data, DEFAULT_HP as hp
        x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
        hpt = torch.tensor([hp[k] for k in ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")], dtype=torch.float64)
        solver = DistributedPredict(n, d, t, 1, sm.StackSpec(3, "relu", "mlp"), "cpu", block=db, backend=NumpyBackend())
        nll, mean, var, info = solver.test_nll(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(xt),
                                               torch.from_numpy(yt), ym, ys, hpt, kind=kind)
        q.put((rank, float(nll[0]), int(info[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,t,db,kind", [(2, 600, 70, 128, "student_t"), (3, 500, 60, 128, "gauss")])
def test_block_row_cyclic_test_nll_matches_oracle(world, n, t, db, kind):
    """SPR.test_nll on P ranks: predictive + second factorisation with the likelihood's jitter + closed-form tail"""
    from oracle import nngp_oracle as orc
    from tests.synth import regression_data, DEFAULT_HP as hp
    d = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + n + t) % 2000
    procs = [ctx.Process(target=_nll_worker, args=(r, world, port, n, t, d, db, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
    ref = orc.spr_test_nll(x, y, xt, yt, ym, ys, num_hiddens=3, act="relu", arch="mlp", w_std=hp["w_std"],
                           b_std=hp["b_std"], last_w_std=hp["last_w_std"], eps=hp["eps"], kind=kind, a=hp["alpha"],
                           b=hp["beta"])
    for rank, nll, info in res:
        assert info == 0 and abs(nll - ref) <= 1e-8 * abs(ref), (rank, nll, ref)


def test_layout_bookkeeping():
    from smnngp_b200.distributed import BlockRowCyclic
    for (m, P, db) in [(701, 2, 128), (1025, 3, 256), (60001, 8, 512), (513, 4, 128), (129, 2, 128)]:
        seen = np.zeros(m, dtype=int)
        for r in range(P):
            lay = BlockRowCyclic(m, m - 1, P, r, db)
            off = 0
            for b in lay.local_blocks():
                assert lay.owner(b) == r and lay.local_offset(b) == off
                seen[b * db: b * db + lay.block_rows(b)] += 1
                off += lay.block_rows(b)
            assert off == lay.local_rows()
            for gb in range(lay.nblocks + 1):
                o, cnt = lay.rows_from_block(gb)
                want = sum(lay.block_rows(b) for b in lay.local_blocks() if b >= gb)
                assert cnt == want and o == lay.local_rows() - want
        assert (seen == 1).all()


def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smnngp_b200.distributed import sweep_eps
        t = 5
        calls = []

        def evaluate(eps):                       # stand-in for GridSearch.point: results are functions of eps only
            calls.append(eps)
            base = torch.arange(t, dtype=torch.float64)
            return base * eps, base + eps, 10.0 * eps, 100.0 * eps, torch.tensor([0 if eps < 1.0 else 7])

        eps_list = [1e-6 * 10 ** (0.5 * k) for k in range(11)] + [2.0]     # find.py:20 has eleven
        res = sweep_eps(eps_list, evaluate, t=t, device="cpu")
        q.put((rank, calls, [(m.tolist(), v.tolist(), ld, qd, info) for m, v, ld, qd, info in res], eps_list))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_eps_sweep_replicas(world):
    """one epsilon per rank (experiments/regression/find.py:141): every rank evaluates only its share and sees all results"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + world
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls, res, eps_list in got:
        assert calls == eps_list[rank::world]
        assert len(res) == len(eps_list)
        for eps, (m, v, ld, qd, info) in zip(eps_list, res):
            assert m == [k * eps for k in range(5)] and v == [k + eps for k in range(5)]
            assert ld == 10.0 * eps and qd == 100.0 * eps and info == (0 if eps < 1.0 else 7)


@pytest.mark.parametrize("world,n,block", [(8, 60000, 512), (4, 60000, 512), (2, 60000, 512), (8, 150000, 512),
                                           (8, 20000, 256), (3, 5000, 128), (2, 700, 128)])
def test_c_driver_block_layouts(world, n, block):
    """host logic of csrc/multigpu.cu (no GPU needed): every block -> rank map gives each rank blocks of the form
    LB * P + (LB odd ? odd_off : even_off) - the property the scatter epilogue and the update mask rely on - and the
    default (auto) never loads the busiest rank more than the plain cyclic map does"""
    import ctypes as C
    import smnngp_b200 as sm
    lib = sm._lib.load()
    nb = -(-(n + 1) // block)
    res = {}
    for layout in (0, 1, 2, 3):
        owner = (C.c_int * nb)()
        load = (C.c_double * world)()
        assert lib.smnngp_mg_layout(world, n, 1, block, layout, owner, load) == nb
        owner = list(owner)
        assert set(owner) <= set(range(world))
        for r in range(world):
            mine = [b for b in range(nb) if owner[b] == r]
            if len(mine) > 1:
                even, odd = mine[0], mine[1] - world
                assert all(b == lb * world + (odd if lb & 1 else even) for lb, b in enumerate(mine))
            # one block per cycle of `world` consecutive blocks (so local block index = position in the list)
            assert all(mine[i + 1] - mine[i] < 2 * world for i in range(len(mine) - 1))
        res[layout] = max(load) / (sum(load) / world)
    if world > 2:
        assert res[3] <= res[0] + 1e-12 and res[3] <= res[1] + 1e-12 and res[3] <= res[2] + 1e-12
    else:
        assert res[3] == res[0]                    # two ranks: auto = plain cyclic (measured faster)
    if world == 8 and n == 60000:
        assert res[0] > 1.08 and res[3] < 1.03          # 8.6 % -> 2.5 % over the mean (profiles/r02_layouts_8gpu.txt)


def test_stream_watchdog_releases_waits_and_poisons_info():
    """dead-peer protection of the Python panel loop: an event that never completes makes the watchdog write every flag
    word far ahead of any sequence number and set info = INT_MAX; a completed event leaves everything alone"""
    from smnngp_b200.distributed import StreamWatchdog

    class Ev:
        def __init__(self, done):
            self.done = done

        def query(self):
            return self.done

    flags = torch.zeros(32, dtype=torch.int64)
    info = torch.zeros(1, dtype=torch.int32)
    dog = StreamWatchdog(flags, timeout_s=0.2)
    dog.watch(Ev(True), info, 1234).join(timeout=5)
    assert not dog.fired and int(info.item()) == 0 and int(flags.abs().sum()) == 0
    dog.watch(Ev(False), info, 1234, poll_s=0.02).join(timeout=5)
    assert dog.fired and int(info.item()) == 0x7fffffff and bool((flags >= 1234 + (1 << 40)).all())


def test_c_driver_block_layouts_property():
    """property test (hypothesis) of the host-side layout function over random group sizes / orders / block sizes: every
    map is a partition with one block per rank per cycle in the alternating form, and `auto` is never worse than the
    other maps under the load model"""
    import ctypes as C
    from hypothesis import given, settings, strategies as st
    import smnngp_b200 as sm
    lib = sm._lib.load()

    @settings(max_examples=60, deadline=None)
    @given(world=st.integers(1, 8), n=st.integers(1, 40000), block=st.sampled_from([128, 256, 384, 512]),
           extra=st.integers(0, 700))
    def check(world, n, block, extra):
        nb = -(-(n + extra) // block)
        worst = {}
        for layout in (0, 1, 2, 3):
            owner = (C.c_int * nb)()
            load = (C.c_double * world)()
            assert lib.smnngp_mg_layout(world, n, extra, block, layout, owner, load) == nb
            owner = list(owner)
            for r in range(world):
                mine = [b for b in range(nb) if owner[b] == r]
                if len(mine) > 1:
                    even, odd = mine[0], mine[1] - world
                    assert all(b == lb * world + (odd if lb & 1 else even) for lb, b in enumerate(mine))
            # every window of `world` consecutive blocks that is aligned to a cycle of the map holds each rank at most
            # twice (snake turn-around) and the union over all blocks is a partition by construction
            assert all(0 <= o < world for o in owner)
            worst[layout] = max(load)
        if world > 2:
            assert worst[3] <= min(worst[0], worst[1], worst[2]) * (1 + 1e-12)
        else:
            assert worst[3] == worst[0]

    check()
