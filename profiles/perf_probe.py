"""Quick stage timings on one B200 (development tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smnngp_b200 as sm
from tests.synth import pixel_data, regression_data, DEFAULT_HP

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

hp = sm.make_hp(1.0, 1e-8, 1.0, 1e-6, 2.0, 2.0)
spec = sm.StackSpec(3, "relu", "mlp")
which = sys.argv[1:] or ["potrf", "gram", "lml"]
sm._lib.load().smnngp_set_tile_variant(int(os.environ.get("TILE", "0")))
sm._lib.load().smnngp_set_lookahead(int(os.environ.get("LOOKAHEAD", "1")))
nbs = [int(v) for v in os.environ.get("NBS", "0").split(",")]
if "TAIL" in os.environ:
    sm._lib.load().smnngp_set_tail_cols(int(os.environ["TAIL"]))
if "RESERVE" in os.environ:
    a_, b_ = (int(v) for v in os.environ["RESERVE"].split(","))
    sm._lib.load().smnngp_set_lookahead_reserve(a_, b_)
if "potrf" in which:
    for n in (8192, 16384, 32768):
        a = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        for nb in nbs:
            sm.device.set_panel_width(nb)
            def run():
                a.zero_(); a.diagonal().fill_(float(n)); a[:, 0] = 1.0; a[0, 0] = float(n)
                sm.device.potrf_(a)
            def base():
                a.zero_(); a.diagonal().fill_(float(n)); a[:, 0] = 1.0; a[0, 0] = float(n)
            t = timed(run) - timed(base)
            print(f"potrf N={n} nb={nb}: {t:9.3f} ms  {n**3/3/t*1e-9:7.2f} TFLOP/s", flush=True)
        del a
    sm.device.set_panel_width(0)
if "cusolver" in which:
    # library bar on the same GPU: cuSOLVER Dpotrf through torch.linalg.cholesky_ex (lower), same SPD matrix as above
    torch.backends.cuda.preferred_linalg_library("cusolver")
    for n in [int(v) for v in os.environ.get("CUSOLVER_N", "8192,16384,32768,60000").split(",")]:
        a = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        out = torch.empty_like(a)
        info = torch.empty((), dtype=torch.int32, device="cuda")
        def base():
            a.zero_(); a.diagonal().fill_(float(n)); a[:, 0] = 1.0; a[0, :] = 1.0; a[0, 0] = float(n)
        def run():
            base()
            torch.linalg.cholesky_ex(a, upper=False, out=(out, info))
        t = timed(run) - timed(base)
        print(f"cusolver Dpotrf (torch.linalg.cholesky_ex) N={n}: {t:9.3f} ms  {n**3/3/t*1e-9:7.2f} TFLOP/s", flush=True)
        del a, out
        torch.cuda.empty_cache()
if "potrf60k" in which:
    n = 60000
    a = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    def run():
        a.zero_(); a.diagonal().fill_(float(n)); a[:, 0] = 1.0; a[0, 0] = float(n)
        sm.device.potrf_(a)
    def base():
        a.zero_(); a.diagonal().fill_(float(n)); a[:, 0] = 1.0; a[0, 0] = float(n)
    t = timed(run) - timed(base)
    print(f"potrf N={n}: {t:9.3f} ms  {n**3/3/t*1e-9:7.2f} TFLOP/s", flush=True)
    del a
    torch.cuda.empty_cache()
if "gram" in which:
    for (n, d) in ((20000, 784), (10000, 8), (20000, 3072)):
        x = torch.from_numpy(pixel_data(n, d)[0]).cuda()
        out = torch.empty((n, n), dtype=torch.float64, device="cuda")
        for L in (0, 3):
            sp = sm.StackSpec(L, "relu", "mlp")
            t = timed(lambda: sm.device.gram(x, spec=sp, hp=hp, lower_only=True, out=out))
            print(f"gram lower N={n} D={d} L={L}: {t:9.3f} ms  contraction {n*(n+1)*d/t*1e-9:7.2f} TFLOP/s  {L*n*(n+1)/2/t*1e-6:8.2f} Geval/s", flush=True)
        t = timed(lambda: sm.device.gram(x, spec=spec, hp=hp, out=out))
        print(f"gram full(mirror) N={n} D={d} L=3: {t:9.3f} ms", flush=True)
        del x, out
if "gram60k" in which:
    n, d = 60000, 784
    x = torch.from_numpy(pixel_data(n, d)[0]).cuda()
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    for L in (0, 1, 3):
        sp = sm.StackSpec(L, "relu", "mlp")
        t = timed(lambda: sm.device.gram(x, spec=sp, hp=hp, lower_only=True, out=out))
        print(f"gram lower N={n} D={d} L={L}: {t:9.3f} ms  contraction {n*(n+1)*d/t*1e-9:7.2f} TFLOP/s  {L*n*(n+1)/2/t*1e-6:8.2f} Geval/s", flush=True)
    del x, out
if "grid" in which:
    for (n, t, d) in ((10000, 1000, 8), (10000, 1000, 784)):
        xs, ys, xts, *_ = regression_data(n, d, t=t)
        x, y, xt = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda(), torch.from_numpy(xts).cuda()
        gs = sm.device.GridSearch(x, y, xt, spec=spec)
        t_cached = timed(lambda: gs.point(hp))
        t_scratch = timed(lambda: sm.device.grid_point(x, y, xt, spec=spec, hp=hp))
        print(f"find.py grid point N={n} T={t} D={d}: cached base {t_cached:8.3f} ms, from scratch {t_scratch:8.3f} ms", flush=True)
        del gs
if "lml" in which:
    for (n, d) in ((10000, 8), (30000, 784), (60000, 784)):
        xs, ys, *_ = pixel_data(n, d) if d > 100 else regression_data(n, d)[:2] + (None,)
        x, y = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
        res = {}
        def run():
            res["o"] = sm.device.lml(x, y, spec=spec, hp=hp)
        t = timed(run, reps=2)
        out, info = res["o"]
        f = n * (n + 1) * d + n ** 3 / 3 + n * n
        print(f"lml N={n} D={d}: {t:9.3f} ms  {f/t*1e-9:7.2f} TFLOP/s  loss={out[1].item():.12f} info={info.item()}", flush=True)
        del x, y
        sm.device.release_workspaces(); torch.cuda.empty_cache()
if "dist1" in which:
    from smnngp_b200.distributed import DistributedLML
    for (n, d) in ((10000, 8), (30000, 784), (60000, 784)):
        xs, ys, *_ = pixel_data(n, d) if d > 100 else regression_data(n, d)[:2] + (None,)
        x, y = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
        for res in (0, 4, 8):
            os.environ["SMNNGP_SM_RESERVE"] = str(res)
            solver = DistributedLML(n, d, spec, "cuda")
            r = {}
            def run():
                r["o"] = solver.lml(x, y, hp)
            t = timed(run, reps=2)
            f = n * (n + 1) * d + n ** 3 / 3 + n * n
            print(f"stage-path P=1 reserve={res} N={n} D={d} db={solver.db}: {t:9.3f} ms {f/t*1e-9:7.2f} TFLOP/s loss={r['o'][0][1].item():.12f}", flush=True)
            del solver
            torch.cuda.empty_cache()
        del x, y

if "graph" in which:
    for (n, d) in ((2000, 8), (5000, 8), (10000, 8), (20000, 784)):
        xs, ys, *_ = pixel_data(n, d) if d > 100 else regression_data(n, d)[:2] + (None,)
        x, y = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
        t0 = timed(lambda: sm.device.lml(x, y, spec=spec, hp=hp), reps=5)
        g = sm.device.LmlGraph(n, d, spec=spec)
        t1 = timed(lambda: g(x, y, hp), reps=5)
        print(f"lml N={n} D={d}: eager {t0:8.3f} ms   cuda-graph replay {t1:8.3f} ms", flush=True)
