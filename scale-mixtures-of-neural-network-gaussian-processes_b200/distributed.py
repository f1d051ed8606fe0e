"""Multi-GPU exact-GP log marginal likelihood: one process per GPU (torch.distributed, NCCL over NVLink),
block-row-cyclic layout, right-looking blocked Cholesky with the solves fused in.

Layout (P ranks, distribution block DB = outer panel width, rows 0..N-1 = K + eps I, row N = y^T):
  global block b = rows [b DB, (b+1) DB) lives on rank b mod P, full width, lower part only.  Every rank holds
  all of X (376 MB at C3) and generates exactly the Gram rows it owns - the N x N matrix never exists in one
  place and there is no exchange step in the Gram stage.
Per panel p (columns [c0, c1)):
  owner factors the DB x DB diagonal block           -> ONE broadcast of (L_pp, block inverses)  [2.6 MB]
  every rank TRSMs its own panel rows (tensor pipe)  -> ONE all-gather of the panel             [<= 245 MB]
  every rank updates its own trailing rows with a single kernel launch (block-row-cyclic lower mask).
This is the P x 1 case of a 2-D block-cyclic grid.  With NVSwitch every rank receives the whole panel at full
link bandwidth (14.4 GB per rank over the whole factorisation at C3, ~20 ms at 700 GB/s vs ~300 ms of math), so a
second grid dimension would only shrink a term that is already negligible, while P x 1 spreads the TRSM over all
ranks and needs two collectives per panel instead of four.

The orchestration below is backend-agnostic: ``CudaBackend`` drives the stage-level C-ABI (product path);
the CPU test-suite injects a NumPy backend to exercise the ownership / exchange logic under gloo.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from .device import ACT, ARCH, KIND, SHIFT, StackSpec

PB = 128


def _cdiv(a, b):
    return (a + b - 1) // b


class BlockRowCyclic:
    """Pure host logic: which global rows live where."""

    def __init__(self, m_total: int, n_cols: int, world: int, rank: int, db: int):
        assert db % PB == 0
        self.m_total, self.n, self.P, self.rank, self.db = m_total, n_cols, world, rank, db
        self.nblocks = _cdiv(m_total, db)

    def owner(self, b):
        return b % self.P

    def block_rows(self, b):
        return min((b + 1) * self.db, self.m_total) - b * self.db

    def local_blocks(self, rank=None):
        r = self.rank if rank is None else rank
        return list(range(r, self.nblocks, self.P))

    def local_rows(self, rank=None):
        return sum(self.block_rows(b) for b in self.local_blocks(rank))

    def local_offset(self, b, rank=None):
        """local row offset of global block b on its owner"""
        r = self.owner(b) if rank is None else rank
        return sum(self.block_rows(x) for x in range(r, b, self.P))

    def first_block_from(self, gb, rank=None):
        """smallest block index >= gb owned by rank (may be >= nblocks)"""
        r = self.rank if rank is None else rank
        return gb + ((r - gb) % self.P)

    def rows_from_block(self, gb, rank=None):
        """(local row offset, row count) of this rank's rows in global blocks >= gb"""
        r = self.rank if rank is None else rank
        fb = self.first_block_from(gb, r)
        if fb >= self.nblocks:
            return self.local_rows(r), 0
        off = self.local_offset(fb, r)
        return off, self.local_rows(r) - off


class CudaBackend:
    """Stage-level C-ABI on torch CUDA tensors (views keep their row pitch)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.lib = _lib.load()

    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"libsmnngp stage '{what}' failed with status {rc}")

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def qtable(self, x, spec: StackSpec, hp):
        nh, act, arch = spec.ids()
        n, d = x.shape
        n_act = max(nh + (1 if arch == 1 else 0), 1)
        tab, q, scal = self.empty(n_act, n), self.empty(n), self.zeros(16)
        self._ck(self.lib.smnngp_stage_qtable_f64(self._s(), self._p(x), n, d, nh, act, arch, self._p(hp), self._p(tab),
                                                  n, self._p(q), self._p(scal)), "qtable")
        return tab, q, scal

    def gram_block(self, x1, x2, spec, hp, tab1, tab2, scal, shift, symmetric_lower, out):
        nh, act, arch = spec.ids()
        self._ck(self.lib.smnngp_stage_gram_f64(self._s(), self._p(x1), x1.shape[0], self._p(x2), x2.shape[0],
                                                x1.shape[1], nh, act, arch, self._p(hp), self._p(tab1), tab1.stride(0),
                                                self._p(tab2), tab2.stride(0), self._p(scal), SHIFT[shift],
                                                1 if symmetric_lower else 0, self._p(out), out.stride(0)), "gram")

    def factor_diag(self, a, linv, logdet, info, gcol0):
        self._ck(self.lib.smnngp_stage_factor_diag_f64(self._s(), self._p(a), a.stride(0), a.shape[0], self._p(linv),
                                                       self._p(logdet), self._p(info), gcol0), "factor_diag")

    def trsm(self, r, ldiag, linv):
        self._ck(self.lib.smnngp_stage_trsm_f64(self._s(), self._p(r), r.stride(0), r.shape[0], r.shape[1],
                                                self._p(ldiag), ldiag.stride(0), self._p(linv)), "trsm")

    def update(self, a, b, c, lower, cyc_db, cyc_p, base_shift, sm_reserve=0):
        self._ck(self.lib.smnngp_stage_update_f64(self._s(), self._p(a), a.stride(0), self._p(b), b.stride(0),
                                                  self._p(c), c.stride(0), c.shape[0], c.shape[1], a.shape[1],
                                                  1 if lower else 0, cyc_db, cyc_p, base_shift, int(sm_reserve)),
                 "update")

    def sumsq(self, z, out):
        self._ck(self.lib.smnngp_stage_sumsq_f64(self._s(), self._p(z), z.shape[0], self._p(out)), "sumsq")

    def predict_finalize(self, v, z, ktt, info):
        """V [T, n] = K_td L^-T rows, Z [C, n] = (L^-1 Y)^T rows -> (mean [T, C], var [T])"""
        t, n = v.shape
        c = z.shape[0]
        mean, var = self.empty(t, c), self.empty(t)
        if t > 0:
            self._ck(self.lib.smnngp_stage_predict_finalize_f64(self._s(), self._p(v), v.stride(0), self._p(z),
                                                                z.stride(0), self._p(ktt), t, c, n, self._p(info),
                                                                self._p(mean), self._p(var)), "predict_finalize")
        return mean, var

    def test_nll_finalize(self, mean, var, ytest, n, y_mean, y_std, hp, kind, quad2, info):
        out = self.empty(1)
        q2 = self._p(quad2) if quad2 is not None else C.c_void_p(0)
        self._ck(self.lib.smnngp_stage_test_nll_finalize_f64(self._s(), self._p(mean), self._p(var), self._p(ytest),
                                                             mean.shape[0], n, float(y_mean), float(y_std),
                                                             self._p(hp), KIND[kind], q2, self._p(info), C.c_void_p(0),
                                                             self._p(out)), "test_nll_finalize")
        return out

    def lml_finalize(self, sums, hp, kind, n, info):
        out = self.empty(4)
        self._ck(self.lib.smnngp_stage_lml_finalize_f64(self._s(), self._p(sums), self._p(hp), KIND[kind], n,
                                                        self._p(info), self._p(out)), "lml_finalize")
        return out


def _on_own_device(method):
    """run a driver method with the driver's device current (the C library launches on the current device when it is
    handed torch's default stream, which is the legacy stream and carries no device of its own)"""
    import functools

    @functools.wraps(method)
    def wrapped(self, *a, **k):
        if self.a.is_cuda:
            with torch.cuda.device(self.a.device):
                return method(self, *a, **k)
        return method(self, *a, **k)
    return wrapped


class StreamWatchdog:
    """Dead-peer protection of the Python panel loop (the C driver has its own, csrc/multigpu.cu): the flag waits are
    stream memory operations without a time-out, so after every evaluation an event is recorded and a daemon thread
    polls it; if it has not completed after ``timeout_s`` the thread writes every flag word of this rank (the waits
    compare cyclically: a value far ahead of any sequence number satisfies them all) and poisons ``info`` - the stream
    drains and the results are NaN instead of the GPU hanging for ever."""

    def __init__(self, flags, timeout_s):
        self.flags, self.timeout_s = flags, float(timeout_s)          # flags: int64 tensor view of the local flag words
        self.fired = False

    def _release(self, info, seq_hint):
        self.fired = True
        side = torch.cuda.Stream(device=self.flags.device) if self.flags.is_cuda else None
        ctx = torch.cuda.stream(side) if side is not None else _NullCtx()
        with ctx:
            if info is not None:
                info.fill_(0x7fffffff)
            self.flags.fill_(int(seq_hint) + (1 << 40))
        if side is not None:
            side.synchronize()

    def watch(self, event, info, seq_hint, poll_s=0.05):
        import threading
        import time

        def run():
            t0 = time.monotonic()
            while not event.query():
                if time.monotonic() - t0 > self.timeout_s:
                    self._release(info, seq_hint)
                    return
                time.sleep(poll_s)

        th = threading.Thread(target=run, daemon=True)
        th.start()
        return th


class PeerUnavailable(RuntimeError):
    """raised on EVERY rank when the peer-memory exchange cannot be set up on some rank"""


class _RawCuda:
    """torch view of a raw device pointer (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3)


class PeerExchange:
    """Peer-visible buffers of the panel exchange (csrc/exchange.cu): every rank cudaMalloc's one region
    [W | flags | 3 panel slots], exports it as a CUDA IPC handle, and opens everybody else's, so kernels can store
    straight into the other GPUs over NVLink.  emulate: all "peers" alias the local region (timing dry-run)."""

    FLAG_WORDS = 16                       # word 0: W-ready (set by the panel owner); words 8..15: panel-ready per rank

    def __init__(self, lib, device, world, rank, group, n, db, emulate=False):
        self.lib, self.world, self.rank, self.db, self.n = lib, world, rank, db, n
        self.off_w = 0
        self.off_flags = db * db * 8
        self.off_panel = self.off_flags + 256
        self.slot_bytes = n * db * 8
        total = self.off_panel + 3 * self.slot_bytes
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        ok = lib.smnngp_peer_alloc(total, C.byref(ptr), handle) == 0
        self.local = ptr.value if ok else 0
        self._opened = []
        if ok:
            torch.as_tensor(_RawCuda(self.local + self.off_flags, (32,), "<i8"), device=device).zero_()
        if emulate or world == 1:
            if not ok:
                raise RuntimeError("smnngp_peer_alloc failed: " + lib.smnngp_last_error().decode())
            self.bases = [self.local] * world
        else:
            # every rank takes part in every collective below even if its own step failed, so that a failure on
            # one rank (no memory, CUDA IPC not permitted in this container) is seen by ALL ranks, which then
            # fall back to the NCCL exchange together instead of dead-locking
            mine = torch.tensor(list(handle) if ok else [0] * 64, dtype=torch.uint8, device=device)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine, group=group)          # also orders the zeroing above before any peer store
            self.bases = []
            for r in range(world):
                if r == rank:
                    self.bases.append(self.local)
                    continue
                q, hh = C.c_void_p(), (C.c_ubyte * 64)(*allh[r].cpu().tolist())
                if ok and any(hh) and lib.smnngp_peer_open(hh, C.byref(q)) == 0:
                    self._opened.append(q.value)
                    self.bases.append(q.value)
                else:
                    ok = False
                    self.bases.append(0)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            torch.cuda.synchronize(device)
            if int(flag.item()) == 0:
                self.close()
                raise PeerUnavailable("CUDA IPC peer memory could not be set up on every rank")
        self.w_local = torch.as_tensor(_RawCuda(self.local + self.off_w, (db, db), "<f8"), device=device)
        self.panel_local = [torch.as_tensor(_RawCuda(self.local + self.off_panel + k * self.slot_bytes, (n, db), "<f8"),
                                            device=device) for k in range(3)]
        self.flags_local = C.c_void_p(self.local + self.off_flags)
        self.w_ptrs = self._arr(self.off_w)
        self.flag_ptrs = self._arr(self.off_flags)
        self.panel_ptrs = [self._arr(self.off_panel + k * self.slot_bytes) for k in range(3)]
        # own-buffers-only views for the copy-engine flavour (the solve kernel then stores locally only)
        # (a NULL entry = "skip this rank")
        self.panel_ptrs_self = [(C.c_void_p * self.world)(*[(self.local + self.off_panel + k * self.slot_bytes)
                                                            if r == self.rank else None for r in range(self.world)])
                                for k in range(3)]
        self.flag_ptrs_none = (C.c_void_p * self.world)(*[None] * self.world)

    def _arr(self, off):
        return (C.c_void_p * self.world)(*[b + off for b in self.bases])

    def close(self):
        self.w_local = None                     # views of the region: must not outlive it
        self.panel_local = []
        for p in self._opened:
            self.lib.smnngp_peer_close(C.c_void_p(p))
        self._opened = []
        if self.local:
            self.lib.smnngp_peer_free(C.c_void_p(self.local))
            self.local = 0


def default_block(n, world):
    """distribution block = outer panel width: wide enough for the update kernel, small enough to balance"""
    if os.environ.get("SMNNGP_BLOCK"):                       # tuning experiments
        return int(os.environ["SMNNGP_BLOCK"])
    per_rank = n / max(world, 1)
    if per_rank >= 4096:
        return 512
    if per_rank >= 1024:
        return 256
    return 128


class DistributedLML:
    """SPR.loss (spax/models.py:93-98) sharded over the ranks of a process group.  Strong scaling: the problem
    is fixed, every rank owns ~1/P of the rows."""

    def __init__(self, n, d, spec: StackSpec, device, group=None, block=None, backend=None, emulate=None,
                 exchange="auto", extra_rows=1):
        """emulate=(world, rank): timing dry-run of ONE rank's work of a `world`-rank job on a single device - the
        collectives are replaced by local copies of the same size, so the numbers it produces are meaningless but
        every kernel launch has the shape it has in the real job (used to profile the schedule at 1 GPU cost)."""
        self.n, self.d, self.spec = int(n), int(d), spec
        self.group = group
        self.emulate = emulate is not None
        if self.emulate:
            self.world, self.rank = int(emulate[0]), int(emulate[1])
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.timeline = None          # list of (panel, label, event) when profiling is switched on
        self._tl_filter = None
        self._gidx = {}
        self.db = int(block) if block else default_block(self.n, self.world)
        self.be = backend if backend is not None else CudaBackend(device)
        # rows below the square part are carried through the factorisation (rows * L^-T): y^T here, the right-hand
        # sides and the test-train cross-Gram in DistributedPredict
        self.n_extra = int(extra_rows)
        self.gram_shift = "eps_abs"                                      # K + eps I (spax/models.py:96)
        self.lay = BlockRowCyclic(self.n + self.n_extra, self.n, self.world, self.rank, self.db)
        self.ld = _cdiv(self.n, 16) * 16
        self.mloc = self.lay.local_rows()
        self.a = self.be.empty(1, 16)             # device marker until the driver is chosen (row shard: see below)
        nblk = self.db // PB
        # largest per-rank panel piece over all panels (panel p -> rows in blocks >= p + 1)
        self.max_m = max(self.lay.rows_from_block(1, r)[1] for r in range(self.world)) if self.world > 1 else 0
        self._nccl_bufs = False
        # panel exchange: "peer" = stores into the other ranks' buffers over NVLink (csrc/exchange.cu), "nccl" =
        # broadcast + all-gather (also the path the CPU suite drives under gloo with the NumPy backend)
        # (a one-rank job may use it too when asked for explicitly: emulate=(1, 0) is the real computation)
        peer_ok = self.a.is_cuda and isinstance(self.be, CudaBackend) and (1 < self.world <= 8 or self.emulate)
        if exchange == "auto":
            exchange = os.environ.get("SMNNGP_EXCHANGE", "peer" if peer_ok else "nccl")
        if exchange == "peer" and not peer_ok:
            raise ValueError("exchange='peer' needs CUDA, the C-ABI backend and 2..8 ranks")
        self.exchange = exchange
        self.px = None
        self.mg = None
        self._dog = None
        # exchange == "peer": the whole evaluation is ONE C call per rank (csrc/multigpu.cu, smnngp_lml_mg_f64); the
        # Python panel loop below remains for the NCCL exchange, the NumPy backend of the CPU suite, the predictive
        # driver (DistributedPredict) and as a cross-check (SMNNGP_MG_DRIVER=python)
        use_c = exchange == "peer" and os.environ.get("SMNNGP_MG_DRIVER", "c") == "c"
        if use_c:
            try:
                self._create_mg(group)
            except PeerUnavailable as e:
                import warnings
                warnings.warn(f"smnngp: {e}; using the NCCL panel exchange")
                self.exchange = exchange = "nccl"
        if self.mg is None:                       # Python panel loop: the row shard lives here
            self.a = self.be.empty(max(self.mloc, 1), self.ld)
            self.diag = self.be.empty(self.db * self.db + nblk * PB * PB)
        if exchange == "peer" and self.mg is None:
            try:
                self.px = PeerExchange(self.be.lib, self.a.device, self.world, self.rank, group, self.n, self.db,
                                       emulate=self.emulate)
            except PeerUnavailable as e:                  # collective decision: every rank lands here together
                import warnings
                warnings.warn(f"smnngp: {e}; using the NCCL panel exchange")
                self.exchange = exchange = "nccl"
        if exchange == "peer" and self.mg is None:
            self.linv4 = self.be.empty(nblk * PB * PB)
            self.ploc = [self.be.empty(max(self.mloc, 1), self.db) for _ in range(2)]
            self.counters = torch.zeros(8, dtype=torch.int32, device=self.a.device)
            self.zvec = self.be.zeros(self.n)
            self.seq_base = 0
            self.send = self.gath = self.pfull = None
            self.wait_timeout_s = float(os.environ.get("SMNNGP_PEER_TIMEOUT_S", "20"))
            # "store": the solve's epilogue stores into every rank's buffer (default); "ce": the solve stores locally and
            # the copy engines push whole blocks (cudaMemcpy2DAsync).  Measured equal at 4 GPUs (610 ms) and 338 vs
            # 333 ms at 8 GPUs: the solve is bound by the math on its few reserved SMs, not by the remote stores.
            self.push = os.environ.get("SMNNGP_PUSH", "store")
            self.be.lib.smnngp_set_peer_wait_mode(int(os.environ.get("SMNNGP_PEER_WAIT", "0")))
        self.side = torch.cuda.Stream(device=self.a.device, priority=-1) if self.a.is_cuda else None
        # SMs the bulk update leaves free so the look-ahead chain (diagonal block, TRSM, NCCL) really overlaps:
        # the persistent update kernel otherwise occupies every SM until it ends
        self.sm_reserve = int(os.environ.get("SMNNGP_SM_RESERVE", "8" if self.world > 1 else "0"))
        self.reserve_below_s = float(os.environ.get("SMNNGP_RESERVE_BELOW_MS", "18")) * 1e-3
        self.sm_reserve_auto = "SMNNGP_SM_RESERVE" not in os.environ

    def _create_mg(self, group):
        """one smnngp_mg handle per rank; the 64-byte IPC handles travel through one all-gather"""
        lib, dev = self.be.lib, self.a.device
        h = C.c_void_p()
        with torch.cuda.device(dev):
            tc = getattr(self, "_tc", None)          # DistributedPredict: (test points, right-hand sides)
            if getattr(self, "_grad", False):        # DistributedGrad: identity rows ride along
                ok = lib.smnngp_mg_create_grad(C.byref(h), self.rank, self.world, self.n, self.db) == 0
            elif tc is None:
                ok = lib.smnngp_mg_create(C.byref(h), self.rank, self.world, self.n, self.db) == 0
            else:
                ok = lib.smnngp_mg_create_predict(C.byref(h), self.rank, self.world, self.n, tc[0], tc[1], self.db) == 0
            handle = (C.c_ubyte * 64)()
            if ok:
                lib.smnngp_mg_ipc_handle(h, handle)
            if self.emulate or self.world == 1:
                if not ok:
                    raise RuntimeError("smnngp_mg_create failed: " + lib.smnngp_mg_last_error().decode())
                if self.emulate:
                    lib.smnngp_mg_connect_emulated(h)
            else:
                # every rank takes part in both collectives even if its own step failed, so a failure anywhere makes
                # ALL ranks fall back to the NCCL exchange together
                mine = torch.tensor(list(handle) if ok else [0] * 64, dtype=torch.uint8, device=dev)
                allh = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(allh, mine, group=group)
                flat = torch.cat(allh).cpu().numpy().tobytes()
                if ok and all(any(t.cpu().tolist()) for t in allh):
                    ok = lib.smnngp_mg_connect_ipc(h, flat) == 0
                else:
                    ok = False
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                torch.cuda.synchronize(dev)
                if int(flag.item()) == 0:
                    if h.value:
                        lib.smnngp_mg_destroy(h)
                    raise PeerUnavailable("CUDA IPC peer memory could not be set up on every rank")
            lib.smnngp_mg_set_timeout(h, float(os.environ.get("SMNNGP_PEER_TIMEOUT_S", "20")))
            if "SMNNGP_SM_RESERVE" in os.environ:
                lib.smnngp_mg_set_sm_reserve(h, int(os.environ["SMNNGP_SM_RESERVE"]))
            if "SMNNGP_RESERVE_MARGIN" in os.environ:
                lib.smnngp_mg_set_reserve_margin(h, float(os.environ["SMNNGP_RESERVE_MARGIN"]))
        self.mg = h                               # the row shard lives inside the handle

    LABELS = ("diag", "bcast", "trsm", "gather", "main_start", "update_a", "update_b", "start")

    def timeline_read(self, cap=8192):
        """[(panel, label, ms since the first mark)] recorded by the C driver after ``enable_timeline()``"""
        pan, lab, ms = (C.c_int * cap)(), (C.c_int * cap)(), (C.c_double * cap)()
        k = self.be.lib.smnngp_mg_timeline_read(self.mg, cap, pan, lab, ms)
        return [(pan[i], self.LABELS[lab[i]], ms[i]) for i in range(k)]

    def enable_timeline(self, on=True):
        self.be.lib.smnngp_mg_timeline(self.mg, 1 if on else 0)

    def close(self):
        """release the peer-visible buffers (CUDA IPC mappings + the local cudaMalloc region)"""
        if getattr(self, "_lik", None) is not None:
            self._lik.close()
            self._lik = None
        if self.mg is not None:
            with torch.cuda.device(self.a.device):
                torch.cuda.synchronize(self.a.device)
                self.be.lib.smnngp_mg_destroy(self.mg)
            self.mg = None
        if self.px is not None:
            if self.a.is_cuda:
                torch.cuda.synchronize(self.a.device)
            self.px.close()
            self.px = None

    # ---- stages ----------------------------------------------------------------------------------------------
    def _build_gram(self, x, y, hp):
        be, lay, n, db = self.be, self.lay, self.n, self.db
        tab, q, scal = be.qtable(x, self.spec, hp)
        for b in lay.local_blocks():
            g0 = b * db
            lo = lay.local_offset(b)
            rows = min(g0 + lay.block_rows(b), n) - g0                 # rows of the square part in this block
            if rows > 0:
                xb = x[g0:g0 + rows]
                if g0 > 0:                                               # rectangle left of the diagonal block
                    be.gram_block(xb, x[:g0], self.spec, hp, tab[:, g0:], tab, scal, "none", False,
                                  self.a[lo:lo + rows, :g0])
                be.gram_block(xb, xb, self.spec, hp, tab[:, g0:], tab[:, g0:], scal, self.gram_shift, True,
                              self.a[lo:lo + rows, g0:g0 + rows])
            if g0 + lay.block_rows(b) > n:                               # this block holds carried rows
                self._fill_extra_rows(b, g0, lo, x, y, hp, tab, scal)

    def _fill_extra_rows(self, b, g0, lo, x, y, hp, tab, scal):
        """the appended row y^T (global row n)"""
        n = self.n
        if g0 <= n < g0 + self.lay.block_rows(b):
            self.a[lo + (n - g0), :n].copy_(y)

    def _gather_index(self, p, c1):
        """position of global rows [c1, N) inside the padded all-gather buffer (rank-major)"""
        idx = self._gidx.get(p)
        if idx is None:                 # depends on the layout only: built once, reused by every evaluation
            db, P = self.db, self.world
            gr = torch.arange(c1, self.n, device=self.a.device, dtype=torch.int64)
            gb = gr // db
            r = gb % P
            fb = (p + 1) + ((r - (p + 1)) % P)
            idx = r * max(self.max_m, 1) + ((gb - fb) // P) * db + gr % db
            self._gidx[p] = idx
        return idx

    def _mark(self, p, label):
        if self.timeline is not None and self.a.is_cuda and (self._tl_filter is None or label in self._tl_filter):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timeline.append((p, label, ev))

    def _alloc_nccl_buffers(self):
        if self._nccl_bufs:
            return
        multi = self.world > 1
        self.send = self.be.empty(max(self.max_m, 1), self.db) if multi else None
        self.gath = self.be.empty(self.world * max(self.max_m, 1), self.db) if multi else None
        # panel in global row order, double buffered (look-ahead prepares panel p + 1 while panel p is in use)
        self.pfull = [self.be.empty(self.n, self.db) for _ in range(2)] if multi else None
        self._nccl_bufs = True

    # ---- one panel, peer-store exchange (csrc/exchange.cu) ---------------------------------------------------
    def _panel_peer(self, p, sums, info):
        """Runs on the CURRENT stream.  Returns (ls, m, arows, pfull): this rank's rows below the diagonal block,
        their solved panel part in local order (A operand of the update) and the whole panel in global order."""
        lib, lay, n, db, P, px = self.be.lib, self.lay, self.n, self.db, self.world, self.px
        s = self.be._s()
        c0, c1 = p * db, min((p + 1) * db, n)
        w = c1 - c0
        owner = lay.owner(p)
        seq = self.seq_base + p + 1
        ck = self.be._ck
        if self.rank == owner:
            # diagonal block factored in place (L_pp + the inverses of its 128-blocks), then its full inverse W is
            # assembled straight into every rank's W buffer and the W-ready flag is raised (csrc/trtri.cu)
            lo = lay.local_offset(p)
            blk = self.a[lo:lo + w, c0:c1]
            self.be.factor_diag(blk, self.linv4, sums[0:1], info, c0)
            ck(lib.smnngp_stage_assemble_inverse_f64(s, C.c_void_p(blk.data_ptr()), blk.stride(0), w,
                                                     C.c_void_p(self.linv4.data_ptr()), px.w_ptrs, P, db, px.flag_ptrs,
                                                     0, seq, C.c_void_p(self.counters.data_ptr())), "assemble_inverse")
        self._mark(p, "diag")
        if not self.emulate or self.rank == owner:           # (dry-run: nobody else is there to raise flags)
            ck(lib.smnngp_stage_wait_flags_f64(s, px.flags_local, 0, 1, seq, self.wait_timeout_s,
                                               C.c_void_p(info.data_ptr())), "wait W")
        self._mark(p, "bcast")
        if self.rank == owner:
            ls = lay.local_offset(p) + w
            m = self.mloc - ls
        else:
            ls, m = lay.rows_from_block(p + 1)
        ploc = self.ploc[p & 1]
        slot = p % 3
        r = self.a[ls:ls + max(m, 1), c0:c1]
        if self.push == "ce" and w == db and c1 < n and ls % db == 0:
            # solve with local stores only (own panel buffer + local-order copy, no flag), then the copy engines
            # push whole blocks to the other ranks and the flag follows in stream order
            ck(lib.smnngp_stage_trsm_scatter_f64(s, C.c_void_p(r.data_ptr()), r.stride(0), m, w,
                                                 C.c_void_p(px.w_local.data_ptr()), db, C.c_void_p(ploc.data_ptr()), db,
                                                 px.panel_ptrs_self[slot], P, self.rank, db, ls, c1, n, db,
                                                 px.flag_ptrs_none, 8 + self.rank, seq,
                                                 C.c_void_p(self.counters.data_ptr() + 16)), "trsm (local)")
            ck(lib.smnngp_stage_push_panel_f64(s, C.c_void_p(ploc.data_ptr()), m, w, db, P, self.rank, ls, c1, n,
                                               px.panel_ptrs[slot], px.flag_ptrs, 8 + self.rank, seq), "push_panel")
        else:
            ck(lib.smnngp_stage_trsm_scatter_f64(s, C.c_void_p(r.data_ptr()), r.stride(0), m, w,
                                                 C.c_void_p(px.w_local.data_ptr()), db, C.c_void_p(ploc.data_ptr()), db,
                                                 px.panel_ptrs[slot], P, self.rank, db, ls, c1, n, db, px.flag_ptrs,
                                                 8 + self.rank, seq, C.c_void_p(self.counters.data_ptr() + 16)),
               "trsm_scatter")
        self._mark(p, "trsm")
        bn = n // db                                                   # block of the appended row y^T
        if m > 0 and self.rank == lay.owner(bn):
            lrow = lay.local_offset(bn) + (n - bn * db)
            if lrow >= ls:
                self.zvec[c0:c1].copy_(ploc[lrow - ls, :w])           # z = L^-1 y, one panel's worth
        if c1 >= n:
            return ls, m, None, None
        first, count = (8 + self.rank, 1) if self.emulate else (8, P)
        ck(lib.smnngp_stage_wait_flags_f64(s, px.flags_local, first, count, seq, self.wait_timeout_s,
                                           C.c_void_p(info.data_ptr())), "wait panel")
        self._mark(p, "gather")
        return ls, m, ploc[:max(m, 1), :w], px.panel_local[slot][:n - c1, :w]

    # ---- one panel: diagonal block, broadcast, TRSM of the local rows, all-gather of the panel ---------------
    def _panel(self, p, sums, info, slot):
        if self.px is not None:
            return self._panel_peer(p, sums, info)
        self._alloc_nccl_buffers()
        ls, m, pfull = self._panel_nccl(p, sums, info, slot)
        arows = self.a[ls:ls + max(m, 1), p * self.db:min((p + 1) * self.db, self.n)]
        return ls, m, arows, pfull

    def _panel_nccl(self, p, sums, info, slot):
        """Runs on the CURRENT stream.  Returns (ls, m, pfull): this rank's rows below the diagonal block and the
        whole panel for global rows [c1, N) in global order (None for the last panel)."""
        be, lay, n, db, P = self.be, self.lay, self.n, self.db, self.world
        nblk = db // PB
        c0, c1 = p * db, min((p + 1) * db, n)
        w = c1 - c0
        owner = lay.owner(p)
        ldiag = self.diag[:w * w].view(w, w)
        linv = self.diag[db * db:db * db + nblk * PB * PB]
        if self.rank == owner:
            lo = lay.local_offset(p)
            blk = self.a[lo:lo + w, c0:c1]
            be.factor_diag(blk, linv, sums[0:1], info, c0)
            ldiag.copy_(blk)
        self._mark(p, "diag")
        if P > 1 and not self.emulate:
            dist.broadcast(self.diag, src=dist.get_global_rank(self.group, owner) if self.group else owner,
                           group=self.group)
        self._mark(p, "bcast")
        # this rank's rows with global index >= c1: on the owner whatever follows the diagonal rows (the rest of
        # block p, i.e. the appended row when the block straddles N, then its later blocks), elsewhere all
        # local blocks >= p + 1.  Both are suffixes of the local storage.
        if self.rank == owner:
            ls = lay.local_offset(p) + w
            m = self.mloc - ls
        else:
            ls, m = lay.rows_from_block(p + 1)
        if m > 0:
            be.trsm(self.a[ls:ls + m, c0:c1], ldiag, linv)
        self._mark(p, "trsm")
        if c1 >= n:
            return ls, m, None
        if P > 1:
            if m > 0:
                self.send[:m, :w].copy_(self.a[ls:ls + m, c0:c1])
            if self.emulate:
                self.gath.view(P, -1, self.db).copy_(self.send.unsqueeze(0).expand(P, -1, -1))
            else:
                dist.all_gather_into_tensor(self.gath, self.send, group=self.group)
            self._mark(p, "gather")
            pfull = self.pfull[slot][:n - c1]
            torch.index_select(self.gath, 0, self._gather_index(p, c1), out=pfull)
            self._mark(p, "reorder")
        else:
            pfull = self.a[ls:ls + (n - c1), c0:c1]
        return ls, m, pfull

    @_on_own_device
    def lml(self, x, y, hp, kind="student_t"):
        """SPR.loss pieces: (out[4] = {log p, loss, sum log L_ii, ||L^-1 y||^2}, info), identical on every rank."""
        be, lay, n, db, P = self.be, self.lay, self.n, self.db, self.world
        if self.mg is not None:                                    # one C call per rank (csrc/multigpu.cu)
            nh, act, arch = self.spec.ids()
            x, y = x.contiguous(), y.contiguous()
            out = be.empty(4)
            info = be.zeros(1, dtype=torch.int32)
            rc = be.lib.smnngp_lml_mg_f64(self.mg, be._s(), be._p(x), be._p(y), x.shape[1], nh, act, arch, be._p(hp),
                                          KIND[kind], SHIFT[self.gram_shift], be._p(out), be._p(info))
            if rc != 0:
                raise RuntimeError("smnngp_lml_mg_f64 failed: " + be.lib.smnngp_mg_last_error().decode())
            return out, info
        sums, info, npanels = self._factor(x, y, hp)
        # z = (L^-1 y)^T sits in global row N on its owner
        bn = n // db
        if self.rank == lay.owner(bn):
            if self.px is not None:
                be.sumsq(self.zvec, sums[1:2])
            else:
                lrow = lay.local_offset(bn) + (n - bn * db)
                be.sumsq(self.a[lrow, :n], sums[1:2])
        if self.px is not None:
            self.seq_base += npanels + 1
            if self.a.is_cuda and not self.emulate:                     # dead-peer watchdog of the Python panel loop
                ev = torch.cuda.Event()
                ev.record()
                if self._dog is None:
                    flags = torch.as_tensor(_RawCuda(self.px.local + self.px.off_flags, (32,), "<i8"), device=self.a.device)
                    self._dog = StreamWatchdog(flags, self.wait_timeout_s)
                self._dog.watch(ev, info, self.seq_base)
        if P > 1 and not self.emulate:
            dist.all_reduce(sums, group=self.group)
            dist.all_reduce(info, op=dist.ReduceOp.MAX, group=self.group)
        return be.lml_finalize(sums, hp, kind, n, info), info

    def _factor(self, x, y, hp):
        """Gram build + right-looking factorisation with one panel of look-ahead: while the bulk of panel p's trailing
        update runs on the main stream, the next panel (diagonal block, broadcast, TRSM, all-gather) is prepared on a
        side stream as soon as its block column has been updated.  Returns (sums, info, npanels); sums[0] = this rank's
        part of sum log L_ii."""
        be, lay, n, db, P = self.be, self.lay, self.n, self.db, self.world
        self._build_gram(x, y, hp)
        sums = be.zeros(2)
        info = be.zeros(1, dtype=torch.int32)
        npanels = _cdiv(n, db)
        cuda = self.a.is_cuda
        main = torch.cuda.current_stream(self.a.device) if cuda else None
        side = self.side if cuda else None

        def on_side():
            return torch.cuda.stream(side) if cuda else _NullCtx()

        if cuda:
            side.wait_stream(main)
        with on_side():
            cur = self._panel(0, sums, info, 0)
        if cuda:
            ev_panel = torch.cuda.Event()
            ev_panel.record(side)
        for p in range(npanels):
            c0, c1 = p * db, min((p + 1) * db, n)
            ls, m, arows, pfull = cur
            if cuda:
                main.wait_event(ev_panel)                                  # panel p is factored and gathered
            self._mark(p, "main_start")
            if c1 >= n:
                break
            w = c1 - c0
            gb0 = lay.first_block_from(p + 1)
            shift = gb0 * db - c1
            na = min(db, n - c1)                                           # next panel's block column first
            if m > 0:
                be.update(arows[:m], pfull[:na], self.a[ls:ls + m, c1:c1 + na], True, db, P, shift)
            self._mark(p, "update_a")
            if cuda:
                ev_a = torch.cuda.Event()
                ev_a.record(main)
                side.wait_event(ev_a)
            with on_side():                                                # look-ahead: prepare panel p + 1
                nxt = self._panel(p + 1, sums, info, (p + 1) & 1)
            if cuda:
                ev_panel = torch.cuda.Event()
                ev_panel.record(side)
            if m > 0 and c1 + na < n:                                      # the rest of the trailing matrix
                # leave SMs to the look-ahead chain only when it is long relative to this update (~1 ms of chain vs
                # 5 % of the update): estimated update time at 33 TFLOP/s below 18 ms
                t_est = 2.0 * m * (n - c1 - na) * w / 33e12
                if self.px is not None and self.sm_reserve_auto:
                    # SMs the fused panel solve needs to keep pace with this update: solve flops m w^2 at ~0.2 TF/s
                    # per SM against update flops 2 m ncols w at 33 TF/s -> 82.5 w / ncols, independent of m and P;
                    # x4.5 margin (measured: remote stores + tile quantisation make the solve ~3x slower per SM than the
                    # model, and the chain also holds the owner's diagonal block + the slowest rank) + 3
                    reserve = min(32, max(2, int(4.5 * 82.5 * w / (n - c1 - na)) + 3))
                else:
                    reserve = self.sm_reserve if t_est < self.reserve_below_s else 0
                be.update(arows[:m], pfull[na:], self.a[ls:ls + m, c1 + na:n], True, db, P, shift - na, reserve)
            self._mark(p, "update_b")
            cur = nxt
        if cuda:
            main.wait_stream(side)
        return sums, info, npanels


class DistributedGrad(DistributedLML):
    """SPR.loss and its gradient w.r.t. the six scalars (what objax.GradValues(model.loss, model.vars()) returns in the
    reference's train step, experiments/regression/train.py:62-66, before the softplus chain rule) on the ranks of a
    process group: one C call per rank (smnngp_lml_grad_mg_f64, csrc/multigpu.cu).  Needs the peer-store exchange
    (CUDA, 2..8 ranks of one node, or emulate=(1, 0) for a one-rank job)."""

    def __init__(self, n, d, spec: StackSpec, device, **kw):
        self._grad = True
        kw.setdefault("exchange", "peer")
        super().__init__(n, d, spec, device, extra_rows=int(n) + 1, **kw)
        if self.mg is None:
            raise RuntimeError("DistributedGrad needs the C multi-GPU driver (peer-store exchange)")

    @_on_own_device
    def lml_grad(self, x, y, hp, kind="student_t"):
        """(out[4] = {log p, loss, sum log L_ii, ||L^-1 y||^2}, grad[6] = d loss / d hp, info), identical on every rank"""
        be = self.be
        nh, act, arch = self.spec.ids()
        x, y = x.contiguous(), y.contiguous()
        out, grad = be.empty(4), be.empty(6)
        info = be.zeros(1, dtype=torch.int32)
        rc = be.lib.smnngp_lml_grad_mg_f64(self.mg, be._s(), be._p(x), be._p(y), x.shape[1], nh, act, arch, be._p(hp),
                                           KIND[kind], be._p(out), be._p(grad), be._p(info))
        if rc != 0:
            raise RuntimeError("smnngp_lml_grad_mg_f64 failed: " + be.lib.smnngp_mg_last_error().decode())
        return out, grad, info

    def lml(self, x, y, hp, kind="student_t"):
        out, _, info = self.lml_grad(x, y, hp, kind)
        return out, info


class DistributedPredict(DistributedLML):
    """NNGPKernel.predict (spax/kernels.py:29-32: neural_tangents gradient_descent_mse_ensemble, relative regulariser
    eps tr(K)/N) on the ranks of a process group.  The C right-hand sides Y^T and the T rows of the test-train
    cross-Gram are carried through the distributed factorisation as extra global rows n .. n+C+T-1 of the same
    block-row-cyclic layout (they are never exchanged: only square rows enter the panel all-gather), so every rank
    ends up with (L^-1 K_dt)^T for ITS test points; the C rows (L^-1 Y)^T are summed to every rank and the predictive
    tail is the single-GPU kernel on local rows."""

    def __init__(self, n, d, t, c, spec: StackSpec, device, **kw):
        self._tc = (int(t), int(c))
        super().__init__(n, d, spec, device, extra_rows=int(c) + int(t), **kw)
        self.t, self.c = int(t), int(c)
        self.gram_shift = "eps_rel"
        self._lik = None
        if self.mg is not None:                   # C driver (smnngp_predict_mg_f64): nothing else to set up
            return
        lay = self.lay
        g = [torch.arange(b * self.db, b * self.db + lay.block_rows(b)) for b in lay.local_blocks()]
        g = torch.cat(g) if g else torch.zeros(0, dtype=torch.int64)
        self.extra_lo = int((g < self.n).sum())                  # local storage keeps global order: carried rows = tail
        self.extra_idx = (g[self.extra_lo:] - self.n).to(self.a.device)     # position inside [Y^T rows | test rows]
        if self.px is not None:
            self.carried = self.be.zeros(max(int(self.extra_idx.numel()), 1), self.n)
        self._lik = None

    def _fill_extra_rows(self, b, g0, lo, x, y, hp, tab, scal):
        n, c = self.n, self.c
        g1 = g0 + self.lay.block_rows(b)
        for j in range(max(g0, n), min(g1, n + c)):             # right-hand sides: global rows n .. n+c-1
            self.a[lo + (j - g0), :n].copy_(self._Y[:, j - n])
        t0, t1 = max(g0, n + c) - (n + c), g1 - (n + c)         # test points held by this block
        if t1 > t0:
            r0 = lo + (n + c + t0 - g0)
            self.be.gram_block(self._xt[t0:t1], x, self.spec, hp, self._tab_t[:, t0:], tab, scal, "none", False,
                               self.a[r0:r0 + (t1 - t0), :n])

    def _panel(self, p, sums, info, slot):
        res = super()._panel(p, sums, info, slot)
        if self.px is not None:                                   # peer exchange: solved rows live in ploc only
            ls, m = res[0], res[1]
            k = self.extra_lo - ls
            if m > k:
                c0, c1 = p * self.db, min((p + 1) * self.db, self.n)
                self.carried[:m - k, c0:c1].copy_(self.ploc[p & 1][k:m, :c1 - c0])
        return res

    @_on_own_device
    def predict(self, x, y, x_test, hp):
        """(mean [T, C], var [T] = diag of the posterior covariance, info) - identical on every rank."""
        be, n, c, t, P = self.be, self.n, self.c, self.t, self.world
        if self.mg is not None:                                    # one C call per rank (csrc/multigpu.cu)
            nh, act, arch = self.spec.ids()
            y2 = (y if y.ndim == 2 else y[:, None]).contiguous()
            x, x_test = x.contiguous(), x_test.contiguous()
            mean, var = be.empty(t, c), be.empty(t)
            info = be.zeros(1, dtype=torch.int32)
            rc = be.lib.smnngp_predict_mg_f64(self.mg, be._s(), be._p(x), be._p(y2), be._p(x_test), x.shape[1], nh, act,
                                              arch, be._p(hp), SHIFT[self.gram_shift], be._p(mean), be._p(var),
                                              be._p(info))
            if rc != 0:
                raise RuntimeError("smnngp_predict_mg_f64 failed: " + be.lib.smnngp_mg_last_error().decode())
            return mean, var, info
        self._Y = y if y.ndim == 2 else y[:, None]
        self._xt = x_test
        self._tab_t, q_t, _ = be.qtable(x_test, self.spec, hp)
        sums, info, npanels = self._factor(x, self._Y[:, 0], hp)
        if self.px is not None:
            self.seq_base += npanels + 1
        rows = self.carried if self.px is not None else self.a[self.extra_lo:self.mloc, :n]
        idx = self.extra_idx
        z = be.zeros(c, n)
        is_rhs = idx < c
        if bool(is_rhs.any()):
            z[idx[is_rhs]] = rows[:idx.numel()][is_rhs]
        multi = P > 1 and not self.emulate
        if multi:
            dist.all_reduce(z, group=self.group)
            dist.all_reduce(info, op=dist.ReduceOp.MAX, group=self.group)
        ti = idx[~is_rhs] - c                                      # this rank's test points
        v = rows[:idx.numel()][~is_rhs].contiguous()
        mean_l, var_l = be.predict_finalize(v, z, q_t[ti].contiguous(), info)
        mean, var = be.zeros(t, c), be.zeros(t)
        mean[ti] = mean_l
        var[ti] = var_l
        if multi:
            dist.all_reduce(mean, group=self.group)
            dist.all_reduce(var, group=self.group)
        return mean, var, info

    @_on_own_device
    def test_nll(self, x, y, x_test, y_test, y_mean, y_std, hp, kind="student_t"):
        """SPR.test_nll (spax/models.py:100-120): the predictive above (relative regulariser, one right-hand side) +,
        for the Student-t likelihood, the scale d = 2a + y^T ((b/a) K + 1e-6 I)^-1 y (spax/likelihoods.py:60-61) from a
        SECOND distributed factorisation of K + 1e-6 (a/b) I, then the closed-form tail.  Returns (nll [1], mean [T],
        var [T], info), identical on every rank."""
        if self.c != 1:
            raise ValueError("test_nll needs a DistributedPredict built with c = 1")
        if self.mg is not None:                                    # one C call per rank
            be = self.be
            nh, act, arch = self.spec.ids()
            lik = None
            if kind == "student_t":
                if self._lik is None:
                    self._lik = DistributedLML(self.n, self.d, self.spec, self.a.device, group=self.group, block=self.db,
                                               backend=self.be, exchange=self.exchange,
                                               emulate=(self.world, self.rank) if self.emulate else None)
                lik = self._lik.mg
            y1 = (y if y.ndim == 1 else y[:, 0]).contiguous()
            x, x_test, y_test = x.contiguous(), x_test.contiguous(), y_test.contiguous()
            nll, mean, var = be.empty(1), be.empty(self.t), be.empty(self.t)
            info = be.zeros(1, dtype=torch.int32)
            rc = be.lib.smnngp_test_nll_mg_f64(self.mg, lik, be._s(), be._p(x), be._p(y1), be._p(x_test), be._p(y_test),
                                               x.shape[1], nh, act, arch, be._p(hp), KIND[kind], float(y_mean),
                                               float(y_std), be._p(nll), be._p(mean), be._p(var), be._p(info))
            if rc != 0:
                raise RuntimeError("smnngp_test_nll_mg_f64 failed: " + be.lib.smnngp_mg_last_error().decode())
            return nll, mean, var, info
        mean, var, info = self.predict(x, y, x_test, hp)
        quad2 = None
        if kind == "student_t":
            if self._lik is None:                                 # plain LML driver with the likelihood's jitter
                self._lik = DistributedLML(self.n, self.d, self.spec, self.a.device, group=self.group, block=self.db,
                                           backend=self.be, exchange=self.exchange,
                                           emulate=(self.world, self.rank) if self.emulate else None)
                self._lik.gram_shift = "lik"
            out2, info2 = self._lik.lml(x, y if y.ndim == 1 else y[:, 0], hp, kind="gauss")
            quad2 = out2[3:4].contiguous()                        # ||L2^-1 y||^2
            info = torch.maximum(info, info2)
        nll = self.be.test_nll_finalize(mean[:, 0].contiguous(), var, y_test, self.n, y_mean, y_std, hp, kind, quad2,
                                        info)
        return nll, mean[:, 0], var, info


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def sweep_eps(eps_list, evaluate, *, t, device, group=None):
    """The epsilon sweep of the reference's grid search (experiments/regression/find.py:20, :141: eleven regularisers
    per (w_std, b_std) point, each a new pair of factorisations of the SAME Gram matrix) as replica parallelism: the
    path does not shard below one factorisation here, so rank r evaluates eps_list[r::world] on its own GPU (its own
    cached base Gram, no data-path collective) and one all-gather of the small results makes every rank see all of them.

    evaluate(eps) -> (mean [T], var [T], logdet, quad, info) on `device` (``GridSearch.point`` with that eps).
    Returns a list aligned with eps_list of (mean [T], var [T], logdet, quad, info) device tensors / Python numbers."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_eps = len(eps_list)
    per = _cdiv(n_eps, world)
    # fixed-size slots so one all_gather_into_tensor moves everything: [mean | var | logdet quad info]
    mine = torch.zeros((per, 2 * t + 3), dtype=torch.float64, device=device)
    for k, i in enumerate(range(rank, n_eps, world)):
        mean, var, logdet, quad, info = evaluate(float(eps_list[i]))
        mine[k, :t] = mean
        mine[k, t:2 * t] = var
        mine[k, 2 * t] = logdet
        mine[k, 2 * t + 1] = quad
        mine[k, 2 * t + 2] = info.to(torch.float64).reshape(()) if torch.is_tensor(info) else float(info)
    if world > 1:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)                 # (list form: also available under gloo)
        allr = torch.stack(parts)
    else:
        allr = mine[None]
    out = []
    for i in range(n_eps):
        row = allr[i % world, i // world]
        out.append((row[:t], row[t:2 * t], float(row[2 * t]), float(row[2 * t + 1]), int(row[2 * t + 2])))
    return out
