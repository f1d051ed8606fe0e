"""Peer-store panel exchange (csrc/exchange.cu) at the C-ABI level on ONE device: every "peer" pointer aliases the
local buffer, which exercises the solve, the global-row placement of the block-row-cyclic layout and the flag
protocol without a second GPU.  The real 2-rank run is tests/test_distributed_gpu.py."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    import torch
    import smnngp_b200 as s
    assert torch.cuda.is_available()
    s._lib.load()
    return s


def _vp(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("w", [512, 384, 96, 130])
def test_factor_diag_inv_and_scatter(sm, w):
    import torch
    lib = sm._lib.load()
    rng = np.random.default_rng(w)
    b = rng.standard_normal((w, w + 8))
    a = b @ b.T / (w + 8) + 1e-2 * np.eye(w)
    lda = 640
    abuf = np.full((w, lda), np.nan)
    abuf[:, :w] = np.tril(a) + np.triu(np.full((w, w), 7.0), 1)      # strict upper part must never be read
    ad = torch.from_numpy(abuf).cuda()
    T = torch.empty(2 * w * w, dtype=torch.float64, device="cuda")
    linv = torch.empty(4 * 128 * 128, dtype=torch.float64, device="cuda")
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.smnngp_stage_factor_diag_inv_f64(s, _vp(ad), lda, w, _vp(T), _vp(linv), _vp(sums), _vp(info), 0) == 0
    P, ldw = 3, 512
    wbuf = torch.full((512, ldw), np.nan, dtype=torch.float64, device="cuda")
    flags = torch.zeros(16, dtype=torch.int64, device="cuda")
    counter = torch.zeros(4, dtype=torch.int32, device="cuda")
    dst = (C.c_void_p * P)(*[wbuf.data_ptr()] * P)
    fl = (C.c_void_p * P)(*[flags.data_ptr()] * P)
    ut = C.c_void_p(T.data_ptr() + w * w * 8)
    assert lib.smnngp_stage_scatter_inverse_f64(s, ut, w, w, dst, P, ldw, fl, 0, 41, _vp(counter)) == 0
    assert lib.smnngp_stage_wait_flags_f64(s, _vp(flags), 0, 1, 41, 5.0, _vp(info)) == 0
    torch.cuda.synchronize()
    L = sla.cholesky(a, lower=True)
    Winv = np.linalg.inv(L)
    got = wbuf.cpu().numpy()[:w, :w]
    assert int(info.item()) == 0 and int(flags[0].item()) == 41 and int(counter[0].item()) == 0
    assert np.all(np.triu(got, 1) == 0.0)
    assert np.abs(got - Winv).max() <= 1e-11 * np.abs(Winv).max()
    assert abs(sums[0].item() - np.log(np.diag(L)).sum()) <= 1e-11 * w
    assert np.array_equal(ad.cpu().numpy()[:, :w], abuf[:, :w])            # input block untouched


@pytest.mark.parametrize("m,w,P,rank,db,ls", [(700, 512, 2, 1, 512, 512), (1300, 256, 3, 2, 256, 256),
                                               (65, 128, 8, 5, 128, 0), (1, 96, 2, 0, 512, 1024), (0, 128, 2, 0, 128, 0)])
def test_trsm_scatter_places_rows_globally(sm, m, w, P, rank, db, ls):
    import torch
    lib = sm._lib.load()
    rng = np.random.default_rng(m + w)
    b = rng.standard_normal((w, w + 8))
    L = sla.cholesky(b @ b.T / (w + 8) + 1e-2 * np.eye(w), lower=True)
    Winv = np.tril(np.linalg.inv(L))
    R = rng.standard_normal((max(m, 1), w))
    ldr = 1024
    rbuf = np.zeros((max(m, 1), ldr))
    rbuf[:, :w] = R
    # global rows of the local rows ls .. ls+m-1 (block-row-cyclic)
    lr = ls + np.arange(m)
    g = ((lr // db) * P + rank) * db + lr % db
    c1 = int(g.min()) - 3 if m else 0
    n = int(g.max()) + 1 - (1 if m > 1 else 0) if m else 10          # the last row plays the appended y^T row
    rows_total = (int(g.max()) - c1 + 1) if m else 1
    rd, wd = torch.from_numpy(rbuf).cuda(), torch.from_numpy(np.ascontiguousarray(Winv)).cuda()
    ploc = torch.full((max(m, 1), db), np.nan, dtype=torch.float64, device="cuda")
    panel = torch.full((rows_total, db), np.nan, dtype=torch.float64, device="cuda")
    flags = torch.zeros(16, dtype=torch.int64, device="cuda")
    counter = torch.zeros(4, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    peers = (C.c_void_p * P)(*[panel.data_ptr()] * P)
    fl = (C.c_void_p * P)(*[flags.data_ptr()] * P)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.smnngp_stage_trsm_scatter_f64(s, _vp(rd), ldr, m, w, _vp(wd), w, _vp(ploc), db, peers, P, rank, db, ls, c1,
                                           n, db, fl, 8 + rank, 7, _vp(counter))
    assert rc == 0, sm._lib.load().smnngp_last_error()
    assert lib.smnngp_stage_wait_flags_f64(s, _vp(flags), 8 + rank, 1, 7, 5.0, _vp(info)) == 0
    torch.cuda.synchronize()
    assert int(info.item()) == 0 and int(flags[8 + rank].item()) == 7 and int(counter[0].item()) == 0
    if m == 0:
        return
    want = sla.solve_triangular(L, R.T, lower=True).T                 # R L^-T
    got = ploc.cpu().numpy()[:m, :w]
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    pan = panel.cpu().numpy()
    for i in range(m):
        row = pan[g[i] - c1, :w]
        if g[i] < n:
            assert np.array_equal(row, got[i]), f"local row {i} -> global {g[i]}"
        else:
            assert np.isnan(row).all()                                # rows >= n are never sent
    touched = np.zeros(rows_total, bool)
    touched[g[g < n] - c1] = True
    assert np.isnan(pan[~touched]).all()


def test_wait_flags_times_out_instead_of_hanging(sm):
    import torch
    lib = sm._lib.load()
    flags = torch.zeros(16, dtype=torch.int64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.smnngp_set_peer_wait_mode(1)                 # the spin-kernel flavour (the default stream-wait has no timeout)
    try:
        assert lib.smnngp_stage_wait_flags_f64(s, _vp(flags), 0, 2, 5, 0.05, _vp(info)) == 0
        torch.cuda.synchronize()
    finally:
        lib.smnngp_set_peer_wait_mode(0)
    assert int(info.item()) == 0x7fffffff


@pytest.mark.parametrize("m,db,P,rank,ls", [(1300, 256, 3, 2, 256), (1024, 512, 2, 1, 512), (700, 128, 8, 5, 0), (0, 128, 2, 0, 128)])
def test_push_panel_copy_engine_flavour(sm, m, db, P, rank, ls):
    """solve with local stores only (NULL peer entries are skipped, no flag), then cudaMemcpy2DAsync of whole
    distribution blocks to the 'other ranks' (here: one second buffer) at their global rows + flag."""
    import torch
    lib = sm._lib.load()
    w = db
    rng = np.random.default_rng(m + db)
    b = rng.standard_normal((w, w + 8))
    L = sla.cholesky(b @ b.T / (w + 8) + 1e-2 * np.eye(w), lower=True)
    Winv = np.tril(np.linalg.inv(L))
    R = rng.standard_normal((max(m, 1), w))
    lr = ls + np.arange(m)
    g = ((lr // db) * P + rank) * db + lr % db
    c1 = (int(g.min()) // db) * db - db if m else 0
    n = int(g.max()) if m else 10                                     # the last row plays the appended y^T row
    rows_total = (int(g.max()) - c1 + 1) if m else 1
    rd, wd = torch.from_numpy(np.ascontiguousarray(R)).cuda(), torch.from_numpy(np.ascontiguousarray(Winv)).cuda()
    ploc = torch.full((max(m, 1), db), np.nan, dtype=torch.float64, device="cuda")
    own = torch.full((rows_total, db), np.nan, dtype=torch.float64, device="cuda")
    other = torch.full((rows_total, db), np.nan, dtype=torch.float64, device="cuda")
    flags = torch.zeros(16, dtype=torch.int64, device="cuda")
    counter = torch.zeros(4, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    self_only = (C.c_void_p * P)(*[own.data_ptr() if q == rank else None for q in range(P)])
    none = (C.c_void_p * P)(*[None] * P)
    everyone = (C.c_void_p * P)(*[own.data_ptr() if q == rank else other.data_ptr() for q in range(P)])
    fl = (C.c_void_p * P)(*[flags.data_ptr()] * P)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.smnngp_stage_trsm_scatter_f64(s, _vp(rd), w, m, w, _vp(wd), w, _vp(ploc), db, self_only, P, rank, db, ls,
                                             c1, n, db, none, 8 + rank, 9, _vp(counter)) == 0
    torch.cuda.synchronize()
    assert int(flags[8 + rank].item()) == 0                           # no flag from the solve itself
    assert lib.smnngp_stage_push_panel_f64(s, _vp(ploc), m, w, db, P, rank, ls, c1, n, everyone, fl, 8 + rank, 9) == 0
    assert lib.smnngp_stage_wait_flags_f64(s, _vp(flags), 8 + rank, 1, 9, 5.0, _vp(info)) == 0
    torch.cuda.synchronize()
    assert int(flags[8 + rank].item()) == 9
    if m == 0:
        return
    want = sla.solve_triangular(L, R.T, lower=True).T
    got = ploc.cpu().numpy()[:m, :w]
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    o1, o2 = own.cpu().numpy(), other.cpu().numpy()
    sent = g < n
    assert np.array_equal(o1[g[sent] - c1], got[sent]) and np.array_equal(o2[g[sent] - c1], got[sent])
    untouched = np.ones(rows_total, bool)
    untouched[g[sent] - c1] = False
    assert np.isnan(o1[untouched]).all() and np.isnan(o2[untouched]).all()
