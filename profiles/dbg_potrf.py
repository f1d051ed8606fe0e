import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.linalg as sla
import smnngp_b200 as sm
lib = sm._lib.load()
for variant in (0, 2, 1, 0, 0):
    lib.smnngp_set_tile_variant(variant)
    for n in (2100, 2500, 3000, 5000):
        rng = np.random.default_rng(n)
        b = rng.standard_normal((n, n + 8)); a = b @ b.T / (n + 8) + 1e-3 * np.eye(n)
        r = rng.standard_normal((1, n)); buf = np.vstack([a, r])
        L = sla.cholesky(a, lower=True)
        errs = []
        for rep in range(3):
            ad = torch.from_numpy(buf).cuda()
            sm.device.potrf_(ad, n)
            got = ad.cpu().numpy()
            errs.append(np.abs(np.tril(got[:n]) - L).max() / np.abs(L).max())
            if rep == 0:
                d = np.abs(np.tril(got[:n]) - L); bad = np.argwhere(d > 1e-9 * np.abs(L).max())
                first = tuple(bad[0]) if len(bad) else None
        print(f"variant {variant} n={n}: errs {['%.1e' % e for e in errs]} first bad {first} nbad {len(bad)}", flush=True)
