"""NumPy stand-in for the stage-level C-ABI - TEST INFRASTRUCTURE ONLY (never imported by the product).
It lets the multi-rank host logic of smnngp_b200.distributed (ownership maps, broadcast / all-gather / index
bookkeeping, block-row-cyclic masks) run under gloo on CPU, with the oracle as the arithmetic."""
import numpy as np
import scipy.linalg as sla
import torch

from oracle import nngp_oracle as orc


class NumpyBackend:
    def __init__(self):
        self.device = torch.device("cpu")

    def empty(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    zeros = empty

    def qtable(self, x, spec, hp):
        n = x.shape[0]
        w, b, v = (float(hp[i]) for i in range(3))
        q = torch.from_numpy(orc.nngp_diag(x.numpy(), num_hiddens=spec.num_hiddens, act=spec.act, arch=spec.arch,
                                           w_std=w, b_std=b, last_w_std=v))
        scal = torch.zeros(16, dtype=torch.float64)
        scal[0] = q.mean()                                    # tr(K) / N, as the device scalar block carries it
        return torch.zeros(1, n, dtype=torch.float64), q, scal

    def gram_block(self, x1, x2, spec, hp, tab1, tab2, scal, shift, symmetric_lower, out):
        w, b, v, eps = (float(hp[i]) for i in range(4))
        k = orc.nngp_gram(x1.numpy(), None if symmetric_lower else x2.numpy(), num_hiddens=spec.num_hiddens,
                          act=spec.act, arch=spec.arch, w_std=w, b_std=b, last_w_std=v)
        o = out.numpy()
        if symmetric_lower:
            if shift == "eps_abs":
                k = k + eps * np.eye(k.shape[0])
            elif shift == "eps_rel":                              # neural_tangents diag_reg: eps tr(K) / N
                k = k + eps * float(scal[0]) * np.eye(k.shape[0])
            elif shift == "lik":                                  # (b/a) K + 1e-6 I = (b/a) (K + 1e-6 (a/b) I)
                k = k + 1e-6 * float(hp[4]) / float(hp[5]) * np.eye(k.shape[0])
            il = np.tril_indices(k.shape[0])
            o[il] = k[il]
        else:
            o[...] = k

    def factor_diag(self, a, linv, logdet, info, gcol0):
        an = a.numpy()
        w = an.shape[0]
        full = np.tril(an) + np.tril(an, -1).T
        L = sla.cholesky(full, lower=True)
        an[np.tril_indices(w)] = L[np.tril_indices(w)]
        logdet += float(np.log(np.diag(L)).sum())
        lv = linv.numpy().reshape(-1, 128, 128)
        for k in range((w + 127) // 128):
            j0, j1 = k * 128, min((k + 1) * 128, w)
            blk = np.eye(128)
            blk[: j1 - j0, : j1 - j0] = sla.solve_triangular(L[j0:j1, j0:j1], np.eye(j1 - j0), lower=True)
            lv[k] = blk

    def trsm(self, r, ldiag, linv):
        rn, L = r.numpy(), np.tril(ldiag.numpy())
        rn[...] = sla.solve_triangular(L, rn.T, lower=True).T

    def update(self, a, b, c, lower, cyc_db, cyc_p, base_shift, sm_reserve=0):
        an, bn, cn = a.numpy(), b.numpy(), c.numpy()
        full = an @ bn.T
        rows = np.arange(cn.shape[0])
        if lower:
            lim = rows + (base_shift + (rows // cyc_db) * (cyc_p - 1) * cyc_db if cyc_db else 0)
            mask = np.arange(cn.shape[1])[None, :] <= lim[:, None]
            cn[mask] -= full[mask]
        else:
            cn -= full

    def sumsq(self, z, out):
        out[0] = float((z.numpy() ** 2).sum())

    def predict_finalize(self, v, z, ktt, info):
        vn, zn = v.numpy(), z.numpy()
        mean = torch.from_numpy(vn @ zn.T)
        var = torch.from_numpy(ktt.numpy() - (vn * vn).sum(axis=1))
        if int(info[0]) != 0:
            mean, var = mean * float("nan"), var * float("nan")
        return mean, var

    def test_nll_finalize(self, mean, var, ytest, n, y_mean, y_std, hp, kind, quad2, info):
        a, b = float(hp[4]), float(hp[5])
        xx = ytest.numpy() * y_std + y_mean
        mm = mean.numpy() * y_std + y_mean
        cv = var.numpy() * y_std ** 2
        if kind == "student_t":
            df = 2 * a
            cond_df = df + n
            d = df + (a / b) * float(quad2[0])
            lp = orc._t_logpdf(xx, cond_df, mm, np.sqrt(d / cond_df * b / a * cv))
        else:
            sg = np.sqrt(cv)
            lp = -0.5 * np.log(2 * np.pi) - np.log(sg) - 0.5 * ((xx - mm) / sg) ** 2
        nll = -float(np.mean(lp)) if int(info[0]) == 0 else float("nan")
        return torch.tensor([nll], dtype=torch.float64)

    def lml_finalize(self, sums, hp, kind, n, info):
        from scipy.special import gammaln
        logdet, zz = float(sums[0]), float(sums[1])
        a, b = float(hp[4]), float(hp[5])
        if kind == "student_t":
            c, df = b / a, 2 * a
            t = 0.5 * (df + n)
            lml = (-t * np.log(1 + (zz / c) / df) - n / 2 * np.log(df * np.pi) + gammaln(t) - gammaln(df / 2)
                   - (logdet + 0.5 * n * np.log(c)))
        else:
            lml = -0.5 * zz - n / 2 * np.log(2 * np.pi) - logdet
        if int(info[0]) != 0:
            lml = float("nan")
        return torch.tensor([lml, -lml / n, logdet, zz], dtype=torch.float64)
