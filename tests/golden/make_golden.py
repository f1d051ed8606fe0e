"""Generates tests/golden/nngp_golden.npz: known answers for the hot path computed INDEPENDENTLY of the NumPy
oracle, in 50-digit mpmath arithmetic (recursion, Cholesky, solves, lgamma all in mp), rounded to float64.

The reference cannot be imported in this image (no jax / neural_tangents / objax), so these vectors pin the
oracle to the *mathematical definition* of the path (experiments/nt_kernels.py:21-31, :83-103;
spax/utils.py:178-183; spax/likelihoods.py:45-65; neural_tangents predict with relative diag_reg), not to
outputs of the reference binary - "parity unpinned" in the sense of the task statement.

    python tests/golden/make_golden.py        # rewrites nngp_golden.npz (deterministic)
"""
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
PI = mp.pi


def act_diag(u, act):
    return u / 2 if act == "relu" else (2 / PI) * mp.asin(2 * u / (1 + 2 * u))


def act_off(k, u1, u2, act):
    if act == "relu":
        s = mp.sqrt(max(u1 * u2 - k * k, mp.mpf(0)))
        th = PI / 2 if (s == 0 and k == 0) else mp.atan2(s, k)
        return s / (2 * PI) + (mp.mpf(1) / 2 - th / (2 * PI)) * k
    return (2 / PI) * mp.asin(2 * k / mp.sqrt((1 + 2 * u1) * (1 + 2 * u2)))


def gram(x1, x2, L, act, arch, w, b, v):
    n, m, d = len(x1), len(x2), len(x1[0])
    w2, b2, v2 = mp.mpf(w) ** 2, mp.mpf(b) ** 2, mp.mpf(v) ** 2
    out = mp.matrix(n, m)
    for i in range(n):
        for j in range(m):
            k = sum(mp.mpf(x1[i][t]) * mp.mpf(x2[j][t]) for t in range(d)) / d
            q1 = sum(mp.mpf(x1[i][t]) ** 2 for t in range(d)) / d
            q2 = sum(mp.mpf(x2[j][t]) ** 2 for t in range(d)) / d
            if arch == "mlp":
                for _ in range(L):
                    k, q1, q2 = w2 * k + b2, w2 * q1 + b2, w2 * q2 + b2
                    k = act_off(k, q1, q2, act)
                    q1, q2 = act_diag(q1, act), act_diag(q2, act)
            else:
                k, q1, q2 = w2 * k + b2, w2 * q1 + b2, w2 * q2 + b2
                for _ in range(L):
                    k = k + (w2 * act_off(k, q1, q2, act) + b2)
                    q1, q2 = q1 + (w2 * act_diag(q1, act) + b2), q2 + (w2 * act_diag(q2, act) + b2)
                k = act_off(k, q1, q2, act)
            out[i, j] = v2 * k
    return out


def chol_solve_stats(S, y):
    """(sum log L_ii, ||L^-1 y||^2, L) in mp."""
    L = mp.cholesky(S)
    z = mp.lu_solve(L, y)          # L is triangular: exact forward substitution up to mp precision
    return sum(mp.log(L[i, i]) for i in range(S.rows)), sum(z[i] ** 2 for i in range(S.rows)), L


def loss_mp(X, ymp, L, act, arch, w, bs, v, eps, a, b, kind):
    """SPR.loss (spax/models.py:93-98) entirely in mp; every scalar argument may be an mpf (differentiated below)."""
    n = len(X)
    K = gram(X, X, L, act, arch, w, bs, v)
    S = K + eps * mp.eye(n)
    if kind == "t":
        ld, zz, _ = chol_solve_stats((b / a) * S, ymp)
        df = 2 * a
        tt = (df + n) / 2
        lml = -tt * mp.log(1 + zz / df) - mp.mpf(n) / 2 * mp.log(df * PI) + mp.loggamma(tt) - mp.loggamma(df / 2) - ld
    else:
        ld, zz, _ = chol_solve_stats(S, ymp)
        lml = -zz / 2 - mp.mpf(n) / 2 * mp.log(2 * PI) - ld
    return -lml / n


def to_np(m):
    return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)], dtype=np.float64)


def main():
    rng = np.random.default_rng(20221018)
    n, t, d = 10, 4, 5
    x = rng.standard_normal((n, d))
    x[3] = 0.5 * x[1]                      # a collinear pair (cosine exactly 1 up to rounding)
    xt = rng.standard_normal((t, d))
    y = rng.standard_normal(n)
    yt = rng.standard_normal(t)
    y_mean, y_std = 0.3, 1.7
    eps, a, b = 1e-3, 2.5, 1.5
    out = dict(x=x, xt=xt, y=y, yt=yt, y_mean=y_mean, y_std=y_std, eps=eps, a=a, b=b)
    variants = [("relu", "mlp", 3, 1.1, 0.2, 0.9), ("erf", "mlp", 2, 1.4, 0.3, 1.0),
                ("relu", "resnet", 2, 1.0, 0.1, 1.2), ("erf", "resnet", 1, 0.8, 0.5, 0.7),
                ("relu", "mlp", 1, 1.0, 0.0, 1.0)]
    out["variants"] = np.array([f"{v[0]},{v[1]},{v[2]},{v[3]},{v[4]},{v[5]}" for v in variants])
    X, XT = x.tolist(), xt.tolist()
    ymp = mp.matrix([mp.mpf(float(v)) for v in y])
    for vi, (act, arch, L, w, bs, v) in enumerate(variants):
        K = gram(X, X, L, act, arch, w, bs, v)
        Ktd = gram(XT, X, L, act, arch, w, bs, v)
        Ktt = gram(XT, XT, L, act, arch, w, bs, v)
        out[f"K{vi}"], out[f"Ktd{vi}"] = to_np(K), to_np(Ktd)
        # --- SPR.loss: Student-t and Gaussian
        S = K + mp.mpf(eps) * mp.eye(n)
        c = mp.mpf(b) / mp.mpf(a)
        ld, zz, _ = chol_solve_stats(c * S, ymp)
        df = 2 * mp.mpf(a)
        tt = (df + n) / 2
        lml_t = -tt * mp.log(1 + zz / df) - mp.mpf(n) / 2 * mp.log(df * PI) + mp.loggamma(tt) - mp.loggamma(df / 2) - ld
        ld, zz, _ = chol_solve_stats(S, ymp)
        lml_g = -zz / 2 - mp.mpf(n) / 2 * mp.log(2 * PI) - ld
        out[f"loss_t{vi}"], out[f"loss_g{vi}"] = float(-lml_t / n), float(-lml_g / n)
        # --- d loss / d (w_std, b_std, last_w_std, eps, alpha, beta): what objax.GradValues(model.loss, vars)
        #     yields before the softplus chain rule (regression/train.py:62-66); 50-digit central differences
        base = [mp.mpf(w), mp.mpf(bs), mp.mpf(v), mp.mpf(eps), mp.mpf(a), mp.mpf(b)]
        for key in ("t", "g"):
            g = []
            for pi in range(6):
                def f(z, pi=pi, key=key):
                    q = list(base)
                    q[pi] = z
                    return loss_mp(X, ymp, L, act, arch, *q, key)
                g.append(float(mp.diff(f, base[pi])))
            out[f"grad_{key}{vi}"] = np.array(g)
        # --- predict (relative regulariser) + test_nll
        tr = sum(K[i, i] for i in range(n)) / n
        A1 = K + mp.mpf(eps) * tr * mp.eye(n)
        A1inv = mp.inverse(A1)
        mean = Ktd * (A1inv * ymp)
        cov = Ktt - Ktd * A1inv * Ktd.T
        out[f"mean{vi}"] = np.array([float(mean[i]) for i in range(t)])
        out[f"var{vi}"] = np.array([float(cov[i, i]) for i in range(t)])
        A2inv = mp.inverse(c * K + mp.mpf("1e-6") * mp.eye(n))
        dd = df + (ymp.T * A2inv * ymp)[0]
        cond_df = df + n
        lp_t, lp_g = [], []
        for i in range(t):
            xx = mp.mpf(float(yt[i])) * y_std + y_mean
            mm = mean[i] * y_std + y_mean
            cv = cov[i, i] * mp.mpf(y_std) ** 2
            sig = mp.sqrt(dd / cond_df * c * cv)
            z = (xx - mm) / sig
            lp_t.append(mp.loggamma((cond_df + 1) / 2) - mp.loggamma(cond_df / 2) - mp.log(cond_df * PI) / 2
                        - mp.log(sig) - (cond_df + 1) / 2 * mp.log(1 + z * z / cond_df))
            sg = mp.sqrt(cv)
            lp_g.append(-mp.log(2 * PI) / 2 - mp.log(sg) - ((xx - mm) / sg) ** 2 / 2)
        out[f"nll_t{vi}"], out[f"nll_g{vi}"] = float(-sum(lp_t) / t), float(-sum(lp_g) / t)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nngp_golden.npz")
    np.savez(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
