"""Seeded synthetic inputs shaped like the BASELINE configs.  Shared by tests, bench
and smoke so every arm sees the same numbers."""
import numpy as np

DEFAULT_HP = dict(w_std=1.0, b_std=1e-8, last_w_std=1.0, eps=1e-6, alpha=2.0, beta=2.0)  # regression/train.py:37-45


def regression_data(n, d, t=0, seed=10):
    """UCI/Boston-shaped: X ~ N(0,1) z-scored per column on the train part (data.py:266-276), y = sin(Xw) + noise,
    z-scored (data.py:278-284)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n + t, d))
    w = rng.standard_normal(d) / np.sqrt(d)
    y = np.sin(x @ w) + 0.1 * rng.standard_normal(n + t)
    if n < 2:                       # degenerate sizes (edge-case tests): nothing to standardise
        return (np.ascontiguousarray(x[:n]), np.ascontiguousarray(y[:n]), np.ascontiguousarray(x[n:]),
                np.ascontiguousarray(y[n:]), 0.0, 1.0)
    xm, xs = x[:n].mean(0), x[:n].std(0)
    x = (x - xm) / xs
    ym, ys = y[:n].mean(), y[:n].std()
    y = (y - ym) / ys
    return (np.ascontiguousarray(x[:n]), np.ascontiguousarray(y[:n]), np.ascontiguousarray(x[n:]),
            np.ascontiguousarray(y[n:]), float(ym), float(ys))


def pixel_data(n, d, seed=10, t=0):
    """MNIST-shaped: 80 % zeros, U(0,1) elsewhere, then (x - 0.5) / 0.5 (classification/data.py:134-144)."""
    rng = np.random.default_rng(seed)
    x = rng.random((n + t, d)) * (rng.random((n + t, d)) < 0.2)
    x = (x - 0.5) / 0.5
    y = rng.standard_normal(n + t)
    return np.ascontiguousarray(x[:n]), np.ascontiguousarray(y[:n]), np.ascontiguousarray(x[n:]), np.ascontiguousarray(y[n:])
