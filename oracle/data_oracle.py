"""CPU oracle for the data side of the hot path (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.

numpy restatement of ``/root/reference/data_finance.py``:

* ``standardization_stats``  <- compute_standardization_stats  data_finance.py:211-240
* ``standardize``            <- standardize_returns            data_finance.py:243-259 (+ f32 cast :331)
* ``time_delay_embedding``   <- time_delay_embedding           data_finance.py:262-300
* ``embedding_index``        <- the gather map implied by      data_finance.py:290-298
* ``split_rows``             <- create_finance_splits masks    data_finance.py:333-351
* ``extract_current_returns`` / ``destandardize`` <- FinanceEnv data_finance.py:717-742

Pinned against the reference itself: tests/golden/make_golden.py imports /root/reference/data_finance.py
and stores its outputs (tests/golden/data_*.npz); tests/test_oracle_data.py re-checks this file against
them and repeats the reference's own assertions (tests/test_finance_data.py:129-178, 199-257).
"""
from __future__ import annotations

import numpy as np


def standardization_stats(log_returns: np.ndarray, n_train_rows: int):
    """mean / std (ddof=1, floored at 1e-8) over the first ``n_train_rows`` rows (rows with
    index <= train_end), data_finance.py:224-234.  pandas' mean/std on a float64 frame reduce each
    column with numpy's pairwise summation, which ``np.mean/np.std(axis=0)`` on the column-contiguous
    copy reproduces."""
    tr = np.asarray(log_returns, dtype=np.float64)[:n_train_rows]
    if tr.shape[0] == 0:
        raise ValueError("No training data")
    mean = np.array([np.mean(np.ascontiguousarray(tr[:, j])) for j in range(tr.shape[1])])
    std = np.array([np.std(np.ascontiguousarray(tr[:, j]), ddof=1) for j in range(tr.shape[1])])
    std = np.maximum(std, 1e-8)
    return mean, std


def standardize(log_returns: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    """(y - mean) / std in float64, then cast to float32 (data_finance.py:258, 331)."""
    z = (np.asarray(log_returns, dtype=np.float64) - mean) / std
    return z.astype(np.float32)


def embedding_index(T: int, n_assets: int, d: int) -> np.ndarray:
    """int32 [T-d+1, d*n_assets]: flat index into data.ravel() of every embedded element:
    emb[i, j*N + a] = data[i + d - 1 - j, a]  (j = 0 is the most recent day)."""
    if T < d:
        raise ValueError(f"Time series length {T} < embedding_dim {d}")
    i = np.arange(T - d + 1, dtype=np.int64)[:, None, None]
    j = np.arange(d, dtype=np.int64)[None, :, None]
    a = np.arange(n_assets, dtype=np.int64)[None, None, :]
    idx = (i + d - 1 - j) * n_assets + a
    return idx.reshape(T - d + 1, d * n_assets).astype(np.int32)


def time_delay_embedding(data: np.ndarray, d: int) -> np.ndarray:
    T, n_assets = data.shape
    idx = embedding_index(T, n_assets, d)
    return np.ascontiguousarray(data).ravel()[idx]


def split_rows(n_rows: int, n_train_days: int, n_val_days: int, d: int):
    """Row ranges of the three splits in the embedded array, given how many *raw* days fall in
    (-inf, train_end] and (train_end, val_end].  The embedding is built over the whole series first
    and embedded row i carries the date of raw day i + d - 1 (data_finance.py:334-343), so the first
    d-1 training days are lost and test rows' lags reach back into the validation period."""
    n_emb = n_rows - d + 1
    tr_end = max(0, min(n_emb, n_train_days - (d - 1)))
    va_end = max(tr_end, min(n_emb, n_train_days + n_val_days - (d - 1)))
    return (0, tr_end), (tr_end, va_end), (va_end, n_emb)


def extract_current_returns(obs: np.ndarray, n_assets: int) -> np.ndarray:
    return obs[..., :n_assets]


def destandardize(x: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    """x * std.float() + mean.float() in float32, two roundings (torch mul then add)."""
    x = np.asarray(x, dtype=np.float32)
    return (x * std.astype(np.float32)).astype(np.float32) + mean.astype(np.float32)


def gbm_log_returns(seed: int, T: int, n_assets: int, mu: float = 3e-4, one_factor: bool = False) -> np.ndarray:
    """Synthetic GBM log-returns of SURVEY.md §8(d): logret ~ N(mu - sigma^2/2, sigma^2),
    sigma ~ U[0.008, 0.02] per asset, PCG64(seed)."""
    rng = np.random.default_rng(seed)
    sigma = rng.uniform(0.008, 0.02, size=n_assets)
    eps = rng.standard_normal((T, n_assets))
    lr = (mu - 0.5 * sigma ** 2) + sigma * eps
    if one_factor:
        beta = rng.uniform(0.5, 1.5, size=n_assets)
        lr = lr + beta * (0.008 * rng.standard_normal((T, 1)))
    return lr
