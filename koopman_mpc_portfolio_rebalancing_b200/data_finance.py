"""Finance data pipeline of the hot path — names and semantics of /root/reference/data_finance.py.

In scope: standardisation statistics and transform (data_finance.py:211-259), the time-delay
embedding (262-300), chronological splits (303-353), ``FinanceDataset`` (361-419) and the ``FinanceEnv``
accessors the backtest touches (582-742).  Out of scope: yfinance download / cleaning / caching / analysis
utilities (90-192, 515-574): host-side ingestion that needs the network.  A ``FinanceEnv`` is therefore built
from a log-return array or DataFrame (``create_finance_env_from_returns``).

The embedding gather and the standardise / de-standardise transforms run on the GPU (csrc/data_kernels.cu);
the once-per-dataset statistics and the date masks are host logic, as in the reference.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _capi


@dataclass
class FinanceStats:
    """Statistics for standardization, from training data only (data_finance.py:71-82)."""
    mean: np.ndarray
    std: np.ndarray
    tickers: List[str] = field(default_factory=list)


def _dev(device):
    import torch
    return torch.device(device)


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def clean_price_data(prices, max_missing_ratio: float = 0.1, max_gap_days: int = 5):
    """Drop assets with more than ``max_missing_ratio`` missing prices, forward-fill gaps of up to ``max_gap_days``,
    drop the rows that still hold a NaN (data_finance.py:147-192).  Host pandas, like the reference."""
    missing = prices.isna().mean()
    prices = prices[missing[missing <= max_missing_ratio].index].copy()
    return prices.ffill(limit=max_gap_days).dropna()


def compute_log_returns(prices):
    """y_t = log(p_t) - log(p_{t-1}), first row dropped (data_finance.py:195-208).  DataFrame or ndarray."""
    if hasattr(prices, "iloc"):
        lp = np.log(prices)
        return lp.diff().iloc[1:]
    lp = np.log(np.asarray(prices, dtype=np.float64))
    return lp[1:] - lp[:-1]


def compute_standardization_stats(log_returns, train_end=None, n_train_rows: Optional[int] = None) -> FinanceStats:
    """mean / std(ddof=1) floored at 1e-8 over rows with index <= train_end (data_finance.py:211-240).
    ``log_returns`` may be a DataFrame (then ``train_end`` is a date string) or an ndarray with ``n_train_rows``."""
    tickers: List[str] = []
    if hasattr(log_returns, "index"):
        tickers = list(log_returns.columns)
        n_train_rows = int((log_returns.index <= train_end).sum())
        values = log_returns.values
    else:
        values = np.asarray(log_returns)
    if not n_train_rows:
        raise ValueError(f"No training data before {train_end}")
    tr = np.asarray(values, dtype=np.float64)[:n_train_rows]
    cols = [np.ascontiguousarray(tr[:, j]) for j in range(tr.shape[1])]
    mean = np.array([c.mean() for c in cols])
    std = np.maximum(np.array([c.std(ddof=1) for c in cols]), 1e-8)
    return FinanceStats(mean=mean, std=std, tickers=tickers)


def standardize_device(log_returns, mean, std, device="cuda"):
    """(y - mean)/std in fp64 then float32 (data_finance.py:258, 331) on the GPU.
    log_returns [T,N] or [B,T,N] (numpy float64 or CUDA tensor); mean/std [N] or [B,N].
    Returns a float32 CUDA tensor [.., T, pad4(N)] (padding columns are zero)."""
    import torch
    dev = _dev(device)
    y = torch.as_tensor(log_returns).to(device=dev, dtype=torch.float64).contiguous()
    squeeze = y.dim() == 2
    if squeeze:
        y = y.unsqueeze(0)
    B, T, N = y.shape
    m = torch.as_tensor(mean).to(device=dev, dtype=torch.float64).contiguous()
    s = torch.as_tensor(std).to(device=dev, dtype=torch.float64).contiguous()
    per_path = m.dim() == 2
    ld = pad4(N)
    out = torch.empty((B, T, ld), dtype=torch.float32, device=dev)
    h = _capi.Handle.get(dev.index or 0)
    _capi.check(_capi.lib().kmpc_standardize(h.ptr, _capi.ptr(y), _capi.ptr(m), _capi.ptr(s), int(per_path), B, T, N,
                                            _capi.ptr(out), ld, _capi.stream_ptr(dev.index or 0)))
    return out[0] if squeeze else out


def standardize_returns(log_returns, stats: FinanceStats):
    """z = (y - mean) / std (data_finance.py:243-259).  Returns the same kind of object as the input."""
    z = standardize_device(log_returns.values if hasattr(log_returns, "values") else log_returns, stats.mean, stats.std)
    N = len(stats.mean)
    # the reference returns float64 here and casts to float32 later (:331); the device result is that float32 array
    arr = z[..., :N].cpu().numpy()
    if hasattr(log_returns, "index"):
        import pandas as pd
        return pd.DataFrame(arr, index=log_returns.index, columns=log_returns.columns)
    return arr


def embedding_index(T: int, n_assets: int, embedding_dim: int) -> np.ndarray:
    """int32 [T-d+1, d*N] gather map of the embedding (kmpc_embed_index_host)."""
    if T < embedding_dim:
        raise ValueError(f"Time series length {T} < embedding_dim {embedding_dim}")
    idx = np.empty((T - embedding_dim + 1, embedding_dim * n_assets), dtype=np.int32)
    _capi.check(_capi.lib().kmpc_embed_index_host(T, n_assets, embedding_dim, _capi.ptr(idx)))
    return idx


def embed_device(data, embedding_dim: int, n_assets: Optional[int] = None):
    """time_delay_embedding on a float32 CUDA tensor [T,ld] or [B,T,ld] (ld >= n_assets) -> [.., T-d+1, d*N]."""
    import torch
    squeeze = data.dim() == 2
    x = data.unsqueeze(0) if squeeze else data
    x = x.contiguous()
    B, T, ld = x.shape
    N = ld if n_assets is None else n_assets
    if T < embedding_dim:
        raise ValueError(f"Time series length {T} < embedding_dim {embedding_dim}")
    out = torch.empty((B, T - embedding_dim + 1, embedding_dim * N), dtype=torch.float32, device=x.device)
    h = _capi.Handle.get(x.device.index or 0)
    _capi.check(_capi.lib().kmpc_embed_gather(h.ptr, _capi.ptr(x), ld, B, T, N, embedding_dim, _capi.ptr(out),
                                             _capi.stream_ptr(x.device.index or 0)))
    return out[0] if squeeze else out


def time_delay_embedding(data: np.ndarray, embedding_dim: int) -> np.ndarray:
    """Y_t = [y_t, y_{t-1}, ..., y_{t-d+1}] (data_finance.py:262-300).  numpy in, numpy out (dtype preserved for
    float32; other dtypes are gathered through the int32 index map so that the values stay bit-identical)."""
    data = np.asarray(data)
    T, n_assets = data.shape
    if T < embedding_dim:
        raise ValueError(f"Time series length {T} < embedding_dim {embedding_dim}")
    if data.dtype == np.float32:
        import torch
        x = torch.from_numpy(np.ascontiguousarray(data)).cuda()
        return embed_device(x, embedding_dim).cpu().numpy()
    return np.ascontiguousarray(data).ravel()[embedding_index(T, n_assets, embedding_dim)]


def verify_embedding_shift(embedded, n_assets: int, embedding_dim: int) -> bool:
    """Y_{t+1}[1:d] == Y_t[0:d-1] for every consecutive pair of embedded rows (data_finance.py:515-540), vectorised."""
    e = np.asarray(embedded.detach().cpu() if hasattr(embedded, "detach") else embedded)
    e = e.reshape(len(e), embedding_dim, n_assets)
    return bool(np.allclose(e[1:, 1:], e[:-1, :-1]))


def split_rows(n_rows: int, n_train_days: int, n_val_days: int, embedding_dim: int):
    """Row ranges (train, val, test) in the embedded array; embedded row i carries the date of raw day i+d-1
    (data_finance.py:334-343)."""
    n_emb = n_rows - embedding_dim + 1
    tr_end = max(0, min(n_emb, n_train_days - (embedding_dim - 1)))
    va_end = max(tr_end, min(n_emb, n_train_days + n_val_days - (embedding_dim - 1)))
    return (0, tr_end), (tr_end, va_end), (va_end, n_emb)


def create_finance_splits(log_returns, stats: FinanceStats, train_end: str, val_end: str, embedding_dim: int):
    """Chronological leak-free splits with time-delay embedding (data_finance.py:303-353).
    Returns (train_data, train_dates, val_data, val_dates, test_data, test_dates) as numpy / DatetimeIndex."""
    dates = log_returns.index
    z = standardize_device(log_returns.values, stats.mean, stats.std)
    emb = embed_device(z, embedding_dim, n_assets=log_returns.shape[1]).cpu().numpy()
    embedded_dates = dates[embedding_dim - 1:]
    train_mask = np.asarray(embedded_dates <= train_end)
    val_mask = np.asarray((embedded_dates > train_end) & (embedded_dates <= val_end))
    test_mask = np.asarray(embedded_dates > val_end)
    return (emb[train_mask], embedded_dates[train_mask], emb[val_mask], embedded_dates[val_mask],
            emb[test_mask], embedded_dates[test_mask])


class FinanceDataset:
    """Embedded observations of one split (data_finance.py:361-419).  ``data`` is a float32 tensor
    [n_rows, embedding_size]; ``len(ds) = n_rows - sequence_length``."""

    def __init__(self, data, dates=None, sequence_length: int = 1):
        import torch
        self.data = torch.as_tensor(data).float()
        self.dates = dates
        self.sequence_length = sequence_length
        self.n_samples = len(self.data) - sequence_length
        if self.n_samples <= 0:
            raise ValueError(f"Data length {len(self.data)} too short for sequence_length {sequence_length}")

    def __len__(self) -> int:
        return self.n_samples

    def __getitem__(self, idx: int):
        if self.sequence_length == 1:
            return self.data[idx], self.data[idx + 1]
        return self.data[idx:idx + self.sequence_length + 1]

    @property
    def observation_size(self) -> int:
        return self.data.shape[1]


class FinanceEnv:
    """Environment-like wrapper (data_finance.py:582-742): datasets, stats, the accessors the backtest uses, plus
    the device-resident standardised series the GPU path reads in place."""

    def __init__(self, train_dataset, val_dataset, test_dataset, stats: FinanceStats, metadata: Dict,
                 series_std=None, test_row0: int = 0):
        self.train_dataset, self.val_dataset, self.test_dataset = train_dataset, val_dataset, test_dataset
        self.stats = stats
        self.metadata = metadata
        self._observation_size = test_dataset.observation_size
        self._series_std = series_std       # [T, pad4(N)] float32 CUDA, whole standardised series
        self.test_row0 = test_row0          # embedded-row index of the first test row in the whole series
        self._stat_cache = None

    @property
    def observation_size(self) -> int:
        return self._observation_size

    @property
    def n_assets(self) -> int:
        return self.metadata["n_assets"]

    @property
    def embedding_dim(self) -> int:
        return self.metadata["embedding_dim"]

    def extract_current_returns(self, observations):
        """Y_t -> y_t: the first n_assets elements (data_finance.py:717-729)."""
        return observations[..., :self.n_assets]

    def destandardize_returns(self, standardized):
        """standardized * std.float() + mean.float() (data_finance.py:731-742), on the tensor's device."""
        import torch
        mean = torch.from_numpy(self.stats.mean).float().to(standardized.device)
        std = torch.from_numpy(self.stats.std).float().to(standardized.device)
        return standardized * std + mean

    def get_test_sequences(self, num_sequences: int = 100, max_length: int = 200):
        """(initial_states [S, obs], future_states [L, S, obs]) of consecutive test observations, start rows spread
        evenly over the test split (data_finance.py:672-715): the inputs of ``evaluation.evaluate_finance``.  One
        gather on the device the test split lives on instead of the reference's per-sequence stack."""
        import torch
        test_data = self.test_dataset.data
        n_samples = len(test_data)
        actual_length = min(max_length, n_samples - 1)
        actual_num_seq = min(num_sequences, n_samples - actual_length)
        if actual_num_seq <= 0:
            raise ValueError(f"Not enough test data for {num_sequences} sequences of length {max_length}")
        step = (n_samples - actual_length) // actual_num_seq
        starts = torch.arange(actual_num_seq, device=test_data.device) * step
        offs = torch.arange(1, actual_length + 1, device=test_data.device)
        return test_data[starts], test_data[(offs[:, None] + starts[None, :])]

    # ---- device views for the batch-resident path ----
    def series_device(self):
        import torch
        if self._series_std is None:
            raise RuntimeError("this FinanceEnv was built without a device series (use create_finance_env_from_returns)")
        if self._stat_cache is None:
            dev = self._series_std.device
            self._stat_cache = (torch.as_tensor(self.stats.mean, dtype=torch.float64, device=dev),
                                torch.as_tensor(self.stats.std, dtype=torch.float64, device=dev))
        return self._series_std, self._stat_cache[0], self._stat_cache[1]

    def realized_test_returns_device(self):
        """all_returns of backtest.py:169-171 for every test row, float32 CUDA [rows, N] (kmpc_current_returns)."""
        import torch
        z, mean, std = self.series_device()
        T, ld = z.shape
        rows = len(self.test_dataset.data)
        out = torch.empty((rows, self.n_assets), dtype=torch.float32, device=z.device)
        h = _capi.Handle.get(z.device.index or 0)
        _capi.check(_capi.lib().kmpc_current_returns(h.ptr, _capi.ptr(z), ld, _capi.ptr(mean), _capi.ptr(std), 0, 1, T,
                                                    self.n_assets, self.embedding_dim, self.test_row0, rows, _capi.ptr(out),
                                                    _capi.stream_ptr(z.device.index or 0)))
        return out


def create_finance_env_from_returns(log_returns, train_end=None, val_end=None, embedding_dim: int = 20,
                                    sequence_length: int = 1, n_train_days: Optional[int] = None,
                                    n_val_days: Optional[int] = None, device="cuda") -> FinanceEnv:
    """load_finance_data / create_finance_env (data_finance.py:427-507, 745-792) minus the download: builds the three
    datasets, stats and metadata from a log-return DataFrame (date split) or ndarray (row-count split)."""
    import pandas as pd
    if hasattr(log_returns, "index"):
        dates = log_returns.index
        values = np.asarray(log_returns.values, dtype=np.float64)
        n_train_days = int((dates <= train_end).sum())
        n_val_days = int(((dates > train_end) & (dates <= val_end)).sum())
        tickers = list(log_returns.columns)
    else:
        values = np.asarray(log_returns, dtype=np.float64)
        dates = pd.bdate_range("2012-01-02", periods=values.shape[0])
        tickers = [f"A{i}" for i in range(values.shape[1])]
    T, N = values.shape
    stats = compute_standardization_stats(values, n_train_rows=n_train_days)
    stats.tickers = tickers
    z = standardize_device(values, stats.mean, stats.std, device)                     # [T, pad4(N)]
    emb = embed_device(z, embedding_dim, n_assets=N)                                   # [T-d+1, d*N]
    (a0, a1), (b0, b1), (c0, c1) = split_rows(T, n_train_days, n_val_days, embedding_dim)
    ed = dates[embedding_dim - 1:]
    train_ds = FinanceDataset(emb[a0:a1], ed[a0:a1], sequence_length)
    val_ds = FinanceDataset(emb[b0:b1], ed[b0:b1], sequence_length)
    test_ds = FinanceDataset(emb[c0:c1], ed[c0:c1], sequence_length)
    metadata = {"tickers": tickers, "n_assets": N, "embedding_dim": embedding_dim,
                "observation_size": train_ds.observation_size, "train_samples": len(train_ds),
                "val_samples": len(val_ds), "test_samples": len(test_ds), "log_returns_shape": values.shape}
    return FinanceEnv(train_ds, val_ds, test_ds, stats, metadata, series_std=z, test_row0=c0)


def create_finance_env(prices, train_end: str, val_end: str, embedding_dim: int = 20, sequence_length: int = 1,
                       device="cuda") -> FinanceEnv:
    """``load_finance_data`` + ``create_finance_env`` of the reference (data_finance.py:427-507, 745-792) from a frame of
    adjusted close prices the caller already holds: clean -> log-returns -> train-only statistics -> standardise ->
    embed -> date splits.  The yfinance download in front of it (data_finance.py:90-144) is outside this library."""
    log_returns = compute_log_returns(clean_price_data(prices))
    env = create_finance_env_from_returns(log_returns, train_end=train_end, val_end=val_end, embedding_dim=embedding_dim,
                                          sequence_length=sequence_length, device=device)
    env.metadata["prices_shape"] = tuple(prices.shape)
    return env
