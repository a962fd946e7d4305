"""The drop-in boundary on CPU: the C-ABI library loads and exports every symbol include/kmpc.h declares, the
product never routes through oracle/, and it fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "koopman_mpc_portfolio_rebalancing_b200")


def header_functions():
    src = open(os.path.join(ROOT, "include", "kmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "run `python -m koopman_mpc_portfolio_rebalancing_b200.build` first"
    L = ctypes.CDLL(_capi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/kmpc.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), set(names) ^ set(_capi.SIGNATURES)
    assert _capi.lib().kmpc_version() == 100


def test_embed_index_host_matches_oracle_and_rejects_short_series():
    import numpy as np
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, data_finance
    from oracle import data_oracle
    idx = data_finance.embedding_index(12, 3, 4)
    assert idx.dtype == np.int32 and np.array_equal(idx, data_oracle.embedding_index(12, 3, 4))
    with pytest.raises(ValueError):
        data_finance.embedding_index(3, 2, 5)
    buf = np.zeros(4, np.int32)
    assert _capi.lib().kmpc_embed_index_host(3, 2, 5, _capi.ptr(buf)) == -1
    assert b"embedding_dim" in _capi.lib().kmpc_last_error()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle\b", txt, flags=re.M), fn
                assert not re.search(r"import_module\(.*oracle|CDLL\(.*oracle|#include\s+\".*oracle", txt), fn


def test_variant_table():
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    L = _capi.lib()
    assert L.kmpc_mpc_supported(5, 50) == 1 and L.kmpc_mpc_supported(5, 10) == 1 and L.kmpc_mpc_supported(3, 64) == 1
    assert L.kmpc_mpc_supported(5, 5000) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, mpc
    with pytest.raises(_capi.KmpcError):
        mpc.solve_mpc_log_utility(np.ones(3) / 3, np.zeros((2, 3), np.float32), mpc.MPCConfig(horizon=2))


def test_split_rows_and_stats_host_logic(golden):
    import numpy as np
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    g = golden("data_small.npz")
    st = df.compute_standardization_stats(g["log_returns"], n_train_rows=int(g["n_train_days"]))
    assert np.array_equal(st.mean, g["mean"]) and np.array_equal(st.std, g["std"])
    (a0, a1), (b0, b1), (c0, c1) = df.split_rows(g["log_returns"].shape[0], int(g["n_train_days"]), int(g["n_val_days"]), int(g["d"]))
    assert (a1 - a0, b1 - b0, c1 - c0) == (g["train"].shape[0], g["val"].shape[0], g["test"].shape[0])


def test_shard_range_covers_everything():
    from koopman_mpc_portfolio_rebalancing_b200.engine import shard_range
    for n, w in [(10, 1), (10, 3), (4096, 8), (5, 8), (65536, 4)]:
        seen = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_result_frames_and_parquet(tmp_path):
    import numpy as np
    """result formats of the batched engine: reference column names, parquet round trip (host-only logic)"""
    import pandas as pd
    from koopman_mpc_portfolio_rebalancing_b200 import engine, backtest as bt
    rng = np.random.default_rng(0)
    met = rng.standard_normal((7, 5)); hist = rng.standard_normal((7, 4, 4))
    mf = engine.metrics_frame(met)
    assert list(mf.columns) == ["Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return"] == list(bt.METRIC_KEYS)
    hf = engine.history_frames(hist, dates=list(pd.bdate_range("2020-01-01", periods=8)), rebalance_freq=2)
    assert len(hf) == 7 and list(hf[0].columns) == ["date", "portfolio_value", "return", "turnover", "cost"]
    assert hf[0]["date"].iloc[1] == pd.bdate_range("2020-01-01", periods=8)[2]
    p = engine.write_results(str(tmp_path / "out.parquet"), met, hist, ids=np.arange(10, 17))
    back = pd.read_parquet(p)
    assert np.array_equal(back["backtest"].values, np.arange(10, 17)) and np.allclose(back[list(bt.METRIC_KEYS)].values, met)
    long = pd.read_parquet(str(tmp_path / "out.history.parquet"))
    assert len(long) == 28 and np.allclose(long[long.backtest == 12][list(bt.HISTORY_COLS)].values, hist[2])


def test_test_sequences_and_shift_check_vs_reference(golden):
    """FinanceEnv.get_test_sequences (data_finance.py:672-715: the feeder of evaluate_finance) and
    verify_embedding_shift (:515-540) on host tensors, against the reference's own outputs for the data_small env."""
    import numpy as np
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    g = golden("data_small.npz")
    r = golden("test_sequences_small.npz")
    ds = lambda a: df.FinanceDataset(a, None, 1)
    env = df.FinanceEnv(ds(g["train"]), ds(g["val"]), ds(g["test"]), df.FinanceStats(g["mean"], g["std"], []),
                        {"n_assets": 3, "embedding_dim": int(g["d"])})
    i1, f1 = env.get_test_sequences(num_sequences=5, max_length=7)
    assert np.array_equal(i1.numpy(), r["init_5_7"]) and np.array_equal(f1.numpy(), r["future_5_7"])
    i2, f2 = env.get_test_sequences()
    assert np.array_equal(i2.numpy(), r["init_default"]) and np.array_equal(f2.numpy(), r["future_default"])
    tiny = df.FinanceEnv(ds(g["train"]), ds(g["val"]), ds(g["test"][:2]), env.stats, env.metadata)
    i3, f3 = tiny.get_test_sequences(3, 5)                  # clamps to what the split holds
    assert i3.shape == (1, 12) and f3.shape == (1, 1, 12)
    assert df.verify_embedding_shift(g["embedded"], 3, int(g["d"])) == bool(r["shift_ok"]) is True
    assert df.verify_embedding_shift(g["embedded"][::2], 3, int(g["d"])) == bool(r["shift_broken"]) is False


def _price_frame(g):
    import pandas as pd
    return pd.DataFrame(g["prices"], index=pd.bdate_range("2015-01-05", periods=g["prices"].shape[0]),
                        columns=[f"P{i}" for i in range(g["prices"].shape[1])])


def test_clean_prices_and_log_returns_vs_reference(golden):
    """clean_price_data / compute_log_returns (data_finance.py:147-208) bit-exact against the reference on a price frame
    with short gaps (filled), a long gap and a leading NaN (rows dropped) and a sparse asset (dropped)."""
    import numpy as np
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    g = golden("prices_small.npz")
    prices = _price_frame(g)
    clean = df.clean_price_data(prices)
    assert np.array_equal(clean.values, g["clean"])
    assert [prices.index.get_loc(i) for i in clean.index] == list(g["clean_rows"])
    assert [list(prices.columns).index(c) for c in clean.columns] == list(g["clean_cols"])
    assert np.array_equal(df.compute_log_returns(clean).values, g["log_returns"])
    st = df.compute_standardization_stats(df.compute_log_returns(clean), str(g["train_end"]))
    assert np.array_equal(st.mean, g["mean"]) and np.array_equal(st.std, g["std"])


def test_bootstrap_indices_are_shard_reproducible_and_uniform():
    import numpy as np
    """config 5's path generator: counter-based, so a rank regenerates exactly its shard; splitmix64 checked against a
    pure-Python evaluation; indices uniform over the historical block"""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import engine
    a = engine.bootstrap_indices(9, 40, 3000, 1234)
    b = engine.bootstrap_indices(4, 40, 3000, 1234, offset=5)
    assert torch.equal(a[5:], b) and a.dtype == torch.int64 and int(a.min()) >= 0 and int(a.max()) < 3000
    M = (1 << 64) - 1

    def ref(seed, p, t, n):
        x = ((p * 0x100000001B3 + t) * 0x9E3779B97F4A7C15 + ((seed * 2 + 1) * 0x632BE59BD9B4E019 % (1 << 63))) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        x ^= x >> 31
        return (x >> 1) % n
    for (p, t) in [(0, 0), (5, 17), (8, 39)]:
        assert int(a[p, t]) == ref(1234, p, t, 3000)
    big = engine.bootstrap_indices(4000, 271, 3000, 7).numpy().ravel()
    counts = np.bincount(big, minlength=3000)
    assert abs(counts.std() / np.sqrt(len(big) / 3000) - 1.0) < 0.1
    assert not torch.equal(engine.bootstrap_indices(2, 40, 3000, 1), engine.bootstrap_indices(2, 40, 3000, 2))
