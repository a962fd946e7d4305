"""Oracle (numpy) vs the golden vectors produced by /root/reference/data_finance.py, plus the reference's
own embedding/split assertions (reference tests/test_finance_data.py:129-178, 199-257)."""
import numpy as np
import pytest

from oracle import data_oracle as do


def test_stats_and_standardize_bit_exact(golden):
    g = golden("data_small.npz")
    mean, std = do.standardization_stats(g["log_returns"], int(g["n_train_days"]))
    assert np.array_equal(mean, g["mean"])
    assert np.array_equal(std, g["std"])
    assert np.array_equal(do.standardize(g["log_returns"], mean, std), g["standardized"])


def test_embedding_bit_exact_vs_reference(golden):
    g = golden("data_small.npz")
    emb = do.time_delay_embedding(g["standardized"], int(g["d"]))
    assert emb.dtype == np.float32
    assert np.array_equal(emb, g["embedded"])


def test_splits_bit_exact_vs_reference(golden):
    g = golden("data_small.npz")
    d = int(g["d"])
    emb = do.time_delay_embedding(g["standardized"], d)
    (a0, a1), (b0, b1), (c0, c1) = do.split_rows(g["log_returns"].shape[0], int(g["n_train_days"]), int(g["n_val_days"]), d)
    assert np.array_equal(emb[a0:a1], g["train"])
    assert np.array_equal(emb[b0:b1], g["val"])
    assert np.array_equal(emb[c0:c1], g["test"])
    assert a1 == b0 and b1 == c0 and c1 == emb.shape[0]          # no overlap, no gap
    assert int(g["test_len"]) == (c1 - c0) - 1                    # len(ds) = rows - sequence_length


def test_embedding_shape_and_content_like_reference_tests():
    data = np.random.default_rng(0).standard_normal((100, 5)).astype(np.float32)
    assert do.time_delay_embedding(data, 5).shape == (96, 25)
    data = np.arange(20).reshape(10, 2).astype(np.float32)
    emb = do.time_delay_embedding(data, 3)
    assert np.array_equal(emb[0], np.concatenate([data[2], data[1], data[0]]))
    assert np.array_equal(emb[1], np.concatenate([data[3], data[2], data[1]]))
    # shift property: Y_{t+1}[N:] == Y_t[:-N]
    data = np.random.default_rng(1).standard_normal((50, 3)).astype(np.float32)
    emb = do.time_delay_embedding(data, 4)
    assert np.array_equal(emb[1:, 3:], emb[:-1, :-3])
    with pytest.raises(ValueError):
        do.time_delay_embedding(np.zeros((3, 2), np.float32), 5)


def test_embedding_index_matches_gather():
    idx = do.embedding_index(12, 3, 4)
    assert idx.dtype == np.int32 and idx.shape == (9, 12)
    assert idx[0, 0] == 3 * 3 and idx[0, -1] == 2 and idx[-1, 0] == 11 * 3


def test_destandardize_two_roundings():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 5)).astype(np.float32)
    mean, std = rng.normal(0, 1e-3, 5), rng.uniform(0.01, 0.02, 5)
    want = (x * std.astype(np.float32)).astype(np.float32) + mean.astype(np.float32)
    assert np.array_equal(do.destandardize(x, mean, std), want)
