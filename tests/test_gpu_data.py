"""Data kernels vs the golden vectors produced by /root/reference/data_finance.py: bit-exact embedding, splits,
standardisation and current-return extraction."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_standardize_and_embedding_bit_exact(golden):
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    g = golden("data_small.npz")
    N, d = g["log_returns"].shape[1], int(g["d"])
    z = df.standardize_device(g["log_returns"], g["mean"], g["std"])
    assert z.shape[1] == df.pad4(N) and torch.all(z[:, N:] == 0)
    assert np.array_equal(z[:, :N].cpu().numpy(), g["standardized"])
    emb = df.embed_device(z, d, n_assets=N).cpu().numpy()
    assert emb.dtype == np.float32 and np.array_equal(emb, g["embedded"])
    # numpy front-end (time_delay_embedding signature of the reference), float32 and float64
    assert np.array_equal(df.time_delay_embedding(g["standardized"], d), g["embedded"])
    x64 = np.arange(20, dtype=np.float64).reshape(10, 2)
    e64 = df.time_delay_embedding(x64, 3)
    assert e64.dtype == np.float64 and np.array_equal(e64[0], np.concatenate([x64[2], x64[1], x64[0]]))
    with pytest.raises(ValueError):
        df.time_delay_embedding(np.zeros((3, 2), np.float32), 5)


def test_env_splits_and_accessors_vs_reference(golden):
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    from oracle import data_oracle as do
    g = golden("data_small.npz")
    env = df.create_finance_env_from_returns(g["log_returns"], embedding_dim=int(g["d"]), n_train_days=int(g["n_train_days"]),
                                             n_val_days=int(g["n_val_days"]))
    assert np.array_equal(env.train_dataset.data.cpu().numpy(), g["train"])
    assert np.array_equal(env.val_dataset.data.cpu().numpy(), g["val"])
    assert np.array_equal(env.test_dataset.data.cpu().numpy(), g["test"])
    assert len(env.test_dataset) == int(g["test_len"])
    want = do.destandardize(do.extract_current_returns(g["test"], env.n_assets), g["mean"], g["std"])
    got = env.realized_test_returns_device().cpu().numpy()
    assert np.array_equal(got, want)
    got2 = env.destandardize_returns(env.extract_current_returns(env.test_dataset.data)).cpu().numpy()
    assert np.array_equal(got2, want)


def test_batched_embedding_large_ragged():
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    from oracle import data_oracle as do
    rng = np.random.default_rng(0)
    B, T, N, d = 3, 57, 7, 5
    lr = rng.standard_normal((B, T, N)) * 0.01
    mean = rng.normal(0, 1e-3, (B, N)); std = rng.uniform(0.005, 0.02, (B, N))
    z = df.standardize_device(lr, mean, std)
    for b in range(B):
        zb = do.standardize(lr[b], mean[b], std[b])
        assert np.array_equal(z[b, :, :N].cpu().numpy(), zb)
    emb = df.embed_device(z, d, n_assets=N).cpu().numpy()
    for b in range(B):
        assert np.array_equal(emb[b], do.time_delay_embedding(z[b, :, :N].cpu().numpy(), d))


def test_env_from_price_frame_vs_reference(golden):
    """create_finance_env(prices, ...): clean -> log-returns -> train-only statistics -> standardise -> embed -> date
    splits, bit-exact against the reference pipeline (data_finance.py:147-353) on the prices_small frame."""
    from test_boundary import _price_frame
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df
    g = golden("prices_small.npz")
    env = df.create_finance_env(_price_frame(g), str(g["train_end"]), str(g["val_end"]), embedding_dim=int(g["d"]))
    assert env.n_assets == 3 and env.metadata["tickers"] == ["P0", "P1", "P2"] and env.metadata["prices_shape"] == (90, 4)
    assert np.array_equal(env.train_dataset.data.cpu().numpy(), g["train"])
    assert np.array_equal(env.val_dataset.data.cpu().numpy(), g["val"])
    assert np.array_equal(env.test_dataset.data.cpu().numpy(), g["test"])
    assert df.verify_embedding_shift(env.test_dataset.data, 3, int(g["d"]))
