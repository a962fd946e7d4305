"""DMDStrategy (reference baselines.py:109-187), host part: the operator fit against the golden K of the unmodified
reference (tests/golden/dmd_small.npz, written by tests/golden/make_golden.py)."""
import numpy as np


def test_dmd_fit_matches_reference(golden):
    from koopman_mpc_portfolio_rebalancing_b200 import baselines, synthetic
    from oracle import data_oracle as do
    g = golden("dmd_small.npz")
    T, N, d = int(g["T"]), int(g["N"]), int(g["d"])
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), T, N)
    z = do.standardize(lr, g["mean"], g["std"])
    emb = do.time_delay_embedding(z, d)
    (a0, a1), _, _ = do.split_rows(T, int(g["n_train_days"]), int(g["n_val_days"]), d)
    K = baselines.DMDStrategy._fit_dmd(emb[a0:a1])
    assert K.dtype == np.float32 and K.shape == g["K"].shape          # the reference fits in the dataset's float32
    assert np.abs(K - g["K"]).max() <= 2e-5 * np.abs(g["K"]).max()
