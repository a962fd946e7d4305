"""Seeded synthetic inputs for the hot path: GBM log-return paths and random-init
Koopman weights keyed like the reference ``state_dict``.  numpy only; used by bench.py, the tests and
tests/golden/make_golden.py so that large weights never have to be stored as fixtures.

The reference ships no data (prices come from yfinance, data_finance.py:90-144) and no checkpoint, so
every benchmark/parity input is generated here.
"""
from __future__ import annotations

import numpy as np


def gbm_log_returns(seed: int, T: int, n_assets: int, mu: float = 3e-4, one_factor: bool = False) -> np.ndarray:
    """log-returns [T, N] f64: N((mu - sigma^2/2), sigma^2), sigma ~ U[0.008, 0.02] per asset."""
    rng = np.random.default_rng(seed)
    sigma = rng.uniform(0.008, 0.02, size=n_assets)
    eps = rng.standard_normal((T, n_assets))
    lr = (mu - 0.5 * sigma ** 2) + sigma * eps
    if one_factor:
        beta = rng.uniform(0.5, 1.5, size=n_assets)
        lr = lr + beta * (0.008 * rng.standard_normal((T, 1)))
    return lr


def gbm_log_returns_batch(seed0: int, B: int, T: int, n_assets: int, mu: float = 3e-4) -> np.ndarray:
    """[B, T, N] f64, path b uses seed0 + b (vectorised: one Generator per 256 paths would change the
    stream, so this draws path by path only for small B and block-wise otherwise)."""
    if B <= 64:
        return np.stack([gbm_log_returns(seed0 + b, T, n_assets, mu) for b in range(B)])
    rng = np.random.default_rng([seed0, B, T, n_assets])
    sigma = rng.uniform(0.008, 0.02, size=(B, 1, n_assets))
    eps = rng.standard_normal((B, T, n_assets), dtype=np.float32)
    return (mu - 0.5 * sigma ** 2) + sigma * eps


def _linear(rng, out_f: int, in_f: int, bias: bool):
    """nn.Linear-like default init: U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    k = 1.0 / np.sqrt(in_f)
    w = rng.uniform(-k, k, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-k, k, size=(out_f,)).astype(np.float32) if bias else None
    return w, b


def koopman_matrix(rng, Z: int, radius: float = 0.98) -> np.ndarray:
    """K = I + 0.05 G / sqrt(Z), G ~ N(0,1) non-symmetric, rescaled to spectral radius ``radius``
    so that z @ K vs K @ z transposition bugs cannot hide (the reference default is K = I, model.py:736)."""
    G = rng.standard_normal((Z, Z))
    K = np.eye(Z) + 0.05 * G / np.sqrt(Z)
    rho = np.max(np.abs(np.linalg.eigvals(K)))
    return (K * (radius / rho)).astype(np.float32)


def generic_km_weights(seed: int, obs: int, enc_layers, Z: int, dec_layers=(), enc_bias: bool = True,
                       dec_bias: bool = False) -> dict:
    """state_dict of GenericKM/SparseKM: kmat, encoder.network.{0,2,..}.{weight,bias},
    decoder.network.{0,2,..}.{weight,bias} (model.py:701-737)."""
    rng = np.random.default_rng([seed, obs, Z])
    sd = {}
    prev = obs
    for li, h in enumerate(list(enc_layers) + [Z]):
        w, b = _linear(rng, h, prev, enc_bias)
        sd[f"encoder.network.{2 * li}.weight"] = w
        if b is not None:
            sd[f"encoder.network.{2 * li}.bias"] = b
        prev = h
    prev = Z
    for li, h in enumerate(list(dec_layers) + [obs]):
        w, b = _linear(rng, h, prev, dec_bias)
        sd[f"decoder.network.{2 * li}.weight"] = w
        if b is not None:
            sd[f"decoder.network.{2 * li}.bias"] = b
        prev = h
    sd["kmat"] = koopman_matrix(rng, Z)
    return sd


def lista_km_weights(seed: int, obs: int, Z: int, alpha: float = 5e-3, enc_layers=None, enc_bias: bool = True):
    """state_dict of LISTAKM (model.py:801-826) with a contractive S: dict ~ 0.01 randn,
    L := 1.05 * ||Wd^T Wd||_2, We = Wd^T / L, S = I - Wd^T Wd / L perturbed non-symmetrically by 1e-3.
    Returns (state_dict, L)."""
    rng = np.random.default_rng([seed, obs, Z, 7])
    Wd = (rng.standard_normal((obs, Z)) * 0.01).astype(np.float32)      # [xdim, zdim]
    gram = Wd.T.astype(np.float64) @ Wd.astype(np.float64)
    L = 1.05 * float(np.linalg.norm(gram, 2))
    sd = {"dict": np.ascontiguousarray(Wd.T), "dict_init": Wd.copy()}
    if enc_layers is None:
        sd["lista.We.weight"] = (Wd.T / L).astype(np.float32)
    else:
        prev = obs
        for li, h in enumerate(list(enc_layers) + [Z]):
            w, b = _linear(rng, h, prev, enc_bias)
            sd[f"lista.We.network.{2 * li}.weight"] = w
            if b is not None:
                sd[f"lista.We.network.{2 * li}.bias"] = b
            prev = h
    S = np.eye(Z) - gram / L + 1e-3 * rng.standard_normal((Z, Z)) / np.sqrt(Z)
    sd["lista.S"] = S.astype(np.float32)
    sd["kmat"] = koopman_matrix(rng, Z)
    return sd, L
