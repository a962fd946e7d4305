"""CUDA forecast path vs golden vectors of /root/reference/model.py (torch CPU fp32) and the numpy oracle.
Bar: 1e-5 relative, measured norm-wise per row (||dy||_inf / ||y||_inf) because standardised returns cross zero."""
import numpy as np
import pytest

from conftest import sd_from_npz

pytestmark = pytest.mark.gpu
FORECAST_RTOL = 1e-5


def rowwise_rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    return float(np.max(np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)))


def build(name, g, meta=None):
    from koopman_mpc_portfolio_rebalancing_b200 import model as km
    obs = g["obs"].shape[1]
    if name == "generic_small" or name == "cfg1":
        h = [16, 16] if name == "generic_small" else [1024, 1024]
        Z = 8 if name == "generic_small" else 128
        cfg = km.model_config("GenericKM", Z, h, enc_bias=True)
    elif name == "generic_tanh_ball":
        cfg = km.model_config("GenericKM", 8, [16], enc_act="tanh", norm_fn="ball")
    elif name == "generic_gelu_mlpdec":
        cfg = km.model_config("SparseKM", 8, [16, 12], dec_layers=[10], enc_bias=True, dec_bias=True, enc_act="gelu", last_relu=True)
    elif name == "lista_linear":
        cfg = km.model_config("LISTAKM", 16, lista_loops=int(meta["loops"]), lista_L=float(meta["L"]), lista_alpha=float(meta["alpha"]), lista_linear=True)
    elif name == "lista_mlp":
        cfg = km.model_config("LISTAKM", 16, [12, 12], enc_bias=True, last_relu=True, lista_loops=int(meta["loops"]),
                              lista_L=float(meta["L"]), lista_alpha=float(meta["alpha"]), lista_linear=False)
    return km.make_model(cfg, obs)


def check_model(m, g):
    import torch
    N, d, H = int(g["N"]), int(g["d"]), int(g["H"])
    obs = torch.from_numpy(g["obs"]).cuda()
    z0 = m.encode(obs)
    assert rowwise_rel(z0.cpu().numpy(), g["z0"]) < FORECAST_RTOL
    # step-by-step API (encode / step_latent / decode), the loop of backtest.py:99-121
    z = z0
    std32 = torch.from_numpy(g["std"]).float().cuda(); mean32 = torch.from_numpy(g["mean"]).float().cuda()
    ys = []
    for _ in range(H):
        z = m.step_latent(z)
        ys.append(m.decode(z)[..., :N] * std32 + mean32)
    y_steps = torch.stack(ys, dim=1).cpu().numpy()
    assert rowwise_rel(y_steps, g["yhat"]) < FORECAST_RTOL
    # fused rollout
    pred = m.rollout(obs, H, n_cols=N)            # [H, M, N] standardised
    y_roll = (pred.permute(1, 0, 2) * std32 + mean32).cpu().numpy()
    assert rowwise_rel(y_roll, g["yhat"]) < FORECAST_RTOL
    return y_steps


@pytest.mark.parametrize("name", ["generic_small", "generic_tanh_ball", "generic_gelu_mlpdec"])
def test_generic_vs_reference(golden, name):
    g = golden(f"forecast_{name}.npz")
    m = build(name, g)
    m.load_state_dict(sd_from_npz(g))
    check_model(m, g)


@pytest.mark.parametrize("name", ["lista_linear", "lista_mlp"])
def test_lista_vs_reference(golden, name):
    g = golden(f"forecast_{name}.npz")
    m = build(name, g, golden(f"forecast_{name}_meta.npz"))
    m.load_state_dict(sd_from_npz(g))
    check_model(m, g)


def test_cfg1_vs_reference_and_strict_keys(golden):
    from koopman_mpc_portfolio_rebalancing_b200 import synthetic
    g = golden("forecast_cfg1.npz")
    m = build("cfg1", g)
    sd = synthetic.generic_km_weights(0, 200, [1024, 1024], 128)
    with pytest.raises(RuntimeError):
        m.load_state_dict({**sd, "bogus": np.zeros(1, np.float32)})
    m.load_state_dict(sd)
    check_model(m, g)


def test_config2_full_architecture_vs_reference_golden(golden):
    """BASELINE config 2 at its full architecture (1000 -> 1024 -> 1024 -> 1024 + linear decoder, the bench weights and
    the first two bench paths) against the reference model.py run on CPU in fp32 (tests/golden/forecast_cfg2.npz, made
    by make_golden.py::gen_forecast_cfg2): the default chain (fp16-pair tcgen05 GEMMs + folded [1024 x 250] multi-horizon
    read-out), the 3xTF32 chain with the folded read-out, and the unfolded step-by-step 3xTF32 chain all sit within
    1e-5 row-wise of the reference — and of the reference's own float64 model, which bounds the reference's rounding."""
    import torch
    import bench
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, data_finance as df, model as km, synthetic
    g = golden("forecast_cfg2.npz")
    w = bench.WORKLOADS["cfg2"]
    N, d, H, Z, rows = w["N"], w["d"], w["H"], w["Z"], w["rows"]
    ns = int(g["ns"])
    lr, mean, std, T = bench.make_inputs(w, int(g["paths"]), int(g["seed"]))
    # 128 more paths behind the two golden ones: full 128-row tiles and several of them, like the bench
    lr2, mean2, std2, _ = bench.make_inputs(w, 64, 777)
    lr_all = np.concatenate([lr, lr2]); mean_all = np.concatenate([mean, mean2]); std_all = np.concatenate([std, std2])
    m = km.make_model(km.model_config("GenericKM", Z, w["enc"], enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(0, N * d, w["enc"], Z))
    z = df.standardize_device(lr_all, mean_all, std_all)
    mean_d, std_d = torch.from_numpy(mean_all).cuda(), torch.from_numpy(std_all).cuda()
    L = _capi.lib()
    got = {}
    got["fp16 pairs, folded"] = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, ns, H)[:2].cpu().numpy()
    try:
        L.kmpc_set_gemm_fp16_pairs(0)
        got["3xTF32, folded"] = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, ns, H)[:2].cpu().numpy()
        L.kmpc_set_forecast_fold(0)
        got["3xTF32, step by step"] = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, ns, H)[:2].cpu().numpy()
    finally:
        L.kmpc_set_gemm_fp16_pairs(1); L.kmpc_set_forecast_fold(1)
    for name, y in got.items():
        assert y.shape == g["yhat"].shape
        r32 = max(rowwise_rel(y[b], g["yhat"][b]) for b in range(2))
        r64 = max(rowwise_rel(y[b], g["yhat_f64model"][b]) for b in range(2))
        print(f"{name}: vs reference fp32 {r32:.2e}, vs reference model in fp64 {r64:.2e}")
        assert r32 < FORECAST_RTOL and r64 < FORECAST_RTOL, (name, r32, r64)


def test_window_forecast_equals_explicit_embedding(golden):
    """forecast_series reads the delay window in place (permuted first-layer weights): must equal the forecast of
    the materialised embedding, for per-path statistics and a ragged N (padding columns)."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng(8)
    B, T, N, d, H, Z = 3, 40, 10, 6, 4, 32
    lr = rng.standard_normal((B, T, N)) * 0.012
    mean = rng.normal(3e-4, 1e-4, (B, N)); std = rng.uniform(0.008, 0.02, (B, N))
    sd = synthetic.generic_km_weights(5, N * d, [48, 40], Z)
    m = km.make_model(km.model_config("GenericKM", Z, [48, 40], enc_bias=True), N * d)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std)
    row0, t0, t1 = 3, 2, 30
    y = m.forecast_series(z, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), N, d, row0, t0, t1, H).cpu().numpy()
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    for b in range(B):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)[row0 + t0: row0 + t1]
        want = fo.forecast(emb, sd, spec, H, N, mean[b], std[b])
        assert rowwise_rel(y[b], want) < FORECAST_RTOL, b


def test_lista_window_forecast_large_latent():
    """LISTAKM at a non-trivial size (Z=256, obs=120): window path vs oracle."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng(2)
    T, N, d, H, Z = 60, 12, 10, 5, 256
    lr = rng.standard_normal((T, N)) * 0.012
    mean = rng.normal(3e-4, 1e-4, N); std = rng.uniform(0.008, 0.02, N)
    sd, L = synthetic.lista_km_weights(3, N * d, Z)
    m = km.make_model(km.model_config("LISTAKM", Z, lista_loops=10, lista_L=L, lista_alpha=5e-3, lista_linear=True), N * d)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std).unsqueeze(0)
    y = m.forecast_series(z, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), N, d, 0, 0, T - d + 1, H)[0].cpu().numpy()
    spec = fo.ModelSpec(kind="lista", linear_encoder=True, alpha=5e-3, L=L, loops=10, act="relu", last_relu=False)
    want = fo.forecast(do.time_delay_embedding(do.standardize(lr, mean, std), d), sd, spec, H, N, mean, std)
    assert rowwise_rel(y, want) < FORECAST_RTOL


def test_multi_chunk_forecast_with_partial_last_chunk():
    """More rows than one activation chunk (32768) and a partial last chunk, per-path statistics, tensor-core
    eligible widths: every row must still match the oracle (guards the chunk / group addressing)."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng(11)
    B, N, d, H, Z = 181, 12, 8, 3, 64
    rpp = 190
    T = rpp + d - 1
    lr = rng.standard_normal((B, T, N)) * 0.012
    mean = rng.normal(3e-4, 1e-4, (B, N)); std = rng.uniform(0.008, 0.02, (B, N))
    sd = synthetic.generic_km_weights(9, N * d, [128, 128], Z)
    m = km.make_model(km.model_config("GenericKM", Z, [128, 128], enc_bias=True), N * d)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std)
    y = m.forecast_series(z, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), N, d, 0, 0, rpp, H).cpu().numpy()
    assert B * rpp > 32768
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    worst = 0.0
    for b in [0, 1, 90, 171, 172, 173, 179, 180]:          # paths on both sides of the chunk boundary
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)[:rpp]
        want = fo.forecast(emb, sd, spec, H, N, mean[b], std[b])
        worst = max(worst, rowwise_rel(y[b], want))
    assert worst < FORECAST_RTOL, worst


@pytest.mark.parametrize("kind", ["generic", "lista"])
def test_folded_readout_matches_sequential_steps(kind):
    """kmpc_forecast evaluates all horizons with one GEMM against D_N (K^T)^(k+1) when step and read-out are linear
    (kmpc_set_forecast_fold, default on).  Must agree with the step-by-step chain (fold off) and with the oracle."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng(21)
    B, T, N, d, H = 4, 200, 12, 8, 5
    lr = rng.standard_normal((B, T, N)) * 0.012
    mean = rng.normal(3e-4, 1e-4, (B, N)); std = rng.uniform(0.008, 0.02, (B, N))
    if kind == "generic":
        Z = 128
        sd = synthetic.generic_km_weights(4, N * d, [128, 128], Z)
        m = km.make_model(km.model_config("GenericKM", Z, [128, 128], enc_bias=True), N * d)
        spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    else:
        Z = 256
        sd, L = synthetic.lista_km_weights(3, N * d, Z)
        m = km.make_model(km.model_config("LISTAKM", Z, lista_loops=10, lista_L=L, lista_alpha=5e-3, lista_linear=True), N * d)
        spec = fo.ModelSpec(kind="lista", linear_encoder=True, alpha=5e-3, L=L, loops=10, act="relu", last_relu=False)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std)
    mean_d, std_d = torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    rows = T - d + 1
    try:
        _capi.lib().kmpc_set_forecast_fold(0)
        y_seq = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, rows, H).cpu().numpy()
    finally:
        _capi.lib().kmpc_set_forecast_fold(1)
    y_fold = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, rows, H).cpu().numpy()
    assert not np.array_equal(y_seq, y_fold) or H == 1      # two different evaluation orders really ran
    for b in range(B):
        assert rowwise_rel(y_fold[b], y_seq[b]) < FORECAST_RTOL
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        want = fo.forecast(emb, sd, spec, H, N, mean[b], std[b])
        assert rowwise_rel(y_fold[b], want) < FORECAST_RTOL
        assert rowwise_rel(y_seq[b], want) < FORECAST_RTOL


def test_fp16_pair_chain_matches_tf32_chain_and_oracle():
    """GenericKM forecast on the fp16-pair tensor-core kernel (default) vs the 3xTF32 chain (kmpc_set_gemm_fp16_pairs(0))
    vs the oracle, tensor-core eligible widths, ragged N (padding to the 8-column fp16 row stride), per-path stats."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng(31)
    B, T, N, d, H, Z = 6, 150, 10, 12, 5, 128
    lr = rng.standard_normal((B, T, N)) * 0.012
    mean = rng.normal(3e-4, 1e-4, (B, N)); std = rng.uniform(0.008, 0.02, (B, N))
    sd = synthetic.generic_km_weights(12, N * d, [256, 128], Z)
    m = km.make_model(km.model_config("GenericKM", Z, [256, 128], enc_bias=True), N * d)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std)
    mean_d, std_d = torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    rows = T - d + 1
    y16 = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, rows, H).cpu().numpy()
    try:
        _capi.lib().kmpc_set_gemm_fp16_pairs(0)
        y32 = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, rows, H).cpu().numpy()
    finally:
        _capi.lib().kmpc_set_gemm_fp16_pairs(1)
    assert not np.array_equal(y16, y32)                      # two different kernels really ran
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    for b in range(B):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        want = fo.forecast(emb, sd, spec, H, N, mean[b], std[b])
        assert rowwise_rel(y16[b], want) < FORECAST_RTOL
        assert rowwise_rel(y32[b], want) < FORECAST_RTOL


@pytest.mark.parametrize("B,T,d,enc,act,last_relu", [
    (40, 18, 12, [256, 128], "relu", False),       # 7 rows per path: the read-out looks its statistics row up per output row
    (5, 61, 12, [256, 128], "relu", True),         # 50 rows per path: two statistics rows inside one warp's 32 rows
    (6, 150, 12, [264, 136], "relu", False),       # widths 8-aligned but not 16-aligned: ragged last column block of a pair layer
    (6, 150, 12, [256, 128], "tanh", False),       # tanh / gelu layers: staged fp16-pair output, out-of-line activation
    (6, 150, 12, [256, 128], "gelu", False),
])
def test_fp16_pair_chain_epilogue_paths(B, T, d, enc, act, last_relu):
    """Every store path of the fp16-pair GEMM epilogue (csrc/gemm_tc16.cu: branch-free interior blocks, the staged /
    transposed path for fp32 output, slow activations and ragged column blocks, both statistics look-ups of the
    de-standardising read-out) against the oracle and against the 3xTF32 chain."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng([41, B, T])
    N, H, Z = 10, 5, 128
    lr = rng.standard_normal((B, T, N)) * 0.012
    mean = rng.normal(3e-4, 1e-4, (B, N)); std = rng.uniform(0.008, 0.02, (B, N))
    sd = synthetic.generic_km_weights(13, N * d, enc, Z)
    m = km.make_model(km.model_config("GenericKM", Z, enc, enc_bias=True, enc_act=act, last_relu=last_relu), N * d)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std)
    mean_d, std_d = torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    rows = T - d + 1
    assert B * rows >= 128                                    # tensor-core eligible
    y16 = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, rows, H).cpu().numpy()
    try:
        _capi.lib().kmpc_set_gemm_fp16_pairs(0)
        y32 = m.forecast_series(z, mean_d, std_d, N, d, 0, 0, rows, H).cpu().numpy()
    finally:
        _capi.lib().kmpc_set_gemm_fp16_pairs(1)
    assert not np.array_equal(y16, y32)                      # two different kernels really ran
    spec = fo.ModelSpec(kind="generic", act=act, last_relu=last_relu, norm_fn="id", dec_act="relu")
    for b in range(B):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        want = fo.forecast(emb, sd, spec, H, N, mean[b], std[b])
        assert rowwise_rel(y16[b], want) < FORECAST_RTOL
        assert rowwise_rel(y32[b], want) < FORECAST_RTOL


def test_fp16_pair_chain_falls_back_when_values_leave_the_fp16_range():
    """an input beyond 65504 standard deviations cannot travel as an fp16 pair: the range flag sends the call to the
    3xTF32 chain, and the result still matches the oracle"""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df, model as km, synthetic
    from oracle import forecast_oracle as fo, data_oracle as do
    rng = np.random.default_rng(32)
    B, T, N, d, H, Z = 4, 80, 8, 8, 3, 64
    lr = rng.standard_normal((B, T, N)) * 0.012
    lr[1, 40, 3] = 2000.0                                       # /std 0.014 -> 1.4e5 standardised
    mean = np.full((B, N), 3e-4); std = np.full((B, N), 0.014)
    sd = synthetic.generic_km_weights(13, N * d, [128, 128], Z)
    m = km.make_model(km.model_config("GenericKM", Z, [128, 128], enc_bias=True), N * d)
    m.load_state_dict(sd)
    z = df.standardize_device(lr, mean, std)
    rows = T - d + 1
    y = m.forecast_series(z, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), N, d, 0, 0, rows, H).cpu().numpy()
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    for b in range(B):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        want = fo.forecast(emb, sd, spec, H, N, mean[b], std[b])
        assert np.all(np.isfinite(y[b]))
        assert rowwise_rel(y[b], want) < FORECAST_RTOL


@pytest.mark.parametrize("name", ["generic_small", "lista_linear"])
def test_rollout_generators_vs_reference(golden, name):
    """evaluation.rollout_{no,every_step,periodic}_reencode (evaluation.py:44-134) against the reference's outputs"""
    from koopman_mpc_portfolio_rebalancing_b200 import evaluation as ev
    g = golden(f"forecast_{name}.npz")
    r = golden("rollouts_small.npz")
    meta = golden(f"forecast_{name}_meta.npz") if name.startswith("lista") else None
    m = build(name, g, meta)
    m.load_state_dict(sd_from_npz(g))
    import torch
    x0 = torch.from_numpy(g["obs"]).cuda()
    got = {"no_reencode": ev.rollout_no_reencode(m, x0, 6), "every_step": ev.rollout_every_step_reencode(m, x0, 6),
           "periodic2": ev.rollout_periodic_reencode(m, x0, 6, 2)}
    for k, v in got.items():
        want = r[f"{name}::{k}"]
        v = v.cpu().numpy()
        assert v.shape == want.shape
        for h in range(6):
            assert rowwise_rel(v[h], want[h]) < FORECAST_RTOL, (k, h)
    with pytest.raises(ValueError):
        ev.rollout_periodic_reencode(m, x0, 3, 0)


@pytest.mark.parametrize("name", ["generic_small", "lista_linear"])
def test_evaluate_finance_vs_reference(golden, name):
    """evaluation.evaluate_finance against train.evaluate_finance of the reference (train.py:221-300) on the same model,
    initial states and seeded future: per-horizon MSE / L2 curves of every rollout mode, scalars and best mode."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import evaluation as ev
    g = golden(f"forecast_{name}.npz")
    r = golden("rollouts_small.npz")
    meta = golden(f"forecast_{name}_meta.npz") if name.startswith("lista") else None
    m = build(name, g, meta)
    m.load_state_dict(sd_from_npz(g))
    out = ev.evaluate_finance(m, torch.from_numpy(g["obs"]), torch.from_numpy(r[f"{name}::eval_future"]), max_horizon=6,
                              periodic_reencode_periods=[2, 3])
    modes = ["every_step", "no_reencode", "periodic_2", "periodic_3"]
    assert list(out["mse_curves"]) == modes and out["true"].shape == (6,) + g["obs"].shape
    for mode in modes:
        assert np.allclose(out["mse_curves"][mode].numpy(), r[f"{name}::eval_mse::{mode}"], rtol=1e-5)
        assert np.allclose(out["l2_curves"][mode].numpy(), r[f"{name}::eval_l2::{mode}"], rtol=1e-5)
        assert not out["predictions"][mode].is_cuda
    assert out["best_mode"] == modes[int(r[f"{name}::eval_best"])]
    want = r[f"{name}::eval_scalars"]
    got = [out["mean_mse_reencode"], out["mean_mse_no_reencode"], out["final_mse_reencode"], out["final_mse_no_reencode"],
           out["best_mse"]]
    assert np.allclose(got, want, rtol=1e-5)
    assert torch.equal(out["mse_reencode"], out["mse_curves"]["every_step"])
    assert torch.equal(out["pred_no_reencode"], out["predictions"]["no_reencode"])


@pytest.mark.parametrize("name", ["generic_small", "generic_tanh_ball", "lista_linear"])
def test_rollout_sequence_vs_reference(golden, name):
    """model.rollout_latent_discrete / rollout_sequence (model.py:527-585): the unroll is the plain z @ K, without the
    latent normalisation step_latent applies for NORM_FN='ball'."""
    import torch
    g = golden(f"forecast_{name}.npz")
    r = golden("sequences_small.npz")
    meta = golden(f"forecast_{name}_meta.npz") if name.startswith("lista") else None
    m = build(name, g, meta)
    m.load_state_dict(sd_from_npz(g))
    x0 = torch.from_numpy(g["obs"]).cuda()
    lat = m.rollout_latent_discrete(m.encode(x0), 4).cpu().numpy()
    seq = m.rollout_sequence(x0, 4).cpu().numpy()
    assert lat.shape == r[f"{name}::latent"].shape and seq.shape == r[f"{name}::sequence"].shape
    for k in range(5):
        assert rowwise_rel(lat[:, k], r[f"{name}::latent"][:, k]) < FORECAST_RTOL, k
        assert rowwise_rel(seq[:, k], r[f"{name}::sequence"][:, k]) < FORECAST_RTOL, k
    if name == "generic_tanh_ball":                      # step_latent normalises, the discrete unroll does not
        z1 = m.step_latent(m.encode(x0)).cpu().numpy()
        assert not np.allclose(z1, lat[:, 1], rtol=1e-3)
