"""Oracle forecast (numpy fp32) vs golden vectors from /root/reference/model.py (torch CPU fp32).
Tolerance: 1e-5 norm-wise per row (||dy||_inf / ||y||_inf), the north-star bar for forecasts."""
import numpy as np
import pytest

from conftest import sd_from_npz
from oracle import forecast_oracle as fo
from koopman_mpc_portfolio_rebalancing_b200 import synthetic

FORECAST_RTOL = 1e-5


def rowwise_rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    return float(np.max(np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)))


SPECS = {
    "generic_small": fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu"),
    "generic_tanh_ball": fo.ModelSpec(kind="generic", act="tanh", last_relu=False, norm_fn="ball", dec_act="relu"),
    "generic_gelu_mlpdec": fo.ModelSpec(kind="generic", act="gelu", last_relu=True, norm_fn="id", dec_act="relu"),
}


@pytest.mark.parametrize("name", list(SPECS))
def test_generic_forecast_vs_reference(golden, name):
    g = golden(f"forecast_{name}.npz")
    sd = sd_from_npz(g)
    z0 = fo.encode(g["obs"], sd, SPECS[name])
    assert rowwise_rel(z0, g["z0"]) < FORECAST_RTOL
    y = fo.forecast(g["obs"], sd, SPECS[name], int(g["H"]), int(g["N"]), g["mean"], g["std"])
    assert y.shape == g["yhat"].shape and y.dtype == np.float32
    assert rowwise_rel(y, g["yhat"]) < FORECAST_RTOL


@pytest.mark.parametrize("name,linear", [("lista_linear", True), ("lista_mlp", False)])
def test_lista_forecast_vs_reference(golden, name, linear):
    g = golden(f"forecast_{name}.npz")
    m = golden(f"forecast_{name}_meta.npz")
    spec = fo.ModelSpec(kind="lista", linear_encoder=linear, alpha=float(m["alpha"]), L=float(m["L"]),
                        loops=int(m["loops"]), act="relu", last_relu=True)
    sd = sd_from_npz(g)
    assert rowwise_rel(fo.encode(g["obs"], sd, spec), g["z0"]) < FORECAST_RTOL
    y = fo.forecast(g["obs"], sd, spec, int(g["H"]), int(g["N"]), g["mean"], g["std"])
    assert rowwise_rel(y, g["yhat"]) < FORECAST_RTOL


def test_cfg1_forecast_from_seeded_weights(golden):
    g = golden("forecast_cfg1.npz")
    sd = synthetic.generic_km_weights(0, 200, [1024, 1024], 128)
    y = fo.forecast(g["obs"], sd, SPECS["generic_small"], 5, 10, g["mean"], g["std"])
    assert rowwise_rel(y, g["yhat"]) < FORECAST_RTOL


def test_shrink_known_answers():
    # reference tests/test_model.py:20-42
    x = np.array([-2.0, -0.5, 0.0, 0.5, 2.0], dtype=np.float32)
    assert np.allclose(fo.shrink(x, 1.0), [-1.0, 0.0, 0.0, 0.0, 1.0])
    assert np.allclose(fo.shrink(x, 0.0), x)
