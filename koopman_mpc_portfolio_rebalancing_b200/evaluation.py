"""Batched rollout generators of the reference (/root/reference/evaluation.py:44-134) on the device kernels.

``rollout_no_reencode`` is the batched twin of the strategy's forecast loop (encode once, then H x step_latent +
decode) and runs as one fused chain (``kmpc_rollout``); the re-encoding variants compose ``encode`` /
``step_latent`` / ``decode`` (each one launch of the GEMM kernels over the whole batch).  Like the reference, a
non-finite prediction marks the remaining steps as NaN.  ``evaluate_finance`` is the multi-horizon scoring of
train.py:221-300 over those rollouts (per-horizon MSE / L2 curves of every mode, best mode): what a sweep uses to pick
a model on the device.  Plots and the ODE-system evaluation driver of evaluation.py are outside the hot path.
"""
from __future__ import annotations


def _finish(preds, horizon):
    import torch
    out = torch.stack(preds, dim=0)
    if len(preds) < horizon:
        pad = torch.full((horizon - len(preds),) + tuple(out.shape[1:]), float("nan"), dtype=out.dtype, device=out.device)
        out = torch.cat([out, pad], dim=0)
    return out


def rollout_no_reencode(model, x0, horizon: int):
    """[horizon, batch, state_dim] (evaluation.py:44-74)."""
    import torch
    out = model.rollout(x0, horizon)                       # fused: encode, H x (z K, decode)
    finite = torch.isfinite(out).reshape(horizon, -1).all(dim=1)
    bad = torch.nonzero(~finite)
    if bad.numel():
        out[int(bad[0]) + 1:] = float("nan")               # the exploding step itself is kept, like the reference
    return out


def rollout_every_step_reencode(model, x0, horizon: int):
    """state <- decode(step_latent(encode(state))) at every step (evaluation.py:77-99, model.step_env)."""
    import torch
    state = torch.as_tensor(x0).to(model.device, dtype=torch.float32)
    preds = []
    for _ in range(horizon):
        state = model.decode(model.step_latent(model.encode(state)))
        preds.append(state)
        if not bool(torch.isfinite(state).all()):
            break
    return _finish(preds, horizon)


def rollout_periodic_reencode(model, x0, horizon: int, period: int):
    """re-encode the prediction every ``period`` steps (evaluation.py:102-134)."""
    import torch
    if period <= 0:
        raise ValueError("period must be a positive integer")
    latent = model.encode(torch.as_tensor(x0).to(model.device, dtype=torch.float32))
    preds = []
    for step in range(horizon):
        latent = model.step_latent(latent)
        x_pred = model.decode(latent)
        preds.append(x_pred)
        if not bool(torch.isfinite(x_pred).all()):
            break
        if (step + 1) % period == 0:
            latent = model.encode(x_pred)
    return _finish(preds, horizon)


def evaluate_finance(model, initial_states, future_states, max_horizon: int = 50, periodic_reencode_periods=(5, 10, 25)):
    """Multi-step prediction error of a model on finance test data (train.py:221-300): same arguments, same keys in the
    returned dictionary.  initial_states [batch, obs], future_states [horizon, batch, obs]; curves are per horizon step
    (MSE over batch and features, L2 norm over features averaged over the batch), returned on the host like the
    reference returns them."""
    import torch
    model.eval()
    dev = model.device
    horizon = min(int(max_horizon), int(future_states.shape[0]))
    initial_states = torch.as_tensor(initial_states).to(dev, dtype=torch.float32)
    true = torch.as_tensor(future_states)[:horizon].to(dev, dtype=torch.float32)
    predictions, mse_curves, l2_curves = {}, {}, {}

    def score(name, pred):
        predictions[name] = pred
        mse_curves[name] = ((pred - true) ** 2).mean(dim=(1, 2))
        l2_curves[name] = torch.norm(pred - true, dim=-1).mean(dim=1)

    score("every_step", rollout_every_step_reencode(model, initial_states, horizon))
    score("no_reencode", rollout_no_reencode(model, initial_states, horizon))
    for period in periodic_reencode_periods:
        score(f"periodic_{period}", rollout_periodic_reencode(model, initial_states, horizon, period=period))
    mean_mses = {mode: curve.mean().item() for mode, curve in mse_curves.items()}
    best_mode = min(mean_mses, key=mean_mses.get)
    return {
        "mse_reencode": mse_curves["every_step"].cpu(),
        "mse_no_reencode": mse_curves["no_reencode"].cpu(),
        "l2_reencode": l2_curves["every_step"].cpu(),
        "l2_no_reencode": l2_curves["no_reencode"].cpu(),
        "mean_mse_reencode": mean_mses["every_step"],
        "mean_mse_no_reencode": mean_mses["no_reencode"],
        "final_mse_reencode": mse_curves["every_step"][-1].item(),
        "final_mse_no_reencode": mse_curves["no_reencode"][-1].item(),
        "pred_reencode": predictions["every_step"].cpu(),
        "pred_no_reencode": predictions["no_reencode"].cpu(),
        "true": true.cpu(),
        "mse_curves": {k: v.cpu() for k, v in mse_curves.items()},
        "l2_curves": {k: v.cpu() for k, v in l2_curves.items()},
        "mean_mses": mean_mses,
        "predictions": {k: v.cpu() for k, v in predictions.items()},
        "best_mode": best_mode,
        "best_mse": mean_mses[best_mode],
    }
