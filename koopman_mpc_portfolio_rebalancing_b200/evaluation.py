"""Batched rollout generators of the reference (/root/reference/evaluation.py:44-134) on the device kernels.

``rollout_no_reencode`` is the batched twin of the strategy's forecast loop (encode once, then H x step_latent +
decode) and runs as one fused chain (``kmpc_rollout``); the re-encoding variants compose ``encode`` /
``step_latent`` / ``decode`` (each one launch of the GEMM kernels over the whole batch).  Like the reference, a
non-finite prediction marks the remaining steps as NaN.  Plots and the ODE-system evaluation driver of
evaluation.py are outside the hot path.
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
