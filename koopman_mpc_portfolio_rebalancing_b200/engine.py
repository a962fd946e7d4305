"""Batch-resident hot path: for many independent backtests at once
    standardise -> (virtual) delay embedding -> Koopman forecast of every step -> persistent MPC/portfolio loop -> metrics.

This is the data-parallel form of run_experiment.py:124-125 / README.md:45-68 (``KoopmanMPCStrategy`` +
``run_backtest`` + ``calculate_metrics``) for scenario paths, lambda/tau sweeps and Monte-Carlo stress tests.
Independent backtests shard across ranks with no communication; ``gather_metrics`` is the single collective.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _capi
from .backtest import BacktestConfig, run_backtest_batched
from .data_finance import pad4, standardize_device
from .mpc import MPCConfig


@dataclass
class PathBatch:
    """Inputs of a batch of scenario backtests: log-returns of the window [first lag day .. last test day]."""
    log_returns: "object"       # [B, T, N] float64, torch tensor (CUDA, or pinned host) or numpy
    mean: "object"              # [B, N] or [N] float64: training-split statistics (data_finance.py:211-240)
    std: "object"
    row0: int = 0               # embedded-row index of the first test row inside the window
    rows: int = 252             # rows of the test split


class BatchedBacktester:
    def __init__(self, model, n_assets: int, delay: int, mpc_config: Optional[MPCConfig] = None,
                 bt_config: Optional[BacktestConfig] = None, sequence_length: int = 1, device="cuda"):
        import torch
        self.model = model
        self.N, self.d = int(n_assets), int(delay)
        self.mpc = mpc_config or MPCConfig()
        self.bt = bt_config or BacktestConfig()
        self.seq_len = sequence_length
        self.device = torch.device(device)
        self._buf = {}

    def _tensor(self, name, shape, dtype):
        import torch
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    def n_steps(self, rows: int) -> int:
        return (rows - self.seq_len) - self.bt.horizon          # backtest.py:150, data_finance.py:389

    def run_device(self, log_returns, mean, std, row0: int, rows: int, lam=None, tau=None, want_history=False,
                   timings: Optional[dict] = None):
        """All inputs already on the device.  Returns dict(metrics [B,5], history, stats, yhat)."""
        import torch
        B, T, N = log_returns.shape
        H = self.mpc.horizon
        ns = self.n_steps(rows)
        ev = None
        if timings is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        z = self._tensor("z", (B, T, pad4(N)), torch.float32)
        h = _capi.Handle.get(self.device.index or 0)
        per_path = int(mean.dim() == 2)
        sp = _capi.stream_ptr(self.device.index or 0)
        _capi.check(_capi.lib().kmpc_standardize(h.ptr, _capi.ptr(log_returns), _capi.ptr(mean), _capi.ptr(std), per_path,
                                                B, T, N, _capi.ptr(z), z.shape[2], sp))
        realized = self._tensor("realized", (B, rows, N), torch.float32)
        _capi.check(_capi.lib().kmpc_current_returns(h.ptr, _capi.ptr(z), z.shape[2], _capi.ptr(mean), _capi.ptr(std),
                                                    per_path, B, T, N, self.d, row0, rows, _capi.ptr(realized), sp))
        if ev: ev[1].record()
        yhat = self._tensor("yhat", (B, ns, H, N), torch.float32)
        self.model.forecast_series(z, mean, std, N, self.d, row0, 0, ns, H, out=yhat)
        if ev: ev[2].record()
        metrics = self._tensor("metrics", (B, 5), torch.float64)
        out = run_backtest_batched(yhat, realized, n_steps=ns, horizon=H, lam=lam, tau=tau, lam0=self.mpc.cost_coeff,
                                   tau0=self.mpc.max_turnover, cost_coeff0=self.bt.cost_coeff,
                                   capital0=self.bt.initial_capital, rebalance_freq=self.bt.rebalance_freq,
                                   allow_short=self.mpc.allow_short, want_history=want_history, out_metrics=metrics)
        if ev:
            ev[3].record()
            timings["_events"] = ev
        out["yhat"] = yhat
        out["realized"] = realized
        return out

    def run(self, batch: PathBatch, lam=None, tau=None, want_history=False, copy_chunks: int = 8):
        """Host or device inputs -> metrics as a numpy array [B,5] (one H2D of the inputs, one D2H of the result).

        Host inputs are copied in ``copy_chunks`` slices of paths on a side stream while the main stream standardises
        and forecasts the slices that have already arrived (the forecast of a path needs nothing but that path), so
        that the host-to-device copy (447 MB per config-2 step) hides behind the forecast; the persistent MPC kernel
        then runs over all backtests at once (its work queue balances best with the whole batch)."""
        import torch
        lr = torch.as_tensor(batch.log_returns)
        mean = torch.as_tensor(batch.mean).to(self.device, dtype=torch.float64, non_blocking=True).contiguous()
        std = torch.as_tensor(batch.std).to(self.device, dtype=torch.float64, non_blocking=True).contiguous()
        B = lr.shape[0]
        if lr.is_cuda or lr.dtype != torch.float64 or copy_chunks <= 1 or B < 2 * copy_chunks or not lr.is_contiguous():
            lr_d = lr.to(self.device, dtype=torch.float64, non_blocking=True).contiguous()
            out = self.run_device(lr_d, mean, std, batch.row0, batch.rows, lam, tau, want_history)
        else:
            out = self._run_pipelined(lr, mean, std, batch.row0, batch.rows, lam, tau, want_history, copy_chunks)
        res = {"metrics": out["metrics"].cpu().numpy(), "stats": out["stats"].cpu().numpy() if out["stats"] is not None else None}
        if want_history:
            res["history"] = out["history"].cpu().numpy()
        return res

    def _run_pipelined(self, lr_host, mean, std, row0, rows, lam, tau, want_history, n_chunks):
        import torch
        B, T, N = lr_host.shape
        H = self.mpc.horizon
        ns = self.n_steps(rows)
        dev = self.device
        lr_d = self._tensor("lr_stage", (B, T, N), torch.float64)
        z = self._tensor("z", (B, T, pad4(N)), torch.float32)
        realized = self._tensor("realized", (B, rows, N), torch.float32)
        yhat = self._tensor("yhat", (B, ns, H, N), torch.float32)
        metrics = self._tensor("metrics", (B, 5), torch.float64)
        per_path = int(mean.dim() == 2)
        h = _capi.Handle.get(dev.index or 0)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        side = self._copy_stream
        side.wait_stream(main)                       # the staging buffer may still be read by the previous call
        per = (B + n_chunks - 1) // n_chunks
        events = []
        with torch.cuda.stream(side):
            for c0 in range(0, B, per):
                c1 = min(B, c0 + per)
                lr_d[c0:c1].copy_(lr_host[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
                events.append((c0, c1, ev))
        sp = _capi.stream_ptr(dev.index or 0)
        for (c0, c1, ev) in events:
            main.wait_event(ev)
            nb = c1 - c0
            mc = mean[c0:c1] if per_path else mean
            sc = std[c0:c1] if per_path else std
            _capi.check(_capi.lib().kmpc_standardize(h.ptr, _capi.ptr(lr_d[c0:c1]), _capi.ptr(mc), _capi.ptr(sc), per_path,
                                                    nb, T, N, _capi.ptr(z[c0:c1]), z.shape[2], sp))
            _capi.check(_capi.lib().kmpc_current_returns(h.ptr, _capi.ptr(z[c0:c1]), z.shape[2], _capi.ptr(mc), _capi.ptr(sc),
                                                        per_path, nb, T, N, self.d, row0, rows, _capi.ptr(realized[c0:c1]), sp))
            self.model.forecast_series(z[c0:c1], mc, sc, N, self.d, row0, 0, ns, H, out=yhat[c0:c1])
        out = run_backtest_batched(yhat, realized, n_steps=ns, horizon=H, lam=lam, tau=tau, lam0=self.mpc.cost_coeff,
                                   tau0=self.mpc.max_turnover, cost_coeff0=self.bt.cost_coeff,
                                   capital0=self.bt.initial_capital, rebalance_freq=self.bt.rebalance_freq,
                                   allow_short=self.mpc.allow_short, want_history=want_history, out_metrics=metrics)
        out["yhat"] = yhat
        out["realized"] = realized
        return out


def run_grid(models, n_assets: int, delay: int, log_returns, mean, std, lam_grid, tau_grid, *, row0: int = 0, rows: int = 252,
             horizon: int = 5, bt_config: Optional[BacktestConfig] = None, allow_short: bool = False, device="cuda",
             shard=None, want_history: bool = False):
    """Sweep grid (BASELINE config 4): every (model, lambda, tau) combination is one backtest on ONE price path.

    The forecast depends on (model, path, t) only, so it is computed once per model ([S, n_steps, H, N]) and the
    S * L * T backtests share it through ``yhat_index``; lambda and tau are per-backtest arrays of the persistent
    MPC kernel.  Backtest id = (s * L + l) * T + t.  ``shard = (rank, world)`` restricts the run to a contiguous
    shard of ids (every rank holds all S forecast sets; no data-path communication, see gather_metrics).
    Returns dict(metrics [n_local, 5] CUDA f64, ids (lo, hi), stats, history)."""
    import torch
    dev = torch.device(device)
    bt = bt_config or BacktestConfig(horizon=horizon)
    if bt.horizon != horizon:
        raise ValueError(f"run_grid: bt_config.horizon ({bt.horizon}) and horizon ({horizon}) must agree "
                         "(the number of steps and the forecasts are both derived from it)")
    lr = torch.as_tensor(log_returns).to(dev, dtype=torch.float64).reshape(1, -1, n_assets).contiguous()
    mean_d = torch.as_tensor(mean).to(dev, dtype=torch.float64).reshape(-1).contiguous()
    std_d = torch.as_tensor(std).to(dev, dtype=torch.float64).reshape(-1).contiguous()
    ns = (rows - 1) - bt.horizon
    S, L, T = len(models), len(lam_grid), len(tau_grid)
    n_total = S * L * T
    lo, hi = shard_range(n_total, *shard) if shard is not None else (0, n_total)
    yh, realized = [], None
    for m in models:
        eng = BatchedBacktester(m, n_assets, delay, MPCConfig(horizon=horizon), bt, device=dev)
        z = eng._tensor("z", (1, lr.shape[1], pad4(n_assets)), torch.float32)
        h = _capi.Handle.get(dev.index or 0)
        sp = _capi.stream_ptr(dev.index or 0)
        _capi.check(_capi.lib().kmpc_standardize(h.ptr, _capi.ptr(lr), _capi.ptr(mean_d), _capi.ptr(std_d), 0, 1, lr.shape[1],
                                                n_assets, _capi.ptr(z), z.shape[2], sp))
        if realized is None:
            realized = torch.empty((1, rows, n_assets), dtype=torch.float32, device=dev)
            _capi.check(_capi.lib().kmpc_current_returns(h.ptr, _capi.ptr(z), z.shape[2], _capi.ptr(mean_d), _capi.ptr(std_d), 0,
                                                        1, lr.shape[1], n_assets, delay, row0, rows, _capi.ptr(realized), sp))
        yh.append(m.forecast_series(z, mean_d, std_d, n_assets, delay, row0, 0, ns, horizon)[0].clone())
    yhat = torch.stack(yh, dim=0)
    ids = torch.arange(lo, hi, device=dev)
    lam_t = torch.as_tensor(np.asarray(lam_grid, dtype=np.float64), device=dev)[(ids // T) % L]
    tau_t = torch.as_tensor(np.asarray(tau_grid, dtype=np.float64), device=dev)[ids % T]
    out = run_backtest_batched(yhat, realized, n_steps=ns, horizon=horizon, lam=lam_t, tau=tau_t,
                               cost_coeff0=bt.cost_coeff, capital0=bt.initial_capital, rebalance_freq=bt.rebalance_freq,
                               allow_short=allow_short, yhat_index=(ids // (L * T)).to(torch.int32),
                               realized_index=torch.zeros_like(ids, dtype=torch.int32), B=hi - lo, want_history=want_history)
    out["ids"] = (lo, hi)
    out["yhat"] = yhat
    return out


def run_markowitz_batched(realized, risk_aversion: float = 1.0, cost_coeff: float = 1e-3, bt_config: Optional[BacktestConfig] = None,
                          allow_short: bool = False, lookback_window: int = 60, want_history: bool = False):
    """B Markowitz backtests advanced together: ``MarkowitzStrategy.rebalance`` (baselines.py:55-106) + the step loop of
    ``run_backtest`` (backtest.py:161-217) for every path at once.

    ``realized`` [B, rows, N] float32 CUDA: de-standardised log-returns of every test row (``kmpc_current_returns``).
    Per step t: rolling window ``realized[:, max(0, t+1-lookback):t+1]``, ``mu`` = its float32 mean, ``Sigma`` = its float64
    sample covariance + 1e-6 I (``np.cov`` semantics: float64 arithmetic on the float32 data, ddof = 1), one batched
    mean-variance solve (``kmpc_mpc_mean_variance`` with a covariance per problem, H = 1), then the portfolio step in
    float64.  The first four steps hold the weights (fewer than 5 observations, baselines.py:72-74).  The rolling
    statistics are plain batched reductions / one DGEMM per step (library calls); the solve is this library's kernel.
    Returns dict(metrics [B,5] f64 CUDA, history [B, n_hist, 4] or None, status_counts [3])."""
    import torch
    from .mpc import solve_mean_variance_batch
    bt = bt_config or BacktestConfig(horizon=1)
    assert realized.is_cuda and realized.dtype == torch.float32 and realized.dim() == 3
    B, rows, N = realized.shape
    dev = realized.device
    n_steps = (rows - 1) - bt.horizon                                    # len(test_dataset) - horizon, backtest.py:150
    steps = list(range(0, max(n_steps, 0), bt.rebalance_freq))
    w = torch.full((B, N), 1.0 / N, dtype=torch.float64, device=dev)     # backtest.py:161
    V = torch.full((B,), float(bt.initial_capital), dtype=torch.float64, device=dev)
    simple = (torch.exp(realized.double()).float() - 1.0)               # float32 simple returns, backtest.py:193
    hist = torch.empty((B, len(steps), 4), dtype=torch.float64, device=dev)
    eye = torch.eye(N, dtype=torch.float64, device=dev) * 1e-6
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    for k, t in enumerate(steps):
        target = w
        if t + 1 >= 5:
            win = realized[:, max(0, t + 1 - lookback_window):t + 1]     # [B, W, N] float32
            mu = win.double().mean(dim=1).float().double()              # the float32 mean, carried in float64
            X = win.double()
            Xc = X - X.mean(dim=1, keepdim=True)
            sigma = torch.bmm(Xc.transpose(1, 2), Xc) / float(win.shape[1] - 1) + eye
            out = solve_mean_variance_batch(w, mu.unsqueeze(1), sigma, risk_aversion, cost_coeff=cost_coeff, allow_short=allow_short)
            target = out["w"][:, 0]                                      # tile(w_cur) on a failed solve (mpc.py:179-180)
            st = out["status"]
            counts += torch.stack([(st == 0).sum(), (st == 1).sum(), (st >= 2).sum()])
        traded = (target - w).abs().sum(dim=1)
        fee = bt.cost_coeff * traded * V
        V = V - fee
        gain = torch.zeros_like(V)
        w = target
        if t + 1 < rows:
            r = simple[:, t + 1].double()
            gain = (target * r).sum(dim=1)
            V = V * (1.0 + gain)
            scale = 1.0 + gain
            scale = torch.where(scale.abs() < 1e-8, torch.full_like(scale, 1e-8), scale)
            w = target * (1.0 + simple[:, t + 1]).double() / scale.unsqueeze(1)
        hist[:, k, 0] = V; hist[:, k, 1] = gain; hist[:, k, 2] = traded; hist[:, k, 3] = fee
    ret = hist[:, :, 1]
    n = max(len(steps), 1)
    mean = ret.mean(dim=1)
    std = ((ret - mean.unsqueeze(1)) ** 2).mean(dim=1).sqrt()
    cum = torch.cumprod(1.0 + ret, dim=1)
    peak = torch.cummax(cum, dim=1).values
    metrics = torch.stack([np.sqrt(252.0) * mean / (std + 1e-8), ((cum - peak) / peak).min(dim=1).values,
                           hist[:, :, 2].mean(dim=1), hist[:, -1, 0], hist[:, -1, 0] / hist[:, 0, 0] - 1.0], dim=1)
    return {"metrics": metrics, "history": hist if want_history else None, "status_counts": counts}


def bootstrap_indices(n_paths: int, length: int, n_hist: int, seed: int, offset: int = 0, device="cpu"):
    """Row indices of the iid bootstrap: idx[p, t] = mix64(seed, offset + p, t) mod n_hist, a counter-based generator
    (the splitmix64 finaliser on a per-(path, day) counter) evaluated for all paths at once.  Path `offset + p` gets the
    same indices whatever the shard boundaries are, on any device, so a rank regenerates exactly its own shard."""
    import torch
    dev = torch.device(device)
    p = torch.arange(offset, offset + n_paths, dtype=torch.int64, device=dev).unsqueeze(1)
    t = torch.arange(length, dtype=torch.int64, device=dev).unsqueeze(0)
    # int64 arithmetic wraps modulo 2^64 (two's complement): exactly splitmix64 on unsigned counters
    x = (p * 0x100000001B3 + t) * (-7046029254386353131) + (int(seed) * 2 + 1) * 0x632BE59BD9B4E019 % (1 << 63)
    def shr(v, k):                                   # logical shift right of a signed int64
        return (v >> k) & ((1 << (64 - k)) - 1)
    x = (x ^ shr(x, 30)) * (-4658895280553007687)    # 0xBF58476D1CE4E5B9
    x = (x ^ shr(x, 27)) * (-7723592293110705685)    # 0x94D049BB133111EB
    x = x ^ shr(x, 31)
    return shr(x, 1) % int(n_hist)                   # 63 uniform bits; modulo bias < n_hist / 2^63


def bootstrap_paths(hist_log_returns, n_paths: int, length: int, seed: int, device="cuda", offset: int = 0):
    """Monte-Carlo stress paths (BASELINE config 5): iid row bootstrap of one historical block [T_hist, N].
    Indices come from the counter-based generator of ``bootstrap_indices`` (seed, path number, day), so that a rank
    regenerates exactly its own shard with ``offset`` = its first path (no scatter).  Returns (log_returns
    [n_paths, length, N] f64 on ``device``, idx int64)."""
    import torch
    dev = torch.device(device)
    hist = torch.as_tensor(hist_log_returns).to(dev, dtype=torch.float64)
    idx = bootstrap_indices(n_paths, length, hist.shape[0], seed, offset, dev)
    return hist[idx], idx


def metrics_frame(metrics, index=None):
    """[B,5] metric rows (device or host) -> DataFrame with the reference's column names (calculate_metrics,
    backtest.py:243-249; the layout of full_comparison_metrics.csv, run_experiment.py:133-137, one row per backtest)."""
    import pandas as pd
    from .backtest import METRIC_KEYS
    m = metrics.detach().cpu().numpy() if hasattr(metrics, "detach") else np.asarray(metrics)
    return pd.DataFrame(m, columns=list(METRIC_KEYS), index=index)


def history_frames(history, dates=None, rebalance_freq: int = 1):
    """[B, n_hist, 4] histories -> one DataFrame per backtest with the columns of run_backtest (backtest.py:211-219):
    date, portfolio_value, return, turnover, cost."""
    import pandas as pd
    from .backtest import HISTORY_COLS
    h = history.detach().cpu().numpy() if hasattr(history, "detach") else np.asarray(history)
    out = []
    for b in range(h.shape[0]):
        df = pd.DataFrame(h[b], columns=list(HISTORY_COLS))
        if dates is not None:
            df.insert(0, "date", [dates[t] for t in range(0, h.shape[1] * rebalance_freq, rebalance_freq)])
        out.append(df)
    return out


def write_results(path: str, metrics, history=None, ids=None):
    """Columnar dump of a batch for 10^4-10^6 backtests: metrics (and, optionally, the per-step history in long form)
    as parquet through pyarrow."""
    import pandas as pd
    mf = metrics_frame(metrics)
    mf.insert(0, "backtest", np.arange(len(mf)) if ids is None else np.asarray(ids))
    mf.to_parquet(path, index=False)
    if history is not None:
        from .backtest import HISTORY_COLS
        h = history.detach().cpu().numpy() if hasattr(history, "detach") else np.asarray(history)
        B, n, _ = h.shape
        long = pd.DataFrame(h.reshape(B * n, 4), columns=list(HISTORY_COLS))
        long.insert(0, "step", np.tile(np.arange(n), B))
        long.insert(0, "backtest", np.repeat(mf["backtest"].values, n))
        long.to_parquet(path.replace(".parquet", "") + ".history.parquet", index=False)
    return path


def compare_strategies(model, env, bt_config=None, mpc_config=None, risk_aversion: float = 1.0, out_dir=None,
                       strategies=("Buy & Hold", "Markowitz", "DMD-MPC", "Koopman-MPC")):
    """The four-strategy comparison of run_experiment.py:85-137 on one environment: Buy & Hold, Markowitz
    (risk_aversion, cost 1e-3), DMD fitted on the training split, and the Koopman model, all through
    ``run_backtest`` / ``calculate_metrics``.  Defaults are the reference's (capital 1e4, H = 5, cost 1e-3,
    max_turnover 0.5).  Returns ``(results, metrics_df)``: per-strategy history frames and the table the reference
    saves as ``full_comparison_metrics.csv`` (written to ``out_dir`` when given; no plotting)."""
    import os
    import pandas as pd
    from .backtest import BacktestConfig, BuyAndHoldStrategy, KoopmanMPCStrategy, run_backtest, calculate_metrics
    from .baselines import DMDStrategy, MarkowitzStrategy
    from .mpc import MPCConfig
    bt_config = bt_config or BacktestConfig(initial_capital=10000.0, horizon=5, rebalance_freq=1, cost_coeff=0.001)
    mpc_config = mpc_config or MPCConfig(horizon=5, gamma=0.0, cost_coeff=0.001, max_turnover=0.5)
    make = {
        "Buy & Hold": lambda: BuyAndHoldStrategy(),
        "Markowitz": lambda: MarkowitzStrategy(risk_aversion=risk_aversion, cost_coeff=0.001),
        "DMD-MPC": lambda: DMDStrategy(env.train_dataset.data, mpc_config),
        "Koopman-MPC": lambda: KoopmanMPCStrategy(model, mpc_config),
    }
    results, metrics = {}, {}
    for name in strategies:
        if name not in make:
            raise ValueError(f"unknown strategy '{name}'; available: {list(make)}")
        results[name] = run_backtest(make[name](), env, bt_config, verbose=False)
        metrics[name] = calculate_metrics(results[name])
    metrics_df = pd.DataFrame(metrics).T
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        metrics_df.to_csv(os.path.join(out_dir, "full_comparison_metrics.csv"))
    return results, metrics_df


def shard_range(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n independent backtests for one rank."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_metrics(local_metrics, n_total: int, rank: int, world: int):
    """The only collective of the path: all ranks' [B_local,5] metric rows -> [n_total,5] on every rank
    (torch.distributed all_gather over NCCL on GPUs, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_metrics
    per = (n_total + world - 1) // world
    pad = torch.zeros((per, local_metrics.shape[1]), dtype=local_metrics.dtype, device=local_metrics.device)
    pad[:local_metrics.shape[0]] = local_metrics
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat(bufs, dim=0)[:n_total]
