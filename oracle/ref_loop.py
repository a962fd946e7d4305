"""Times the reference's OWN rebalancing loop on the host CPU (TEST / BENCH INFRASTRUCTURE, not product code).

BASELINE config 1 (finance_sparse SparseKM target_size = 128, 10 assets, d = 20, H = 5, 246 decisions of one backtest)
through the unmodified reference code — ``backtest.run_backtest`` + ``backtest.KoopmanMPCStrategy`` (backtest.py:67-219),
``model.make_model`` (torch CPU forecasts, batch 1 per decision as the reference runs them), ``data_finance`` splits —
imported from ``baseline/_ref/`` (a git-ignored copy of /root/reference made by ``__graft_entry__.build()``; it ships to
the GPU box with the snapshot).  The one substitution: the reference's ``mpc`` module needs cvxpy + SCS/ECOS, which
cannot be installed offline, so ``tests/golden/_shims/mpc.py`` (same names and signatures) routes the solve to the fp64
oracle interior-point solver.  Reported as kind "reference-loop+substitute-solver".

Run as a script by bench.py (a child process, so that the reference's module names — ``model``, ``backtest``,
``config`` — never enter the bench's own interpreter).  Prints one JSON line.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_shims"))      # mpc, matplotlib, cvxpy stand-ins
    import numpy as np
    import pandas as pd
    import torch
    import config as ref_config
    import data_finance as ref_data
    import model as ref_model
    import backtest as ref_backtest
    import mpc as shim_mpc
    from koopman_mpc_portfolio_rebalancing_b200 import synthetic

    assert os.path.dirname(os.path.abspath(ref_backtest.__file__)) == REF, ref_backtest.__file__
    N, d, H, T = 10, 20, 5, 1200
    lr = synthetic.gbm_log_returns(0, T, N)
    frame = pd.DataFrame(lr, index=pd.bdate_range("2012-01-02", periods=T), columns=[f"A{i}" for i in range(N)])
    val_end = str(frame.index[T - 253].date())
    train_end = str(frame.index[T - 253 - 200].date())
    stats = ref_data.compute_standardization_stats(frame, train_end)
    tr, trd, va, vad, te, ted = ref_data.create_finance_splits(frame, stats, train_end, val_end, d)
    env = ref_data.FinanceEnv(ref_data.FinanceDataset(tr, trd, 1), ref_data.FinanceDataset(va, vad, 1),
                              ref_data.FinanceDataset(te, ted, 1), stats, {"n_assets": N, "embedding_dim": d})
    cfg = ref_config.get_config("finance_sparse")
    cfg.MODEL.TARGET_SIZE = 128
    model = ref_model.make_model(cfg, N * d)
    sd = synthetic.generic_km_weights(0, N * d, [1024, 1024], 128)
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    mpc_cfg = shim_mpc.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2)
    bt_cfg = ref_backtest.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3)
    strat = ref_backtest.KoopmanMPCStrategy(model, mpc_cfg, device="cpu")
    # forecast-only share: the strategy's forecast loop with the solve stubbed out
    t0 = time.perf_counter()
    df = ref_backtest.run_backtest(strat, env, bt_cfg, verbose=False)
    dt = time.perf_counter() - t0
    n = len(df)
    metrics = ref_backtest.calculate_metrics(df)
    real_solve = shim_mpc.solve_mpc_log_utility
    ref_backtest.solve_mpc_log_utility = lambda w, y, c: (np.tile(w, (y.shape[0], 1)), {"status": "optimal", "value": 0.0})
    t0 = time.perf_counter()
    ref_backtest.run_backtest(strat, env, bt_cfg, verbose=False)
    dt_fc = time.perf_counter() - t0
    ref_backtest.solve_mpc_log_utility = real_solve
    print(json.dumps({
        "value": n / dt, "unit": "decisions/s", "cores": torch.get_num_threads(), "kind": "reference-loop+substitute-solver",
        "sample": f"config 1: one backtest x {n} decisions through the unmodified reference run_backtest + KoopmanMPCStrategy "
                  f"(baseline/_ref, torch CPU forecasts, batch 1) with the oracle solver in place of cvxpy/SCS: {dt:.2f} s; the "
                  f"same loop with the solve stubbed out (forecast + loop only): {dt_fc:.2f} s = {n / dt_fc:.0f} decisions/s",
        "seconds": dt, "forecast_loop_only_decisions_per_s": n / dt_fc, "final_value": float(metrics["Final Value"])}))


if __name__ == "__main__":
    main()
