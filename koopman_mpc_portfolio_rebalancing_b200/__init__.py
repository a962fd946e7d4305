"""koopman_mpc_portfolio_rebalancing_b200 — B200-native implementation of the Koopman-forecast + MPC rebalancing
hot path of yli421/koopman-mpc-portfolio-rebalancing.

Modules mirror the reference's for this path: ``mpc`` (MPCConfig, solve_mpc_log_utility, solve_mpc_mean_variance), ``backtest``
(BacktestConfig, Strategy, KoopmanMPCStrategy, run_backtest, calculate_metrics), ``model`` (GenericKM / SparseKM /
LISTAKM forward path), ``data_finance`` (embedding, splits, FinanceDataset, FinanceEnv), ``baselines`` (DMDStrategy, MarkowitzStrategy), ``evaluation`` (batched rollout generators), plus ``engine`` (the
batch-resident data-parallel form: scenario batches, sweep grids, bootstrap paths).  All compute goes through libkmpc.so (include/kmpc.h); nothing here imports
``oracle/`` and there is no CPU fallback.
"""
__all__ = ["mpc", "backtest", "baselines", "evaluation", "model", "data_finance", "engine", "synthetic"]
