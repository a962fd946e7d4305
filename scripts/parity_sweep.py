"""Wide random parity sweep of the CUDA MPC solver against the fp64 oracle (test infrastructure, like tests/): many
shapes (N 2..64, H 1..5), lambda in {0, 1e-5..1e-1}, tau in {0, 0.01..1}, concentrated and diffuse current weights,
calm and wild forecasts.  Prints the worst objective / weight gaps per shape and every instance outside the parity bar
(objective 1e-6 relative, weights 1e-4).  Usage: python scripts/parity_sweep.py [instances_per_shape] [seed]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from koopman_mpc_portfolio_rebalancing_b200 import mpc
from oracle import mpc_oracle as mo

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
shapes = [(2, 1), (3, 5), (7, 2), (10, 5), (17, 3), (32, 5), (33, 4), (50, 5), (64, 5), (64, 1)]
bad = 0
t0 = time.time()
for (N, H) in shapes:
    w0 = np.stack([rng.dirichlet(np.ones(N) * rng.choice([0.05, 0.3, 1.0, 5.0])) for _ in range(P)])
    y = np.stack([(3e-4 + rng.standard_normal((H, N)) * rng.choice([0.001, 0.003, 0.01, 0.03, 0.1])) for _ in range(P)]).astype(np.float32)
    lam = rng.choice([0.0, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1], P)
    tau = rng.choice([0.0, 0.01, 0.05, 0.2, 0.5, 1.0], P)
    out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda(), lam=torch.from_numpy(lam).cuda(),
                              tau=torch.from_numpy(tau).cuda())
    W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy(); its = out["iterations"].cpu().numpy()
    wo = ww = 0.0; nbad = 0; ninacc = 0
    for p in range(P):
        ref = mo.solve_structured(w0[p], y[p], float(lam[p]), float(tau[p]))
        if ref.status != mo.STATUS_OPTIMAL:
            continue
        og = abs(val[p] - ref.value) / max(abs(ref.value), 1e-3) if st[p] <= 1 else np.inf
        wg = np.abs(W[p] - ref.w).max()
        ninacc += int(st[p] == 1)
        turn = np.abs(W[p][0] - w0[p]).sum()
        capbad = tau[p] > 0 and turn > tau[p] + 1e-9
        if st[p] > 1 or og > 1e-6 or capbad:
            nbad += 1
            print(f"  OUTSIDE N={N} H={H} p={p} lam={lam[p]:g} tau={tau[p]:g} status={st[p]} iters={its[p]} obj gap {og:.2e} w gap {wg:.2e} turn {turn:.6f}")
        if st[p] <= 1:
            wo = max(wo, og); ww = max(ww, wg)
    bad += nbad
    print(f"N={N:2d} H={H}: worst rel obj gap {wo:.2e}, worst |dw| {ww:.2e}, mean iterations {its.mean():.1f}, inaccurate {ninacc}, outside {nbad}")
print(f"{len(shapes) * P} instances, {bad} outside the bar, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
