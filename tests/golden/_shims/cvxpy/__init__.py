"""Empty stand-in so that /root/reference/baselines.py (`import cvxpy as cp`, line 16) can be imported by
tests/golden/make_golden.py: DMDStrategy never touches cvxpy itself (its MPC call goes through the `mpc` module, which
the sibling shim substitutes).  cvxpy cannot be installed offline."""
