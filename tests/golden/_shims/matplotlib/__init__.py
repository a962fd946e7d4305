"""Empty stand-in: /root/reference/backtest.py:11 imports matplotlib.pyplot but never uses it, and
matplotlib is not installed in this image.  Used only by tests/golden/make_golden.py."""
