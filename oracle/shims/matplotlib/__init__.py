"""oracle shim: the reference imports matplotlib.pylab (multigrid.py:17,
stoch_trace.py:23, matrix.py:5) and never uses it; matplotlib is absent here."""
