"""oracle shim: empty stand-in for matplotlib.pylab (never used by the reference)."""
