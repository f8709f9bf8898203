"""oracle shim: empty package (multigrid.py:6 imports pyamg.aggregation.adaptive)."""
