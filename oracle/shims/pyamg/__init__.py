"""TEST INFRASTRUCTURE ONLY (oracle shim).

Stand-in for the un-vendored third-party dependency `pyamg` (version unpinned by
the reference: no requirements file).  The reference imports it at
multigrid.py:4,6 and only ever calls `pyamg.krylov.fgmres` (multigrid.py:362).
pyamg is not installable in this image (no network), so `krylov.fgmres` is a
restatement of pyamg's published algorithm -> this part of the oracle is
"parity unpinned" (see oracle/README.md).
"""
