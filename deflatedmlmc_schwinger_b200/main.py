"""Drop-in for the reference's main.py (main.py:12-22): `python -m deflatedmlmc_schwinger_b200.main`."""
import os

from .gateway import G101, G102, G201, G202  # noqa: F401

if __name__ == '__main__':
    # Schwinger 128^2, deflated MLMC -- the reference's shipped default (main.py:20-21)
    os.environ['OMP_NUM_THREADS'] = '1'
    G202()
