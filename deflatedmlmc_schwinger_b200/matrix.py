"""Drop-in for the reference's matrix.py (loadMatrix, matrix.py:14-31)."""
import os
import warnings

import numpy as np
from scipy.sparse import identity

from . import lattice

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def loadMatrix(matrix_name, params):
    """A = S + m I as a CSC complex128 matrix, like the reference.

    Sources, in order: the MATLAB file `matrix_name` if it exists (same handling as
    matrix.py:21-29, including the gamma3 flip of schwinger16.mat); else the packaged link
    field data/<name>_links.npy (bit-identical matrix); `synthetic:L:seed[:sigma]` builds a
    random-U(1) lattice."""
    warnings.simplefilter("ignore")
    m = params['mass']
    if matrix_name.startswith("synthetic:"):
        parts = matrix_name.split(":")
        L, seed = int(parts[1]), int(parts[2])
        sigma = float(parts[3]) if len(parts) > 3 else 0.204
        return lattice.wilson_matrix(lattice.random_u1_links(L, seed, sigma), m)
    if os.path.isfile(matrix_name):
        import scipy.io as sio
        A = sio.loadmat(matrix_name)['S']
        if os.path.basename(matrix_name) == 'schwinger16.mat':
            mat_size = int(A.shape[0] / 2)
            A = A.tolil()
            A[mat_size:, :] = -A[mat_size:, :]
            A = A.tocsc()
        A = A + m * identity(A.shape[0], dtype=A.dtype)
        return A.tocsc()
    base = os.path.basename(matrix_name)
    base = base[:-4] if base.endswith(".mat") else base
    path = os.path.join(_DATA, base + "_links.npy")
    if not os.path.isfile(path):
        raise Exception("matrix file not found: " + matrix_name)
    return lattice.wilson_matrix(np.load(path), m)
