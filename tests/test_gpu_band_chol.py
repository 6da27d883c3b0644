"""The single-CTA register-window band(+arrow) Cholesky of the exact-Schur mode
(csrc/ba_band_chol.cuh) against numpy on random SPD systems, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from particlesfm_b200 import _lib

pytestmark = pytest.mark.gpu


def _system(nb, bw, seed, arrow=True):
    rng = np.random.default_rng(seed)
    n = nb + 3
    B = rng.standard_normal((n, n))
    A = B @ B.T
    i, j = np.indices((n, n))
    A[(np.abs(i - j) > bw) & (i < nb) & (j < nb)] = 0.0
    if not arrow:                       # inactive intrinsics: identity rows, as k_band_assemble writes them
        A[nb:, :] = 0.0
        A[:, nb:] = 0.0
    A += np.eye(n) * (np.abs(A).sum(axis=1).max() + 1.0)
    b = rng.standard_normal(n)
    if not arrow:
        A[nb:, nb:] = np.eye(3)
        b[nb:] = 0.0
    return A, b


def _solve(A, b, nb, bw):
    x = np.zeros_like(b)
    rc = _lib.lib().psfm_ba_band_solve(_lib.dptr(np.ascontiguousarray(A)), _lib.dptr(b), nb, bw, _lib.dptr(x))
    return rc, x


# (nb, bw): tiny window, window wider than the matrix, nb not a multiple of 4, the bench shape
# (6 * 200 images, span 11/12), the 1024-thread instantiation, the widest supported window
CASES = [(30, 9), (18, 17), (18, 40), (90, 59), (1200, 71), (1200, 77), (600, 127), (3000, 71), (700, 150), (100, 31), (1203, 95),
         (64, 6), (41, 14), (500, 38)]


@pytest.mark.parametrize("nb,bw", CASES)
def test_band_solve_matches_numpy(gpu, nb, bw):
    A, b = _system(nb, bw, seed=nb + bw)
    rc, x = _solve(A, b, nb, bw)
    assert rc == 0, _lib.lib().psfm_last_error()
    ref = np.linalg.solve(A, b)
    assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max()


def test_band_solve_inactive_arrow_and_failure(gpu):
    A, b = _system(120, 35, seed=1, arrow=False)
    rc, x = _solve(A, b, 120, 35)
    assert rc == 0
    ref = np.linalg.solve(A, b)
    assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.array_equal(x[120:], np.zeros(3))
    A[57, 57] = -1.0                    # not positive definite -> reported, no garbage accepted
    rc, _ = _solve(A, b, 120, 35)
    assert rc == -1
    rc, _ = _solve(A, b, 120, 400)      # clamped to nb - 1 = 119 <= window limit: still solvable shape
    assert rc == -1
    A2, b2 = _system(400, 300, seed=2)                # window would be 304 > 152
    rc, _ = _solve(A2, b2, 400, 300)
    assert rc == -4                     # PSFM_ERR_UNSUPPORTED: wider than the register window
