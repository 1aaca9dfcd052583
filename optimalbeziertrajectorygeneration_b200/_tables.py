"""Host-side constant tables, built with the same scipy.special.binom
expressions as the reference so that the device tables hold
bit-identical values:
  elevMatrix (bezier.py:1127-1147), prodMatrix / bezProductCoefficients
  (bezier.py:1151-1208).
"""
import functools

import numpy as np
from scipy.special import binom


@functools.lru_cache(maxsize=None)
def elev_matrix(N, R):
    """T[j, i] = C(N,j) C(R,i-j) / C(N+R,i), shape (N+1, N+R+1)."""
    j = np.arange(N + 1)[:, None]
    i = np.arange(N + R + 1)[None, :]
    T = binom(N, j) * binom(R, i - j) / binom(N + R, i)
    T.setflags(write=False)
    return T


@functools.lru_cache(maxsize=None)
def prod_weights(m, n=None):
    """W[i, j] = C(m,i) C(n,j) / C(m+n,i+j), shape (m+1, n+1)."""
    if n is None:
        n = m
    i = np.arange(m + 1)[:, None]
    j = np.arange(n + 1)[None, :]
    W = binom(m, i) * binom(n, j) / binom(m + n, i + j)
    W.setflags(write=False)
    return W
