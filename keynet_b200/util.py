"""Small host helpers used by key generation (reference: keynet/util.py:16-45)."""
import numpy as np
from numpy.lib.stride_tricks import as_strided


def find_closest_positive_divisor(a, b):
    """Divisor d > 1 of a that minimises |d - b| (ties resolved upward); a itself when a <= b.
    Used to snap tile / block sizes to the image size (keynet/util.py:16-28)."""
    assert a > 0 and b > 0
    if a <= b:
        return a
    for delta in range(0, a - b + 1):
        for cand in (b + delta, b - delta):
            if cand > 1 and a % cand == 0:
                return cand
    return a


def blockview(A, n):
    """View a (H,W) array as (H//n, W//n, n, n) blocks: blockview(A,n)[i,j] == A[i*n:(i+1)*n, j*n:(j+1)*n]."""
    assert A.ndim == 2
    (s0, s1) = A.strides
    return as_strided(A, shape=(A.shape[0] // n, A.shape[1] // n, n, n), strides=(n * s0, n * s1, s0, s1))
