"""Shared helpers for the parity tests."""
import numpy as np


def relerr(a, b):
    """max |a-b| / max(|a|,|b|) elementwise (0 where both are 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    m = np.maximum(np.abs(a), np.abs(b))
    m[m == 0] = 1.0
    return float((d / m).max()) if d.size else 0.0


def colerr(a, b, axis=None):
    """max |a-b| / max|b| : error relative to the largest entry (Jacobian columns)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / (s if s > 0 else 1.0))


def cpu(t):
    return t.detach().cpu().numpy()
