"""The parity metric of north_star ("relative L1 / Linf error <= 1e-12 per conserved variable") made precise.

Used by the tests of `arithmetic=fast` and by bench.py's self-check; the strict build needs no tolerance (it is
compared bit for bit).

For rho and E:  L1 = sum|a - b| / sum|b|,  Linf = max|a - b| / max|b|  over the interior cells.
The momentum components are measured against the momentum scale of the state, sqrt(2 rho E) (>= |rho u|, |rho v| in
every cell, and ~ rho c in a gas at rest): a component that vanishes in exact arithmetic — by symmetry
(shocked_bubble's rho*v) or because the gas is at rest (the `discontinuity` deck: a contact at uniform pressure) —
holds nothing but round-off of size eps * rho * c, so its own norm is not a scale.
"""
from __future__ import annotations

import numpy as np

VAR_NAMES = ("rho", "E", "mx", "my")


def state_deviation(a: np.ndarray, b: np.ndarray):
    """a, b: [4, ny, nx] interior states (b = the reference).  Returns [(name, relL1, relLinf)] per conserved variable."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape and a.shape[0] == 4
    l1 = [float(np.sum(np.abs(b[v]))) for v in range(4)]
    li = [float(np.max(np.abs(b[v]))) for v in range(4)]
    mom = np.sqrt(2.0 * np.abs(b[0] * b[1]))
    l1[2] = l1[3] = float(np.sum(mom))
    li[2] = li[3] = float(np.max(mom))
    out = []
    for v in range(4):
        d = np.abs(a[v] - b[v])
        out.append((VAR_NAMES[v], float(np.sum(d)) / l1[v] if l1[v] > 0 else float(np.sum(d)),
                    float(np.max(d)) / li[v] if li[v] > 0 else float(np.max(d))))
    return out


def max_deviation(a: np.ndarray, b: np.ndarray) -> float:
    return max(max(l1, linf) for _, l1, linf in state_deviation(a, b))
