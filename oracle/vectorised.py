"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy-vectorised restatement of the fermion -> spin
transforms for electronic-structure Hamiltonians, used where the dictionary algebra of
oracle/transforms.py (the faithful restatement of /root/reference/qat/fermion/transforms.py:82-257)
is too slow (n >= 16).  Validated against oracle.transforms in tests/test_oracle_goldens.py.

Same mathematics: every ladder operator is 1/2 A_j -+ i/2 B_j, a term is the ordered product over
its operators, like terms are merged per (x, z) key; here the 2^k A/B choices of all terms are
expanded as uint64 array operations.
"""

from __future__ import annotations

import numpy as np

from .models import TOL
from .transforms import ladder_strings


def _popc(v):
    return np.bitwise_count(v).astype(np.int64)


def expand_electronic(hpq: np.ndarray, hpqrs: np.ndarray | None, constant, encoding: str, drop_eps: float = 1e-12):
    """(x, z, c_yform, constant) canonical arrays of ElectronicStructureHamiltonian(hpq, hpqrs, constant).to_spin(encoding)."""
    n = hpq.shape[0]
    lad = ladder_strings(n, encoding)
    ax = np.array([l[0][0] for l in lad], dtype=np.uint64)
    az = np.array([l[0][1] for l in lad], dtype=np.uint64)
    bx = np.array([l[1][0] for l in lad], dtype=np.uint64)
    bz = np.array([l[1][1] for l in lad], dtype=np.uint64)
    xs, zs, cs = [], [], []

    def expand(idx, w, dagger):
        k = idx.shape[1]
        for choice in range(1 << k):
            x = np.zeros(len(w), dtype=np.uint64)
            z = np.zeros(len(w), dtype=np.uint64)
            ph = np.zeros(len(w), dtype=np.int64)
            for l in range(k):
                q = idx[:, l]
                if (choice >> l) & 1:
                    sx, sz = bx[q], bz[q]
                    ph += 1 + (3 if dagger[l] else 1)  # B carries i (its Y); the ladder adds -i or +i
                else:
                    sx, sz = ax[q], az[q]
                ph += 2 * _popc(z & sx)
                x ^= sx
                z ^= sz
            xs.append(x)
            zs.append(z)
            cs.append(w * (0.5**k) * np.array([1, 1j, -1, -1j])[ph % 4])

    i2 = np.argwhere(np.abs(hpq) > TOL)
    if len(i2):
        expand(i2, hpq[tuple(i2.T)].astype(np.complex128), [True, False])
    if hpqrs is not None:
        i4 = np.argwhere(np.abs(hpqrs) > TOL)
        if len(i4):
            expand(i4, 0.5 * hpqrs[tuple(i4.T)].astype(np.complex128), [True, True, False, False])
    if not xs:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.complex128), complex(constant)
    x = np.concatenate(xs)
    z = np.concatenate(zs)
    c = np.concatenate(cs)
    order = np.lexsort((z, x))
    x, z, c = x[order], z[order], c[order]
    head = np.ones(len(x), dtype=bool)
    head[1:] = (x[1:] != x[:-1]) | (z[1:] != z[:-1])
    starts = np.flatnonzero(head)
    csum = np.add.reduceat(c, starts)
    ux, uz = x[starts], z[starts]
    const = complex(constant)
    ident = (ux == 0) & (uz == 0)
    if ident.any():
        const += complex(csum[ident][0])
    cy = csum * np.array([1, -1j, -1, 1j])[_popc(ux & uz) % 4]
    keep = (~ident) & (np.abs(cy) > drop_eps)
    return ux[keep], uz[keep], cy[keep], const
