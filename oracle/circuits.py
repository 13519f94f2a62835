"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  State preparation used by two reference goldens.

Restates the 8-parameter, 4-qubit RY/CNOT circuit of
/root/reference/qat/fermion/circuits.py:207-232 (``make_shallow_circ``) as a dense statevector
simulation.  Qubit 0 is the most significant index bit, as everywhere in the reference.
"""

from __future__ import annotations

import numpy as np


def _apply_1q(psi: np.ndarray, n: int, q: int, u: np.ndarray) -> np.ndarray:
    t = psi.reshape((2**q, 2, 2 ** (n - q - 1)))
    return np.einsum("ab,xby->xay", u, t).reshape(-1)


def _apply_cnot(psi: np.ndarray, n: int, ctrl: int, tgt: int) -> np.ndarray:
    idx = np.arange(2**n)
    cb = 1 << (n - 1 - ctrl)
    tb = 1 << (n - 1 - tgt)
    src = np.where(idx & cb, idx ^ tb, idx)
    return psi[src]


def ry(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def shallow_circuit_state(theta) -> np.ndarray:
    """|psi(theta)> of make_shallow_circ on |0000>."""
    n = 4
    psi = np.zeros(2**n, dtype=complex)
    psi[0] = 1.0
    for i in range(4):
        psi = _apply_1q(psi, n, i, ry(theta[i]))
    psi = _apply_cnot(psi, n, 2, 3)
    psi = _apply_1q(psi, n, 2, ry(theta[4]))
    psi = _apply_1q(psi, n, 3, ry(theta[5]))
    psi = _apply_cnot(psi, n, 0, 2)
    psi = _apply_1q(psi, n, 0, ry(theta[6]))
    psi = _apply_1q(psi, n, 2, ry(theta[7]))
    psi = _apply_cnot(psi, n, 0, 1)
    return psi
