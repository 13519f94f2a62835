"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  Evaluation loops of the reference's plugins,
restated on dense matrices.

  * ADAPT-VQE gradient sweep, /root/reference/qat/plugins/adapt_vqe.py:50-62,181-195:
    commutators ``[H, tau_k]`` built as expanded observables, then
    ``g_k = -1j * <psi|[H, tau_k]|psi>``, ``||g||_2`` and ``argmax |g|``.
  * Rotosolve sequential optimiser, /root/reference/qat/plugins/sequential_optimization.py:104-146:
    three energy evaluations per parameter, closed-form angle update.
"""

from __future__ import annotations

import math

import numpy as np

from .spin import PauliSum, csr_from_packed


def commutator_gradients(h: PauliSum, pool: list[PauliSum], psi: np.ndarray) -> np.ndarray:
    """g_k = -i <psi|[H, tau_k]|psi>, each commutator expanded term by term like the reference does."""
    out = []
    for tau in pool:
        comm = h.commutator(tau)
        xs, zs, cs, const = comm.canonical(eps=0.0)
        mat = csr_from_packed(h.n, xs, zs, cs, const)
        out.append(-1j * np.vdot(psi, mat.dot(psi)))
    return np.array(out)


def select_operator(grads: np.ndarray, tol: float = 1e-3):
    """(index or None, ||g||) as adapt_vqe.py:186-195."""
    norm = float(np.linalg.norm(grads))
    if norm < tol:
        return None, norm
    return int(np.argmax(np.abs(grads))), norm


def rotosolve(energy, x0, ncycles: int = 10):
    """SeqOptim.optimize with coeff = 1: returns (final energy, parameters)."""
    params = list(x0)
    cf = np.real(energy(params))
    for _ in range(ncycles):
        for d in range(len(params)):
            plus = list(params)
            plus[d] = params[d] + math.pi / 2
            cf_plus = np.real(energy(plus))
            minus = list(params)
            minus[d] = params[d] - math.pi / 2
            cf_minus = np.real(energy(minus))
            new = list(params)
            new[d] = params[d] - math.pi / 2 - math.atan2(2 * cf - cf_plus - cf_minus, cf_plus - cf_minus)
            cf = np.real(energy(new))
            params = new
    return cf, params
