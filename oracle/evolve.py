"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  CPU restatement of the "next" rows of SURVEY.md 8(f): Trotter slices / ansatz
layers as ordered products of Pauli rotations, quantum-subspace-expansion matrices and the natural-gradient update.

* ``qat/fermion/trotterisation.py:85-150``  make_spin_hamiltonian_trotter_slice: per term a basis change (H on X, RX(pi/2) on Y),
  a CNOT ladder, RZ(2 coeff Re c_t) and the inverse.  RZ(2 t) = exp(-i t Z) conjugated by the basis change and the ladder is
  exp(-i t P_t): the circuit is restated gate by gate in ``trotter_slice_circuit`` and as rotations in ``trotter_slice``.
* ``qat/fermion/chemistry/qse.py:191-214``  H_ij = Re <psi|O_i^dag H O_j|psi>, S_ij = Re <psi|O_i^dag O_j|psi>, entries with
  modulus <= threshold set to 0.
* ``qat/fermion/naturalgradient/qfim_plugin.py:96-251``  theta <- theta + solve(g, -lambda grad E), g = 4 Re(<dk|dl> - <dk|psi><psi|dl>),
  entries of g and grad below 1e-10 truncated, pinv fallback, stopping criteria.

Parity pinning: the reference holds no golden vector for these rows beyond the optimisation end points of
``tests/integration/test_integration_natural_gradient_plugin.py`` (4 decimals, needs the closed-source simulator); the restatement
is checked against ``scipy.linalg.expm`` / dense matrices in tests/test_oracle_evolve.py.
"""

from __future__ import annotations

import numpy as np

from . import spin


def pauli_matrix(n: int, x: int, z: int) -> np.ndarray:
    """Dense i^{|x&z|} X^x Z^z (the packed-term operator with coefficient 1)."""
    return spin.csr_from_packed(n, [x], [z], [1.0], 0.0).toarray()


def rotate(n: int, x: int, z: int, theta: float, psi: np.ndarray) -> np.ndarray:
    """exp(-i theta P) psi = cos(theta) psi - i sin(theta) P psi."""
    return np.cos(theta) * psi - 1j * np.sin(theta) * spin.apply_packed(n, [x], [z], [1.0], 0.0, psi)


def rotations(n: int, xs, zs, thetas, psi: np.ndarray) -> np.ndarray:
    for x, z, t in zip(xs, zs, thetas):
        psi = rotate(n, int(x), int(z), float(t), psi)
    return psi


def trotter_slice(n: int, xs, zs, cs, coeff: float, psi: np.ndarray) -> np.ndarray:
    """prod_t exp(-i coeff Re(c_t) P_t) psi, first term applied first (trotterisation.py:139-150)."""
    return rotations(n, xs, zs, coeff * np.real(np.asarray(cs)), psi)


# ---- gate-level restatement of the reference's Trotter-slice circuit (independent check of ``rotate``) ----------------------
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)


def _rx(t):
    return np.array([[np.cos(t / 2), -1j * np.sin(t / 2)], [-1j * np.sin(t / 2), np.cos(t / 2)]], dtype=np.complex128)


def _rz(t):
    return np.array([[np.exp(-1j * t / 2), 0], [0, np.exp(1j * t / 2)]], dtype=np.complex128)


def _apply_1q(n, g, q, psi):
    t = psi.reshape([2] * n)
    t = np.moveaxis(np.tensordot(g, t, axes=([1], [q])), 0, q)  # qubit 0 = most significant = axis 0
    return t.reshape(-1)


def _apply_cnot(n, ctrl, tgt, psi):
    t = psi.reshape([2] * n).copy()
    idx = [slice(None)] * n
    idx[ctrl] = 1
    sub = t[tuple(idx)]
    axis = tgt if tgt < ctrl else tgt - 1
    t[tuple(idx)] = np.flip(sub, axis=axis)
    return t.reshape(-1)


def trotter_slice_circuit(n: int, terms, coeff: float, psi: np.ndarray) -> np.ndarray:
    """``terms``: [(coeff, op, qbits)].  Gate by gate as trotterisation.py:103-150: H on X, RX(pi/2) on Y, CNOT(previous, qb)
    walking the string backwards, RZ(2 coeff Re c) on the last visited qubit, then the inverse."""
    for c, op, qbits in terms:
        fwd = []
        for letter, q in zip(op, qbits):
            if letter == "X":
                fwd.append(("1q", _H, q))
            if letter == "Y":
                fwd.append(("1q", _rx(np.pi / 2), q))
        prev = len(qbits) - 1
        for k in range(len(qbits) - 2, -1, -1):
            if op[k] != "I":
                fwd.append(("cx", qbits[prev], qbits[k]))
                prev = k
        for g in fwd:
            psi = _apply_1q(n, g[1], g[2], psi) if g[0] == "1q" else _apply_cnot(n, g[1], g[2], psi)
        psi = _apply_1q(n, _rz(2 * coeff * np.real(c)), qbits[prev], psi)
        for g in reversed(fwd):
            psi = _apply_1q(n, g[1].conj().T, g[2], psi) if g[0] == "1q" else _apply_cnot(n, g[1], g[2], psi)
    return psi


def qse_matrices(n: int, h, ops, psi: np.ndarray, threshold: float = 1e-15):
    """h, ops[i]: (xs, zs, cs, constant) packed operators.  Dense restatement of qse.py:191-214."""
    hm = spin.csr_from_packed(n, *h).toarray()
    om = [spin.csr_from_packed(n, *o).toarray() for o in ops]
    dim = len(om)
    mh = np.zeros((dim, dim))
    ms = np.zeros((dim, dim))
    for i in range(dim):
        for j in range(dim):
            sh = np.vdot(psi, om[i].conj().T @ (hm @ (om[j] @ psi)))
            ss = np.vdot(psi, om[i].conj().T @ (om[j] @ psi))
            mh[i, j] = sh.real if abs(sh) > threshold else 0.0
            ms[i, j] = ss.real if abs(ss) > threshold else 0.0
    return mh, ms


def qfim(dpsis, psi) -> np.ndarray:
    d = np.array(dpsis)
    g = d.conj() @ d.T
    b = d.conj() @ psi
    return 4.0 * np.real(g - np.outer(b, b.conj()))


def gradient_descent(hmat, state_and_derivatives, x0, maxiter=1000, natural_gradient=True, lambda_step=0.2, stop_crit="grad_norm", tol=1e-10):
    """qfim_plugin.py:124-251 on dense states: returns (energy, params, trace)."""
    params = np.array(x0, dtype=float)
    psi, dps = state_and_derivatives(params)
    trace = [np.vdot(psi, hmat @ psi).real]
    it = 0
    stop = False
    crit = None
    while it < maxiter and not stop:
        phi = hmat @ psi
        grad = np.array([2.0 * np.vdot(d, phi).real for d in dps])
        grad = np.real(np.where(np.abs(grad) < 1e-10, 0.0, grad))
        if natural_gradient:
            g = qfim(dps, psi)
            g[np.abs(g) < 1e-10] = 0
            try:
                params = np.real(params + np.linalg.solve(g, -lambda_step * grad))
            except np.linalg.LinAlgError:
                params = np.real(params - lambda_step * np.linalg.pinv(g).dot(grad))
        else:
            params = np.real(params - lambda_step * grad)
        psi, dps = state_and_derivatives(params)
        e = np.vdot(psi, hmat @ psi).real
        if stop_crit == "energy_dist" and it > 0:
            crit = abs(e - trace[-1])
        trace.append(e)
        if stop_crit == "grad_norm":
            crit = np.linalg.norm(grad)
        if stop_crit != "none" and crit is not None and crit < tol:
            stop = True
        it += 1
    return trace[-1], params, trace
