"""Observable-evaluation surface of the reference's plugins, on device statevectors.

The reference's plugins talk to a QPU through ``self.execute(job).value`` with
``job = circuit.to_job(observable=O)`` (/root/reference/qat/plugins/adapt_vqe.py:182,
/root/reference/qat/fermion/vqe.py:53-59, /root/reference/qat/plugins/sequential_optimization.py:96-100).
State preparation from circuits belongs to the QPU and is not part of this path: here the state
arrives as a device buffer (a torch CUDA tensor) or as a callable ``params -> state``, and the
observable evaluations are done by libqfe.

* ``StatevectorQPU.submit(observable, psi) -> Result``      <->  ``qpu.submit(job) -> Result`` for OBS jobs
* ``AdaptVQEGradients``                                      <->  AdaptVQEPlugin._compute_commutators + the gradient loop
                                                                 (adapt_vqe.py:50-62, 181-195)
* ``SeqOptim.optimize``                                      <->  SeqOptim.optimize (sequential_optimization.py:104-146)
* ``VQE``, ``vqe_energy``                                    <->  VQE and its cost function (vqe.py:13-62)
* ``AdaptVQE.run``                                           <->  AdaptVQEPlugin.run (adapt_vqe.py:143-240), state preparation on the device
* ``qse_matrices``, ``apply_quantum_subspace_expansion``,
  ``build_linear_pauli_expansion``                           <->  chemistry/qse.py:17-241
* ``GradientDescentOptimizer``                               <->  GradientDescentOptimizer.optimize (naturalgradient/qfim_plugin.py:96-251)
"""

from __future__ import annotations

import math
import warnings
from typing import Callable, List, Optional, Sequence

import numpy as np

from .hamiltonians import SpinHamiltonian


class Result:
    """``qat.core.Result`` look-alike: ``value``, ``meta_data`` (dict of strings), ``parameter_map``."""

    def __init__(self, value=None, meta_data=None, parameter_map=None):
        self.value = value
        self.meta_data = meta_data if meta_data is not None else {}
        self.parameter_map = parameter_map


class StatevectorQPU:
    """Evaluates OBS jobs on a given device statevector: ``submit(observable, psi).value = <psi|O|psi>``."""

    def __init__(self, engine=None):
        from .engine import get_engine

        self.engine = engine or get_engine()

    def _terms(self, observable):
        if isinstance(observable, SpinHamiltonian):
            return observable.packed(self.engine)
        return observable  # PackedTerms / Plan

    def submit(self, observable, psi) -> Result:
        val = self.engine.expectation(self._terms(observable), psi)
        return Result(value=val.real if abs(val.imag) <= 1e-12 * max(1.0, abs(val)) else val)

    def submit_batch(self, observables: Sequence, psi) -> List[Result]:
        """Many observables on ONE state (what the ADAPT / natural-gradient loops do job by job): one batched sweep."""
        batch = self.engine.concat([self._terms(o) for o in observables])
        vals = np.atleast_1d(self.engine.expectation(batch, psi))
        return [Result(value=complex(v)) for v in vals]


class AdaptVQEGradients:
    """Gradient sweep of ADAPT-VQE: ``g_k = -i <psi|[H, tau_k]|psi>`` for every pool operator.

    ``mode="matrix-free"`` (default) evaluates 2 Im <H psi|tau_k|psi> with one H|psi> and one batched
    pass over the pool; ``mode="commutators"`` follows the reference literally: the observables
    ``H | tau_k`` are expanded (``_compute_commutators``) and each is measured on the state."""

    def __init__(self, operator_pool: Sequence[SpinHamiltonian], tol_vanishing_grad: float = 1e-3, commutators=None, engine=None, mode: str = "matrix-free"):
        from .engine import get_engine

        if mode not in ("matrix-free", "commutators"):
            raise ValueError(f"unknown mode {mode!r}")
        self.engine = engine or get_engine()
        self.pool = list(operator_pool)
        self.tol_vanishing_grad = tol_vanishing_grad
        self.commutators = commutators
        self.mode = mode
        self._pool_terms = None
        self._commutator_batch = None  # (observable the commutators were expanded for, concatenated term set)

    @staticmethod
    def _compute_commutators(observable: SpinHamiltonian, pool: Sequence[SpinHamiltonian]):
        return [observable | op for op in pool]

    def gradients(self, observable: SpinHamiltonian, psi) -> np.ndarray:
        eng = self.engine
        if self.mode == "commutators":
            # commutators handed in by the caller are used as they are (adapt_vqe.py:40,169-170); the ones expanded here belong to one
            # observable (held by reference, so its identity cannot be recycled) and are expanded again for another
            if self.commutators is None or (self._commutator_batch is not None and self._commutator_batch[0] is not observable):
                self.commutators = self._compute_commutators(observable, self.pool)
                self._commutator_batch = None
            if self._commutator_batch is None:
                self._commutator_batch = (observable, eng.concat([c.packed(eng) for c in self.commutators]))
            vals = np.atleast_1d(eng.expectation(self._commutator_batch[1], psi))
            return np.real(-1j * vals)
        if self._pool_terms is None:
            self._pool_terms = eng.concat([op.packed(eng) for op in self.pool])
        # observable.packed() is cached on the object and rebuilt when its constant changed: ask every time
        return eng.commutator_gradients(observable.packed(eng), self._pool_terms, psi)

    def select(self, energy_gradients: np.ndarray):
        """(operator index or None, gradient norm) as adapt_vqe.py:186-195."""
        norm = float(np.linalg.norm(energy_gradients))
        if norm < self.tol_vanishing_grad:
            warnings.warn("Norm of the energy gradient is below the set threshold. Ending calculation.", stacklevel=2)
            return None, norm
        return int(np.argmax(np.abs(energy_gradients))), norm


class SeqOptimResult:
    def __init__(self, x, fun):
        self.x = x
        self.fun = fun


class SeqOptim:
    """Rotosolve (sequential_optimization.py:53-146): three energy evaluations per angle and cycle."""

    def __init__(self, ncycles: Optional[int] = 10, coeff: Optional[float] = 1, x0: Optional[np.ndarray] = None, verbose: Optional[bool] = False):
        self.ncycles_roto = ncycles
        self.coeff = coeff
        self.x0 = x0
        self.verbose = verbose
        self.n_evaluations = 0

    def optimize(self, var_names: Sequence[str], evaluate: Callable[[dict], float]):
        """``evaluate(params: dict) -> energy`` plays the role of ``Optimizer.evaluate``.

        One coordinate update of rotosolve: E(theta_d) = a sin(theta_d + b) + c along every angle, so the three values
        E(theta), E(theta + pi/2), E(theta - pi/2) fix the minimiser theta - pi/2 - atan2(2E - E+ - E-, E+ - E-)
        (sequential_optimization.py:125-135); the energy at the new angle becomes the E of the next coordinate."""
        names = [str(v) for v in var_names]
        theta = np.array(self.x0 if self.x0 is not None else np.random.random(len(names)), dtype=object)

        def energy_at(angles) -> float:
            self.n_evaluations += 1
            try:
                scaled = {name: angle * self.coeff for name, angle in zip(names, angles)}
            except Exception as exc:
                raise ValueError("Parameters must be real-valued!") from exc
            return np.real(evaluate(scaled))

        def shifted(d: int, delta: float):
            out = theta.copy()
            out[d] = out[d] + delta
            return out

        energy = evaluate({name: angle * self.coeff for name, angle in zip(names, theta)})
        self.n_evaluations += 1
        for cycle in range(self.ncycles_roto):
            if self.verbose:
                print("-------Optimization cycle #%i-----------" % (cycle + 1))
            for d in range(len(names)):
                e_plus, e_minus = energy_at(shifted(d, math.pi / 2)), energy_at(shifted(d, -math.pi / 2))
                theta = shifted(d, -math.pi / 2 - math.atan2(2 * energy - e_plus - e_minus, e_plus - e_minus))
                energy = energy_at(theta)
            if self.verbose:
                print("current cost:", energy)
        result = SeqOptimResult(list(theta), energy)
        return result.fun, list(result.x), result


def vqe_energy(hamiltonian: SpinHamiltonian, state_routine: Callable, engine=None) -> Callable:
    """``VQE.fun`` (vqe.py:53-59): theta -> <psi(theta)|H|psi(theta)> with ``state_routine(theta)`` a device state."""
    from .engine import get_engine

    eng = engine or get_engine()
    plan = eng.plan(hamiltonian.packed(eng))

    def fun(theta):
        return eng.energy(plan, state_routine(theta))

    return fun


def VQE(hamiltonian: SpinHamiltonian, optimizer: Callable, state_routine: Callable, theta0, engine=None, n_shots=None):
    """``VQE`` (/root/reference/qat/fermion/vqe.py:13-62): ``optimizer(fun, theta0) -> (theta, energy, nfev, trace)`` minimises
    theta -> <psi(theta)|H|psi(theta)>, with ``state_routine(theta)`` returning the device state the reference would prepare from
    ``ansatz_routine(theta)``.  Returns (energy, theta, None, None) like the reference; ``n_shots`` is accepted and ignored (exact
    expectation values, the reference's nbshots = 0)."""
    theta, energy, _, _ = optimizer(vqe_energy(hamiltonian, state_routine, engine=engine), theta0)
    return energy, theta, None, None


def qse_matrices(hamiltonian: SpinHamiltonian, psi, expansion_operators: Sequence[SpinHamiltonian], engine=None, threshold: float = 1e-15):
    """Quantum-subspace-expansion matrices H_ij = <psi|O_i^dag H O_j|psi>, S_ij = <psi|O_i^dag O_j|psi>
    (/root/reference/qat/fermion/chemistry/qse.py:191-214: the reference expands O_i^dag H O_j as an observable and submits dim^2 jobs
    per matrix).  Here phi_j = O_j|psi> and H|phi_j> are formed on the device once per j and the matrices are overlaps.  Returns real
    matrices with entries below ``threshold`` in modulus set to zero, like the reference."""
    from .engine import get_engine

    eng = engine or get_engine()
    h_plan = eng.plan(hamiltonian.packed(eng))
    phis = [eng.apply(eng.plan(op.packed(eng)), psi) for op in expansion_operators]
    chis = [eng.apply(h_plan, phi) for phi in phis]
    # two batched overlap passes instead of 2 dim^2 dot products: S = <phi_i|phi_j>, H = <phi_i|H phi_j>
    s_mat = eng.gram(phis)
    h_mat = eng.gram(phis, chis)
    matrix_h = np.where(np.abs(h_mat) > threshold, h_mat.real, 0.0)
    matrix_s = np.where(np.abs(s_mat) > threshold, s_mat.real, 0.0)
    return matrix_h, matrix_s


class GradientDescentOptimizer:
    """``GradientDescentOptimizer.optimize`` (/root/reference/qat/fermion/naturalgradient/qfim_plugin.py:96-251) with the ancilla
    circuits replaced by overlaps of device states: theta <- theta - lambda_step * g^+ grad E, g the quantum Fisher information
    metric (natural gradient) or the identity.  ``state_and_derivatives(theta) -> (psi, [d psi / d theta_k])`` supplies the state
    and its parameter derivatives (state preparation is outside this path; ``Engine.pauli_rotations`` can build them)."""

    def __init__(self, maxiter: int = 1000, natural_gradient: bool = True, lambda_step: float = 0.2, stop_crit: str = "grad_norm",
                 tol: float = 1e-10, x0=None, engine=None):
        if stop_crit not in ("grad_norm", "energy_dist", "none"):
            raise ValueError(f"unknown stopping criterion {stop_crit}")
        self.maxiter = maxiter
        self.natural_gradient = natural_gradient
        self.lambda_step = lambda_step
        self.stop_crit = stop_crit
        self.tolerance = tol
        self.initial_parameters = None if x0 is None else np.array(x0, dtype=float)
        self.engine = engine
        self.trace: List[float] = []
        self.iterations = 0
        self.check_crit_val = None

    def optimize(self, var_names: Sequence[str], hamiltonian: SpinHamiltonian, state_and_derivatives: Callable):
        from .engine import get_engine

        eng = self.engine or get_engine()
        plan = eng.plan(hamiltonian.packed(eng))
        n_par = len(var_names)
        if self.initial_parameters is None:
            self.initial_parameters = np.random.uniform(0, 2 * np.pi, n_par)
        if len(self.initial_parameters) != n_par:
            raise ValueError(f"Initial parameters ({len(self.initial_parameters)}) do not match the job parameter names {list(var_names)}")
        cur_params = np.array(self.initial_parameters, dtype=float)
        angles = [list(cur_params)]
        psi, dpsis = state_and_derivatives(cur_params)
        energy = eng.energy(plan, psi)
        self.trace = [energy]
        self.iterations = 0
        stopping = False
        while self.iterations < self.maxiter and not stopping:
            cur_grad = eng.energy_gradient(plan, dpsis, psi)
            cur_grad = np.real(np.where(np.abs(cur_grad) < 1e-10, 0.0, cur_grad))
            if self.natural_gradient:
                qfim = eng.qfim(dpsis, psi)
                qfim[np.abs(qfim) < 1e-10] = 0
                try:
                    cur_params = np.real(cur_params + np.linalg.solve(qfim, -self.lambda_step * cur_grad))
                except np.linalg.LinAlgError:
                    cur_params = np.real(cur_params - self.lambda_step * np.linalg.pinv(qfim).dot(cur_grad))
            else:
                cur_params = np.real(cur_params - self.lambda_step * cur_grad)
            angles.append(list(cur_params))
            psi, dpsis = state_and_derivatives(cur_params)
            energy = eng.energy(plan, psi)
            if self.stop_crit == "energy_dist" and self.iterations > 0:
                self.check_crit_val = abs(energy - self.trace[-1])
            self.trace.append(energy)
            if self.stop_crit == "grad_norm":
                self.check_crit_val = float(np.linalg.norm(cur_grad))
            if self.stop_crit != "none" and self.check_crit_val is not None and self.check_crit_val < self.tolerance:
                stopping = True
            self.iterations += 1
        angle_energy = [(angles[i], self.trace[i]) for i in range(len(self.trace))]
        return self.trace[-1], list(cur_params), {"iterations": self.iterations, "energies": angle_energy}


class AdaptVQE:
    """``AdaptVQEPlugin.run`` (/root/reference/qat/plugins/adapt_vqe.py:143-240) on a device state: per iteration the gradients
    -i<psi|[H, tau_k]|psi> of the whole pool, the operator with the largest one appended to the ansatz as exp(-i theta tau) term by
    term (``_grow_ansatz``, :64-141), all angles re-optimised by ``minimizer`` (the role of the optimizer plugin stacked under the
    reference's plugin; default scipy COBYLA, the reference's test setting), until the gradient norm falls below
    ``tol_vanishing_grad`` or ``n_iterations`` is reached.  ``Result.meta_data`` holds the reference's four string fields."""

    def __init__(self, operator_pool: Sequence[SpinHamiltonian], n_iterations: int = 300, tol_vanishing_grad: float = 1e-3, minimizer: Optional[Callable] = None,
                 engine=None):
        from .engine import get_engine

        self.engine = engine or get_engine()
        self.pool = list(operator_pool)
        self.n_iterations = n_iterations
        self.gradients = AdaptVQEGradients(self.pool, tol_vanishing_grad=tol_vanishing_grad, engine=self.engine)
        self.minimizer = minimizer or self._cobyla

    @staticmethod
    def _cobyla(fun, x0):
        import scipy.optimize

        trace = []

        def wrapped(x):
            v = fun(x)
            trace.append(v)
            return v

        res = scipy.optimize.minimize(wrapped, x0, method="COBYLA", tol=1e-6, options={"maxiter": 500})
        return float(res.fun), np.atleast_1d(res.x), trace

    def run(self, observable: SpinHamiltonian, initial_basis_state: int) -> Result:
        eng = self.engine
        n = observable.nbqbits
        plan = eng.plan(observable.packed(eng))
        operator_idx: List[int] = []
        thetas = np.zeros(0)
        energy_trace, optimization_trace, n_iters_optim = [], [], []

        def state(params):
            psi = eng.basis_state(n, initial_basis_state)
            eng.evolve_operators(psi, [self.pool[k] for k in operator_idx], params)
            return psi

        def energy(params):
            return eng.energy(plan, state(params))

        value = energy(thetas)
        for _ in range(self.n_iterations):
            grads = self.gradients.gradients(observable, state(thetas))
            op_ind, _ = self.gradients.select(grads)
            if op_ind is None:
                break
            operator_idx.append(op_ind)
            value, thetas, trace = self.minimizer(energy, np.append(thetas, 0.0))
            energy_trace.append(value)
            n_iters_optim.append(len(trace))
            optimization_trace += list(trace)
        res = Result(value=value, parameter_map={f"theta_{i}": float(t) for i, t in enumerate(thetas)})
        res.meta_data = {
            "operator_order": str(operator_idx), "energy_trace": str(energy_trace), "optimization_trace": str(optimization_trace),
            "n_iters_optim": str(n_iters_optim),
        }
        return res


def apply_quantum_subspace_expansion(hamiltonian: SpinHamiltonian, psi, expansion_operators: Sequence[SpinHamiltonian], engine=None, threshold: float = 1e-15,
                                     return_matrices: bool = False):
    """``apply_quantum_subspace_expansion`` (/root/reference/qat/fermion/chemistry/qse.py:17-135) for a device state: the lowest real
    part of the generalised eigenvalues of (H^(s), S) built by ``qse_matrices``."""
    import scipy.linalg

    matrix_h, matrix_s = qse_matrices(hamiltonian, psi, expansion_operators, engine=engine, threshold=threshold)
    e_qse = min(x.real for x in scipy.linalg.eig(matrix_h, matrix_s)[0])
    return (e_qse, matrix_h, matrix_s) if return_matrices else e_qse


def build_linear_pauli_expansion(pauli_gates: Sequence[str], nb_qubits: int) -> List[SpinHamiltonian]:
    """First-order expansion operators: every listed Pauli letter on every qubit, letters outermost (qse.py:216-241)."""
    from .hamiltonians import Term

    return [SpinHamiltonian(nb_qubits, [Term(1.0, pauli, [qbit])]) for pauli in pauli_gates for qbit in range(nb_qubits)]
