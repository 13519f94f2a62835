"""Device engine: the Python host side of libqfe (kernels 1-3) on ONE GPU.

Call shapes follow what the reference's plugins ask of a QPU:
``energy(psi, H)`` <-> ``qpu.submit(circ.to_job(observable=H)).value``
(/root/reference/qat/fermion/vqe.py:53-59), ``commutator_gradients`` <-> the gradient loop of
/root/reference/qat/plugins/adapt_vqe.py:181-185, ``ground_state`` <-> ``eigvalsh(H.get_matrix())``.
State vectors are torch CUDA tensors (complex128 / complex64); torch is buffer plumbing only.
"""

from __future__ import annotations

import ctypes as C
import weakref
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import C64, C128, ENC, as_ptr, check
from .packed import PackedTerms

_engines: dict = {}


def _torch():
    import torch

    return torch


def _dtype_code(t) -> int:
    torch = _torch()
    if t.dtype == torch.complex128:
        return C128
    if t.dtype == torch.complex64:
        return C64
    raise TypeError(f"state vectors must be complex128 or complex64, got {t.dtype}")


class Plan:
    """Owner of a ``qfe_plan`` handle (sweep schedule of a term set for one shard geometry)."""

    def __init__(self, engine: "Engine", terms: PackedTerms, n_local: int | None = None, shard: int = 0, tile_bits: int = 0, _owned_by_terms: bool = False):
        self.engine = engine
        # a plan cached on its term set must not keep that set alive (the set owns the plan); a free-standing plan does
        self._terms_ref = weakref.ref(terms) if _owned_by_terms else None
        self._terms_strong = None if _owned_by_terms else terms
        self.n_slots = terms.n_slots
        self.nqbits = terms.nqbits
        self.n_local = terms.nqbits if n_local is None else n_local
        self.shard = shard
        h = C.c_void_p()
        check(engine.lib.qfe_plan_create(engine.ctx, terms.handle, self.n_local, shard, tile_bits, C.byref(h)))
        self._h = h
        ns, ng, nt = C.c_int64(), C.c_int64(), C.c_int64()
        nc, tb = C.c_int32(), C.c_int32()
        check(engine.lib.qfe_plan_info(h, C.byref(ns), C.byref(ng), C.byref(nt), C.byref(nc), C.byref(tb)))
        self.n_sweeps, self.n_groups, self.n_terms, self.n_classes, self.tile_bits = ns.value, ng.value, nt.value, nc.value, tb.value
        cls = np.zeros(max(self.n_classes, 1), dtype=np.int32)
        check(engine.lib.qfe_plan_classes(h, as_ptr(cls, C.c_int32), len(cls)))
        self.classes = [int(v) for v in cls[: self.n_classes]]
        herm = C.c_int()
        check(engine.lib.qfe_plan_hermitian(h, C.byref(herm)))
        self.hermitian = bool(herm.value)  # all coefficients real: every x_hi class is a Hermitian operator
        w = [C.c_int64() for _ in range(4)]
        check(engine.lib.qfe_plan_work(h, *[C.byref(v) for v in w]))
        # groups of the tiled layout visited on half / all rows of a tile in <psi|H|psi>, and their staged terms
        self.half_groups, self.full_groups, self.half_terms, self.full_terms = (v.value for v in w)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self.engine.lib.qfe_plan_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    @property
    def terms(self) -> PackedTerms:
        t = self._terms_strong if self._terms_ref is None else self._terms_ref()
        if t is None:
            raise RuntimeError("the term set of this plan has been garbage-collected")
        return t


class Engine:
    """One per (process, device).  Owns the ``qfe_ctx``."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("myqlm_fermion_b200 needs a CUDA device: the engine has no CPU path")
        self.device = device
        torch.cuda.set_device(device)
        h = C.c_void_p()
        check(self.lib.qfe_ctx_create(C.byref(h), device))
        self.ctx = h

    def __del__(self):
        h, self.ctx = getattr(self, "ctx", None), None
        if h:
            try:
                self.lib.qfe_ctx_destroy(h)
            except Exception:
                pass

    # ---- bookkeeping -------------------------------------------------------------------------
    def sync(self) -> None:
        check(self.lib.qfe_ctx_sync(self.ctx))

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self.lib.qfe_ctx_launch_count(self.ctx, C.byref(v)))
        return v.value

    def use_torch_stream(self) -> None:
        """Launch every libqfe kernel on torch's current stream (so that ``torch.cuda.Event`` brackets
        them).  The legacy default stream cannot be handed over as a handle: a fresh torch stream is
        made current in that case."""
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == 0:
            cur = torch.cuda.Stream(self.device)
            torch.cuda.set_stream(cur)
        self._torch_stream = cur  # keep it alive
        check(self.lib.qfe_ctx_set_stream(self.ctx, C.c_void_p(cur.cuda_stream)))

    def _sync_torch(self) -> None:
        # libqfe runs on its own non-blocking stream: order it after whatever torch queued
        _torch().cuda.current_stream(self.device).synchronize()

    # ---- kernel 1 ----------------------------------------------------------------------------
    def terms_from_integrals(self, hpq, hpqrs=None, constant=0.0, encoding="jordan-wigner", tol=1e-12, drop_eps=1e-12) -> PackedTerms:
        """ElectronicStructureHamiltonian(hpq, hpqrs, constant).to_spin(encoding) on the GPU."""
        enc = ENC[encoding]  # KeyError like the reference's transform_map (hamiltonians.py:454)
        hpq = np.ascontiguousarray(hpq, dtype=np.complex128)
        n = hpq.shape[0]
        if hpq.shape != (n, n):
            raise ValueError("hpq must be a square 2D array")
        p4 = None
        if hpqrs is not None:
            hpqrs = np.ascontiguousarray(hpqrs, dtype=np.complex128)
            if hpqrs.shape != (n, n, n, n):
                raise ValueError("hpqrs must have shape (n, n, n, n)")
            p4 = as_ptr(hpqrs.view(np.float64), C.c_double)
        k = np.array([complex(constant)], dtype=np.complex128)
        h = C.c_void_p()
        check(
            self.lib.qfe_terms_from_integrals(
                self.ctx, n, as_ptr(hpq.view(np.float64), C.c_double), p4, as_ptr(k.view(np.float64), C.c_double), tol, drop_eps, enc, C.byref(h)
            )
        )
        return PackedTerms(h)

    def terms_from_fermion(self, nqbits: int, terms, constant=0.0, encoding="jordan-wigner", drop_eps=1e-12) -> PackedTerms:
        """transform_to_*_basis of a general fermionic term list [(coeff, "CCcc", [p,q,r,s]) | Term]."""
        enc = ENC[encoding]
        ops, qbits, offs, coeffs = bytearray(), [], [0], []
        for t in terms:
            coeff, op, qb = (t.coeff, t.op, t.qbits) if hasattr(t, "coeff") else t
            if len(op) != len(qb):
                raise ValueError("operator string and qubit list differ in length")
            ops.extend(op.encode())
            qbits.extend(int(q) for q in qb)
            offs.append(len(qbits))
            coeffs.append(complex(coeff))
        ops_a = np.frombuffer(bytes(ops) or b"\0", dtype=np.uint8).copy()
        qb_a = np.array(qbits or [0], dtype=np.int32)
        of_a = np.array(offs, dtype=np.int64)
        co_a = np.array(coeffs or [0j], dtype=np.complex128)
        k = np.array([complex(constant)], dtype=np.complex128)
        h = C.c_void_p()
        check(
            self.lib.qfe_terms_from_fermion(
                self.ctx, nqbits, len(coeffs), as_ptr(ops_a, C.c_uint8), as_ptr(qb_a, C.c_int32), as_ptr(of_a, C.c_int64),
                as_ptr(co_a.view(np.float64), C.c_double), as_ptr(k.view(np.float64), C.c_double), drop_eps, enc, C.byref(h),
            )
        )
        return PackedTerms(h)

    def terms_from_pauli(self, nqbits: int, x, z, c, constant=0.0, drop_eps=0.0) -> PackedTerms:
        x = np.ascontiguousarray(x, dtype=np.uint64)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.complex128)
        if not (len(x) == len(z) == len(c)):
            raise ValueError("x, z, c must have the same length")
        k = np.array([complex(constant)], dtype=np.complex128)
        h = C.c_void_p()
        T = len(x)
        if T == 0:
            x, z, c = np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.zeros(1, np.complex128)
        check(
            self.lib.qfe_terms_from_pauli(
                self.ctx, nqbits, T, as_ptr(x, C.c_uint64), as_ptr(z, C.c_uint64), as_ptr(c.view(np.float64), C.c_double),
                as_ptr(k.view(np.float64), C.c_double), drop_eps, C.byref(h),
            )
        )
        return PackedTerms(h)

    def concat(self, ops: Sequence[PackedTerms]) -> PackedTerms:
        """Batch K operators (an operator pool) into one term set with K output slots."""
        arr = (C.c_void_p * len(ops))(*[o.handle for o in ops])
        h = C.c_void_p()
        check(self.lib.qfe_terms_concat(self.ctx, arr, len(ops), C.byref(h)))
        out = PackedTerms(h)
        out._keepalive = list(ops)
        return out

    # ---- kernel 2 ----------------------------------------------------------------------------
    def plan(self, terms: PackedTerms, n_local: int | None = None, shard: int = 0, tile_bits: int = 0) -> Plan:
        """Sweep schedule of ``terms`` for a shard geometry.  Plans are cached ON the term set (one per engine and geometry), so a plan
        and its device tables die with the PackedTerms they were built for: callers that make a fresh term set per call (commutator
        batches, get_matrix, QSE) do not accumulate plans."""
        cache = terms.__dict__.setdefault("_plans", {})
        key = (id(self), n_local, shard, tile_bits)
        p = cache.get(key)
        if p is None:
            p = Plan(self, terms, n_local, shard, tile_bits, _owned_by_terms=True)
            cache[key] = p
        return p

    def _check_state(self, psi, n_amps: int):
        torch = _torch()
        if not (isinstance(psi, torch.Tensor) and psi.is_cuda and psi.is_contiguous()):
            raise TypeError("state vector must be a contiguous torch CUDA tensor")
        if psi.numel() != n_amps:
            raise ValueError(f"state vector has {psi.numel()} amplitudes, expected {n_amps}")
        return _dtype_code(psi)

    def expectation(self, terms_or_plan, psi, bra=None) -> complex:
        """<bra|H|psi> (bra = psi when omitted): ``Result.value`` of an OBS job on state psi."""
        plan = terms_or_plan if isinstance(terms_or_plan, Plan) else self.plan(terms_or_plan)
        if plan.n_local != plan.nqbits:
            raise ValueError("use ShardedEngine for a sharded plan")
        dt = self._check_state(psi, 1 << plan.nqbits)
        if bra is not None and self._check_state(bra, 1 << plan.nqbits) != dt:
            raise TypeError("bra and ket dtypes differ")
        self._sync_torch()
        out = np.zeros(2 * plan.n_slots, dtype=np.float64)
        b = psi if bra is None else bra
        check(self.lib.qfe_expectation_class(self.ctx, plan.handle, 0, C.c_void_p(b.data_ptr()), C.c_void_p(psi.data_ptr()), dt, as_ptr(out, C.c_double)))
        vals = out.view(np.complex128)
        return complex(vals[0]) if plan.n_slots == 1 else vals.copy()

    def energy(self, terms_or_plan, psi) -> float:
        return float(np.real(self.expectation(terms_or_plan, psi)))

    def expectation_host(self, terms_or_plan, psi_host: np.ndarray) -> complex:
        """Same through HOST memory: the upload happens inside the library call (end-to-end path)."""
        plan = terms_or_plan if isinstance(terms_or_plan, Plan) else self.plan(terms_or_plan)
        if psi_host.dtype == np.complex128:
            dt = C128
        elif psi_host.dtype == np.complex64:
            dt = C64
        else:
            raise TypeError("host state must be complex128 or complex64")
        if psi_host.size != 1 << plan.nqbits or not psi_host.flags.c_contiguous:
            raise ValueError("host state must be a contiguous vector of 2^n amplitudes")
        out = np.zeros(2, dtype=np.float64)
        check(self.lib.qfe_expectation_host(self.ctx, plan.handle, C.c_void_p(psi_host.ctypes.data), dt, as_ptr(out, C.c_double)))
        return complex(out[0], out[1])

    def apply(self, terms_or_plan, psi, out=None):
        """H|psi> as a new (or given) device vector."""
        torch = _torch()
        plan = terms_or_plan if isinstance(terms_or_plan, Plan) else self.plan(terms_or_plan)
        dt = self._check_state(psi, 1 << plan.nqbits)
        if out is None:
            out = torch.empty_like(psi)
        self._check_state(out, 1 << plan.nqbits)
        self._sync_torch()
        check(self.lib.qfe_apply(self.ctx, plan.handle, C.c_void_p(psi.data_ptr()), C.c_void_p(out.data_ptr()), dt))
        self.sync()
        return out

    # ---- shard-level entry points (one x_hi class at a time; ShardedEngine drives these) ------
    def expectation_class0_host(self, plan: Plan, shard_host: np.ndarray, shard_dev) -> complex:
        """Upload this process's shard from HOST memory into ``shard_dev`` and return the class-0 (local) share of <psi|H|psi>; the
        copy runs on a second stream and hides behind the first sweeps (qfe_expectation_class0_host)."""
        n_amps = 1 << plan.n_local
        dt = self._check_state(shard_dev, n_amps)
        if shard_host.dtype != (np.complex128 if dt == C128 else np.complex64) or shard_host.size != n_amps or not shard_host.flags.c_contiguous:
            raise ValueError("host shard must be a contiguous vector of 2^n_local amplitudes of the device buffer's dtype")
        self._sync_torch()
        out = np.zeros(2, dtype=np.float64)
        check(self.lib.qfe_expectation_class0_host(self.ctx, plan.handle, C.c_void_p(shard_host.ctypes.data), C.c_void_p(shard_dev.data_ptr()), dt,
                                                   as_ptr(out, C.c_double)))
        return complex(out[0], out[1])

    def expectation_class(self, plan: Plan, x_hi: int, bra_shard, ket_shard, part: int = 0, n_parts: int = 1):
        """Class x_hi's share of <bra|H|ket>: ``bra_shard`` is this rank's shard of the bra,
        ``ket_shard`` the shard (shard ^ x_hi) of the ket.  ``part`` / ``n_parts``: every n_parts-th sweep of the class only."""
        n_amps = 1 << plan.n_local
        dt = self._check_state(ket_shard, n_amps)
        if self._check_state(bra_shard, n_amps) != dt:
            raise TypeError("bra and ket dtypes differ")
        self._sync_torch()
        out = np.zeros(2 * plan.n_slots, dtype=np.float64)
        check(
            self.lib.qfe_expectation_class_part(
                self.ctx, plan.handle, x_hi, C.c_void_p(bra_shard.data_ptr()), C.c_void_p(ket_shard.data_ptr()), dt, part, n_parts, as_ptr(out, C.c_double)
            )
        )
        vals = out.view(np.complex128)
        return complex(vals[0]) if plan.n_slots == 1 else vals.copy()

    def apply_class(self, plan: Plan, x_hi: int, ket_shard, y_shard, accumulate: bool, sync: bool = True) -> None:
        """y_shard (+)= H_class(x_hi) ket_shard, where ket_shard is the shard (shard ^ x_hi) of the ket."""
        n_amps = 1 << plan.n_local
        dt = self._check_state(ket_shard, n_amps)
        if self._check_state(y_shard, n_amps) != dt:
            raise TypeError("ket and output dtypes differ")
        self._sync_torch()
        check(
            self.lib.qfe_apply_class(
                self.ctx, plan.handle, x_hi, C.c_void_p(ket_shard.data_ptr()), C.c_void_p(y_shard.data_ptr()), dt, 1 if accumulate else 0
            )
        )
        if sync:
            self.sync()

    def commutator_gradients(self, h_plan, pool_plan, psi, scratch=None) -> np.ndarray:
        """g_k = -i <psi|[H, tau_k]|psi> for every pool operator (adapt_vqe.py:181-185)."""
        torch = _torch()
        h_plan = h_plan if isinstance(h_plan, Plan) else self.plan(h_plan)
        pool_plan = pool_plan if isinstance(pool_plan, Plan) else self.plan(pool_plan)
        dt = self._check_state(psi, 1 << h_plan.nqbits)
        if scratch is None:
            scratch = torch.empty_like(psi)
        self._sync_torch()
        g = np.zeros(pool_plan.n_slots, dtype=np.float64)
        check(
            self.lib.qfe_commutator_gradients(
                self.ctx, h_plan.handle, pool_plan.handle, C.c_void_p(psi.data_ptr()), C.c_void_p(scratch.data_ptr()), dt, as_ptr(g, C.c_double)
            )
        )
        return g

    # ---- kernel 3 ----------------------------------------------------------------------------
    def ground_state(self, terms_or_plan, v0=None, max_iter: int = 300, tol: float = 1e-10, dtype=None, want_vector: bool = True, seed: int = 0):
        """Lowest eigenpair by Lanczos on the matrix-free apply.  Returns (E0, psi0, info)."""
        torch = _torch()
        plan = terms_or_plan if isinstance(terms_or_plan, Plan) else self.plan(terms_or_plan)
        N = 1 << plan.nqbits
        if v0 is None:
            v0 = self.random_state(plan.nqbits, dtype or torch.complex128, seed=seed)
        else:
            v0 = v0.clone()
        dt = self._check_state(v0, N)
        work = torch.empty(3 * N, dtype=v0.dtype, device=v0.device)
        self._sync_torch()
        ev, it, res = C.c_double(), C.c_int(), C.c_double()
        check(
            self.lib.qfe_lanczos(
                self.ctx, plan.handle, C.c_void_p(v0.data_ptr()), C.c_void_p(work.data_ptr()), max_iter, tol, dt, 1 if want_vector else 0,
                C.byref(ev), C.byref(it), C.byref(res),
            )
        )
        return ev.value, (v0 if want_vector else None), {"iterations": it.value, "residual": res.value}

    # ---- state helpers -----------------------------------------------------------------------
    def random_state(self, nqbits: int, dtype=None, seed: int = 0, normalize: bool = True, n_local: int | None = None, shard: int = 0):
        torch = _torch()
        dtype = dtype or torch.complex128
        n_local = nqbits if n_local is None else n_local
        N = 1 << n_local
        psi = torch.empty(N, dtype=dtype, device=f"cuda:{self.device}")
        self._sync_torch()
        check(self.lib.qfe_vec_fill_random(self.ctx, C.c_void_p(psi.data_ptr()), N, _dtype_code(psi), seed, shard * N))
        if normalize and n_local == nqbits:
            nrm = self.dot(psi, psi).real
            self.scale(psi, 1.0 / np.sqrt(nrm))
        self.sync()
        return psi

    def product_state(self, nqbits: int, alpha_beta: np.ndarray, dtype=None, n_local: int | None = None, shard: int = 0):
        """(x)_q (alpha_q|0> + beta_q|1>); alpha_beta has shape (n, 2) complex."""
        torch = _torch()
        dtype = dtype or torch.complex128
        n_local = nqbits if n_local is None else n_local
        ab = np.ascontiguousarray(alpha_beta, dtype=np.complex128).reshape(nqbits, 2)
        psi = torch.empty(1 << n_local, dtype=dtype, device=f"cuda:{self.device}")
        self._sync_torch()
        check(self.lib.qfe_vec_product_state(self.ctx, C.c_void_p(psi.data_ptr()), n_local, shard, nqbits, as_ptr(ab.view(np.float64), C.c_double), _dtype_code(psi)))
        self.sync()
        return psi

    def basis_state(self, nqbits: int, index: int, dtype=None, n_local: int | None = None, shard: int = 0):
        """|index> with qubit 0 = most significant bit (e.g. the Hartree-Fock determinant: the first n_e qubits set)."""
        torch = _torch()
        dtype = dtype or torch.complex128
        n_local = nqbits if n_local is None else n_local
        N = 1 << n_local
        local = index - shard * N
        psi = torch.empty(N, dtype=dtype, device=f"cuda:{self.device}")
        self._sync_torch()
        check(self.lib.qfe_vec_basis_state(self.ctx, C.c_void_p(psi.data_ptr()), N, _dtype_code(psi), local if 0 <= local < N else -1))
        self.sync()  # like random_state / product_state: the vector is complete when the call returns, whatever stream reads it next
        return psi

    # ---- state evolution ----------------------------------------------------
    def pauli_rotations(self, psi, x, z, theta) -> None:
        """In place psi <- exp(-i theta[K-1] P[K-1]) ... exp(-i theta[0] P[0]) psi with P[k] = i^{|x&z|} X^x Z^z (packed masks,
        bit n-1-q <-> qubit q).  Consecutive rotations on nearby qubits are fused into one pass over the state."""
        nqbits = int(psi.numel()).bit_length() - 1
        x = np.ascontiguousarray(x, dtype=np.uint64)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if not (len(x) == len(z) == len(theta)):
            raise ValueError("x, z and theta must have the same length")
        self._check_state(psi, 1 << nqbits)
        self._sync_torch()
        check(self.lib.qfe_pauli_rotations(self.ctx, nqbits, len(x), as_ptr(x, C.c_uint64), as_ptr(z, C.c_uint64), as_ptr(theta, C.c_double),
                                           C.c_void_p(psi.data_ptr()), _dtype_code(psi)))

    def trotter_slice(self, hamiltonian, psi, coeff: float = 1.0, n_steps: int = 1) -> None:
        """In place first-order Trotter slice(s) prod_t exp(-i coeff Re(c_t) P_t) in the term order of ``hamiltonian``, the first
        term applied first: what make_spin_hamiltonian_trotter_slice(hamiltonian, coeff) builds as a circuit
        (/root/reference/qat/fermion/trotterisation.py:85-150).  ``hamiltonian``: a SpinHamiltonian or PackedTerms."""
        from .packed import PackedTerms, pack_terms

        if isinstance(hamiltonian, PackedTerms):
            x, z, c, _, _ = hamiltonian.arrays()
        else:
            x, z, c = pack_terms(hamiltonian.nbqbits, hamiltonian.terms)[:3]
        theta = coeff * np.real(np.asarray(c))
        for _ in range(n_steps):
            self.pauli_rotations(psi, x, z, theta)

    def evolve_operators(self, psi, operators, thetas) -> None:
        """In place psi <- prod_k exp(-i theta_k O_k) psi, first operator first, each exp(-i theta O) factorised term by term as
        prod_t exp(-i theta Re(c_t) P_t) in the term order of O: the parameterised Trotter slices AdaptVQEPlugin._grow_ansatz appends
        per selected pool operator (/root/reference/qat/plugins/adapt_vqe.py:64-141; like the reference only the real part of a
        coefficient is used).  ``operators``: Hermitian SpinHamiltonians, e.g. the cluster operators of get_cluster_ops
        (/root/reference/qat/fermion/chemistry/ucc.py:801-852) mapped to spins."""
        from .packed import pack_terms

        xs, zs, th = [], [], []
        for op, theta in zip(operators, thetas):
            x, z, c = pack_terms(op.nbqbits, op.terms)[:3]
            xs.append(np.asarray(x, dtype=np.uint64))
            zs.append(np.asarray(z, dtype=np.uint64))
            th.append(float(theta) * np.real(np.asarray(c)))
        if xs:
            self.pauli_rotations(psi, np.concatenate(xs), np.concatenate(zs), np.concatenate(th))

    # ---- overlaps ------------------------------------------------------
    def gram(self, bras, kets=None) -> np.ndarray:
        """G[i, j] = <bras[i]|kets[j]> (kets = None: the Hermitian Gram matrix of ``bras``) in one pass per panel of 32 x 32 vectors
        (qfe_vec_gram): every vector is read once per panel pair instead of once per entry."""
        bras = list(bras)
        if not bras:
            return np.zeros((0, 0 if kets is None else len(kets)), dtype=np.complex128)
        n_amps = bras[0].numel()
        dt = self._check_state(bras[0], n_amps)
        kets_l = None if kets is None else list(kets)
        for v in bras + (kets_l or []):
            if self._check_state(v, n_amps) != dt:
                raise TypeError("all vectors of a Gram matrix must have the same dtype")
        if kets_l is not None and not kets_l:
            return np.zeros((len(bras), 0), dtype=np.complex128)
        if n_amps % 32:  # registers of fewer than 5 qubits: entry by entry
            kk = bras if kets_l is None else kets_l
            return np.array([[self.dot(b, k) for k in kk] for b in bras], dtype=np.complex128)
        pb = (C.c_void_p * len(bras))(*[v.data_ptr() for v in bras])
        nk = len(bras) if kets_l is None else len(kets_l)
        pk = None if kets_l is None else (C.c_void_p * nk)(*[v.data_ptr() for v in kets_l])
        out = np.zeros(2 * len(bras) * nk, dtype=np.float64)
        self._sync_torch()
        check(self.lib.qfe_vec_gram(self.ctx, pb, len(bras), pk, nk, n_amps, dt, as_ptr(out, C.c_double)))
        return out.view(np.complex128).reshape(len(bras), nk).copy()

    def qfim(self, dpsis, psi) -> np.ndarray:
        """Quantum Fisher information matrix g_kl = 4 Re(<d_k psi|d_l psi> - <d_k psi|psi><psi|d_l psi>) from the derivative states,
        the quantity GradientDescentOptimizer assembles from Hadamard-test circuits
        (/root/reference/qat/plugins/qfim_plugin.py:159-199): ONE batched overlap pass over (d_1 psi, ..., d_P psi, psi)."""
        dpsis = list(dpsis)
        full = self.gram(dpsis + [psi])
        P = len(dpsis)
        g, b = full[:P, :P], full[:P, P]  # b_k = <d_k psi|psi>
        return 4.0 * np.real(g - np.outer(b, np.conj(b)))

    def energy_gradient(self, terms_or_plan, dpsis, psi) -> np.ndarray:
        """dE/dtheta_k = 2 Re <d_k psi|H|psi> (/root/reference/qat/fermion/auto_derivatives.py:473-551 evaluates the same number
        term by term through ancilla circuits)."""
        phi = self.apply(terms_or_plan, psi)
        return 2.0 * np.real(self.gram(list(dpsis), [phi])[:, 0])

    def dot(self, a, b) -> complex:
        out = np.zeros(2, dtype=np.float64)
        self._sync_torch()
        check(self.lib.qfe_vec_dot(self.ctx, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), a.numel(), _dtype_code(a), as_ptr(out, C.c_double)))
        return complex(out[0], out[1])

    def scale(self, x, alpha) -> None:
        a = np.array([complex(alpha)], dtype=np.complex128)
        self._sync_torch()
        check(self.lib.qfe_vec_scale(self.ctx, as_ptr(a.view(np.float64), C.c_double), C.c_void_p(x.data_ptr()), x.numel(), _dtype_code(x)))
        self.sync()

    def axpy(self, alpha, x, y) -> None:
        a = np.array([complex(alpha)], dtype=np.complex128)
        self._sync_torch()
        check(self.lib.qfe_vec_axpy(self.ctx, as_ptr(a.view(np.float64), C.c_double), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), x.numel(), _dtype_code(x)))
        self.sync()


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine of a device (default: torch's current device)."""
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    eng = _engines.get(device)
    if eng is None:
        eng = Engine(device)
        _engines[device] = eng
    return eng
