"""Statevector sharded over the ranks of one node: one process per GPU.

On the GPU box (``torch.distributed`` backend "nccl") ``ShardedEngine`` is a ctypes shim: the class loop, the pairwise
``ncclSend``/``ncclRecv`` exchanges, the overlap of the next exchange with the current sweeps, the scalar all-reduce and the
sharded Lanczos all live behind the C ABI (``qfe_*_sharded`` in include/qfe.h, csrc/sharded.cu), on an NCCL communicator the
``qfe_ctx`` owns; torch.distributed only carries the 128-byte NCCL id to the ranks.  With any other backend (gloo in the CPU
tests, where the compute backend is the plan interpreter of tests/cpu_backend.py) the same class loop runs here in Python over
``torch.distributed`` point-to-point calls: that mirror is what the world-size-2/4 CPU tests cover.

Rank r holds the 2^(n-g) amplitudes whose g most significant index bits (= qubits 0..g-1, because
qubit 0 is the MSB: /root/reference/qat/fermion/hamiltonians.py:231-235) equal r.  A Pauli term whose
X mask has high part x_hi maps shard r onto shard r ^ x_hi, so the terms fall into at most 2^g
*classes*; class 0 is local, every other class needs ONE pairwise exchange of the shard with rank
r ^ x_hi (``isend``/``irecv``), after which all sweeps of that class run locally on
(own shard, partner shard).  Scalars are combined with one all-reduce.

The reference has no distributed path at all; this module is new structure around the same numbers.
"""

from __future__ import annotations

import ctypes as C

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


class ShardedEngine:
    """Drives a per-rank compute backend (``Engine`` on the GPU box) over the x_hi classes.

    ``backend`` must provide ``plan(terms, n_local, shard)``, ``expectation_class``, ``apply_class``,
    ``dot``, ``axpy``, ``scale`` and the state constructors; ``Engine`` does.
    """

    def __init__(self, backend, nqbits: int, group=None):
        dist = _dist()
        if not dist.is_initialized():
            raise RuntimeError("ShardedEngine needs an initialised torch.distributed process group")
        self.backend = backend
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        g = self.world.bit_length() - 1
        if 1 << g != self.world:
            raise ValueError(f"world size {self.world} is not a power of two")
        if g > nqbits:
            raise ValueError("more ranks than amplitudes")
        self.g = g
        self.nqbits = nqbits
        self.n_local = nqbits - g
        self._partner = None
        self._work = None
        self._exchanges = 0
        self._exchanged_bytes = 0
        # native path: the backend is the CUDA engine and the ranks talk NCCL -> everything below the call is C (csrc/sharded.cu)
        self.native = hasattr(backend, "lib") and hasattr(backend, "ctx") and dist.get_backend(group) == "nccl"
        if self.native:
            self._init_native_comm()

    def _init_native_comm(self) -> None:
        """Rank 0 draws an NCCL unique id (qfe_comm_unique_id), torch.distributed hands it to every rank, every rank joins
        (qfe_comm_init).  A context keeps its communicator: a second ShardedEngine on the same engine reuses it."""
        import torch

        from ._lib import check

        dist = _dist()
        be = self.backend
        if getattr(be, "_comm_world", None) is not None:
            if be._comm_world != (self.rank, self.world):
                raise RuntimeError(f"the engine's communicator is rank {be._comm_world[0]} of {be._comm_world[1]}, this group is rank {self.rank} of {self.world}")
            return
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            check(be.lib.qfe_comm_unique_id(uid.ctypes.data_as(C.POINTER(C.c_uint8))))
        t = torch.from_numpy(uid).to(f"cuda:{be.device}")
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(t, src=src, group=self.group)
        uid = t.cpu().numpy()
        check(be.lib.qfe_comm_init(be.ctx, uid.ctypes.data_as(C.POINTER(C.c_uint8)), self.rank, self.world))
        be._comm_world = (self.rank, self.world)

    @property
    def exchanges(self) -> int:
        if not self.native:
            return self._exchanges
        v = C.c_int64()
        self.backend.lib.qfe_comm_info(self.backend.ctx, None, None, None, C.byref(v), None)
        return v.value

    @property
    def exchanged_bytes(self) -> int:
        if not self.native:
            return self._exchanged_bytes
        v = C.c_int64()
        self.backend.lib.qfe_comm_info(self.backend.ctx, None, None, None, None, C.byref(v))
        return v.value

    def comm_info(self) -> dict:
        if not self.native:
            return {"native": False, "rank": self.rank, "nranks": self.world}
        r, n, ver = C.c_int(), C.c_int(), C.c_int()
        self.backend.lib.qfe_comm_info(self.backend.ctx, C.byref(r), C.byref(n), C.byref(ver), None, None)
        return {"native": True, "rank": r.value, "nranks": n.value, "nccl_version": ver.value}

    # ---- plumbing ------------------------------------------------------------------------------
    def plan(self, terms):
        return self.backend.plan(terms, n_local=self.n_local, shard=self.rank)

    def _partner_buffer(self, like):
        import torch

        if self._partner is None or self._partner.shape != like.shape or self._partner.dtype != like.dtype or self._partner.device != like.device:
            self._partner = torch.empty_like(like)
        return self._partner

    def _partner_native(self, like):
        """Device scratch for the partner shards: two shards when the device has room (the exchange of the next class then overlaps the
        sweeps of the current one), else one.  Allocated once per shard shape."""
        import torch

        p = self._partner
        if p is None or p.dtype != like.dtype or p.device != like.device or p.numel() % like.numel() or p.numel() == 0:
            self._partner = None
            free_b = torch.cuda.mem_get_info(like.device)[0]
            n_buf = 2 if free_b > 2 * like.numel() * like.element_size() + (8 << 30) else 1
            self._partner = torch.empty(n_buf * like.numel(), dtype=like.dtype, device=like.device)
        return self._partner

    def _native_args(self, like):
        from .engine import _dtype_code

        buf = self._partner_native(like)
        return C.c_void_p(buf.data_ptr()), C.c_int64(buf.numel() * buf.element_size()), _dtype_code(like)

    def release_buffers(self) -> None:
        self._partner = None
        self._work = None

    def exchange(self, shard, x_hi: int):
        """Partner shard (rank ^ x_hi) of the vector ``shard`` belongs to; pairwise send/recv."""
        if x_hi == 0:
            return shard
        dist = _dist()
        import torch

        peer = self.rank ^ x_hi
        buf = self._partner_buffer(shard)
        # complex tensors travel as their real view (NCCL has no complex dtype)
        send_t, recv_t = torch.view_as_real(shard), torch.view_as_real(buf)
        peer_global = dist.get_global_rank(self.group, peer) if self.group is not None else peer
        ops = [dist.P2POp(dist.isend, send_t, peer_global, group=self.group), dist.P2POp(dist.irecv, recv_t, peer_global, group=self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if shard.is_cuda:
            torch.cuda.current_stream(shard.device).synchronize()
        self._exchanges += 1
        self._exchanged_bytes += shard.numel() * shard.element_size()
        return buf

    def allreduce(self, values: np.ndarray) -> np.ndarray:
        import torch

        if self.native:
            from ._lib import as_ptr, check

            v = np.ascontiguousarray(values, dtype=np.float64).copy()
            check(self.backend.lib.qfe_allreduce_sum(self.backend.ctx, as_ptr(v, C.c_double), v.size))
            return v
        dist = _dist()
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    # ---- kernel 2 on shards ----------------------------------------------------------------------
    def expectation_host(self, plan, shard_host, ket_shard):
        """<psi|H|psi> with this rank's shard arriving from HOST memory: it is uploaded into ``ket_shard`` behind the sweeps of the
        local class (Engine.expectation_class0_host); the other classes then run on the device copy as in ``expectation``."""
        if plan.terms.n_slots != 1:
            raise ValueError("expectation_host needs a single-operator term set")
        if self.native:
            from ._lib import as_ptr, check

            be = self.backend
            be._check_state(ket_shard, 1 << self.n_local)
            if shard_host.size != ket_shard.numel() or not shard_host.flags.c_contiguous or shard_host.itemsize != ket_shard.element_size():
                raise ValueError("host shard must be a contiguous vector of 2^n_local amplitudes of the device buffer's dtype")
            pp, pb, dt = self._native_args(ket_shard)
            be._sync_torch()
            out = np.zeros(2, dtype=np.float64)
            check(be.lib.qfe_expectation_sharded_host(be.ctx, plan.handle, C.c_void_p(shard_host.ctypes.data), C.c_void_p(ket_shard.data_ptr()), pp, pb, dt,
                                                      as_ptr(out, C.c_double)))
            return complex(out[0], out[1])
        local = self.backend.expectation_class0_host(plan, shard_host, ket_shard)
        return self.expectation(plan, ket_shard, _class0=local)

    def expectation(self, plan, ket_shard, bra_shard=None, _class0=None):
        """<bra|H|ket> of the whole register (identical on every rank)."""
        bra = ket_shard if bra_shard is None else bra_shard
        n_slots = plan.terms.n_slots
        if self.native and _class0 is None:
            from ._lib import as_ptr, check

            be = self.backend
            be._check_state(ket_shard, 1 << self.n_local)
            be._check_state(bra, 1 << self.n_local)
            pp, pb, dt = self._native_args(ket_shard)
            be._sync_torch()
            out = np.zeros(2 * n_slots, dtype=np.float64)
            check(be.lib.qfe_expectation_sharded(be.ctx, plan.handle, C.c_void_p(bra.data_ptr()) if bra_shard is not None else None,
                                                 C.c_void_p(ket_shard.data_ptr()), pp, pb, dt, as_ptr(out, C.c_double)))
            vals = out.view(np.complex128)
            return complex(vals[0]) if n_slots == 1 else vals.copy()
        acc = np.zeros(n_slots, dtype=np.complex128)
        if _class0 is not None:
            acc[0] += _class0  # the local class was evaluated during the upload
        # <psi|H|psi> with every class Hermitian: the shard pairs (r, r^x_hi) and (r^x_hi, r) of a class give conjugate numbers.  The
        # two ranks of a pair split the sweeps of the class between them (each with its own shard as the bra) and count their half
        # twice: every rank does half of the work of every non-local class.
        split = bra_shard is None and n_slots == 1 and getattr(plan, "hermitian", False)
        for x_hi in plan.classes:  # the class list is the same on every rank (it depends on the X masks only)
            if x_hi == 0 and _class0 is not None:
                continue
            if x_hi != 0 and split:
                pivot = (x_hi & -x_hi).bit_length() - 1
                partner = self.exchange(ket_shard, x_hi)
                half = self.backend.expectation_class(plan, x_hi, bra, partner, part=(self.rank >> pivot) & 1, n_parts=2)
                acc += 2.0 * np.real(np.atleast_1d(half))
                continue
            partner = self.exchange(ket_shard, x_hi)
            acc += np.atleast_1d(self.backend.expectation_class(plan, x_hi, bra, partner))
        tot = self.allreduce(acc.view(np.float64)).view(np.complex128)
        return complex(tot[0]) if n_slots == 1 else tot

    def energy(self, plan, ket_shard) -> float:
        return float(np.real(self.expectation(plan, ket_shard)))

    def apply(self, plan, ket_shard, out_shard=None):
        """This rank's shard of H|ket>."""
        import torch

        y = torch.empty_like(ket_shard) if out_shard is None else out_shard
        if self.native:
            from ._lib import check

            be = self.backend
            be._check_state(ket_shard, 1 << self.n_local)
            be._check_state(y, 1 << self.n_local)
            pp, pb, dt = self._native_args(ket_shard)
            be._sync_torch()
            check(be.lib.qfe_apply_sharded(be.ctx, plan.handle, C.c_void_p(ket_shard.data_ptr()), C.c_void_p(y.data_ptr()), pp, pb, dt))
            be.sync()
            return y
        first = True
        for x_hi in plan.classes:
            partner = self.exchange(ket_shard, x_hi)
            self.backend.apply_class(plan, x_hi, partner, y, accumulate=not first)
            first = False
        if first:
            y.zero_()
        return y

    def commutator_gradients(self, h_plan, pool_plan, psi_shard, scratch=None) -> np.ndarray:
        """g_k = -i<psi|[H, tau_k]|psi> = 2 Im <H psi|tau_k|psi> (adapt_vqe.py:181-185), sharded."""
        if self.native:
            import torch

            from ._lib import as_ptr, check

            be = self.backend
            be._check_state(psi_shard, 1 << self.n_local)
            phi = torch.empty_like(psi_shard) if scratch is None else scratch
            pp, pb, dt = self._native_args(psi_shard)
            be._sync_torch()
            g = np.zeros(pool_plan.terms.n_slots, dtype=np.float64)
            check(be.lib.qfe_commutator_gradients_sharded(be.ctx, h_plan.handle, pool_plan.handle, C.c_void_p(psi_shard.data_ptr()), C.c_void_p(phi.data_ptr()),
                                                          pp, pb, dt, as_ptr(g, C.c_double)))
            return g
        phi = self.apply(h_plan, psi_shard, out_shard=scratch)
        vals = np.atleast_1d(self.expectation(pool_plan, psi_shard, bra_shard=phi))
        return 2.0 * vals.imag

    # ---- vector helpers ----------------------------------------------------------------------------
    def dot(self, a_shard, b_shard) -> complex:
        v = np.array([self.backend.dot(a_shard, b_shard)], dtype=np.complex128)
        return complex(self.allreduce(v.view(np.float64)).view(np.complex128)[0])

    def random_state(self, dtype=None, seed: int = 0):
        psi = self.backend.random_state(self.nqbits, dtype=dtype, seed=seed, normalize=False, n_local=self.n_local, shard=self.rank)
        nrm = self.dot(psi, psi).real
        self.backend.scale(psi, 1.0 / np.sqrt(nrm))
        return psi

    def product_state(self, alpha_beta, dtype=None):
        return self.backend.product_state(self.nqbits, alpha_beta, dtype=dtype, n_local=self.n_local, shard=self.rank)

    # ---- kernel 3 on shards ------------------------------------------------------------------------
    def ground_state(self, plan, v0_shard=None, max_iter: int = 300, tol: float = 1e-10, dtype=None, want_vector: bool = True, seed: int = 0):
        """Lanczos on the sharded apply; same recurrence as ``qfe_lanczos`` with the dots all-reduced.
        Returns (E0, psi0 shard or None, info)."""
        if self.native:
            import torch

            from ._lib import check

            be = self.backend
            v = self.random_state(dtype=dtype, seed=seed) if v0_shard is None else v0_shard.clone()
            be._check_state(v, 1 << self.n_local)
            if self._work is None or self._work.numel() != 3 * v.numel() or self._work.dtype != v.dtype:
                self._work = None
                self._work = torch.empty(3 * v.numel(), dtype=v.dtype, device=v.device)  # kept: no allocation per call, none per iteration
            pp, pb, dt = self._native_args(v)
            be._sync_torch()
            ev, it, res = C.c_double(), C.c_int(), C.c_double()
            check(be.lib.qfe_lanczos_sharded(be.ctx, plan.handle, C.c_void_p(v.data_ptr()), C.c_void_p(self._work.data_ptr()), pp, pb, max_iter, tol, dt,
                                             1 if want_vector else 0, C.byref(ev), C.byref(it), C.byref(res)))
            return ev.value, (v if want_vector else None), {"iterations": it.value, "residual": res.value}
        from scipy.linalg import eigh_tridiagonal

        be = self.backend
        start = self.random_state(dtype=dtype, seed=seed) if v0_shard is None else v0_shard.clone()
        nrm = np.sqrt(self.dot(start, start).real)
        if nrm == 0.0:
            raise ValueError("start vector has zero norm")
        be.scale(start, 1.0 / nrm)
        alpha, beta = [], []
        theta, y, res, m_done = 0.0, None, 0.0, 0

        def solve(m):
            a = np.array(alpha[:m])
            if m == 1:
                return a[0], np.ones(1)
            w, v = eigh_tridiagonal(a, np.array(beta[: m - 1]), select="i", select_range=(0, 0))
            return float(w[0]), v[:, 0]

        for pass_no in range(2 if want_vector else 1):
            cur = start.clone()
            prev = None
            out = None
            if pass_no == 1:
                out = cur.clone()
                be.scale(out, y[0])
            m_target = max_iter if pass_no == 0 else m_done
            for j in range(m_target):
                w = self.apply(plan, cur)
                if pass_no == 0:
                    a_j = self.dot(cur, w).real
                    alpha.append(a_j)
                else:
                    a_j = alpha[j]
                be.axpy(-a_j, cur, w)
                if prev is not None:
                    be.axpy(-beta[j - 1], prev, w)
                if pass_no == 0:
                    b_j = np.sqrt(max(self.dot(w, w).real, 0.0))
                    beta.append(b_j)
                    m = j + 1
                    m_done = m
                    tiny = b_j <= 1e-14 * max(1.0, abs(a_j))
                    if m % 5 == 0 or m == max_iter or tiny:
                        theta, y = solve(m)
                        res = abs(b_j * y[m - 1])
                        if res <= tol * max(1.0, abs(theta)) or tiny:
                            break
                else:
                    b_j = beta[j]
                    if j + 1 >= m_done:
                        break
                if b_j == 0.0:
                    break
                be.scale(w, 1.0 / b_j)
                prev, cur = cur, w
                if pass_no == 1:
                    be.axpy(y[j + 1], cur, out)
            if pass_no == 0:
                theta, y = solve(m_done)
                res = abs(beta[m_done - 1] * y[m_done - 1])
        return theta, (out if want_vector else None), {"iterations": m_done, "residual": res}
