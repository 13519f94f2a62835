"""TEST INFRASTRUCTURE.  A CPU stand-in for ``Engine`` so that the multi-rank host logic of
``ShardedEngine`` (class loop, pairwise exchange, all-reduce, sharded Lanczos) can run under
``gloo`` with world_size > 1 on a machine without GPUs.  The per-class numbers come from the
planner interpreter (tests/csrc/plan_emul.cpp), i.e. from the same plan arrays the CUDA kernel
consumes.  Never imported by the package."""

import ctypes as C
import subprocess
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent


def _lib():
    subprocess.run(["make", "-C", str(HERE / "csrc")], check=True, capture_output=True)
    lib = C.CDLL(str(HERE / "csrc" / "libplan_emul.so"))
    P = C.POINTER
    lib.plan_emul_run.restype = C.c_int
    lib.plan_emul_run.argtypes = [
        C.c_int, C.c_int64, P(C.c_uint64), P(C.c_uint64), P(C.c_double), P(C.c_int32), C.c_int64, P(C.c_double),
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_int64), C.c_int,
    ]
    return lib


def make_terms(nqbits, x, z, c, constant=0.0, owner=None, n_slots=1):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    return SimpleNamespace(
        nqbits=nqbits, x=x, z=np.ascontiguousarray(z, dtype=np.uint64), c=np.ascontiguousarray(c, dtype=np.complex128),
        owner=np.zeros(len(x), dtype=np.int32) if owner is None else np.ascontiguousarray(owner, dtype=np.int32),
        constant=np.ascontiguousarray(np.broadcast_to(np.asarray(constant, dtype=np.complex128), (n_slots,))).copy(), n_slots=n_slots,
    )


class CpuBackend:
    def __init__(self, tile_bits=9):
        self.lib = _lib()
        self.tile_bits = tile_bits

    def plan(self, terms, n_local=None, shard=0, tile_bits=0):
        n_local = terms.nqbits if n_local is None else n_local
        classes = set((terms.x >> np.uint64(n_local)).tolist()) if n_local < 64 else {0}
        if np.any(terms.constant != 0):
            classes.add(0)
        hermitian = bool(np.all(terms.c.imag == 0) and np.all(terms.constant.imag == 0))
        return SimpleNamespace(terms=terms, nqbits=terms.nqbits, n_local=n_local, shard=shard, classes=sorted(int(v) for v in classes), hermitian=hermitian)

    def _run(self, plan, x_hi, bra, ket):
        t = plan.terms
        P = C.POINTER
        out = np.zeros(t.n_slots, dtype=np.complex128)
        y = np.zeros(1 << plan.n_local, dtype=np.complex128)
        stats = np.zeros(4, dtype=np.int64)
        f = lambda a: a.view(np.float64).ctypes.data_as(P(C.c_double))
        bra_a = np.ascontiguousarray(bra.numpy().astype(np.complex128))
        ket_a = np.ascontiguousarray(ket.numpy().astype(np.complex128))
        rc = self.lib.plan_emul_run(
            t.nqbits, len(t.x), t.x.ctypes.data_as(P(C.c_uint64)), t.z.ctypes.data_as(P(C.c_uint64)), f(t.c), t.owner.ctypes.data_as(P(C.c_int32)),
            t.n_slots, f(t.constant), plan.n_local, plan.shard, min(self.tile_bits, plan.n_local), 2048, 512, x_hi, f(bra_a), f(ket_a), f(out), f(y),
            stats.ctypes.data_as(P(C.c_int64)), 0,
        )
        assert rc == 0, rc
        return out, y

    def expectation_class(self, plan, x_hi, bra_shard, ket_shard, part=0, n_parts=1):
        # the interpreter has no notion of sweep subsets: part 0 carries the whole class, the other parts nothing (they still add up)
        if part != 0:
            return 0.0 if plan.terms.n_slots == 1 else np.zeros(plan.terms.n_slots, dtype=np.complex128)
        out, _ = self._run(plan, x_hi, bra_shard, ket_shard)
        return complex(out[0]) if plan.terms.n_slots == 1 else out

    def expectation_class0_host(self, plan, shard_host, shard_dev):
        shard_dev.copy_(torch.from_numpy(shard_host))
        return self.expectation_class(plan, 0, shard_dev, shard_dev) if 0 in plan.classes else 0.0

    def apply_class(self, plan, x_hi, ket_shard, y_shard, accumulate, sync=True):
        _, y = self._run(plan, x_hi, ket_shard, ket_shard)
        y = torch.from_numpy(y).to(y_shard.dtype)
        if accumulate:
            y_shard += y
        else:
            y_shard.copy_(y)

    def dot(self, a, b):
        return complex(torch.vdot(a, b))

    def axpy(self, alpha, x, y):
        y += complex(alpha) * x

    def scale(self, x, alpha):
        x *= complex(alpha)

    def random_state(self, nqbits, dtype=None, seed=0, normalize=True, n_local=None, shard=0):
        n_local = nqbits if n_local is None else n_local
        rng = np.random.default_rng(seed)
        full = rng.standard_normal(1 << nqbits) + 1j * rng.standard_normal(1 << nqbits)
        part = full[shard << n_local : (shard + 1) << n_local]
        if normalize and n_local == nqbits:
            part = part / np.linalg.norm(part)
        return torch.from_numpy(np.ascontiguousarray(part)).to(dtype or torch.complex128)
