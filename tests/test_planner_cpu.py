"""CPU check of the sweep planner (host C++ of kernel 2): the plan it emits -- tile masks covering
every X mask, tile geometry, compressed z masks, folded coefficients -- is interpreted by a scalar
CPU walker (tests/csrc/plan_emul.cpp) and compared with the oracle's matrix-free formula.
No GPU involved; the CUDA kernel consumes exactly the same plan arrays."""

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import models, spin, transforms

HERE = Path(__file__).resolve().parent
LIB = HERE / "csrc" / "libplan_emul.so"


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", str(HERE / "csrc")], check=True, capture_output=True)
    lib = C.CDLL(str(LIB))
    P = C.POINTER
    lib.plan_emul_run.restype = C.c_int
    lib.plan_emul_run.argtypes = [
        C.c_int, C.c_int64, P(C.c_uint64), P(C.c_uint64), P(C.c_double), P(C.c_int32), C.c_int64, P(C.c_double),
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_int64), C.c_int,
    ]
    return lib


def _run(lib, n, x, z, c, owner, n_slots, constant, n_local, shard, tile_bits, x_hi, bra, ket, max_terms=2048, max_groups=512, half_visit=0):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    c = np.ascontiguousarray(c, dtype=np.complex128)
    owner = np.ascontiguousarray(owner, dtype=np.int32)
    constant = np.ascontiguousarray(constant, dtype=np.complex128)
    out = np.zeros(n_slots, dtype=np.complex128)
    y = np.zeros(1 << n_local, dtype=np.complex128)
    stats = np.zeros(4, dtype=np.int64)
    ptr = lambda a, t: a.view(np.float64).ctypes.data_as(C.POINTER(t)) if a.dtype == np.complex128 else a.ctypes.data_as(C.POINTER(t))
    rc = lib.plan_emul_run(
        n, len(x), ptr(x, C.c_uint64), ptr(z, C.c_uint64), ptr(c, C.c_double), ptr(owner, C.c_int32), n_slots, ptr(constant, C.c_double),
        n_local, shard, tile_bits, max_terms, max_groups, x_hi, ptr(np.ascontiguousarray(bra), C.c_double), ptr(np.ascontiguousarray(ket), C.c_double),
        ptr(out, C.c_double), ptr(y, C.c_double), ptr(stats, C.c_int64), half_visit,
    )
    assert rc == 0, f"plan_emul_run failed with code {rc}"
    return out, y, stats


def _state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)


def _molecule(m, enc, constant=0.0):
    one, two = models.synthetic_integrals(m, seed=1234)
    hpq, hpqrs = models.convert_to_h_integrals(one, two)
    return transforms.transform(2 * m, models.electronic_terms(hpq, hpqrs), constant, enc).canonical()


@pytest.mark.parametrize("enc", ["jordan-wigner", "bravyi-kitaev", "parity"])
@pytest.mark.parametrize("tile_bits", [9, 10, 12])
def test_unsharded_plan_reproduces_apply_and_expectation(emul, enc, tile_bits):
    n = 12
    x, z, c, k = _molecule(6, enc, constant=0.4)
    psi, bra = _state(n, 0), _state(n, 1)
    want_phi = spin.apply_packed(n, x, z, c, k, psi)
    out, y, stats = _run(emul, n, x, z, c, np.zeros(len(x)), 1, [k], n, 0, tile_bits, 0, bra, psi)
    assert np.max(np.abs(y - want_phi)) <= 1e-12 * np.max(np.abs(want_phi))
    assert abs(out[0] - np.vdot(bra, want_phi)) <= 1e-12
    # <psi|H|psi> with the half-visit of Hermitian groups (what the EXPECT kernel does when the bra tile is the ket tile)
    out2, _, _ = _run(emul, n, x, z, c, np.zeros(len(x)), 1, [k], n, 0, tile_bits, 0, psi, psi, half_visit=1)
    want_e = np.vdot(psi, want_phi)
    assert abs(out2[0] - want_e.real) <= 1e-12 * abs(want_e) and abs(want_e.imag) <= 1e-12
    n_groups = len(set(x.tolist())) + (0 if 0 in set(x.tolist()) else 1)  # + identity group if no diagonal term
    assert stats[1] == n_groups and len(x) + 1 <= stats[2] <= len(x) + 1 + n_groups  # + identity term, + even-length padding
    assert 1 <= stats[0] <= n_groups


@pytest.mark.parametrize("split", [2, 3, 16])
def test_ranges_walked_in_slices_like_the_ctas_of_a_small_shard(emul, split):
    """Shards with fewer tiles than SMs: several CTAs stage the same tile and each walks a slice of the groups of every range, entering
    the staged terms through the second header word of the slice's first group.  The interpreter walks the slices the same way."""
    n = 12
    x, z, c, k = _molecule(6, "jordan-wigner", constant=-0.2)
    psi = _state(n, 3)
    want_phi = spin.apply_packed(n, x, z, c, k, psi)
    emul.plan_emul_set_split(split)
    try:
        out, y, _ = _run(emul, n, x, z, c, np.zeros(len(x)), 1, [k], n, 0, 10, 0, psi, psi, half_visit=1)
    finally:
        emul.plan_emul_set_split(1)
    assert np.max(np.abs(y - want_phi)) <= 1e-12 * np.max(np.abs(want_phi))
    assert abs(out[0] - np.vdot(psi, want_phi).real) <= 1e-12 * abs(np.vdot(psi, want_phi))


@pytest.mark.parametrize("n_local", [9, 10])
def test_sharded_classes_recombine(emul, n_local):
    n = 12
    x, z, c, k = _molecule(6, "bravyi-kitaev", constant=-0.3)
    psi = _state(n, 2)
    want_phi = spin.apply_packed(n, x, z, c, k, psi)
    want_e = np.vdot(psi, want_phi)
    g = n - n_local
    phi = np.zeros_like(psi)
    total = 0.0
    for r in range(1 << g):
        for x_hi in range(1 << g):
            own = psi[r << n_local : (r + 1) << n_local]
            partner = psi[(r ^ x_hi) << n_local : ((r ^ x_hi) + 1) << n_local]
            out, y, _ = _run(emul, n, x, z, c, np.zeros(len(x)), 1, [k], n_local, r, 9, x_hi, own, partner)
            phi[r << n_local : (r + 1) << n_local] += y
            total += out[0]
    assert np.max(np.abs(phi - want_phi)) <= 1e-12 * np.max(np.abs(want_phi))
    assert abs(total - want_e) <= 1e-12 * abs(want_e)


def test_wide_masks_and_capacity_splits(emul):
    """X masks wider than a tile (non-zero ket_xor sweeps), complex coefficients, groups larger
    than the per-sweep staging capacity, several operator slots."""
    n = 11
    rng = np.random.default_rng(3)
    T = 300
    x = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
    x[:120] = x[0]  # one big group
    z = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
    c = rng.standard_normal(T) + 1j * rng.standard_normal(T)
    ps = spin.PauliSum(n)
    owner = np.sort(rng.integers(0, 3, size=T)).astype(np.int32)
    # canonical order inside every slot: ascending (x, z), no duplicates
    recs = sorted({(int(o), int(a), int(b)) for o, a, b in zip(owner, x, z) if (a, b) != (0, 0)})
    owner = np.array([r[0] for r in recs], dtype=np.int32)
    x = np.array([r[1] for r in recs], dtype=np.uint64)
    z = np.array([r[2] for r in recs], dtype=np.uint64)
    c = c[: len(recs)]
    consts = np.array([0.5, 0.0, -0.25j])
    psi, bra = _state(n, 4), _state(n, 5)
    out, y, stats = _run(emul, n, x, z, c, owner, 3, consts, n, 0, 9, 0, bra, psi, max_terms=64, max_groups=16)
    want_total = np.zeros_like(psi)
    for s in range(3):
        sel = owner == s
        phi = spin.apply_packed(n, x[sel], z[sel], c[sel], consts[s], psi)
        assert abs(out[s] - np.vdot(bra, phi)) <= 1e-11, s
        want_total += phi
    assert np.max(np.abs(y - want_total)) <= 1e-11
    assert stats[0] > 3  # capacity limits force several sweeps
    del ps


def test_uccsd_pool_batch(emul):
    """Operator pool as one batched term set: every slot's <bra|tau_k|ket> (ADAPT gradient sweep)."""
    n = 10
    pool = [transforms.transform(n, op, 0.0, "jordan-wigner").canonical() for op in models.cluster_ops(4, n)]
    x = np.concatenate([p[0] for p in pool])
    z = np.concatenate([p[1] for p in pool])
    c = np.concatenate([p[2] for p in pool])
    owner = np.concatenate([np.full(len(p[0]), i) for i, p in enumerate(pool)])
    K = len(pool)
    psi, bra = _state(n, 6), _state(n, 7)
    out, _, stats = _run(emul, n, x, z, c, owner, K, np.zeros(K), n, 0, 9, 0, bra, psi, max_terms=1024, max_groups=128)
    for i, p in enumerate(pool):
        want = np.vdot(bra, spin.apply_packed(n, p[0], p[1], p[2], 0.0, psi))
        assert abs(out[i] - want) <= 1e-12, i
    assert stats[1] == K


def test_random_plans_property(emul):
    """Property test over random Pauli sums (hypothesis): any term set, shard geometry, tile size and staging capacity gives a plan
    whose interpretation equals the oracle's matrix-free formula -- wide masks, X bits on register positions, runs of register-row
    Z bits, complex coefficients, several operator slots and the Hermitian half visit included."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(
        n=st.integers(9, 11), tile_bits=st.integers(9, 11), T=st.integers(1, 60), seed=st.integers(0, 10**6), hermitian=st.booleans(),
        sparse_x=st.booleans(), n_slots=st.integers(1, 3), g=st.integers(0, 2), max_terms=st.sampled_from([16, 64, 2048]),
    )
    def run(n, tile_bits, T, seed, hermitian, sparse_x, n_slots, g, max_terms):
        rng = np.random.default_rng(seed)
        x = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
        if sparse_x:  # molecular-like masks: a few X bits, many terms per mask
            x &= rng.integers(0, 1 << n, size=T, dtype=np.uint64) & rng.integers(0, 1 << n, size=T, dtype=np.uint64)
            x[: T // 2] = x[0]
        z = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
        c = rng.standard_normal(T) + (0 if hermitian else 1j * rng.standard_normal(T))
        owner = np.sort(rng.integers(0, n_slots, size=T)).astype(np.int32)
        recs = sorted({(int(o), int(a), int(b)) for o, a, b in zip(owner, x, z) if (a, b) != (0, 0)})
        if not recs:
            return
        owner = np.array([r[0] for r in recs], dtype=np.int32)
        x = np.array([r[1] for r in recs], dtype=np.uint64)
        z = np.array([r[2] for r in recs], dtype=np.uint64)
        c = np.asarray(c[: len(recs)], dtype=np.complex128)
        consts = rng.standard_normal(n_slots) * (rng.random(n_slots) < 0.5)
        n_local = n - g
        tb = min(tile_bits, n_local)
        if tb < 9:
            return
        psi, bra = _state(n, seed % 97), _state(n, seed % 89 + 100)
        same = hermitian and n_slots == 1
        if same:
            bra = psi
        total = np.zeros(n_slots, dtype=np.complex128)
        phi = np.zeros(1 << n, dtype=np.complex128)
        for r in range(1 << g):
            for x_hi in range(1 << g):
                own_b = bra[r << n_local : (r + 1) << n_local]
                own_k = psi[r << n_local : (r + 1) << n_local]
                partner = psi[(r ^ x_hi) << n_local : ((r ^ x_hi) + 1) << n_local]
                half = 1 if (same and x_hi == 0) else 0
                out, y, _ = _run(emul, n, x, z, c, owner, n_slots, consts, n_local, r, tb, x_hi, own_k if half else own_b, partner, max_terms=max_terms,
                                 max_groups=16 if max_terms == 16 else 512, half_visit=half)
                total += out
                phi[r << n_local : (r + 1) << n_local] += y
        want_phi = np.zeros(1 << n, dtype=np.complex128)
        for s in range(n_slots):
            sel = owner == s
            ph = spin.apply_packed(n, x[sel], z[sel], c[sel], consts[s], psi)
            assert abs(total[s] - np.vdot(bra, ph)) <= 1e-10, (s, total[s], np.vdot(bra, ph))
            want_phi += ph
        assert np.max(np.abs(phi - want_phi)) <= 1e-10

    run()


def _run_records(lib, n, x, z, c, n_local, shard, tile_bits, x_hi, bra, ket, half_visit, single_precision=0, max_terms=1920, max_groups=512, max_slots=1999):
    P = C.POINTER
    lib.plan_emul_run_records.restype = C.c_int
    lib.plan_emul_run_records.argtypes = [C.c_int, C.c_int64, P(C.c_uint64), P(C.c_uint64), P(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, P(C.c_double), P(C.c_double), C.c_int, C.c_int, P(C.c_double), P(C.c_int64)]
    x = np.ascontiguousarray(x, dtype=np.uint64)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    c = np.ascontiguousarray(c, dtype=np.complex128)
    bra, ket = np.ascontiguousarray(bra), np.ascontiguousarray(ket)
    out, stats = np.zeros(2), np.zeros(3, dtype=np.int64)
    f64 = lambda a: a.view(np.float64).ctypes.data_as(P(C.c_double))
    # bra is ket (the same buffer) when the caller passes the same array: the interpreter's half visit keys on that, like the kernel's same_tile
    rc = lib.plan_emul_run_records(n, len(x), x.ctypes.data_as(P(C.c_uint64)), z.ctypes.data_as(P(C.c_uint64)), f64(c), n_local, shard, tile_bits, max_terms,
                                   max_groups, max_slots, x_hi, f64(ket) if bra is ket else f64(bra), f64(ket), half_visit, single_precision,
                                   out.ctypes.data_as(P(C.c_double)), stats.ctypes.data_as(P(C.c_int64)))
    assert rc == 0, f"plan_emul_run_records failed with code {rc}"
    return complex(out[0], out[1]), stats


@pytest.mark.parametrize("enc", ["jordan-wigner", "bravyi-kitaev", "parity"])
def test_record_stream_of_the_second_generation_expect_kernel(emul, enc):
    """The record stream tile_expect2_kernel reads from its kernel parameter (plan_host.cpp: build_sweep_records -- range words, one header
    record per group followed by its term records, a closing zero record) interpreted on the CPU the way the kernel walks it: <psi|H|psi>
    with the Hermitian half visit, the transition <phi|H|psi> with full visits, and single-precision coefficients."""
    n = 12
    x, z, c, k = _molecule(6, enc, constant=0.0)
    nz = (x != 0) | (z != 0)
    x, z, c = x[nz], z[nz], c[nz]
    psi, phi = _state(n, 0), _state(n, 1)
    h_psi = spin.apply_packed(n, x, z, c, 0.0, psi)
    want_e, want_t = np.vdot(psi, h_psi), np.vdot(phi, h_psi)
    for tile_bits in (9, 11):
        e, stats = _run_records(emul, n, x, z, c, n, 0, tile_bits, 0, psi, psi, half_visit=1)
        assert abs(e - want_e.real) <= 1e-12 * abs(want_e)  # (the full visits of the groups without a row-block X bit leave rounding in Im)
        t, _ = _run_records(emul, n, x, z, c, n, 0, tile_bits, 0, phi, psi, half_visit=1)  # another bra: no half visit whatever the flag says
        assert abs(t - want_t) <= 1e-12 * abs(want_e)
        e32, _ = _run_records(emul, n, x, z, c, n, 0, tile_bits, 0, psi, psi, half_visit=1, single_precision=1)
        assert abs(e32 - want_e.real) <= 1e-6 * abs(want_e) and e32 != e
        assert stats[0] >= 1 and stats[2] <= 48
    # sharded: the classes of a register split over 4 ranks recombine (the pair ranks' split of a class is host logic, tests/test_sharded_gloo.py)
    n_local, total = 10, 0.0
    for r in range(4):
        for x_hi in range(4):
            own = psi[r << n_local : (r + 1) << n_local]
            partner = own if x_hi == 0 else psi[(r ^ x_hi) << n_local : ((r ^ x_hi) + 1) << n_local]
            v, _ = _run_records(emul, n, x, z, c, n_local, r, 9, x_hi, own, partner, half_visit=1)
            total += v
    assert abs(total - want_e) <= 1e-12 * abs(want_e)


def test_planner_keeps_headers_plus_terms_within_the_record_capacity(emul):
    """max_slots: terms + one header per group of a sweep stay within the record stream of the kernel parameter (the interpreter returns
    code 5 otherwise); with a tight bound the plan needs more sweeps but gives the same number."""
    n = 12
    x, z, c, k = _molecule(6, "jordan-wigner")
    nz = (x != 0) | (z != 0)
    x, z, c = x[nz], z[nz], c[nz]
    psi = _state(n, 4)
    want = np.vdot(psi, spin.apply_packed(n, x, z, c, 0.0, psi)).real
    loose, s_loose = _run_records(emul, n, x, z, c, n, 0, 10, 0, psi, psi, 1, max_slots=1999)
    tight, s_tight = _run_records(emul, n, x, z, c, n, 0, 10, 0, psi, psi, 1, max_slots=120)
    assert abs(loose - want) <= 1e-12 * abs(want) and abs(tight - want) <= 1e-12 * abs(want)
    assert s_tight[0] > s_loose[0] and s_tight[1] - 1 <= 120 < s_loose[1]
