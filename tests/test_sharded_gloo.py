"""Multi-rank host logic of ShardedEngine on CPU: world_size 2 and 4 under gloo, compute by the
planner interpreter (tests/cpu_backend.py).  Checks that the x_hi-class loop + pairwise exchanges +
all-reduce reproduce the unsharded oracle numbers."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from threadpoolctl import threadpool_limits

    threadpool_limits(1)  # several ranks share this box's cores: keep BLAS single-threaded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_backend import CpuBackend, make_terms
        from myqlm_fermion_b200.sharded import ShardedEngine
        from oracle import models, spin, transforms

        n = 12  # shards of 2^10 (world 4) or 2^11 amplitudes: at least one 2^9 tile
        one, two = models.synthetic_integrals(n // 2, seed=1234)
        hpq, hpqrs = models.convert_to_h_integrals(one, two)
        x, z, c, k = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.3, "bravyi-kitaev").canonical()
        be = CpuBackend(tile_bits=9)
        se = ShardedEngine(be, n)
        assert se.n_local == n - (world.bit_length() - 1)
        h_terms = make_terms(n, x, z, c, constant=k)
        plan = se.plan(h_terms)
        rng = np.random.default_rng(0)
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        lo, hi = rank << se.n_local, (rank + 1) << se.n_local
        shard = torch.from_numpy(psi[lo:hi].copy())
        want_phi = spin.apply_packed(n, x, z, c, k, psi)
        want_e = np.vdot(psi, want_phi)
        got_e = se.expectation(plan, shard)
        assert abs(got_e - want_e) <= 1e-11 * abs(want_e), (got_e, want_e)
        landing = torch.zeros_like(shard)  # end-to-end form: the shard arrives from host memory, the local class runs during the upload
        got_h = se.expectation_host(plan, psi[lo:hi].copy(), landing)
        assert abs(got_h - want_e) <= 1e-11 * abs(want_e) and torch.equal(landing, shard)
        got_phi = se.apply(plan, shard)
        assert np.max(np.abs(got_phi.numpy() - want_phi[lo:hi])) <= 1e-11 * np.max(np.abs(want_phi))
        assert se.exchanges > 0 and abs(se.dot(shard, shard) - 1.0) < 1e-12

        # ADAPT gradient sweep, pool batched into slots
        pool = [transforms.transform(n, op, 0.0, "jordan-wigner").canonical() for op in models.cluster_ops(4, n)[::5]]
        hx, hz, hc, hk = transforms.transform(n, models.electronic_terms(hpq, hpqrs), 0.0, "jordan-wigner").canonical()
        hj = se.plan(make_terms(n, hx, hz, hc, constant=hk))
        pt = make_terms(
            n, np.concatenate([p[0] for p in pool]), np.concatenate([p[1] for p in pool]), np.concatenate([p[2] for p in pool]),
            constant=0.0, owner=np.concatenate([np.full(len(p[0]), i) for i, p in enumerate(pool)]), n_slots=len(pool),
        )
        g = se.commutator_gradients(hj, se.plan(pt), shard)
        phi_j = spin.apply_packed(n, hx, hz, hc, hk, psi)
        want_g = np.array([2 * np.imag(np.vdot(phi_j, spin.apply_packed(n, p[0], p[1], p[2], 0.0, psi))) for p in pool])
        assert np.max(np.abs(g - want_g)) <= 1e-11 * max(1.0, np.max(np.abs(want_g)))

        # sharded Lanczos on a small Hubbard chain vs dense eigvalsh
        n2 = 12
        t_mat = np.zeros((6, 6))
        for i in range(5):
            t_mat[i, i + 1] = t_mat[i + 1, i] = -1.0
        a, b = models.hubbard_integrals(t_mat, 4.0, 2.0)
        x2, z2, c2, k2 = transforms.transform(n2, models.electronic_terms(a, b), 0.0, "jordan-wigner").canonical()
        se2 = ShardedEngine(be, n2)
        e0, vec, info = se2.ground_state(se2.plan(make_terms(n2, x2, z2, c2, constant=k2)), max_iter=200, tol=1e-12)
        want = np.linalg.eigvalsh(spin.csr_from_packed(n2, x2, z2, c2, k2).toarray()).min()
        assert abs(e0 - want) <= 1e-10 * abs(want), (e0, want, info)
        assert abs(se2.dot(vec, vec) - 1.0) < 1e-8
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_engine_under_gloo(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
