#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: <psi|H|psi> of a synthetic molecular
Jordan-Wigner Hamiltonian (BASELINE.json config 4 family) in terms*amps/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU algorithm

One "step" = one evaluation of <psi|H|psi> over the whole (sharded) statevector.
N=1: 32 qubits (random 8-fold-symmetric integrals on 32 spin-orbitals, 94 848 Pauli terms, complex128,
68.7 GB state).  N=2/4/8: 32/33/34 qubits (a 2^31-amplitude shard = 34.4 GB per GPU), state sharded by its
high qubits, one process per GPU (torchrun), NCCL send/recv per x_hi class.  Prints ONE JSON line on rank 0.

Order of a run (sized so that the driver's --steps 20 --warmup 5 ends inside its per-N limit): the W warm-up
steps come first and the end-to-end leg lives inside them -- one warm-up through the HOST-buffer call, then up
to two timed ones (`e2e`), the rest device-resident -- then the K timed device-resident steps (`value`).  After
the first step a projected wall time goes to stderr.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "<psi|H|psi> terms*amps/s"
UNIT = "terms*amps/s"
BYTES_PER_SWEEP_AMP = {"c128": 16, "c64": 8}  # every amplitude is read once per sweep


def qubits_for(n_gpus: int) -> int:
    # BASELINE.json: 32 qubits on 1 B200 ... 34 qubits on 8 B200; 2 and 4 GPUs keep the 8-GPU shard (2^31 amplitudes per GPU)
    return {1: 32, 2: 32, 4: 33, 8: 34}.get(n_gpus, 32)


def ncu_traffic(n_local: int, dtype: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu --set full capture recorded in
    profiles/ncu_traffic.json (written by tools/ncu_summary.py from the .ncu-rep of the round); None where no capture exists."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        for row in json.loads(p.read_text())["captures"]:
            if row["n_local"] == n_local and row["dtype"] == dtype:
                return float(row["dram_bytes_read"]) + float(row["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port), bounded sample, all host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(n: int, workers: int, steps: int, warmup: int):
    """The reference's CPU algorithm on an n-qubit member of the bench generator: get_matrix(sparse=True) ONCE (build_s), then
    `steps` evaluations vdot(psi, H psi) on the cached matrix (eval_s each).  Two rates come back: `value_cached` counts the
    evaluations only (what a VQE loop on the reference pays per energy once the matrix exists) and `value_rebuild` charges every
    evaluation with a build (SpinHamiltonian.get_matrix recomputes the matrix on every call, hamiltonians.py:163-211)."""
    from concurrent.futures import ProcessPoolExecutor

    from oracle import baseline

    triples, const = baseline.sample_terms(n)
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    pool = ProcessPoolExecutor(workers) if workers > 1 else None
    try:
        t0 = time.perf_counter()
        mat = baseline.reference_matrix(n, triples, const, workers, pool)
        build_s = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.shutdown()
    for _ in range(warmup):
        np.vdot(psi, mat.dot(psi))
    t0 = time.perf_counter()
    for _ in range(steps):
        e = np.vdot(psi, mat.dot(psi))
    eval_s = (time.perf_counter() - t0) / steps
    T = len(triples)
    work = T * float(1 << n)
    return {"T": T, "n": n, "build_s": build_s, "eval_s": eval_s, "value_cached": work / eval_s, "value_rebuild": work / (build_s + eval_s),
            "energy": float(np.real(e)), "nnz": int(mat.nnz)}


def cpu_baseline_dict(r, cores):
    return {
        "value": r["value_cached"], "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{r['n']}-qubit member of the same generator ({r['T']} Pauli terms, nnz {r['nnz']}): oracle port of get_matrix(sparse=True) built once "
                  f"(term list split over {cores} processes), then vdot(psi, H psi) per step on the cached CSR matrix",
        "build_s": r["build_s"], "eval_s": r["eval_s"], "value_cached_matrix": r["value_cached"], "value_rebuild_per_step": r["value_rebuild"],
        "note": "value = evaluations on the cached matrix only (the favourable reading for the CPU path); value_rebuild_per_step charges "
                "every evaluation with a get_matrix() build, which is what calling SpinHamiltonian.get_matrix per energy costs",
    }


def cpu_sample_qubits(cores: int) -> int:
    return 14 if cores >= 8 else 12


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import baseline

    cores = baseline.usable_cores()
    n = cpu_sample_qubits(cores)
    r = cpu_sample(n, cores, max(1, args.steps), max(0, args.warmup))
    cfg = bench_config(args.gpus, qubits_for(args.gpus), None, None)
    cfg["workload"] += f"; THIS ARM RAN A BOUNDED SAMPLE: the {n}-qubit member of the same generator ({r['T']} Pauli terms) on {cores} host cores"
    cfg["sample_n_qubits"] = n
    cfg["sample_pauli_terms"] = r["T"]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value_cached"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["eval_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic", "config": cfg, "cpu_baseline": cpu_baseline_dict(r, cores),
        "e2e": {"value": r["value_cached"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(n_gpus, n, T, plan):
    cfg = {
        "workload": f"config4 family: random 8-fold-symmetric integrals on {n} spin-orbitals (seed 1234), Jordan-Wigner, <psi|H|psi> on a random complex state"
        + ("" if n_gpus == 1 else f"; {n} qubits on {n_gpus} GPUs = one 2^{n - (n_gpus.bit_length() - 1)}-amplitude complex128 shard per GPU"),
        "n_qubits": n, "n_gpus": n_gpus, "sharding": f"statevector split by its {max(n_gpus.bit_length() - 1, 0)} high qubits, one process per GPU",
        "l2": "inputs larger than L2 (each sweep streams the whole shard, >= 34 GB)",
    }
    if T is not None:
        cfg["pauli_terms"] = T
    if plan is not None:
        cfg.update(x_masks=plan.n_groups, sweeps_per_step=plan.n_sweeps, tile_bits=plan.tile_bits, x_hi_classes=plan.n_classes)
    return cfg


# ------------------------------------------------------------------------------------------------
def main():
    t_proc0 = time.perf_counter()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qfe", choices=["qfe", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the register size (development only)")
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (development only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (development only)")
    ap.add_argument("--cpu-sample-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.cpu_sample_only:  # child of the own arm: the reference arm's CPU sample, on the host cores, while the GPU warms up
        from oracle import baseline

        cores = baseline.usable_cores()
        print("CPU_SAMPLE " + json.dumps(cpu_baseline_dict(cpu_sample(cpu_sample_qubits(cores), cores, 20, 2), cores)), flush=True)
        return
    cpu_child = None
    if not args.no_cpu and int(os.environ.get("RANK", "0")) == 0 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        cpu_child = subprocess.Popen([sys.executable, str(Path(__file__).resolve()), "--cpu-sample-only"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    import torch
    import torch.distributed as dist

    import myqlm_fermion_b200 as qfe
    from myqlm_fermion_b200 import workloads
    from myqlm_fermion_b200.sharded import ShardedEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 prints exactly one JSON line: keep NCCL's version banner (written to stdout when the communicator is created) off it
        sys.stdout.flush()
        saved_fd, null_fd = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null_fd, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
            os.close(null_fd)

    n = args.qubits or qubits_for(world)
    tdtype = torch.complex128 if args.dtype == "c128" else torch.complex64
    elt = BYTES_PER_SWEEP_AMP[args.dtype]

    eng = qfe.get_engine(local_rank)
    eng.use_torch_stream()  # kernels on torch's current stream, so torch.cuda.Event brackets them

    # ---- kernel 1: integrals -> packed Pauli terms (on every rank; term tables are replicated) ----
    hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
    t0 = time.perf_counter()
    terms = eng.terms_from_integrals(hpq, hpqrs, encoding="jordan-wigner")
    eng.sync()
    t_terms = time.perf_counter() - t0
    del hpqrs
    T = terms.n_terms

    if world > 1:
        se = ShardedEngine(eng, n)
        plan = se.plan(terms)
        psi = se.random_state(dtype=tdtype, seed=0)
        step = lambda: se.expectation(plan, psi)
    else:
        se = None
        plan = eng.plan(terms, tile_bits=args.tile_bits)
        psi = eng.random_state(n, dtype=tdtype, seed=0)
        step = lambda: eng.expectation(plan, psi)
    n_local = plan.n_local
    shard_bytes = (1 << n_local) * elt
    amps_f = float(1 << n)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        ev0.record()
        val = None
        for _ in range(steps):
            val = fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = eng.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, launches = float(tmax[0]), int(t[1])
        return ms, launches, val

    def stamp(what):
        if rank == 0:
            print(f"[bench] t={time.perf_counter() - t_proc0:6.1f} s  {what}", file=sys.stderr, flush=True)

    stamp(f"state ready: {n} qubits, {T} Pauli terms, {plan.n_sweeps} sweeps per step and rank")

    # ---- end-to-end leg preparation: every rank stages its shard in pinned HOST memory.  Page-locking tens of GB takes a while: it runs
    # in a thread (a plain cudaHostAlloc through ctypes, which releases the GIL) behind the device-resident warm-up steps ----
    e2e_skip = None
    host_np = None
    pinned = False
    pin_thread = None
    pin_box = {}
    if not args.no_e2e:
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        ok = not (avail is not None and avail < 1.15 * shard_bytes * local_world)
        if world > 1:  # one decision for all ranks: the end-to-end leg contains collectives
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if not ok:
            avail = avail or 0
            e2e_skip = f"host memory: {avail / 2**30:.0f} GiB available, {shard_bytes * local_world / 2**30:.0f} GiB needed for the shards of {local_world} ranks"
        else:
            import ctypes
            import threading

            def pin():
                try:
                    rt = ctypes.CDLL("libcudart.so.12")
                    ptr = ctypes.c_void_p()
                    rt.cudaSetDevice(local_rank)
                    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(shard_bytes), ctypes.c_uint(0))
                    if rc == 0 and ptr.value:
                        pin_box["ptr"] = ptr.value
                except OSError:
                    pass

            pin_thread = threading.Thread(target=pin, daemon=True)
            pin_thread.start()

    # ---- warm-up (W steps).  The end-to-end leg is part of it: the last up-to-three warm-up steps go through the host-buffer call, one
    # untimed and up to two timed (`e2e`); the ones before run device-resident ----
    def project(first_step_s, what):
        if rank == 0:
            total = (time.perf_counter() - t_proc0) + first_step_s * (args.warmup + args.steps - 1) + 15.0
            print(f"[bench] first {what} step {first_step_s:.1f} s -> projected wall time of this run ~{total:.0f} s "
                  f"({args.warmup} warm-up + {args.steps} timed steps)", file=sys.stderr, flush=True)

    e2e = None
    want_e2e = pin_thread is not None
    e2e_warm = (1 if args.warmup >= 2 else 0) if want_e2e else 0
    e2e_timed = max(1, min(2, args.warmup - e2e_warm)) if want_e2e else 0
    dev_warm = max(0, args.warmup - e2e_warm - e2e_timed)
    first_timed = False
    for i in range(dev_warm):
        t0 = time.perf_counter()
        step()
        if not first_timed:
            torch.cuda.synchronize()
            project(time.perf_counter() - t0, "device-resident")
            first_timed = True
    if dev_warm:
        stamp(f"{dev_warm} device-resident warm-up steps done")

    if want_e2e:
        pin_thread.join()
        if "ptr" in pin_box:
            import ctypes

            raw = (ctypes.c_byte * shard_bytes).from_address(pin_box["ptr"])
            host_np = np.frombuffer(raw, dtype=np.complex128 if args.dtype == "c128" else np.complex64)
            pinned = True
        else:
            host_np = np.empty(1 << n_local, dtype=np.complex128 if args.dtype == "c128" else np.complex64)
        torch.from_numpy(host_np).copy_(psi)  # page-locked memory: a plain DMA whatever torch believes about the tensor
        torch.cuda.synchronize()
        stamp(f"shard staged in {'pinned' if pinned else 'pageable'} host memory ({shard_bytes / 2**30:.0f} GiB)")
        if world == 1:
            # the user-facing call uploads into a landing buffer of the library; when the device cannot hold a second state the sharded
            # form of the same call uploads into the caller's buffer instead (same pipeline, same kernels)
            free_b = torch.cuda.mem_get_info(dev)[0]
            if free_b > shard_bytes + (4 << 30):
                e2e_call, e2e_step = "qfe_expectation_host", (lambda: eng.expectation_host(plan, host_np))
            else:
                e2e_call, e2e_step = "qfe_expectation_class0_host", (lambda: eng.expectation_class0_host(plan, host_np, psi))
        else:
            e2e_call = "qfe_expectation_sharded_host"

            def e2e_step():  # the shard is uploaded into psi inside the library call, behind the sweeps of the local class
                return se.expectation_host(plan, host_np, psi)

        if e2e_warm:
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            if not first_timed:
                project(time.perf_counter() - t0, "end-to-end")
                first_timed = True
        ms2, _, energy2 = timed(e2e_step, e2e_timed)
        e2e = {
            "value": T * amps_f * e2e_timed / (ms2 / 1e3), "unit": UNIT, "h2d_bytes_per_step": shard_bytes * world, "d2h_bytes_per_step": 16 * world,
            "steps": e2e_timed, "ms_per_step": ms2 / e2e_timed, "host_memory": "pinned" if pinned else "pageable", "call": e2e_call,
            "note": "timed inside the warm-up phase, after one untimed step through the same call",
        }
        stamp("end-to-end leg done")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, energy = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    sec = ms / 1e3
    value = T * amps_f * args.steps / sec
    if e2e is not None:
        e2e["energy_matches_device"] = bool(abs(energy2 - energy) <= 1e-10 * abs(energy))

    # ---- roofline of the dominant kernel (EXPECT sweep): HBM, 16 B (c128) per amplitude per sweep ----
    peak, peak_src = measured_peak()
    sweeps_per_step = plan.n_sweeps  # per rank; every rank launches the same number of sweeps over its own shard
    sweep_launches = sweeps_per_step * args.steps
    avg_launch_s = sec / max(sweep_launches, 1)  # upper bound: includes the (tiny) reduce kernels and, for N>1, the exchanges
    achieved = shard_bytes / avg_launch_s / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(n_local, args.dtype),
        "kernel": "tile_expect2_kernel (real-coefficient EXPECT sweeps over full tiles: every sweep of this workload)", "peak_source": peak_src,
        "bytes_per_launch": shard_bytes, "avg_launch_ms": 1e3 * avg_launch_s,
        "terms_per_sweep": T / max(sweeps_per_step, 1),
        "note": "each sweep evaluates terms_per_sweep Pauli terms per amplitude from one HBM pass, so the kernel is issue/shared-memory bound by design; "
        "equiv_gbs_one_sweep_per_xmask is the HBM rate a one-pass-per-X-mask schedule would need for the same throughput",
        "equiv_gbs_one_sweep_per_xmask": plan.n_groups * shard_bytes * args.steps / sec / 1e9,
    }
    # the resource that actually binds the sweep kernel: shared-memory bandwidth.  Algorithmic shared-memory bytes per step =
    # one 16 B partner read per visited (X mask, amplitude) -- Hermitian X masks visit half the amplitudes -- plus one write and
    # one read of every staged amplitude per sweep.  Peak = 128 B/clk/SM (measured: tools/ubench/smem_paths.cu) x SMs x SM clock.
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    smem_bytes = shard_bytes * (0.5 * plan.half_groups + plan.full_groups + 2.0 * sweeps_per_step)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    smem_peak = 128.0 * sm_count * sm_mhz * 1e6 / 1e9
    smem_achieved = smem_bytes * args.steps / sec / 1e9
    roofline["smem"] = {
        "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s", "frac": smem_achieved / smem_peak,
        "half_visit_x_masks": plan.half_groups, "full_visit_x_masks": plan.full_groups,
        "note": "algorithmic shared-memory bytes (partner rows + tile staging) / time; peak = 128 B/clk/SM x SMs x sampled SM clock",
    }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)

    cpu = None
    if cpu_child is not None:  # the reference arm's sample, run on the host cores beside the GPU warm-up (rank 0, N = 1 only)
        try:
            out_txt, _ = cpu_child.communicate(timeout=300)
            cpu = json.loads([l for l in out_txt.splitlines() if l.startswith("CPU_SAMPLE ")][-1][len("CPU_SAMPLE "):])
        except Exception as exc:  # reported, never fatal for the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"CPU sample failed: {exc!r}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": bench_config(world, n, T, plan), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "energy": energy.real, "energy_imag": energy.imag, "term_expansion_s": t_terms,
        "hbm_gbs_algorithmic": roofline["achieved"] * world,
    }
    if e2e_skip is not None:
        line["e2e_skipped"] = e2e_skip
    if se is not None:
        line["exchanges_per_step"] = se.exchanges // max(args.warmup + args.steps, 1)
    print(json.dumps(line), flush=True)
    stamp("line printed")
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)  # skip the teardown of tens of GB of pinned and device memory: the driver's clock is running


if __name__ == "__main__":
    main()
