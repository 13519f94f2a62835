"""Timing of the batched overlap kernel (qfe_vec_gram): P vectors of n qubits, Gram matrix in one pass per 32 x 32 panel, against the
time P vector reads take at the measured HBM peak.  usage: python tools/bench_gram.py [n] [P]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
P = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eng = qfe.get_engine(0)
eng.use_torch_stream()
vecs = [eng.random_state(n, seed=s) for s in range(P)]
g = eng.gram(vecs)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
g = eng.gram(vecs)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
ev0.record()
d = eng.dot(vecs[0], vecs[1])
ev1.record()
torch.cuda.synchronize()
ms_dot = ev0.elapsed_time(ev1)
peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists() else 6544.3
floor_ms = P * (16 << n) / peak / 1e6
print(json.dumps({"n": n, "P": P, "gram_ms": ms, "floor_ms_P_vector_reads_at_hbm_peak": floor_ms, "ratio_to_floor": ms / floor_ms, "one_dot_ms": ms_dot,
                  "entry_by_entry_ms_estimate": ms_dot * P * (P + 1) / 2, "hermitian": bool((g == g.conj().T).all()), "g01_matches_dot": bool(abs(g[0, 1] - d) <= 1e-12 * abs(d) + 1e-15)}))
