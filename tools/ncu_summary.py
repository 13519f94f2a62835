"""Summarise an .ncu-rep (raw page) into the handful of metrics we track.
usage: ncu_summary.py file.ncu-rep [--traffic n_local dtype]
--traffic appends the capture's per-launch DRAM bytes (mean over the captured launches of the sweep kernel) to profiles/ncu_traffic.json,
which bench.py reads for roofline.traffic."""
import csv, json, subprocess, sys
from pathlib import Path
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed.sum', 'sm__cycles_elapsed.avg', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_no_instructions']
for r in rows[2:]:
    print('----')
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k}: {r[i]} {units[i]}")

if "--traffic" in sys.argv:
    a = sys.argv.index("--traffic")
    n_local, dtype = int(sys.argv[a + 1]), sys.argv[a + 2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    def col(name):
        i = hdr.index(name)
        return [float(r[i].replace(",", "")) * scale[units[i]] for r in rows[2:]]
    rd, wr, names = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), [r[hdr.index('Kernel Name')] for r in rows[2:]]
    keep = [i for i in range(len(rd)) if rd[i] == rd[i] and wr[i] == wr[i]]  # ncu sometimes reports -nan for a replayed launch
    rd, wr = [rd[i] for i in keep], [wr[i] for i in keep]
    out = Path(__file__).resolve().parents[1] / "profiles" / "ncu_traffic.json"
    doc = json.loads(out.read_text()) if out.exists() else {"captures": []}
    doc["captures"] = [c for c in doc["captures"] if not (c["n_local"] == n_local and c["dtype"] == dtype)]
    doc["captures"].append({"n_local": n_local, "dtype": dtype, "kernel": names[0].strip(), "launches": len(rd), "dram_bytes_read": sum(rd) / len(rd),
                            "dram_bytes_write": sum(wr) / len(wr), "source": Path(sys.argv[1]).name})
    out.write_text(json.dumps(doc, indent=1) + "\n")
    print("wrote", out)
