#!/usr/bin/env python
"""One `ncu --set full` capture of a bench workload's dominant kernel -> its entry in profiles/traffic.json
(what bench.py's roofline block reads) and the summary profiles/<name>.md.

    python tools/ncu_traffic.py <report.ncu-rep> <bench-line.json> <key> <profiles/out.md> [note...]

<bench-line.json> is the line `python bench.py --workload W [--extra-flags F]` printed on the same box: its
roofline.live_samples_per_launch belongs to the very launch the capture profiled (same workload, one GPU).
<key> is the workload name, or "<name>+flags<F>" for a flagged variant.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    rep, bench, key, md = sys.argv[1:5]
    note = " ".join(sys.argv[5:])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    r = [x for x in data if x[col["smsp__inst_executed.sum"]] not in ("", "nan", "-nan")][0]

    def num(k):
        return float(r[col[k]].replace(",", ""))

    def scaled(k, units):
        """ncu prints byte counters in a per-metric unit; normalise to bytes"""
        u = rows[1][col[k]]
        return num(k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

    line = json.loads(Path(bench).read_text().strip().splitlines()[-1])
    live = int(line["roofline"]["live_samples_per_launch"])
    req = num("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
    entry = {
        "bytes": int(scaled("dram__bytes_read.sum", 1) + scaled("dram__bytes_write.sum", 1)),
        "kernel": r[col["Kernel Name"]],
        "profile": md,
        "n_gpus": 1,
        "frames": line["config"]["frames"],
        "warp_instructions": int(num("smsp__inst_executed.sum")),
        "live_samples": live,
        "ncu_kernel_ms": scaled_ms(rows, col, r),
        "issue_active_pct": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
        "l1tex_pct": round(num("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), 2),
        "xu_pct": round(num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), 2),
        "fp64_pct": round(num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 2),
        "l1_hit_pct": round(num("l1tex__t_sector_hit_rate.pct"), 2),
        "l2_hit_pct": round(num("lts__t_sector_hit_rate.pct"), 2),
    }
    # L1 data-pipe wavefronts (the unit's peak is one per clock per SM): ncu reports the utilisation, the count is
    # utilisation x elapsed L1 cycles summed over the SMs
    wf = num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0 * num("l1tex__cycles_elapsed.sum")
    entry["l1_wavefronts"] = int(wf)
    entry["l1_wavefront_pct"] = round(num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), 2)
    if req > 0:
        entry["l1_wavefronts_per_global_load"] = round((wf - num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")) / req, 2)
    tj = ROOT / "profiles" / "traffic.json"
    d = json.loads(tj.read_text())
    d[key] = entry
    tj.write_text(json.dumps(d, indent=1) + "\n")
    subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), rep, md] + ([note] if note else []), check=True)
    print(json.dumps({key: entry}, indent=1))


def scaled_ms(rows, col, r):
    u = rows[1][col["gpu__time_duration.sum"]]
    v = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    return round(v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}[u], 4)


if __name__ == "__main__":
    main()
