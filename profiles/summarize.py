#!/usr/bin/env python
"""Turns the raw ncu artefacts brought back in gpurun_out/ into the small text summaries that
are committed under profiles/ (the .ncu-rep files themselves stay in gpurun_out/, untracked).

    python profiles/summarize.py launches gpurun_out/r1_launches_ivfflat_1m.csv  > profiles/r1_launches_ivfflat_1m.txt
    python profiles/summarize.py full     gpurun_out/r1_prof_ivfflat_scan.ncu-rep > profiles/r1_ncu_ivfflat_scan.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us":
                ns *= 1e3
            elif r.get("Metric Unit") == "ms":
                ns *= 1e6
            rows.append((r["Kernel Name"].split("(")[0][-90:], ns))
    agg = OrderedDict()
    for name, ns in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms total (ncu per-launch times: cold-cache, serialised -> compare SHARES)")
    print(f"{'kernel':92s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:92s} {cnt:8d} {ns / 1e6:10.3f} {ns / total:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name[:140]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:80s} {vals[i]:>18s} {units[i]}")


def traffic(out_path):
    """profiles/r1_traffic.json: DRAM bytes per launch of each dominant kernel, read back from the
    committed full-capture summaries (bench.py copies the figure into roofline.traffic)."""
    import json
    import os
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def dram(name):
        t = open(os.path.join(here, name)).read()
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(key + r"\s+([0-9.]+) (\w+)", t)
            tot += float(m.group(1)) * scale[m.group(2)]
        return tot

    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures "
                       "summarised in this directory (profiles/r1_capture.sh)",
           # keyed by bench workload, then by the kernel the bench timed: a capture only speaks for its own workload
           "ivfpq_10m": {"ivfpq_scan_kernel": dram("r1_ncu_ivfpq_scan.txt")},
           "ivfflat_1m": {"ivf_listmajor_tma_kernel": dram("r1_ncu_ivfflat_listmajor.txt"),
                          "ivfflat_scan_warp_kernel": dram("r1_ncu_ivfflat_scan_querymajor.txt")}}
    json.dump(out, open(out_path, "w"), indent=1)
    print(out)


def ncu_json(out_path):
    """profiles/r2_ncu.json: per bench workload, per dominant kernel, the counters of the committed ncu --set full
    summaries that bench.py copies into roofline.traffic / roofline.secondary.
        python profiles/summarize.py ncu_json profiles/r2_ncu.json"""
    import json
    import os
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def read(name, work_units=None):
        t = open(os.path.join(here, name)).read()

        def num(key):
            m = re.search(re.escape(key) + r"\s+([0-9.]+)\s*(\w*)", t)
            return (float(m.group(1)), m.group(2)) if m else (None, "")

        dram = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = num(key)
            dram += v * scale.get(u, 1.0)
        inst = num("smsp__inst_executed.sum")[0]
        wav, conf = num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")[0], num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")[0]
        dur, du = num("gpu__time_duration.sum")
        out = {"source": name, "dram_bytes": dram, "kernel_ms_under_ncu": dur * {"ms": 1.0, "us": 1e-3, "s": 1e3}.get(du, 1.0),
               "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active")[0],
               "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")[0],
               "tensor_pipe_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[0],
               "l2_hit_pct": num("lts__t_sector_hit_rate.pct")[0],
               "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active")[0],
               "shared_wavefronts": wav, "lds_bank_conflict_pct": 100.0 * conf / wav if wav else None,
               "warp_instructions": inst}
        if work_units:
            out["inst_per_entry"] = inst * 32.0 / work_units[0]
            out["inst_per_entry_unit"] = work_units[1]
        return out

    out = {"_comment": "counters per launch from the ncu --set full captures summarised in this directory; keyed by bench "
                       "workload, then by the kernel the bench timed (a capture only speaks for its own workload)",
           "ivfpq_10m": {"pqtc_scan_kernel": read("r2_ncu_pqtc_scan.txt", (3.939e9, "thread instructions per (query, entry) pair filtered")),
                         "ivfpq_scan_kernel": read("r1_ncu_ivfpq_scan.txt", (3.98e9, "thread instructions per scanned code"))},
           "ivfflat_1m": {"ivf_listmajor_tma_kernel": read("r1_ncu_ivfflat_listmajor.txt"),
                          "ivfflat_scan_warp_kernel": read("r1_ncu_ivfflat_scan_querymajor.txt")}}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "ncu_json": ncu_json}[sys.argv[1]](sys.argv[2])
