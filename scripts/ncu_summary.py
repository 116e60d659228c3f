#!/usr/bin/env python
"""Summarise an Nsight Compute report (and optionally a launch list) into profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_fgs_apply.ncu-rep profiles/r01_v1_fgs_apply.md \
        [--launches gpurun_out/launches.csv] [--note "text"]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_"


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    launches = sys.argv[sys.argv.index("--launches") + 1] if "--launches" in sys.argv else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary: {rep}", "", note, "",
             "Captured with `ncu --set full --clock-control none --import-source on` (per-launch, cold cache, serialised).", ""]
    lines.append("| metric | " + " | ".join(f"launch {r[col['ID']]}" for r in data) + " |")
    lines.append("|---|" + "---|" * len(data))
    lines.append("| kernel | " + " | ".join(r[col["Kernel Name"]].split("(")[0] for r in data) + " |")
    for k in KEYS:
        if k in col:
            lines.append(f"| {k} [{units[col[k]]}] | " + " | ".join(r[col[k]] for r in data) + " |")
    lines += ["", "## warp stall reasons (warps stalled per issue-active cycle, launch 0; >= 0.05 only)", ""]
    for h in hdr:
        if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and data:
            v = data[0][col[h]]
            try:
                if float(v) >= 0.05:
                    lines.append(f"- {h[len(STALLS):-len('_per_issue_active.ratio')]}: {v}")
            except ValueError:
                pass
    if launches:
        lines += ["", f"## launch list ({launches}; `--metrics gpu__time_duration.sum --clock-control none`)", "",
                  "| id | kernel | grid | block | ns |", "|---|---|---|---|---|"]
        tot = {}
        with open(launches) as f:
            rd = csv.reader(l for l in f if l.startswith('"'))
            h = next(rd)
            c = {n: i for i, n in enumerate(h)}
            for r in rd:
                if len(r) <= c["Metric Value"]:
                    continue
                k = r[c["Kernel Name"]].split("(")[0]
                ns = float(r[c["Metric Value"]].replace(",", ""))
                tot.setdefault(k, [0, 0.0])
                tot[k][0] += 1; tot[k][1] += ns
                if "vfgs::" in k:
                    lines.append(f"| {r[c['ID']]} | {k} | {r[c['Grid Size']]} | {r[c['Block Size']]} | {ns:.0f} |")
        lines += ["", "| kernel | launches | total ns | share of all profiled launches |", "|---|---|---|---|"]
        alln = sum(v[1] for v in tot.values())
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"| {k} | {v[0]} | {v[1]:.0f} | {100 * v[1] / alln:.1f}% |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
