"""Turns ncu artefacts brought back in gpurun_out/ into the small text/JSON summaries that are
committed under profiles/ (run here, on the CPU box: `ncu -i` needs no GPU).

  python tools/ncu_summary.py full   <report.ncu-rep> <profiles/out_prefix>   # --set full capture
  python tools/ncu_summary.py launch <launches.csv>   <profiles/out.txt>      # gpu__time_duration list
  python tools/ncu_summary.py stalls <report.ncu-rep> <kernel-regex> <profiles/out.txt>
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy % (warps active)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct", "issue slots busy % (issue-slot utilisation)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "inst issued % of peak"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (warp execution efficiency, of 32)"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", "warp execution efficiency %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long scoreboard %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard (warps per issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard (warps per issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (warps per issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle (warps per issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall LG throttle (warps per issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps per issue)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch resolving (warps per issue)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected (warps per issue)"),
    ("local_load_bytes", "local-memory load bytes (spills/stack)"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (warp inst)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (warp inst)"),
]


def ncu_csv(rep: str, page: str) -> list[list[str]]:
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def short(name: str) -> str:
    return re.sub(r"\(.*", "", name.replace("void ", "").replace("tutu::", ""))


def to_bytes(val: str, unit: str) -> float:
    v = float(val.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * mult.get(unit, 1)


def to_us(val: str, unit: str) -> float:
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}.get(unit, 1)


def full(rep: str, prefix: str) -> None:
    rows = ncu_csv(rep, "raw")
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none, report {rep.split('/')[-1]} (read with ncu -i --page raw --csv)",
             "# one block per profiled launch; values are per launch"]
    traffic: dict[str, dict] = defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "l1tex_pct": 0.0, "dram_pct": 0.0,
                                                    "lanes": 0.0, "l1_hit_pct": 0.0, "l2_hit_pct": 0.0, "registers": 0})
    def num(r, key):
        try:
            return float(r[idx[key]].replace(",", ""))
        except (KeyError, ValueError):
            return 0.0
    for r in body:
        name = short(r[idx["Kernel Name"]])
        lines.append("")
        lines.append(f"== {name}   grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        seen = set()
        for m, label in METRICS:
            if m in idx and label not in seen and r[idx[m]] != "":
                seen.add(label)
                lines.append(f"  {label:<72s} {r[idx[m]]} {units[idx[m]]}   [{m}]")
        try:
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            us = to_us(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
            lines.append(f"  {'DRAM traffic (read+write) / duration':<72s} {(rd + wr) / 1e6:.1f} MB / {us:.1f} us = {(rd + wr) / us * 1e-3:.0f} GB/s")
            t = traffic[name]
            t["launches"] += 1
            t["dram_bytes"] += rd + wr
            t["us"] += us
            t["l1tex_pct"] += num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
            t["dram_pct"] += num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or num(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")
            t["lanes"] += num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
            t["l1_hit_pct"] += num(r, "l1tex__t_sector_hit_rate.pct")
            t["l2_hit_pct"] += num(r, "lts__t_sector_hit_rate.pct")
            t["registers"] = int(num(r, "launch__registers_per_thread"))
        except (KeyError, ValueError):
            pass
    open(prefix + ".txt", "w").write("\n".join(lines) + "\n")
    js = {k: {"launches_profiled": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
              "us_per_launch_under_ncu": v["us"] / v["launches"],
              "dram_gbs_under_ncu": v["dram_bytes"] / v["us"] * 1e-3,
              "dram_pct_of_peak": v["dram_pct"] / v["launches"], "l1tex_pct_of_peak": v["l1tex_pct"] / v["launches"],
              "l1_hit_pct": v["l1_hit_pct"] / v["launches"], "l2_hit_pct": v["l2_hit_pct"] / v["launches"],
              "active_threads_per_warp_instruction": v["lanes"] / v["launches"], "registers": v["registers"]}
          for k, v in traffic.items()}
    json.dump(js, open(prefix + "_traffic.json", "w"), indent=1)
    print(f"wrote {prefix}.txt and {prefix}_traffic.json ({len(body)} launches)")


def launch(csv_path: str, out: str) -> None:
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg: dict[str, list[float]] = defaultdict(list)
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        agg[short(r[idx["Kernel Name"]])].append(to_us(r[idx["Metric Value"]], r[idx["Metric Unit"]]))
    total = sum(sum(v) for v in agg.values())
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none, source {csv_path.split('/')[-1]}",
             "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
             f"# {sum(len(v) for v in agg.values())} launches, {total / 1e3:.2f} ms summed",
             f"{'kernel':<34s} {'launches':>8s} {'sum ms':>10s} {'share':>7s} {'avg us':>9s} {'min us':>9s} {'max us':>9s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{k:<34s} {len(v):>8d} {sum(v) / 1e3:>10.3f} {sum(v) / total:>7.1%} {sum(v) / len(v):>9.1f} {min(v):>9.1f} {max(v):>9.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def stalls(rep: str, kernel_re: str, out: str) -> None:
    res = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel_re}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(res)))
    # first table only (first matching launch)
    hdr = None
    table = []
    for r in rows:
        if hdr is None:
            if "Source" in r and any("Sampling" in c for c in r):
                hdr = r
            continue
        if len(r) != len(hdr) or r == hdr:
            if r == hdr:
                break
            continue
        table.append(r)
    if hdr is None:
        raise SystemExit("no source table found")
    idx = {h: i for i, h in enumerate(hdr)}
    samp = next(h for h in hdr if h.startswith("# Samples") or "Warp Stall Sampling (All" in h)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[idx[samp]] or 0) for r in table) or 1.0
    by_reason = {h: sum(float(r[idx[h]] or 0) for r in table) for h in stall_cols}
    lines = [f"# ncu source page (SASS), kernel regex '{kernel_re}', report {rep.split('/')[-1]}; {int(tot)} warp-stall samples",
             "# stall reasons, share of all samples:"]
    rs = sum(by_reason.values()) or 1.0
    for h, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:10]:
        lines.append(f"  {h:<28s} {v / rs:6.1%}")
    lines.append("# top 25 instructions by samples:")
    for r in sorted(table, key=lambda r: -float(r[idx[samp]] or 0))[:25]:
        top = max(stall_cols, key=lambda h: float(r[idx[h]] or 0)) if stall_cols else ""
        lines.append(f"  {float(r[idx[samp]] or 0) / tot:6.2%}  {r[idx['Source']][:90]:<90s} {top}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:16]))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "full":
        full(sys.argv[2], sys.argv[3])
    elif mode == "launch":
        launch(sys.argv[2], sys.argv[3])
    elif mode == "stalls":
        stalls(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        raise SystemExit(__doc__)
