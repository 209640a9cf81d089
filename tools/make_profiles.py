"""profiles/r02_* from the reports that tools/capture_r02.sh (parts a and b) left in gpurun_out/: the --set full
summaries, lane tables, stall and per-source-line tables, the launch list, and the merged r02_traffic.json that
bench.py reads.  Runs here (ncu -i needs no GPU); the library on disk must be the build that was profiled."""
import json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
G, P, T = ROOT / "gpurun_out", ROOT / "profiles", ROOT / "tools"

def run(*a, out=None):
    r = subprocess.run([sys.executable, *map(str, a)], capture_output=True, text=True, cwd=ROOT)
    if r.returncode != 0:
        sys.exit(f"{a}: {r.stderr[-2000:]}")
    if out:
        Path(out).write_text(r.stdout)
    return r.stdout

captures = {
    "rays": "configs[1]: 4 Mi rays of each kind vs 999 698 triangles (tools/prof_run.py rays)",
    "steady": "Cornell 1024^2 wavefront in steady state, iteration 10 of a 256-spp render, one lane of 32 Mi paths (tools/prof_run.py render)",
    "glass": "glass / texture scene 1024^2 @ 8 spp, one lane (tools/prof_glass.py)",
    "bdpt": "Veach room 800x600 @ 4 spp BDPT (tools/prof_run.py bdpt)",
    "bdpt_connect": "Veach room 800x600 @ 4 spp BDPT (tools/prof_run.py bdpt)",
}
merged = {}
for name, workload in captures.items():
    run(T / "ncu_summary.py", "full", G / f"r02_{name}.ncu-rep", P / f"r02_{name}_full")
    for k, v in json.loads((P / f"r02_{name}_full_traffic.json").read_text()).items():
        k = k.split("::")[-1]  # kernels of an anonymous namespace come as "<unnamed>::name"
        key = f"{k}@glass" if name == "glass" and k in merged else k
        merged[key] = dict(v, capture=f"profiles/r02_{name}_full.txt", workload=workload)
        if key != k:
            merged[key]["kernel_profiled"] = k
(P / "r02_traffic.json").write_text(json.dumps(merged, indent=1) + "\n")
lanes = [("rays", "k_trace_closest", "closest"), ("rays", "k_trace_any", "any"), ("steady", "wf_shade", "steady_shade"),
         ("glass", "wf_extend", "glass_extend"), ("glass", "wf_shade", "glass_shade"), ("bdpt", "q_extend", "q_extend"),
         ("bdpt_connect", "bdpt_connect", "bdpt_connect")]
for rep, kre, tag in lanes:
    run(T / "ncu_lanes.py", G / f"r02_{rep}.ncu-rep", kre, out=P / f"r02_{tag}_lanes.txt")
run(T / "ncu_summary.py", "stalls", G / "r02_steady.ncu-rep", "wf_shade", P / "r02_steady_shade_stalls.txt")
for kre, mangled, tag in (("wf_shade", "wf_shadeENS", "shade"), ("wf_extend_small", "wf_extend_smallENS", "extend_small"),
                          ("wf_shadow_small", "wf_shadow_smallENS", "shadow_small")):
    run(T / "ncu_lines.py", G / "r02_steady.ncu-rep", kre, mangled, 60, out=P / f"r02_steady_{tag}_lines.txt")
run(T / "ncu_summary.py", "launch", G / "r02_launches_bench.csv", P / "r02_launches_bench.txt")
print("wrote", len(merged), "kernels to profiles/r02_traffic.json")
