"""Per-source-line instruction and stall shares of one kernel: joins the SASS table of an ncu report (--page source)
with the line table of the built library (cuobjdump -xelf + nvdisasm -g).  The library must be the build that was
profiled.   python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <mangled-name-substring> [top=40]"""
import csv, io, re, subprocess, sys, tempfile, collections, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lib = os.environ.get("TUTU_LIB", str(ROOT / "tuturenderer_b200" / "libtutu_b200.so"))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; idx = {c: i for i, c in enumerate(h)}
end = [i for i, r in enumerate(rows) if r == h]
tab = [r for r in rows[2:(end[1] if len(end) > 1 else len(rows))] if len(r) == len(h)]
num = lambda r, c: float((r[idx[c]] or "0").replace(",", ""))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    lines = None
    for cub in sorted(Path(td).glob("*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
        m = re.search(r"^\.text\.[^\n]*" + re.escape(mangled) + r"[^\n]*:\n", txt, re.M)
        if not m:
            continue
        body = txt[m.end():]
        nxt = re.search(r"^\.text\.|^\t\.section", body, re.M)
        if nxt:
            body = body[:nxt.start()]
        lines, cur = [], ("?", 0)
        for ln in body.splitlines():
            f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if f:
                cur = (Path(f.group(1)).name, int(f.group(2)))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur)
        break
if lines is None:
    sys.exit("kernel not found in " + lib)
if len(lines) != len(tab):
    print(f"# warning: {len(lines)} SASS instructions in the library, {len(tab)} in the report (different build?)")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for (f, l), r in zip(lines, tab):
    a = agg[(f, l)]
    a[0] += num(r, "Instructions Executed"); a[1] += num(r, "Thread Instructions Executed"); a[2] += num(r, "# Samples"); a[3] += 1
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"# {kre}: {ti:.4g} warp instructions, {ts:.0f} stall samples, {len(tab)} SASS instructions; per source line (innermost inlined location)")
print("# file:line  instr%  samples%  lanes  sass")
src_cache = {}
def src(f, l):
    for d in (ROOT / "tuturenderer_b200" / "csrc",):
        p = d / f
        if p.exists():
            if p not in src_cache: src_cache[p] = p.read_text().splitlines()
            s = src_cache[p]
            return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""
    return ""
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f}:{l:<5d} {a[0] / ti * 100:5.2f} {a[2] / ts * 100:6.2f}  {a[1] / max(a[0], 1):5.1f} {a[3]:4d}  {src(f, l)}")
# per file
byf = collections.defaultdict(lambda: [0.0, 0.0])
for (f, l), a in agg.items():
    byf[f][0] += a[0]; byf[f][1] += a[2]
print("# per file: instr% samples%")
for f, a in sorted(byf.items(), key=lambda kv: -kv[1][1]):
    print(f"{f:24s} {a[0] / ti * 100:6.2f} {a[1] / ts * 100:6.2f}")
