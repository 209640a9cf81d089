"""profiles/r02_sass_features.txt: the sm_100a instruction mnemonics the hot kernels of the shipped library rely on
(cuobjdump -sass; needs no GPU).  UTMALDG = cp.async.bulk.tensor TMA tile load, SYNCS = mbarrier, FENCE.VIEW.ASYNC =
proxy fence, FADD2 / FMUL2 = packed fp32, LDG.E...256 = 32-byte loads, FMNMX3 = 3-input min/max."""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "tuturenderer_b200" / "libtutu_b200.so"
KERNELS = ["wf_shadeE", "wf_extend_small", "wf_shadow_small", "k_trace_closestILi0", "k_trace_anyILi0", "k_trace_anyILi4", "wf_extendILi0", "wf_extendILi2",
           "wf_shadowILi0", "wf_shadowILi2", "q_extendILi0", "q_extendILi2", "q_shadow_addILi0", "q_shadow_addILi2", "bdpt_connect", "bdpt_vertex",
           "pt_resident", "sah_small", "sah_bin"]
WANT = re.compile(r"^(UTMALDG|UBLKCP|SYNCS|FENCE\.VIEW\.ASYNC|ELECT|FADD2|FMUL2|FFMA2|FMNMX3|LDG\.E\.[A-Z0-9.]*256|ATOMG|REDG|MUFU\.(RCP|RSQ)|VOTE|MATCH|SHFL|LDL|STL)")
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
counts, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((k for k in KERNELS if k in m.group(1)), None)
        if cur:
            counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and WANT.match(m.group(1)):
            op = m.group(1)
            op = "LDL" if op.startswith("LDL") else "STL" if op.startswith("STL") else "SHFL" if op.startswith("SHFL") else "VOTE" if op.startswith("VOTE") else op
            counts[cur][op] += 1
out = ["# cuobjdump -sass of libtutu_b200.so (tools/sass_features.py), final code of round 2: instruction mnemonics that show the sm_100a features the hot kernels use",
       "# (UTMALDG = cp.async.bulk.tensor TMA tile load, SYNCS = mbarrier, FENCE.VIEW.ASYNC = proxy fence, FADD2/FMUL2 = packed fp32, LDG.E...256 = 32-byte loads,",
       "#  FMNMX3 = 3-input min/max; LDL/STL = local memory: the traversal stack of the <2> / <4> tree kernels, DESIGN.md 5.10; VOTE = the ballots of the warp-level walks)"]
for k in KERNELS:
    if k in counts:
        out.append(f"{k}: " + ", ".join(f"{n} x {op}" for op, n in sorted(counts[k].items())))
(ROOT / "profiles" / "r02_sass_features.txt").write_text("\n".join(out) + "\n")
print("\n".join(out))
