#!/bin/bash
# Evidence runs of round 2 (one GPU): every workload first runs WITHOUT ncu (must exit 0), then under
# `ncu --set full --clock-control none --import-source on` for its kernels; part "a" also takes the launch list of the
# bench command.  Two parts because gpurun brings back at most 64 MiB per call.  Outputs under gpurun_out/ (scratch);
# tools/ncu_summary.py turns them into profiles/r02_*.
#   tools/capture_r02.sh a    rays + Cornell steady state + launch list
#   tools/capture_r02.sh b    glass scene + BDPT (walk and connect kernels)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
run() { echo "== $*"; "$@"; echo "== rc=$?"; }
part=${1:-a}
{
if [ "$part" = a ]; then
run python tools/prof_run.py rays
run $NCU -k regex:'k_trace_closest|k_trace_any' -o $O/r02_rays python tools/prof_run.py rays
TUTU_PROF_SPP=256 TUTU_LANES=1 run python tools/prof_run.py render
TUTU_PROF_SPP=256 TUTU_LANES=1 run $NCU -k regex:'wf_shade|wf_extend_small|wf_shadow_small' --launch-skip 30 -c 3 -o $O/r02_steady python tools/prof_run.py render
run python bench.py --steps 2 --warmup 3 --spp 64 --no-cpu --no-extras --no-rays
run ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --spp 64 --no-cpu --no-extras --no-rays
else
run python tools/prof_glass.py
run $NCU -k regex:'wf_extend|wf_shade|wf_shadow|wf_classify' --launch-skip 16 -c 4 -o $O/r02_glass python tools/prof_glass.py
run python tools/prof_run.py bdpt
run $NCU -k regex:'q_extend|bdpt_vertex' --launch-skip 4 -c 4 -o $O/r02_bdpt python tools/prof_run.py bdpt
run $NCU -k regex:'bdpt_connect|q_shadow_add' --launch-skip 4 -c 4 -o $O/r02_bdpt_connect python tools/prof_run.py bdpt
fi
} > $O/r02_capture_$part.log 2>&1
ls -la $O/r02_*.ncu-rep $O/r02_launches_bench.csv 2>/dev/null
grep -E "^== rc=|==ERROR==" $O/r02_capture_$part.log | sort | uniq -c
