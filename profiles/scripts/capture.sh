#!/bin/bash
# ncu captures of the hot-path kernels (run under gpurun on ONE GPU, after the same commands exited 0 without ncu).
#   bash profiles/scripts/capture.sh <tag>
# Writes gpurun_out/<tag>_launches.csv (per-launch gpu__time_duration of one bench pass) and one --set full report per
# dominant kernel (search, reduce, merge).  The numbers printed by runs under ncu are never bench values.
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python profiles/scripts/prof_overlap.py 1.0 2 > $out/${tag}_prof_overlap.plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ov_search_quad -c 1 -o $out/${tag}_search -f \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ov_reduce -c 1 -o $out/${tag}_reduce -f \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_reduce.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bcr_merge --launch-skip 100 -c 1 -o $out/${tag}_merge -f \
    python profiles/scripts/prof_overlap.py 0.2 1 > $out/${tag}_ncu_merge.log 2>&1
# k-mer correction kernel (siga correct): plain run first, then one --set full capture of the single launch
python profiles/scripts/prof_correct.py 1.0 2 > $out/${tag}_correct.jsonl 2> $out/${tag}_correct.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:corr_kernel -c 1 -o $out/${tag}_correct -f \
    python profiles/scripts/prof_correct.py 1.0 1 > $out/${tag}_ncu_correct.log 2>&1
