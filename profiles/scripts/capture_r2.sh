#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE GPU, each only after the same command has exited 0 without ncu).
#   bash profiles/scripts/capture_r2.sh <tag>
# gpurun_out/<tag>_launches.csv : per-launch gpu__time_duration of one overlap pass + the index builds (cold-cache,
#                                 serialised: shares of the step, not absolutes)
# gpurun_out/<tag>_{search,reduce,merge,scatter}.ncu-rep : --set full of the dominant kernels at full cfg2 size
set -u
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
python profiles/scripts/prof_overlap.py 1.0 2 > $out/${tag}_prof_overlap.plain.log 2>&1 || exit 1
python profiles/scripts/prof_index.py 1.0 1 > $out/${tag}_prof_index.plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ov_search_pq2 -c 1 -o $out/${tag}_search -f \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ov_reduce_lean -c 1 -o $out/${tag}_reduce -f \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_reduce.log 2>&1
# index kernels at FULL scale, one launch from the middle of the column passes (pass ~75 of 150 of the first build)
ncu --set full --clock-control none --import-source on -k regex:bcr_merge --launch-skip 75 -c 1 -o $out/${tag}_merge -f \
    python profiles/scripts/prof_index.py 1.0 1 > $out/${tag}_ncu_merge.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:part_scatter_kernel.*BcrMove --launch-skip 75 -c 1 -o $out/${tag}_scatter -f \
    python profiles/scripts/prof_index.py 1.0 1 > $out/${tag}_ncu_scatter.log 2>&1
for k in search reduce merge scatter; do
  ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_prof_$k.raw.csv 2>/dev/null
done
gzip -f $out/${tag}_launches.csv
