#!/bin/bash
# Final round-2 ncu captures (ONE GPU; each command has exited 0 without ncu first).
#   bash profiles/scripts/capture_r2z.sh <tag>
set -u
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
python profiles/scripts/prof_overlap.py 1.0 2 > $out/${tag}_prof_overlap.plain.log 2>&1 || exit 1
python profiles/scripts/prof_index.py 1.0 1 > $out/${tag}_prof_index.plain.log 2>&1 || exit 1
SGX_SHARD_TRACE=1 python profiles/scripts/prof_sharded_local.py 1.0 4 > $out/${tag}_prof_sharded_local.plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ov_search_pq2 -c 1 -o $out/${tag}_search -f \
    python profiles/scripts/prof_overlap.py 1.0 1 > $out/${tag}_ncu_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bcr_merge --launch-skip 75 -c 1 -o $out/${tag}_merge -f \
    python profiles/scripts/prof_index.py 1.0 1 > $out/${tag}_ncu_merge.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bcr_scatter --launch-skip 75 -c 1 -o $out/${tag}_scatter -f \
    python profiles/scripts/prof_index.py 1.0 1 > $out/${tag}_ncu_scatter.log 2>&1
# exchange kernels of the sharded build, local transport: pass ~75 of shard 1 (launches come in shard order 0..3 per pass)
ncu --set full --clock-control none --import-source on -k regex:shard_send --launch-skip 301 -c 1 -o $out/${tag}_send -f \
    python profiles/scripts/prof_sharded_local.py 1.0 4 > $out/${tag}_ncu_send.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shard_place --launch-skip 301 -c 1 -o $out/${tag}_place -f \
    python profiles/scripts/prof_sharded_local.py 1.0 4 > $out/${tag}_ncu_place.log 2>&1
for k in search merge scatter send place; do
  ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_prof_$k.raw.csv 2>/dev/null
done
gzip -f $out/${tag}_launches.csv
rm -f $out/${tag}_*.ncu-rep.tmp
ls -la $out | grep ${tag}
