"""`siga correct` kernel at config-2 shape: python profiles/scripts/prof_correct.py [scale] [reps] [k]
Prints one JSON line per timed call (device time by CUDA events inside the library, no profiler)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from siga_b200 import api, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k = int(sys.argv[3]) if len(sys.argv) > 3 else 31
reads = synth.config_reads("cfg2", scale)
flat, off = synth.reads_to_flat(reads)
ctx = api.Context(0)
r = ctx.pack_reads(flat, off)
fi = ctx.index_build(r)
fm = ctx.fm_create(fi.runs, fi.n_suffixes)
gather = ctx.bench_gather(8 << 30, 1 << 26)
for it in range(reps):
    t0 = time.perf_counter()
    res = ctx.correct(fm, r, 0, len(reads), kmer_size=k)
    t1 = time.perf_counter()
    s = res.stats
    print(json.dumps(dict(workload="cfg2 x %.3g: %d reads x %d bp, siga correct -k %d" % (scale, len(reads), reads.shape[1], k),
                          wall_ms=round((t1 - t0) * 1e3, 1), ms_total=round(s["ms_total"], 2), ms_kernel=round(s["ms_kernel"], 2),
                          reads_per_s=round(len(reads) / (s["ms_kernel"] * 1e-3)), lookups=s["lookups"],
                          lookups_per_read=round(s["lookups"] / len(reads), 1),
                          lookups_per_s=round(s["lookups"] / (s["ms_kernel"] * 1e-3)),
                          algorithmic_GBps=round(s["lookups"] * 128 / (s["ms_kernel"] * 1e-3) / 1e9, 1),
                          random_gather_128B_per_s=round(gather), written=int(res.valid.sum()), n_fixes=s["n_fixes"],
                          fixed_reads=s["fixed_reads"])))
ctx.close()
