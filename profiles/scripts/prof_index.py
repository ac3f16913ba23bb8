"""Index-build timings on one GPU: sgx_index_build per strand and sgx_index_build_pair (both strands side by side).

  python profiles/scripts/prof_index.py [scale] [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
from siga_b200 import api, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reads = synth.config_reads("cfg2", scale)
flat, off = synth.reads_to_flat(reads)
ctx = api.Context(0)
fmt = lambda st: {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()}
for it in range(reps):
    t0 = time.perf_counter(); r = ctx.pack_reads(flat, off); t1 = time.perf_counter()
    print("pack %.1f ms" % ((t1 - t0) * 1e3))
    single = {}
    for rv in (False, True):
        t0 = time.perf_counter(); o = ctx.index_build(r, reversed=rv); t1 = time.perf_counter()
        single[rv] = o
        print("build rev=%d wall %.1f ms" % (rv, (t1 - t0) * 1e3), fmt(o.stats))
    t0 = time.perf_counter(); a, b = ctx.index_build_pair(r); t1 = time.perf_counter()
    same = all(np.array_equal(x.runs, y.runs) and np.array_equal(x.sai, y.sai) for x, y in ((a, single[False]), (b, single[True])))
    print("pair wall %.1f ms (identical to the single builds: %s)" % ((t1 - t0) * 1e3, same), fmt(a.stats), fmt(b.stats))
    ctx.free_reads(r)
ctx.close()
