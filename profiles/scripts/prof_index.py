import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from siga_b200 import api, synth
reads = synth.config_reads("cfg2", float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
flat, off = synth.reads_to_flat(reads)
ctx = api.Context(0)
for it in range(3):
    t0 = time.perf_counter(); r = ctx.pack_reads(flat, off); t1 = time.perf_counter()
    print("pack %.1f ms" % ((t1 - t0) * 1e3))
    for rv in (False, True):
        t0 = time.perf_counter(); o = ctx.index_build(r, reversed=rv); t1 = time.perf_counter()
        print("build rev=%d wall %.1f ms" % (rv, (t1 - t0) * 1e3), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in o.stats.items()})
    t0 = time.perf_counter(); ctx.free_reads(r); print("free %.1f ms" % ((time.perf_counter() - t0) * 1e3))
ctx.close()
