import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from siga_b200 import api, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reads = synth.config_reads("cfg2", scale)
flat, off = synth.reads_to_flat(reads)
ctx = api.Context(0)
r = ctx.pack_reads(flat, off)
fi = ctx.index_build(r); ri = ctx.index_build(r, reversed=True)
print("index", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in ri.stats.items()})
fwd = ctx.fm_create(fi.runs, fi.n_suffixes); rev = ctx.fm_create(ri.runs, ri.n_suffixes)
ctx.fm_attach_reads(fwd, fi.sai, r); ctx.fm_attach_reads(rev, ri.sai, r)
nr = api.name_ranks(["r%d" % i for i in range(len(reads))])
for it in range(reps):
    t0 = time.perf_counter(); h = ctx.overlap_edges(fwd, rev, r, 0, len(reads), 45, nr, copy=False); t1 = time.perf_counter()
    print("overlap wall %.1f ms" % ((t1 - t0) * 1e3), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in h.stats.items()})
    ctx.free_edges(h)
ctx.close()
