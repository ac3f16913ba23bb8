"""Per-call wall times of the end-to-end overlap step of bench.py (pack + attach x2 + overlap_edges), cfg2."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from siga_b200 import api, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reads = synth.config_reads("cfg2", scale)
n, L = reads.shape
host = torch.empty(reads.size, dtype=torch.uint8).pin_memory()
host.numpy()[:] = reads.reshape(-1)
flat = host.numpy()
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
ctx = api.Context(0)
r = ctx.pack_reads(flat, off)
fi = ctx.index_build(r); ri = ctx.index_build(r, reversed=True)
fwd = ctx.fm_create(fi.runs, fi.n_suffixes); rev = ctx.fm_create(ri.runs, ri.n_suffixes)
ctx.fm_attach_reads(fwd, fi.sai, r); ctx.fm_attach_reads(rev, ri.sai, r)
nr = synth.default_name_ranks(n)
for it in range(5):
    t = [time.perf_counter()]
    rr = ctx.pack_reads(flat, off); t.append(time.perf_counter())
    ctx.fm_attach_reads(fwd, fi.sai, rr); t.append(time.perf_counter())
    ctx.fm_attach_reads(rev, ri.sai, rr); t.append(time.perf_counter())
    e = ctx.overlap_edges(fwd, rev, rr, 0, n, 45, nr, copy=False); t.append(time.perf_counter())
    ctx.free_edges(e); ctx.free_reads(rr); t.append(time.perf_counter())
    print("iter %d: pack %.1f attach %.1f %.1f overlap %.1f (device %.1f) free %.1f ms" % (
        (it,) + tuple((t[i + 1] - t[i]) * 1e3 for i in range(3)) + ((t[4] - t[3]) * 1e3, e.stats["ms_total"], (t[5] - t[4]) * 1e3)))
ctx.close()
