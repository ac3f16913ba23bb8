import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from siga_b200 import api, synth
cfg = sys.argv[1]; m = int(sys.argv[2]) if len(sys.argv) > 2 else 45; scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
t0 = time.time(); reads = synth.config_reads(cfg, scale); n, L = reads.shape
flat, off = synth.reads_to_flat(reads); rank = synth.default_name_ranks(n)
print("generated", reads.shape, "in %.1f s" % (time.time() - t0), flush=True)
ctx = api.Context(0)
t0 = time.perf_counter(); r = ctx.pack_reads(flat, off); print("pack %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
res = {}
for rv in (False, True):
    for it in range(2):
        t0 = time.perf_counter(); o = ctx.index_build(r, reversed=rv); dt = time.perf_counter() - t0
    res[rv] = o
    print("index rev=%d wall %.1f ms" % (rv, dt * 1e3), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in o.stats.items()}, "runs", len(o.runs), flush=True)
    cnt = np.bincount(o.runs >> 5, weights=(o.runs & 31).astype(np.float64), minlength=5).astype(np.int64)
    assert cnt[0] == n and cnt.sum() == o.n_suffixes and np.array_equal(cnt[1:], np.bincount(flat, minlength=256)[[65, 67, 71, 84]]), cnt
    assert np.array_equal(np.sort(o.sai), np.arange(n, dtype=np.uint32))
print("Gbases/s (both strands, device): %.3f" % (n * L / ((res[False].stats["ms_total"] + res[True].stats["ms_total"]) * 1e-3) / 1e9))
fwd = ctx.fm_create(res[False].runs, res[False].n_suffixes); rev = ctx.fm_create(res[True].runs, res[True].n_suffixes)
ctx.fm_attach_reads(fwd, res[False].sai, r); ctx.fm_attach_reads(rev, res[True].sai, r)
for it in range(3):
    t0 = time.perf_counter(); e = ctx.overlap_edges(fwd, rev, r, 0, n, m, rank, copy=False); dt = time.perf_counter() - t0
    print("overlap m=%d wall %.1f ms -> %.2f M reads/s" % (m, dt * 1e3, n / dt / 1e6), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in e.stats.items()}, flush=True)
    ed = e.edges
    assert np.all(rank[ed["query"]] > rank[ed["target"]]) and np.all((ed["length"] >= m) & (ed["length"] <= L))
    ctx.free_edges(e)
ctx.close()
