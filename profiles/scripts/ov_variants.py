"""Overlap-stage experiments on one GPU: the cfg2 read set (or a scaled one), sgx_overlap_edges under a list of
environment-variable variants (the hooks of overlap.cu are read at call time).

  python profiles/scripts/ov_variants.py [scale] [reps] [m] ["VAR=val,VAR2=val;VAR=val;..."]

One JSON line per variant: best wall / device / search / post times, executed lookups, fetched lines, edge digest
(tests/fullsize_common.py) so that variants can be compared with each other and with tests/golden/fullsize."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import fullsize_common as fc  # noqa: E402
from siga_b200 import api, synth  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    variants = sys.argv[4].split(";") if len(sys.argv) > 4 else [""]
    reads = synth.config_reads("cfg2", scale)
    n, L = reads.shape
    flat, off = synth.reads_to_flat(reads)
    ctx = api.Context(0)
    r = ctx.pack_reads(flat, off)
    fi, ri = ctx.index_build(r), ctx.index_build(r, reversed=True)
    fwd, rev = ctx.fm_create(fi.runs, fi.n_suffixes), ctx.fm_create(ri.runs, ri.n_suffixes)
    ctx.fm_attach_reads(fwd, fi.sai, r)
    ctx.fm_attach_reads(rev, ri.sai, r)
    nr = synth.default_name_ranks(n)
    golden = None
    gp = os.path.join(REPO, "tests", "golden", "fullsize", "cfg2_%g.json" % scale)
    if os.path.isfile(gp):
        golden = json.load(open(gp)).get("overlap", {}).get(str(m))
    for v in variants:
        sets = dict(kv.split("=") for kv in v.split(",") if kv)
        for k, val in sets.items():
            os.environ[k] = val
        e = ctx.overlap_edges(fwd, rev, r, 0, n, m, nr)  # warm-up + digest
        digest = fc.edge_digest(e.edges, np.int64(L))
        n_ss = int(e.substring.sum())
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            x = ctx.overlap_edges(fwd, rev, r, 0, n, m, nr, copy=False)
            wall = (time.perf_counter() - t0) * 1e3
            st = dict(x.stats)
            ctx.free_edges(x)
            if best is None or wall < best[0]:
                best = (wall, st)
        wall, st = best
        out = dict(variant=v, m=m, reads=n, wall_ms=round(wall, 2), device_ms=round(st["ms_total"], 2),
                   search_ms=round(st["ms_search"], 2), post_ms=round(st["ms_extract"], 2), edges=st["n_edges"],
                   lookups_per_read=round(st["lookups_search"] / n, 1), lines_per_read=round(st["lines_search"] / n, 1),
                   slow_reads=st["slow_reads"], substring=n_ss, digest=digest[:16])
        if golden:
            out["matches_reference"] = bool(digest == golden["edges_digest"] and st["n_edges"] == golden["n_ed"] and
                                            n_ss == golden["n_substring"])
        print(json.dumps(out), flush=True)
        for k in sets:
            del os.environ[k]
    ctx.close()


if __name__ == "__main__":
    main()
