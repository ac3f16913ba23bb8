"""Where do the slow e2e steps come from?  The e2e step of bench.py (pinned host reads -> sgx_reads_pack_async ->
attach -> sgx_overlap_edges -> edges in pinned host memory), 30 times each: without any clock sampler, with the
`nvidia-smi -lms` sampler bench.py used, and with an in-process NVML sampler (pynvml, two queries per sample).

  python profiles/scripts/e2e_jitter.py [steps]
"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from siga_b200 import api, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reads = synth.config_reads("cfg2", 1.0)
n, L = reads.shape
host = torch.empty(reads.size, dtype=torch.uint8).pin_memory()
host.numpy()[:] = reads.reshape(-1)
flat = host.numpy()


def pinned(a):
    t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype))).pin_memory()
    t.numpy()[...] = a
    return t


ctx = api.Context(0)
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
r = ctx.pack_reads(flat, off)
fi, ri = ctx.index_build_pair(r)
fwd, rev = ctx.fm_create(fi.runs, fi.n_suffixes), ctx.fm_create(ri.runs, ri.n_suffixes)
keep = [pinned(off.view(np.int64)), pinned(fi.sai.view(np.int32)), pinned(ri.sai.view(np.int32)),
        pinned(synth.default_name_ranks(n).view(np.int32))]
off_p = keep[0].numpy().view(np.uint64)
sai_p, rsai_p, rank_p = (t.numpy().view(np.uint32) for t in keep[1:])


PARTS = []


def step():
    t = [time.perf_counter()]
    rr = ctx.pack_reads_async(flat, off_p)
    t.append(time.perf_counter())
    ctx.fm_attach_reads(fwd, sai_p, rr)
    t.append(time.perf_counter())
    ctx.fm_attach_reads(rev, rsai_p, rr)
    t.append(time.perf_counter())
    e = ctx.overlap_edges(fwd, rev, rr, 0, n, 45, rank_p, copy=False)
    t.append(time.perf_counter())
    dev = e.stats["ms_total"]
    ctx.free_edges(e)
    t.append(time.perf_counter())
    ctx.free_reads(rr)
    t.append(time.perf_counter())
    PARTS.append([round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])])  # pack, attach, attach, overlap, free edges, free reads
    return (t[-1] - t[0]) * 1e3, dev


def run(label):
    for _ in range(3):
        step()
    del PARTS[:]
    t = [step() for _ in range(steps)]
    slow_parts = [p for p, x in zip(PARTS, t) if x[0] > 90]
    typical = PARTS[len(PARTS) // 2]
    w = sorted(x[0] for x in t)
    print(json.dumps(dict(sampler=label, steps=steps, median_ms=round(w[len(w) // 2], 2), mean_ms=round(sum(w) / len(w), 2),
                          max_ms=round(w[-1], 2), over_90ms=sum(1 for x in w if x > 90),
                          slow=[(round(a, 1), round(b, 1)) for a, b in t if a > 90], typical_parts=typical, slow_parts=slow_parts)), flush=True)


run("none")
run("none again")
ctx.close()
