"""Bucket-sharded index build over the LOCAL transport (4 virtual shards in one context, one GPU): the same kernels the
multi-GPU build runs -- shard_rank / partition count / shard_send (scatter fused with the push) / shard_wait /
shard_place / merge -- with the "peers" in the same HBM, for ncu captures of the exchange kernels (ncu must not wrap a
multi-rank command).

  python profiles/scripts/prof_sharded_local.py [scale] [shards]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
from siga_b200 import api, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reads = synth.config_reads("cfg2", scale)
flat, off = synth.reads_to_flat(reads)
ctx = api.Context(0)
r = ctx.pack_reads(flat, off)
comm = ctx.comm_init_local(shards)
one = ctx.index_build(r)
for it in range(2):
    t0 = time.perf_counter()
    o = ctx.index_build_sharded(comm, r)
    t1 = time.perf_counter()
    print("sharded x%d wall %.1f ms" % (shards, (t1 - t0) * 1e3), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in o.stats.items()},
          "identical:", bool(np.array_equal(o.runs, one.runs) and np.array_equal(o.sai, one.sai)))
ctx.free_comm(comm)
ctx.free_reads(r)
ctx.close()
