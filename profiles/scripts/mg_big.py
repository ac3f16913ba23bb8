"""Configs 4 / 5 on N GPUs of one box: bucket-sharded index build over NCCL + read-sharded overlap.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
      profiles/scripts/mg_big.py cfg4 [min_overlap] [scale] [single]

Every rank generates the same synthetic read set and packs it on its GPU (replicated reads).  Index: both strands
with sgx_index_build_sharded (all ranks together).  With `single` rank 0 also runs sgx_index_build alone and the two
results are compared byte for byte.  Overlap: the run bytes are broadcast, every rank builds both FM-indices and
searches its contiguous range of reads; edge records are gathered on rank 0 (the "final gather of ASQG records").
Times are device/wall maxima over ranks.  Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import shard
    from siga_b200 import api, synth

    cfg = sys.argv[1]
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    single = len(sys.argv) > 4 and sys.argv[4] == "single"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def mx(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync():
        dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    reads = synth.config_reads(cfg, scale)
    n, L = reads.shape
    flat, off = synth.reads_to_flat(reads)
    name_rank = synth.default_name_ranks(n)
    out = {"cfg": cfg, "scale": scale, "reads": int(n), "read_len": int(L), "world": world, "min_overlap": m,
           "generate_s": round(time.time() - t0, 1)}
    ctx = api.Context(local)
    uid = torch.from_numpy(ctx.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
    dist.broadcast(uid, 0)
    comm = ctx.comm_init(uid.cpu().numpy(), rank, world)
    r = ctx.pack_reads(flat, off)
    del reads

    # ---- index, both strands, sharded over all ranks
    idx = {}
    for rv in (False, True):
        for it in range(2):
            sync()
            t0 = time.perf_counter()
            o = ctx.index_build_sharded(comm, r, reversed=rv)
            torch.cuda.synchronize()
            wall = mx((time.perf_counter() - t0) * 1e3)
        st = {k: round(mx(float(o.stats[k])), 2) for k in ("ms_total", "ms_sort", "ms_bcr", "ms_emit", "ms_merge")}
        st["wall_ms"] = round(wall, 2)
        st["launches"] = o.stats["launches"]
        idx[rv] = o
        out["index_sharded_rev%d" % rv] = st
        if rank == 0:
            cnt = np.bincount(o.runs >> 5, weights=(o.runs & 31).astype(np.float64), minlength=5).astype(np.int64)
            assert cnt[0] == n and cnt.sum() == o.n_suffixes
            assert np.array_equal(cnt[1:], np.bincount(flat, minlength=256)[[65, 67, 71, 84]])
        assert np.array_equal(np.sort(o.sai), np.arange(n, dtype=np.uint32))
    both = out["index_sharded_rev0"]["wall_ms"] + out["index_sharded_rev1"]["wall_ms"]
    out["index_gbases_per_s"] = round(n * L / (both * 1e-3) / 1e9, 3)
    if single:
        same = True
        if rank == 0:
            for rv in (False, True):
                for it in range(2):
                    t0 = time.perf_counter()
                    one = ctx.index_build(r, reversed=rv)
                    w1 = (time.perf_counter() - t0) * 1e3
                out["index_single_rev%d" % rv] = {"wall_ms": round(w1, 2), "ms_total": round(one.stats["ms_total"], 2),
                                                  "ms_sort": round(one.stats["ms_sort"], 2), "ms_merge": round(one.stats["ms_merge"], 2)}
                same = same and np.array_equal(one.runs, idx[rv].runs) and np.array_equal(one.sai, idx[rv].sai)
                del one
            out["sharded_identical_to_single"] = bool(same)
        sync()

    # ---- overlap: run bytes to every rank, reads sharded by contiguous range
    fms = []
    for rv in (False, True):
        nr = torch.tensor([len(idx[rv].runs) if rank == 0 else 0], dtype=torch.int64, device="cuda")
        dist.broadcast(nr, 0)
        buf = torch.empty(int(nr.item()), dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.from_numpy(idx[rv].runs))
        dist.broadcast(buf, 0)
        runs = buf.cpu().numpy()
        f = ctx.fm_create(runs, idx[rv].n_suffixes)
        ctx.fm_attach_reads(f, idx[rv].sai, r)
        fms.append(f)
        del buf
    first, count = shard.read_range(n, rank, world)
    for it in range(3):
        sync()
        t0 = time.perf_counter()
        e = ctx.overlap_edges(fms[0], fms[1], r, first, count, m, name_rank, copy=False)
        torch.cuda.synchronize()
        wall = mx((time.perf_counter() - t0) * 1e3)
        st = dict(e.stats)
        ed = e.edges
        if it == 2:
            assert np.all(name_rank[ed["query"]] > name_rank[ed["target"]]) and np.all((ed["length"] >= m) & (ed["length"] <= L))
            # final gather of the edge records on rank 0 (sgx_edges_gather: pinned buffers + NCCL)
            sync()
            t0 = time.perf_counter()
            allb = ctx.edges_gather(comm, e, root=0, copy=False)
            torch.cuda.synchronize()
            out["gather_ms"] = round(mx((time.perf_counter() - t0) * 1e3), 2)
            if rank == 0:
                out["edges_total"] = int(len(allb.edges))
                assert allb.n_reads == n
            ctx.free_edges(allb)
        ctx.free_edges(e)
    out["overlap"] = {"wall_ms": round(wall, 2), "reads_per_s": round(n / (wall * 1e-3)),
                      "ms_search": round(mx(st["ms_search"]), 2), "ms_extract": round(mx(st["ms_extract"]), 2),
                      "slow_reads_rank0": st["slow_reads"], "lookups_rank0": st["lookups_search"]}
    for f in fms:
        ctx.free_fm(f)
    ctx.free_comm(comm)
    ctx.free_reads(r)
    ctx.close()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
