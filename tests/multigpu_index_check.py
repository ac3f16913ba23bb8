"""NCCL transport check of the bucket-sharded index build (needs >= 2 GPUs; run under torchrun on the GPU box).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tests/multigpu_index_check.py [cfg] [scale] [reps]

Every rank packs the same synthetic read set (cfg2 at `scale` by default), rank 0 also builds the index on its own
GPU with sgx_index_build; then all ranks run sgx_index_build_sharded over NCCL and rank 0 compares run bytes and .sai.
Prints one JSON line with the timings (device time = max over ranks).
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import torch.distributed as dist
    from siga_b200 import api, synth

    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    reads = synth.config_reads(cfg, scale)
    flat, off = synth.reads_to_flat(reads)
    ctx = api.Context(local)
    uid = torch.from_numpy(ctx.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
    dist.broadcast(uid, 0)
    comm = ctx.comm_init(uid.cpu().numpy(), rank, world)
    r = ctx.pack_reads(flat, off)
    # reference hashes of this (config, scale), when tests/golden/make_fullsize.py pinned them
    import hashlib
    gp = os.path.join(REPO, "tests", "golden", "fullsize", "%s_%g.json" % (cfg, scale))
    golden = json.load(open(gp))["index"] if os.path.isfile(gp) else None
    ok = True
    out = {"cfg": cfg, "scale": scale, "reads": int(reads.shape[0]), "read_len": int(reads.shape[1]), "world": world}
    for rev in (False, True):
        one = None
        if rank == 0:
            for _ in range(2):
                one = ctx.index_build(r, reversed=rev)
            out["single_ms_rev%d" % rev] = one.stats["ms_total"]
        for it in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = ctx.index_build_sharded(comm, r, reversed=rev, copy=False)  # views of the library's pinned buffers
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            if it + 1 < reps:
                ctx.free_index(res)
            t = torch.tensor([wall, res.stats["ms_total"], res.stats["ms_sort"], res.stats["ms_bcr"], res.stats["ms_emit"],
                              res.stats["ms_merge"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            same = np.array_equal(res.runs, one.runs) and np.array_equal(res.sai, one.sai)
            ok = ok and same
            if golden is not None:
                ref_same = (hashlib.sha256(res.runs.tobytes()).hexdigest() == golden["rbwt_runs_sha256" if rev else "bwt_runs_sha256"] and
                            hashlib.sha256(res.sai.astype("<u4").tobytes()).hexdigest() == golden["rsai_u32_sha256" if rev else "sai_u32_sha256"])
                out["reference_identical_rev%d" % rev] = bool(ref_same)
                ok = ok and ref_same
            out["sharded_rev%d" % rev] = dict(zip(["wall_ms", "ms_total", "ms_sort", "ms_bcr", "ms_emit", "ms_merge"],
                                                  [round(float(x), 2) for x in t.tolist()]), identical=bool(same),
                                              run_bytes=int(len(res.runs)), launches=res.stats["launches"])
        # the same build without the final gather: every rank's segment must be the slice of the single-GPU run bytes
        # at its offset, and the segments must tile the whole
        ctx.free_index(res)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seg = ctx.index_build_sharded_segments(comm, r, reversed=rev, copy=False)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([wall, seg.stats["ms_total"], seg.stats["ms_emit"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = (int(seg.seg_offset), int(len(seg.runs)), hashlib.sha256(seg.runs.tobytes()).hexdigest(), int(seg.n_runs_total),
                hashlib.sha256(seg.sai.tobytes()).hexdigest())
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            at, seg_ok = 0, True
            for off, ln, dig, tot, sdig in parts:
                seg_ok = seg_ok and off == at and tot == len(one.runs) and \
                    hashlib.sha256(one.runs[off:off + ln].tobytes()).hexdigest() == dig and \
                    sdig == hashlib.sha256(one.sai.tobytes()).hexdigest()
                at += ln
            seg_ok = seg_ok and at == len(one.runs)
            ok = ok and seg_ok
            out["segments_rev%d" % rev] = dict(wall_ms=round(float(t[0]), 2), ms_total=round(float(t[1]), 2),
                                               ms_emit=round(float(t[2]), 2), identical=bool(seg_ok))
        ctx.free_index(seg)
    ctx.free_comm(comm)
    ctx.free_reads(r)
    ctx.close()
    if rank == 0:
        out["ok"] = bool(ok)
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
