"""World-size-2 gloo test of the N>1 host path (read sharding + result gather); no GPU involved."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_reads, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.read_range(n_reads, rank, world)
    # stand-in for the per-rank search: every read yields one record naming itself, and an SS flag
    edges = ["r%d r%d" % (i, (i * 7) % n_reads) for i in range(first, first + count)]
    flags = np.array([(i % 5 == 0) for i in range(first, first + count)], dtype=np.uint8)
    all_edges = shard.gather_edges(edges, dist)
    all_flags = shard.gather_substring_flags(flags, n_reads, dist)
    if rank == 0:
        assert sorted(all_edges) == sorted("r%d r%d" % (i, (i * 7) % n_reads) for i in range(n_reads))
        assert all_flags.tolist() == [int(i % 5 == 0) for i in range(n_reads)]  # input order preserved
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert all_edges is None and all_flags is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 1001, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def _correct_worker(rank, world, port, fixture, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import correct_common as cc
    import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = cc.load(fixture)
    runs = np.frombuffer(bytes.fromhex(fx["bwt_runs_hex"]), dtype=np.uint8)
    n = len(fx["reads"])
    first, count = shard.read_range(n, rank, world)
    # the per-rank call of the sharded correction: the replicated forward index, this rank's contiguous read range
    # (what bench.py and `siga correct --gpus N` hand sgx_correct_batch), here on the single-lane host engine
    reads = fx["reads"][first:first + count]
    quals = fx["quals"][first:first + count] if fx["quals"] else None
    got, valid, _, _ = cc.hostemu_correct(cc.hostemu_lib(), runs, reads, quals, fx["options"])
    part = cc.ec_text(got, quals, valid, names=["r%d" % i for i in range(first, first + count)])
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0)
    if rank == 0:  # rank order = input order: no collective in the data path, only the final concatenation
        assert "".join(parts) == fx["ec_text"]
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_correction_equals_reference_file(tmp_path):
    import glob
    import pytest
    if not os.path.isfile(os.path.join(REPO, "build", "libhostemu.so")):
        pytest.skip("build/libhostemu.so not built (make hostemu)")
    fixture = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "correct", "*fastq_variable*.json")))[0]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_correct_worker, args=(2, port, fixture, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
