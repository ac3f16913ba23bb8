"""CPU tests of the N>1 index path: the bucket-sharded column-insertion algorithm with the product's own planning code
(siga_b200/csrc/shard_plan.h via build/libhostemu.so), emulated in one process and over two gloo ranks, against the
oracle's BWT.  The CUDA data path of the same algorithm is covered by tests/test_gpu_sharded_index.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import pyoracle  # noqa: E402
from siga_b200 import formats  # noqa: E402

HEMU = os.path.join(REPO, "build", "libhostemu.so")
pytestmark = pytest.mark.skipif(not os.path.isfile(HEMU), reason="build/libhostemu.so not built (make hostemu)")


def _read_sets():
    rng = np.random.default_rng(3)
    g = rng.integers(0, 4, 300)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    fixed = [acgt[g[s:s + 24]].tobytes() for s in rng.integers(0, 276, 40)]
    fixed[5] = fixed[9]
    var = [acgt[g[s:s + int(l)]].tobytes() for s, l in zip(rng.integers(0, 250, 30), rng.integers(4, 40, 30))]
    low = [b"A" * 70, b"A" * 70, b"A" * 33, b"AC" * 20, b"T" * 64, b"ACGT"]
    return {"fixed": fixed, "varlen": var, "lowcomplexity": low}


def _expected(reads, model):
    bwt, sai = pyoracle.bwt_and_sai(reads, model)
    return formats.encode_runs(bwt).tobytes(), sai


@pytest.mark.parametrize("slots", [False, True], ids=["packed_receive_buffer", "per_sender_slots"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("name", ["fixed", "varlen", "lowcomplexity"])
def test_sharded_column_passes_match_oracle(world, name, slots):
    """`slots`: the receive layout of the peer-memory transport (one worst-case-sized slot per sender in the receiver's
    region) instead of the packed buffer of the send/recv transport; the same planning code places the inserts."""
    import sharded_emu
    reads = _read_sets()[name]
    for model in ("B", "A"):
        want, sai = _expected(reads, model)
        segs, rows = sharded_emu.sharded_bwt_runs(reads, sharded_emu.sentinel_order(sai, model), sharded_emu.LocalNet(world),
                                                  slots=slots)
        assert sum(rows.values()) == sum(len(r) + 1 for r in reads)
        assert b"".join(segs[g] for g in range(world)) == want, (name, model, world)


def test_bounds_keep_sentinel_rows_on_shard_zero_and_are_monotone():
    import sharded_emu
    K = sharded_emu.consts()
    hist = np.zeros(K["key_space"], dtype=np.uint64)
    hist[0] = 1000  # only '$' rows: every bound collapses to 1, shards > 0 stay empty
    b = np.zeros(K["max_shards"] + 1, dtype=np.uint32)
    sharded_emu.lib().hemu_shard_bounds(sharded_emu._p(hist), 4, sharded_emu._p(b))
    assert b[0] == 0 and all(b[g] >= 1 for g in range(1, 4)) and list(b[:5]) == sorted(b[:5]) and b[4] == K["key_space"]
    rng = np.random.default_rng(1)
    hist = rng.integers(0, 50, K["key_space"]).astype(np.uint64)
    sharded_emu.lib().hemu_shard_bounds(sharded_emu._p(hist), 8, sharded_emu._p(b))
    cum = np.concatenate([[0], np.cumsum(hist)])
    shares = [int(cum[b[g + 1]] - cum[b[g]]) for g in range(8)]
    assert max(shares) - min(shares) <= 2 * int(hist.max())  # balanced up to one bucket


def _worker(rank, world, port, out_dir):
    for p in (REPO, os.path.join(REPO, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sharded_emu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for name, reads in _read_sets().items():
        want, sai = _expected(reads, "B")
        segs, _ = sharded_emu.sharded_bwt_runs(reads, sharded_emu.sentinel_order(sai, "B"), sharded_emu.GlooNet(dist))
        parts = [None] * world
        dist.all_gather_object(parts, segs[rank])  # the final gather of the run-byte segments
        ok = ok and b"".join(parts) == want
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_build_the_same_bwt(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
