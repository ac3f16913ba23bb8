"""GPU parity tests of the bucket-sharded index build (sgx_index_build_sharded).

The driver's GPU test box has ONE device, so these run the sharded algorithm over the `local` transport (virtual
shards on one context: the exchange steps are device copies instead of NCCL send/recv; everything else -- bucket
bounds, per-shard ranks, destination partition, receiver-side row fix-up, seam-aware run-length encoding -- is the code
the NCCL path runs).  The NCCL transport itself is exercised by `bench.py --gpus N` / tests/multigpu_index_check.py.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle
from siga_b200 import api, formats, synth
from test_gpu_parity import _rand_reads

pytestmark = pytest.mark.gpu


def _check_sharded(ctx, reads, shards, label, oracle=True):
    flat, off = synth.reads_to_flat(reads)
    r = ctx.pack_reads(flat, off)
    comm = ctx.comm_init_local(shards)
    try:
        for model, order in (("B", api.ORDER_SAIS2), ("A", api.ORDER_SAIS)):
            for rev in (False, True):
                one = ctx.index_build(r, order=order, reversed=rev)
                res = ctx.index_build_sharded(comm, r, order=order, reversed=rev)
                assert res.n_strings == one.n_strings and res.n_suffixes == one.n_suffixes
                assert np.array_equal(res.sai, one.sai), (label, shards, model, rev, "sai")
                if not np.array_equal(res.runs, one.runs):
                    a, b = formats.decode_runs(res.runs), formats.decode_runs(one.runs)
                    n = min(len(a), len(b))
                    bad = np.flatnonzero(a[:n] != b[:n])
                    raise AssertionError((label, shards, model, rev, "rows", len(a), len(b),
                                          "first bad row", int(bad[0]) if len(bad) else None, "run bytes", len(res.runs), len(one.runs)))
                # without the final gather: one process holds every shard of a local communicator, so its "segment" is the whole
                seg = ctx.index_build_sharded_segments(comm, r, order=order, reversed=rev)
                assert seg.seg_offset == 0 and seg.n_runs_total == len(one.runs)
                assert np.array_equal(seg.runs, one.runs) and np.array_equal(seg.sai, one.sai), (label, shards, model, rev, "segments")
                if oracle:
                    bwt, sai = pyoracle.bwt_and_sai(reads, model, reversed_reads=rev)
                    assert np.array_equal(res.runs, formats.encode_runs(bwt)), (label, shards, model, rev, "oracle")
                    assert np.array_equal(res.sai, sai)
    finally:
        ctx.free_comm(comm)
        ctx.free_reads(r)


@pytest.mark.parametrize("shards", [1, 2, 3, 4, 8, 16])
def test_sharded_index_matches_oracle_small(ctx, shards):
    rng = np.random.default_rng(500 + shards)
    _check_sharded(ctx, _rand_reads(rng, 80, 30, 30, 400, err=0.02), shards, "fixed30")
    _check_sharded(ctx, _rand_reads(rng, 90, 5, 40, 200, dup=0.1), shards, "varlen")


@pytest.mark.parametrize("shards", [2, 5, 8])
def test_sharded_index_runs_across_seams(ctx, shards):
    """Low-complexity reads: runs far longer than 31 that span shard seams, shards that stay empty, whole shards
    made of a single run."""
    rng = np.random.default_rng(77)
    reads = [b"A" * int(rng.integers(1, 70)) for _ in range(12)] + [b"AC" * 20, b"CA" * 20, b"T" * 64, b"T" * 64]
    _check_sharded(ctx, reads, shards, "lowcomplexity")
    _check_sharded(ctx, [b"A" * 200, b"A" * 200, b"A" * 150], shards, "polyA")
    _check_sharded(ctx, [b"ACGT"], shards, "one-read")
    base = _rand_reads(rng, 6, 12, 12, 40, both=False)
    _check_sharded(ctx, [base[int(rng.integers(0, 6))] for _ in range(70)], shards, "dups")


@pytest.mark.parametrize("shards", [2, 8])
def test_sharded_index_medium_matches_single_device(ctx, shards):
    """20 k reads x 100 bp (config 1 shape): sharded == single-device build (itself checked against the reference)."""
    rng = np.random.default_rng(9)
    g = synth.random_genome(200_000, rng)
    reads = synth.sample_reads(g, 20_000, 100, rng, err=0.01)
    _check_sharded(ctx, [x.tobytes() for x in reads], shards, "medium", oracle=False)


def test_sharded_index_wide_positions(ctx, monkeypatch):
    monkeypatch.setenv("SGX_FORCE_WIDE", "1")
    rng = np.random.default_rng(31)
    _check_sharded(ctx, _rand_reads(rng, 120, 10, 50, 500, err=0.01, dup=0.05), 4, "wide")


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("cfg,scale", [("cfg4", 0.01), ("cfg5", 0.01)])
def test_sharded_index_over_nccl_matches_reference(cfg, scale, transport):
    """One process per GPU, both transports of the inserts -- peer memory (cudaIpc-mapped receive regions, stores over
    NVLink, arrival flags) and the fall-back (ncclSend/ncclRecv all-to-all per column pass): run bytes and .sai of the
    sharded build equal the single-GPU build AND the reference's hashes (tests/golden/fullsize); the per-rank segments of
    sgx_index_build_sharded_segments tile the same bytes."""
    here = os.path.dirname(os.path.abspath(__file__))
    n = min(_gpu_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(here, "multigpu_index_check.py"), cfg, str(scale), "1"]
    env = dict(os.environ, SGX_SHARD_TRACE="1")
    if transport == "nccl":
        env["SGX_SHARD_XCHG"] = "nccl"
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    assert ("peer-memory transport" if transport == "peer" else "NCCL transport") in p.stderr, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["world"] == n
    assert out["sharded_rev0"]["identical"] and out["sharded_rev1"]["identical"]
    assert out["segments_rev0"]["identical"] and out["segments_rev1"]["identical"]
    assert out["reference_identical_rev0"] and out["reference_identical_rev1"]
