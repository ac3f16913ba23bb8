"""GPU parity of the k-mer correction row: sgx_correct_batch and `siga correct` against the
compiled reference and the committed golden outputs."""
import gzip
import os
import subprocess
import tempfile

import numpy as np
import pytest

import correct_common as cc
import refrun

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "siga_b200", "bin", "siga")


def _gpu_correct(ctx, runs, n_suffixes, reads, quals, opt, first=0, count=None):
    flat, off, q = cc.flat_inputs(reads, quals)
    count = len(reads) - first if count is None else count
    rd = ctx.pack_reads(flat, off)
    fm = ctx.fm_create(runs, n_suffixes)
    try:
        qq = qo = None
        if q is not None:
            qq, qo = q, off[first:first + count + 1]
        res = ctx.correct(fm, rd, first, count, qual=qq, qual_off=qo, kmer_size=opt["k"], kmer_threshold=opt["x"],
                          kmer_rounds=opt["i"], kmer_count_offset=opt["O"])
    finally:
        ctx.free_fm(fm)
        ctx.free_reads(rd)
    return cc.apply_result(reads, res.apply(flat, off, first), off), res


@pytest.mark.parametrize("path", cc.GOLDEN, ids=[os.path.basename(p)[:-5] for p in cc.GOLDEN])
def test_cuda_correction_reproduces_golden_output(ctx, path):
    fx = cc.load(path)
    runs = np.frombuffer(bytes.fromhex(fx["bwt_runs_hex"]), dtype=np.uint8)
    reads, res = _gpu_correct(ctx, runs, fx["n_suffixes"], fx["reads"], fx["quals"], fx["options"])
    assert cc.ec_text(reads, fx["quals"], res.valid) == fx["ec_text"]
    lib = cc.hostemu_lib()
    if lib is not None:  # the executed-lookup counter of the kernel against the single-lane count
        _, _, nfix, lookups = cc.hostemu_correct(lib, runs, fx["reads"], fx["quals"], fx["options"])
        assert res.stats["lookups"] == lookups
        assert res.stats["n_fixes"] <= int(nfix.sum()) and res.stats["launches"] >= 1


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("case", [
    dict(seed=11, opt=dict(k=31, x=3, i=10, O=1), n_reads=4000, read_len=100, genome=8000, err=0.01),
    dict(seed=12, opt=dict(k=21, x=3, i=10, O=1), fastq=True, n_reads=3000, read_len=80, genome=5000),
    dict(seed=13, opt=dict(k=17, x=4, i=3, O=0), fastq=True, variable=True, n_reads=3000),
    dict(seed=14, opt=dict(k=1, x=3, i=10, O=1)),
    dict(seed=15, opt=dict(k=60, x=1, i=10, O=1)),
    dict(seed=16, opt=dict(k=19, x=2, i=0, O=1), both=True),
    dict(seed=17, opt=dict(k=15, x=-1, i=10, O=1)),
    dict(seed=18, opt=dict(k=25, x=2, i=10, O=1), n_reads=2500, read_len=150, genome=3000, err=0.03),
], ids=lambda c: "seed%d_k%d" % (c["seed"], c["opt"]["k"]))
def test_cuda_correction_equals_reference_binary(ctx, case):
    kw = {k: v for k, v in case.items() if k not in ("seed", "opt")}
    reads, quals = cc.random_case(case["seed"], **kw)
    runs, n_suffixes, want = cc.reference_correct(reads, quals, case["opt"])
    got, res = _gpu_correct(ctx, runs, n_suffixes, reads, quals, case["opt"])
    assert cc.ec_text(got, quals, res.valid) == want


def test_cuda_correction_sub_range_wide_positions_and_errors(ctx, monkeypatch):
    from siga_b200 import api
    fx = cc.load([p for p in cc.GOLDEN if "fastq_variable" in p][0])
    runs = np.frombuffer(bytes.fromhex(fx["bwt_runs_hex"]), dtype=np.uint8)
    full, res_full = _gpu_correct(ctx, runs, fx["n_suffixes"], fx["reads"], fx["quals"], fx["options"])
    # a sub-range of the read store (what a multi-GPU front-end hands each rank) gives the same reads
    part, res = _gpu_correct(ctx, runs, fx["n_suffixes"], fx["reads"], fx["quals"], fx["options"], first=37, count=101)
    assert part[37:138] == full[37:138] and list(res.valid) == list(res_full.valid[37:138])
    # 64-bit position kernel (indices of 2^32 symbols and more, e.g. config 4) forced onto the small fixture
    monkeypatch.setenv("SGX_FORCE_WIDE", "1")
    wide, res_w = _gpu_correct(ctx, runs, fx["n_suffixes"], fx["reads"], fx["quals"], fx["options"])
    monkeypatch.delenv("SGX_FORCE_WIDE")
    assert wide == full and list(res_w.valid) == list(res_full.valid) and res_w.stats["lookups"] == res_full.stats["lookups"]
    # argument errors: k = 0, quality strings that do not match the reads
    with pytest.raises(api.SgxError):
        _gpu_correct(ctx, runs, fx["n_suffixes"], fx["reads"], fx["quals"], dict(fx["options"], k=0))
    bad_q = [q + "I" for q in fx["quals"]]
    with pytest.raises(api.SgxError):
        flat, off, _ = cc.flat_inputs(fx["reads"], None)
        qflat, qoff, _ = cc.flat_inputs(bad_q, None)
        rd = ctx.pack_reads(flat, off)
        fm = ctx.fm_create(runs, fx["n_suffixes"])
        try:
            ctx.correct(fm, rd, 0, len(fx["reads"]), qual=qflat, qual_off=qoff, kmer_size=13)
        finally:
            ctx.free_fm(fm)
            ctx.free_reads(rd)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
@pytest.mark.parametrize("fastq", [False, True], ids=["fasta", "fastq"])
def test_cli_correct_writes_the_reference_file(fastq):
    reads, quals = cc.random_case(21 + int(fastq), n_reads=3000, read_len=75, genome=6000, err=0.015, fastq=fastq, both=True)
    names = ["q%d%s" % (i, " CR:i:%d note" % (i % 5) if i % 3 == 0 else "") for i in range(len(reads))]
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        src = os.path.join(d, "reads.fq" if fastq else "reads.fa")
        cc.write_reads(src, reads, quals, names)
        refrun.run_ref(["index", src], ref)
        refrun.run_ref(["correct", "-t", "3", "-k", "23", src], ref)
        for cwd in (ours,):
            r = subprocess.run([CLI, "index", src], cwd=cwd, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            r = subprocess.run([CLI, "correct", "-k", "23", src], cwd=cwd, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
        want = open(os.path.join(ref, "reads.ec.fa"), "rb").read()
        assert open(os.path.join(ours, "reads.ec.fa"), "rb").read() == want
        # -o FILE.gz and -p PREFIX
        r = subprocess.run([CLI, "correct", "-k", "23", "-p", os.path.join(ref, "reads"), "-o", "out.fa.gz", src], cwd=ours,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert gzip.open(os.path.join(ours, "out.fa.gz"), "rb").read() == want
        r = subprocess.run([CLI, "correct", "-a", "overlap", src], cwd=ours, capture_output=True, text=True)
        assert r.returncode != 0
