"""GPU parity tests: every call goes through the C ABI (siga_b200/api.py -> libsiga_b200.so)."""
import os
import tempfile

import numpy as np
import pytest

import refrun
from oracle import pyoracle
from siga_b200 import api, formats, synth

pytestmark = pytest.mark.gpu


def _rand_reads(rng, n, lmin, lmax, genome_len, err=0.0, dup=0.0, both=True):
    g = synth.random_genome(genome_len, rng)
    reads = []
    for _ in range(n):
        l = int(rng.integers(lmin, lmax + 1))
        s = int(rng.integers(0, genome_len - l + 1))
        r = g[s:s + l].copy()
        if err > 0:
            m = rng.random(l) < err
            r[m] = synth.random_genome(int(m.sum()), rng)
        if both and rng.random() < 0.5:
            r = synth.revcomp(r)
        reads.append(np.ascontiguousarray(r).tobytes())
    for _ in range(int(n * dup)):
        reads[int(rng.integers(0, n))] = reads[int(rng.integers(0, n))]
    return reads


# ---------------------------------------------------------------------------------------------- rank structure

@pytest.mark.parametrize("n", [1, 31, 255, 256, 257, 4096, 100_003])
def test_fm_occ_matches_cumulative_counts(ctx, n):
    rng = np.random.default_rng(n)
    sym = rng.choice(5, size=n, p=[0.02, 0.3, 0.2, 0.2, 0.28]).astype(np.uint8)
    if n > 1000:  # long runs, including > 31
        sym[100:400] = 1
        sym[5000:5100] = 0
    runs = formats.encode_runs(sym)
    fm = ctx.fm_create(runs, n)
    try:
        occ = pyoracle.occ_table(sym)
        pos = np.unique(np.concatenate([np.arange(min(n, 600)), rng.integers(0, n, size=2000), [n - 1]])).astype(np.uint64)
        got = ctx.fm_occ(fm, np.concatenate([pos, np.array([2**64 - 1], dtype=np.uint64)]))
        assert np.array_equal(got[:-1], occ[pos.astype(np.int64)])
        assert not got[-1].any()  # getOcc(-1) == 0 (src/fmindex.cpp:188-231: ++i wraps to 0)
        cnt = np.bincount(sym, minlength=5)
        assert ctx.fm_pred(fm) == [int(x) for x in np.concatenate([[0], np.cumsum(cnt)[:-1]])]
    finally:
        ctx.free_fm(fm)


def test_fm_count_matches_bruteforce(ctx):
    rng = np.random.default_rng(5)
    reads = _rand_reads(rng, 60, 20, 30, 300)
    bwt, _ = pyoracle.bwt_and_sai(reads, "B")
    fm = ctx.fm_create(formats.encode_runs(bwt), len(bwt))
    try:
        qs = [reads[int(rng.integers(0, 60))][a:a + k] for a, k in zip(rng.integers(0, 10, 200), rng.integers(1, 12, 200))]
        qs += [b"ACGTACGTACGTACGT", b"A", b"T"]
        flat, off = synth.reads_to_flat(qs)
        got = ctx.fm_count(fm, flat, off)
        text = b"$".join(reads)
        want = []
        for q in qs:
            c, s = 0, text.find(q)
            while s >= 0:
                c += 1
                s = text.find(q, s + 1)
            want.append(c)
        assert got.tolist() == want
    finally:
        ctx.free_fm(fm)


# ---------------------------------------------------------------------------------------------- index construction

def _check_index(ctx, reads, label):
    flat, off = synth.reads_to_flat(reads)
    r = ctx.pack_reads(flat, off)
    try:
        for model, order in (("B", api.ORDER_SAIS2), ("A", api.ORDER_SAIS)):
            for rev in (False, True):
                res = ctx.index_build(r, order=order, reversed=rev)
                bwt, sai = pyoracle.bwt_and_sai(reads, model, reversed_reads=rev)
                assert res.n_strings == len(reads) and res.n_suffixes == len(bwt)
                got = formats.decode_runs(res.runs)
                assert len(got) == len(bwt), (label, model, rev)
                bad = np.flatnonzero(got != bwt)
                assert len(bad) == 0, (label, model, rev, "first bad row", int(bad[0]), got[max(0, bad[0] - 5):bad[0] + 5], bwt[max(0, bad[0] - 5):bad[0] + 5])
                assert np.array_equal(res.sai, sai), (label, model, rev)
                assert np.array_equal(res.runs, formats.encode_runs(bwt)), (label, model, rev, "run bytes")
    finally:
        ctx.free_reads(r)


def test_index_golden_f1(ctx):
    reads = ("CAGATTTTCATATTATGCAG TTTCATATTATGCAGAAAAT TATTATGCAGAAAATCTACT TGCAGAAAATCTACTTCGCC AAAATCTACTTCGCCTGATA "
             "CTACTTCGCCTGATACGAGT TCGCCTGATACGAGTCGGTT TGATACGAGTCGGTTATCTT CGAGTCGGTTATCTTCGGAT").split()
    flat, off = synth.reads_to_flat(reads)
    r = ctx.pack_reads(flat, off)
    try:
        res = ctx.index_build(r)
        assert res.sai.tolist() == [4, 0, 8, 5, 2, 6, 7, 3, 1]
        assert "".join("$ACGT"[s] for s in formats.decode_runs(res.runs)).startswith("TCTATTTTGTGGG$AAAAAAAATTTTTTTCCCCC")
        assert len(res.runs) == 71
        rres = ctx.index_build(r, reversed=True)
        assert rres.sai.tolist() == [4, 3, 0, 1, 8, 2, 5, 7, 6]  # .rsai
        assert len(rres.runs) == 74
    finally:
        ctx.free_reads(r)
    _check_index(ctx, reads, "f1")


@pytest.mark.parametrize("seed", range(8))
def test_index_random_sets_both_models(ctx, seed):
    rng = np.random.default_rng(100 + seed)
    kind = seed % 4
    if kind == 0:
        reads = _rand_reads(rng, 80, 30, 30, 400, err=0.02)
    elif kind == 1:  # heavy duplication: few distinct strings, ties resolved by the following reads
        base = _rand_reads(rng, 6, 12, 12, 40, both=False)
        reads = [base[int(rng.integers(0, 6))] for _ in range(70)]
    elif kind == 2:  # variable length incl. reads that are prefixes / substrings of others
        reads = _rand_reads(rng, 90, 5, 40, 200, dup=0.1)
    else:  # low complexity, runs > 31
        reads = [b"A" * int(rng.integers(1, 70)) for _ in range(12)] + [b"AC" * 20, b"CA" * 20, b"T" * 64, b"T" * 64]
    _check_index(ctx, reads, "seed%d" % seed)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("case", ["cfg1_small", "cfg2_small", "varlen"])
def test_index_bytes_match_reference_files(ctx, case):
    rng = np.random.default_rng(7)
    if case == "cfg1_small":
        g = synth.random_genome(60_000, rng)
        reads = synth.paired_reads(g, 600, 100, rng)
    elif case == "cfg2_small":
        g = synth.random_genome(20_000, rng)
        reads = synth.sample_reads(g, 13_000, 150, rng, err=0.01)
    else:
        reads = _rand_reads(rng, 3000, 30, 120, 30_000, err=0.01, dup=0.05)
    flat, off = synth.reads_to_flat(reads)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "x.fa")
        synth.write_fasta(fa, reads)
        r = ctx.pack_reads(flat, off)
        try:
            for algo, order in ((None, api.ORDER_SAIS2), ("sais", api.ORDER_SAIS)):
                refrun.ref_index(fa, d, algo=algo)
                for rev, be, se in ((False, ".bwt", ".sai"), (True, ".rbwt", ".rsai")):
                    res = ctx.index_build(r, order=order, reversed=rev)
                    want_bwt = open(os.path.join(d, "x" + be), "rb").read()
                    want_sai = open(os.path.join(d, "x" + se), "rb").read()
                    assert formats.bwt_bytes(res.n_strings, res.n_suffixes, res.runs) == want_bwt, (case, algo, be)
                    assert formats.sai_bytes(res.sai) == want_sai, (case, algo, se)
        finally:
            ctx.free_reads(r)


# ---------------------------------------------------------------------------------------------- overlap search

def _overlap_vs_ref(ctx, reads, m, exhaustive=False, no_rc=False, names=None):
    flat, off = synth.reads_to_flat(reads)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "x.fa")
        synth.write_fasta(fa, reads, names=names)
        refrun.ref_index(fa, d)
        hitfiles, asqg = refrun.ref_overlap(fa, d, m, threads=4, exhaustive=exhaustive, no_rc=no_rc)
        ref_hits = formats.read_hits(hitfiles)
        ref_asqg = formats.read_asqg(asqg)
        n_str, n_suf, fr = formats.read_bwt(os.path.join(d, "x.bwt"))
        _, _, rr = formats.read_bwt(os.path.join(d, "x.rbwt"))
        sai = formats.read_sai(os.path.join(d, "x.sai"))
        rsai = formats.read_sai(os.path.join(d, "x.rsai"))
    fwd = ctx.fm_create(fr, n_suf)
    rev = ctx.fm_create(rr, n_suf)
    r = ctx.pack_reads(flat, off)
    try:
        # general path: extension by BWT walks (no read store attached)
        h_walk = ctx.overlap(fwd, rev, r, 0, len(off) - 1, m, irreducible=not exhaustive, rc=not no_rc)
        assert h_walk.stats["slow_reads"] == len(off) - 1
        # fast path: .sai/.rsai + read store attached, extension read from the target reads
        ctx.fm_attach_reads(fwd, sai, r)
        ctx.fm_attach_reads(rev, rsai, r)
        h = ctx.overlap(fwd, rev, r, 0, len(off) - 1, m, irreducible=not exhaustive, rc=not no_rc)
        hh = ctx.overlap_host(fwd, rev, flat, off, m, irreducible=not exhaustive, rc=not no_rc)
        nm0 = names if names is not None else ["r%d" % i for i in range(len(off) - 1)]
        dev_edges = ctx.overlap_edges(fwd, rev, r, 0, len(off) - 1, m, api.name_ranks(nm0), irreducible=not exhaustive,
                                      rc=not no_rc)
        assert np.array_equal(h.block_off, h_walk.block_off) and np.array_equal(h.substring, h_walk.substring)
        diff = [i for i in range(len(off) - 1) if h.read_blocks(i) != h_walk.read_blocks(i)]
        assert not diff, ("fast path differs from the BWT-walking path", diff[:5], h.read_blocks(diff[0]), h_walk.read_blocks(diff[0]))
    finally:
        ctx.free_reads(r)
        ctx.free_fm(fwd)
        ctx.free_fm(rev)
    n = len(off) - 1
    bad = []
    for i in range(n):
        sub, blocks = ref_hits[i]
        if h.read_blocks(i) != sorted(blocks) or int(h.substring[i]) != sub:
            bad.append(i)
    assert not bad, ("reads with different blocks", bad[:5], h.read_blocks(bad[0]), sorted(ref_hits[bad[0]][1]))
    assert np.array_equal(h.block_off, hh.block_off) and np.array_equal(h.substring, hh.substring)
    assert all(h.read_blocks(i) == hh.read_blocks(i) for i in range(0, n, max(1, n // 50)))
    # hits -> ASQG through the host-side converter: identical to the reference after canonical sorting
    nm = names if names is not None else ["r%d" % i for i in range(n)]
    seqs = [bytes(x).decode() for x in reads]
    lens = [len(s) for s in seqs]
    edges = formats.hits_to_edges(nm, lens, h.substring, h.block_off, h.blocks, sai, rsai)
    mine = formats.asqg_lines(nm, seqs, h.substring, edges, m)
    assert sorted(mine) == sorted(ref_asqg)
    assert mine[:1 + n] == ref_asqg[:1 + n]  # HT + VT lines come in input order
    # device-side hits -> ED conversion (sgx_overlap_edges) yields the same edge set
    assert sorted(formats.edges_to_ed(nm, lens, dev_edges.edges)) == sorted(edges)
    assert np.array_equal(dev_edges.substring, h.substring)
    return h


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("seed", range(6))
def test_overlap_blocks_match_reference_hits(ctx, seed):
    rng = np.random.default_rng(200 + seed)
    g = synth.random_genome(int(rng.integers(500, 3000)), rng)
    L = int(rng.integers(30, 80))
    reads = synth.sample_reads(g, int(rng.integers(200, 900)), L, rng, err=0.01 if seed % 2 else 0.0)
    reads[rng.integers(0, len(reads), 20)] = reads[rng.integers(0, len(reads), 20)]  # duplicates -> containments
    h = _overlap_vs_ref(ctx, reads, int(rng.integers(12, L - 5)), exhaustive=(seed == 4), no_rc=(seed == 5))
    assert h.stats["slow_reads"] <= len(reads) // 20


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
def test_overlap_periodic_and_variable_length(ctx):
    rng = np.random.default_rng(31)
    unit = synth.random_genome(3, rng)
    g = np.tile(unit, 150)
    mut = rng.random(len(g)) < 0.03
    g[mut] = synth.random_genome(int(mut.sum()), rng)
    reads = synth.sample_reads(g, 60, 30, rng)
    h = _overlap_vs_ref(ctx, reads, 10)  # exercises SubMaximalBlockFilter::resolve's tracing path
    assert h.stats["slow_reads"] > 0     # periodic reads are handed to the literal path
    reads = _rand_reads(rng, 150, 12, 40, 400)
    h = _overlap_vs_ref(ctx, reads, 8)   # substring reads: SS flags and the extraction abort order
    assert h.stats["slow_reads"] > 0     # variable-length targets are handed to the literal path


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
def test_golden_f2_names_and_containment(ctx):
    # duplicates, reverse complement, substring; names decide which end is id0
    reads = [b"ACGTTGCAAGGCTTAACCGGATTACA", b"TTGCAAGGCTTAACCGGATTACAGGC", b"ACGTTGCAAGGCTTAACCGGATTACA", b"GCAAGGCTTAACCGG",
             b"TGTAATCCGGTTAAGCCTTGCAACGT", b"CCGGATTACAGGCTTTTGACCA"]
    h = _overlap_vs_ref(ctx, reads, 8, names=list("abcdef"))
    assert h.substring.tolist() == [0, 0, 0, 1, 0, 0]


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
def test_config1_full_pipeline_bit_exact(ctx):
    """BASELINE.json configs[0]: 1 Mbp genome, 20k error-free 100 bp reads, index + overlap -m 45."""
    reads = synth.config_reads("cfg1")
    assert reads.shape == (20_000, 100)
    flat, off = synth.reads_to_flat(reads)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "cfg1.fa")
        synth.write_fasta(fa, reads)
        refrun.ref_index(fa, d)
        _, asqg = refrun.ref_overlap(fa, d, 45, threads=os.cpu_count() or 1)
        ref_asqg = formats.read_asqg(asqg)
        want = {e: open(os.path.join(d, "cfg1" + e), "rb").read() for e in (".bwt", ".rbwt", ".sai", ".rsai")}
    r = ctx.pack_reads(flat, off)
    try:
        fi = ctx.index_build(r)
        ri = ctx.index_build(r, reversed=True)
        assert formats.bwt_bytes(fi.n_strings, fi.n_suffixes, fi.runs) == want[".bwt"]
        assert formats.bwt_bytes(ri.n_strings, ri.n_suffixes, ri.runs) == want[".rbwt"]
        assert formats.sai_bytes(fi.sai) == want[".sai"]
        assert formats.sai_bytes(ri.sai) == want[".rsai"]
        fwd = ctx.fm_create(fi.runs, fi.n_suffixes)
        rev = ctx.fm_create(ri.runs, ri.n_suffixes)
        try:
            ctx.fm_attach_reads(fwd, fi.sai, r)
            ctx.fm_attach_reads(rev, ri.sai, r)
            h = ctx.overlap(fwd, rev, r, 0, 20_000, 45)
            assert h.stats["slow_reads"] == 0  # fixed-length, non-periodic reads all take the read-based path
            e = ctx.overlap_edges(fwd, rev, r, 0, 20_000, 45, synth.default_name_ranks(20_000))
            if os.environ.get("SGX_MAX_BATCH"):
                assert h.stats["search_launches"] > 1 and e.stats["search_launches"] > 1
        finally:
            ctx.free_fm(fwd)
            ctx.free_fm(rev)
    finally:
        ctx.free_reads(r)
    names = ["r%d" % i for i in range(20_000)]
    seqs = [x.tobytes().decode() for x in reads]
    edges = formats.hits_to_edges(names, [100] * 20_000, h.substring, h.block_off, h.blocks, fi.sai, ri.sai)
    assert sorted(formats.asqg_lines(names, seqs, h.substring, edges, 45)) == sorted(ref_asqg)
    assert sorted(formats.edges_to_ed(names, [100] * 20_000, e.edges)) == sorted(edges)
    assert np.array_equal(e.substring, h.substring)


def test_rejects_bad_input(ctx):
    flat, off = synth.reads_to_flat([b"ACGT", b"ACNT"])
    with pytest.raises(api.SgxError) as e:
        ctx.pack_reads(flat, off)
    assert e.value.code == -1
    with pytest.raises(api.SgxError):
        ctx.pack_reads(np.frombuffer(b"ACGT", dtype=np.uint8), np.array([0, 0, 4], dtype=np.uint64))  # empty read


# ---------------------------------------------------------------------------------------------- committed golden vectors

import glob as _glob
import json as _json

_GOLDEN = sorted(_glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.json")))


@pytest.mark.parametrize("path", _GOLDEN, ids=[os.path.basename(p)[:-5] for p in _GOLDEN])
def test_cuda_path_reproduces_golden_fixture(ctx, path):
    """Index bytes, hit blocks and ASQG of the committed reference outputs (tests/golden/make_golden.py)."""
    fx = _json.load(open(path))
    reads = [r.encode() for r in fx["reads"]]
    flat, off = synth.reads_to_flat(reads)
    n = len(reads)
    r = ctx.pack_reads(flat, off)
    try:
        built = {}
        for algo, order in (("sais2", api.ORDER_SAIS2), ("sais", api.ORDER_SAIS)):
            for rev, bk, sk in ((False, "bwt_hex", "sai"), (True, "rbwt_hex", "rsai")):
                res = ctx.index_build(r, order=order, reversed=rev)
                assert formats.bwt_bytes(res.n_strings, res.n_suffixes, res.runs).hex() == fx["index_" + algo][bk], (algo, bk)
                assert res.sai.tolist() == fx["index_" + algo][sk], (algo, sk)
                built[(algo, rev)] = res
        fi, ri = built[("sais2", False)], built[("sais2", True)]
        fwd = ctx.fm_create(fi.runs, fi.n_suffixes)
        rev = ctx.fm_create(ri.runs, ri.n_suffixes)
        try:
            opt = fx["options"]
            kw = dict(irreducible=not opt.get("exhaustive", False), rc=not opt.get("no_rc", False))
            for attach in (False, True):
                if attach:
                    ctx.fm_attach_reads(fwd, fi.sai, r)
                    ctx.fm_attach_reads(rev, ri.sai, r)
                h = ctx.overlap(fwd, rev, r, 0, n, fx["min_overlap"], **kw)
                for i, (sub, want) in enumerate(fx["hits"]):
                    assert h.read_blocks(i) == [tuple(b) for b in want], (fx["name"], "attach" if attach else "walk", i)
                    assert int(h.substring[i]) == sub
            edges = formats.hits_to_edges(fx["names"], [len(x) for x in reads], h.substring, h.block_off, h.blocks, fi.sai, ri.sai)
            assert sorted(formats.asqg_lines(fx["names"], fx["reads"], h.substring, edges, fx["min_overlap"])) == fx["asqg_sorted"]
        finally:
            ctx.free_fm(fwd)
            ctx.free_fm(rev)
    finally:
        ctx.free_reads(r)


def test_wide_position_paths_on_small_input(ctx, monkeypatch):
    """The 64-bit position kernels (BWT >= 2^32 symbols, e.g. config 4) forced onto a golden fixture."""
    monkeypatch.setenv("SGX_FORCE_WIDE", "1")
    for path in _GOLDEN:
        if any(k in path for k in ("random_fixed36_dups", "variable_length", "periodic30")):
            test_cuda_path_reproduces_golden_fixture(ctx, path)


# ------------------------------------------------------------------------- the five BASELINE.json configs, scaled down

@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("cfg,scale,m", [("cfg2", 0.004, 45), ("cfg3", 0.004, 31), ("cfg3", 0.004, 63), ("cfg3", 0.004, 101),
                                         ("cfg4", 0.0002, 45), ("cfg5", 0.0003, 45)])
def test_named_configs_scaled_down_match_reference(ctx, cfg, scale, m):
    """Read shape, coverage, error rate and min-overlap of configs 2-5 (BASELINE.json) on a genome small enough for the
    reference binary: .bwt/.rbwt/.sai/.rsai bytes and the canonically sorted ASQG must be identical, through the
    single-device build, the bucket-sharded build (4 virtual shards) and sgx_overlap_edges."""
    reads = synth.config_reads(cfg, scale)
    n, L = reads.shape
    flat, off = synth.reads_to_flat(reads)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "x.fa")
        synth.write_fasta(fa, reads)
        refrun.ref_index(fa, d)
        _, asqg = refrun.ref_overlap(fa, d, m, threads=os.cpu_count() or 1)
        ref_asqg = formats.read_asqg(asqg)
        want = {e: open(os.path.join(d, "x" + e), "rb").read() for e in (".bwt", ".rbwt", ".sai", ".rsai")}
    r = ctx.pack_reads(flat, off)
    comm = ctx.comm_init_local(4)
    try:
        fi = ctx.index_build(r)
        ri = ctx.index_build(r, reversed=True)
        assert formats.bwt_bytes(fi.n_strings, fi.n_suffixes, fi.runs) == want[".bwt"]
        assert formats.bwt_bytes(ri.n_strings, ri.n_suffixes, ri.runs) == want[".rbwt"]
        assert formats.sai_bytes(fi.sai) == want[".sai"] and formats.sai_bytes(ri.sai) == want[".rsai"]
        for rv, one in ((False, fi), (True, ri)):
            sh = ctx.index_build_sharded(comm, r, reversed=rv)
            assert np.array_equal(sh.runs, one.runs) and np.array_equal(sh.sai, one.sai)
        fwd = ctx.fm_create(fi.runs, fi.n_suffixes)
        rev = ctx.fm_create(ri.runs, ri.n_suffixes)
        try:
            ctx.fm_attach_reads(fwd, fi.sai, r)
            ctx.fm_attach_reads(rev, ri.sai, r)
            names = ["r%d" % i for i in range(n)]
            e = ctx.overlap_edges(fwd, rev, r, 0, n, m, api.name_ranks(names))
        finally:
            ctx.free_fm(fwd)
            ctx.free_fm(rev)
    finally:
        ctx.free_comm(comm)
        ctx.free_reads(r)
    seqs = [x.tobytes().decode() for x in reads]
    ed = formats.edges_to_ed(names, [L] * n, e.edges)
    mine = formats.asqg_lines(names, seqs, e.substring, ed, m)
    assert sorted(mine) == sorted(ref_asqg), (cfg, m, len(mine), len(ref_asqg))


# ------------------------------------------------------------ multi-batch loops, regrow-and-redo, edge-buffer overflow
# parallel::foreach hands the reference's workers `threads * batch` reads at a time (src/parallel_framework.h:35-45,
# src/overlap_builder.cpp:453-480); here a batch is a kernel launch over a slice of the read range.  The production
# batch (~0.8 M reads) never splits a test-sized input, so the hooks of overlap.cu (SGX_MAX_BATCH, SGX_OUT_CAP,
# SGX_EDGE_CAP) force many batches / undersized output regions / an undersized edge buffer onto the same
# reference-checked cases as above.

@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("hook", ["batch1024", "batch1000_outcap4", "edgecap64"])
def test_config1_full_pipeline_multibatch(ctx, monkeypatch, hook):
    if hook.startswith("batch1024"):
        monkeypatch.setenv("SGX_MAX_BATCH", "1024")
    if hook == "batch1000_outcap4":
        monkeypatch.setenv("SGX_MAX_BATCH", "1000")
        monkeypatch.setenv("SGX_OUT_CAP", "4")
    if hook == "edgecap64":
        monkeypatch.setenv("SGX_MAX_BATCH", "3000")
        monkeypatch.setenv("SGX_EDGE_CAP", "64")
    test_config1_full_pipeline_bit_exact(ctx)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("k", ["2", "7", "10"])
def test_config1_with_kmer_start_table(ctx, monkeypatch, k):
    """The opt-in k-mer start table of the lean search (SGX_KTAB=k, fm.cu: fm_build_ktab): every chain starts from the
    tabulated interval of its last k symbols; same files as the reference."""
    monkeypatch.setenv("SGX_KTAB", k)
    test_config1_full_pipeline_bit_exact(ctx)


@pytest.mark.parametrize("batch", ["1", "7"])
def test_golden_fixtures_multibatch(ctx, monkeypatch, batch):
    monkeypatch.setenv("SGX_MAX_BATCH", batch)
    for path in _GOLDEN:
        test_cuda_path_reproduces_golden_fixture(ctx, path)


def test_golden_fixtures_regrow_and_edge_overflow(ctx, monkeypatch):
    monkeypatch.setenv("SGX_MAX_BATCH", "16")
    monkeypatch.setenv("SGX_OUT_CAP", "4")
    monkeypatch.setenv("SGX_EDGE_CAP", "8")
    for path in _GOLDEN:
        test_cuda_path_reproduces_golden_fixture(ctx, path)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("seed,batch", [(0, "64"), (1, "100"), (4, "33"), (5, "50")])
def test_overlap_blocks_match_reference_hits_multibatch(ctx, monkeypatch, seed, batch):
    monkeypatch.setenv("SGX_MAX_BATCH", batch)
    monkeypatch.setenv("SGX_EDGE_CAP", "1000")
    test_overlap_blocks_match_reference_hits(ctx, seed)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
def test_overlap_periodic_and_variable_length_multibatch(ctx, monkeypatch):
    monkeypatch.setenv("SGX_MAX_BATCH", "16")   # the general-path chunks of sgx_overlap_edges split too
    monkeypatch.setenv("SGX_OUT_CAP", "8")
    test_overlap_periodic_and_variable_length(ctx)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("cfg,scale,m", [("cfg2", 0.004, 45), ("cfg3", 0.004, 31), ("cfg4", 0.0002, 45), ("cfg5", 0.0003, 45)])
def test_named_configs_scaled_down_multibatch(ctx, monkeypatch, cfg, scale, m):
    monkeypatch.setenv("SGX_MAX_BATCH", "1024")
    test_named_configs_scaled_down_match_reference(ctx, cfg, scale, m)


# ------------------------------------------------------------------------------------------------------------ rmdup

@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
def test_rmdup_records_match_reference(ctx):
    """`siga rmdup` (src/overlap_builder.cpp:623-704, 1184-1195) through the C ABI: a search at min-overlap = |read|
    leaves the containment blocks; NumDuplicates and the kept / removed split must equal the reference's files."""
    rng = np.random.default_rng(61)
    g = synth.random_genome(2500, rng)
    reads = [r.tobytes() for r in synth.sample_reads(g, 500, 36, rng, err=0.0)]
    for _ in range(80):
        a, b = rng.integers(0, len(reads), 2)
        reads[a] = reads[b] if rng.random() < 0.5 else synth.revcomp(np.frombuffer(reads[b], dtype=np.uint8)).tobytes()
    flat, off = synth.reads_to_flat(reads)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "x.fa")
        synth.write_fasta(fa, reads)
        refrun.ref_index(fa, d)
        refrun.run_ref(["rmdup", "-t", "2", fa], d)
        want = [sorted(">" + r for r in open(os.path.join(d, f)).read().split(">") if r) for f in ("x.rmdup.fa", "x.rmdup.dups.fa")]
    r = ctx.pack_reads(flat, off)
    try:
        fi = ctx.index_build(r)
        ri = ctx.index_build(r, reversed=True)
        fwd = ctx.fm_create(fi.runs, fi.n_suffixes)
        rev = ctx.fm_create(ri.runs, ri.n_suffixes)
        try:
            for attach in (False, True):
                if attach:
                    ctx.fm_attach_reads(fwd, fi.sai, r)
                    ctx.fm_attach_reads(rev, ri.sai, r)
                h = ctx.overlap(fwd, rev, r, 0, len(reads), 36, irreducible=False)
                names = ["r%d" % i for i in range(len(reads))]
                kept, dups = formats.rmdup_records(names, [x.decode() for x in reads], h.substring, h.block_off, h.blocks,
                                                   fi.sai, ri.sai)
                assert sorted(kept) == want[0] and sorted(dups) == want[1], ("attach" if attach else "walk", len(kept), len(dups))
        finally:
            ctx.free_fm(fwd)
            ctx.free_fm(rev)
    finally:
        ctx.free_reads(r)


# ------------------------------------------------------------------------------------------------------- edge cases

@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("case", ["single_read", "two_identical", "min_overlap_above_length", "tiny_reads", "all_same_base",
                                  "read_and_its_revcomp"])
def test_edge_case_inputs_match_reference(ctx, case):
    """Degenerate inputs through the whole chain (index bytes, hit blocks on both post paths, ASQG) against the
    reference binary: one read, duplicates only, a min-overlap no read can reach, reads of 1-3 bases, homopolymers."""
    m = 8
    if case == "single_read":
        reads = [b"ACGTTGCAAGGCTTAACCGG"]
    elif case == "two_identical":
        reads = [b"ACGTTGCAAGGCTTAACCGG"] * 2
    elif case == "min_overlap_above_length":
        reads, m = [b"ACGTTGCAAGGCTTAACCGG", b"TTGCAAGGCTTAACCGGATT", b"GCAAGGCTTAACCGGATTAC"], 64
    elif case == "tiny_reads":
        reads, m = [b"A", b"C", b"AC", b"CA", b"ACG", b"CGT", b"T", b"AC"], 1
    elif case == "all_same_base":
        reads, m = [b"A" * 40, b"A" * 40, b"A" * 35, b"T" * 40], 10
    else:
        reads = [b"ACGTTGCAAGGCTTAACCGGATTACA", b"TGTAATCCGGTTAAGCCTTGCAACGT"]
    _check_index(ctx, reads, case)
    _overlap_vs_ref(ctx, reads, m)


def test_zero_reads_are_rejected(ctx):
    with pytest.raises(api.SgxError):
        ctx.pack_reads(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))


# ------------------------------------------------------------------------------------------- streaming read store

def test_async_read_store_gives_identical_results(ctx, monkeypatch):
    """sgx_reads_pack_async (bases streamed to the device in chunks on a second stream): index, overlap edges (several
    search batches, each waiting only for its own chunks) and correction equal those of the synchronous store."""
    monkeypatch.setenv("SGX_MAX_BATCH", "3000")
    rng = np.random.default_rng(88)
    g = synth.random_genome(40_000, rng)
    reads = synth.sample_reads(g, 30_000, 100, rng, err=0.01)
    flat, off = synth.reads_to_flat(reads)
    n = len(off) - 1
    ranks = synth.default_name_ranks(n)
    r = ctx.pack_reads(flat, off)
    fi, ri = ctx.index_build(r), ctx.index_build(r, reversed=True)
    fwd, rev = ctx.fm_create(fi.runs, fi.n_suffixes), ctx.fm_create(ri.runs, ri.n_suffixes)
    try:
        ctx.fm_attach_reads(fwd, fi.sai, r)
        ctx.fm_attach_reads(rev, ri.sai, r)
        want = ctx.overlap_edges(fwd, rev, r, 0, n, 40, ranks)
        for _ in range(3):
            ra = ctx.pack_reads_async(flat, off)
            ctx.fm_attach_reads(fwd, fi.sai, ra)
            ctx.fm_attach_reads(rev, ri.sai, ra)
            got = ctx.overlap_edges(fwd, rev, ra, 0, n, 40, ranks)
            assert got.stats["search_launches"] > 1
            assert np.array_equal(got.edges, want.edges) and np.array_equal(got.substring, want.substring)
            ctx.free_reads(ra)
        # any other consumer waits for the whole store first
        ra = ctx.pack_reads_async(flat, off)
        fa = ctx.index_build(ra)
        assert np.array_equal(fa.runs, fi.runs) and np.array_equal(fa.sai, fi.sai)
        ctx.free_reads(ra)
        ra = ctx.pack_reads_async(flat, off)
        ctx.free_reads(ra)  # released while chunks may still be in flight
    finally:
        ctx.fm_attach_reads(fwd, fi.sai, r)
        ctx.fm_attach_reads(rev, ri.sai, r)
        ctx.free_fm(fwd)
        ctx.free_fm(rev)
        ctx.free_reads(r)


def test_async_read_store_reports_bad_bases_at_the_consumer(ctx):
    bad = synth.sample_reads(synth.random_genome(5000, np.random.default_rng(1)), 4000, 50, np.random.default_rng(2))
    bad[3210, 7] = ord("N")
    flat, off = synth.reads_to_flat(bad)
    ra = ctx.pack_reads_async(flat, off)
    with pytest.raises(api.SgxError) as e:
        ctx.reads_wait(ra)
    assert e.value.code == -1 and "3210" in str(e.value)
    with pytest.raises(api.SgxError):
        ctx.index_build(ra)  # stays refused
    ctx.free_reads(ra)
    ra = ctx.pack_reads_async(flat, off)
    with pytest.raises(api.SgxError):
        ctx.index_build(ra)
    ctx.free_reads(ra)
