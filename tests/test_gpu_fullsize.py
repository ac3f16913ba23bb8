"""Full-size parity of the named configs against hashes of the reference's own outputs.

tests/golden/fullsize/<config>_<scale>.json were produced ONCE by tests/golden/make_fullsize.py from the compiled
reference (`siga index`, `siga overlap -m M`; config 2 at full size = 661 s + 4 x ~540 s of host time): sha256 of the
four index files, of the run bytes / .sai ids, of the canonically sorted ASQG and a digest of the ED set.  Here the
same seeded read set goes through the C ABI (and, for config 2, through the `siga` CLI file to file) and every hash
must match.  Config 2 at full size runs the production batch loop (4 batches per pass); configs 4 and 5 are pinned at
genome scale 0.01 with both position widths (SGX_FORCE_WIDE = the > 2^32-symbol kernels config 4 needs at full size).
"""
import glob
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import fullsize_common as fc
from siga_b200 import api, formats, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CLI = os.path.join(REPO, "siga_b200", "bin", "siga")
_FIX = sorted(glob.glob(os.path.join(HERE, "golden", "fullsize", "*.json")))
_CASES = [(p, w) for p in _FIX for w in ((False, True) if "cfg2_1" not in os.path.basename(p) else (False,))]


def _sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("path,wide", _CASES, ids=["%s%s" % (os.path.basename(p)[:-5], "_wide" if w else "") for p, w in _CASES])
def test_named_config_matches_reference_hashes(ctx, monkeypatch, path, wide):
    fx = json.load(open(path))
    if wide:
        monkeypatch.setenv("SGX_FORCE_WIDE", "1")
    reads = synth.config_reads(fx["config"], fx["scale"])
    n, L = reads.shape
    assert n == fx["n_reads"] and L == fx["read_len"]
    assert _sha(np.ascontiguousarray(reads).tobytes()) == fx["reads_sha256"], "the generator no longer yields the pinned reads"
    flat, off = synth.reads_to_flat(reads)
    r = ctx.pack_reads(flat, off)
    try:
        fi = ctx.index_build(r)
        ri = ctx.index_build(r, reversed=True)
        ix = fx["index"]
        assert fi.n_suffixes == ix["n_suffixes"] and len(fi.runs) == ix["bwt_n_runs"] and len(ri.runs) == ix["rbwt_n_runs"]
        assert _sha(fi.runs.tobytes()) == ix["bwt_runs_sha256"]
        assert _sha(ri.runs.tobytes()) == ix["rbwt_runs_sha256"]
        assert _sha(fi.sai.astype("<u4").tobytes()) == ix["sai_u32_sha256"]
        assert _sha(ri.sai.astype("<u4").tobytes()) == ix["rsai_u32_sha256"]
        assert _sha(formats.bwt_bytes(fi.n_strings, fi.n_suffixes, fi.runs)) == ix["bwt_sha256"]
        assert _sha(formats.bwt_bytes(ri.n_strings, ri.n_suffixes, ri.runs)) == ix["rbwt_sha256"]
        if n <= 1_000_000:  # the text image of the .sai files (the CLI test below covers config 2)
            assert _sha(formats.sai_bytes(fi.sai)) == ix["sai_sha256"] and _sha(formats.sai_bytes(ri.sai)) == ix["rsai_sha256"]
        fwd = ctx.fm_create(fi.runs, fi.n_suffixes)
        rev = ctx.fm_create(ri.runs, ri.n_suffixes)
        try:
            ctx.fm_attach_reads(fwd, fi.sai, r)
            ctx.fm_attach_reads(rev, ri.sai, r)
            ranks = synth.default_name_ranks(n)
            for m, want in sorted(fx["overlap"].items(), key=lambda kv: int(kv[0])):
                e = ctx.overlap_edges(fwd, rev, r, 0, n, int(m), ranks)
                assert len(e.edges) == want["n_ed"], (m, len(e.edges), want["n_ed"])
                assert int(e.substring.sum()) == want["n_substring"], m
                assert fc.edge_digest(e.edges, np.int64(L)) == want["edges_digest"], m
                if n > 1_000_000:
                    assert e.stats["search_launches"] > 1  # the production multi-batch loop
        finally:
            ctx.free_fm(fwd)
            ctx.free_fm(rev)
    finally:
        ctx.free_reads(r)


@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
@pytest.mark.parametrize("path", [p for p in _FIX if "cfg2_1" in os.path.basename(p) or "cfg5" in os.path.basename(p)],
                         ids=lambda p: os.path.basename(p)[:-5])
def test_cli_files_match_reference_hashes(path):
    """File to file: `siga index` + `siga overlap -m 45` on the FASTA of the config; .bwt/.rbwt/.sai/.rsai byte-identical
    (sha256) and `zcat | LC_ALL=C sort` of the .asqg.gz identical to the reference's."""
    fx = json.load(open(path))
    reads = synth.config_reads(fx["config"], fx["scale"])
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "reads.fa")
        synth.write_fasta(fa, reads)
        for args in (["index", fa], ["overlap", "-m", "45", fa]):
            p = subprocess.run([CLI] + args, cwd=d, capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-2000:]
        for ext in ("bwt", "rbwt", "sai", "rsai"):
            with open(os.path.join(d, "reads." + ext), "rb") as f:
                assert _sha(f.read()) == fx["index"][ext + "_sha256"], ext
        p = subprocess.run("zcat reads.asqg.gz | LC_ALL=C sort -S 4G | sha256sum", shell=True, cwd=d, capture_output=True,
                           text=True, check=True)
        assert p.stdout.split()[0] == fx["overlap"]["45"]["sorted_asqg_sha256"]
