"""The `siga` front-end of this repo against the compiled reference binary: same sub-commands, same files."""
import gzip
import os
import subprocess
import tempfile

import numpy as np
import pytest

import refrun
from siga_b200 import synth

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "siga_b200", "bin", "siga")


def _run(args, cwd):
    r = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def _asqg(path):
    with gzip.open(path, "rt") as f:
        return [l.rstrip("\n") for l in f]


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
@pytest.mark.parametrize("case", ["fasta_fixed", "fastq_comments", "sais_exhaustive_norc"])
def test_cli_outputs_match_reference(case):
    rng = np.random.default_rng(17)
    g = synth.random_genome(4000, rng)
    reads = synth.sample_reads(g, 1500, 60, rng, err=0.01)
    reads[rng.integers(0, 1500, 40)] = reads[rng.integers(0, 1500, 40)]
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        if case == "fastq_comments":
            name = "reads.fq"
            with open(os.path.join(d, name), "w") as f:
                for i, r in enumerate(reads):
                    tag = " CR:i:%d BX:Z:bc%d junk" % (i % 7, i % 3) if i % 2 else ""
                    f.write("@q%d%s\n%s\n+\n%s\n" % (i, tag, r.tobytes().decode(), "I" * 60))
        else:
            name = "reads.fa"
            with open(os.path.join(d, name), "w") as f:
                for i, r in enumerate(reads):
                    s = r.tobytes().decode()
                    f.write(">r%d\n%s\n%s\n" % (i, s[:35], s[35:]))  # multi-line FASTA records are joined
        src = os.path.join(d, name)
        iargs = ["-a", "sais"] if case == "sais_exhaustive_norc" else []
        oargs = ["-x", "--no-opposite-strand"] if case == "sais_exhaustive_norc" else []
        refrun.run_ref(["index"] + iargs + [src], ref)
        refrun.run_ref(["overlap", "-m", "20", "-t", "3"] + oargs + [src], ref)
        _run(["index"] + iargs + [src], ours)
        _run(["overlap", "-m", "20", "-t", "3"] + oargs + [src], ours)
        for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
            a = open(os.path.join(ours, "reads" + ext), "rb").read()
            b = open(os.path.join(ref, "reads" + ext), "rb").read()
            assert a == b, ext
        mine, want = _asqg(os.path.join(ours, "reads.asqg.gz")), _asqg(os.path.join(ref, "reads.asqg.gz"))
        n = len(reads)
        assert mine[:1 + n] == want[:1 + n]          # HT, then VT in input order (incl. SS/CR/BX tags)
        assert sorted(mine[1 + n:]) == sorted(want[1 + n:])  # ED set (the reference emits them in thread order)
        assert all(l.startswith("ED\t") for l in mine[1 + n:])


@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
def test_cli_prefix_and_errors(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_text(">a\nACGTACGTACGTAAAC\n>b\nGTACGTAAACTTTTGG\n")
    _run(["index", "-p", "pfx", "--no-reverse", str(fa)], str(tmp_path))
    assert (tmp_path / "pfx.bwt").exists() and (tmp_path / "pfx.sai").exists() and not (tmp_path / "pfx.rbwt").exists()
    r = subprocess.run([CLI, "index", "-a", "rope", str(fa)], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0  # the reference silently writes nothing here; we refuse
    r = subprocess.run([CLI, "overlap", "-p", "missing", str(fa)], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to load FMIndex" in r.stderr


def _fasta_records(path):
    txt = open(path).read()
    return sorted(">" + r for r in txt.split(">") if r)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
@pytest.mark.parametrize("case", ["fixed", "varlen"])
def test_cli_rmdup_matches_reference(case):
    """`siga rmdup` (src/rmdup.cpp, OverlapBuilder::rmdup): the kept / duplicate FASTA record sets are identical
    (the reference writes them in per-thread hits order, so the comparison is on sorted records)."""
    rng = np.random.default_rng(23)
    g = synth.random_genome(3000, rng)
    if case == "fixed":
        reads = [r.tobytes() for r in synth.sample_reads(g, 700, 40, rng, err=0.0)]
    else:  # variable length: substrings of other reads are removed too
        from test_gpu_parity import _rand_reads
        reads = _rand_reads(rng, 700, 15, 50, 1500)
    for _ in range(90):   # exact duplicates and reverse-complement duplicates
        a, b = rng.integers(0, len(reads), 2)
        reads[a] = reads[b] if rng.random() < 0.5 else synth.revcomp(np.frombuffer(reads[b], dtype=np.uint8)).tobytes()
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        src = os.path.join(d, "reads.fa")
        synth.write_fasta(src, reads)
        refrun.run_ref(["index", src], ref)
        refrun.run_ref(["rmdup", "-t", "3", src], ref)
        _run(["index", src], ours)
        _run(["rmdup", src], ours)
        for f in ("reads.rmdup.fa", "reads.rmdup.dups.fa"):
            mine, want = _fasta_records(os.path.join(ours, f)), _fasta_records(os.path.join(ref, f))
            assert mine == want, (case, f, len(mine), len(want))
        assert len(_fasta_records(os.path.join(ours, "reads.rmdup.dups.fa"))) > 20


def _contigs(path):
    """Contig sequences in canonical orientation (the smaller of a sequence and its reverse complement), sorted."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    out = []
    for rec in open(path, "rb").read().split(b">")[1:]:
        seq = b"".join(rec.split(b"\n")[1:])
        out.append(min(seq, seq.translate(comp)[::-1]))
    return sorted(out)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
@pytest.mark.parametrize("err", [0.0, 0.005])
def test_downstream_assemble_runs_unchanged(err):
    """BASELINE.json north_star: "downstream `siga assemble` runs unchanged".  The reference's own assembler
    (src/assembler.cpp, src/bigraph*.cpp, compiled into oracle/_ref/siga) is run on the .asqg.gz produced by this repo's
    index -> rmdup -> index -> overlap chain and on the reference's; the contig sets must be identical."""
    rng = np.random.default_rng(41)
    g = synth.random_genome(30_000, rng)
    g[20_000:20_400] = g[5_000:5_400]  # repeats longer than a read
    g[25_000:25_800] = g[9_000:9_800]
    reads = synth.sample_reads(g, 15_000, 100, rng, err=err)  # with errors: hundreds of contigs after trimming
    with tempfile.TemporaryDirectory() as d:
        out = {}
        for who in ("ours", "ref"):
            w = os.path.join(d, who)
            os.mkdir(w)
            src = os.path.join(w, "reads.fa")
            synth.write_fasta(src, reads)
            run = (lambda a: _run(a, w)) if who == "ours" else (lambda a: refrun.run_ref(a, w))
            run(["index", src])
            run(["rmdup"] + (["-t", "4"] if who == "ref" else []) + [src])
            dedup = os.path.join(w, "reads.rmdup.fa")
            run(["index", dedup])
            run(["overlap", "-m", "45", "-t", "4", dedup])
            refrun.run_ref(["assemble", "-m", "45", "-p", "asm", os.path.join(w, "reads.rmdup.asqg.gz")], w)
            out[who] = _contigs(os.path.join(w, "asm-contigs.fa"))
        assert sum(len(c) for c in out["ref"]) > 25_000 and (err == 0.0 or len(out["ref"]) > 100)
        assert out["ours"] == out["ref"]


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_cli_multi_gpu_outputs_equal_single_gpu():
    """`siga index --gpus 2` (NCCL, rows sharded by k-mer bucket), `siga overlap --gpus 2` and `siga correct --gpus 2`
    (reads sharded, part files concatenated) write the same index bytes, the same HT/VT lines, the same ED set and the
    same corrected reads as the one-GPU commands."""
    rng = np.random.default_rng(5)
    g = synth.random_genome(20_000, rng)
    reads = synth.sample_reads(g, 9_001, 80, rng, err=0.01)
    with tempfile.TemporaryDirectory() as d:
        one, two = os.path.join(d, "one"), os.path.join(d, "two")
        os.mkdir(one)
        os.mkdir(two)
        src = os.path.join(d, "reads.fa")
        synth.write_fasta(src, reads)
        _run(["index", src], one)
        _run(["overlap", "-m", "30", src], one)
        _run(["index", "--gpus", "2", src], two)
        _run(["overlap", "-m", "30", "--gpus", "2", src], two)
        for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
            assert open(os.path.join(one, "reads" + ext), "rb").read() == open(os.path.join(two, "reads" + ext), "rb").read(), ext
        # one strand sharded over both GPUs (rows by k-mer bucket, NCCL): forced below the size threshold
        # (the GPUs write one .bwt together, each its own segment: a longer stale file of that name must not survive)
        for ext in (".bwt", ".rbwt"):
            with open(os.path.join(two, "sharded" + ext), "wb") as f:
                f.write(b"\xff" * (os.path.getsize(os.path.join(one, "reads" + ext)) + 100000))
        for flag, exts in (("--no-reverse", (".bwt", ".sai")), ("--no-forward", (".rbwt", ".rsai"))):
            r = subprocess.run([CLI, "index", "--gpus", "2", flag, "-p", "sharded", src], cwd=two, capture_output=True, text=True,
                               env=dict(os.environ, SIGA_SHARD_MIN_ROWS="0"))
            assert r.returncode == 0, r.stderr[-2000:]
            for ext in exts:
                assert open(os.path.join(one, "reads" + ext), "rb").read() == open(os.path.join(two, "sharded" + ext), "rb").read(), ext
        a, b = _asqg(os.path.join(one, "reads.asqg.gz")), _asqg(os.path.join(two, "reads.asqg.gz"))
        n = len(reads)
        assert a[:1 + n] == b[:1 + n]
        assert sorted(a[1 + n:]) == sorted(b[1 + n:]) and len(a) > 1 + n
        # `siga correct --gpus 2`: reads sharded by contiguous range, part files concatenated in rank order
        _run(["correct", "-k", "25", src], one)
        _run(["correct", "-k", "25", "--gpus", "2", src], two)
        ec = open(os.path.join(one, "reads.ec.fa"), "rb").read()
        assert ec == open(os.path.join(two, "reads.ec.fa"), "rb").read() and ec.count(b">") > 1000
        assert not [f for f in os.listdir(two) if ".part" in f]


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
def test_cli_match_stdout_equals_reference():
    """`siga match` on the device (sgx_fm_count, and sgx_fm_occ for queries with bytes other than A/C/G/T): stdout byte
    for byte the reference's (src/match.cpp:22-76), whole reads and -l windows, both strands and forward only."""
    rng = np.random.default_rng(77)
    g = synth.random_genome(3000, rng)
    idx_reads = synth.sample_reads(g, 1200, 60, rng, err=0.005)
    q = [r.tobytes().decode() for r in synth.sample_reads(g, 300, 60, rng, err=0.01)]
    q += [r.tobytes().decode()[:int(l)] for r, l in zip(synth.sample_reads(g, 100, 60, rng), rng.integers(5, 61, 100))]
    for i, ch in ((3, "N"), (10, "n"), (11, "R")):
        k = int(rng.integers(0, len(q[i])))
        q[i] = q[i][:k] + ch + q[i][k + 1:]
    q[20] = q[20][:-1] + "N"
    q[21] = "N" + q[21][1:]
    with tempfile.TemporaryDirectory() as d:
        ref_fa = os.path.join(d, "ref.fa")
        synth.write_fasta(ref_fa, idx_reads)
        refrun.run_ref(["index", "--no-reverse", ref_fa], d)
        a = os.path.join(d, "a.fa")
        synth.write_fasta(a, q, names=["a%d" % i for i in range(len(q))])
        for args in (["-p", "ref", a], ["-p", "ref", "--no-opposite-strand", a], ["-p", "ref", "-l", "20", a]):
            want = subprocess.run([refrun.REF_BIN, "match"] + args, cwd=d, capture_output=True, text=True,
                                  env=dict(os.environ, ORACLE_LOG_LEVEL="5"))
            assert want.returncode == 0 and want.stdout.count("\n") >= len(q)
            assert _run(["match"] + args, d).stdout == want.stdout, args


@pytest.mark.skipif(not os.path.isfile(CLI), reason="siga_b200/bin/siga not built")
def test_cli_edge_order_is_reproducible():
    """The ED section is ordered by query read (stable device sort of the edge records): two runs write the same bytes."""
    rng = np.random.default_rng(12)
    g = synth.random_genome(30_000, rng)
    reads = synth.sample_reads(g, 20_000, 100, rng, err=0.01)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "reads.fa")
        synth.write_fasta(src, reads)
        _run(["index", src], d)
        outs = []
        for k in range(2):
            _run(["overlap", "-m", "40", "-p", os.path.join(d, "reads"), src], d)
            with gzip.open(os.path.join(d, "reads.asqg.gz"), "rb") as f:
                outs.append(f.read())
        assert outs[0] == outs[1] and outs[0].count(b"\nED\t") > 10_000
        ed = [int(l.split(b"\t")[1].split(b" ")[0][1:]) for l in outs[0].split(b"\n") if l.startswith(b"ED")]
        assert ed == sorted(ed)
