"""CPU tests of the `siga` front-end's host logic: the UNMODIFIED front-end sources (siga_b200/csrc/host) linked against
a host stand-in of the C ABI (tests/hostemu/sgx_stub.cpp -> build/siga_hostemu, test infrastructure only) and run against
the compiled reference.  Covers what no kernel test sees: option parsing, FASTQ qualities and comments, fix application,
output formatting and compression, and the N > 1 path of `siga correct --gpus N` (fork, per-rank read ranges, part
files concatenated in rank order) on CPU processes."""
import gzip
import os
import subprocess
import tempfile

import numpy as np
import pytest

import correct_common as cc
import refrun
from siga_b200 import synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "build", "siga_hostemu")

pytestmark = [pytest.mark.skipif(not os.path.isfile(CLI), reason="build/siga_hostemu not built (make hostcli)"),
              pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")]


def _cli(args, cwd, ok=True, env=None):
    r = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, env=env)
    assert (r.returncode == 0) == ok, r.stderr[-2000:]
    return r


@pytest.mark.parametrize("fastq", [False, True], ids=["fasta", "fastq"])
def test_correct_front_end_single_and_sharded_equal_reference_file(fastq):
    reads, quals = cc.random_case(21 + int(fastq), n_reads=2001, read_len=75, genome=4000, err=0.015, fastq=fastq, both=True)
    names = ["q%d%s" % (i, " CR:i:%d note" % (i % 5) if i % 3 == 0 else "") for i in range(len(reads))]
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        src = os.path.join(d, "reads.fq" if fastq else "reads.fa")
        cc.write_reads(src, reads, quals, names)
        refrun.run_ref(["index", "--no-reverse", src], ref)
        refrun.run_ref(["correct", "-t", "3", "-k", "23", "-x", "2", "-i", "4", "-O", "2", src], ref)
        want = open(os.path.join(ref, "reads.ec.fa"), "rb").read()
        assert 0 < want.count(b"\n") // (4 if fastq else 2) < len(reads)
        opts = ["-k", "23", "-x", "2", "-i", "4", "-O", "2", "-p", os.path.join(ref, "reads")]
        par = dict(os.environ, SIGA_PARSE_MIN_PARALLEL="1000")  # the multi-threaded in-place reader instead of the serial one
        for extra, out, env in (([], "reads.ec.fa", None), (["--gpus", "2"], "reads.ec.fa", par),
                                (["--gpus", "3", "-o", "x.fa.gz"], "x.fa.gz", None),
                                (["--kmer-size=23", "--outfile", "y.fa"], "y.fa", par)):
            _cli(["correct"] + opts + extra + [src], ours, env=env)
            path = os.path.join(ours, out)
            got = gzip.open(path, "rb").read() if out.endswith(".gz") else open(path, "rb").read()
            assert got == want, extra
            os.remove(path)
            assert not [f for f in os.listdir(ours) if ".part" in f]
        # errors: unknown algorithm, missing index, bzip2 output, a base other than A/C/G/T
        _cli(["correct", "-a", "overlap"] + opts + [src], ours, ok=False)
        _cli(["correct", "-p", "missing", src], ours, ok=False)
        _cli(["correct", "-o", "z.fa.bz2"] + opts + [src], ours, ok=False)
        bad = os.path.join(d, "bad.fa")
        open(bad, "w").write(">a\nACGTNACGTACGTACGTACGTACGTACGT\n")
        _cli(["correct"] + opts + [bad], ours, ok=False)


def _fasta_records(path):
    recs = open(path, "rb").read().split(b">")[1:]
    return sorted(recs)


def test_rmdup_front_end_equals_reference_records():
    """`siga rmdup` host logic (NumDuplicates, the kept / removed split by name order) on hit blocks from the host lane."""
    rng = np.random.default_rng(23)
    g = synth.random_genome(1500, rng)
    reads = [r.tobytes() for r in synth.sample_reads(g, 350, 40, rng, err=0.0)]
    for _ in range(50):
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
        for ext in (".bwt", ".rbwt", ".sai", ".rsai"):  # the front-end reads the reference's index files
            os.symlink(os.path.join(ref, "reads" + ext), os.path.join(ours, "reads" + ext))
        _cli(["rmdup", src], ours)
        for f in ("reads.rmdup.fa", "reads.rmdup.dups.fa"):
            assert _fasta_records(os.path.join(ours, f)) == _fasta_records(os.path.join(ref, f)), f
        assert len(_fasta_records(os.path.join(ours, "reads.rmdup.dups.fa"))) > 10


def _asqg(path):
    with gzip.open(path, "rt") as f:
        return [l.rstrip("\n") for l in f]


@pytest.mark.parametrize("case", ["fasta", "fastq_comments_exhaustive_norc", "fasta_long_names_variable_length"])
def test_overlap_front_end_single_and_sharded_equal_reference_asqg(case):
    """`siga overlap` host logic: index / read loading, name ranks, the ASQG writer (HT, VT in input order with SS / CR /
    BX tags, ED) and the --gpus N path (fork, read ranges, .vt / .ed part files concatenated)."""
    rng = np.random.default_rng(17)
    g = synth.random_genome(2500, rng)
    reads = synth.sample_reads(g, 901, 50, rng, err=0.01)
    reads[rng.integers(0, 901, 30)] = reads[rng.integers(0, 901, 30)]
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        if case == "fasta":
            src = os.path.join(d, "reads.fa")
            synth.write_fasta(src, reads)
            oargs = []
        elif case == "fasta_long_names_variable_length":
            # ED records of very different sizes (names of 2 .. 70 characters, coordinates of reads of 30 .. 50 bases):
            # the writer sizes its record buffer from the longest name and the longest read
            src = os.path.join(d, "reads.fa")
            reads = [r[:int(rng.integers(30, 51))] for r in reads]
            with open(src, "w") as f:
                for i, r in enumerate(reads):
                    f.write(">%s%d\n%s\n" % ("n" * (1 + (i * 7) % 64), i, r.tobytes().decode()))
            oargs = []
        else:
            src = os.path.join(d, "reads.fq")
            with open(src, "w") as f:
                for i, r in enumerate(reads):
                    tag = " CR:i:%d BX:Z:bc%d junk" % (i % 7, i % 3) if i % 2 else ""
                    f.write("@q%d%s\n%s\n+\n%s\n" % (i, tag, r.tobytes().decode(), "I" * 50))
            oargs = ["-x", "--no-opposite-strand"]
        refrun.run_ref(["index", src], ref)
        refrun.run_ref(["overlap", "-m", "18", "-t", "3"] + oargs + [src], ref)
        want = _asqg(os.path.join(ref, "reads.asqg.gz"))
        n = len(reads)
        assert len(want) > 1 + n + 100
        for extra in ([], ["--gpus", "2"], ["--gpus", "3", "--gzip-level", "6"]):
            os.remove(os.path.join(ref, "reads.asqg.gz"))  # -p PREFIX names the output too
            _cli(["overlap", "-m", "18", "-p", os.path.join(ref, "reads")] + oargs + extra + [src], ours)
            got = _asqg(os.path.join(ref, "reads.asqg.gz"))
            assert got[:1 + n] == want[:1 + n], extra
            assert sorted(got[1 + n:]) == sorted(want[1 + n:]), extra
            assert not [f for f in os.listdir(ref) if ".part" in f]


@pytest.mark.parametrize("fail,hang", [(1, 0), (0, 1), (2, 1)])
def test_multi_gpu_worker_failure_ends_the_run_cleanly(fail, hang):
    """`--gpus 3`: one rank fails while another never returns (a peer blocked in a collective).  The command must exit
    with an error within seconds, leave no part files and no worker process behind -- whichever rank fails, the calling
    process (rank 0) included."""
    import time
    rng = np.random.default_rng(3)
    g = synth.random_genome(1500, rng)
    reads = synth.sample_reads(g, 300, 40, rng)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "reads.fa")
        synth.write_fasta(src, reads)
        _cli(["index", src], d)
        env = dict(os.environ, SGX_STUB_FAIL_DEVICE=str(fail), SGX_STUB_HANG_DEVICE=str(hang))
        for cmd in (["overlap", "-m", "15", "--gpus", "3", src], ["correct", "-k", "11", "--gpus", "3", src]):
            t0 = time.time()
            p = subprocess.run([CLI] + cmd, cwd=d, env=env, capture_output=True, text=True, timeout=60)
            assert p.returncode != 0 and time.time() - t0 < 20, (cmd, p.returncode, p.stderr[-500:])
            assert "ERROR" in p.stderr
            assert not [f for f in os.listdir(d) if ".part" in f], os.listdir(d)
            time.sleep(0.3)
            left = subprocess.run(["pgrep", "-f", src], capture_output=True, text=True).stdout.split()
            assert not left, "worker processes left behind: %s" % left


def test_match_front_end_equals_reference_stdout():
    """`siga match` (src/match.cpp:22-76): occurrence counts of whole reads or of their first / last -l bases, both
    strands or forward only, several input files, FASTA and FASTQ -- stdout byte for byte."""
    rng = np.random.default_rng(77)
    g = synth.random_genome(3000, rng)
    idx_reads = synth.sample_reads(g, 1200, 60, rng, err=0.005)
    q1 = [r.tobytes().decode() for r in synth.sample_reads(g, 300, 60, rng, err=0.01)]
    q2 = [r.tobytes().decode()[:int(l)] for r, l in zip(synth.sample_reads(g, 200, 60, rng, err=0.0), rng.integers(5, 61, 200))]
    # un-preprocessed queries: the reference maps any other byte to the sentinel rank and reports a count for every read
    for i, ch in ((3, "N"), (10, "n"), (11, "R"), (50, "N")):
        k = int(rng.integers(0, len(q1[i])))
        q1[i] = q1[i][:k] + ch + q1[i][k + 1:]
    q1[20] = q1[20][:-1] + "N"
    q1[21] = "N" + q1[21][1:]
    with tempfile.TemporaryDirectory() as d:
        ref_fa = os.path.join(d, "ref.fa")
        synth.write_fasta(ref_fa, idx_reads)
        refrun.run_ref(["index", "--no-reverse", ref_fa], d)
        a, b = os.path.join(d, "a.fa"), os.path.join(d, "b.fq")
        cc.write_reads(a, q1, None, ["a%d extra" % i for i in range(len(q1))])
        cc.write_reads(b, q2, ["I" * len(s) for s in q2])
        for args in (["-p", "ref", a], ["-p", "ref", "--no-opposite-strand", a, b], ["-p", "ref", "-l", "20", a, b],
                     ["--prefix=ref", "--max-length", "7", "--no-opposite-strand", b]):
            want = subprocess.run([refrun.REF_BIN, "match"] + args, cwd=d, capture_output=True, text=True,
                                  env=dict(os.environ, ORACLE_LOG_LEVEL="5"))
            assert want.returncode == 0 and want.stdout.count("\n") >= len(q2)
            got = _cli(["match"] + args, d)
            assert got.stdout == want.stdout, args
        _cli(["match", "-p", "missing", a], d, ok=False)


def test_reference_assembler_reads_our_asqg_and_builds_the_same_contigs():
    """BASELINE.json: "downstream `siga assemble` runs unchanged" -- the reference's assembler (compiled into the oracle
    binary) on the .asqg.gz this front-end writes (multi-member gzip, literal-only members for the VT records) and on the
    reference's own file: identical contig sets."""
    rng = np.random.default_rng(41)
    g = synth.random_genome(6000, rng)
    reads = synth.sample_reads(g, 2500, 80, rng, err=0.0)

    def contigs(path):
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        out = []
        for rec in open(path, "rb").read().split(b">")[1:]:
            s = b"".join(rec.split(b"\n")[1:])
            out.append(min(s, s.translate(comp)[::-1]))
        return sorted(out)

    with tempfile.TemporaryDirectory() as d:
        ref, ours = os.path.join(d, "ref"), os.path.join(d, "ours")
        os.mkdir(ref)
        os.mkdir(ours)
        src = os.path.join(ref, "reads.fa")
        synth.write_fasta(src, reads)
        refrun.run_ref(["index", src], ref)
        refrun.run_ref(["rmdup", "-t", "3", src], ref)
        dedup = os.path.join(ref, "reads.rmdup.fa")
        refrun.run_ref(["index", dedup], ref)
        refrun.run_ref(["overlap", "-m", "40", "-t", "3", dedup], ref)
        refrun.run_ref(["assemble", "-m", "40", "-p", "asm", os.path.join(ref, "reads.rmdup.asqg.gz")], ref)
        for ext in (".fa", ".bwt", ".rbwt", ".sai", ".rsai"):
            os.symlink(os.path.join(ref, "reads.rmdup" + ext), os.path.join(ours, "reads.rmdup" + ext))
        _cli(["overlap", "-m", "40", os.path.join(ours, "reads.rmdup.fa")], ours)
        refrun.run_ref(["assemble", "-m", "40", "-p", "asm", os.path.join(ours, "reads.rmdup.asqg.gz")], ours)
        a, b = contigs(os.path.join(ours, "asm-contigs.fa")), contigs(os.path.join(ref, "asm-contigs.fa"))
        assert a == b and sum(len(c) for c in a) > 4000


IRREGULAR_FASTA = {
    # FASTAReader::read (src/kseq.cpp:191-228) keys its records on the text behind '>' being non-empty
    "bare_header_first": ">\nACGTACGTAC\n>b\nGGGTTTAAAC\n",                     # the orphan lines join record b
    "bare_header_mid": ">a\nACGTACGTAC\n>\nGGGTTTAAAC\n>c\nTTTTACGTAA\n>d\nCCCCCAAAAA\n",
    "bare_header_last": ">a\nACGTACGTAC\n>b\nGGGTTTAAAC\n>\nTTTTACGTAA\n",    # ... or are dropped at the end
    "bare_header_without_lines": ">a\nACGTACGTAC\n>\n>c\nTTTTACGTAA\n",
    "named_record_without_lines": ">a\nACGTACGTAC\n>e\n>c\nTTTTACGTAA\n",       # the reference stops reading here
    "name_is_all_comment": "> only a comment\nACGTACGTAC\n>b\nGGGTTTAAAC\n",
    "crlf_blank_lines_no_final_newline": ">a x\r\nACGTAC\r\n\r\nGTAC\r\n>b\r\nGGGTTTAAAC\r\n>c\r\nTTTTACGTAA",
}


@pytest.mark.parametrize("case", sorted(IRREGULAR_FASTA))
def test_irregular_fasta_is_read_like_the_reference_reads_it(case):
    """Index files and ASQG of inputs the reference's reader treats in its own way: same records, same files."""
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        src = os.path.join(d, "reads.fa")
        with open(src, "w", newline="") as f:
            f.write(IRREGULAR_FASTA[case])
        for run, cwd in ((lambda a: refrun.run_ref(a, ref), ref), (lambda a: _cli(a, ours), ours)):
            run(["index", src])
            run(["overlap", "-m", "3", src])
        for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
            assert open(os.path.join(ours, "reads" + ext), "rb").read() == open(os.path.join(ref, "reads" + ext), "rb").read(), ext
        assert sorted(_asqg(os.path.join(ours, "reads.asqg.gz"))) == sorted(_asqg(os.path.join(ref, "reads.asqg.gz")))


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def _nasty_genome(rnd):
    kind, n = rnd.random(), rnd.randint(30, 400)
    if kind < 0.3:
        return "".join(rnd.choice("ACGT") for _ in range(n))
    if kind < 0.55:  # tandem repeats with a few substitutions
        unit = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 7)))
        g = list((unit * (n // len(unit) + 1))[:n])
        for _ in range(rnd.randint(0, 4)):
            g[rnd.randrange(n)] = rnd.choice("ACGT")
        return "".join(g)
    if kind < 0.8:  # low complexity
        return "".join(rnd.choice("AC") if rnd.random() < 0.9 else rnd.choice("GT") for _ in range(n))
    seg = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(10, 40)))  # long repeats and a palindrome
    return seg + "".join(rnd.choice("ACGT") for _ in range(10)) + seg + _rc(seg) + seg[:len(seg) // 2]


def _nasty_reads(rnd, g):
    fixed, l0, out = rnd.random() < 0.4, rnd.randint(6, min(40, len(g))), []
    for _ in range(rnd.randint(2, 120)):
        ln = l0 if fixed else rnd.randint(4, min(45, len(g)))
        p = rnd.randrange(0, len(g) - ln + 1)
        s = g[p:p + ln]
        if rnd.random() < 0.4:
            s = _rc(s)
        if out and rnd.random() < 0.1:
            s = rnd.choice(out)          # duplicate
        if out and rnd.random() < 0.05:
            s = _rc(rnd.choice(out))     # reverse-complement duplicate
        if rnd.random() < 0.1:
            k = rnd.randrange(len(s))
            s = s[:k] + rnd.choice("ACGT") + s[k + 1:]
        out.append(s)
    return out


def test_overlap_logic_on_repetitive_read_sets_equals_the_reference():
    """Seeded fuzz of the whole front-end on read sets made to hurt: tandem repeats, two-letter sequence, long repeats with
    a palindrome, duplicates and reverse-complement duplicates, reads of 4 .. 45 bases, min-overlap 3 .. 16, -x,
    --no-opposite-strand, both ordering models.  The host stand-in runs the SAME overlap_logic.h (find / filter / resolve /
    extract) the general path of the CUDA post kernel compiles: index files byte-identical, ASQG set-identical."""
    import random
    rnd = random.Random(23)
    with tempfile.TemporaryDirectory() as d:
        for it in range(60):
            reads = _nasty_reads(rnd, _nasty_genome(rnd))
            src = os.path.join(d, "in%d.fa" % it)
            pre = rnd.choice(["r", "x9", "Z", "a_"])
            with open(src, "w") as f:
                f.write("".join(">%s%d\n%s\n" % (pre, i, s) for i, s in enumerate(reads)))
            opts = ["-m", str(rnd.randint(3, 16))] + (["-x"] if rnd.random() < 0.3 else []) + \
                (["--no-opposite-strand"] if rnd.random() < 0.25 else [])
            algo = rnd.choice(["sais", "sais2"])
            ours, ref = os.path.join(d, "ours%d" % it), os.path.join(d, "ref%d" % it)
            os.mkdir(ours)
            os.mkdir(ref)
            for tool, cwd in ((refrun.REF_BIN, ref), (CLI, ours)):
                for args in (["index", "-a", algo, "-p", "x", src], ["overlap", "-p", "x", "-t", "2"] + opts + [src]):
                    r = subprocess.run([tool] + args, cwd=cwd, capture_output=True, text=True)
                    assert r.returncode == 0, (it, args, r.stderr[-500:])
            for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
                assert open(os.path.join(ours, "x" + ext), "rb").read() == open(os.path.join(ref, "x" + ext), "rb").read(), (it, ext)
            assert sorted(_asqg(os.path.join(ours, "x.asqg.gz"))) == sorted(_asqg(os.path.join(ref, "x.asqg.gz"))), (it, opts, algo)


def test_correct_and_rmdup_on_random_read_sets_equal_the_reference():
    """Seeded fuzz of `siga correct` (k 3 .. 31, threshold 1 .. 5, 0 .. 12 rounds, count offset 0 .. 3; FASTA and FASTQ with
    low qualities; error rates up to 8 %; random, tandem-repeat and two-letter genomes; fixed and variable read length)
    and `siga rmdup` on the same sets.  The host stand-in runs the SAME correct_logic.h the CUDA kernel compiles:
    corrected files byte-identical, kept / duplicate record sets identical."""
    import random
    rnd = random.Random(31)
    with tempfile.TemporaryDirectory() as d:
        for it in range(40):
            n_g, kind = rnd.randint(60, 500), rnd.random()
            if kind < 0.6:
                g = "".join(rnd.choice("ACGT") for _ in range(n_g))
            elif kind < 0.8:
                u = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(2, 9)))
                g = (u * (n_g // len(u) + 1))[:n_g]
            else:
                g = "".join(rnd.choice("AC") if rnd.random() < 0.85 else rnd.choice("GT") for _ in range(n_g))
            fixed, l0, err = rnd.random() < 0.5, rnd.randint(12, min(60, n_g)), rnd.choice([0.0, 0.01, 0.03, 0.08])
            reads = []
            for _ in range(max(2, n_g * rnd.randint(3, 25) // l0)):
                ln = l0 if fixed else rnd.randint(6, min(60, n_g))
                p = rnd.randrange(0, n_g - ln + 1)
                s = g[p:p + ln] if rnd.random() < 0.5 else _rc(g[p:p + ln])
                reads.append("".join(c if rnd.random() >= err else rnd.choice("ACGT") for c in s))
            fq = rnd.random() < 0.5
            src = os.path.join(d, "in%d.%s" % (it, "fq" if fq else "fa"))
            with open(src, "w") as f:
                for i, s in enumerate(reads):
                    f.write("@r%d%s\n%s\n+\n%s\n" % (i, rnd.choice(["", " c"]), s, "".join(rnd.choice("#+5?I") for _ in s)) if fq
                            else ">r%d\n%s\n" % (i, s))
            opts = ["-k", str(rnd.randint(3, min(31, l0 if fixed else 12))), "-x", str(rnd.randint(1, 5)),
                    "-i", str(rnd.randint(0, 12)), "-O", str(rnd.randint(0, 3))]
            ours, ref = os.path.join(d, "ours%d" % it), os.path.join(d, "ref%d" % it)
            os.mkdir(ours)
            os.mkdir(ref)
            for tool, cwd in ((refrun.REF_BIN, ref), (CLI, ours)):
                for args in (["index", "-p", "x", src], ["correct", "-p", "x", "-o", "out.fa", "-t", "2"] + opts + [src],
                             ["rmdup", "-p", "x", src]):
                    r = subprocess.run([tool] + args, cwd=cwd, capture_output=True, text=True)
                    assert r.returncode == 0, (it, args, r.stderr[-500:])
            assert open(os.path.join(ours, "out.fa")).read() == open(os.path.join(ref, "out.fa")).read(), (it, opts)
            for f in ("x.rmdup.fa", "x.rmdup.dups.fa"):
                assert _fasta_records(os.path.join(ours, f)) == _fasta_records(os.path.join(ref, f)), (it, f)


@pytest.mark.parametrize("lengths,works", [(range(10, 41), True), (range(10, 47), False)], ids=["31_groups", "37_groups"])
def test_known_limit_of_the_duplicate_map_in_resolve(lengths, works):
    """KNOWN LIMITATION, pinned so that it stays loud: SubMaximalBlockFilter::resolve hands out the
    forward rows of duplicated reads through a std::map (src/overlap_builder.cpp:1018-1059); overlap_logic.h:ov_resolve
    keeps 32 entries per call.  Homopolymer reads of N different lengths, each one twice, put N duplicate groups into one
    resolve: up to 32 the ASQG equals the reference's, beyond that the call fails with SGX_EOVERFLOW -- it never writes
    a different graph.  (`siga rmdup`, which the pipeline runs before `overlap`, removes such reads.)"""
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        src = os.path.join(d, "reads.fa")
        with open(src, "w") as f:
            f.write("".join(">r%d\n%s\n" % (i, "A" * ln) for i, ln in enumerate(l for l in lengths for _ in (0, 1))))
        refrun.run_ref(["index", src], ref)
        refrun.run_ref(["overlap", "-m", "8", src], ref)
        _cli(["index", src], ours)
        r = _cli(["overlap", "-m", "8", src], ours, ok=works)
        if works:
            assert sorted(_asqg(os.path.join(ours, "reads.asqg.gz"))) == sorted(_asqg(os.path.join(ref, "reads.asqg.gz")))
        else:
            assert "sgx_overlap_edges failed (-5)" in r.stderr and not os.path.exists(os.path.join(ours, "reads.asqg.gz"))


def _random_fasta(rnd):
    seq = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))  # noqa: E731
    out = []
    for i in range(rnd.randint(1, 12)):
        k, nl = rnd.random(), rnd.choice(["\n", "\n", "\r\n"])
        if k < 0.55:
            out.append(">r%d%s%s" % (i, rnd.choice(["", " c", " CR:i:3 x", "\tt"]), nl))
        elif k < 0.75:
            out.append(rnd.choice([">", "> ", "> lead"]) + nl)
        for _ in range(rnd.randint(0, 3)):
            r = rnd.random()
            out.append(rnd.choice(["", " ", "  "]) + seq(rnd.randint(1, 25)) + rnd.choice(["", " ", "\t"]) + nl if r < 0.7 else
                       nl if r < 0.85 else "  " + nl)
    s = "".join(out)
    if rnd.random() < 0.3:
        s = s.rstrip("\r\n")
    return s if s.startswith(">") else ">first\n" + s


def _random_fastq(rnd):
    out = []
    for i in range(rnd.randint(1, 10)):
        s = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 25)))
        name, nl = "q%d%s" % (i, rnd.choice(["", " c1", " BX:Z:a"])), rnd.choice(["\n", "\n", "\r\n"])
        q, k = "".join(rnd.choice("I@+>5") for _ in s), rnd.random()
        plus = "-" if k < 0.07 else "+other" if 0.12 < k < 0.16 else "+" if rnd.random() < 0.6 else "+" + name
        rec = ["@" + name, s, plus, q[:-1] if 0.07 < k < 0.12 else q]
        if rnd.random() < 0.1:
            rec.insert(rnd.randint(0, 4), "")
        if rnd.random() < 0.05:
            rec = rec[:rnd.randint(1, 3)]
        out.append(nl.join(rec) + nl)
    return "".join(out)


def test_random_irregular_inputs_parse_like_the_reference():
    """Seeded fuzz of the FASTA / FASTQ reader: bare and comment-only headers, records without lines, padded lines, blank
    lines, CRLF, missing final newline; FASTQ with a wrong separator, a short quality line, a repeated or foreign name
    behind '+', cut records.  Wherever the reference builds an index, ours is byte-identical; where the reference gives
    up (it aborts on an empty read set), ours exits with an error."""
    import random
    rnd = random.Random(7)
    compared = 0
    with tempfile.TemporaryDirectory() as d:
        for it in range(80):
            fq = rnd.random() < 0.4
            src = os.path.join(d, "in%d.%s" % (it, "fq" if fq else "fa"))
            with open(src, "w", newline="") as f:
                f.write(_random_fastq(rnd) if fq else _random_fasta(rnd))
            ours, ref = os.path.join(d, "ours%d" % it), os.path.join(d, "ref%d" % it)
            os.mkdir(ours)
            os.mkdir(ref)
            r = subprocess.run([refrun.REF_BIN, "index", "-p", "x", src], cwd=ref, capture_output=True)
            o = subprocess.run([CLI, "index", "-p", "x", src], cwd=ours, capture_output=True)
            if r.returncode != 0 or not os.path.isfile(os.path.join(ref, "x.sai")):
                assert o.returncode != 0, open(src).read()
                continue
            assert o.returncode == 0, open(src).read()
            for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
                assert open(os.path.join(ours, "x" + ext), "rb").read() == open(os.path.join(ref, "x" + ext), "rb").read(), open(src).read()
            compared += 1
    assert compared >= 40


@pytest.mark.parametrize("algo", ["sais2", "sais"])
def test_index_front_end_equals_reference_files(algo):
    """`siga index` host logic (reads in, .bwt/.sai/.rbwt/.rsai writers, -a / -p / --no-forward / --no-reverse) with the
    brute-force suffix sort of the host stand-in: both ordering models, variable lengths, duplicates, multi-line FASTA."""
    rng = np.random.default_rng(5)
    g = synth.random_genome(1200, rng)
    reads = [r.tobytes().decode()[:int(l)] for r, l in zip(synth.sample_reads(g, 400, 50, rng, err=0.01), rng.integers(8, 51, 400))]
    for _ in range(40):
        a, b = rng.integers(0, len(reads), 2)
        reads[a] = reads[b]
    with tempfile.TemporaryDirectory() as d:
        ours, ref = os.path.join(d, "ours"), os.path.join(d, "ref")
        os.mkdir(ours)
        os.mkdir(ref)
        src = os.path.join(d, "reads.fa")
        with open(src, "w") as f:
            for i, s in enumerate(reads):
                f.write(">r%d some comment\n%s\n%s\n" % (i, s[:20], s[20:]) if len(s) > 20 else ">r%d\n%s\n" % (i, s))
        refrun.run_ref(["index", "-a", algo, src], ref)
        _cli(["index", "-a", algo, src], ours)
        for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
            assert open(os.path.join(ours, "reads" + ext), "rb").read() == open(os.path.join(ref, "reads" + ext), "rb").read(), ext
        _cli(["index", "-a", algo, "-p", "fwd", "--no-reverse", src], ours)
        assert sorted(f for f in os.listdir(ours) if f.startswith("fwd")) == ["fwd.bwt", "fwd.sai"]
        assert open(os.path.join(ours, "fwd.bwt"), "rb").read() == open(os.path.join(ref, "reads.bwt"), "rb").read()
        _cli(["index", "-a", algo, "-p", "rev", "--no-forward", src], ours)
        assert sorted(f for f in os.listdir(ours) if f.startswith("rev")) == ["rev.rbwt", "rev.rsai"]
        assert open(os.path.join(ours, "rev.rsai"), "rb").read() == open(os.path.join(ref, "reads.rsai"), "rb").read()
        _cli(["index", "-a", "rope", src], ours, ok=False)
        # --gpus 2: one strand per worker process (strand groups of one GPU, no collective); each group's first rank
        # writes its two files
        _cli(["index", "-a", algo, "-p", "two", "--gpus", "2", src], ours)
        for ext in (".bwt", ".rbwt", ".sai", ".rsai"):
            assert open(os.path.join(ours, "two" + ext), "rb").read() == open(os.path.join(ref, "reads" + ext), "rb").read(), ext
        # a strand sharded over the workers of its group (forced below the size threshold): every worker writes the run
        # bytes of its own share into the one .bwt (pwrite at its offset, header and final length by the group's first
        # worker) -- three workers on one strand, then two groups of two; a longer stale file must not survive
        for prefix, extra, exts in (("s3", ["--gpus", "3", "--no-reverse"], (".bwt", ".sai")),
                                    ("s4", ["--gpus", "4"], (".bwt", ".rbwt", ".sai", ".rsai"))):
            for ext in exts:
                if "bwt" in ext:
                    with open(os.path.join(ours, prefix + ext), "wb") as f:
                        f.write(b"\xff" * (os.path.getsize(os.path.join(ref, "reads" + ext)) + 4096))
            _cli(["index", "-a", algo, "-p", prefix] + extra + [src], ours, env=dict(os.environ, SIGA_SHARD_MIN_ROWS="0"))
            assert sorted(f for f in os.listdir(ours) if f.startswith(prefix)) == sorted(prefix + e for e in exts)
            for ext in exts:
                assert open(os.path.join(ours, prefix + ext), "rb").read() == open(os.path.join(ref, "reads" + ext), "rb").read(), (prefix, ext)


def test_whole_pipeline_through_the_front_end_gives_the_reference_contigs():
    """index -> rmdup -> index -> overlap with this repo's front-end (host stand-in of the device), then the reference's
    assembler: the same contigs as the all-reference pipeline."""
    rng = np.random.default_rng(43)
    g = synth.random_genome(5000, rng)
    reads = synth.sample_reads(g, 2000, 70, rng, err=0.0)

    def contigs(path):
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        out = []
        for rec in open(path, "rb").read().split(b">")[1:]:
            s = b"".join(rec.split(b"\n")[1:])
            out.append(min(s, s.translate(comp)[::-1]))
        return sorted(out)

    with tempfile.TemporaryDirectory() as d:
        res = {}
        for who in ("ours", "ref"):
            w = os.path.join(d, who)
            os.mkdir(w)
            src = os.path.join(w, "reads.fa")
            synth.write_fasta(src, reads)
            run = (lambda a: _cli(a, w)) if who == "ours" else (lambda a: refrun.run_ref(a, w))
            run(["index", src])
            run(["rmdup", src])
            dedup = os.path.join(w, "reads.rmdup.fa")
            run(["index", dedup])
            run(["overlap", "-m", "35", dedup])
            refrun.run_ref(["assemble", "-m", "35", "-p", "asm", os.path.join(w, "reads.rmdup.asqg.gz")], w)
            res[who] = contigs(os.path.join(w, "asm-contigs.fa"))
        assert res["ours"] == res["ref"] and sum(len(c) for c in res["ref"]) > 3000


def test_ini_file_and_bzip2_input():
    """-s/--ini FILE: options as key=value lines, overridden by the command line (src/main.cpp:63-77); .bz2 input
    (src/utils.cpp:92-95) gives the same files as the plain input."""
    rng = np.random.default_rng(8)
    g = synth.random_genome(1200, rng)
    reads = synth.sample_reads(g, 200, 40, rng)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "reads.fa")
        synth.write_fasta(src, reads)
        refrun.run_ref(["index", src], d)
        want = {e: open(os.path.join(d, "reads" + e), "rb").read() for e in (".bwt", ".rbwt", ".sai", ".rsai")}
        refrun.run_ref(["overlap", "-m", "17", src], d)
        want_asqg = _asqg(os.path.join(d, "reads.asqg.gz"))
        # bzip2 input
        subprocess.run(["bzip2", "-k", src], check=True)
        sub = os.path.join(d, "bz")
        os.mkdir(sub)
        _cli(["index", "-p", os.path.join(sub, "reads"), src + ".bz2"], d)
        for e, b in want.items():
            assert open(os.path.join(sub, "reads" + e), "rb").read() == b, e
        # ini file: min-overlap from the file; a second run overrides it on the command line
        ini = os.path.join(d, "opts.ini")
        with open(ini, "w") as f:
            f.write("; options\n[overlap]\nmin-overlap = 17\n")
        os.remove(os.path.join(d, "reads.asqg.gz"))
        _cli(["overlap", "-s", ini, src], d)
        got = _asqg(os.path.join(d, "reads.asqg.gz"))
        assert got[0] == want_asqg[0] and sorted(got) == sorted(want_asqg)
        _cli(["overlap", "--ini", ini, "-m", "25", src], d)
        assert _asqg(os.path.join(d, "reads.asqg.gz"))[0].split("\t")[2] == "OL:i:25"
        r = _cli(["overlap", "-s", os.path.join(d, "missing.ini"), src], d, ok=False)
        assert "load" in r.stderr and "failed" in r.stderr
