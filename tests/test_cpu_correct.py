"""CPU tests of the k-mer correction row: correct_logic.h -- the per-read control logic the
CUDA kernel runs -- instantiated with the single-lane host engine (tests/hostemu), against the committed outputs of the
compiled reference (`siga correct`, tests/golden/make_golden_correct.py) and against the reference binary itself."""
import os

import numpy as np
import pytest

import correct_common as cc
import refrun


def _lib():
    lib = cc.hostemu_lib()
    if lib is None:
        pytest.skip("build/libhostemu.so not built (make hostemu)")
    return lib


@pytest.mark.parametrize("path", cc.GOLDEN, ids=[os.path.basename(p)[:-5] for p in cc.GOLDEN])
def test_correction_logic_reproduces_reference_output(path):
    fx = cc.load(path)
    runs = np.frombuffer(bytes.fromhex(fx["bwt_runs_hex"]), dtype=np.uint8)
    reads, valid, nfix, lookups = cc.hostemu_correct(_lib(), runs, fx["reads"], fx["quals"], fx["options"])
    assert cc.ec_text(reads, fx["quals"], valid) == fx["ec_text"]
    assert int(nfix.sum()) > 0 and 0 < int(valid.sum()) < len(reads)  # the fixture exercises fixes and dropped reads
    # dropped reads keep their input bases; reads shorter than k are always dropped (src/correct_processor.cpp:84-88)
    for i, r in enumerate(fx["reads"]):
        if not valid[i]:
            assert reads[i] == r
        if len(r) < fx["options"]["k"]:
            assert not valid[i]


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
@pytest.mark.parametrize("case", [
    dict(seed=1, opt=dict(k=21, x=3, i=10, O=1)),
    dict(seed=2, opt=dict(k=31, x=3, i=10, O=1), fastq=True),
    dict(seed=3, opt=dict(k=17, x=4, i=3, O=0), fastq=True, variable=True),
    dict(seed=4, opt=dict(k=1, x=3, i=10, O=1)),
    dict(seed=5, opt=dict(k=60, x=1, i=10, O=1)),               # k = read length: one k-mer per read
    dict(seed=6, opt=dict(k=19, x=2, i=0, O=1), both=True),     # no correction rounds at all
    dict(seed=7, opt=dict(k=15, x=-1, i=10, O=1)),              # negative threshold wraps to a huge size_t in the reference
    # k > 64 and reads of more than 32 x 32 bases: multi-word shifts in the solid-base / quality dilation, word loops
    dict(seed=31, opt=dict(k=100, x=3, i=10, O=1), fastq=True, n_reads=300, read_len=300, genome=2000, err=0.01),
    dict(seed=33, opt=dict(k=70, x=2, i=5, O=1), fastq=True, variable=True, n_reads=250, read_len=400, genome=3000),
    dict(seed=36, opt=dict(k=31, x=3, i=10, O=1), fastq=True, n_reads=150, read_len=1100, genome=4000, err=0.01),
    # -x -1 with qualities: requiredSupport is 2^64 - 1 below the phred cut-off and 0 from it (every such k-mer "solid")
    dict(seed=34, opt=dict(k=33, x=-1, i=10, O=1), fastq=True, n_reads=300, read_len=120, genome=1200, err=0.03),
    dict(seed=35, opt=dict(k=32, x=0, i=10, O=0), fastq=True, n_reads=300, read_len=120, genome=1200, err=0.03),
], ids=lambda c: "seed%d_k%d" % (c["seed"], c["opt"]["k"]))
def test_correction_logic_equals_reference_binary(case):
    kw = {k: v for k, v in case.items() if k not in ("seed", "opt")}
    reads, quals = cc.random_case(case["seed"], **kw)
    runs, _, want = cc.reference_correct(reads, quals, case["opt"])
    got_reads, valid, _, _ = cc.hostemu_correct(_lib(), runs, reads, quals, case["opt"])
    assert cc.ec_text(got_reads, quals, valid) == want


def test_lookup_count_of_clean_reads_is_the_nominal_figure():
    """Error-free reads at high coverage: every k-mer is solid after one counting round, so the executed rank lookups
    are exactly 2 (k-1) (n-k+1) per read -- the algorithmic-work figure DESIGN.md quotes for the correction kernel."""
    rng = np.random.default_rng(9)
    from siga_b200 import formats, synth
    from oracle import pyoracle
    g = synth.random_genome(300, rng)
    reads = [x.tobytes().decode() for x in synth.sample_reads(g, 400, 40, rng, err=0.0, both_strands=False)]
    bwt, _ = pyoracle.bwt_and_sai(reads, "B")
    runs = formats.encode_runs(bwt)
    k = 17
    out, valid, nfix, lookups = cc.hostemu_correct(_lib(), runs, reads, None, dict(k=k, x=2, i=10, O=1))
    solid = int(valid.sum())
    assert int(nfix.sum()) == 0 and out == reads and solid > 0.9 * len(reads)
    if solid == len(reads):
        assert lookups == len(reads) * 2 * (k - 1) * (40 - k + 1)


@pytest.mark.skipif(not refrun.have_ref(), reason="compiled reference (oracle/_ref/siga) not present")
def test_correction_logic_fuzz_against_reference_binary():
    """Random shapes and options (read length 20-260, k 1-90, thresholds 0-6, 0-30 rounds, offsets 0-5, error rates up to
    10 %, FASTA / FASTQ, variable length, periodic genomes): the written file must always equal the reference's.  A longer
    run of the same generator (3 568 cases) found no difference."""
    from siga_b200 import synth
    lib = _lib()
    rng = np.random.default_rng(20261018)
    for case in range(24):
        seed = int(rng.integers(1 << 30))
        L = int(rng.integers(20, 260))
        G = int(rng.integers(L + 50, 3000))
        N = int(rng.integers(100, 500))
        err = float(rng.choice([0.0, 0.005, 0.02, 0.05, 0.1]))
        fastq, var, both = bool(rng.integers(2)), bool(rng.integers(3) == 0), bool(rng.integers(2))
        opt = dict(k=int(rng.integers(1, min(L, 90) + 1)), x=int(rng.choice([0, 1, 2, 3, 4, 6])), i=int(rng.choice([0, 1, 3, 10, 30])),
                   O=int(rng.choice([0, 1, 2, 5])))
        mode = int(rng.integers(4))
        if mode == 0:
            reads, quals = cc.random_case(seed, n_reads=N, read_len=L, genome=G, err=err, fastq=fastq, variable=var, both=both)
        else:  # periodic / low-complexity genome with a few mutations
            r2 = np.random.default_rng(seed)
            unit = synth.random_genome(int(r2.integers(1, 40)), r2)
            g = np.tile(unit, G // len(unit) + 2)[:G]
            mut = r2.random(G) < 0.02 * mode
            g[mut] = synth.random_genome(int(mut.sum()), r2)
            reads = [a.tobytes().decode() for a in synth.sample_reads(g, N, L, r2, err=err, both_strands=both)]
            quals = ["".join(chr(int(v)) for v in r2.integers(33, 33 + 42, size=L)) for _ in reads] if fastq else None
        runs, _, want = cc.reference_correct(reads, quals, opt)
        got, valid, _, _ = cc.hostemu_correct(lib, runs, reads, quals, opt)
        assert cc.ec_text(got, quals, valid) == want, (case, seed, mode, L, G, N, err, fastq, var, both, opt)
