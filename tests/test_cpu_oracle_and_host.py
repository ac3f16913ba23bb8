"""CPU tests (run with -m "not gpu"): the oracle against the golden vectors, the host-side format logic, the
test-only host stand-in of the device control logic, and the C-ABI library surface (no compute without a GPU)."""
import ctypes
import glob
import json
import os
import re

import numpy as np
import pytest

from oracle import pyoracle
import shard
from siga_b200 import api, formats, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "*.json")))


def _load(path):
    with open(path) as f:
        return json.load(f)


# ------------------------------------------------------------------------------------------------ reference unit tests

def test_alphabet_and_rle_known_answers():
    # test/index_test.cpp:13-31 : '$ACGT' <-> 0..4 ; test/index_test.cpp:33-84 : AAACGGGTA -> 5 greedy runs
    sym = np.array([1, 1, 1, 2, 3, 3, 3, 4, 1], dtype=np.uint8)
    runs = formats.encode_runs(sym)
    assert [(int(r) >> 5, int(r) & 31) for r in runs] == [(1, 3), (2, 1), (3, 3), (4, 1), (1, 1)]
    assert np.array_equal(formats.decode_runs(runs), sym)
    long = np.array([4] * 70 + [0] + [4] * 31, dtype=np.uint8)  # runs are cut at 31 (src/bwt.cpp:16-27)
    assert [(int(r) >> 5, int(r) & 31) for r in formats.encode_runs(long)] == [(4, 31), (4, 31), (4, 8), (0, 1), (4, 31)]


def test_stem_rule_of_output_names():
    # test/utils_test.cpp:32-36 : Utils::stem("a.txt.gz") == "a" ; restated by the CLI host code, mirrored here
    def stem(p):
        for ext in (".gz", ".bz2"):
            if p.endswith(ext):
                return stem(p[:-len(ext)])
        b = os.path.basename(p)
        return b[:b.rfind(".")] if b.rfind(".") > 0 else b
    assert stem("a.txt.gz") == "a" and stem("/x/y/reads.fa") == "reads" and stem("reads") == "reads"


# ------------------------------------------------------------------------------------------------ oracle vs golden

@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_oracle_index_matches_reference_bytes(path):
    fx = _load(path)
    for algo, model in (("sais2", "B"), ("sais", "A")):
        for rev, bk, sk in ((False, "bwt_hex", "sai"), (True, "rbwt_hex", "rsai")):
            bwt, sai = pyoracle.bwt_and_sai(fx["reads"], model, reversed_reads=rev)
            got = formats.bwt_bytes(len(fx["reads"]), len(bwt), formats.encode_runs(bwt))
            assert got.hex() == fx["index_" + algo][bk], (algo, bk)
            assert sai.tolist() == fx["index_" + algo][sk], (algo, sk)


@pytest.mark.parametrize("path", [p for p in GOLDEN if any(k in p for k in ("f1_", "random_fixed36", "long_runs"))],
                         ids=lambda p: os.path.basename(p)[:-5])
def test_oracle_overlap_model_matches_reference_edges(path):
    fx = _load(path)
    if fx["options"].get("no_rc"):
        pytest.skip("the brute-force model restates the default (both strands) search")
    if len(set(len(r) for r in fx["reads"])) != 1:
        pytest.skip("the brute-force model is stated for fixed-length reads (SURVEY.md Appendix B)")
    want = sorted(l[3:] for l in fx["asqg_sorted"] if l.startswith("ED\t"))
    got = pyoracle.overlap_edges(fx["reads"], fx["names"], fx["min_overlap"], exhaustive=fx["options"].get("exhaustive", False))
    assert got == want


# ------------------------------------------------------------------------------------------------ host stand-in of the device logic

class _Block(ctypes.Structure):
    _fields_ = [("capped", ctypes.c_uint64 * 4), ("raw", ctypes.c_uint64 * 4), ("length", ctypes.c_uint32), ("af", ctypes.c_uint32)]


def _hostemu():
    p = os.path.join(REPO, "build", "libhostemu.so")
    if not os.path.isfile(p):
        pytest.skip("build/libhostemu.so not built (make hostemu)")
    lib = ctypes.CDLL(p)
    lib.hemu_fm_from_rle.restype = ctypes.c_void_p
    lib.hemu_fm_from_rle.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    lib.hemu_fm_free.argtypes = [ctypes.c_void_p]
    lib.hemu_fm_occ.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
    lib.hemu_overlap.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                 ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(api._Hits)]
    lib.hemu_hits_free.argtypes = [ctypes.POINTER(api._Hits)]
    return lib


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_device_control_logic_on_host_matches_reference_hits(path):
    """overlap_logic.h (the text the CUDA slow path runs) instantiated for one lane, against the reference's .hits."""
    lib = _hostemu()
    fx = _load(path)
    idx = fx["index_sais2"]
    fr = np.frombuffer(bytes.fromhex(idx["bwt_hex"]), dtype=np.uint8)[30:].copy()
    rr = np.frombuffer(bytes.fromhex(idx["rbwt_hex"]), dtype=np.uint8)[30:].copy()
    fwd = lib.hemu_fm_from_rle(fr.ctypes.data, len(fr))
    rev = lib.hemu_fm_from_rle(rr.ctypes.data, len(rr))
    # rank structure first: getOcc at every position against the cumulative counts of the decoded BWT
    sym = formats.decode_runs(fr)
    pos = np.arange(len(sym), dtype=np.uint64)
    got = np.zeros((len(sym), 5), dtype=np.uint64)
    lib.hemu_fm_occ(fwd, pos.ctypes.data, len(pos), got.ctypes.data)
    assert np.array_equal(got, pyoracle.occ_table(sym))
    flat, off = synth.reads_to_flat([r.encode() for r in fx["reads"]])
    h = api._Hits()
    rc = lib.hemu_overlap(fwd, rev, flat.ctypes.data, off.ctypes.data, len(off) - 1, fx["min_overlap"],
                          0 if fx["options"].get("exhaustive") else 1, 0 if fx["options"].get("no_rc") else 1, ctypes.byref(h))
    assert rc == 0
    blocks = ctypes.cast(h.blocks, ctypes.POINTER(_Block))
    for i, (sub, want) in enumerate(fx["hits"]):
        mine = sorted(tuple(blocks[k].capped) + tuple(blocks[k].raw) + (blocks[k].length, blocks[k].af)
                      for k in range(h.block_off[i], h.block_off[i + 1]))
        assert mine == [tuple(b) for b in want], (fx["name"], i)
        assert h.substring[i] == sub
    lib.hemu_hits_free(ctypes.byref(h))
    lib.hemu_fm_free(fwd)
    lib.hemu_fm_free(rev)


# ------------------------------------------------------------------------------------------------ host format logic

@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_hits_to_asqg_conversion_matches_reference(path):
    """formats.hits_to_edges + asqg_lines = Hit2OverlapConverter + the ASQG writers, fed with the reference's own hits."""
    fx = _load(path)
    n = len(fx["reads"])
    blocks, off = [], [0]
    for sub, bl in fx["hits"]:
        for b in bl:
            blocks.append((tuple(b[0:4]), tuple(b[4:8]), b[8], b[9]))
        off.append(len(blocks))
    arr = np.array(blocks, dtype=api.BLOCK_DTYPE) if blocks else np.zeros(0, dtype=api.BLOCK_DTYPE)
    sub = [s for s, _ in fx["hits"]]
    idx = fx["index_sais2"]
    edges = formats.hits_to_edges(fx["names"], [len(r) for r in fx["reads"]], sub, off, arr, idx["sai"], idx["rsai"])
    lines = formats.asqg_lines(fx["names"], fx["reads"], sub, edges, fx["min_overlap"])
    assert sorted(lines) == fx["asqg_sorted"]
    assert lines[0] == "HT\tVN:i:1\tOL:i:%d\tCN:i:1" % fx["min_overlap"] and len([l for l in lines if l.startswith("VT")]) == n


def test_bwt_and_sai_file_images_round_trip(tmp_path):
    fx = _load(GOLDEN[0])
    raw = bytes.fromhex(fx["index_sais2"]["bwt_hex"])
    p = tmp_path / "x.bwt"
    p.write_bytes(raw)
    ns, nf, runs = formats.read_bwt(str(p))
    assert formats.bwt_bytes(ns, nf, runs) == raw
    q = tmp_path / "x.sai"
    q.write_bytes(formats.sai_bytes(fx["index_sais2"]["sai"]))
    assert formats.read_sai(str(q)).tolist() == fx["index_sais2"]["sai"]


def test_synthetic_configs_have_the_named_shapes():
    r = synth.config_reads("cfg1")
    assert r.shape == (20_000, 100) and set(np.unique(r)) <= set(b"ACGT")
    pairs = np.concatenate([r[:10_000], r[10_000:]], axis=1)
    assert len(np.unique(pairs, axis=0)) == 10_000  # unique start positions => unique pairs (paired_read_gen.py:38-54)
    r2 = synth.config_reads("cfg2", 0.01)
    assert r2.shape == (30_666, 150)
    assert np.array_equal(r2, synth.config_reads("cfg2", 0.01))  # seeded


def test_read_ranges_partition_the_read_set():
    for n, w in ((10, 3), (3_066_666, 8), (5, 8), (0, 2)):
        spans = [shard.read_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


# ------------------------------------------------------------------------------------------------ C-ABI surface

def test_header_and_binding_declare_the_same_symbols():
    hdr = open(os.path.join(REPO, "include", "siga_b200.h")).read()
    declared = set(re.findall(r"\b(sgx_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol_and_refuses_to_run_without_a_gpu():
    if not os.path.isfile(api.LIB_PATH):
        pytest.skip("libsiga_b200.so not built (run __graft_entry__.build())")
    lib = api.load_library()
    for s in api.ABI_SYMBOLS:
        assert hasattr(lib, s), s
    assert lib.sgx_abi_version() == 2
    import torch
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.sgx_init(0, ctypes.byref(h)) == -2  # SGX_ENODEV: no CPU fallback
        assert b"no CUDA device" in lib.sgx_last_error(None) or b"cannot run" in lib.sgx_last_error(None)
        with pytest.raises(api.SgxError):
            api.Context(0)


# ------------------------------------------------------------------------------------ host front-end: parallel parser

def test_parallel_sequence_reader_equals_serial_reader(tmp_path):
    """siga_b200/csrc/host/seqio.cpp: the multi-threaded FASTA/FASTQ reader (plain files) must return exactly what the
    serial reader returns -- multi-line records, comments, CRLF, blank lines, '@' at the start of quality lines, and the
    irregular inputs that make it fall back (empty record, truncated FASTQ)."""
    import subprocess
    exe = tmp_path / "seqio_check"
    src = [os.path.join(REPO, "tests", "hostemu", "seqio_check.cpp"), os.path.join(REPO, "siga_b200", "csrc", "host", "seqio.cpp")]
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-o", str(exe)] + src + ["-lz", "-lpthread"], check=True)
    rng = np.random.default_rng(12)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def seq(n):
        return acgt[rng.integers(0, 4, n)].tobytes().decode()

    files = {}
    files["plain.fa"] = "".join(">r%d\n%s\n" % (i, seq(60)) for i in range(3000))
    files["multiline_comments_crlf.fa"] = "".join(
        ">q%d CR:i:%d  extra\r\n%s\r\n\r\n%s\r\n" % (i, i % 5, seq(35), seq(int(rng.integers(1, 40)))) for i in range(2500))
    files["quality_at.fq"] = "".join("@s%d c%d\n%s\n+%s\n%s\n" % (i, i, s, "" if i % 2 else "s%d c%d" % (i, i), "@" + "I" * (len(s) - 1))
                                     for i, s in ((i, seq(50)) for i in range(2500)))
    files["empty_record.fa"] = files["plain.fa"][:40000] + ">empty\n>next\nACGT\n" + files["plain.fa"][40000:]
    files["truncated.fq"] = files["quality_at.fq"][:-30]
    cut = files["plain.fa"].index("\n>", 50000) + 1
    files["bare_header.fa"] = files["plain.fa"][:cut] + ">\nACGT\n" + files["plain.fa"][cut:]  # its line joins the next record
    paths = []
    for name, text in files.items():
        p = tmp_path / name
        p.write_text(text, newline="")
        paths.append(str(p))
    env = dict(os.environ, SIGA_PARSE_MIN_PARALLEL="1000")
    r = subprocess.run([str(exe)] + paths, capture_output=True, text=True, env=env)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0 and len(lines) == len(paths) and all(" same " in l for l in lines), r.stdout + r.stderr
    assert int(lines[0].split()[2]) == 3000 and int(lines[2].split()[2]) == 2500
    assert int(lines[5].split()[2]) == 3000  # the bare header adds no record
    assert lines[2].endswith(" q") and not lines[0].endswith(" q")  # qualities kept for FASTQ only


# ------------------------------------------------------------------------- the roofline numerator: executed rank lookups

def _longest_occurring_suffix(s, corpus):
    """Largest k such that every suffix of s of length <= k occurs in the corpus text (occurrence is monotone)."""
    lo, hi = 0, len(s)
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if s[len(s) - mid:] in corpus:
            lo = mid
        else:
            hi = mid - 1
    return lo


def expected_search_lookups(reads):
    """executed 128-byte rank lookups of the four OverlapBlockFinder::find passes of every read --
    two per step while the interval of the current suffix is non-empty, four more for the containment test when the
    whole string occurs.  Counted from substring occurrence alone, independent of any FM-index."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fwd = b"$".join(reads)
    rev = b"$".join(r[::-1] for r in reads)
    total = 0
    for r in reads:
        for s, corpus in ((r, fwd), (r.translate(comp)[::-1], fwd), (r[::-1], rev), (r.translate(comp), rev)):
            k = _longest_occurring_suffix(s, corpus)
            total += 2 * min(len(s) - 1, k) + (4 if k == len(s) else 0)
    return total


def test_executed_lookup_counter_matches_a_count_from_first_principles():
    """The counter that feeds `roofline.achieved` (the same counting rule runs in the CUDA search kernel and in
    overlap_logic.h) against a count derived from substring occurrence only."""
    lib = _hostemu()
    fx = _load([p for p in GOLDEN if "random_fixed36_dups" in p][0])
    idx = fx["index_sais2"]
    fr = np.frombuffer(bytes.fromhex(idx["bwt_hex"]), dtype=np.uint8)[30:].copy()
    rr = np.frombuffer(bytes.fromhex(idx["rbwt_hex"]), dtype=np.uint8)[30:].copy()
    fwd = lib.hemu_fm_from_rle(fr.ctypes.data, len(fr))
    rev = lib.hemu_fm_from_rle(rr.ctypes.data, len(rr))
    reads = [r.encode() for r in fx["reads"]]
    flat, off = synth.reads_to_flat(reads)
    h = api._Hits()
    assert lib.hemu_overlap(fwd, rev, flat.ctypes.data, off.ctypes.data, len(off) - 1, fx["min_overlap"], 1, 1, ctypes.byref(h)) == 0
    assert int(h.lookups_search) == expected_search_lookups(reads)
    lib.hemu_hits_free(ctypes.byref(h))
    lib.hemu_fm_free(fwd)
    lib.hemu_fm_free(rev)


def test_single_symbol_lane_arithmetic_equals_full_rank_query():
    """fm_lane_sym_count (fm_layout.h) is the per-lane arithmetic of quad_rank_sym, the rank query of the correction
    kernel; summed the way a quad sums it, it must give getOcc(c, i) of the all-symbol query for every symbol and
    position (block boundaries, the last block, position -1)."""
    lib = _hostemu()
    lib.hemu_fm_occ_sym.restype = ctypes.c_uint64
    lib.hemu_fm_occ_sym.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
    lib.hemu_fm_length.restype = ctypes.c_uint64
    lib.hemu_fm_length.argtypes = [ctypes.c_void_p]
    fx = _load([p for p in GOLDEN if "random_fixed36_dups" in p][0])
    runs = np.frombuffer(bytes.fromhex(fx["index_sais2"]["bwt_hex"]), dtype=np.uint8)[30:].copy()
    fm = lib.hemu_fm_from_rle(runs.ctypes.data, len(runs))
    n = int(lib.hemu_fm_length(fm))
    rng = np.random.default_rng(3)
    pos = np.concatenate([np.arange(0, min(n, 600)), rng.integers(0, n, 1500), [n - 1]]).astype(np.uint64)
    pos = np.concatenate([pos, np.array([2 ** 64 - 1], dtype=np.uint64)])
    want = np.zeros((len(pos), 5), dtype=np.uint64)
    lib.hemu_fm_occ(fm, pos.ctypes.data, len(pos), want.ctypes.data)
    for i, p in enumerate(pos.tolist()):
        for c in range(1, 5):
            assert lib.hemu_fm_occ_sym(fm, p, c) == int(want[i, c]), (p, c)
    lib.hemu_fm_free(fm)


def test_literal_only_gzip_writer_round_trips(tmp_path):
    """siga_b200/csrc/host/fastgz.cpp (the encoder of the VT members of .asqg.gz and of `siga correct -o x.gz`): every
    member must be a valid gzip stream that decodes to its input -- empty and one-byte inputs, one and two symbols, all
    256 byte values, a frequency profile that forces the 15-bit length limit, block boundaries, ASQG-like text."""
    import gzip
    import subprocess
    import zlib
    exe = tmp_path / "fastgz_check"
    src = [os.path.join(REPO, "tests", "hostemu", "fastgz_check.cpp"), os.path.join(REPO, "siga_b200", "csrc", "host", "fastgz.cpp")]
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-o", str(exe)] + src + ["-lz"], check=True)
    rng = np.random.default_rng(3)
    skew = np.concatenate([np.full(2 ** i, i, dtype=np.uint8) for i in range(20)] + [np.arange(20, 256, dtype=np.uint8)])
    rng.shuffle(skew)
    g = synth.random_genome(20000, rng)
    reads = synth.sample_reads(g, 6000, 150, rng, err=0.01)
    files = {
        "empty": b"", "one": b"A", "same": b"A" * 100000, "two": b"AB" * 70000,
        "random256": rng.integers(0, 256, 700000, dtype=np.uint8).tobytes(),
        "length_limit": skew.tobytes(),
        "block_edge": b"ACGT" * 65536, "block_edge_plus_one": b"ACGT" * 65536 + b"N",
        "vt": b"".join(b"VT\tr%d\t%s\tSS:i:0\n" % (i, r.tobytes()) for i, r in enumerate(reads)),
    }
    for k in range(12):  # random lengths around the block size, random alphabets
        n = int(rng.integers(1, 3 * 262144))
        files["fuzz%d" % k] = rng.integers(0, int(rng.integers(1, 257)), n, dtype=np.uint8).tobytes()
    # inputs for the LZ77 variant: repeated phrases, repeats beyond the 32 KB window, maximal matches, ED-like records
    phrases = [rng.integers(65, 91, int(rng.integers(1, 400)), dtype=np.uint8).tobytes() for _ in range(40)]
    files["phrases"] = b"".join(phrases[int(i)] for i in rng.integers(0, 40, 6000))
    far = rng.integers(0, 4, 50000, dtype=np.uint8).tobytes()
    files["far_repeats"] = far + b"x" * 40000 + far + b"A" * 3000 + far
    files["ed"] = b"".join(b"ED\tr%d r%d %d 149 150 0 %d 150 %d 0\n" % (rng.integers(3e6), rng.integers(3e6), a, 149 - a, rng.integers(2))
                           for a in rng.integers(0, 105, 60000))
    files["three"] = b"ACG"
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
    sizes = {}
    for mode in ([], ["lz"]):  # literal-only encoder, then the greedy-LZ77 one
        r = subprocess.run([str(exe)] + mode + [str(tmp_path / n) for n in files], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for name, data in files.items():
            blob = (tmp_path / (name + ".gz")).read_bytes()
            assert gzip.decompress(blob) == data, (mode, name)
            d = zlib.decompressobj(31)
            assert d.decompress(blob) == data and d.eof and d.unused_data == b"", (mode, name)
            sizes[(bool(mode), name)] = len(blob)

    def zlib1(data):
        z = zlib.compressobj(1, zlib.DEFLATED, 31)
        return len(z.compress(data) + z.flush())
    # within a few percent of zlib level 1: literal-only on sequence records, LZ77 on ED records
    assert sizes[(False, "vt")] < 1.05 * zlib1(files["vt"])
    assert sizes[(True, "ed")] < 1.05 * zlib1(files["ed"])


def test_default_name_ranks_equal_string_sort():
    """synth.default_name_ranks (one integer key per name) == ranks from sorting the names as strings."""
    for n in (1, 2, 9, 10, 11, 99, 100, 101, 1234, 20000, 100001):
        assert np.array_equal(synth.default_name_ranks(n), api.name_ranks(["r%d" % i for i in range(n)])), n
