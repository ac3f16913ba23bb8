"""Shared helpers of the k-mer correction tests: fixtures, the reference run, the host stand-in and the output text."""
import ctypes
import glob
import json
import os
import tempfile

import numpy as np

import refrun
from siga_b200 import formats, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "correct", "*.json")))


def load(path):
    with open(path) as f:
        return json.load(f)


def write_reads(path, reads, quals, names=None):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            nm = names[i] if names else "r%d" % i
            if quals is None:
                f.write(">%s\n%s\n" % (nm, r))
            else:
                f.write("@%s\n%s\n+\n%s\n" % (nm, r, quals[i]))


def ec_text(reads, quals, valid, names=None):
    """The .ec.fa text PostCorrector writes (src/correct_processor.cpp:258-264, operator<< src/kseq.cpp:106-125)."""
    out = []
    for i, r in enumerate(reads):
        if not valid[i]:
            continue
        nm = names[i] if names else "r%d" % i
        out.append(">%s\n%s\n" % (nm, r) if quals is None else "@%s\n%s\n+\n%s\n" % (nm, r, quals[i]))
    return "".join(out)


def reference_correct(reads, quals, opt, names=None):
    """-> (forward run bytes, n_suffixes, .ec.fa text) from the compiled reference."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "x.fq" if quals is not None else "x.fa")
        write_reads(src, reads, quals, names)
        refrun.run_ref(["index", "--no-reverse", src], d)
        refrun.run_ref(["correct", "-t", "2", "-k", str(opt["k"]), "-x", str(opt["x"]), "-i", str(opt["i"]), "-O", str(opt["O"]), src], d)
        _, n_suffixes, runs = formats.read_bwt(os.path.join(d, "x.bwt"))
        with open(os.path.join(d, "x.ec.fa")) as f:
            return runs, int(n_suffixes), f.read()


def flat_inputs(reads, quals):
    flat, off = synth.reads_to_flat([r.encode() for r in reads])
    flat = np.ascontiguousarray(flat).copy()
    q = None if quals is None else np.frombuffer("".join(quals).encode(), dtype=np.uint8).copy()
    return flat, off, q


def apply_result(reads, flat_out, off):
    return [flat_out[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(reads))]


def hostemu_lib():
    p = os.path.join(REPO, "build", "libhostemu.so")
    if not os.path.isfile(p):
        return None
    lib = ctypes.CDLL(p)
    lib.hemu_fm_from_rle.restype = ctypes.c_void_p
    lib.hemu_fm_from_rle.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    lib.hemu_fm_free.argtypes = [ctypes.c_void_p]
    lib.hemu_correct.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                 ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def hostemu_correct(lib, runs, reads, quals, opt):
    """correct_logic.h on the single-lane host engine -> (corrected reads, valid flags, fixes per read, lookups)."""
    runs = np.ascontiguousarray(runs, dtype=np.uint8)
    fm = lib.hemu_fm_from_rle(runs.ctypes.data, len(runs))
    flat, off, q = flat_inputs(reads, quals)
    out = np.zeros_like(flat)
    valid = np.zeros(len(reads), dtype=np.uint8)
    nfix = np.zeros(len(reads), dtype=np.uint32)
    look = ctypes.c_uint64(0)
    rc = lib.hemu_correct(fm, flat.ctypes.data, q.ctypes.data if q is not None else None, off.ctypes.data, len(reads),
                          opt["k"], opt["x"], opt["i"], opt["O"], out.ctypes.data, valid.ctypes.data, nfix.ctypes.data,
                          ctypes.byref(look))
    lib.hemu_fm_free(fm)
    assert rc == 0
    return apply_result(reads, out, off), valid, nfix, look.value


def random_case(seed, n_reads=600, read_len=60, genome=1500, err=0.02, fastq=False, variable=False, both=False):
    rng = np.random.default_rng(seed)
    g = synth.random_genome(genome, rng)
    reads = [x.tobytes().decode() for x in synth.sample_reads(g, n_reads, read_len, rng, err=err, both_strands=both)]
    quals = None
    if fastq:
        quals = ["".join(chr(int(v)) for v in rng.integers(33 + 2, 33 + 41, size=read_len)) for _ in reads]
    if variable:
        lens = rng.integers(max(4, read_len // 6), read_len + 1, size=n_reads)
        reads = [s[:int(l)] for s, l in zip(reads, lens)]
        if quals is not None:
            quals = [s[:int(l)] for s, l in zip(quals, lens)]
    return reads, quals
