"""Regenerates tests/golden/correct/*.json by running the compiled reference (oracle/_ref/siga: `siga index` then
`siga correct`) on small seeded inputs.  Run from the repo root:  python tests/golden/make_golden_correct.py

Each fixture stores the input reads (and FASTQ qualities), the options, the run bytes of the forward .bwt the reference
built, and the exact text of the reference's .ec.fa output.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import refrun  # noqa: E402
from siga_b200 import formats, synth  # noqa: E402


def write_reads(path, reads, quals, names=None):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            nm = names[i] if names else "r%d" % i
            if quals is None:
                f.write(">%s\n%s\n" % (nm, r))
            else:
                f.write("@%s\n%s\n+\n%s\n" % (nm, r, quals[i]))


def cases():
    rng = np.random.default_rng(31)
    g = synth.random_genome(600, rng)
    r = [x.tobytes().decode() for x in synth.sample_reads(g, 240, 40, rng, err=0.02, both_strands=False)]
    yield "correct_fasta_k15", r, None, dict(k=15, x=3, i=10, O=1)
    q = ["".join(chr(int(v)) for v in rng.integers(33 + 2, 33 + 41, size=40)) for _ in r]
    yield "correct_fastq_k15", r, q, dict(k=15, x=3, i=10, O=1)
    yield "correct_fastq_k11_x4_i2_O0", r, q, dict(k=11, x=4, i=2, O=0)
    lens = rng.integers(8, 41, size=len(r))
    vr = [s[:int(l)] for s, l in zip(r, lens)]
    vq = [s[:int(l)] for s, l in zip(q, lens)]
    yield "correct_fastq_variable_k13", vr, vq, dict(k=13, x=2, i=10, O=1)
    rb = [x.tobytes().decode() for x in synth.sample_reads(g, 240, 40, rng, err=0.02, both_strands=True)]
    yield "correct_fasta_both_strands_k13", rb, None, dict(k=13, x=2, i=10, O=2)


def run_case(reads, quals, opt, names=None):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "x.fq" if quals is not None else "x.fa")
        write_reads(src, reads, quals, names)
        refrun.run_ref(["index", "--no-reverse", src], d)
        refrun.run_ref(["correct", "-t", "2", "-k", str(opt["k"]), "-x", str(opt["x"]), "-i", str(opt["i"]), "-O", str(opt["O"]), src], d)
        _, n_suffixes, runs = formats.read_bwt(os.path.join(d, "x.bwt"))
        with open(os.path.join(d, "x.ec.fa")) as f:
            return runs, n_suffixes, f.read()


def main():
    if not refrun.have_ref():
        raise SystemExit("oracle/_ref/siga missing: run oracle/build_ref.sh first")
    for name, reads, quals, opt in cases():
        runs, n_suffixes, text = run_case(reads, quals, opt)
        fx = {"name": name, "reads": reads, "quals": quals, "options": opt, "bwt_runs_hex": runs.tobytes().hex(),
              "n_suffixes": int(n_suffixes), "ec_text": text}
        with open(os.path.join(HERE, "correct", name + ".json"), "w") as f:
            json.dump(fx, f, indent=0)
        print(name, len(reads), "reads ->", text.count("\n") // (4 if quals else 2), "written")


if __name__ == "__main__":
    main()
