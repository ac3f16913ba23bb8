"""Regenerates tests/golden/*.json by running the compiled reference (oracle/_ref/siga, built from /root/reference by
oracle/build_ref.sh) on small seeded inputs.  Run from the repo root:  python tests/golden/make_golden.py

Each fixture stores the input reads and, for `siga index` (both `-a sais2` and `-a sais`) the exact .bwt/.rbwt bytes and
.sai/.rsai ids, and for `siga overlap` the parsed per-read hit blocks (the reference's .hits.gz) and the ASQG lines.
F1/F2 are the vectors of SURVEY.md Appendix F.
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


def cases():
    rng = np.random.default_rng(2024)
    f1 = ("CAGATTTTCATATTATGCAG TTTCATATTATGCAGAAAAT TATTATGCAGAAAATCTACT TGCAGAAAATCTACTTCGCC AAAATCTACTTCGCCTGATA "
          "CTACTTCGCCTGATACGAGT TCGCCTGATACGAGTCGGTT TGATACGAGTCGGTTATCTT CGAGTCGGTTATCTTCGGAT").split()
    yield "f1_tiling", f1, ["R%d" % i for i in range(9)], 10, {}
    f2 = ["ACGTTGCAAGGCTTAACCGGATTACA", "TTGCAAGGCTTAACCGGATTACAGGC", "ACGTTGCAAGGCTTAACCGGATTACA", "GCAAGGCTTAACCGG",
          "TGTAATCCGGTTAAGCCTTGCAACGT", "CCGGATTACAGGCTTTTGACCA"]
    yield "f2_dup_rc_substring", f2, list("abcdef"), 8, {}
    yield "f2_exhaustive", f2, list("abcdef"), 8, {"exhaustive": True}
    g = synth.random_genome(400, rng)
    reads = synth.sample_reads(g, 120, 36, rng, err=0.01)
    reads[rng.integers(0, 120, 12)] = reads[rng.integers(0, 120, 12)]
    yield "random_fixed36_dups", [r.tobytes().decode() for r in reads], None, 12, {}
    yield "random_fixed36_norc", [r.tobytes().decode() for r in reads], None, 12, {"no_rc": True}
    unit = synth.random_genome(3, rng)
    gp = np.tile(unit, 120)
    mut = rng.random(len(gp)) < 0.03
    gp[mut] = synth.random_genome(int(mut.sum()), rng)
    yield "periodic30", [r.tobytes().decode() for r in synth.sample_reads(gp, 50, 30, rng)], None, 10, {}
    var = []
    for _ in range(90):
        l = int(rng.integers(12, 40))
        s = int(rng.integers(0, 400 - l))
        r = g[s:s + l]
        var.append((synth.revcomp(r) if rng.random() < 0.5 else r).tobytes().decode())
    yield "variable_length", var, None, 8, {}
    yield "long_runs", ["A" * 40, "A" * 70, "AC" * 20, "CA" * 20, "T" * 64, "T" * 64, "A" * 33 + "C", "G" + "A" * 35], None, 6, {}


def main():
    if not refrun.have_ref():
        raise SystemExit("oracle/_ref/siga missing: run oracle/build_ref.sh first")
    for name, reads, names, m, opt in cases():
        names = names or ["r%d" % i for i in range(len(reads))]
        fx = {"name": name, "reads": reads, "names": names, "min_overlap": m, "options": opt}
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "x.fa")
            synth.write_fasta(fa, [r.encode() for r in reads], names=names)
            for algo in ("sais", "sais2"):
                refrun.ref_index(fa, d, algo=algo)
                fx["index_" + algo] = {
                    "bwt_hex": open(os.path.join(d, "x.bwt"), "rb").read().hex(),
                    "rbwt_hex": open(os.path.join(d, "x.rbwt"), "rb").read().hex(),
                    "sai": formats.read_sai(os.path.join(d, "x.sai")).tolist(),
                    "rsai": formats.read_sai(os.path.join(d, "x.rsai")).tolist(),
                }
            # the overlap stage runs on the default (sais2) index, built last
            hitfiles, asqg = refrun.ref_overlap(fa, d, m, threads=2, exhaustive=opt.get("exhaustive", False),
                                                no_rc=opt.get("no_rc", False))
            hits = formats.read_hits(hitfiles)
            fx["hits"] = [[hits[i][0], sorted(hits[i][1])] for i in range(len(reads))]
            fx["asqg_sorted"] = sorted(formats.read_asqg(asqg))
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fx, f, separators=(",", ":"))
        print(name, len(reads), "reads,", sum(1 for l in fx["asqg_sorted"] if l.startswith("ED")), "edges")


if __name__ == "__main__":
    main()
