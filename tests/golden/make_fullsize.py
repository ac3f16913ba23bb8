"""Pins the named configs to outputs of the compiled reference (oracle/_ref/siga) by hash.

  python tests/golden/make_fullsize.py cfg2 1.0 31,45,63,101      # ~45 min of host time on 8 cores, run once
  python tests/golden/make_fullsize.py cfg4 0.01 45
  python tests/golden/make_fullsize.py cfg5 0.01 45

For one (config, scale) it generates the seeded read set (siga_b200/synth.py), runs the reference's `siga index`
(default sais2) and `siga overlap -m M` for every M, and writes tests/golden/fullsize/<config>_<scale>.json with
  * sha256 of the .bwt/.rbwt/.sai/.rsai files, of the bare run bytes and of the .sai ids as little-endian u32
    (the buffers bench.py and the C ABI hold),
  * per M: sha256 of `zcat x.asqg.gz | LC_ALL=C sort`, record counts, and `edges_digest`: sha256 over the ED records as
    int32 rows (query index, target index, s1, e1, s2, e2, rc) sorted lexicographically -- the form a caller can
    compute from sgx_overlap_edges records without formatting text (tests/fullsize_common.py:edge_digest).
Test infrastructure only; the working directory (/tmp/sigaref/<config>_<scale>) keeps the reference files.
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import fullsize_common as fc  # noqa: E402
import refrun  # noqa: E402
from siga_b200 import formats, synth  # noqa: E402


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    cfg, scale = sys.argv[1], float(sys.argv[2])
    ms = [int(x) for x in sys.argv[3].split(",")]
    threads = os.cpu_count() or 1
    tag = "%s_%g" % (cfg, scale)
    work = os.path.join("/tmp/sigaref", tag)
    os.makedirs(work, exist_ok=True)
    out_path = os.path.join(HERE, "fullsize", tag + ".json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    rec = json.load(open(out_path)) if os.path.isfile(out_path) else {}
    reads = synth.config_reads(cfg, scale)
    n, l = reads.shape
    rec.update({"config": cfg, "scale": scale, "n_reads": int(n), "read_len": int(l),
                "reads_sha256": hashlib.sha256(np.ascontiguousarray(reads).tobytes()).hexdigest(),
                "generator": "tests/golden/make_fullsize.py %s %g (oracle/_ref/siga, index -a sais2)" % (cfg, scale)})
    fa = os.path.join(work, "reads.fa")
    if not os.path.isfile(fa):
        synth.write_fasta(fa, reads)
    if "index" not in rec or not os.path.isfile(os.path.join(work, "reads.rsai")):
        t0 = time.time()
        refrun.ref_index(fa, work)
        idx = {"seconds": round(time.time() - t0, 1)}
        for ext in ("bwt", "rbwt", "sai", "rsai"):
            idx[ext + "_sha256"] = sha_file(os.path.join(work, "reads." + ext))
        for ext in ("bwt", "rbwt"):
            ns, nf, runs = formats.read_bwt(os.path.join(work, "reads." + ext))
            idx[ext + "_runs_sha256"] = hashlib.sha256(runs.tobytes()).hexdigest()
            idx[ext + "_n_runs"] = int(len(runs))
            idx["n_suffixes"] = int(nf)
        for ext in ("sai", "rsai"):
            ids = formats.read_sai(os.path.join(work, "reads." + ext))
            idx[ext + "_u32_sha256"] = hashlib.sha256(ids.astype("<u4").tobytes()).hexdigest()
        rec["index"] = idx
        json.dump(rec, open(out_path, "w"), indent=1, sort_keys=True)
    rec.setdefault("overlap", {})
    for m in ms:
        if str(m) in rec["overlap"]:
            continue
        t0 = time.time()
        refrun.run_ref(["overlap", "-m", str(m), "-t", str(threads), fa], work)
        secs = time.time() - t0
        asqg = os.path.join(work, "reads.asqg.gz")
        p = subprocess.run("zcat %s | LC_ALL=C sort -S 4G | sha256sum" % asqg, shell=True, capture_output=True, text=True,
                           check=True)
        q, t, s1, e1, s2, e2, rc, n_vt, n_ss = fc.parse_asqg_edges(asqg)
        rec["overlap"][str(m)] = {"seconds": round(secs, 1), "threads": threads,
                                  "sorted_asqg_sha256": p.stdout.split()[0],
                                  "n_vt": int(n_vt), "n_substring": int(n_ss), "n_ed": int(len(q)),
                                  "edges_digest": fc.edge_digest_rows(q, t, s1, e1, s2, e2, rc)}
        os.rename(asqg, os.path.join(work, "reads.m%d.asqg.gz" % m))
        for f in os.listdir(work):
            if f.endswith(".hits.gz"):
                os.remove(os.path.join(work, f))
        json.dump(rec, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
