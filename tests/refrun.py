"""Helpers to run the compiled reference (oracle/_ref/siga) -- test infrastructure only."""
import glob
import os
import subprocess
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(REPO, "oracle", "_ref", "siga")


def have_ref():
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(args, cwd, threads_env=None):
    env = dict(os.environ)
    env.setdefault("ORACLE_LOG_LEVEL", "4")
    r = subprocess.run([REF_BIN] + list(args), cwd=cwd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference failed: %s\n%s" % (args, r.stderr[-2000:]))
    return r


def ref_index(fasta, cwd, algo=None):
    args = ["index"]
    if algo:
        args += ["-a", algo]
    run_ref(args + [fasta], cwd)


def ref_overlap(fasta, cwd, min_overlap, threads=1, exhaustive=False, no_rc=False):
    args = ["overlap", "-m", str(min_overlap), "-t", str(threads)]
    if exhaustive:
        args.append("-x")
    if no_rc:
        args.append("--no-opposite-strand")
    run_ref(args + [fasta], cwd)
    stem = os.path.splitext(os.path.basename(fasta))[0]
    return sorted(glob.glob(os.path.join(cwd, stem + "-thread*.hits.gz"))), os.path.join(cwd, stem + ".asqg.gz")
