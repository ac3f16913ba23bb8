"""Experiment: does `siga overlap` start faster when the CUDA context is created BEFORE the file loaders start
(SIGA_CUDA_FIRST=1) instead of beside them?  cfg2 on tmpfs, three runs of each.

  python profiles/scripts/cli_cuda_first.py
"""
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from siga_b200 import synth  # noqa: E402

cli = os.path.join(REPO, "siga_b200", "bin", "siga")
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    reads = synth.config_reads("cfg2", 1.0)
    fa = os.path.join(d, "reads.fa")
    synth.write_fasta(fa, reads)
    for env_extra in ({}, {"SIGA_CUDA_FIRST": "1"}, {}, {"SIGA_CUDA_FIRST": "1"}):
        for cmd in (["index", fa], ["overlap", "-m", "45", fa]):
            t0 = time.perf_counter()
            r = subprocess.run([cli] + cmd, cwd=d, capture_output=True, text=True, env=dict(os.environ, **env_extra))
            dt = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr[-1000:]
            print("%-24s siga %-8s %.3f s" % (env_extra or "default", cmd[0], dt))
            for l in r.stderr.splitlines():
                if "load" in l or "wall" in l:
                    print("      " + l)
