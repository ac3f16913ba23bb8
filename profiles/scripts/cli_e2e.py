"""End-to-end CLI timing (file in -> files out) of this repo's `siga` front-end on a named config.

  python profiles/scripts/cli_e2e.py cfg2 [scale] [ref_scale]

Writes the synthetic FASTA, runs `siga index`, `siga overlap -m 45` (and `siga rmdup`) from siga_b200/bin and prints
their own wall-clock breakdown lines.  With ref_scale > 0 the reference binary (oracle/_ref/siga, all host threads)
runs the same commands on a scaled-down read set for the CPU line beside it.
"""
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from siga_b200 import synth  # noqa: E402


def run(binary, args, cwd, env=None):
    t0 = time.perf_counter()
    r = subprocess.run([binary] + args, cwd=cwd, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit("failed: %s %s\n%s" % (binary, args, r.stderr[-1500:]))
    return dt, [l for l in r.stderr.splitlines() if "wall" in l or "device" in l]


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    ref_scale = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    cli = os.path.join(REPO, "siga_b200", "bin", "siga")
    ref = os.path.join(REPO, "oracle", "_ref", "siga")
    with tempfile.TemporaryDirectory() as d:
        reads = synth.config_reads(cfg, scale)
        fa = os.path.join(d, "reads.fa")
        synth.write_fasta(fa, reads)
        n, L = reads.shape
        print("%s scale %g: %d reads x %d bp, FASTA %.1f MB" % (cfg, scale, n, L, os.path.getsize(fa) / 1e6))
        for rep in range(2):  # the first run pays the page-in of the binary / driver start-up
            for args in (["index", fa], ["overlap", "-m", "45", fa], ["rmdup", fa], ["correct", fa]):
                dt, lines = run(cli, args, d)
                unit = n * L / dt / 1e9 if args[0] == "index" else n / dt / 1e6
                print("run %d: siga %-8s %8.3f s  (%.3f %s)" % (rep, args[0], dt, unit, "Gbases/s" if args[0] == "index" else "M reads/s"))
                for l in lines:
                    print("      " + l)
        print("output sizes: " + ", ".join("%s %.1f MB" % (f, os.path.getsize(os.path.join(d, f)) / 1e6)
                                           for f in sorted(os.listdir(d)) if not f.endswith(".fa") or "rmdup" in f))
    if ref_scale > 0 and os.path.isfile(ref):
        with tempfile.TemporaryDirectory() as d:
            reads = synth.config_reads(cfg, ref_scale)
            fa = os.path.join(d, "reads.fa")
            synth.write_fasta(fa, reads)
            n, L = reads.shape
            env = dict(os.environ, ORACLE_LOG_LEVEL="5")
            th = str(os.cpu_count() or 1)
            print("reference binary, %s scale %g: %d reads, %s host threads" % (cfg, ref_scale, n, th))
            for args in (["index", fa], ["overlap", "-m", "45", "-t", th, fa], ["rmdup", "-t", th, fa], ["correct", "-t", th, fa]):
                dt, _ = run(ref, args, d, env)
                unit = n * L / dt / 1e9 if args[0] == "index" else n / dt / 1e6
                print("      siga %-8s %8.3f s  (%.4f %s)" % (args[0], dt, unit, "Gbases/s" if args[0] == "index" else "M reads/s"))


if __name__ == "__main__":
    main()
