#!/usr/bin/env python
"""Benchmark of the SIGA hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU code (oracle/_ref/siga)

Workload: BASELINE.json configs[1] -- synthetic E. coli-sized 4.6 Mbp genome, 100x coverage of 150 bp reads with
1 % substitution errors, both strands (3 066 666 reads), `index` + `overlap -m 45`.
Headline metric: overlap reads/s (a step = one overlap pass over the whole read set, index resident);
the BWT-build Gbases/s of the same read set is reported in the "index" object of the same JSON line.
Further objects of the line: "strong" (BASELINE.json configs[2]: ONE cfg2 read set split over the N ranks by contiguous
range -- what `siga overlap --gpus N` does -- for min-overlap 31 / 45 / 63 / 101), "cli" (N = 1: `siga index` and
`siga overlap` file to file), "cfg4" (N = 8: the 100 Mbp config, index sharded over the 8 GPUs + overlap, each with its
1-GPU time), "parity" (full-size outputs against hashes of the reference's own files, tests/golden/fullsize).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from siga_b200 import synth  # noqa: E402

MIN_OVERLAP = 45
WORKLOAD = "cfg2: 4.6 Mbp random genome, 100x, 150 bp reads, 1% substitutions, both strands, 3066666 reads; index + overlap -m 45"


def read_range(n_reads, rank, world):
    """Contiguous, balanced read range of a rank (first, count): the split `siga overlap --gpus N` makes."""
    first = (n_reads * rank) // world
    return first, (n_reads * (rank + 1)) // world - first


def golden_fullsize(cfg, scale):
    """Hashes of the reference's own outputs for (config, scale) (tests/golden/make_fullsize.py), or None."""
    p = os.path.join(REPO, "tests", "golden", "fullsize", "%s_%g.json" % (cfg, scale))
    return json.load(open(p)) if os.path.isfile(p) else None


def sha256(b):
    import hashlib
    return hashlib.sha256(b).hexdigest()


def check_index_hashes(golden, fi, ri):
    """Run bytes and .sai ids of both indices against the reference's files (sha256 of buffers already held)."""
    ix = golden["index"]
    assert fi.n_suffixes == ix["n_suffixes"]
    assert sha256(fi.runs.tobytes()) == ix["bwt_runs_sha256"], ".bwt run bytes differ from the reference's"
    assert sha256(ri.runs.tobytes()) == ix["rbwt_runs_sha256"], ".rbwt run bytes differ from the reference's"
    assert sha256(fi.sai.astype("<u4").tobytes()) == ix["sai_u32_sha256"], ".sai differs from the reference's"
    assert sha256(ri.sai.astype("<u4").tobytes()) == ix["rsai_u32_sha256"], ".rsai differs from the reference's"


# A strand group shards its strand over its GPUs only when the strand is big: a column pass of a config-2-sized build
# (463 M rows) takes 0.3 ms on one GPU, less than the collectives of a sharded pass cost, so one GPU per strand is fastest
# there (measured: 51.7 ms on 2 GPUs, 60.3 ms on 4 with groups of 2); config 4 (5.03 G rows) and 5 (2.57 G) shard.
SHARD_MIN_ROWS = int(os.environ.get("SIGA_SHARD_MIN_ROWS", str(1 << 31)))


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Lifeline:
    """Keeps the one JSON line printable while an OPTIONAL multi-rank object (cfg4 / cfg5 at N = 8) runs.

    Those objects are long chains of collectives: a rank that throws leaves its peers inside a collective that never
    completes, and the line with the headline numbers -- complete by then -- would be lost with them.  Every rank runs a
    watcher thread; a rank that fails drops a flag file (tmpfs, keyed by the job); a watcher that sees the flag, or whose
    object overruns its deadline, ends its process with status 0 -- rank 0 first prints the line, with
    {"error": ...} in the object's place.  Nothing here touches a timed region."""

    def __init__(self, rank, line, key, poll_s=0.25, out=None):
        self.rank, self.line, self.poll_s = rank, line, poll_s
        self.out = out if out is not None else sys.stdout
        d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        self.flag = os.path.join(d, "siga_b200_bench_abort_%s" % key)
        self.lock = threading.Lock()
        self.name, self.deadline = None, None
        if rank == 0:
            self.clear()
        threading.Thread(target=self._watch, daemon=True).start()

    def clear(self):
        try:
            os.remove(self.flag)
        except OSError:
            pass

    def failed(self):
        return os.path.exists(self.flag)

    def arm(self, name, seconds):
        with self.lock:
            self.name, self.deadline = name, time.monotonic() + seconds

    def disarm(self):
        with self.lock:
            self.name, self.deadline = None, None

    def fail(self, name, text):
        """Called by the rank whose object raised: publishes the failure and waits for its own watcher to end the process."""
        tmp = "%s.%d" % (self.flag, os.getpid())
        with open(tmp, "w") as f:
            f.write("%s: rank %d: %s" % (name, self.rank, text[:800]))
        os.replace(tmp, self.flag)
        while True:
            time.sleep(1.0)

    def _watch(self):
        while True:
            time.sleep(self.poll_s)
            with self.lock:  # disarm() waits while the decision below is being made: the line is printed once
                if self.name is None:
                    continue
                why = None
                if os.path.exists(self.flag):
                    try:
                        why = open(self.flag).read()
                    except OSError:
                        why = "%s: a rank failed" % self.name
                elif time.monotonic() > self.deadline:
                    why = "%s: rank %d gave up waiting (deadline)" % (self.name, self.rank)
                if why is None:
                    continue
                if self.rank == 0:
                    self.line[self.name] = {"error": why}
                    self.out.write(json.dumps(self.line) + "\n")
                    self.out.flush()
                os._exit(0)


def run_reference_sample(scale, threads, workdir, want_index=True):
    """Times the compiled reference (`siga index`, `siga overlap -m 45 -t threads`) on a scaled-down cfg2 sample."""
    ref = os.path.join(REPO, "oracle", "_ref", "siga")
    if not os.path.isfile(ref):
        return None
    reads = synth.config_reads("cfg2", scale)
    fa = os.path.join(workdir, "sample.fa")
    synth.write_fasta(fa, reads)
    env = dict(os.environ, ORACLE_LOG_LEVEL="5")
    out = {"reads": int(reads.shape[0]), "bases": int(reads.size), "threads": threads}
    t0 = time.perf_counter()
    subprocess.run([ref, "index", fa], cwd=workdir, env=env, check=True, capture_output=True)
    out["index_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    subprocess.run([ref, "overlap", "-m", str(MIN_OVERLAP), "-t", str(threads), fa], cwd=workdir, env=env, check=True,
                   capture_output=True)
    out["overlap_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    subprocess.run([ref, "correct", "-t", str(threads), fa], cwd=workdir, env=env, check=True, capture_output=True)
    out["correct_s"] = time.perf_counter() - t0
    return out


def reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    scale = args.ref_scale
    ref = os.path.join(REPO, "oracle", "_ref", "siga")
    if not os.path.isfile(ref):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/siga was not built (no /root/reference at build time)"}))
        return 0
    with tempfile.TemporaryDirectory() as d:
        reads = synth.config_reads("cfg2", scale)
        fa = os.path.join(d, "sample.fa")
        synth.write_fasta(fa, reads)
        env = dict(os.environ, ORACLE_LOG_LEVEL="5")
        t0 = time.perf_counter()
        subprocess.run([ref, "index", fa], cwd=d, env=env, check=True, capture_output=True)
        index_s = time.perf_counter() - t0
        cmd = [ref, "overlap", "-m", str(MIN_OVERLAP), "-t", str(threads), fa]
        for _ in range(args.warmup):
            subprocess.run(cmd, cwd=d, env=env, check=True, capture_output=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            subprocess.run(cmd, cwd=d, env=env, check=True, capture_output=True)
        dt = (time.perf_counter() - t0) / args.steps
    n = int(reads.shape[0])
    value = n / dt
    sample = "cfg2 at genome scale %g: %d reads x 150 bp, same coverage and error rate; `siga overlap -m 45 -t %d` end to end" % (scale, n, threads)
    line = {
        "impl": "reference", "metric": "overlap_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "index": {"metric": "bwt_build_gbases_per_s", "value": reads.size / index_s / 1e9, "unit": "Gbases/s",
                  "seconds": index_s, "cores": 1, "note": "`siga index` (default sais2 is single-threaded), fwd + rev index of the sample"},
    }
    golden = golden_fullsize("cfg2", 1.0)
    if golden and str(MIN_OVERLAP) in golden.get("overlap", {}):
        # what the same binary did ONCE on the full-size workload (tests/golden/make_fullsize.py, 8-vCPU container): the
        # per-step sample above is bounded, this is the configuration the metric is quoted on
        g = golden["overlap"][str(MIN_OVERLAP)]
        line["reference_full_size"] = {
            "what": "the compiled reference on the full cfg2 read set (3066666 reads), run once when the golden hashes were made",
            "overlap_seconds": g["seconds"], "overlap_threads": g["threads"], "overlap_reads_per_s": 3066666 / g["seconds"],
            "index_seconds": golden["index"]["seconds"], "index_threads": 1}
    print(json.dumps(line))
    return 0


def lf_check_sample(ctx, fm, sai, reads, sample, rng):
    """Size-independent exact check of a full-size index: for `sample` random reads, walk the BWT backwards from the
    read's sentinel row (LF mapping through sgx_fm_occ); the symbols met must spell the read from its last base to its
    first, and the walk must end on the '$' row whose .sai entry is the read."""
    n, L = reads.shape
    sai = np.asarray(sai, dtype=np.int64)
    ids = np.sort(rng.choice(n, size=min(sample, n), replace=False)).astype(np.int64)
    # sentinel rows under ordering model B: the '$' of read i sorts by the suffix that follows it, i.e. by the rank of
    # read i+1 among the read-start suffixes; the last read's '$' (end of text) comes first
    inv = np.empty(n, dtype=np.int64)
    inv[sai] = np.arange(n, dtype=np.int64)
    k0 = inv[0]
    nxt = np.minimum(ids + 1, n - 1)
    kk = inv[nxt]
    row = np.where(ids == n - 1, 0, 1 + kk - (kk > k0)).astype(np.uint64)
    pred = np.array(ctx.fm_pred(fm), dtype=np.uint64)
    code = np.zeros(256, dtype=np.int64)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i + 1
    for j in range(L - 1, -2, -1):
        occ_hi = ctx.fm_occ(fm, row)
        occ_lo = ctx.fm_occ(fm, row - np.uint64(1))  # row 0 - 1 wraps to getOcc(-1) = 0
        d = (occ_hi - occ_lo).astype(np.int64)
        assert np.all(d.sum(axis=1) == 1)
        sym = d.argmax(axis=1)
        if j >= 0:
            assert np.array_equal(sym, code[reads[ids, j]]), "LF walk does not spell the read (column %d)" % j
            row = pred[sym] + occ_hi[np.arange(len(ids)), sym] - np.uint64(1)
        else:
            assert np.all(sym == 0)
            assert np.array_equal(sai[occ_hi[:, 0].astype(np.int64) - 1], ids), "LF walk ends on another read's row"
    return int(len(ids))


def run_cli(args, cwd):
    t0 = time.perf_counter()
    p = subprocess.run(args, cwd=cwd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("%s failed: %s" % (args[:2], p.stderr[-1500:]))
    return dt, p.stderr


def cli_object(reads, golden, min_overlap):
    """`siga index` + `siga overlap -m M` file to file (FASTA in /dev/shm -> .bwt/.rbwt/.sai/.rsai and .asqg.gz)."""
    cli = os.path.join(REPO, "siga_b200", "bin", "siga")
    if not os.path.isfile(cli):
        return {"unavailable": "siga_b200/bin/siga not built"}
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    n, L = reads.shape
    with tempfile.TemporaryDirectory(dir=base) as d:
        fa = os.path.join(d, "reads.fa")
        synth.write_fasta(fa, reads)
        t_index, err_i = run_cli([cli, "index", fa], d)
        t_overlap, err_o = run_cli([cli, "overlap", "-m", str(min_overlap), fa], d)
        out = {"index_s": t_index, "overlap_s": t_overlap, "index_gbases_per_s": n * L / t_index / 1e9,
               "overlap_reads_per_s": n / t_overlap, "input": "FASTA, %d bytes, tmpfs" % os.path.getsize(fa),
               "asqg_gz_bytes": os.path.getsize(os.path.join(d, "reads.asqg.gz")),
               "phases_index": [l.strip() for l in err_i.splitlines() if "wall" in l or " ms" in l or "load (" in l][-6:],
               "phases_overlap": [l.strip() for l in err_o.splitlines() if "wall" in l or " ms" in l or " s," in l][-8:]}
        if golden is not None:
            ok = True
            for ext in ("bwt", "rbwt", "sai", "rsai"):
                with open(os.path.join(d, "reads." + ext), "rb") as f:
                    ok = ok and sha256(f.read()) == golden["index"][ext + "_sha256"]
            p = subprocess.run("zcat reads.asqg.gz | LC_ALL=C sort -S 2G | sha256sum", shell=True, cwd=d, capture_output=True, text=True)
            ok = ok and p.stdout.split()[0] == golden["overlap"][str(min_overlap)]["sorted_asqg_sha256"]
            out["files_identical_to_reference"] = bool(ok)
            assert ok, "CLI output files differ from the reference's (sha256)"
    return out


T_START = time.monotonic()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale of cfg2 (1.0 = the named workload)")
    ap.add_argument("--ref-scale", type=float, default=0.01, help="genome scale of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-correct", action="store_true")
    ap.add_argument("--cfg4", default="auto", choices=["auto", "on", "off"], help="the 100 Mbp and 64 Mbp configs 4 and 5 (auto: at 8 GPUs)")
    ap.add_argument("--cfg4-scale", type=float, default=1.0)
    ap.add_argument("--big-deadline", type=float, default=300.0, help="seconds one of the cfg4 / cfg5 objects may take before the line is printed without it")
    ap.add_argument("--big-budget", type=float, default=480.0, help="a cfg4 / cfg5 object is skipped when the run is already this many seconds old")
    ap.add_argument("--index-steps", type=int, default=0, help="timed index builds (default: min(steps, 3))")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from siga_b200 import api, formats

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's version banner must not share stdout with the JSON line; NCCL honours NCCL_DEBUG_FILE only above
        # the VERSION level, so raise VERSION (the launcher's default) to WARN and send the log to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX)

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM)

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    golden = golden_fullsize("cfg2", args.scale)
    parity = {"reference_hashes": "tests/golden/fullsize/cfg2_%g.json" % args.scale if golden else None}

    # ---- synthetic input, resident in pinned host memory (every rank holds the same read set)
    reads = synth.config_reads("cfg2", args.scale)
    n_reads, read_len = reads.shape
    if golden:
        assert sha256(np.ascontiguousarray(reads).tobytes()) == golden["reads_sha256"], "generator drifted from the pinned reads"
    host = torch.empty(reads.size, dtype=torch.uint8).pin_memory()
    host.numpy()[:] = reads.reshape(-1)
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    flat = host.numpy()
    h2d_bytes = int(flat.nbytes + off.nbytes)

    ctx = api.Context(local)
    sampler = ClockSampler(local)

    # ---- index (both strands) : build once for the overlap stage, then time it
    r_dev = ctx.pack_reads(flat, off)
    fi = ctx.index_build(r_dev, order=api.ORDER_SAIS2, reversed=False)
    ri = ctx.index_build(r_dev, order=api.ORDER_SAIS2, reversed=True)
    if golden:  # full-size parity: byte-identical to the reference's .bwt/.rbwt/.sai/.rsai (hashes of one full reference run)
        check_index_hashes(golden, fi, ri)
        parity["index_identical_to_reference"] = True
    fwd = ctx.fm_create(fi.runs, fi.n_suffixes)
    rev = ctx.fm_create(ri.runs, ri.n_suffixes)
    ctx.fm_attach_reads(fwd, fi.sai, r_dev)
    ctx.fm_attach_reads(rev, ri.sai, r_dev)
    # size-independent checks on the full-size index: '$' count, symbol counts = base counts, sai is a permutation
    for idx in (fi, ri):
        sym_counts = np.bincount(idx.runs >> 5, weights=(idx.runs & 31).astype(np.float64), minlength=5).astype(np.int64)
        assert sym_counts[0] == n_reads and sym_counts.sum() == idx.n_suffixes
        assert np.array_equal(np.sort(idx.sai), np.arange(n_reads, dtype=np.uint32))
        base_counts = np.bincount(flat, minlength=256)[[65, 67, 71, 84]]
        assert np.array_equal(sym_counts[1:], base_counts)

    name_rank = synth.default_name_ranks(n_reads)

    sampler.start()
    # ---- overlap, inputs resident in HBM ("value"): search + post + hits->edges on the device, edges copied out
    for _ in range(args.warmup):
        e = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, MIN_OVERLAP, name_rank, copy=False)
        ctx.free_edges(e)
    barrier()
    t0 = time.perf_counter()
    dev_ms = search_ms = extract_ms = 0.0
    look_s = look_e = lines_s = launches = search_launches = 0
    for _ in range(args.steps):
        e = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, MIN_OVERLAP, name_rank, copy=False)
        st = e.stats
        dev_ms += st["ms_total"]
        search_ms += st["ms_search"]
        extract_ms += st["ms_extract"]
        look_s += st["lookups_search"]
        look_e += st["lookups_extract"]
        lines_s += st["lines_search"]
        launches += st["launches"]
        search_launches += st["search_launches"]
        n_blocks = st["n_blocks"]
        n_edges = st["n_edges"]
        slow_reads = st["slow_reads"]
        ctx.free_edges(e)
    barrier()
    ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    total_reads = n_reads * world
    value = total_reads / (ms_step * 1e-3)

    # ---- parity at full size: the ED set (canonical digest) and the SS flags against the reference's .asqg.gz; plus the
    #      size-independent property that exactly one direction of every overlap survives the id0 >= id1 rule
    chk = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, MIN_OVERLAP, name_rank)
    ed = chk.edges
    assert len(ed) == n_edges and np.all(name_rank[ed["query"]] > name_rank[ed["target"]])
    assert np.all((ed["length"] >= MIN_OVERLAP) & (ed["length"] <= read_len))
    if golden and str(MIN_OVERLAP) in golden.get("overlap", {}):
        g = golden["overlap"][str(MIN_OVERLAP)]
        assert len(ed) == g["n_ed"] and int(chk.substring.sum()) == g["n_substring"]
        if rank == 0:
            assert formats.edge_digest(ed, np.int64(read_len)) == g["edges_digest"], "ED set differs from the reference's"
        parity["edges_identical_to_reference_m%d" % MIN_OVERLAP] = True
    del chk, ed

    # ---- overlap end to end through the C ABI with HOST buffers ("e2e"): pack the pinned host reads, attach the
    #      host .sai/.rsai, search, convert, copy the edges out
    def pinned(a):  # every e2e input lives in pinned host memory (the reads already do)
        t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype))).pin_memory()
        t.numpy()[...] = a
        return t  # keep the tensor: its numpy view does not own the memory
    pin_keep = [pinned(off.view(np.int64)), pinned(fi.sai.view(np.int32)), pinned(ri.sai.view(np.int32)), pinned(name_rank.view(np.int32))]
    off_p = pin_keep[0].numpy().view(np.uint64)
    sai_p, rsai_p, rank_p = (t.numpy().view(np.uint32) for t in pin_keep[1:])

    def e2e_step():
        rr = ctx.pack_reads_async(flat, off_p)  # the bases stream in while the index attaches and the first batches search
        ctx.fm_attach_reads(fwd, sai_p, rr)
        ctx.fm_attach_reads(rev, rsai_p, rr)
        ee = ctx.overlap_edges(fwd, rev, rr, 0, n_reads, MIN_OVERLAP, rank_p, copy=False)
        nb = int(ee.edges.nbytes + ee.substring.nbytes)
        e2e_dev.append(ee.stats["ms_total"])
        ctx.free_edges(ee)
        ctx.free_reads(rr)
        return nb
    e2e_dev = []  # device time of the overlap call inside every e2e step (a slow step with a normal device time is the host's)
    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = []
    del e2e_dev[:]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        e2e_d2h = e2e_step()
        e2e_steps.append((time.perf_counter() - ts) * 1e3)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    e2e_value = total_reads / (e2e_ms * 1e-3)
    ctx.fm_attach_reads(fwd, fi.sai, r_dev)
    ctx.fm_attach_reads(rev, ri.sai, r_dev)
    e2e_h2d = int(flat.nbytes + off.nbytes + 2 * fi.sai.nbytes + name_rank.nbytes)

    # ---- BASELINE.json configs[2]: the same read set, min-overlap sweep, overlap stage sharded over the N ranks by
    #      contiguous read range (strong scaling of ONE read set; FM-indices replicated; what `siga overlap --gpus N`
    #      does: every rank formats and writes the records of its own range, nothing is gathered)
    first, count = read_range(n_reads, rank, world)
    strong = {"scaling": "strong", "split": "one cfg2 read set, contiguous ranges of %d reads per rank, FM-indices replicated" % count,
              "by_min_overlap": {}}
    for m in (31, 45, 63, 101):
        ctx.free_edges(ctx.overlap_edges(fwd, rev, r_dev, first, count, m, name_rank, copy=False))
        times = []
        for _ in range(max(3, min(args.steps, 5))):
            barrier()
            t0 = time.perf_counter()
            e = ctx.overlap_edges(fwd, rev, r_dev, first, count, m, name_rank, copy=False)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            st = dict(e.stats)
            ctx.free_edges(e)
            times.append(max_over_ranks(dt))
        ms = sorted(times)[len(times) // 2]
        ne = int(sum_over_ranks(st["n_edges"]))
        rec = {"ms": ms, "reads_per_s": n_reads / (ms * 1e-3), "edges": ne,
               "search_ms_max_rank": max_over_ranks(st["ms_search"]), "post_ms_max_rank": max_over_ranks(st["ms_extract"]),
               "lookups_per_read": sum_over_ranks(st["lookups_search"]) / n_reads,
               "lines_per_read": sum_over_ranks(st["lines_search"]) / n_reads}
        if golden and str(m) in golden.get("overlap", {}):
            g = golden["overlap"][str(m)]
            assert ne == g["n_ed"], ("edge count differs from the reference's", m, ne, g["n_ed"])
            if world == 1:
                ck = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, m, name_rank)
                assert formats.edge_digest(ck.edges, np.int64(read_len)) == g["edges_digest"] and \
                    int(ck.substring.sum()) == g["n_substring"], ("ED set differs from the reference's", m)
                del ck
                rec["identical_to_reference"] = "ED set digest + SS flags"
            else:
                rec["identical_to_reference"] = "ED count (the digest is checked at N = 1 and by tests/test_gpu_fullsize.py)"
            rec["reference_seconds"] = g["seconds"]
            rec["reference_threads"] = g["threads"]
        strong["by_min_overlap"][str(m)] = rec

    # ---- index build (fwd + rev), resident and end to end.  N = 1: sgx_index_build_pair (both strands side by side on two
    #      streams).  N > 1: ONE index pair of the same read set built by all N GPUs together (strong scaling): the ranks
    #      split into two groups, one per strand -- the two indices are independent jobs, so the first factor of two costs
    #      no communication -- and inside a group of N/2 > 1 GPUs the rows are sharded by leading k-mer bucket, the
    #      inserts of every column pass written straight into the receivers' memory over NVLink (peer-memory exchange, NCCL
    #      send/recv as the fall-back) and every rank keeping the run bytes of its shard (sgx_index_build_sharded_segments).  Checked byte for byte against the
    #      single-GPU build of the setup phase (itself checked against the reference's hashes above).
    isteps = args.index_steps or min(args.steps, 3)
    group = 0 if world == 1 else rank // (world // 2)   # which strand this rank works on
    group_size = 1 if world == 1 else world // 2
    sub_world = group_size if fi.n_suffixes >= SHARD_MIN_ROWS else 1  # GPUs of the group that take part (SHARD_MIN_ROWS)
    sub_rank = rank % group_size
    idle = sub_rank >= sub_world  # a small build leaves the group's other GPUs out
    sub_comm = None  # the library's own NCCL communicator of this rank's strand group
    if sub_world > 1:
        ids = []
        for leader in (0, world // 2):  # each group's leader creates its NCCL id
            uid = torch.from_numpy(ctx.comm_unique_id() if rank == leader else np.zeros(128, dtype=np.uint8)).cuda()
            dist.broadcast(uid, leader)
            ids.append(uid.cpu().numpy())
        if not idle:
            sub_comm = ctx.comm_init(ids[group], sub_rank, sub_world)

    def build_index(rd, copy=True):
        """-> list of results this rank produced (both strands at N = 1, its group's strand otherwise)"""
        if idle:
            return []
        if world == 1:
            return list(ctx.index_build_pair(rd, order=api.ORDER_SAIS2, copy=copy))
        if sub_comm is None:
            return [ctx.index_build(rd, order=api.ORDER_SAIS2, reversed=bool(group), copy=copy)]
        # no gather: every rank of the group ends up with the run bytes of its own shard (sgx_index_build_sharded_segments)
        return [ctx.index_build_sharded_segments(sub_comm, rd, order=api.ORDER_SAIS2, reversed=bool(group), copy=copy)]

    res = build_index(r_dev)  # warm-up + parity at full size
    refs = [fi, ri] if world == 1 else [ri if group else fi]
    for o, ref_idx in zip(res, refs):
        assert np.array_equal(o.sai, ref_idx.sai)
        if sub_comm is not None:  # this rank's segment against its place in the single-GPU build's run bytes
            assert o.n_runs_total == len(ref_idx.runs) and \
                np.array_equal(o.runs, ref_idx.runs[o.seg_offset:o.seg_offset + len(o.runs)]), "index segment differs from the single-GPU build"
        elif sub_rank == 0:
            assert np.array_equal(o.runs, ref_idx.runs), "index differs from the single-GPU build"
    if golden:
        parity["timed_index_build_identical_to_reference"] = True  # == the setup build, which matched the hashes above
    barrier()
    t0 = time.perf_counter()
    idx_dev_ms = idx_merge_ms = 0.0
    idx_merge_bytes = idx_launches = idx_merge_launches = 0
    for _ in range(isteps):
        step_dev = 0.0
        for o in build_index(r_dev, copy=False):  # results stay in the library's pinned buffers (no numpy copy)
            step_dev = max(step_dev, o.stats["ms_total"])
            idx_merge_ms += o.stats["ms_merge"]
            idx_merge_bytes += o.stats["merge_bytes"]
            idx_launches += o.stats["launches"]
            idx_merge_launches += o.stats["merge_launches"]
            ctx.free_index(o)
        idx_dev_ms += step_dev
    barrier()
    idx_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / isteps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(isteps):
        rr = ctx.pack_reads(flat, off)
        for o in build_index(rr, copy=False):
            ctx.free_index(o)
        ctx.free_reads(rr)
    barrier()
    idx_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / isteps)
    idx_dev_ms = max_over_ranks(idx_dev_ms)
    if sub_comm is not None:
        ctx.free_comm(sub_comm)

    # ---- k-mer error correction (`siga correct`, defaults -k 31 -x 3 -i 10 -O 1) of the same reads against the forward
    #      index: resident (reads + index in HBM) and end to end (pinned host reads in, flags + fixes out)
    correct_obj = None
    if not args.no_correct:
        csteps = min(args.steps, 3)
        ctx.correct(fwd, r_dev, 0, n_reads)
        barrier()
        t0 = time.perf_counter()
        corr_kernel_ms = 0.0
        corr_look = corr_launches = 0
        for _ in range(csteps):
            cr = ctx.correct(fwd, r_dev, 0, n_reads)
            corr_kernel_ms += cr.stats["ms_kernel"]
            corr_look += cr.stats["lookups"]
            corr_launches += cr.stats["launches"]
        barrier()
        corr_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / csteps)
        # size-independent properties at full size: a written read differs from its input exactly at its fixes, every fix
        # changes the base, dropped reads carry no fixes
        fx_pos = off[cr.fixes["read"].astype(np.int64)].astype(np.int64) + (cr.fixes["pos_code"] >> 2).astype(np.int64)
        assert np.all(cr.valid[cr.fixes["read"]] == 1) and len(np.unique(fx_pos)) == len(fx_pos)
        assert np.all(flat[fx_pos] != np.frombuffer(b"ACGT", dtype=np.uint8)[cr.fixes["pos_code"] & 3])
        barrier()
        t0 = time.perf_counter()
        for _ in range(csteps):
            rr = ctx.pack_reads(flat, off)
            c2 = ctx.correct(fwd, rr, 0, n_reads)
            ctx.free_reads(rr)
        barrier()
        corr_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / csteps)
        assert np.array_equal(c2.valid, cr.valid) and len(c2.fixes) == len(cr.fixes)
        correct_obj = {
            "metric": "correct_reads_per_s", "value": total_reads / (corr_ms * 1e-3), "unit": "reads/s",
            "ms_per_step": corr_ms, "steps": csteps, "kernel_ms_per_step": corr_kernel_ms / csteps,
            "scaling": "weak", "options": "-k 31 -x 3 -i 10 -O 1 (the reference's defaults), forward index only",
            "written_reads": int(cr.valid.sum()), "fixed_bases": int(len(cr.fixes)), "gpu_launches": int(corr_launches),
            "e2e": {"value": total_reads / (corr_e2e_ms * 1e-3), "unit": "reads/s", "ms_per_step": corr_e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(cr.valid.nbytes + cr.fixes.nbytes),
                    "entry": "sgx_reads_pack + sgx_correct_batch (pinned host reads in, written flags + fixes out)"},
            "roofline": {"bound": "hbm", "kernel": "corr_kernel", "unit": "GB/s", "peak": hbm_peak,
                         "achieved": corr_look * 128 / (corr_kernel_ms * 1e-3) / 1e9 if corr_kernel_ms else 0.0,
                         "frac": corr_look * 128 / (corr_kernel_ms * 1e-3) / 1e9 / hbm_peak if corr_kernel_ms else 0.0,
                         "note": "numerator = executed rank lookups x 128 B (two lookups of one step often share a line and L2 "
                                 "absorbs the top of the index: not an HBM fraction, see DESIGN.md 4.8)",
                         # dram__bytes_read.sum + dram__bytes_write.sum of the one launch at full cfg2 size, ncu --set
                         # full capture profiles/r1h_prof_correct.raw.csv
                         "traffic": 75.20e9 if args.scale == 1.0 else None, "traffic_source": "profiles/r1h_summary.md",
                         "lookups_per_read": corr_look / csteps / n_reads,
                         "reference_lookups_per_read_nominal": 2 * 30 * (read_len - 30),
                         "lookups_per_s": corr_look / (corr_kernel_ms * 1e-3) if corr_kernel_ms else 0.0}}
    clocks = sampler.stop()

    bases = n_reads * read_len
    # random-access roofline of the search kernel: distinct 128-byte rank-block lines fetched over its device time
    search_s = search_ms * 1e-3 / args.steps
    lines_per_s = (lines_s / args.steps) / search_s if search_s > 0 else 0.0
    ach_gbs = lines_per_s * 128 / 1e9
    gather = None
    try:
        gather = ctx.bench_gather(8 << 30, 1 << 27)
    except Exception:
        gather = None
    traffic, traffic_src, merge_traffic, merge_traffic_src = None, None, None, None
    tp = os.path.join(REPO, "profiles", "r2_traffic.json")
    if args.scale == 1.0 and os.path.isfile(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("search_dram_bytes_per_launch"), tj.get("source")
        merge_traffic, merge_traffic_src = tj.get("merge_dram_bytes_of_captured_launch"), tj.get("merge_source")
    idx_suffix_bytes = {100: 31, 150: 37, 250: 50}.get(read_len, (read_len + 1 + 7) // 8 + 18)
    merge_s = idx_merge_ms * 1e-3 / max(idx_merge_launches, 1)
    merge_ach = (idx_merge_bytes / max(idx_merge_launches, 1)) / merge_s / 1e9 if merge_s > 0 else 0.0

    line = {
        "metric": "overlap_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.scale == 1.0 else WORKLOAD + " [genome scale %g]" % args.scale,
                   "reads_per_gpu": n_reads, "read_len": read_len, "min_overlap": MIN_OVERLAP,
                   "sharding": "FM-indices replicated per GPU, every rank searches its own %d-read shard, no data-path collective" % n_reads,
                   "l2": "index (2 x %.0f MB rank blocks) + candidate buffers exceed the 126 MB L2; no flush needed" %
                         (fi.n_suffixes / 256 * 128 / 1e6)},
        "device_ms_per_step": dev_ms / args.steps,
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": e2e_ms, "step_ms": [round(x, 2) for x in e2e_steps],
                "step_overlap_device_ms": [round(x, 2) for x in e2e_dev],
                "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": e2e_d2h,
                "entry": "sgx_reads_pack_async + sgx_fm_attach_reads x2 + sgx_overlap_edges (pinned host reads / .sai / name ranks in, "
                         "edges + SS flags out); every rank keeps the records of its own reads, as `siga overlap --gpus N` does"},
        "roofline": {"bound": "hbm", "kernel": "ov_search_lean_kernel", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach_gbs / hbm_peak,
                     "numerator": "distinct 128-byte rank-block lines fetched (a step whose interval ends share a block counts one line)",
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at full cfg2 size from the ncu --set full
                     # capture named in profiles/r2_traffic.json
                     "traffic": traffic, "traffic_source": traffic_src,
                     "launches_per_step": search_launches / args.steps,
                     "algorithmic_bytes_per_launch": lines_s * 128 / max(search_launches, 1),
                     "peak_kind": peak_kind + " STREAM copy",
                     "lines_per_read": lines_s / args.steps / n_reads, "lines_per_s": lines_per_s,
                     "lookups_per_read": look_s / args.steps / n_reads, "lookups_per_s": look_s / args.steps / search_s if search_s else 0,
                     "random_gather_128B_per_s": gather,
                     "frac_of_random_gather": lines_per_s / gather if gather and search_s else None,
                     "search_ms": search_ms / args.steps, "extract_ms": extract_ms / args.steps,
                     "extract_lookups_per_read": look_e / args.steps / n_reads, "blocks_per_read": n_blocks / n_reads, "edges_per_read": n_edges / n_reads,
                     "slow_path_reads": slow_reads},
        "strong": strong,
        "parity": parity,
        "index": {"metric": "bwt_build_gbases_per_s", "value": bases / (idx_ms * 1e-3) / 1e9, "unit": "Gbases/s",
                  "ms_per_step": idx_ms, "steps": isteps, "device_ms_per_step": idx_dev_ms / isteps,
                  "scaling": "strong" if world > 1 else "single GPU",
                  "path": ("one cfg2 index pair built by %d GPUs: one group of %d per strand, %d of them building%s; byte-identical to "
                           "the single-GPU build" % (world, group_size, sub_world, " (rows sharded by k-mer bucket, inserts exchanged through "
                           "peer memory over NVLink, sgx_index_build_sharded_segments)" if sub_world > 1 else " (sgx_index_build, no collective: a strand of %d rows is below the sharding "
                           "threshold of %d)" % (fi.n_suffixes, SHARD_MIN_ROWS))) if world > 1
                          else "sgx_index_build_pair (both strands side by side on two streams)",
                  "e2e": {"value": bases / (idx_e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": idx_e2e_ms,
                          "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(len(fi.runs) + len(ri.runs) + 8 * n_reads)},
                  "gpu_launches": int(idx_launches),
                  "roofline": {"bound": "hbm", "kernel": "bcr_merge64_kernel", "achieved": merge_ach, "peak": hbm_peak, "unit": "GB/s",
                               "frac": merge_ach / hbm_peak,
                               # DRAM bytes of ONE captured launch (pass 76 of 150: launches grow with the pass number, the
                               # average launch moves (n_old + n_new) / 2 + 5 B per insert = merge_bytes / merge_launches)
                               "traffic": merge_traffic, "traffic_source": merge_traffic_src,
                               "algorithmic_bytes_per_launch": idx_merge_bytes / max(idx_merge_launches, 1),
                               "merge_ms_per_build": idx_merge_ms / ((2 if world == 1 else 1) * isteps),
                               "floor_bytes_per_suffix": idx_suffix_bytes,
                               "floor_frac": idx_suffix_bytes * 2 * fi.n_suffixes / (idx_dev_ms / isteps * 1e-3) / 1e9 / hbm_peak}},
        "clocks": clocks,
    }
    if golden:
        line["reference_full_size"] = {
            "what": "the compiled reference on this exact read set, run once (tests/golden/make_fullsize.py)",
            "index_seconds": golden["index"]["seconds"], "index_threads": 1,
            "overlap_seconds": {m: v["seconds"] for m, v in golden.get("overlap", {}).items()},
            "overlap_threads": {m: v["threads"] for m, v in golden.get("overlap", {}).items()}}
    if correct_obj:
        line["correct"] = correct_obj

    if rank == 0 and world == 1 and not args.no_cli:
        ctx.free_fm(fwd)
        ctx.free_fm(rev)
        ctx.free_reads(r_dev)
        fwd = rev = r_dev = None
        try:
            line["cli"] = cli_object(reads, golden, MIN_OVERLAP)
        except Exception as ex:  # the CLI is a separate process: report, do not lose the line
            line["cli"] = {"error": str(ex)[:500]}

    want_cfg4 = args.cfg4 == "on" or (args.cfg4 == "auto" and world == 8)
    life = None
    if want_cfg4:
        if fwd is not None:
            ctx.free_fm(fwd)
            ctx.free_fm(rev)
            ctx.free_reads(r_dev)
            fwd = rev = r_dev = None
        del host, flat
        # the line is complete without these two objects: a failure or a hang inside them must not take it along
        life = Lifeline(rank, line, "%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()))
        barrier()  # rank 0 has removed a stale flag before anybody can raise a new one
        for big in ("cfg4", "cfg5"):  # BASELINE.json configs[3] and configs[4]
            life.arm(big, args.big_deadline)  # armed before the first collective: a peer may be gone since the last object
            late = torch.tensor([1 if time.monotonic() - T_START > args.big_budget else 0], device="cuda")
            if world > 1:
                dist.broadcast(late, 0)  # one decision for all ranks
            if int(late.item()):
                line[big] = {"skipped": "%.0f s had passed when this object was due (--big-budget %g)" % (time.monotonic() - T_START, args.big_budget)}
                continue
            try:
                obj = cfg4_object(args, ctx, api, formats, torch, dist, rank, world, local, barrier, max_over_ranks,
                                  sum_over_ranks, cfg=big)
                with life.lock:  # the watcher serialises the line under the same lock
                    line[big] = obj
            except BaseException as ex:  # noqa: B036 -- the peers sit in collectives this rank will never join
                life.fail(big, "%s: %s" % (type(ex).__name__, ex))
        life.disarm()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as d:
            cpu = run_reference_sample(args.ref_scale * 2.5, os.cpu_count() or 1, d)
        if cpu:
            sample = "cfg2 at genome scale %g: %d reads x 150 bp (same coverage/error); reference binary, end to end" % (
                args.ref_scale * 2.5, cpu["reads"])
            line["cpu_baseline"] = {"value": cpu["reads"] / cpu["overlap_s"], "unit": "reads/s", "cores": cpu["threads"],
                                    "kind": "reference", "sample": sample, "seconds": cpu["overlap_s"],
                                    "index_gbases_per_s": cpu["bases"] / cpu["index_s"] / 1e9, "index_seconds": cpu["index_s"],
                                    "index_cores": 1, "correct_reads_per_s": cpu["reads"] / cpu["correct_s"],
                                    "correct_seconds": cpu["correct_s"]}
    if fwd is not None:
        ctx.free_fm(fwd)
        ctx.free_fm(rev)
        ctx.free_reads(r_dev)
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if life is not None and life.failed():  # a peer went away after its last collective: do not wait for it
        sys.stdout.flush()
        os._exit(0)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cfg4_object(args, ctx, api, formats, torch, dist, rank, world, local, barrier, max_over_ranks, sum_over_ranks, cfg="cfg4"):
    """BASELINE.json configs[3] (cfg4: 100 Mbp genome, 50x 150 bp reads, 33.3 M reads, 5.03 G symbols per index) or
    configs[4] (cfg5: 64 Mbp, 40x 250 bp, 10.2 M reads, 2.57 G symbols): the index built by all ranks together (strand
    groups, rows sharded by k-mer bucket, inserts exchanged through peer memory over NVLink every column pass) and the overlap stage sharded by read
    range, each next to the time ONE GPU takes for the same job with the same kernels (rank 0, the others idle)."""
    scale = args.cfg4_scale
    shm = "/dev/shm/siga_b200_%s_%g.u8" % (cfg, scale)
    c = synth.CONFIGS[cfg]
    g = max(int(c["genome"] * scale), 4 * c["read_len"])
    n = int(g * c["coverage"] // c["read_len"])
    L = c["read_len"]
    t0 = time.perf_counter()
    if rank == 0:  # one generation, shared through tmpfs (33.3 M x 150 bytes)
        reads0 = synth.config_reads(cfg, scale)
        assert reads0.shape == (n, L)
        mm = np.lib.format.open_memmap(shm, mode="w+", dtype=np.uint8, shape=(n, L))
        mm[:] = reads0
        mm.flush()
        del mm, reads0
    barrier()
    reads = np.load(shm, mmap_mode="r")
    gen_s = time.perf_counter() - t0
    flat = reads.reshape(-1)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    out = {"workload": "%s%s: %d Mbp random genome, %dx, %d bp error-free reads, both strands, %d reads; index + overlap -m %d" %
                       (cfg, "" if scale == 1.0 else " [genome scale %g]" % scale, g // 1_000_000, c["coverage"], L, n, MIN_OVERLAP),
           "generate_s": gen_s, "n_gpus": world}
    # strand groups, as in the cfg2 index object: ranks [0, N/2) build the forward index, ranks [N/2, N) the reverse one
    group = 0 if world == 1 else rank // (world // 2)
    group_size = 1 if world == 1 else world // 2
    sub_world = group_size if n * (L + 1) >= SHARD_MIN_ROWS else 1  # see SHARD_MIN_ROWS
    sub_rank = rank % group_size
    idle = sub_rank >= sub_world
    comm = None
    if sub_world > 1:
        ids = []
        for leader in (0, world // 2):
            uid = torch.from_numpy(ctx.comm_unique_id() if rank == leader else np.zeros(128, dtype=np.uint8)).cuda()
            dist.broadcast(uid, leader)
            ids.append(uid.cpu().numpy())
        if not idle:
            comm = ctx.comm_init(ids[group], sub_rank, sub_world)
    r_dev = ctx.pack_reads(flat, off)

    def build_mine():  # results stay in the library's pinned buffers (views, no numpy copy of ~0.6 GB of run bytes)
        if idle:
            return []
        if world == 1:
            return list(ctx.index_build_pair(r_dev, order=api.ORDER_SAIS2, copy=False))
        if comm is None:
            return [ctx.index_build(r_dev, order=api.ORDER_SAIS2, reversed=bool(group), copy=False)]
        return [ctx.index_build_sharded_segments(comm, r_dev, order=api.ORDER_SAIS2, reversed=bool(group), copy=False)]

    times = []
    mine = []
    for it in range(2):
        for o in mine:
            ctx.free_index(o)
        barrier()
        t0 = time.perf_counter()
        mine = build_mine()
        torch.cuda.synchronize()
        times.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
    out["index_ms"] = min(times)
    out["index_ms_runs"] = times
    out["index_gbases_per_s"] = n * L / (out["index_ms"] * 1e-3) / 1e9
    out["index_path"] = "one group of %d GPU(s) per strand, %d of them building%s" % (
        group_size, sub_world, " (rows sharded by k-mer bucket, inserts exchanged through peer memory over NVLink)" if sub_world > 1 else " (strand below the sharding threshold)")
    # the finished indices travel through tmpfs: run bytes are complete on each group's leader, .sai everywhere in the group
    runs_path, sai_path = "/dev/shm/siga_b200_" + cfg + "_runs_%d.npy", "/dev/shm/siga_b200_" + cfg + "_sai_%d.npy"
    seg_path = "/dev/shm/siga_b200_" + cfg + "_seg_%d_%d.npy"
    if comm is not None:  # every rank of a strand group holds the run bytes of its own shard (nothing was gathered)
        np.save(seg_path % (group, sub_rank), mine[0].runs)
        seg_at = torch.tensor([mine[0].seg_offset, len(mine[0].runs), mine[0].n_runs_total], dtype=torch.int64, device="cuda")
    else:
        seg_at = torch.zeros(3, dtype=torch.int64, device="cuda")
    seg_all = [seg_at]
    if world > 1:
        seg_all = [torch.zeros(3, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(seg_all, seg_at)
    barrier()
    if sub_rank == 0:
        for k, o in enumerate(mine):
            rv = k if world == 1 else group
            if comm is not None:  # the group's segments laid end to end are the run bytes of the strand
                parts, at = [], 0
                for q in range(sub_world):
                    off_q, len_q, tot_q = (int(x) for x in seg_all[group * group_size + q].tolist())
                    assert off_q == at, "%s: the shards' segments do not tile the run bytes" % cfg
                    parts.append(np.load(seg_path % (group, q)))
                    assert len(parts[-1]) == len_q
                    at += len_q
                assert at == tot_q
                np.save(runs_path % rv, np.concatenate(parts))
                del parts
            else:
                np.save(runs_path % rv, o.runs)
            np.save(sai_path % rv, o.sai)
    n_suffixes = n * (L + 1)
    for o in mine:
        ctx.free_index(o)
    del mine
    barrier()
    if rank == 0 and world > 1:  # the same job on ONE GPU with the same kernels (both strands side by side)
        for o in ctx.index_build_pair(r_dev, order=api.ORDER_SAIS2, copy=False):
            ctx.free_index(o)
        t0 = time.perf_counter()
        a1, b1 = ctx.index_build_pair(r_dev, order=api.ORDER_SAIS2, copy=False)
        out["index_ms_one_gpu"] = (time.perf_counter() - t0) * 1e3
        for rv, one in ((0, a1), (1, b1)):
            assert np.array_equal(one.runs, np.load(runs_path % rv, mmap_mode="r")) and \
                np.array_equal(one.sai, np.load(sai_path % rv, mmap_mode="r")), "%s: multi-GPU index differs from the single-GPU build" % cfg
        ctx.free_index(a1)
        ctx.free_index(b1)
        out["index_speedup"] = out["index_ms_one_gpu"] / out["index_ms"]
        out["index_efficiency"] = out["index_speedup"] / world
        out["index_identical_to_one_gpu"] = True
    barrier()
    fms, sais = [], []
    for rv in (0, 1):
        sai = np.load(sai_path % rv)
        fm = ctx.fm_create(np.load(runs_path % rv, mmap_mode="r"), n_suffixes)
        ctx.fm_attach_reads(fm, sai, r_dev)
        fms.append(fm)
        sais.append(sai)
    fwd, rev = fms
    res = {False: type("R", (), {"sai": sais[0]})(), True: type("R", (), {"sai": sais[1]})()}
    if rank == 0:
        # exact, size-independent check of the full-size index: LF-walk a sample of reads out of the BWT
        out["lf_walk_reads_checked"] = lf_check_sample(ctx, fwd, res[False].sai, reads, 20000, np.random.default_rng(4))
    name_rank = synth.default_name_ranks(n)
    first, count = read_range(n, rank, world)
    ctx.free_edges(ctx.overlap_edges(fwd, rev, r_dev, first, count, MIN_OVERLAP, name_rank, copy=False))
    times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e = ctx.overlap_edges(fwd, rev, r_dev, first, count, MIN_OVERLAP, name_rank, copy=False)
        torch.cuda.synchronize()
        times.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
        st = dict(e.stats)
        ctx.free_edges(e)
    out["overlap_ms"] = sorted(times)[1]
    out["overlap_reads_per_s"] = n / (out["overlap_ms"] * 1e-3)
    out["edges"] = int(sum_over_ranks(st["n_edges"]))
    out["slow_path_reads"] = int(sum_over_ranks(st["slow_reads"]))
    if rank == 0 and world > 1:
        ctx.free_edges(ctx.overlap_edges(fwd, rev, r_dev, 0, n, MIN_OVERLAP, name_rank, copy=False))
        t0 = time.perf_counter()
        e = ctx.overlap_edges(fwd, rev, r_dev, 0, n, MIN_OVERLAP, name_rank, copy=False)
        out["overlap_ms_one_gpu"] = (time.perf_counter() - t0) * 1e3
        assert e.stats["n_edges"] == out["edges"], "%s: sharded overlap found a different number of edges" % cfg
        ctx.free_edges(e)
        out["overlap_speedup"] = out["overlap_ms_one_gpu"] / out["overlap_ms"]
        out["overlap_efficiency"] = out["overlap_speedup"] / world
    barrier()
    ctx.free_fm(fwd)
    ctx.free_fm(rev)
    ctx.free_reads(r_dev)
    if comm is not None:
        ctx.free_comm(comm)
    del reads, flat
    if rank == 0:
        for f in [shm] + [runs_path % rv for rv in (0, 1)] + [sai_path % rv for rv in (0, 1)] + \
                [seg_path % (gq, q) for gq in (0, 1) for q in range(group_size)]:
            try:
                os.remove(f)
            except OSError:
                pass
    return out


if __name__ == "__main__":
    sys.exit(main())
