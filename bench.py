#!/usr/bin/env python
"""Benchmark of the SIGA hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU code (oracle/_ref/siga)

Workload: BASELINE.json configs[1] -- synthetic E. coli-sized 4.6 Mbp genome, 100x coverage of 150 bp reads with
1 % substitution errors, both strands (3 066 666 reads), `index` + `overlap -m 45`.
Headline metric: overlap reads/s (a step = one overlap pass over the whole read set, index resident);
the BWT-build Gbases/s of the same read set is reported in the "index" object of the same JSON line.
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale of cfg2 (1.0 = the named workload)")
    ap.add_argument("--ref-scale", type=float, default=0.01, help="genome scale of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- synthetic input, resident in pinned host memory (every rank holds the same read set)
    reads = synth.config_reads("cfg2", args.scale)
    n_reads, read_len = reads.shape
    host = torch.empty(reads.size, dtype=torch.uint8).pin_memory()
    host.numpy()[:] = reads.reshape(-1)
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    flat = host.numpy()
    h2d_bytes = int(flat.nbytes + off.nbytes)

    ctx = api.Context(local)
    sampler = ClockSampler(local)
    comm = None  # the library's own NCCL communicator: sharded index build + final gather of edge records
    if world > 1:
        uid = torch.from_numpy(ctx.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
        dist.broadcast(uid, 0)
        comm = ctx.comm_init(uid.cpu().numpy(), rank, world)

    # ---- index (both strands) : build once for the overlap stage, then time it
    r_dev = ctx.pack_reads(flat, off)
    fi = ctx.index_build(r_dev, order=api.ORDER_SAIS2, reversed=False)
    ri = ctx.index_build(r_dev, order=api.ORDER_SAIS2, reversed=True)
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

    names = ["r%d" % i for i in range(n_reads)]
    name_rank = api.name_ranks(names)
    lengths = np.full(n_reads, read_len, dtype=np.int64)

    sampler.start()
    # ---- overlap, inputs resident in HBM ("value"): search + post + hits->edges on the device, edges copied out
    for _ in range(args.warmup):
        e = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, MIN_OVERLAP, name_rank, copy=False)
        ctx.free_edges(e)
    barrier()
    t0 = time.perf_counter()
    dev_ms = search_ms = extract_ms = 0.0
    look_s = look_e = launches = search_launches = 0
    d2h_bytes = 0
    for _ in range(args.steps):
        e = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, MIN_OVERLAP, name_rank, copy=False)
        st = e.stats
        dev_ms += st["ms_total"]
        search_ms += st["ms_search"]
        extract_ms += st["ms_extract"]
        look_s += st["lookups_search"]
        look_e += st["lookups_extract"]
        launches += st["launches"]
        search_launches += st["search_launches"]
        d2h_bytes = int(e.edges.nbytes + e.substring.nbytes)
        n_blocks = st["n_blocks"]
        n_edges = st["n_edges"]
        slow_reads = st["slow_reads"]
        ctx.free_edges(e)
    barrier()
    ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    total_reads = n_reads * world
    value = total_reads / (ms_step * 1e-3)

    # size-independent parity property at full size: the edge set is closed under the reference's symmetry -- every
    # overlap is found from both reads, and exactly one of the two directions survives the id0 >= id1 rule
    chk = ctx.overlap_edges(fwd, rev, r_dev, 0, n_reads, MIN_OVERLAP, name_rank)
    ed = chk.edges
    assert len(ed) == n_edges and np.all(name_rank[ed["query"]] > name_rank[ed["target"]])
    assert np.all((ed["length"] >= MIN_OVERLAP) & (ed["length"] <= read_len))

    # ---- overlap end to end through the C ABI with HOST buffers ("e2e"): pack the pinned host reads, attach the
    #      host .sai/.rsai, search, convert, copy the edges out
    def e2e_step():
        rr = ctx.pack_reads(flat, off)
        ctx.fm_attach_reads(fwd, fi.sai, rr)
        ctx.fm_attach_reads(rev, ri.sai, rr)
        ee = ctx.overlap_edges(fwd, rev, rr, 0, n_reads, MIN_OVERLAP, name_rank, copy=False)
        nb = int(ee.edges.nbytes + ee.substring.nbytes)
        if comm is not None:  # N > 1: the edge records of every rank's shard end up in rank 0's host memory
            ga = ctx.edges_gather(comm, ee, root=0, copy=False)
            nb += int(ga.edges.nbytes + ga.substring.nbytes)
            ctx.free_edges(ga)
        ctx.free_edges(ee)
        ctx.free_reads(rr)
        return nb
    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = []
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

    # ---- index build (fwd + rev), resident and end to end.  N = 1: sgx_index_build.  N > 1: ONE index of the same
    #      read set built by all N GPUs together (sgx_index_build_sharded: rows sharded by leading k-mer bucket,
    #      inserts exchanged with NCCL all-to-all every column pass) -- strong scaling, checked against the
    #      single-GPU build of the setup phase byte for byte.
    isteps = args.index_steps or min(args.steps, 3)
    def build_index(rd, rv, copy=True):
        if comm is None:
            return ctx.index_build(rd, order=api.ORDER_SAIS2, reversed=rv, copy=copy)
        return ctx.index_build_sharded(comm, rd, order=api.ORDER_SAIS2, reversed=rv, copy=copy)

    for rv, ref_idx in ((False, fi), (True, ri)):  # warm-up + parity of the sharded build at full size
        o = build_index(r_dev, rv)
        assert np.array_equal(o.sai, ref_idx.sai)
        if rank == 0:
            assert np.array_equal(o.runs, ref_idx.runs), "sharded index differs from the single-GPU build"
    barrier()
    t0 = time.perf_counter()
    idx_dev_ms = idx_merge_ms = 0.0
    idx_merge_bytes = idx_launches = idx_merge_launches = 0
    for _ in range(isteps):
        for rv in (False, True):
            o = build_index(r_dev, rv, copy=False)  # results stay in the library's pinned buffers (no numpy copy)
            idx_dev_ms += o.stats["ms_total"]
            idx_merge_ms += o.stats["ms_merge"]
            idx_merge_bytes += o.stats["merge_bytes"]
            idx_launches += o.stats["launches"]
            idx_merge_launches += o.stats["merge_launches"]
            ctx.free_index(o)
    barrier()
    idx_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / isteps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(isteps):
        rr = ctx.pack_reads(flat, off)
        for rv in (False, True):
            o = build_index(rr, rv, copy=False)
            ctx.free_index(o)
        ctx.free_reads(rr)
    barrier()
    idx_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / isteps)
    idx_dev_ms = max_over_ranks(idx_dev_ms)

    # ---- k-mer error correction (`siga correct`, defaults -k 31 -x 3 -i 10 -O 1) of the same reads against the forward
    #      index: resident (reads + index in HBM) and end to end (pinned host reads in, flags + fixes out)
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
    if comm is not None:
        ctx.free_comm(comm)
    clocks = sampler.stop()

    bases = n_reads * read_len
    # random-access roofline of the search kernel: executed rank lookups x 128 B over its device time
    search_s = search_ms * 1e-3 / args.steps
    ach_gbs = (look_s / args.steps) * 128 / search_s / 1e9 if search_s > 0 else 0.0
    gather = None
    try:
        gather = ctx.bench_gather(8 << 30, 1 << 27)
    except Exception:
        gather = None
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
                "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": e2e_d2h,
                "entry": "sgx_reads_pack + sgx_fm_attach_reads x2 + sgx_overlap_edges (pinned host reads / .sai / name ranks in, edges + SS flags out)"
                         + ("; + sgx_edges_gather of all shards' edge records to rank 0 (NCCL)" if world > 1 else "")},
        "roofline": {"bound": "hbm", "kernel": "ov_search_kernel", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach_gbs / hbm_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch (mode 0 of 4) at full cfg2 size,
                     # ncu --set full capture profiles/r1c_prof_search.raw.csv; its algorithmic bytes are 89.5 GB
                     "traffic": 42.14e9 if args.scale == 1.0 else None, "traffic_source": "profiles/r1c_summary.md",
                     "launches_per_step": search_launches / args.steps,
                     "algorithmic_bytes_per_launch": look_s * 128 / max(search_launches, 1),
                     "peak_kind": peak_kind + " STREAM copy",
                     "lookups_per_read": look_s / args.steps / n_reads, "lookups_per_s": look_s / args.steps / search_s if search_s else 0,
                     "random_gather_128B_per_s": gather,
                     "frac_of_random_gather": (look_s / args.steps / search_s) / gather if gather and search_s else None,
                     "search_ms": search_ms / args.steps, "extract_ms": extract_ms / args.steps,
                     "extract_lookups_per_read": look_e / args.steps / n_reads, "blocks_per_read": n_blocks / n_reads, "edges_per_read": n_edges / n_reads,
                     "slow_path_reads": slow_reads},
        "index": {"metric": "bwt_build_gbases_per_s", "value": bases / (idx_ms * 1e-3) / 1e9, "unit": "Gbases/s",
                  "ms_per_step": idx_ms, "steps": isteps, "device_ms_per_step": idx_dev_ms / isteps,
                  "scaling": "strong" if world > 1 else "single GPU",
                  "path": ("sgx_index_build_sharded over NCCL: one cfg2 index (fwd + rev) built by %d GPUs, byte-identical "
                           "to the single-GPU build" % world) if world > 1 else "sgx_index_build",
                  "e2e": {"value": bases / (idx_e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": idx_e2e_ms,
                          "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(len(fi.runs) + len(ri.runs) + 8 * n_reads)},
                  "gpu_launches": int(idx_launches),
                  "roofline": {"bound": "hbm", "kernel": "bcr_merge_kernel", "achieved": merge_ach, "peak": hbm_peak, "unit": "GB/s",
                               "frac": merge_ach / hbm_peak, "traffic": None,
                               "merge_ms_per_build": idx_merge_ms / (2 * isteps),
                               "floor_bytes_per_suffix": idx_suffix_bytes,
                               "floor_frac": idx_suffix_bytes * 2 * fi.n_suffixes / (idx_dev_ms / isteps * 1e-3) / 1e9 / hbm_peak}},
        "correct": {"metric": "correct_reads_per_s", "value": total_reads / (corr_ms * 1e-3), "unit": "reads/s",
                    "ms_per_step": corr_ms, "steps": csteps, "kernel_ms_per_step": corr_kernel_ms / csteps,
                    "scaling": "weak", "options": "-k 31 -x 3 -i 10 -O 1 (the reference's defaults), forward index only",
                    "written_reads": int(cr.valid.sum()), "fixed_bases": int(len(cr.fixes)), "gpu_launches": int(corr_launches),
                    "e2e": {"value": total_reads / (corr_e2e_ms * 1e-3), "unit": "reads/s", "ms_per_step": corr_e2e_ms,
                            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(cr.valid.nbytes + cr.fixes.nbytes),
                            "entry": "sgx_reads_pack + sgx_correct_batch (pinned host reads in, written flags + fixes out)"},
                    "roofline": {"bound": "hbm", "kernel": "corr_kernel", "unit": "GB/s", "peak": hbm_peak,
                                 "achieved": corr_look * 128 / (corr_kernel_ms * 1e-3) / 1e9 if corr_kernel_ms else 0.0,
                                 "frac": corr_look * 128 / (corr_kernel_ms * 1e-3) / 1e9 / hbm_peak if corr_kernel_ms else 0.0,
                                 # dram__bytes_read.sum + dram__bytes_write.sum of the one launch at full cfg2 size, ncu --set
                                 # full capture profiles/r1h_prof_correct.raw.csv (the r1h kernel: same rank-block loads in
                                 # the same order as the current one, which only spends fewer instructions per step)
                                 "traffic": 75.20e9 if args.scale == 1.0 else None, "traffic_source": "profiles/r1h_summary.md",
                                 "lookups_per_read": corr_look / csteps / n_reads,
                                 "reference_lookups_per_read_nominal": 2 * 30 * (read_len - 30),
                                 "lookups_per_s": corr_look / (corr_kernel_ms * 1e-3) if corr_kernel_ms else 0.0}},
        "clocks": clocks,
    }

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
    ctx.free_fm(fwd)
    ctx.free_fm(rev)
    ctx.free_reads(r_dev)
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
