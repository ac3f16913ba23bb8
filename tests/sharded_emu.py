"""CPU emulation of the bucket-sharded index build (siga_b200/csrc/index_sharded.cu) -- test infrastructure only.

The per-rank DATA steps (local rank, send buckets, placement, merge, run-length encoding) are restated in numpy /
Python; every PLANNING step the ranks derive from all-gathered values (bucket bounds, send/receive offsets, the
per-(sender, symbol) row constants, the seam rule of the RLE) is the product's own C++ (siga_b200/csrc/shard_plan.h),
called through build/libhostemu.so.  Transport: `LocalNet` (all shards in one process) or `GlooNet` (one shard per
torch.distributed rank, gloo backend) -- the same all-gather + all-to-all pattern the NCCL path uses.
"""
import ctypes
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(os.path.join(REPO, "build", "libhostemu.so"))
        L.hemu_shard_pass_counts.restype = ctypes.c_uint64
        L.hemu_shard_recv_plan.restype = ctypes.c_uint64
        _lib = L
    return _lib


def consts():
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    ks, kt = ctypes.c_uint32(), ctypes.c_uint32()
    lib().hemu_plan_consts(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(ks), ctypes.byref(kt))
    return dict(max_shards=a.value, send_buckets=b.value, meta_n=c.value, key_space=ks.value, key_top=kt.value)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


CODE = {65: 1, 67: 2, 71: 3, 84: 4}  # 'A' 'C' 'G' 'T' -> symbol rank


class LocalNet:
    """All `world` shards live in this process; collectives are list shuffles."""

    def __init__(self, world):
        self.world, self.shards = world, list(range(world))

    def allgather(self, per_shard):  # {g: value} for my shards -> [value of shard 0, 1, ...]
        return [per_shard[g] for g in range(self.world)]

    def alltoall(self, per_shard):  # {s: [items for d = 0..W-1]} -> {d: [block from s = 0..W-1]}
        return {d: [per_shard[s][d] for s in range(self.world)] for d in range(self.world)}


class GlooNet:
    """One shard per torch.distributed rank."""

    def __init__(self, dist):
        self.dist, self.world, self.shards = dist, dist.get_world_size(), [dist.get_rank()]

    def allgather(self, per_shard):
        out = [None] * self.world
        self.dist.all_gather_object(out, per_shard[self.shards[0]])
        return out

    def alltoall(self, per_shard):
        out = [None] * self.world
        self.dist.all_gather_object(out, per_shard[self.shards[0]])  # out[s][d] = items s sends to d
        me = self.shards[0]
        return {me: [out[s][me] for s in range(self.world)]}


def sharded_bwt_runs(reads, sord, net, slots=False):
    """Run bytes of the BWT of `reads` (list of bytes, already in the indexed orientation) given the sentinel order.
    Returns (segments per local shard {g: bytes}, rows per shard)."""
    K = consts()
    W, N = net.world, len(reads)
    L = lib()
    lens = [len(r) for r in reads]
    # ---- bucket bounds from the key histogram (all suffixes; the product samples reads, any bounds are valid)
    hist = np.zeros(K["key_space"], dtype=np.uint64)
    for r in reads:
        key = 0
        hist[0] += 1
        for col in range(len(r) - 1, -1, -1):
            key = CODE[r[col]] * K["key_top"] + key // 5
            hist[key] += 1
    bounds = np.zeros(K["max_shards"] + 1, dtype=np.uint32)
    L.hemu_shard_bounds(_p(hist), W, _p(bounds))
    # ---- pass 0: the sentinel rows on shard 0
    B = {g: [] for g in net.shards}
    items = {g: [] for g in net.shards}  # (id, row, sym, key) in row order
    if 0 in net.shards:
        for k, i in enumerate(sord):
            sym = CODE[reads[i][-1]]
            B[0].append(sym)
            items[0].append((int(i), k, sym, 0))
    n_cur = np.zeros(W, dtype=np.uint64)
    n_cur[0] = N
    for j in range(1, max(lens) + 1):
        meta, send = {}, {}
        for g in net.shards:
            b = np.asarray(B[g], dtype=np.int64)
            tot = [int((b == c + 1).sum()) for c in range(4)]
            buckets = [[] for _ in range(K["send_buckets"])]
            for (i, row, sym, key) in items[g]:
                if sym == 0:
                    continue  # the read ended in the previous pass
                c = sym - 1
                part = sum(tot[:c]) + int((b[:row] == sym).sum())
                nkey = sym * K["key_top"] + key // 5
                d = L.hemu_shard_owner(_p(bounds), W, ctypes.c_uint32(nkey))
                buckets[d * 4 + c].append((part, i, nkey))
            meta[g] = [len(x) for x in buckets] + tot
            send[g] = [sum((buckets[d * 4 + c] for c in range(4)), []) for d in range(W)]
        M = np.ascontiguousarray(np.array(net.allgather(meta), dtype=np.uint64).reshape(W, K["meta_n"]))
        cnt, soff, roff = (np.zeros(W * W, dtype=np.uint64) for _ in range(3))
        n_in, n_new, first_row = np.zeros(W, dtype=np.uint64), np.zeros(W, dtype=np.uint64), np.zeros(W + 1, dtype=np.uint64)
        moved = L.hemu_shard_pass_counts(_p(M), W, _p(n_cur), _p(cnt), _p(soff), _p(roff), _p(n_in), _p(n_new), _p(first_row))
        if moved == 0:
            break
        recv = net.alltoall(send)
        for g in net.shards:
            buf = sum(recv[g], [])
            assert len(buf) == int(n_in[g])
            for s in range(W):  # the offsets the NCCL path would use match the sender-major concatenation
                assert int(roff[g * W + s]) == sum(len(recv[g][q]) for q in range(s))
            if slots:
                # the peer-memory transport (index_sharded.cu): the receive "buffer" is the shard's region, one slot of
                # `slot_cap` items per sender (worst case: every read), whatever the counts -- the plan only needs to
                # be told where each sender's block starts
                slot_cap = N
                buf = []
                for q in range(W):
                    buf += list(recv[g][q]) + [None] * (slot_cap - len(recv[g][q]))
                roff = roff.copy()
                for q in range(W):
                    roff[g * W + q] = q * slot_cap
            out0 = np.zeros(K["send_buckets"] + 1, dtype=np.uint64)
            src0 = np.zeros(K["send_buckets"], dtype=np.uint64)
            Kc = np.zeros(K["send_buckets"], dtype=np.int64)
            covered = L.hemu_shard_recv_plan(_p(M), W, g, ctypes.c_uint64(N), _p(roff), _p(first_row), _p(out0), _p(src0), _p(Kc))
            assert covered == int(n_in[g])
            ent = 4 * W
            new_items = []
            for d in range(int(n_in[g])):
                e = int(np.searchsorted(out0[:ent], d, side="right")) - 1
                part, i, nkey = buf[int(src0[e]) + d - int(out0[e])]
                row = part + int(Kc[e])
                sym2 = 0 if lens[i] == j else CODE[reads[i][lens[i] - j - 1]]
                new_items.append((i, row, sym2, nkey))
            rows = [it[1] for it in new_items]
            assert rows == sorted(rows) and len(set(rows)) == len(rows), "placement must yield strictly increasing rows"
            total = len(B[g]) + len(new_items)
            assert all(0 <= r < total for r in rows)
            merged, old, ins = [], iter(B[g]), {it[1]: it[2] for it in new_items}
            for r in range(total):
                merged.append(ins[r] if r in ins else next(old))
            B[g], items[g] = merged, new_items
        n_cur = n_new.copy()
    # ---- run-length encoding with the seam rule
    seam_local, runs_local = {}, {}
    for g in net.shards:
        b = B[g]
        runs = []
        for s in b:
            if runs and runs[-1][0] == s:
                runs[-1][1] += 1
            else:
                runs.append([s, 1])
        runs_local[g] = runs
        seam_local[g] = [len(b), runs[0][0], runs[0][1], runs[-1][0], runs[-1][1]] if b else [0, 0, 0, 0, 0]
    seam = np.ascontiguousarray(np.array(net.allgather(seam_local), dtype=np.uint64).reshape(W, 5))
    segs = {}
    for g in net.shards:
        runs = [list(r) for r in runs_local[g]]
        skip, ext = ctypes.c_int(), ctypes.c_uint64()
        L.hemu_shard_seam_rule(_p(seam), W, g, ctypes.c_uint64(len(runs)), ctypes.byref(skip), ctypes.byref(ext))
        if runs:
            runs[-1][1] += ext.value
            if skip.value:
                runs[0][1] = 0
        out = bytearray()
        for s, n in runs:
            while n > 0:
                c = min(31, n)
                out.append((s << 5) | c)
                n -= c
        segs[g] = bytes(out)
    return segs, {g: len(B[g]) for g in net.shards}


def sentinel_order(sai, model):
    """Order of the N sentinel rows (index.cu: sentinel_order_kernel; model A: read index order)."""
    n = len(sai)
    if model == "A":
        return list(range(n))
    k0 = list(sai).index(0)
    sord = [0] * n
    sord[0] = n - 1
    for k in range(n):
        if k != k0:
            sord[1 + k - (1 if k > k0 else 0)] = int(sai[k]) - 1
    return sord
