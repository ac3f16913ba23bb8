"""ctypes binding of libsiga_b200.so (include/siga_b200.h) -- the only way Python reaches the hot path.

There is no CPU fallback: if the shared library is missing, or no B200-class device is usable, this module raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsiga_b200.so")

ORDER_SAIS, ORDER_SAIS2 = 0, 1
AF_QUERYREV, AF_TARGETREV, AF_QUERYCOMP = 1, 2, 4

BLOCK_DTYPE = np.dtype([("capped", "<u8", (4,)), ("raw", "<u8", (4,)), ("length", "<u4"), ("af", "<u4")])
assert BLOCK_DTYPE.itemsize == 72

# Every symbol include/siga_b200.h declares (checked by the CPU tests against the built library).
ABI_SYMBOLS = [
    "sgx_abi_version", "sgx_init", "sgx_shutdown", "sgx_last_error", "sgx_sm_count",
    "sgx_reads_pack", "sgx_reads_pack_async", "sgx_reads_wait", "sgx_reads_free", "sgx_reads_count", "sgx_reads_bases",
    "sgx_index_build", "sgx_index_build_pair", "sgx_index_out_free",
    "sgx_comm_unique_id", "sgx_comm_init", "sgx_comm_init_local", "sgx_comm_free", "sgx_comm_world", "sgx_comm_rank",
    "sgx_index_build_sharded", "sgx_index_build_sharded_segments", "sgx_edges_gather",
    "sgx_fm_create", "sgx_fm_free", "sgx_fm_attach_reads", "sgx_fm_length", "sgx_fm_pred", "sgx_fm_occ", "sgx_fm_count",
    "sgx_overlap_batch", "sgx_overlap_host", "sgx_hits_free", "sgx_overlap_edges", "sgx_edges_free",
    "sgx_correct_batch", "sgx_corrected_free",
    "sgx_bench_gather", "sgx_flush_l2",
]


class SgxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("siga_b200 error %d: %s" % (code, msg))
        self.code = code


class _IndexOut(ctypes.Structure):
    _fields_ = [("n_strings", ctypes.c_uint64), ("n_suffixes", ctypes.c_uint64), ("n_runs", ctypes.c_uint64),
                ("rle", ctypes.POINTER(ctypes.c_uint8)), ("sai", ctypes.POINTER(ctypes.c_uint32)),
                ("ms_total", ctypes.c_double), ("ms_sort", ctypes.c_double), ("ms_bcr", ctypes.c_double),
                ("ms_emit", ctypes.c_double), ("launches", ctypes.c_uint64), ("merge_launches", ctypes.c_uint64),
                ("ms_merge", ctypes.c_double), ("merge_bytes", ctypes.c_uint64), ("seg_offset", ctypes.c_uint64),
                ("seg_bytes", ctypes.c_uint64)]


class _Hits(ctypes.Structure):
    _fields_ = [("n_reads", ctypes.c_uint64), ("substring", ctypes.POINTER(ctypes.c_uint8)),
                ("block_off", ctypes.POINTER(ctypes.c_uint64)), ("blocks", ctypes.c_void_p),
                ("n_blocks", ctypes.c_uint64), ("ms_total", ctypes.c_double), ("ms_search", ctypes.c_double),
                ("ms_extract", ctypes.c_double), ("lookups_search", ctypes.c_uint64),
                ("lookups_extract", ctypes.c_uint64), ("launches", ctypes.c_uint64),
                ("search_launches", ctypes.c_uint64), ("slow_reads", ctypes.c_uint64), ("lines_search", ctypes.c_uint64)]


class _Edges(ctypes.Structure):
    _fields_ = [("n_reads", ctypes.c_uint64), ("substring", ctypes.POINTER(ctypes.c_uint8)),
                ("n_edges", ctypes.c_uint64), ("edges", ctypes.c_void_p), ("n_blocks", ctypes.c_uint64),
                ("ms_total", ctypes.c_double), ("ms_search", ctypes.c_double), ("ms_extract", ctypes.c_double),
                ("lookups_search", ctypes.c_uint64), ("lookups_extract", ctypes.c_uint64), ("launches", ctypes.c_uint64),
                ("search_launches", ctypes.c_uint64), ("slow_reads", ctypes.c_uint64), ("lines_search", ctypes.c_uint64)]


EDGE_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("length", "<u4"), ("af", "<u4")])


class _CorrectParams(ctypes.Structure):
    _fields_ = [("kmer_size", ctypes.c_uint32), ("kmer_threshold", ctypes.c_int32), ("kmer_rounds", ctypes.c_uint32),
                ("kmer_count_offset", ctypes.c_uint64)]


class _Corrected(ctypes.Structure):
    _fields_ = [("n_reads", ctypes.c_uint64), ("valid", ctypes.POINTER(ctypes.c_uint8)), ("n_fixes", ctypes.c_uint64),
                ("fixes", ctypes.c_void_p), ("ms_total", ctypes.c_double), ("ms_kernel", ctypes.c_double),
                ("lookups", ctypes.c_uint64), ("fixed_reads", ctypes.c_uint64), ("launches", ctypes.c_uint64)]


FIX_DTYPE = np.dtype([("read", "<u4"), ("pos_code", "<u4")])

_lib = None


def load_library():
    """Load libsiga_b200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError("%s is missing: build it with `make` (or __graft_entry__.build()); "
                          "siga_b200 has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    lib.sgx_abi_version.restype = i32
    lib.sgx_init.argtypes = [i32, ctypes.POINTER(vp)]
    lib.sgx_shutdown.argtypes = [vp]
    lib.sgx_shutdown.restype = None
    lib.sgx_last_error.argtypes = [vp]
    lib.sgx_last_error.restype = ctypes.c_char_p
    lib.sgx_sm_count.argtypes = [vp]
    lib.sgx_reads_pack.argtypes = [vp, vp, vp, u64, ctypes.POINTER(vp)]
    lib.sgx_reads_pack_async.argtypes = [vp, vp, vp, u64, ctypes.POINTER(vp)]
    lib.sgx_reads_wait.argtypes = [vp, vp]
    lib.sgx_reads_free.argtypes = [vp]
    lib.sgx_reads_free.restype = None
    lib.sgx_reads_count.argtypes = [vp]
    lib.sgx_reads_count.restype = u64
    lib.sgx_reads_bases.argtypes = [vp]
    lib.sgx_reads_bases.restype = u64
    lib.sgx_index_build.argtypes = [vp, vp, i32, i32, ctypes.POINTER(_IndexOut)]
    lib.sgx_index_build_pair.argtypes = [vp, vp, i32, ctypes.POINTER(_IndexOut), ctypes.POINTER(_IndexOut)]
    lib.sgx_index_out_free.argtypes = [ctypes.POINTER(_IndexOut)]
    lib.sgx_index_out_free.restype = None
    lib.sgx_comm_unique_id.argtypes = [vp]
    lib.sgx_comm_init.argtypes = [vp, vp, i32, i32, ctypes.POINTER(vp)]
    lib.sgx_comm_init_local.argtypes = [vp, i32, ctypes.POINTER(vp)]
    lib.sgx_comm_free.argtypes = [vp]
    lib.sgx_comm_free.restype = None
    lib.sgx_comm_world.argtypes = [vp]
    lib.sgx_comm_rank.argtypes = [vp]
    lib.sgx_index_build_sharded.argtypes = [vp, vp, vp, i32, i32, ctypes.POINTER(_IndexOut)]
    lib.sgx_index_build_sharded_segments.argtypes = [vp, vp, vp, i32, i32, ctypes.POINTER(_IndexOut)]
    lib.sgx_fm_create.argtypes = [vp, vp, u64, u64, ctypes.POINTER(vp)]
    lib.sgx_fm_free.argtypes = [vp]
    lib.sgx_fm_free.restype = None
    lib.sgx_fm_attach_reads.argtypes = [vp, vp, vp, u64, vp]
    lib.sgx_fm_length.argtypes = [vp]
    lib.sgx_fm_length.restype = u64
    lib.sgx_fm_pred.argtypes = [vp, ctypes.POINTER(u64 * 5)]
    lib.sgx_fm_occ.argtypes = [vp, vp, vp, u64, vp]
    lib.sgx_fm_count.argtypes = [vp, vp, vp, vp, u64, vp]
    lib.sgx_overlap_batch.argtypes = [vp, vp, vp, vp, u64, u64, u32, i32, i32, ctypes.POINTER(_Hits)]
    lib.sgx_overlap_host.argtypes = [vp, vp, vp, vp, vp, u64, u32, i32, i32, ctypes.POINTER(_Hits)]
    lib.sgx_hits_free.argtypes = [ctypes.POINTER(_Hits)]
    lib.sgx_hits_free.restype = None
    lib.sgx_overlap_edges.argtypes = [vp, vp, vp, vp, u64, u64, u32, i32, i32, vp, ctypes.POINTER(_Edges)]
    lib.sgx_edges_gather.argtypes = [vp, vp, ctypes.POINTER(_Edges), i32, ctypes.POINTER(_Edges)]
    lib.sgx_edges_free.argtypes = [ctypes.POINTER(_Edges)]
    lib.sgx_edges_free.restype = None
    lib.sgx_correct_batch.argtypes = [vp, vp, vp, u64, u64, vp, vp, ctypes.POINTER(_CorrectParams), ctypes.POINTER(_Corrected)]
    lib.sgx_corrected_free.argtypes = [ctypes.POINTER(_Corrected)]
    lib.sgx_corrected_free.restype = None
    lib.sgx_bench_gather.argtypes = [vp, u64, u64, ctypes.POINTER(ctypes.c_double)]
    lib.sgx_flush_l2.argtypes = [vp]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class IndexResult:
    """Run bytes + read-start order of one index (what the .bwt / .sai files are written from)."""

    def __init__(self, o, copy=True):
        self.n_strings = int(o.n_strings)
        self.n_suffixes = int(o.n_suffixes)
        self.runs = np.ctypeslib.as_array(o.rle, shape=(max(int(o.n_runs), 1),))[:int(o.n_runs)]
        self.sai = np.ctypeslib.as_array(o.sai, shape=(self.n_strings,))
        if copy:  # otherwise views of the library's pinned output buffers, valid until free_index()
            self.runs, self.sai = self.runs.copy(), self.sai.copy()
        self.stats = dict(ms_total=o.ms_total, ms_sort=o.ms_sort, ms_bcr=o.ms_bcr, ms_emit=o.ms_emit,
                          launches=int(o.launches), merge_launches=int(o.merge_launches), ms_merge=o.ms_merge,
                          merge_bytes=int(o.merge_bytes))


class Hits:
    def __init__(self, h, copy=True):
        n = int(h.n_reads)
        nb = int(h.n_blocks)
        self.n_reads = n
        sub = np.ctypeslib.as_array(h.substring, shape=(max(n, 1),))[:n]
        off = np.ctypeslib.as_array(h.block_off, shape=(n + 1,))
        buf = (ctypes.c_uint8 * (max(nb, 1) * BLOCK_DTYPE.itemsize)).from_address(h.blocks)
        blocks = np.frombuffer(buf, dtype=BLOCK_DTYPE)[:nb]
        if copy:
            sub, off, blocks = sub.copy(), off.copy(), blocks.copy()
        self.substring, self.block_off, self.blocks = sub, off, blocks
        self.stats = dict(ms_total=h.ms_total, ms_search=h.ms_search, ms_extract=h.ms_extract,
                          lookups_search=int(h.lookups_search), lookups_extract=int(h.lookups_extract), lines_search=int(h.lines_search),
                          launches=int(h.launches), search_launches=int(h.search_launches), n_blocks=nb,
                          slow_reads=int(h.slow_reads))

    def read_blocks(self, i):
        """Sorted tuples (capped x4, raw x4, length, af) of read i -- the form the reference's .hits lines parse to."""
        b = self.blocks[int(self.block_off[i]):int(self.block_off[i + 1])]
        return sorted(tuple(int(x) for x in r["capped"]) + tuple(int(x) for x in r["raw"]) + (int(r["length"]), int(r["af"]))
                      for r in b)


class Edges:
    """Compact edge records (sgx_edge) + per-read SS flags of one sgx_overlap_edges call."""

    def __init__(self, e, copy=True):
        n, ne = int(e.n_reads), int(e.n_edges)
        if not e.substring or not e.edges:  # empty result (e.g. a non-root rank of sgx_edges_gather)
            n = ne = 0
            sub, edges = np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=EDGE_DTYPE)
        else:
            sub = np.ctypeslib.as_array(e.substring, shape=(max(n, 1),))[:n]
            buf = (ctypes.c_uint8 * (max(ne, 1) * EDGE_DTYPE.itemsize)).from_address(e.edges)
            edges = np.frombuffer(buf, dtype=EDGE_DTYPE)[:ne]
        if copy:
            sub, edges = sub.copy(), edges.copy()
        self.n_reads, self.substring, self.edges = n, sub, edges
        self.stats = dict(ms_total=e.ms_total, ms_search=e.ms_search, ms_extract=e.ms_extract,
                          lookups_search=int(e.lookups_search), lookups_extract=int(e.lookups_extract), lines_search=int(e.lines_search),
                          launches=int(e.launches), search_launches=int(e.search_launches),
                          slow_reads=int(e.slow_reads), n_blocks=int(e.n_blocks), n_edges=ne)


class Corrected:
    """Result of sgx_correct_batch: per-read `valid` flags (written / dropped) and the substituted bases."""

    def __init__(self, c):
        n, nf = int(c.n_reads), int(c.n_fixes)
        self.n_reads = n
        self.valid = np.ctypeslib.as_array(c.valid, shape=(max(n, 1),))[:n].copy()
        buf = (ctypes.c_uint8 * (max(nf, 1) * FIX_DTYPE.itemsize)).from_address(c.fixes)
        self.fixes = np.frombuffer(buf, dtype=FIX_DTYPE)[:nf].copy()
        self.stats = dict(ms_total=c.ms_total, ms_kernel=c.ms_kernel, lookups=int(c.lookups), n_fixes=nf,
                          fixed_reads=int(c.fixed_reads), launches=int(c.launches))

    def apply(self, flat, off, first=0):
        """Copy of the ASCII bases `flat` (layout of pack_reads) with the fixes of reads [first, first+n) applied."""
        out = np.array(flat, dtype=np.uint8, copy=True)
        if len(self.fixes):
            at = off[first + self.fixes["read"].astype(np.int64)].astype(np.int64) + (self.fixes["pos_code"] >> 2).astype(np.int64)
            out[at] = np.frombuffer(b"ACGT", dtype=np.uint8)[self.fixes["pos_code"] & 3]
        return out


def name_ranks(names):
    """Rank of every read name in std::string (byte-wise) order, equal names sharing a rank."""
    enc = [n.encode() if isinstance(n, str) else bytes(n) for n in names]
    order = sorted(set(enc))
    lut = {nm: i for i, nm in enumerate(order)}
    return np.array([lut[nm] for nm in enc], dtype=np.uint32)


class Context:
    """One sgx_ctx = one GPU."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.sgx_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise SgxError(rc, self.lib.sgx_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.sgx_shutdown(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise SgxError(rc, self.lib.sgx_last_error(self.h).decode())

    @property
    def sm_count(self):
        return int(self.lib.sgx_sm_count(self.h))

    # ---- reads
    def pack_reads(self, flat, off):
        flat = np.ascontiguousarray(flat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        r = ctypes.c_void_p()
        self._check(self.lib.sgx_reads_pack(self.h, _ptr(flat), _ptr(off), len(off) - 1, ctypes.byref(r)))
        return r

    def pack_reads_async(self, flat, off):
        """sgx_reads_pack_async: returns at once, the bases stream to the device in chunks on a second stream.  `flat` must
        stay alive (and should be pinned) until a consumer of the store has returned."""
        flat = np.ascontiguousarray(flat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        r = ctypes.c_void_p()
        self._check(self.lib.sgx_reads_pack_async(self.h, _ptr(flat), _ptr(off), len(off) - 1, ctypes.byref(r)))
        return r

    def reads_wait(self, r):
        self._check(self.lib.sgx_reads_wait(self.h, r))

    def free_reads(self, r):
        self.lib.sgx_reads_free(r)

    # ---- index
    def index_build(self, reads, order=ORDER_SAIS2, reversed=False, copy=True):
        o = _IndexOut()
        self._check(self.lib.sgx_index_build(self.h, reads, int(order), 1 if reversed else 0, ctypes.byref(o)))
        return self._index_result(o, copy)

    def index_build_pair(self, reads, order=ORDER_SAIS2, copy=True):
        """Forward and reversed index built concurrently (sgx_index_build_pair) -> (forward, reversed) results."""
        a, b = _IndexOut(), _IndexOut()
        self._check(self.lib.sgx_index_build_pair(self.h, reads, int(order), ctypes.byref(a), ctypes.byref(b)))
        return self._index_result(a, copy), self._index_result(b, copy)

    def _index_result(self, o, copy):
        res = IndexResult(o, copy=copy)
        if copy:
            self.lib.sgx_index_out_free(ctypes.byref(o))
        else:
            res._raw = o  # caller releases with free_index
        return res

    def free_index(self, res):
        if getattr(res, "_raw", None) is not None:
            self.lib.sgx_index_out_free(ctypes.byref(res._raw))
            res._raw = None
            res.runs = res.sai = None

    # ---- multi-GPU index
    def comm_unique_id(self):
        """128-byte NCCL unique id (rank 0 creates it; the launcher carries it to the other ranks)."""
        buf = np.zeros(128, dtype=np.uint8)
        rc = self.lib.sgx_comm_unique_id(_ptr(buf))
        if rc != 0:
            raise SgxError(rc, self.lib.sgx_last_error(None).decode())
        return buf

    def comm_init(self, unique_id, rank, world):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        c = ctypes.c_void_p()
        self._check(self.lib.sgx_comm_init(self.h, _ptr(uid), int(rank), int(world), ctypes.byref(c)))
        return c

    def comm_init_local(self, shards):
        """`shards` virtual shards on this context (device copies instead of NCCL)."""
        c = ctypes.c_void_p()
        self._check(self.lib.sgx_comm_init_local(self.h, int(shards), ctypes.byref(c)))
        return c

    def free_comm(self, c):
        self.lib.sgx_comm_free(c)

    def index_build_sharded(self, comm, reads, order=ORDER_SAIS2, reversed=False, copy=True):
        """Collective bucket-sharded build; `runs` is complete on rank 0 only, `sai` everywhere."""
        o = _IndexOut()
        self._check(self.lib.sgx_index_build_sharded(self.h, comm, reads, int(order), 1 if reversed else 0, ctypes.byref(o)))
        rank = int(self.lib.sgx_comm_rank(comm))
        n_runs = int(o.n_runs)
        if rank != 0:
            o.n_runs = 0
        res = self._index_result(o, copy)
        res.n_runs_total = n_runs
        return res

    def index_build_sharded_segments(self, comm, reads, order=ORDER_SAIS2, reversed=False, copy=True):
        """Collective bucket-sharded build without the final gather: `runs` holds this rank's segment of the run bytes,
        which belongs at byte `seg_offset` of the `n_runs_total` bytes of the whole index; `sai` everywhere."""
        o = _IndexOut()
        self._check(self.lib.sgx_index_build_sharded_segments(self.h, comm, reads, int(order), 1 if reversed else 0,
                                                              ctypes.byref(o)))
        n_runs, seg_offset, seg_bytes = int(o.n_runs), int(o.seg_offset), int(o.seg_bytes)
        o.n_runs = seg_bytes  # IndexResult views n_runs bytes of rle
        res = self._index_result(o, copy)
        res.n_runs_total = n_runs
        res.seg_offset = seg_offset
        return res

    # ---- FM-index
    def fm_create(self, runs, n_suffixes):
        runs = np.ascontiguousarray(runs, dtype=np.uint8)
        f = ctypes.c_void_p()
        self._check(self.lib.sgx_fm_create(self.h, _ptr(runs), len(runs), int(n_suffixes), ctypes.byref(f)))
        return f

    def fm_attach_reads(self, f, sai, reads):
        """Attach the .sai order + indexed read set (enables the read-based extraction path)."""
        sai = np.ascontiguousarray(sai, dtype=np.uint32)
        self._check(self.lib.sgx_fm_attach_reads(self.h, f, _ptr(sai), len(sai), reads))

    def free_fm(self, f):
        self.lib.sgx_fm_free(f)

    def fm_pred(self, f):
        p = (ctypes.c_uint64 * 5)()
        self._check(self.lib.sgx_fm_pred(f, ctypes.byref(p)))
        return [int(x) for x in p]

    def fm_occ(self, f, pos):
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros((len(pos), 5), dtype=np.uint64)
        self._check(self.lib.sgx_fm_occ(self.h, f, _ptr(pos), len(pos), _ptr(out)))
        return out

    def fm_count(self, f, flat, off):
        flat = np.ascontiguousarray(flat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        out = np.zeros(len(off) - 1, dtype=np.uint64)
        self._check(self.lib.sgx_fm_count(self.h, f, _ptr(flat), _ptr(off), len(off) - 1, _ptr(out)))
        return out

    # ---- overlap
    def overlap(self, fwd, rev, reads, first, count, min_overlap, irreducible=True, rc=True, copy=True):
        h = _Hits()
        self._check(self.lib.sgx_overlap_batch(self.h, fwd, rev, reads, int(first), int(count), int(min_overlap),
                                               1 if irreducible else 0, 1 if rc else 0, ctypes.byref(h)))
        res = Hits(h, copy=copy)
        if copy:
            self.lib.sgx_hits_free(ctypes.byref(h))
        else:
            res._raw = h  # caller releases with free_hits
        return res

    def overlap_host(self, fwd, rev, flat, off, min_overlap, irreducible=True, rc=True, copy=True):
        flat = np.ascontiguousarray(flat, dtype=np.uint8) if not isinstance(flat, int) else flat
        off = np.ascontiguousarray(off, dtype=np.uint64)
        h = _Hits()
        p = ctypes.c_void_p(flat) if isinstance(flat, int) else _ptr(flat)
        self._check(self.lib.sgx_overlap_host(self.h, fwd, rev, p, _ptr(off), len(off) - 1, int(min_overlap),
                                              1 if irreducible else 0, 1 if rc else 0, ctypes.byref(h)))
        res = Hits(h, copy=copy)
        if copy:
            self.lib.sgx_hits_free(ctypes.byref(h))
        else:
            res._raw = h
        return res

    def overlap_edges(self, fwd, rev, reads, first, count, min_overlap, name_rank, irreducible=True, rc=True, copy=True):
        name_rank = np.ascontiguousarray(name_rank, dtype=np.uint32)
        e = _Edges()
        self._check(self.lib.sgx_overlap_edges(self.h, fwd, rev, reads, int(first), int(count), int(min_overlap),
                                               1 if irreducible else 0, 1 if rc else 0, _ptr(name_rank), ctypes.byref(e)))
        res = Edges(e, copy=copy)
        if copy:
            self.lib.sgx_edges_free(ctypes.byref(e))
        else:
            res._raw = e
        return res

    def edges_gather(self, comm, edges, root=0, copy=True):
        """Collective: every rank passes the Edges of its read range (obtained with copy=False); `root` gets the
        concatenation (SS flags in input order), the other ranks an empty result."""
        assert getattr(edges, "_raw", None) is not None, "edges_gather needs the raw result (overlap_edges(..., copy=False))"
        out = _Edges()
        self._check(self.lib.sgx_edges_gather(self.h, comm, ctypes.byref(edges._raw), int(root), ctypes.byref(out)))
        res = Edges(out, copy=copy)
        if copy:
            self.lib.sgx_edges_free(ctypes.byref(out))
        else:
            res._raw = out
        return res

    # ---- k-mer error correction
    def correct(self, fm, reads, first, count, qual=None, qual_off=None, kmer_size=31, kmer_threshold=3, kmer_rounds=10,
                kmer_count_offset=1):
        """KmerCorrector::process for reads [first, first+count); qual/qual_off = the quality strings of those reads."""
        prm = _CorrectParams(int(kmer_size), int(kmer_threshold), int(kmer_rounds), int(kmer_count_offset))
        c = _Corrected()
        qp = qo = None
        if qual is not None:
            qual = np.ascontiguousarray(qual, dtype=np.uint8)
            qual_off = np.ascontiguousarray(qual_off, dtype=np.uint64)
            qp, qo = _ptr(qual), _ptr(qual_off)
        self._check(self.lib.sgx_correct_batch(self.h, fm, reads, int(first), int(count), qp, qo, ctypes.byref(prm),
                                               ctypes.byref(c)))
        res = Corrected(c)
        self.lib.sgx_corrected_free(ctypes.byref(c))
        return res

    def free_edges(self, edges):
        if getattr(edges, "_raw", None) is not None:
            self.lib.sgx_edges_free(ctypes.byref(edges._raw))
            edges._raw = None

    def free_hits(self, hits):
        if getattr(hits, "_raw", None) is not None:
            self.lib.sgx_hits_free(ctypes.byref(hits._raw))
            hits._raw = None

    # ---- measurement helpers
    def bench_gather(self, nbytes, n_loads):
        out = ctypes.c_double()
        self._check(self.lib.sgx_bench_gather(self.h, int(nbytes), int(n_loads), ctypes.byref(out)))
        return out.value

    def flush_l2(self):
        self._check(self.lib.sgx_flush_l2(self.h))
