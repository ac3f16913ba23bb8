"""SGA/SIGA on-disk formats of the hot path: readers, writers and the hits -> ED step.

Host-side mirror in Python of what the reference's writers/readers do; used by the tests and bench.  The product
CLI (siga_b200/csrc/host) implements the same formats in C++.
"""
import gzip
import struct

import numpy as np

BWT_MAGIC = 0xCACA
AF_QUERYREV, AF_TARGETREV, AF_QUERYCOMP = 1, 2, 4


def read_bwt(path):
    """-> (n_strings, n_suffixes, runs uint8 array).  Layout: src/bwt.cpp:34-39, 121-178."""
    with open(path, "rb") as f:
        data = f.read()
    magic, n_strings, n_suffixes, n_runs, flag = struct.unpack_from("<HQQQi", data, 0)
    if magic != BWT_MAGIC:
        raise ValueError("bad .bwt magic")
    runs = np.frombuffer(data, dtype=np.uint8, offset=30, count=n_runs).copy()
    return n_strings, n_suffixes, runs


def bwt_bytes(n_strings, n_suffixes, runs):
    """The exact byte image BWTWriter produces (src/bwt.cpp:121-178)."""
    runs = np.ascontiguousarray(runs, dtype=np.uint8)
    return struct.pack("<HQQQi", BWT_MAGIC, n_strings, n_suffixes, len(runs), 0) + runs.tobytes()


def decode_runs(runs):
    """Run bytes -> symbol ranks (uint8, 0..4)."""
    runs = np.asarray(runs, dtype=np.uint8)
    return np.repeat(runs >> 5, runs & 31).astype(np.uint8)


def encode_runs(sym):
    """Greedy run-length coding with runs capped at 31 (src/bwt.cpp:16-27, src/rlstring.h:10-13)."""
    sym = np.asarray(sym, dtype=np.uint8)
    if len(sym) == 0:
        return np.zeros(0, dtype=np.uint8)
    starts = np.flatnonzero(np.concatenate(([True], sym[1:] != sym[:-1])))
    lens = np.diff(np.concatenate((starts, [len(sym)])))
    full, rem = lens // 31, lens % 31
    per = full + (rem > 0)
    out = np.empty(int(per.sum()), dtype=np.uint8)
    pos = np.concatenate(([0], np.cumsum(per)))[:-1]
    s = sym[starts].astype(np.uint8) << 5
    idx = np.repeat(np.arange(len(starts)), per)
    k = np.arange(len(out)) - np.repeat(pos, per)
    cnt = np.where(k < full[idx], 31, rem[idx]).astype(np.uint8)
    out[:] = s[idx] | cnt
    return out


def read_sai(path):
    """-> uint32 array of read ids (the `<i> 0` lines).  Layout: src/suffix_array.cpp:16-43."""
    with open(path, "rb") as f:
        tok = f.read().split()
    if int(tok[0]) != BWT_MAGIC:
        raise ValueError("bad .sai magic")
    n = int(tok[1])
    ids = np.array(tok[3::2], dtype=np.uint64).astype(np.uint32)
    assert len(ids) == n
    return ids


def sai_bytes(ids):
    """The exact text image SAWriter produces: magic, n, n, then `<i> 0` per read (src/suffix_array.cpp:16-43)."""
    ids = np.asarray(ids)
    head = b"%d\n%d\n%d\n" % (BWT_MAGIC, len(ids), len(ids))
    body = b"".join(b"%d 0\n" % int(i) for i in ids)
    return head + body


def read_hits(paths):
    """Parse reference `<prefix>-thread<t>.hits.gz` files (src/overlap_builder.cpp:138-141,198-201,234-254).

    -> dict read_idx -> (substring flag, list of (c0lo,c0hi,c1lo,c1hi,r0lo,r0hi,r1lo,r1hi,len,af))"""
    out = {}
    for p in paths:
        with gzip.open(p, "rt") as f:
            for line in f:
                t = line.split()
                if not t:
                    continue
                idx, sub, nb = int(t[0]), int(t[1]), int(t[2])
                blocks = []
                k = 3
                for _ in range(nb):
                    v = [int(x) for x in t[k:k + 9]]
                    af = int(t[k + 9], 2)  # bitset<3> prints QC TR QR (MSB first)
                    blocks.append(tuple(v) + (af,))
                    k += 10
                out[idx] = (sub, blocks)
    return out


def hits_to_edges(names, lengths, substring, block_off, blocks, sai, rsai):
    """Hit2OverlapConverter::convert + OverlapBlock::overlap (src/overlap_builder.cpp:158-175, 345-375).

    blocks: structured array with fields capped(4), raw(4), length, af.  -> list of ED payload strings."""
    edges = []
    for q in range(len(names)):
        qn, ql = names[q], int(lengths[q])
        for b in blocks[int(block_off[q]):int(block_off[q + 1])]:
            ln, af = int(b["length"]), int(b["af"])
            lo, hi = int(b["capped"][0]), int(b["capped"][1])
            sa = rsai if af & AF_TARGETREV else sai
            for j in range(lo, hi + 1):
                t = int(sa[j])
                tn, tl = names[t], int(lengths[t])
                if qn == tn:
                    continue
                s0, e0 = ql - ln, ql - 1
                s1, e1 = 0, ln - 1
                if af & AF_QUERYREV:
                    s0, e0 = ql - 1 - e0, ql - 1 - s0
                if af & AF_TARGETREV:
                    s1, e1 = tl - 1 - e1, tl - 1 - s1
                contained = (s0 == 0 and e0 + 1 == ql) or (s1 == 0 and e1 + 1 == tl)
                if qn < tn or (contained and (af & AF_QUERYREV)):
                    continue
                edges.append("%s %s %d %d %d %d %d %d %d 0" % (qn, tn, s0, e0, ql, s1, e1, tl, 1 if af & AF_QUERYCOMP else 0))
    return edges


def rmdup_records(names, seqs, substring, block_off, blocks, sai, rsai):
    """`siga rmdup` from hit blocks (OverlapBuilder::duplicate + Hits2FastaConverter::convert,
    src/overlap_builder.cpp:1184-1195, 563-616): the blocks of a read searched with min-overlap = |read| are its
    containment blocks; duplicate() keeps those of the seq / forward-index and complement / reverse-index searches
    (AlignFlags 0 and TARGETREV|QUERYCOMP).  NumDuplicates = number of ranks in their capped[0] intervals (the read
    itself included); the read is a duplicate when it is a substring of another read or an identical / reverse-
    complement read has a smaller name.  -> (kept FASTA records, duplicate FASTA records), each ">header\nseq\n"."""
    kept, dups = [], []
    for q in range(len(names)):
        qn = names[q]
        copies, contained = 0, bool(substring[q])
        for b in blocks[int(block_off[q]):int(block_off[q + 1])]:
            af = int(b["af"])
            if af not in (0, AF_TARGETREV | AF_QUERYCOMP):
                continue
            lo, hi = int(b["capped"][0]), int(b["capped"][1])
            sa = rsai if af & AF_TARGETREV else sai
            copies += hi - lo + 1
            for j in range(lo, hi + 1):
                tn = names[int(sa[j])]
                if tn != qn and not qn < tn:  # kept overlap (id0 >= id1): mutually contained, containedIdx() == 0
                    contained = True
        meta = "%s NumDuplicates=%d" % (qn, copies)
        if contained:
            dups.append(">%s,seqrank=%d %s\n%s\n" % (qn, q, meta, seqs[q]))
        else:
            kept.append(">%s %s\n%s\n" % (qn, meta, seqs[q]))
    return kept, dups


def edges_to_ed(names, lengths, edges):
    """ED payload strings of compact edge records (sgx_edge): the coordinates of OverlapBlock::overlap
    (src/overlap_builder.cpp:158-175) printed like Overlap's writer (src/coord.cpp:66-80)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    q, t = edges["query"].astype(np.int64), edges["target"].astype(np.int64)
    k, af = edges["length"].astype(np.int64), edges["af"].astype(np.int64)
    lq, lt = lengths[q], lengths[t]
    s0, e0 = lq - k, lq - 1
    qr = (af & AF_QUERYREV) != 0
    s0, e0 = np.where(qr, lq - 1 - e0, s0), np.where(qr, lq - 1 - (lq - k), e0)
    s1, e1 = np.zeros_like(k), k - 1
    tr = (af & AF_TARGETREV) != 0
    s1, e1 = np.where(tr, lt - 1 - e1, s1), np.where(tr, lt - 1, e1)
    rc = ((af & AF_QUERYCOMP) != 0).astype(np.int64)
    return ["%s %s %d %d %d %d %d %d %d 0" % (names[a], names[b], c, d, e, f, g, h, i)
            for a, b, c, d, e, f, g, h, i in zip(q.tolist(), t.tolist(), s0.tolist(), e0.tolist(), lq.tolist(),
                                                 s1.tolist(), e1.tolist(), lt.tolist(), rc.tolist())]


def edge_rows(edges, lengths):
    """int64 columns (query, target, s1, e1, s2, e2, rc) of compact edge records: the numbers of the ED line
    (OverlapBlock::overlap, src/overlap_builder.cpp:158-175).  `lengths`: per-read array or one common length."""
    lengths = np.asarray(lengths, dtype=np.int64)
    q, t = edges["query"].astype(np.int64), edges["target"].astype(np.int64)
    k, af = edges["length"].astype(np.int64), edges["af"].astype(np.int64)
    lq, lt = (lengths[q], lengths[t]) if lengths.ndim else (np.full_like(q, lengths), np.full_like(q, lengths))
    qr = (af & AF_QUERYREV) != 0
    tr = (af & AF_TARGETREV) != 0
    return (q, t, np.where(qr, 0, lq - k), np.where(qr, k - 1, lq - 1), np.where(tr, lt - k, 0), np.where(tr, lt - 1, k - 1),
            ((af & AF_QUERYCOMP) != 0).astype(np.int64))


def edge_digest_rows(q, t, s1, e1, s2, e2, rc):
    """sha256 over the ED set as int32 rows (query index, target index, s1, e1, s2, e2, rc) sorted lexicographically --
    the canonical form full-size parity compares (the reference emits ED records in thread order)."""
    import hashlib
    cols = [np.asarray(x, dtype=np.int64) for x in (q, t, s1, e1, s2, e2, rc)]
    if all(len(c) == 0 or (c.min() >= 0 and c.max() < (1 << 15)) for c in cols[2:]) and \
            all(len(c) == 0 or (c.min() >= 0 and c.max() < (1 << 31)) for c in cols[:2]):
        k1 = (cols[0] << 31) | cols[1]  # same order as the seven-key sort, two packed keys
        k2 = (cols[2] << 46) | (cols[3] << 31) | (cols[4] << 16) | (cols[5] << 1) | cols[6]
        order = np.lexsort((k2, k1))
    else:
        order = np.lexsort(tuple(cols[::-1]))
    rows = np.stack(cols, axis=1)[order].astype("<i4")
    return hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest()


def edge_digest(edges, lengths):
    return edge_digest_rows(*edge_rows(edges, lengths))


def asqg_lines(names, seqs, substring, edges, min_overlap):
    """HT / VT / ED lines exactly as OverlapBuilder::build emits them (src/overlap_builder.cpp:428-437, 301-322,
    410-413; src/asqg.cpp:93-124, 171-186, 228-237)."""
    out = ["HT\tVN:i:1\tOL:i:%d\tCN:i:1" % min_overlap]
    for n, s, ss in zip(names, seqs, substring):
        out.append("VT\t%s\t%s\tSS:i:%d" % (n, s, 1 if ss else 0))
    out.extend("ED\t" + e for e in edges)
    return out


def read_asqg(path):
    with gzip.open(path, "rt") as f:
        return [l.rstrip("\n") for l in f]
