"""TEST INFRASTRUCTURE ONLY: brute-force Python restatement of the reference's two suffix-ordering models.

Pure-Python sorting, so only for small cases (a few thousand symbols).  Follows SURVEY.md Appendix B, which was
validated against the compiled reference:
  model B (`sais2`, src/suffix_array_builder.cpp:474-673 + 3rdparty/sais.hxx:552): plain suffix array of
          T = R0 $ R1 $ ... R(n-1) $ with every '$' = 0 and end-of-text smallest;
  model A (`sais`,  src/suffix_array_builder.cpp:31-282): suffixes ordered by (R_i[j:] $, i).
BWT symbol of a row = the symbol before the suffix, '$' at a read start (src/suffix_array_builder.cpp:152-160, 548-556).
Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import numpy as np

_RANK = {ord("A"): 1, ord("C"): 2, ord("G"): 3, ord("T"): 4}


def _as_bytes_list(reads):
    out = []
    for r in reads:
        if isinstance(r, str):
            r = r.encode()
        out.append(bytes(r))
    return out


def bwt_and_sai(reads, model="B", reversed_reads=False):
    """-> (bwt symbol ranks uint8[n_suffixes], sai read ids uint32[n_reads])"""
    reads = _as_bytes_list(reads)
    if reversed_reads:
        reads = [r[::-1] for r in reads]
    ranks = [bytes(_RANK[c] for c in r) for r in reads]
    if model == "B":
        text = b"".join(r + b"\0" for r in ranks)
        starts = {}
        p = 0
        for i, r in enumerate(ranks):
            starts[p] = i
            p += len(r) + 1
        order = sorted(range(len(text)), key=lambda q: text[q:])
        bwt = np.array([0 if (q == 0 or text[q - 1] == 0) else text[q - 1] for q in order], dtype=np.uint8)
        sai = np.array([starts[q] for q in order if q in starts], dtype=np.uint32)
        return bwt, sai
    items = []
    for i, r in enumerate(ranks):
        for j in range(len(r) + 1):
            items.append((r[j:] + b"\0", i, j))
    items.sort(key=lambda t: (t[0], t[1]))
    bwt = np.array([0 if j == 0 else ranks[i][j - 1] for _, i, j in items], dtype=np.uint8)
    sai = np.array([i for _, i, j in items if j == 0], dtype=np.uint32)
    return bwt, sai


def occ_table(bwt):
    """Inclusive cumulative counts: occ[i, c] = #c in bwt[0..i]."""
    onehot = (np.asarray(bwt)[:, None] == np.arange(5)[None, :]).astype(np.uint64)
    return np.cumsum(onehot, axis=0)


# ---------------------------------------------------------------------------------------------------------------------
# Overlap models (validated there against the compiled reference on random fixed-length sets).
# Brute force over all read pairs: small inputs only.
# ---------------------------------------------------------------------------------------------------------------------

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _rc(s):
    return s.translate(_COMP)[::-1]


def _candidates(reads, names, m):
    """Per query: containment edges and the four candidate lists of OverlapBuilder::overlap
    (src/overlap_builder.cpp:1113-1158) for FIXED-LENGTH reads: (orientation, target index, k), largest k only."""
    out = []
    for qi, Q in enumerate(reads):
        rq = _rc(Q)
        lists = {"SP": [], "PP": [], "PS": [], "SS": []}
        contain = []
        for ti, T in enumerate(reads):
            if names[ti] == names[qi]:
                continue  # self hits are dropped by name (src/overlap_builder.cpp:358)
            if T == Q:
                contain.append((ti, 0))
                continue  # its length-|Q| block makes every shorter overlap with T sub-maximal (1137-1158)
            if _rc(T) == Q:
                contain.append((ti, 1))
                continue
            hi = min(len(Q), len(T)) - 1
            for name, a, b in (("SP", Q, T), ("PP", rq, T)):
                for k in range(hi, m - 1, -1):
                    if a[len(a) - k:] == b[:k]:
                        lists[name].append((ti, k))
                        break
            for name, a, b in (("PS", Q, T), ("SS", rq, T)):
                for k in range(hi, m - 1, -1):
                    if a[:k] == b[len(b) - k:]:
                        lists[name].append((ti, k))
                        break
        out.append((lists, contain))
    return out


def _ed(names, lens, qi, ti, k, kind):
    """ED payload of OverlapBlock::overlap + Overlap's writer (src/overlap_builder.cpp:158-175, src/coord.cpp:66-80)."""
    lq, lt = lens[qi], lens[ti]
    if kind == "SP":
        c = (lq - k, lq - 1, lq, 0, k - 1, lt, 0)
    elif kind == "PP":
        c = (0, k - 1, lq, 0, k - 1, lt, 1)
    elif kind == "PS":
        c = (0, k - 1, lq, lt - k, lt - 1, lt, 0)
    else:
        c = (lq - k, lq - 1, lq, lt - k, lt - 1, lt, 1)
    return "%s %s %d %d %d %d %d %d %d 0" % ((names[qi], names[ti]) + c)


def _irreducible(cands, reads):
    """IrreducibleBlockListExtractor::extract (src/overlap_builder.cpp:711-809) on explicit extension strings:
    cands = [(kind, target, k, ext)]; groups split by the next extension symbol; a group ends when a top-length
    candidate has no extension left, emitting the top-length candidates."""
    keep = []
    groups = [(list(cands), 0)]
    while groups:
        alive, p = groups.pop()
        if not alive:
            continue
        top = max(c[2] for c in alive)
        tl = [c for c in alive if c[2] == top]
        if any(len(c[3]) == p for c in tl):
            keep.extend(c for c in tl if len(c[3]) == p)
            continue
        by = {}
        for c in alive:
            if len(c[3]) > p:
                by.setdefault(c[3][p], []).append(c)
        for sym in by:
            groups.append((by[sym], p + 1))
    return keep


def overlap_edges(reads, names, m, exhaustive=False):
    """Sorted ED payloads `siga overlap -m m [-x]` writes for FIXED-LENGTH reads."""
    reads = _as_bytes_list(reads)
    lens = [len(r) for r in reads]
    edges = []
    for qi, (lists, contain) in enumerate(_candidates(reads, names, m)):
        for ti, is_rc in contain:
            if names[qi] > names[ti]:
                l = lens[qi]
                edges.append("%s %s 0 %d %d 0 %d %d %d 0" % (names[qi], names[ti], l - 1, l, l - 1, l, is_rc))
        if exhaustive:
            chosen = [(kind, ti, k) for kind in ("SP", "SS", "PP", "PS") for ti, k in lists[kind]]
        else:
            right, left = [], []
            for kind in ("SP", "SS"):
                for ti, k in lists[kind]:
                    t = reads[ti] if kind == "SP" else _rc(reads[ti])
                    right.append((kind, ti, k, t[k:]))
            for kind in ("PS", "PP"):
                for ti, k in lists[kind]:
                    t = reads[ti] if kind == "PS" else _rc(reads[ti])
                    left.append((kind, ti, k, t[:len(t) - k][::-1]))
            chosen = [(c[0], c[1], c[2]) for c in _irreducible(right, reads) + _irreducible(left, reads)]
        for kind, ti, k in chosen:
            if names[qi] > names[ti]:
                edges.append(_ed(names, lens, qi, ti, k, kind))
    return sorted(edges)
