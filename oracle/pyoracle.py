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
