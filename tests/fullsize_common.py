"""Canonical digests of `.asqg.gz` edge sets and of sgx_overlap_edges records -- test infrastructure only.

The reference writes ED records in thread order (src/overlap_builder.cpp:383-421), so full-size parity compares SETS:
every ED record becomes an int32 row (query index, target index, s1, e1, s2, e2, rc), rows are sorted
lexicographically and hashed.  Read names must be the generator's `r<idx>`.
"""
import hashlib
import io
import subprocess

import numpy as np

AF_QUERYREV, AF_TARGETREV, AF_QUERYCOMP = 1, 2, 4


def edge_digest_rows(q, t, s1, e1, s2, e2, rc):
    rows = np.stack([np.asarray(x, dtype=np.int64) for x in (q, t, s1, e1, s2, e2, rc)], axis=1)
    order = np.lexsort(rows.T[::-1])
    return hashlib.sha256(np.ascontiguousarray(rows[order].astype("<i4")).tobytes()).hexdigest()


def edge_rows_from_records(edges, lengths):
    """Coordinates of OverlapBlock::overlap (src/overlap_builder.cpp:158-175) from compact sgx_edge records."""
    lengths = np.asarray(lengths, dtype=np.int64)
    q, t = edges["query"].astype(np.int64), edges["target"].astype(np.int64)
    k, af = edges["length"].astype(np.int64), edges["af"].astype(np.int64)
    lq, lt = (lengths[q], lengths[t]) if lengths.ndim else (np.full_like(q, lengths), np.full_like(q, lengths))
    qr = (af & AF_QUERYREV) != 0
    tr = (af & AF_TARGETREV) != 0
    s1 = np.where(qr, 0, lq - k)
    e1 = np.where(qr, k - 1, lq - 1)
    s2 = np.where(tr, lt - k, 0)
    e2 = np.where(tr, lt - 1, k - 1)
    rc = ((af & AF_QUERYCOMP) != 0).astype(np.int64)
    return q, t, s1, e1, s2, e2, rc


def edge_digest(edges, lengths):
    return edge_digest_rows(*edge_rows_from_records(edges, lengths))


def parse_asqg_edges(path):
    """-> (q, t, s1, e1, s2, e2, rc arrays, number of VT records, number of VT records with SS:i:1)."""
    import pandas as pd
    txt = subprocess.run("zcat %s | grep '^ED' | cut -f2 | tr -d r" % path, shell=True, capture_output=True, check=True).stdout
    if txt:
        df = pd.read_csv(io.BytesIO(txt), sep=" ", header=None, dtype=np.int64)
        cols = [df[i].to_numpy() for i in range(10)]
    else:
        cols = [np.zeros(0, dtype=np.int64)] * 10
    q, t, s1, e1, _l1, s2, e2, _l2, rc, _nd = cols
    n_vt = int(subprocess.run("zcat %s | grep -c '^VT'" % path, shell=True, capture_output=True, text=True).stdout or 0)
    n_ss = int(subprocess.run("zcat %s | grep '^VT' | grep -c 'SS:i:1'" % path, shell=True, capture_output=True, text=True).stdout or 0)
    return q, t, s1, e1, s2, e2, rc, n_vt, n_ss
