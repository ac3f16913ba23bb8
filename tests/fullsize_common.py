"""Canonical digests of `.asqg.gz` edge sets and of sgx_overlap_edges records -- test infrastructure only.

The reference writes ED records in thread order (src/overlap_builder.cpp:383-421), so full-size parity compares SETS:
every ED record becomes an int32 row (query index, target index, s1, e1, s2, e2, rc), rows are sorted
lexicographically and hashed.  Read names must be the generator's `r<idx>`.
"""
import hashlib
import io
import subprocess

import numpy as np

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siga_b200.formats import edge_digest, edge_digest_rows, edge_rows  # noqa: E402,F401  (the canonical form lives with the formats)


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
