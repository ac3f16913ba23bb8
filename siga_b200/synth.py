"""Seeded synthetic read sets of the shapes BASELINE.json names.

Python-3 restatement of the *shape* of the reference's generator benchmark/paired_read_gen.py:25-61 (which is
Python 2, unseeded and needs a reference FASTA): uniform random genome, fixed-length reads, FASTA with one
sequence line per read, names r<idx>, alphabet strictly ACGT.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b

# name -> (genome bp, read length, coverage, substitution error rate, both strands, seed)
CONFIGS = {
    "cfg1": dict(genome=1_000_000, read_len=100, n_reads=20_000, err=0.0, seed=101, paired=True),
    "cfg2": dict(genome=4_600_000, read_len=150, coverage=100, err=0.01, seed=202),
    "cfg3": dict(genome=4_600_000, read_len=150, coverage=100, err=0.01, seed=202),  # same reads, min-overlap sweep
    "cfg4": dict(genome=100_000_000, read_len=150, coverage=50, err=0.0, seed=404),
    "cfg5": dict(genome=64_000_000, read_len=250, coverage=40, err=0.0, seed=505),
}


def random_genome(n, rng):
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(arr):
    """Reverse complement of a 2-D uint8 array of reads (rows) or a 1-D read."""
    return _COMP[arr][..., ::-1]


def sample_reads(genome, n_reads, read_len, rng, err=0.0, both_strands=True, chunk=262144):
    """n_reads x read_len uint8 matrix of ASCII bases: uniform starts, strand uniform, iid substitutions.

    Generated in fixed-size chunks (bounded memory at config-2..5 sizes); the chunk size is part of the recipe."""
    g = len(genome)
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    ar = np.arange(read_len, dtype=np.int64)[None, :]
    for s in range(0, n_reads, chunk):
        e = min(n_reads, s + chunk)
        starts = rng.integers(0, g - read_len + 1, size=e - s)
        reads = genome[starts[:, None] + ar]
        if err > 0:
            mask = rng.random(reads.shape, dtype=np.float32) < err
            k = int(mask.sum())
            if k:
                # substitute with one of the three other bases, uniformly
                code = np.searchsorted(_ACGT, reads[mask])
                code = (code + rng.integers(1, 4, size=k)) & 3
                reads[mask] = _ACGT[code]
        if both_strands:
            flip = rng.random(e - s) < 0.5
            reads[flip] = revcomp(reads[flip])
        out[s:e] = reads
    return out


def config_reads(name, scale=1.0):
    """Read matrix of a named config; `scale` shrinks the genome (and read count) keeping coverage and length."""
    c = CONFIGS[name]
    rng = np.random.default_rng(c["seed"])
    g = max(int(c["genome"] * scale), 4 * c["read_len"])
    genome = random_genome(g, rng)
    if c.get("paired"):
        return paired_reads(genome, int(c["n_reads"] * scale) // 2, c["read_len"], rng)
    n = int(g * c["coverage"] // c["read_len"])
    return sample_reads(genome, n, c["read_len"], rng, err=c["err"], both_strands=True)


def paired_reads(genome, n_pairs, read_len, rng, insert=1000):
    """Config-1 shape (paired_read_gen.py:33-54): forward-strand pairs, mate 1 = ref[pos:pos+K], mate 2 =
    ref[pos+K+insert : pos+2K+insert], pos uniform in [0, G-2K-insert] without repetition (so pairs are unique;
    single reads may still coincide across pairs, exactly as with the reference generator).
    Mate 1 of pair i is read i, mate 2 is read i + n_pairs."""
    g = len(genome)
    starts = rng.choice(g - 2 * read_len - insert + 1, size=n_pairs, replace=False)
    ar = np.arange(read_len)[None, :]
    m1 = genome[starts[:, None] + ar]
    m2 = genome[(starts + read_len + insert)[:, None] + ar]
    return np.ascontiguousarray(np.concatenate([m1, m2], axis=0))


def reads_to_flat(reads):
    """(ascii bytes, uint64 offsets) in the layout sgx_reads_pack takes."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, l = reads.shape
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(l)).astype(np.uint64)
        return np.ascontiguousarray(reads).reshape(-1), off
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    flat = np.frombuffer(b"".join(bytes(r) if not isinstance(r, str) else r.encode() for r in reads), dtype=np.uint8)
    return flat, off


def write_fasta(path, reads, names=None, prefix="r"):
    """One sequence line per read, names r<idx> by default."""
    with open(path, "wb") as f:
        if isinstance(reads, np.ndarray) and reads.ndim == 2 and names is None:
            n, l = reads.shape
            # vectorised: b'>r123\nACGT...\n'
            chunk = 200_000
            for s in range(0, n, chunk):
                e = min(n, s + chunk)
                parts = []
                for i in range(s, e):
                    parts.append(b">%s%d\n" % (prefix.encode(), i))
                    parts.append(reads[i].tobytes())
                    parts.append(b"\n")
                f.write(b"".join(parts))
        else:
            for i, r in enumerate(reads):
                nm = names[i] if names is not None else "%s%d" % (prefix, i)
                seq = r if isinstance(r, (bytes, bytearray)) else (r.encode() if isinstance(r, str) else bytes(r))
                f.write(b">" + nm.encode() + b"\n" + seq + b"\n")


def default_name_ranks(n, prefix="r"):
    """Rank of the default names r0..r(n-1) in std::string order (equal to api.name_ranks of the list).

    Byte-wise order of decimal strings = order of the numbers left-aligned to the longest width, a shorter string
    before the longer ones it is a prefix of ("r1" < "r10" < "r2"): one integer key per name, no string sort."""
    if n <= 0:
        return np.zeros(0, dtype=np.uint32)
    x = np.arange(n, dtype=np.uint64)
    width = len(str(n - 1))
    digits = np.ones(n, dtype=np.uint64)
    p = np.uint64(10)
    for d in range(2, width + 1):
        digits[x >= p] = d
        p = p * np.uint64(10)
    pow10 = np.array([10 ** k for k in range(width + 1)], dtype=np.uint64)
    key = x * pow10[(np.uint64(width) - digits).astype(np.int64)] * np.uint64(width + 1) + digits
    order = np.argsort(key, kind="stable")
    rank = np.empty(n, dtype=np.uint32)
    rank[order] = np.arange(n, dtype=np.uint32)
    return rank
