"""Synthetic workloads of the BASELINE.json shapes, generated directly as packed
oligo records (no Python strings on the fast path).

config 4: N pairs (A_i, B_i); each length iid uniform 18..30;
bases iid uniform ACGT; numpy.random.Generator(PCG64(20240607)), lengths drawn
before bases.
"""
import numpy as np

from ._native import OLIGO_DTYPE

CONFIG4_SEED = 20240607
LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def pack_codes(codes, lens):
    """codes: (n, L<=60) uint8 array of 0..3; lens: (n,) -> OLIGO_DTYPE records."""
    n, width = codes.shape
    assert width <= 60
    full = np.zeros((n, 64), dtype=np.uint32)
    mask = np.arange(width)[None, :] < lens[:, None]
    full[:, :width] = np.where(mask, codes, 0)
    shifts = (2 * (np.arange(64) % 16)).astype(np.uint32)
    words = (full << shifts[None, :]).reshape(n, 4, 16)
    w = np.bitwise_or.reduce(words, axis=2).astype(np.uint32)
    w[:, 3] |= lens.astype(np.uint32) << np.uint32(24)
    out = np.zeros(n, dtype=OLIGO_DTYPE)
    out["w"] = w
    return out


def random_oligos(rng, n, lo=18, hi=30, chunk=1 << 20):
    """n random oligos with lengths uniform in [lo, hi]; returns (records, lens)."""
    lens = rng.integers(lo, hi + 1, n)
    out = np.zeros(n, dtype=OLIGO_DTYPE)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        codes = rng.integers(0, 4, (e - s, hi), dtype=np.uint8)
        out[s:e] = pack_codes(codes, lens[s:e])
    return out, lens


def config4(n, seed=CONFIG4_SEED):
    """BASELINE config 4 pair batch: (A records, B records)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    la = rng.integers(18, 31, n)
    lb = rng.integers(18, 31, n)
    a = np.zeros(n, dtype=OLIGO_DTYPE)
    b = np.zeros(n, dtype=OLIGO_DTYPE)
    chunk = 1 << 20
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        a[s:e] = pack_codes(rng.integers(0, 4, (e - s, 30), dtype=np.uint8), la[s:e])
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        b[s:e] = pack_codes(rng.integers(0, 4, (e - s, 30), dtype=np.uint8), lb[s:e])
    return a, b


def to_strings(records):
    """Packed records -> list[str] (host only; used to feed the CPU oracle)."""
    w = records["w"]
    lens = (w[:, 3] >> 24) & 0x3F
    out = []
    for k in range(len(records)):
        codes = [(int(w[k, p >> 4]) >> (2 * (p & 15))) & 3 for p in range(int(lens[k]))]
        out.append(LETTERS[codes].tobytes().decode("ascii"))
    return out
