"""Interval helpers of the mask rows, mirroring the behaviour of
reference src/polyoligo/lib_utils.py:144-164 (list_2_ranges) and :233-261
(reduce_ivs), written over numpy arrays."""
import numpy as np


def list_2_ranges(ixs):
    """Sorted runs of consecutive indices -> [[start, length], ...]."""
    ixs = np.sort(np.asarray(ixs, dtype=np.int64))
    if ixs.size == 0:
        return []
    breaks = np.flatnonzero(np.diff(ixs) != 1) + 1      # a duplicate index also breaks a run, as upstream
    starts = np.concatenate([[0], breaks])
    stops = np.concatenate([breaks, [ixs.size]])
    return [[int(ixs[s]), int(e - s)] for s, e in zip(starts, stops)]


def reduce_ivs(ivs, n):
    """Merge the two closest neighbouring intervals until at most n remain
    (primer3 rejects more than 200 excluded regions)."""
    ivs = [list(iv) for iv in ivs]
    if len(ivs) <= n or not ivs:
        return ivs
    gaps = [ivs[k + 1][0] - (ivs[k][0] + ivs[k][1] - 1) - 1 for k in range(len(ivs) - 1)]
    while len(ivs) > n:
        k = int(np.argmin(gaps))          # first minimum, like np.argmin on the reference's list
        gaps.pop(k)
        left, right = ivs[k], ivs.pop(k + 1)
        ivs[k] = [left[0], right[0] + right[1] - left[0]]
    return ivs


def round_tidy(x, n):
    """x rounded to n significant digits (reference lib_utils.py:21-27)."""
    import sys
    y = abs(x)
    if y <= sys.float_info.min:
        return 0.0
    return np.round(x, int(n - np.ceil(np.log10(y))))


# Biopython IUPACData.ambiguous_dna_values, inverted as the reference does (lib_utils.py:15-18)
_IUPAC_DNA = {"A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT",
              "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC"}
IUPAC_DNA_I = {v: k for k, v in _IUPAC_DNA.items()}


def seqs2ambiguous_dna(seqs):
    """Same-length sequences -> one IUPAC-ambiguous sequence (reference lib_utils.py:216-230);
    a column whose letter set has no code (e.g. all four bases, whose table key is 'GATC') is N."""
    out = []
    for column in zip(*seqs):
        out.append(IUPAC_DNA_I.get("".join(sorted(set(column))), "N"))
    return "".join(out)


_COMP = str.maketrans("ACGTUacgtuRYKMBVDHrykmbvdhNnSsWw", "TGCAAtgcaaYRMKVBHDyrmkvbhdNnSsWw")


def reverse_complement(seq):
    """Bio.Seq.reverse_complement for (ambiguous) DNA strings."""
    return seq[::-1].translate(_COMP)
