"""Shared test helpers: seeded synthetic oligo sets of the BASELINE.json shapes."""
import numpy as np

VIC = "GAAGGTCGGAGTCAACGGATT"   # reference src/polyoligo/cli_kasp.py:65 default reporter 1
FAM = "GAAGGTGACCAAGTTCATGCT"   # reference src/polyoligo/cli_kasp.py:72 default reporter 2


def rand_seqs(rng, n, lo, hi, alphabet="ACGT"):
    lens = rng.integers(lo, hi + 1, n)
    letters = np.array(list(alphabet))
    return ["".join(letters[rng.integers(0, len(letters), l)]) for l in lens]


def config4_pairs(n, seed=20240607):
    """BASELINE config 4: iid uniform 18..30 nt oligo pairs (lengths drawn first, then bases)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    la = rng.integers(18, 31, n)
    lb = rng.integers(18, 31, n)
    letters = np.array(list("ACGT"))
    a = ["".join(letters[rng.integers(0, 4, l)]) for l in la]
    b = ["".join(letters[rng.integers(0, 4, l)]) for l in lb]
    return a, b


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def adversarial_pairs(rng):
    """Edge cases the domain offers: low complexity, perfect duplexes, palindromes,
    extreme lengths, reporter-tailed KASP primers."""
    a, b = [], []
    for s in ["A" * 60, "AT" * 30, "GC" * 30, "ACGT" * 15, "GCGCGCGCGCGCGCGCGCGC", "A", "AC", "ACG",
              "GGGGGGGGGGGGGGGGGGGG", "CCCCCATCCGATCAGGGGG", "GTAAAACGACGGCCAGT"]:
        a.append(s)
        b.append(revcomp(s))
        a.append(s)
        b.append(s)
    x = rand_seqs(rng, 300, 1, 60)
    a += x
    b += [revcomp(s) for s in x]                       # perfect duplexes
    a += rand_seqs(rng, 300, 2, 60, "AT")
    b += rand_seqs(rng, 300, 2, 60, "AT")               # every cell of the DP table finite
    a += rand_seqs(rng, 300, 40, 60, "GC")
    b += rand_seqs(rng, 300, 40, 60, "GC")
    a += [VIC + s for s in rand_seqs(rng, 200, 18, 26)] + [FAM + s for s in rand_seqs(rng, 200, 18, 26)]
    b += rand_seqs(rng, 400, 18, 26)                    # tailed KASP shape (<= 47 x 26)
    return a, b
