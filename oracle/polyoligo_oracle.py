"""CPU ORACLE (test infrastructure, NOT product code) for the rows whose reference
arithmetic is Python: small pure-Python restatements.

* tm_nn: Biopython 1.76 Bio.SeqUtils.MeltingTemp.Tm_NN(seq, strict=False) with its
  default arguments and a perfect complement, the call polyoligo makes at reference
  src/polyoligo/lib_markers.py:586-587.  Biopython is absent from /root/reference and
  not installable offline; pinned by the value its documentation prints,
  Tm_NN('CGTTCCAAAGATGTGGGCATGAGCTTAC') -> 60.32.
* annotate_mutations: numeric part of PrimerPair.add_mutations, reference
  src/polyoligo/lib_primer3.py:166-210.
"""
import math

# Allawi & SantaLucia (1997), Biochemistry 36: 10581-10594 (Biopython DNA_NN3)
DNA_NN3 = {"init": (0, 0), "init_A/T": (2.3, 4.1), "init_G/C": (0.1, -2.8), "init_oneG/C": (0, 0),
           "init_allA/T": (0, 0), "init_5T/A": (0, 0), "sym": (0, -1.4),
           "AA/TT": (-7.9, -22.2), "AT/TA": (-7.2, -20.4), "TA/AT": (-7.2, -21.3), "CA/GT": (-8.5, -22.7),
           "GT/CA": (-8.4, -22.4), "CT/GA": (-7.8, -21.0), "GA/CT": (-8.2, -22.2), "CG/GC": (-10.6, -27.2),
           "GC/CG": (-9.8, -24.4), "GG/CC": (-8.0, -19.9)}
_COMP = str.maketrans("ACGT", "TGCA")


def tm_nn(seq, dnac1=25, dnac2=25, Na=50, K=0, Tris=0):
    # MeltingTemp._check(seq, "Tm_NN"): upper-case, back-transcribe, then DROP every base that is
    # not in ("A", "C", "G", "T", "I") -- ambiguous / N bases are removed, nothing is raised.
    seq = "".join(seq.split()).upper().replace("U", "T")
    seq = "".join(b for b in seq if b in "ACGTI")
    if not seq:
        raise ValueError("Base not allowed for Tm_NN")     # Biopython: IndexError on seq[0]
    c_seq = seq.translate(_COMP)                            # I stays I
    nn = DNA_NN3
    delta_h = 0
    delta_s = 0
    delta_h += nn["init"][0]
    delta_s += nn["init"][1]
    gc = sum(seq.count(x) for x in "GC")
    key = "init_allA/T" if gc == 0 else "init_oneG/C"
    delta_h += nn[key][0]
    delta_s += nn[key][1]
    if seq.startswith("T"):
        delta_h += nn["init_5T/A"][0]
        delta_s += nn["init_5T/A"][1]
    if seq.endswith("A"):
        delta_h += nn["init_5T/A"][0]
        delta_s += nn["init_5T/A"][1]
    ends = seq[0] + seq[-1]
    AT = ends.count("A") + ends.count("T")
    GC = ends.count("G") + ends.count("C")
    delta_h += nn["init_A/T"][0] * AT
    delta_s += nn["init_A/T"][1] * AT
    delta_h += nn["init_G/C"][0] * GC
    delta_s += nn["init_G/C"][1] * GC
    for k in range(len(seq) - 1):
        neighbors = seq[k:k + 2] + "/" + c_seq[k:k + 2]
        if neighbors in nn:
            delta_h += nn[neighbors][0]
            delta_s += nn[neighbors][1]
        elif neighbors[::-1] in nn:
            delta_h += nn[neighbors[::-1]][0]
            delta_s += nn[neighbors[::-1]][1]
        # else: strict=False -> _key_error only warns (a neighbour pair with an inosine against
        # an inosine is in none of the tables) and the pair contributes nothing
    k = (dnac1 - (dnac2 / 2.0)) * 1e-9
    R = 1.987
    mon = (Na + K + Tris / 2.0) * 1e-3
    corr = 0.368 * (len(seq) - 1) * math.log(mon)
    delta_s += corr
    return (1000 * delta_h) / (delta_s + (R * (math.log(k)))) - 273.15


def annotate_mutations(mutations, windows, product):
    """mutations: [(pos, aaf, indel_size)]; windows: [(start, stop)] per primer;
    product: (start, stop).  Returns ([max_aaf per primer], max_indel)."""
    aafs = []
    for start, stop in windows:
        best = 0
        for pos, aaf, _ in mutations:
            if start <= pos <= stop:
                best = max(best, aaf)
        aafs.append(best)
    indel = 0
    for pos, _, size in mutations:
        if product[0] <= pos <= product[1]:
            indel = max(indel, size)
    return aafs, indel
