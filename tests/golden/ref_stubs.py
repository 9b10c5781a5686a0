"""Stand-ins for the third-party modules the reference imports, so that the
reference's OWN Python (``/root/reference/src/polyoligo``) can be imported in
the build container to generate golden vectors (see make_golden.py).

* ``primer3``: ThermoAnalysis is backed by the CPU oracle (the real primer3-py
  is not installable offline); designPrimers / setP3Globals are inert.
* ``Bio``: the handful of Biopython 1.76 helpers the hot path touches
  (Seq slicing / reverse_complement, SeqUtils.GC, IUPAC table), restated.
* ``vcf``: empty module (VCF parsing is out of scope).

This file is test infrastructure; nothing in the product imports it.
"""
import sys
import types


def install(oracle):
    class _Res:
        def __init__(self, r):
            self.tm, self.dg, self.dh, self.ds = r.tm, r.dg, r.dh, r.ds
            self.structure_found = bool(r.structure_found)

    class ThermoAnalysis:
        def calcTm(self, seq):
            return oracle.calcTm(seq)

        def calcHairpin(self, seq):
            return _Res(oracle.calcHairpin(seq))

        def calcHomodimer(self, seq):
            return _Res(oracle.calcHomodimer(seq))

        def calcHeterodimer(self, s1, s2):
            return _Res(oracle.calcHeterodimer(s1, s2))

    primer3 = types.ModuleType("primer3")
    ta = types.ModuleType("primer3.thermoanalysis")
    ta.ThermoAnalysis = ThermoAnalysis
    bindings = types.ModuleType("primer3.bindings")
    bindings.designPrimers = lambda *a, **k: {}
    primer3.thermoanalysis = ta
    primer3.bindings = bindings
    primer3.setP3Globals = lambda g: None

    comp = str.maketrans("ACGTUacgtuRYKMBVDHrykmbvdhNnSsWw", "TGCAAtgcaaYRMKVBHDyrmkvbhdNnSsWw")

    class Seq(str):
        def reverse_complement(self):
            return Seq(str(self)[::-1].translate(comp))

        def __getitem__(self, k):
            r = str.__getitem__(self, k)
            return Seq(r) if isinstance(k, slice) else r

    bio = types.ModuleType("Bio")
    bio_seq = types.ModuleType("Bio.Seq")
    bio_seq.Seq = Seq
    iupac = types.SimpleNamespace(IUPACData=types.SimpleNamespace(ambiguous_dna_values={
        "A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT",
        "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC"}))
    bio_seq.IUPAC = iupac
    seq_utils = types.ModuleType("Bio.SeqUtils")

    def GC(seq):  # Biopython 1.76 SeqUtils.GC
        gc = sum(seq.count(x) for x in ["G", "C", "g", "c", "S", "s"])
        try:
            return gc * 100.0 / len(seq)
        except ZeroDivisionError:
            return 0.0

    seq_utils.GC = GC
    seq_utils.MeltingTemp = types.ModuleType("Bio.SeqUtils.MeltingTemp")
    from oracle.polyoligo_oracle import tm_nn
    # Biopython 1.76 MeltingTemp.Tm_NN with default arguments (restated in oracle/polyoligo_oracle.py, pinned
    # by the value Biopython's documentation prints); the reference calls Tm_NN(seq, strict=False)
    seq_utils.MeltingTemp.Tm_NN = lambda seq, strict=True, **kw: tm_nn(str(seq))
    bio.Seq = bio_seq
    bio.SeqUtils = seq_utils
    mods = {"primer3": primer3, "primer3.thermoanalysis": ta, "primer3.bindings": bindings, "Bio": bio,
            "Bio.Seq": bio_seq, "Bio.SeqUtils": seq_utils, "Bio.SeqUtils.MeltingTemp": seq_utils.MeltingTemp,
            "vcf": types.ModuleType("vcf")}
    sys.modules.update(mods)
    if "/root/reference/src" not in sys.path:
        sys.path.insert(0, "/root/reference/src")
