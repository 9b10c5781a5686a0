"""CPU tests of the oracle (oracle/): anchors, table provenance, invariants.

The reference's own tests hold no numeric vectors, so the
oracle is pinned where a published value exists (primer3-py README) and checked
through properties of the model elsewhere.
"""
import os

import numpy as np
import pytest

from helpers import rand_seqs, revcomp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_calctm_readme_anchor_bit_exact(oracle):
    # primer3-py README: calcTm('GTAAAACGACGGCCAGT') == 49.16808228911765
    assert oracle.calcTm("GTAAAACGACGGCCAGT") == 49.16808228911765


def test_calctm_matches_independent_python_restatement(oracle):
    # independent pure-Python restatement of oligotm
    import math
    DH = dict(zip([a + b for a in "ACGT" for b in "ACGT"],
                  [79, 84, 78, 72, 85, 80, 106, 78, 82, 98, 80, 84, 72, 82, 85, 79]))
    DS = dict(zip([a + b for a in "ACGT" for b in "ACGT"],
                  [222, 224, 210, 204, 227, 199, 272, 210, 222, 244, 199, 224, 213, 222, 227, 222]))

    def tm(s, mv=50.0, dna=50.0):
        sym = len(s) % 2 == 0 and s == revcomp(s)
        dh = ds = 0
        if sym:
            ds += 14
        for c in (s[0], s[-1]):
            if c in "AT":
                ds += -41
                dh += -23
            else:
                ds += 28
                dh += -1
        for i in range(len(s) - 1):
            dh += DH[s[i:i + 2]]
            ds += DS[s[i:i + 2]]
        delta_h = dh * -100.0
        delta_s = ds * -0.1
        delta_s = delta_s + 0.368 * (len(s) - 1) * math.log(mv / 1000.0)
        return delta_h / (delta_s + 1.987 * math.log(dna / (1e9 if sym else 4e9))) - 273.15

    rng = np.random.default_rng(3)
    seqs = rand_seqs(rng, 500, 2, 60) + ["ACGT", "GAATTC", "AT"]
    got = oracle.tm_batch(seqs)
    for s, g in zip(seqs, got):
        assert g == tm(s), s


def test_calctm_invalid_and_long(oracle):
    assert oracle.calcTm("ACGTNACGT") == -999999.9999
    long_seq = "ACGT" * 20  # 80 nt > max_nn_length: GC formula
    import math
    expect = 81.5 + (16.6 * math.log10(50.0 / 1000.0)) + (41.0 * (40 / 80)) - (600.0 / 80)
    assert oracle.calcTm(long_seq) == expect


README_HAIRPIN = "CCCCCATCCGATCAGGGGG"   # primer3-py README: tm=34.15, dg=337.09, dh=-36300.00, ds=-118.13


def _anchor_oracle(oracle):
    """The oracle on upstream's primer3_config/ tables when POLYOLIGO_P3_CONFIG points at them,
    else on the embedded literature tables."""
    d = os.environ.get("POLYOLIGO_P3_CONFIG")
    if d:
        from oracle.oracle import Oracle
        return Oracle(d)
    return oracle


def test_hairpin_readme_anchor_structure_and_dh(oracle):
    # The part of the only published thal anchor that the literature tables DO reproduce:
    # same structure (5 bp stem, 9 nt loop) and dH = 4 CC/GG stacks + the CA/GA terminal
    # mismatch = -36300 exactly; dg / tm are consistent with the reported ds.
    r = _anchor_oracle(oracle).calcHairpin(README_HAIRPIN)
    assert r.structure_found == 1
    assert r.dh == -36300.0
    assert abs(r.dg - (r.dh - 310.15 * r.ds)) < 1e-9
    assert abs(r.tm - (r.dh / r.ds - 273.15)) < 1e-9


@pytest.mark.xfail(condition=not os.environ.get("POLYOLIGO_P3_CONFIG"), strict=True,
                   reason="thal PARITY UNPINNED: embedded literature tables give ds=-108.11 / tm=62.63; upstream's "
                          "tstack2.ds / loops.ds are not available offline (scripts/localise_anchor.py names the "
                          "entries).  Passes -- and, being strict, demands removal of this mark -- the moment "
                          "the real primer3_config/ tables are the default; with POLYOLIGO_P3_CONFIG=<dir> it "
                          "is a plain assertion.")
def test_hairpin_readme_anchor_upstream_values(oracle):
    # upstream's published numbers at print precision (2 dp)
    r = _anchor_oracle(oracle).calcHairpin(README_HAIRPIN)
    assert r.structure_found == 1
    assert round(r.tm, 2) == 34.15
    assert round(r.dg, 2) == 337.09
    assert round(r.dh, 2) == -36300.00
    assert round(r.ds, 2) == -118.13


def test_default_tables_match_literature_spot_checks(oracle):
    stack_h, stack_s = oracle.table(0)

    def q(a, b, c, d):
        code = "ACGT".index
        return ((code(a) * 5 + code(b)) * 5 + code(c)) * 5 + code(d)

    # SantaLucia (1998) unified nearest neighbours
    assert (stack_h[q("A", "A", "T", "T")], stack_s[q("A", "A", "T", "T")]) == (-7900.0, -22.2)
    assert (stack_h[q("C", "G", "G", "C")], stack_s[q("C", "G", "G", "C")]) == (-10600.0, -27.2)
    assert (stack_h[q("G", "C", "C", "G")], stack_s[q("G", "C", "C", "G")]) == (-9800.0, -24.4)
    assert (stack_h[q("T", "A", "A", "T")], stack_s[q("T", "A", "A", "T")]) == (-7200.0, -21.3)
    assert np.isinf(stack_h[q("A", "A", "A", "T")])
    # the strand-exchange symmetry every WC stack must obey: ab/cd == dc/ba
    for a in "ACGT":
        for b in "ACGT":
            c, d = revcomp(a), revcomp(b)
            assert stack_h[q(a, b, c, d)] == stack_h[q(d, c, b, a)]
            assert stack_s[q(a, b, c, d)] == stack_s[q(d, c, b, a)]
    mm_h, mm_s = oracle.table(1)
    assert (mm_h[q("A", "A", "A", "T")], mm_s[q("A", "A", "A", "T")]) == (4700.0, 12.9)   # TA/AA, Peyret 1999
    assert (mm_h[q("C", "G", "G", "T")], mm_s[q("C", "G", "G", "T")]) == (-4100.0, -11.7)  # CG/GT, Allawi 1997
    ts_h, ts_s = oracle.table(3)
    assert (ts_h[q("C", "A", "G", "A")], ts_s[q("C", "A", "G", "A")]) == (-4300.0, -10.7)
    inter_h, inter_s = oracle.table(6)
    bulge_h, bulge_s = oracle.table(7)
    hp_h, hp_s = oracle.table(8)
    assert bulge_s[0] == -12.89 and inter_s[2] == -10.31 and hp_s[2] == -11.28 and hp_s[8] == -14.5
    assert inter_s[29] == -21.28 and bulge_s[29] == -19.02 and hp_s[29] == -20.31
    assert np.all(hp_h[2:] == 0)


def test_param_dir_loader_equals_embedded_defaults(oracle):
    from oracle.oracle import Oracle
    o2 = Oracle(os.path.join(ROOT, "polyoligo_b200", "params"))
    for t in range(9):
        h1, s1 = oracle.table(t)
        h2, s2 = o2.table(t)
        assert np.array_equal(h1, h2) and np.array_equal(s1, s2)


def test_generated_param_files_are_current():
    import subprocess
    import sys
    import filecmp
    import tempfile
    import shutil
    pdir = os.path.join(ROOT, "polyoligo_b200", "params")
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copytree(pdir, os.path.join(tmp, "before"))
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_default_params.py")],
                              stdout=subprocess.DEVNULL)
        cmp = filecmp.dircmp(os.path.join(tmp, "before"), pdir)
        assert not cmp.diff_files and not cmp.left_only and not cmp.right_only


def test_perfect_duplex_equals_nn_sum(oracle):
    # A perfect duplex's dH must be the nearest-neighbour sum + initiation + end terms:
    # check against oligotm-style accounting on a GC-clamped oligo (no AT penalty).
    s = "GCATCGATCGGATCGATGCC"
    r = oracle.calcHeterodimer(s, revcomp(s))
    assert r.structure_found == 1
    nn = {"AA": -7900, "AC": -8400, "AG": -7800, "AT": -7200, "CA": -8500, "CC": -8000, "CG": -10600,
          "CT": -7800, "GA": -8200, "GC": -9800, "GG": -8000, "GT": -8400, "TA": -7200, "TC": -8200,
          "TG": -8500, "TT": -7900}
    expect = sum(nn[s[i:i + 2]] for i in range(len(s) - 1)) + 200.0
    assert r.dh == expect
    assert (r.align_end_1, r.align_end_2) == (len(s), len(s))


def test_no_structure_gives_zero_tm(oracle):
    r = oracle.calcHeterodimer("A" * 20, "C" * 20)
    assert r.structure_found == 0 and r.tm == 0.0
    r = oracle.calcHairpin("A" * 20)
    assert r.structure_found == 0 and r.tm == 0.0


def test_salt_and_concentration_monotonic(oracle):
    s = "GCATCGATCGGATCGATGCC"
    lo = oracle.thal(s, revcomp(s), 1, mv=20.0).tm
    hi = oracle.thal(s, revcomp(s), 1, mv=200.0).tm
    assert hi > lo
    lo = oracle.thal(s, revcomp(s), 1, dna_conc=5.0).tm
    hi = oracle.thal(s, revcomp(s), 1, dna_conc=500.0).tm
    assert hi > lo


def test_errors(oracle):
    assert oracle.thal("", "ACGT", 1).status == 1
    assert oracle.thal("A" * 61, "C" * 61, 1).status == 2
    assert oracle.thal("A" * 61, "A" * 61, 4).status == 2


def test_batch_equals_scalar_and_is_thread_invariant(oracle):
    rng = np.random.default_rng(5)
    a, b = rand_seqs(rng, 200, 10, 40), rand_seqs(rng, 200, 10, 40)
    one = oracle.thal_batch(a, b, 1, nthreads=1)
    many = oracle.thal_batch(a, b, 1, nthreads=4)
    assert one.tobytes() == many.tobytes()
    for k in range(0, 200, 17):
        r = oracle.calcHeterodimer(a[k], b[k])
        assert (r.tm, r.dg, r.dh, r.ds) == (one["tm"][k], one["dg"][k], one["dh"][k], one["ds"][k])
