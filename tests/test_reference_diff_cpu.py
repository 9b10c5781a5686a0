"""Randomised differential tests of the host logic against the REFERENCE's own Python, imported from
/root/reference where that tree exists (the build container); skipped elsewhere (the GPU box), where
the committed fixtures under tests/golden/ -- made by the same code -- stand in.  Third-party modules
the reference imports (primer3, Bio, vcf) are the stand-ins of tests/golden/ref_stubs.py; nothing
here evaluates thermodynamics."""
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "polyoligo")),
                                reason="the reference tree is not available on this machine")


@pytest.fixture(scope="module")
def ref(oracle):
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for p in (here, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ref_stubs
    ref_stubs.install(oracle)
    import polyoligo.lib_utils as lu
    import polyoligo.lib_primer3 as lp
    import polyoligo.lib_markers as lm
    import polyoligo.scoring as sc
    return types.SimpleNamespace(lu=lu, lp=lp, lm=lm, sc=sc)


def rand_seq(rng, n):
    return "".join(np.array(list("ACGT"))[rng.integers(0, 4, n)])


def test_interval_helpers(ref):
    from polyoligo_b200 import lib_utils as mine
    rng = np.random.default_rng(11)
    for _ in range(300):
        n = int(rng.integers(0, 60))
        ixs = rng.integers(0, 80, n)                   # unsorted, with duplicates
        assert mine.list_2_ranges(ixs) == ref.lu.list_2_ranges(list(ixs))
    for _ in range(100):
        m = int(rng.integers(0, 40))
        starts = np.sort(rng.choice(400, m, replace=False))
        ivs = [[int(s), int(rng.integers(1, 6))] for s in starts]
        keep = int(rng.integers(1, 45))
        got = mine.reduce_ivs([list(x) for x in ivs], keep)
        want = ref.lu.reduce_ivs([list(x) for x in ivs], n=keep)
        assert [[int(a), int(b)] for a, b in got] == [[int(a), int(b)] for a, b in want]


def test_round_tidy_and_ambiguous_dna(ref):
    from polyoligo_b200 import lib_utils as mine
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.normal(0, 50, 300), rng.random(100) * 1e-4, [0.0, 1.0, -1.0, 99.95, 0.5, 1e-300, 59.99999]])
    for x in xs:
        for n in (2, 3, 4):
            assert mine.round_tidy(float(x), n) == ref.lu.round_tidy(float(x), n), (x, n)
    for _ in range(200):
        ln = int(rng.integers(1, 30))
        seqs = [rand_seq(rng, ln) for _ in range(int(rng.integers(1, 5)))]
        assert mine.seqs2ambiguous_dna(seqs) == ref.lu.seqs2ambiguous_dna(seqs)


def test_scoring_functions(ref):
    from polyoligo_b200 import scoring as mine
    rng = np.random.default_rng(13)
    for _ in range(400):
        def pp_like():
            prim = {d: types.SimpleNamespace(tm=float(t), max_aaf=float(a))
                    for d, t, a in zip("FRA", 60 + rng.normal(0, 3, 3), rng.choice([0.0, 0.05, 0.0999, 0.1, 0.3], 3))}
            return types.SimpleNamespace(
                primers=prim, offtargets=["x"] * int(rng.integers(0, 2)), dimer=bool(rng.integers(0, 2)),
                ref_dimer=bool(rng.integers(0, 2)), alt_dimer=bool(rng.integers(0, 2)),
                max_indel_size=int(rng.choice([0, 1, 49, 50, 120])),
                hrm={"delta": float(rng.choice([0.0, 0.49, 0.5, 0.99, 1.0, 2.5]))}, goodness=None, qcode=None)
        pp = pp_like()
        two = types.SimpleNamespace(**{**pp.__dict__, "primers": {d: pp.primers[d] for d in "FR"}})
        for name, obj in (("PCR", two), ("CAPS", two), ("HRM", two), ("KASP", pp)):
            want = ref.sc.get_score_fnc(name)(obj)
            got = mine.get_score_fnc(name)(obj)
            assert (int(got[0]), got[1]) == (int(want[0]), want[1]), name


def test_merge_with_primers_and_masks(ref):
    from polyoligo_b200 import lib_markers as mine_lm, lib_primer3 as mine_lp
    rng = np.random.default_rng(14)
    genome = rand_seq(rng, 3000)

    class Fetch:
        def fetch(self, query):
            q = query[0]
            return {"%s:%d-%d" % (q["chr"], q["start"], q["stop"]): genome[q["start"] - 1:q["stop"]]}

    for trial in range(40):
        ts = int(rng.integers(700, 1500))
        tlen = int(rng.integers(1, 400))
        flank = int(rng.integers(60, 300))
        overlap = bool(rng.integers(0, 2))             # HRM / CAPS style: the flanks reach the marker
        lstart = ts - flank + (5 if overlap else -int(rng.integers(0, 30)))
        lstop = lstart + flank - 1
        rstart = ts + tlen - 1 - (5 if overlap else -int(rng.integers(0, 30)))
        rstop = rstart + flank - 1
        lstop = min(lstop, rstop)
        muts_pos = sorted(set(int(p) for p in rng.integers(lstart, rstop + 1, 25)))

        def build(lm_mod, lp_mod):
            rois = []
            state = np.random.default_rng(1000 + trial)
            for a, b in ((lstart, lstop), (rstart, rstop)):
                roi = lm_mod.ROI(chrom="c", start=a, stop=b, blast_hook=Fetch(), vcf_hook=True, name="t")
                n = roi.n
                mism = state.random(n) < 0.2
                roi.p3_sequence_included_maps = {"all": np.repeat(True, n), "partial_mism": mism | (state.random(n) < 0.3),
                                                 "mism": mism}
                roi.mutations = [types.SimpleNamespace(pos=p) for p in muts_pos if a <= p <= b]
                lp_mod.include_mut_in_included_maps(roi)
                rois.append(roi)
            region = lm_mod.Region(chrom="c", start=ts, stop=ts + tlen - 1, blast_hook=Fetch(), vcf_hook=True, name="r",
                                   left_roi=rois[0], right_roi=rois[1])
            merged = region.merge_with_primers()
            lp_mod.get_sequence_excluded_regions(merged)
            return merged

        got, want = build(mine_lm, mine_lp), build(ref.lm, ref.lp)
        assert (got.start, got.stop, got.n, got.seq) == (want.start, want.stop, want.n, want.seq)
        assert [list(x) for x in got.p3_sequence_target] == [list(x) for x in want.p3_sequence_target]
        assert set(got.p3_sequence_included_maps) == set(want.p3_sequence_included_maps)
        for k in want.p3_sequence_included_maps:
            assert np.array_equal(got.p3_sequence_included_maps[k], want.p3_sequence_included_maps[k]), (trial, k)
            assert [[int(a), int(b)] for a, b in got.p3_sequence_excluded_regions[k]] == \
                   [[int(a), int(b)] for a, b in want.p3_sequence_excluded_regions[k]]
        assert [m.pos for m in got.mutations] == [m.pos for m in want.mutations]


def test_offtarget_listing(ref):
    from polyoligo_b200 import lib_primer3 as mine_lp
    rng = np.random.default_rng(15)
    ref.lp.set_offtarget_size(30, 400)
    mine_lp.set_offtarget_size(30, 400)
    try:
        for _ in range(100):
            hits = {}
            for d in "FR":
                hits[d] = {}
                for chrom in ("c1", "c2", "c3")[:int(rng.integers(1, 4))]:
                    hits[d][chrom] = [{"sstart": int(x)} for x in rng.integers(1, 1500, int(rng.integers(0, 8)))]
            assert mine_lp.list_offtargets(hits) == ref.lp.PCR.list_offtargets(hits)
    finally:
        ref.lp.set_offtarget_size(0, 1000)
        mine_lp.set_offtarget_size(0, 1000)


def test_add_mutations_host_form(ref):
    # PrimerPair.add_mutations (labels, IUPAC-ambiguous sequences, max_aaf, indel size): the host form
    # of this package (one pair, no device) against the reference's method on random primers / variants
    from polyoligo_b200 import lib_primer3 as mine_lp
    rng = np.random.default_rng(16)
    for trial in range(150):
        kasp = bool(rng.integers(0, 2))
        direction = "R" if rng.integers(0, 2) else "F"
        fs = int(rng.integers(100, 200))
        fl, rl = int(rng.integers(18, 26)), int(rng.integers(18, 26))
        rstop = fs + int(rng.integers(60, 160))
        spec = {"F": (rand_seq(rng, fl), fs, fs + fl - 1), "R": (rand_seq(rng, rl), rstop - rl + 1, rstop)}
        if kasp:
            d = direction
            spec["A"] = (rand_seq(rng, len(spec[d][0])), spec[d][1], spec[d][2])
        muts = []
        for pos in sorted(set(int(p) for p in rng.integers(fs - 5, rstop + 6, 12))):
            kind = int(rng.integers(0, 4))
            r0 = rand_seq(rng, 1)
            if kind == 0:
                m = dict(ref=r0 + rand_seq(rng, int(rng.integers(1, 4))), alt=[r0])                  # deletion
            elif kind == 1:
                m = dict(ref=r0, alt=[r0 + rand_seq(rng, int(rng.integers(1, 70)))])                 # insertion
            else:
                m = dict(ref=r0, alt=[rand_seq(rng, 1) for _ in range(int(rng.integers(1, 3)))])     # SNP, 1-2 alleles
            indel = max(abs(len(m["ref"]) - len(a)) for a in m["alt"])
            muts.append(types.SimpleNamespace(pos=pos, aaf=float(rng.choice([0.0, 0.013, 0.1, 0.25, 0.5])),
                                              indel_size=indel, **m))

        def make(mod):
            pp = mod.PrimerPair()
            if kasp:
                pp.primers["A"] = mod.Primer()
            pp.dir = direction
            for d, (seq, a, b) in spec.items():
                pp.primers[d].sequence, pp.primers[d].start, pp.primers[d].stop = seq, a, b
            return pp

        got, want = make(mine_lp), make(ref.lp)
        got.add_mutations(muts)
        want.add_mutations(muts)
        assert got.max_indel_size == want.max_indel_size
        for d in spec:
            g, w = got.primers[d], want.primers[d]
            assert (g.sequence, g.sequence_ambiguous, list(g.mutations), float(g.max_aaf)) == \
                   (w.sequence, w.sequence_ambiguous, list(w.mutations), float(w.max_aaf)), (trial, d)
