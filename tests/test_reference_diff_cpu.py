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


def _fill_pairs(rng, mod, pair_cls, n, kasp=False, hrm=False, caps=False):
    """n primer pairs with every attribute the classify / prune / report code reads, set directly."""
    pps = []
    for k in range(n):
        pp = pair_cls()
        pp.dir = "R" if rng.integers(0, 2) else "F"
        start = int(rng.integers(1000, 5000))
        size = int(rng.integers(60, 400))
        keys = ["F", "R", "A"] if kasp else ["F", "R"]
        if kasp and "A" not in pp.primers:
            pp.primers["A"] = mod.Primer()
        for d in keys:
            p = pp.primers[d]
            ln = int(rng.integers(18, 26))
            p.sequence = rand_seq(rng, ln)
            p.sequence_ambiguous = p.sequence[:3] + "R" + p.sequence[4:]
            p.length = ln
            p.start, p.stop = (start, start + ln - 1) if d != "R" else (start + size - ln, start + size - 1)
            p.tm = float(55 + 10 * rng.random())
            p.gc_content = float(rng.choice([40.0, 45.454545, 52.17391304, 60.0]))
            p.max_aaf = float(rng.choice([0.0, 0.0, 0.04, 0.13, 0.5]))
            p.mutations = ["A%dG:0.1" % (p.start + 2)] if rng.integers(0, 3) == 0 else []
            if kasp:
                p.reporter = ""
        pp.product_size = size
        pp.offtargets = ["chr1:%d-%d" % (start, start + 300)] * int(rng.integers(0, 2))
        pp.max_indel_size = int(rng.choice([0, 0, 3, 60]))
        pp.dimer = bool(rng.integers(0, 2))
        if kasp:
            pp.ref_dimer, pp.alt_dimer = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        if hrm:
            pp.hrm = {"delta": round(0.05 * k + 0.013, 3), "ref": 75.1234, "alt": 75.9}   # distinct deltas: no ties
            pp.seq_x = "ACGTxACGT"
        if caps:
            pp.enzymes = [{"name": "EcoRI", "supplier": "BN", "allele": "REF", "fragl": 100 + k, "fragr": 200}]
            if k % 3 == 0:
                pp.enzymes.append({"name": "TaqI", "supplier": "X", "allele": "ALT", "fragl": 77, "fragr": 88 + k})
            pp.seq_x = "ACGTxACGT"
        pps.append(pp)
    return pps


@pytest.mark.parametrize("assay", ["PCR", "HRM", "CAPS", "KASP"])
def test_classify_prune_and_report_writers(ref, assay, tmp_path):
    # classify + prune + print_report on random repositories, files compared byte for byte
    import importlib
    from polyoligo_b200 import caps, hrm, kasp, lib_primer3 as mine_lp, pcr
    ref_mod = importlib.import_module("polyoligo._lib_" + assay.lower())
    mine_mod = {"PCR": pcr, "HRM": hrm, "CAPS": caps, "KASP": kasp}[assay]
    for trial in range(12):
        out = {}
        for tag, mod, lpm in (("mine", mine_mod, mine_lp), ("ref", ref_mod, ref.lp)):
            rng = np.random.default_rng(500 + trial)           # the same repository on both sides
            pair_cls = mod.PrimerPair if assay == "KASP" else lpm.PrimerPair
            n = int(rng.integers(0, 14)) if trial else 0
            pps = _fill_pairs(rng, lpm, pair_cls, n, kasp=assay == "KASP", hrm=assay == "HRM", caps=assay == "CAPS")
            if assay == "PCR":
                p = lpm.PCR(chrom="chr7", name="reg_%d" % trial)
            else:
                p = mod.PCR(snp_id="mk%d" % trial, chrom="chr7", pos=4321, ref="A", alt="" if trial % 2 else "GT")
            p.pps = pps
            if assay == "KASP":
                p.add_reporter_dyes(["GAAGGTCGGAGTCAACGGATT", "GAAGGTGACCAAGTTCATGCT"])
            p.classify(assay_name=assay)
            kept = p.prune(5, assay_name="HRM") if assay == "HRM" else p.prune(5)
            base = str(tmp_path / ("%s_%s_%d" % (tag, assay, trial)))
            mod.print_report(p, base)
            mod.print_report_header(base + "_hdr.txt")
            out[tag] = (int(kept), open(base + ".txt").read(), open(base + ".bed").read(), open(base + "_hdr.txt").read())
        assert out["mine"] == out["ref"], (assay, trial)


def test_get_primers_result_parsing(ref, monkeypatch):
    # lib_primer3.get_primers' dict -> PrimerPair conversion on random designPrimers-style answers
    from polyoligo_b200 import lib_primer3 as mine_lp
    rng = np.random.default_rng(17)
    for trial in range(60):
        n = int(rng.integers(0, 12))
        d = {"PRIMER_LEFT_EXPLAIN": "considered 10", "PRIMER_PAIR_NUM_RETURNED": n, "PRIMER_LEFT_NUM_RETURNED": n}
        order = list(rng.permutation(n))
        for i in order:
            i = int(i)
            ls, ll, rs, rl = int(rng.integers(0, 200)), int(rng.integers(18, 26)), int(rng.integers(300, 500)), int(rng.integers(18, 26))
            d["PRIMER_PAIR_%d_PENALTY" % i] = float(rng.random())
            d["PRIMER_LEFT_%d_SEQUENCE" % i] = rand_seq(rng, ll)
            d["PRIMER_RIGHT_%d_SEQUENCE" % i] = rand_seq(rng, rl)
            d["PRIMER_LEFT_%d" % i] = (ls, ll)
            d["PRIMER_RIGHT_%d" % i] = (rs, rl)
            for side in ("LEFT", "RIGHT"):
                for f in ("TM", "GC_PERCENT", "SELF_ANY_TH", "SELF_END_TH", "HAIRPIN_TH", "END_STABILITY", "PENALTY"):
                    d["PRIMER_%s_%d_%s" % (side, i, f)] = float(rng.random() * 60)
            d["PRIMER_PAIR_%d_COMPL_ANY_TH" % i] = float(rng.random())
            d["PRIMER_PAIR_%d_COMPL_END_TH" % i] = float(rng.random())
            d["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] = rs - ls + 1
        if trial % 10 == 9:
            d["PRIMER_WARNING" if trial % 20 == 9 else "PRIMER_ERROR"] = "something"
        ts = int(rng.integers(1, 10 ** 6))
        sys.modules["primer3"].bindings.designPrimers = lambda seq_args, _d=d: dict(_d)
        want = ref.lp.get_primers({"SEQUENCE_TEMPLATE": "ACGT"}, target_start=ts)
        got = mine_lp.parse_p3_result(dict(d), target_start=ts)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            gd = {k: v for k, v in g.__dict__.items() if k != "primers"}
            wd = {k: v for k, v in w.__dict__.items() if k != "primers"}
            for k in wd:                     # every attribute the reference's pair carries
                if k in ("enzymes", "seq_x"):
                    continue
                assert gd[k] == wd[k], (trial, k)
            for side in "FR":
                wp = w.primers[side].__dict__
                gp = g.primers[side].__dict__
                for k, v in wp.items():
                    if k in ("is_marker", "marker_pos", "seed"):
                        continue
                    assert gp[k] == v, (trial, side, k)
    sys.modules["primer3"].bindings.designPrimers = lambda *a, **k: {}


def test_kasp_merge_and_round_robin(ref, monkeypatch):
    # _lib_kasp.design_primers (merge_primers + the round-robin over the marker primers + dedup + the
    # validity of the common primer) on random canned picker answers and a canned validity rule
    import zlib
    import polyoligo._lib_kasp as ref_kasp
    from polyoligo_b200 import kasp as mine_kasp, lib_primer3 as mine_lp
    rng = np.random.default_rng(18)

    def valid_rule(seq):
        return zlib.crc32(seq.encode()) % 4 != 0

    monkeypatch.setattr(ref.lp.Primer, "is_valid", lambda self: valid_rule(self.sequence))
    monkeypatch.setattr(mine_lp, "is_valid_batch", lambda primers, device=None: [valid_rule(p.sequence) for p in primers])
    for trial in range(40):
        template = rand_seq(rng, 400)
        roi = types.SimpleNamespace(seq=template, start=int(rng.integers(1, 10 ** 5)), chrom="c")
        n_mpp = int(rng.integers(1, 7))
        spec = []
        for m in range(n_mpp):
            d = "R" if rng.integers(0, 2) else "F"
            ln = int(rng.integers(18, 26))
            spec.append(dict(dir=d, seq=rand_seq(rng, ln), alt=rand_seq(rng, ln), start=roi.start + 200 - ln, stop=roi.start + 199))
        pool = [rand_seq(rng, int(rng.integers(18, 26))) for _ in range(12)]      # common primers, shared: duplicates occur
        answers = {}
        for passno in range(2):
            for m, sp in enumerate(spec):
                d = {}
                for i in range(int(rng.integers(0, 7))):
                    other = pool[int(rng.integers(0, len(pool)))]
                    left, right = (sp["seq"], other) if sp["dir"] == "F" else (other, sp["seq"])
                    d["PRIMER_PAIR_%d_PENALTY" % i] = float(i)
                    d["PRIMER_LEFT_%d_SEQUENCE" % i] = left
                    d["PRIMER_RIGHT_%d_SEQUENCE" % i] = right
                    d["PRIMER_LEFT_%d" % i] = (int(rng.integers(0, 150)), len(left))
                    d["PRIMER_RIGHT_%d" % i] = (int(rng.integers(250, 399)), len(right))
                    for side in ("LEFT", "RIGHT"):
                        for f in ("END_STABILITY", "HAIRPIN_TH", "PENALTY", "SELF_ANY_TH", "SELF_END_TH"):
                            d["PRIMER_%s_%d_%s" % (side, i, f)] = float(rng.random())
                    d["PRIMER_PAIR_%d_COMPL_ANY_TH" % i] = 1.0
                    d["PRIMER_PAIR_%d_COMPL_END_TH" % i] = 2.0
                    d["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] = 150
                answers[(passno, m)] = d
        state = {"pass": 0}

        def find(args):
            key = "SEQUENCE_PRIMER" if "SEQUENCE_PRIMER" in args else "SEQUENCE_PRIMER_REVCOMP"
            d = "F" if key == "SEQUENCE_PRIMER" else "R"
            for m, sp in enumerate(spec):
                if sp["dir"] == d and sp["seq"] == args[key]:
                    return dict(answers[(state["pass"], m)])
            raise AssertionError("unknown marker primer")

        sys.modules["primer3"].bindings.designPrimers = lambda seq_args: find(seq_args)
        monkeypatch.setattr(mine_lp, "get_primers_batch",
                            lambda seq_args_list, target_starts, device=None:
                            [mine_lp.parse_p3_result(find(a), ts) for a, ts in zip(seq_args_list, target_starts)])

        def mpps_of(kmod):
            out = []
            for sp in spec:
                pp = kmod.PrimerPair()
                pp.dir = sp["dir"]
                p = pp.primers[sp["dir"]]
                p.sequence, p.start, p.stop = sp["seq"], sp["start"], sp["stop"]
                pp.primers["A"].sequence = sp["alt"]
                out.append(pp)
            return out

        got, want = [], []
        for passno, n_primers in ((0, int(rng.integers(1, 12))), (1, int(rng.integers(6, 20)))):
            state["pass"] = passno
            got = mine_kasp.design_primers(got, mpps_of(mine_kasp), roi, sequence_excluded_region=[[0, 5]], n_primers=n_primers)
            want = ref_kasp.design_primers(pps_repo=want, mpps=mpps_of(ref_kasp), roi=roi, sequence_excluded_region=[[0, 5]],
                                           n_primers=n_primers)
            rows = lambda repo: [(pp.dir, pp.penalty, pp.product_size) + tuple(
                (pp.primers[d].sequence, pp.primers[d].start, pp.primers[d].stop, pp.primers[d].id) for d in "FRA") for pp in repo]
            assert rows(got) == rows(want), (trial, passno)
    sys.modules["primer3"].bindings.designPrimers = lambda *a, **k: {}


def test_pcr_selection_with_max_unique(ref, monkeypatch):
    # _lib_pcr.design_primers: validity of both primers + the max_unique rule (incl. its counting of rejected
    # pairs) over several search passes into one repository, on random canned picker answers
    import zlib
    import polyoligo._lib_pcr as ref_pcr
    from polyoligo_b200 import lib_primer3 as mine_lp, pcr as mine_pcr
    rng = np.random.default_rng(19)

    def valid_rule(seq):
        return zlib.crc32(seq.encode()) % 5 != 0

    monkeypatch.setattr(ref.lp.Primer, "is_valid", lambda self: valid_rule(self.sequence))
    monkeypatch.setattr(mine_lp, "is_valid_batch", lambda primers, device=None: [valid_rule(p.sequence) for p in primers])
    for trial in range(40):
        lefts = [rand_seq(rng, int(rng.integers(18, 26))) for _ in range(5)]
        rights = [rand_seq(rng, int(rng.integers(18, 26))) for _ in range(5)]
        roi = types.SimpleNamespace(seq=rand_seq(rng, 300), chrom="c", left_roi=types.SimpleNamespace(start=int(rng.integers(1, 9999))))
        answers = []
        for passno in range(3):
            d = {}
            for i in range(int(rng.integers(0, 10))):
                left, right = lefts[int(rng.integers(0, 5))], rights[int(rng.integers(0, 5))]
                d["PRIMER_PAIR_%d_PENALTY" % i] = float(i)
                d["PRIMER_LEFT_%d_SEQUENCE" % i] = left
                d["PRIMER_RIGHT_%d_SEQUENCE" % i] = right
                d["PRIMER_LEFT_%d" % i] = (int(rng.integers(0, 100)), len(left))
                d["PRIMER_RIGHT_%d" % i] = (int(rng.integers(200, 299)), len(right))
                d["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] = 120
            if passno == 2 and trial % 7 == 0:
                d["PRIMER_ERROR"] = "boom"
            answers.append(d)
        state = {"pass": 0}
        sys.modules["primer3"].bindings.designPrimers = lambda seq_args: dict(answers[state["pass"]])
        monkeypatch.setattr(mine_lp, "get_primers_batch",
                            lambda seq_args_list, target_starts, device=None:
                            [mine_lp.parse_p3_result(dict(answers[state["pass"]]), ts) for ts in target_starts])
        max_unique = [np.inf, 1, 2, 3][trial % 4]
        n_primers = int(rng.integers(1, 9))
        got, want = [], []
        for passno in range(3):
            state["pass"] = passno
            got = mine_pcr.design_primers(got, roi, sequence_target=[[100, 50]], sequence_excluded_region=[[0, 3]],
                                          n_primers=n_primers, max_unique=max_unique)
            want = ref_pcr.design_primers(pps_repo=want, roi=roi, sequence_target=[[100, 50]], sequence_excluded_region=[[0, 3]],
                                          n_primers=n_primers, max_unique=max_unique)
            rows = lambda repo: [(pp.chrom, pp.penalty, pp.product_size) + tuple(
                (pp.primers[d].sequence, pp.primers[d].start, pp.primers[d].stop, pp.primers[d].id) for d in "FR") for pp in repo]
            assert rows(got) == rows(want), (trial, passno, max_unique)
    sys.modules["primer3"].bindings.designPrimers = lambda *a, **k: {}


def test_will_it_cut_on_the_reference_enzyme_table(ref):
    # PCRproduct (seq / seq_alt / seq_x) and will_it_cut with every enzyme of the reference's type2.json
    # (degenerate recognition sites included), shared-enzyme-dict quirk included
    import polyoligo._lib_caps as ref_caps
    from polyoligo_b200 import caps as mine_caps, lib_markers as mine_lm
    rng = np.random.default_rng(20)
    table = ref_caps.get_enzymes(included_enzymes=None)
    assert mine_caps.get_enzymes(filename=ref_caps.ENZYME_FILENAME) == table
    assert [e["name"] for e in mine_caps.get_enzymes(["EcoRI", "TaqI"], ref_caps.ENZYME_FILENAME)] == \
           [e["name"] for e in ref_caps.get_enzymes(included_enzymes=["EcoRI", "TaqI"])]
    genome = rand_seq(rng, 3000)

    class Fetch:
        def fetch(self, query):
            q = query[0]
            return {"%s:%d-%d" % (q["chr"], q["start"], q["stop"]): genome[q["start"] - 1:q["stop"]]}

    for trial in range(25):
        pos1 = int(rng.integers(800, 2200))
        refa = genome[pos1 - 1:pos1 - 1 + int(rng.choice([1, 1, 1, 2]))]
        alt = rand_seq(rng, int(rng.choice([1, 1, 2, 3])))
        while alt == refa:
            alt = rand_seq(rng, len(alt))
        marker = types.SimpleNamespace(name="m%d" % trial, chrom="c", pos1=pos1, ref=refa, alt=alt, n=len(refa))
        mine_enz, ref_enz = [dict(e) for e in table], [dict(e) for e in table]
        for k in range(6):                                   # several products share one enzyme list (the quirk)
            fstart = pos1 - int(rng.integers(30, 400))
            rstop = pos1 + int(rng.integers(30, 400))
            pp = types.SimpleNamespace(primers={"F": types.SimpleNamespace(start=fstart), "R": types.SimpleNamespace(stop=rstop)})
            roi = types.SimpleNamespace(chrom="c", blast_hook=Fetch(), malign_hook=None, vcf_hook=None, marker=marker,
                                        do_print_alt_subjects=False)
            got, want = mine_lm.PCRproduct(roi=roi, pp=pp), ref.lm.PCRproduct(roi=roi, pp=pp)
            assert (got.roi.seq, got.roi.seq_alt, got.roi.seq_x, got.roi.n) == \
                   (want.roi.seq, want.roi.seq_alt, want.roi.seq_x, want.roi.n)
            mfs = int(rng.choice([0, 20, 60]))
            got.will_it_cut(mine_enz, mfs)
            want.will_it_cut(enzymes=ref_enz, min_fragment_size=mfs)
            assert [(e["name"], e["allele"], float(e["fragl"]), float(e["fragr"])) for e in got.enzymes] == \
                   [(e["name"], e["allele"], float(e["fragl"]), float(e["fragr"])) for e in want.enzymes], (trial, k)
        # the shared dicts ended up with the same last-writer values
        key = lambda e: (e["name"], e.get("allele"), None if "fragl" not in e else float(e["fragl"]))
        assert [key(e) for e in mine_enz] == [key(e) for e in ref_enz]
