"""PCR / HRM / CAPS design flows (reference _lib_pcr.main :222-356, _lib_hrm.main :190-341,
_lib_caps.main :214-389) from the homolog-mapped flank ROIs to the report files, against
tests/golden/pcr_main.json -- produced by the REFERENCE's own code (Region.merge_with_primers,
include_mut_in_included_maps, get_sequence_excluded_regions, design_primers, heterodimers,
add_mutations, PCRproduct, classify, prune, print_report) on a synthetic genome; the picker's answers
in the fixture come from the picker oracle (tests/golden/make_golden.py: golden_pcr_main)."""
import json
import os
import tempfile
import types

import numpy as np
import pytest

from polyoligo_b200 import caps, hrm, lib_markers as lm, lib_primer3 as lp, pcr, primer3_bindings as p3b

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pcr_main.json")
MODULES = {"PCR": pcr, "HRM": hrm, "CAPS": caps}
PRIMER_FIELDS = ("sequence", "start", "stop", "id", "penalty", "end_stability", "hairpin_th", "self_any_th",
                 "self_end_th")


def gold():
    with open(GOLD) as f:
        return json.load(f)


def row_of(pp):
    row = {"dir": pp.dir, "penalty": pp.penalty, "product_size": pp.product_size,
           "compl_any_th": pp.compl_any_th, "compl_end_th": pp.compl_end_th}
    for d, p in pp.primers.items():
        row[d] = {f: getattr(p, f) for f in PRIMER_FIELDS}
    return row


def build_region(case):
    """The state _lib_*.main is in after map_homologs + upload_mutations: a target region with two
    flank ROIs that carry their inclusion maps and mutations."""
    store = lm.SequenceStore({"chrG": case["genome"]})
    marker = None
    if case["marker"]:
        m = case["marker"]
        marker = types.SimpleNamespace(name=m["name"], chrom=m["chrom"], pos1=m["pos1"], ref=m["ref"], alt=m["alt"],
                                       n=len(m["ref"]))
    flanks = {}
    for side, inp in case["inputs"].items():
        roi = lm.ROI(chrom="chrG", start=inp["start"], stop=inp["stop"], blast_hook=store, vcf_hook=True, name="t",
                     marker=marker)
        maps = {}
        for k, excluded in inp["maps"].items():
            v = np.ones(roi.n, dtype=bool)
            v[excluded] = False
            maps[k] = v
        roi.p3_sequence_included_maps = maps
        roi.mutations = [types.SimpleNamespace(pos=p, ref=r, alt=a, aaf=f, indel_size=i) for p, r, a, f, i in
                         inp["mutations"]]
        flanks[side] = roi
    name = "reg%d" % case["index"]
    if case["assay"] == "CAPS":      # _lib_caps.main rebuilds the region around the marker itself
        start = stop = marker.pos1
    else:
        start, stop = case["target"]["start"], case["target"]["stop"]
    return lm.Region(chrom="chrG", start=start, stop=stop, blast_hook=store, vcf_hook=True, name=name, marker=marker,
                     left_roi=flanks["left"], right_roi=flanks["right"])


def cases_of(assay):
    fx = gold()
    out = []
    for k, c in enumerate(fx["cases"]):
        c["index"] = k
        if c["assay"] == assay:
            out.append(c)
    return fx, out


def set_globals(fx, case):
    lp.PRIMER3_GLOBALS.clear()
    p3b.resetP3Globals()
    lp.set_globals(**fx["globals"])
    lp.set_globals(PRIMER_NUM_RETURN=case["depth"])
    if case["assay"] == "HRM":
        lp.set_globals(PRIMER_PRODUCT_SIZE_RANGE=case["size_ranges"])
    lp.set_tm_delta(fx["tm_delta"])


@pytest.mark.parametrize("assay", ["PCR", "HRM", "CAPS"])
def test_merged_region_and_masks_match_reference(assay):
    # Region.merge_with_primers + the '_mut' twins + excluded intervals (host logic, no device)
    fx, cases = cases_of(assay)
    assert cases
    for case in cases:
        region = build_region(case)
        merged = pcr.prepare_region(region, case["fragment_min_size"])
        want = case["merged"]
        assert (merged.start, merged.stop, merged.n, merged.seq) == (want["start"], want["stop"], want["n"], want["seq"])
        assert [list(x) for x in merged.p3_sequence_target] == want["target"]
        assert set(merged.p3_sequence_included_maps) == set(want["excluded_idx"])
        for k, idx in want["excluded_idx"].items():
            assert np.flatnonzero(~merged.p3_sequence_included_maps[k]).tolist() == idx, k
            assert [[int(a), int(b)] for a, b in merged.p3_sequence_excluded_regions[k]] == want["excluded_regions"][k]
        assert [[m.pos, m.ref, m.alt, m.aaf, m.indel_size] for m in merged.mutations] == case["mutations"]


def test_will_it_cut_quirks():
    # cut position: half the site rounded down, except a forward hit on the ALT allele (rounded up);
    # the accepted enzyme dict is the caller's (shared), reference lib_markers.py:471-516
    seq = "ACGT" * 10 + "GGATC" + "TTGCA" * 8
    store = lm.SequenceStore({"c": "N" * 9 + seq})
    marker = types.SimpleNamespace(name="m", chrom="c", pos1=10 + 42, ref="A", alt="C", n=1)
    pp = types.SimpleNamespace(primers={"F": types.SimpleNamespace(start=10), "R": types.SimpleNamespace(stop=9 + len(seq))})
    roi = types.SimpleNamespace(chrom="c", blast_hook=store, malign_hook=None, vcf_hook=None, marker=marker,
                                do_print_alt_subjects=False)
    prod = lm.PCRproduct(roi=roi, pp=pp)
    assert prod.roi.seq == seq and prod.roi.seq_alt[42] == "C" and prod.roi.seq_x[42] == "x"
    enz = [{"name": "five", "supplier": "X", "regex": {"F": "GGATC", "R": "GATCC"}},
           {"name": "none", "supplier": "X", "regex": {"F": "GGGGGG", "R": "CCCCCC"}}]
    prod.will_it_cut(enz, 10)
    assert [e["name"] for e in prod.enzymes] == ["five"]
    assert (enz[0]["allele"], enz[0]["fragl"], enz[0]["fragr"]) == ("REF", 42, len(seq) - 42)     # 40 + 5 // 2
    marker2 = types.SimpleNamespace(name="m", chrom="c", pos1=10 + 42, ref="A", alt="T", n=1)
    store2 = lm.SequenceStore({"c": "N" * 9 + seq[:42] + "C" + seq[43:]})
    roi2 = types.SimpleNamespace(chrom="c", blast_hook=store2, malign_hook=None, vcf_hook=None,
                                 marker=types.SimpleNamespace(name="m", chrom="c", pos1=52, ref="C", alt="A", n=1),
                                 do_print_alt_subjects=False)
    prod2 = lm.PCRproduct(roi=roi2, pp=pp)
    prod2.will_it_cut(enz, 10)
    assert (enz[0]["allele"], enz[0]["fragl"]) == ("ALT", 43)                                      # 40 + (5 + 1) // 2
    assert prod.enzymes[0] is enz[0] and prod.enzymes[0]["allele"] == "ALT"                        # shared dict
    del marker2


def run_flow(assay, regions, case0, per_region_range=False):
    if assay == "PCR":
        return pcr.design_regions(regions, n_primers=5, per_region_range=per_region_range)
    if assay == "HRM":
        return hrm.design_regions(regions, n_primers=5)
    enzymes = [dict(e) for e in case0["enzymes"]]
    return caps.design_regions(regions, enzymes, case0["fragment_min_size"], n_primers=5)


def check_case(assay, case, pcr_, merged, canned=True):
    mod = MODULES[assay]
    got, want = [row_of(pp) for pp in pcr_.pps], json.loads(json.dumps(case["repo"]))
    if not canned:        # the canned answers carry a placeholder end stability (it is not reported)
        for row in got + want:
            for d in ("F", "R"):
                del row[d]["end_stability"]
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g == w, (assay, case["index"], k)
    assert [int(pp.goodness) for pp in pcr_.pps] == case["goodness"]
    assert sum(len(v) for v in pcr_.pps_pruned.values()) == case["n_kept"]
    if assay == "HRM":
        assert [pp.hrm for pp in pcr_.pps] == case["hrm"]
    if assay == "CAPS":
        assert [[e["name"] for e in pp.enzymes] for pp in pcr_.pps] == case["caps"]
    with tempfile.TemporaryDirectory() as tmp:
        mod.print_report(pcr_, os.path.join(tmp, "rep"))
        mod.print_report_header(os.path.join(tmp, "hdr.txt"))
        for name in ("rep.txt", "rep.bed", "hdr.txt"):
            assert open(os.path.join(tmp, name)).read() == case["report"][name], (assay, case["index"], name)


@pytest.mark.gpu
@pytest.mark.parametrize("assay", ["PCR", "HRM", "CAPS"])
def test_flow_matches_reference_on_canned_picker_answers(monkeypatch, assay):
    """Everything around the picker on the device (validity, heterodimers, mutations, Tm_NN, scoring,
    top-n) and the report files byte-identical to the reference's, one batch over the assay's cases."""
    fx, cases = cases_of(assay)
    answers = {}
    for case in cases:
        for c in case["canned"]:
            key = (case["merged"]["seq"], tuple(tuple(x) for x in c["excluded"]))
            answers[key] = {k: (tuple(v) if isinstance(v, list) else v) for k, v in c["answer"]}
    asked = []

    def fake_batch(seq_args_list, global_args=None, device=None):
        out = []
        for a in seq_args_list:
            key = (a["SEQUENCE_TEMPLATE"], tuple((int(x), int(y)) for x, y in a["SEQUENCE_EXCLUDED_REGION"]))
            asked.append((key, [list(x) for x in a["SEQUENCE_TARGET"]], p3b._globals["PRIMER_PRODUCT_SIZE_RANGE"]))
            out.append(dict(answers[key]))
        return out

    monkeypatch.setattr(p3b, "design_primers_batch", fake_batch)
    set_globals(fx, cases[0])
    regions = [build_region(c) for c in cases]
    pcrs, merged = run_flow(assay, regions, cases[0], per_region_range=True)
    for case, pcr_, reg in zip(cases, pcrs, merged):
        check_case(assay, case, pcr_, reg)
    # every picker call carried the region's own target and product size range
    by_template = {c["merged"]["seq"]: c for c in cases}
    assert asked
    for (template, _), target, ranges in asked:
        assert target == by_template[template]["merged"]["target"]
        assert [list(x) for x in ranges] == by_template[template]["size_ranges"]
    if assay == "PCR":      # the reference's process-global range: the first region's, kept for the rest
        del asked[:]
        set_globals(fx, cases[0])
        pcr.design_regions([build_region(c) for c in cases[:2]], n_primers=5)
        assert {tuple(map(tuple, r)) for _, _, r in asked} == {tuple(map(tuple, cases[0]["size_ranges"]))}


@pytest.mark.gpu
@pytest.mark.parametrize("assay", ["PCR", "HRM", "CAPS"])
def test_flow_with_the_gpu_picker_reproduces_the_reference_reports(assay):
    """The same flows with the real GPU picker under the fixture's picker settings (ThermoAnalysis()'s
    salt): the reports are byte-identical to the ones the reference wrote from the picker oracle's answers."""
    fx, cases = cases_of(assay)
    set_globals(fx, cases[0])
    lp.set_globals(PRIMER_SALT_DIVALENT=0.0, PRIMER_DNTP_CONC=0.0)
    try:
        regions = [build_region(c) for c in cases]
        pcrs, merged = run_flow(assay, regions, cases[0], per_region_range=True)
        for case, pcr_, reg in zip(cases, pcrs, merged):
            check_case(assay, case, pcr_, reg, canned=False)
    finally:
        p3b.resetP3Globals()
        lp.PRIMER3_GLOBALS.clear()


def test_marker_and_roi_match_reference():
    # lib_markers.Marker / ROI / upload_mutations against the reference's own objects (host logic)
    import pandas as pd
    with open(os.path.join(os.path.dirname(GOLD), "lib_markers.json")) as f:
        fx = json.load(f)
    store = lm.SequenceStore({"chrM": fx["genome"]})
    for case in fx["cases"]:
        name, pos, ref, alt = case["args"]
        m = lm.Marker(name=name, chrom="chrM", pos=pos, ref_allele=ref, alt_allele=alt)
        assert dict(name=m.name, pos=m.pos, pos1=m.pos1, ref=m.ref, alt=m.alt, variant=m.variant, n=m.n) == case["before"]
        assert m.assert_marker(store) == case["err"]
        m.parse_indels(store)
        assert dict(ref=m.ref, alt=m.alt, variant=m.variant, n=m.n, is_indel=m.is_indel) == case["after"]
        for want in case["rois"]:
            roi = lm.ROI(chrom="chrM", start=want["start"], stop=want["stop"], blast_hook=store, marker=m)
            assert dict(start=roi.start, stop=roi.stop, name=roi.name, seq=roi.seq, seq_alt=roi.seq_alt, seq_x=roi.seq_x,
                        n=roi.n) == want
    name, pos, ref, alt = fx["bad_args"]
    assert lm.Marker(name=name, chrom="chrM", pos=pos, ref_allele=ref, alt_allele=alt).assert_marker(store) == fx["bad_err"]

    class Vcf:
        def fetch_genotypes(self, chrom, start, stop):
            pos = [p for p in (95, 101, 110, 140) if start <= p <= stop]
            return pd.DataFrame({p: [0, 1, 2] for p in pos}), [types.SimpleNamespace(pos=p) for p in pos]

    for want in fx["upload"]:
        mpos = want["marker_pos"]
        m = lm.Marker(name="u", chrom="chrM", pos=mpos, ref_allele=fx["genome"][mpos].upper(), alt_allele="T")
        roi = lm.ROI(chrom="chrM", start=90, stop=150, blast_hook=store, vcf_hook=Vcf(), marker=m)
        roi.upload_mutations()
        assert [x.pos for x in roi.mutations] == want["mutation_pos"]
        assert [int(c) for c in roi.G.columns] == want["columns"]


def test_get_enzymes_reads_the_reference_format(tmp_path):
    table = [{"name": "EcoRI", "supplier": "B", "regex": {"F": "GAATTC", "R": "GAATTC"}},
             {"name": "BsaI", "supplier": "N", "regex": {"F": "GGTCTC", "R": "GAGACC"}}]
    fn = tmp_path / "type2.json"
    fn.write_text(json.dumps(table))
    assert caps.get_enzymes(filename=str(fn)) == table
    assert [e["name"] for e in caps.get_enzymes(included_enzymes=["BsaI"], filename=str(fn))] == ["BsaI"]
    with pytest.raises(FileNotFoundError):
        caps.get_enzymes(filename=str(tmp_path / "absent.json"))


def test_reports_of_empty_designs_match_reference(tmp_path):
    # a region / marker for which no pair survives: the writers' NA line (host only)
    fx = gold()
    for assay, mod in MODULES.items():
        if assay == "PCR":
            p = lp.PCR(chrom="chrG", name="nothing")
        else:
            p = mod.PCR(snp_id="mk0", chrom="chrG", pos=1234, ref="A", alt="")
        mod.print_report(p, str(tmp_path / assay))
        for ext in (".txt", ".bed"):
            assert open(str(tmp_path / assay) + ext).read() == fx["empty_reports"][assay][ext], (assay, ext)
