"""Selection logic around the primer picker (SURVEY 8a rows a5 result parsing, a6, a7) against
fixtures produced by the REFERENCE's own get_primers / design_primers code run on canned
designPrimers answers (tests/golden/make_golden.py: golden_design)."""
import json
import os
import types

import numpy as np
import pytest

from polyoligo_b200 import kasp, lib_primer3 as lp, pcr, primer3_bindings as p3b

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRIMER_FIELDS = ("sequence", "start", "stop", "id", "penalty", "end_stability", "hairpin_th", "self_any_th",
                 "self_end_th")


def gold():
    with open(os.path.join(GOLD, "design.json")) as f:
        return json.load(f)


def row_of(pp):
    row = {"dir": pp.dir, "penalty": pp.penalty, "product_size": pp.product_size,
           "compl_any_th": pp.compl_any_th, "compl_end_th": pp.compl_end_th}
    for d, p in pp.primers.items():
        row[d] = {f: getattr(p, f) for f in PRIMER_FIELDS}
    return row


def canned_dict(items):
    return {k: (tuple(v) if isinstance(v, list) else v) for k, v in items}


def test_result_parsing_matches_reference():
    # get_primers' dict -> PrimerPair conversion: pairs by id, coordinates shifted, *_TM dropped
    fx = gold()
    case = fx["pcr"][0]
    pps = lp.parse_p3_result(canned_dict(case["canned"]), target_start=case["left_start"])
    assert len(pps) == 30 and [pp.primers["F"].id for pp in pps] == list(range(30))
    by_key = {(pp.primers["F"].sequence, pp.primers["R"].sequence): pp for pp in pps}
    assert case["repo"]
    for want in case["repo"]:
        pp = by_key[(want["F"]["sequence"], want["R"]["sequence"])]
        got = row_of(pp)
        got["dir"] = want["dir"]
        assert got == want
        assert pp.primers["F"].tm is None          # excluded from the import
    assert lp.parse_p3_result({"PRIMER_LEFT_0_SEQUENCE": "ACGT", "PRIMER_ERROR": "x"}) == []
    assert p3b.end_stability("AAAGCGCG") == 6.86 and p3b.end_stability("TATAT") == 0.86   # primer3 manual


@pytest.mark.gpu
def test_kasp_design_primers_matches_reference(monkeypatch):
    fx = gold()
    lp.PRIMER3_GLOBALS.clear()
    lp.set_globals(**fx["globals"])
    lp.set_tm_delta(fx["tm_delta"])
    for case in fx["kasp"]:
        marker = types.SimpleNamespace(pos1=case["pos1"], ref=case["ref"], alt=case["alt"])
        hroi = types.SimpleNamespace(seq=case["seq"], start=case["start"], stop=case["stop"], marker=marker,
                                     chrom="chrT", name="m")
        mpps = kasp.get_marker_primers(hroi)
        repo = []
        for canned, excluded, n_primers, want in zip(case["canned"], case["excluded"], case["n_primers"],
                                                     case["repo"]):
            assert len(canned) == len(mpps)
            seen = []

            def fake_batch(seq_args_list, global_args=None, device=None, _c=canned, _seen=seen):
                _seen.extend(seq_args_list)
                return [canned_dict(items) for items in _c]

            monkeypatch.setattr(p3b, "design_primers_batch", fake_batch)
            repo = kasp.design_primers(repo, mpps, hroi, sequence_excluded_region=excluded, n_primers=n_primers)
            # the picker was asked what the reference asks (_lib_kasp.py:571-588)
            for args, mpp in zip(seen, mpps):
                key = "SEQUENCE_PRIMER" if mpp.dir == "F" else "SEQUENCE_PRIMER_REVCOMP"
                assert args[key] == mpp.primers[mpp.dir].sequence and args["SEQUENCE_TEMPLATE"] == case["seq"]
                assert [list(x) for x in args["SEQUENCE_EXCLUDED_REGION"]] == excluded
            assert [row_of(pp) for pp in repo] == want
        # the rest of _lib_kasp.main on that repository: dyes, heterodimers, mutations, classify,
        # prune and the report writer -- byte-identical to the files the reference wrote
        from helpers import VIC, FAM
        muts = [types.SimpleNamespace(**m) for m in case["mutations"]]
        pcr_ = kasp.PCR(snp_id="mk%d" % fx["kasp"].index(case), chrom="chrT", pos=case["pos1"], ref=case["ref"],
                        alt=case["alt"])
        pcr_.pps = repo
        pcr_.add_reporter_dyes([VIC, FAM])
        lp.add_mutations_batch([pcr_], [muts])
        for pp in pcr_.pps:
            lp.add_mutation_strings(pp, muts)
        lp.rank_batch([pcr_], 6, "KASP", reporters=(VIC, FAM))
        assert sum(len(v) for v in pcr_.pps_pruned.values()) == case["n_kept"]
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            kasp.print_report(pcr_, os.path.join(tmp, "rep"))
            kasp.print_report_header(os.path.join(tmp, "hdr.txt"))
            for name in ("rep.txt", "rep.bed", "hdr.txt"):
                assert open(os.path.join(tmp, name)).read() == case["report"][name], name


@pytest.mark.gpu
def test_pcr_design_primers_matches_reference(monkeypatch):
    fx = gold()
    lp.PRIMER3_GLOBALS.clear()
    lp.set_globals(**fx["globals"])
    lp.set_tm_delta(fx["tm_delta"])
    n = 0
    for case in fx["pcr"]:
        monkeypatch.setattr(p3b, "design_primers_batch",
                            lambda seq_args_list, global_args=None, device=None, _c=case: [canned_dict(_c["canned"])])
        roi = types.SimpleNamespace(seq=case["seq"], chrom="chrP",
                                    left_roi=types.SimpleNamespace(start=case["left_start"]))
        repo = pcr.design_primers([], roi, sequence_target=case["target"], sequence_excluded_region=case["excluded"],
                                  n_primers=case["n_primers"], max_unique=case["max_unique"])
        got = [row_of(pp) for pp in repo]
        for g in got:
            g["dir"] = None
        assert got == case["repo"]
        n += len(got)
    assert n > 5


@pytest.mark.gpu
def test_array_pipeline_matches_object_pipeline():
    # BASELINE config 5 shape: the throughput path (flat arrays) makes the decisions of the object
    # path (kasp.design_markers), whose parts the reference fixtures pin
    from polyoligo_b200 import _native, kasp_pipeline as kp
    from helpers import VIC, FAM
    n = 10
    data = kp.synth_config5(n, seed=77)
    lp.PRIMER3_GLOBALS.clear()
    p3b.resetP3Globals()
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
                   PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
                   PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0, PRIMER_PAIR_MAX_DIFF_TM=6.0,
                   PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]], PRIMER_NUM_RETURN=12,
                   # salt as in ThermoAnalysis() so that enough picked primers survive Primer.is_valid
                   PRIMER_SALT_DIVALENT=0.0, PRIMER_DNTP_CONC=0.0)
    lp.set_tm_delta(5.0)
    res = kp.design_markers_arrays(data, n_primers=5, reporters=(VIC, FAM), chunk=4)
    hrois = []
    for k in range(n):
        seq = data["templates"][k].tobytes().decode("ascii")
        start = int(data["starts"][k])
        idx = int(data["marker_idx"][k])
        a0, a1 = data["mut_seg"][k], data["mut_seg"][k + 1]
        muts = [types.SimpleNamespace(pos=int(p), aaf=float(f), indel_size=int(i))
                for p, f, i in zip(data["mut_pos"][a0:a1], data["mut_aaf"][a0:a1], data["mut_indel"][a0:a1])]
        hrois.append(types.SimpleNamespace(
            seq=seq, start=start, stop=start + len(seq) - 1, chrom="chrS", name="m%d" % k, vcf_hook=True,
            mutations=muts, marker=types.SimpleNamespace(pos1=start + idx, ref=seq[idx], alt=data["alt"][k], name="m%d" % k),
            p3_sequence_included_maps={name: data["maps"][name][k].copy() for name in ("mism", "partial_mism", "all")}))
    pcrs = kasp.design_markers(hrois, n_primers=5, reporters=(VIC, FAM))
    total = 0
    for k, pcr in enumerate(pcrs):
        lo, hi = res["seg"][k], res["seg"][k + 1]
        got_repo = list(zip(_native.unpack(res["f"][lo:hi]), _native.unpack(res["r"][lo:hi]),
                            _native.unpack(res["a"][lo:hi]), ["FR"[d] for d in res["dir"][lo:hi]]))
        want_repo = [(pp.primers["F"].sequence, pp.primers["R"].sequence, pp.primers["A"].sequence, pp.dir)
                     for pp in pcr.pps]
        assert got_repo == want_repo, k
        assert list(res["goodness"][lo:hi]) == [pp.goodness for pp in pcr.pps]
        want_sel = [id(pp) for score in sorted(pcr.pps_pruned, reverse=True) for pp in pcr.pps_pruned[score]]
        got_sel = [id(pcr.pps[int(i - lo)]) for i in res["sel"][k, :res["n_sel"][k]]]
        assert got_sel == want_sel, k
        total += len(want_repo)
    assert total > 30
    p3b.resetP3Globals()


@pytest.mark.gpu
def test_array_pipeline_matches_cpu_oracle_pipeline():
    # the whole config-5 pipeline, GPU context against the CPU oracle context (oracle/cpu_pipeline.py:
    # C-oracle thermodynamics, lazy picker, Python scoring), reference defaults (depth 100, top 10)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from oracle.cpu_pipeline import CpuContext
    from polyoligo_b200 import kasp_pipeline as kp
    from helpers import VIC, FAM
    lp.PRIMER3_GLOBALS.clear()
    p3b.resetP3Globals()
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
                   PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
                   PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0, PRIMER_PAIR_MAX_DIFF_TM=6.0,
                   PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]], PRIMER_NUM_RETURN=100)
    lp.set_tm_delta(5.0)
    data = kp.synth_config5(12, seed=31)
    gpu = kp.design_markers_arrays(data, n_primers=10, reporters=(VIC, FAM), chunk=5)
    cpu = kp.design_markers_arrays(data, n_primers=10, reporters=(VIC, FAM), chunk=5, ctx=CpuContext(nthreads=8))
    assert len(gpu["f"]) == len(cpu["f"]) > 100
    for k in ("seg", "dir", "goodness", "qbits", "sel", "n_sel"):
        assert np.array_equal(gpu[k], cpu[k]), k
    for k in ("f", "r", "a"):
        assert gpu[k].tobytes() == cpu[k].tobytes(), k
    p3b.resetP3Globals()


@pytest.mark.gpu
def test_pcr_design_primers_end_to_end_with_the_real_picker():
    # no canned answers: picker + validity + max_unique through the public mirror (properties only)
    rng = np.random.default_rng(4)
    seq = "".join(np.array(list("ACGT"))[rng.integers(0, 4, 400)])
    lp.PRIMER3_GLOBALS.clear()
    p3b.resetP3Globals()
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
                   PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
                   PRIMER_PAIR_MAX_DIFF_TM=6.0, PRIMER_PRODUCT_SIZE_RANGE=[[80, 200]], PRIMER_NUM_RETURN=40,
                   PRIMER_SALT_DIVALENT=0.0, PRIMER_DNTP_CONC=0.0)
    lp.set_tm_delta(5.0)
    roi = types.SimpleNamespace(seq=seq, chrom="chrQ", left_roi=types.SimpleNamespace(start=7000))
    repo = pcr.design_primers([], roi, sequence_target=[190, 10], sequence_excluded_region=[[0, 10]], n_primers=6,
                              max_unique=2)
    assert 0 < len(repo) <= 6
    used = {}
    for pp in repo:
        f, r = pp.primers["F"], pp.primers["R"]
        assert seq[f.start - 7000:f.stop - 7000 + 1] == f.sequence
        assert p3b._revcomp(seq[r.start - 7000:r.stop - 7000 + 1]) == r.sequence
        assert f.stop - 7000 < 190 and r.start - 7000 >= 200 and f.start - 7000 >= 10
        assert 80 <= pp.product_size <= 200 and pp.product_size == r.stop - f.start + 1
        assert 55.0 <= f.tm <= 65.0 and 55.0 <= r.tm <= 65.0 and pp.chrom == "chrQ"
        for p in (f, r):
            used[p.sequence] = used.get(p.sequence, 0) + 1
    assert max(used.values()) <= 2
    assert [pp.penalty for pp in repo] == sorted(pp.penalty for pp in repo)
    p3b.resetP3Globals()
