"""GPU tests at the SHAPES BASELINE.json names ("Next round" 1d/1e):

* config 2: ``polyoligo-pcr`` with sample_data/primer3_example.yaml -- kb-scale templates (5.5 kb and
  15.5 kb: 5 / 15 kb targets + 2 x 250 nt flanks, reference cli_pcr.py:161), product 500-1000,
  ``SEQUENCE_PRIMER_REVCOMP`` fixed;
* CAPS / HRM: both primers picked on 1000-nt templates with the assay's own product ranges
  (reference data/PRIMER3_CAPS.yaml, PRIMER3_HRM.yaml);
* config 5 at the reference's defaults (depth 100, top 10) on 100+ markers, array pipeline against
  the object pipeline whose parts the reference fixtures pin, and the chromosome-start edge case
  (``hroi.start <= PRIMER_MAX_SIZE``, reference _lib_kasp.py:681-684, test_data/markers_lim.txt).

The picker is compared with oracle/p3_oracle.py (parity with libprimer3 itself is unpinned)."""
import os
import sys
import types

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import p3_oracle  # noqa: E402
from polyoligo_b200 import _native, kasp, lib_primer3 as lp, primer3_bindings as p3b  # noqa: E402
from test_p3 import oracle_settings, rand_template  # noqa: E402

pytestmark = pytest.mark.gpu

# reference src/polyoligo/data/PRIMER3_PCR.yaml with sample_data/primer3_example.yaml laid over it
# (cli_pcr.py merges the user YAML over the packaged defaults); PRIMER_NUM_RETURN = depth = 100
PCR_EXAMPLE = {"PRIMER_OPT_SIZE": 20, "PRIMER_MIN_SIZE": 18, "PRIMER_MAX_SIZE": 25, "PRIMER_OPT_TM": 60.0,
               "PRIMER_MIN_TM": 55.0, "PRIMER_MAX_TM": 65.0, "PRIMER_MIN_GC": 20.0, "PRIMER_MAX_GC": 80.0,
               "PRIMER_MAX_POLY_X": 100, "PRIMER_SALT_MONOVALENT": 50.0, "PRIMER_DNA_CONC": 50.0,
               "PRIMER_MAX_NS_ACCEPTED": 0, "PRIMER_PAIR_MAX_DIFF_TM": 6.0,
               "PRIMER_PRODUCT_SIZE_RANGE": [[500, 1000]], "PRIMER_NUM_RETURN": 100}
HRM = {**PCR_EXAMPLE, "PRIMER_PRODUCT_SIZE_RANGE": [[40, 75], [40, 150]]}        # data/PRIMER3_HRM.yaml
CAPS = {**PCR_EXAMPLE, "PRIMER_PRODUCT_SIZE_RANGE": [[500, 1000], [250, 2000]]}  # data/PRIMER3_CAPS.yaml


def _check(res, ref, two_sided):
    assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref)
    for i, e in enumerate(ref):
        assert res["PRIMER_LEFT_%d" % i] == (e["left"]["start"], e["left"]["len"]), i
        assert res["PRIMER_RIGHT_%d" % i] == (e["right"]["start"], e["right"]["len"]), i
        assert res["PRIMER_LEFT_%d_SEQUENCE" % i] == e["left"]["seq"]
        assert res["PRIMER_PAIR_%d_PENALTY" % i] == e["pair_penalty"]
        assert res["PRIMER_PAIR_%d_COMPL_ANY_TH" % i] == e["compl_any_th"]
        assert res["PRIMER_PAIR_%d_COMPL_END_TH" % i] == e["compl_end_th"]
        assert res["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] == e["product_size"]


@pytest.mark.parametrize("length,seed", [(5500, 11), (15500, 12)])
def test_config2_shape_fixed_right_primer_on_kb_templates(ctx, oracle, length, seed):
    rng = np.random.default_rng(seed)
    t = rand_template(rng, length, gc=0.42)
    so = oracle_settings(PCR_EXAMPLE, 100)
    full = {**p3_oracle.DEFAULTS, **so}
    # a right primer that is itself acceptable, far enough inside for 500-1000 nt products
    right = None
    for start in range(length // 2, length // 2 + 400):
        for ln in range(18, 26):
            r = p3_oracle.cheap_record(full, t, [0] * len(t), "R", start, ln)
            if r:
                p3_oracle.add_self(oracle, full, [r], 4)
                if r["self_ok"] and t.count(t[start - ln + 1:start + 1]) == 1:
                    right = r
                    break
        if right:
            break
    assert right is not None
    ref = p3_oracle.design(oracle, so, t, right=right["seq"], nthreads=16)
    p3b.resetP3Globals()
    p3b.setP3Globals(PCR_EXAMPLE)
    try:
        res = p3b.designPrimers({"SEQUENCE_TEMPLATE": t, "SEQUENCE_PRIMER_REVCOMP": right["seq"]})
    finally:
        p3b.resetP3Globals()
    assert len(ref) == 100                      # depth reached: the first 100 of thousands of valid pairs
    _check(res, ref, False)
    assert all(500 <= res["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] <= 1000 for i in range(100))
    assert {res["PRIMER_RIGHT_%d_SEQUENCE" % i] for i in range(100)} == {right["seq"]}


@pytest.mark.parametrize("name,g,excluded", [("HRM", HRM, []), ("CAPS", CAPS, [(0, 180), (820, 180)])])
def test_caps_hrm_shape_two_sided_on_1000nt(ctx, oracle, name, g, excluded):
    # marker in the middle of a 1000-nt region (reference _lib_caps.py:256-275, _lib_hrm.py), target = the marker
    rng = np.random.default_rng(5 if name == "HRM" else 6)
    t = rand_template(rng, 1000, gc=0.45)
    target = (499, 1)
    so = oracle_settings(g, 100)
    ref = p3_oracle.design(oracle, so, t, excluded=excluded, target=target, nthreads=16)
    p3b.resetP3Globals()
    p3b.setP3Globals(g)
    try:
        args = {"SEQUENCE_TEMPLATE": t, "SEQUENCE_TARGET": list(target)}
        if excluded:
            args["SEQUENCE_EXCLUDED_REGION"] = [list(x) for x in excluded]
        res = p3b.designPrimers(args)
    finally:
        p3b.resetP3Globals()
    assert len(ref) == 100
    _check(res, ref, True)


def _hrois(data, n):
    out = []
    for k in range(n):
        seq = data["templates"][k].tobytes().decode("ascii")
        start = int(data["starts"][k])
        idx = int(data["marker_idx"][k])
        a0, a1 = data["mut_seg"][k], data["mut_seg"][k + 1]
        muts = [types.SimpleNamespace(pos=int(p), aaf=float(f), indel_size=int(i))
                for p, f, i in zip(data["mut_pos"][a0:a1], data["mut_aaf"][a0:a1], data["mut_indel"][a0:a1])]
        out.append(types.SimpleNamespace(
            seq=seq, start=start, stop=start + len(seq) - 1, chrom="chrS", name="m%d" % k, vcf_hook=True,
            mutations=muts, marker=types.SimpleNamespace(pos1=start + idx, ref=seq[idx], alt=data["alt"][k], name="m%d" % k),
            p3_sequence_included_maps={name: data["maps"][name][k].copy() for name in ("mism", "partial_mism", "all")}))
    return out


def _kasp_defaults():
    lp.PRIMER3_GLOBALS.clear()
    p3b.resetP3Globals()
    # reference data/PRIMER3_KASP.yaml + cli_kasp.py defaults: depth 100 (PRIMER_NUM_RETURN), n 10, tm_delta 5
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
                   PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
                   PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0, PRIMER_MAX_NS_ACCEPTED=0,
                   PRIMER_PAIR_MAX_DIFF_TM=6.0, PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]],
                   PRIMER_NUM_RETURN=100)
    lp.set_tm_delta(5.0)


def _compare_array_and_object(data, n, res, pcrs):
    total = sel_total = 0
    for k, pcr in enumerate(pcrs):
        lo, hi = res["seg"][k], res["seg"][k + 1]
        got_repo = list(zip(_native.unpack(res["f"][lo:hi]), _native.unpack(res["r"][lo:hi]),
                            _native.unpack(res["a"][lo:hi]), ["FR"[d] for d in res["dir"][lo:hi]]))
        want_repo = [(pp.primers["F"].sequence, pp.primers["R"].sequence, pp.primers["A"].sequence, pp.dir)
                     for pp in pcr.pps]
        assert got_repo == want_repo, k
        assert list(res["goodness"][lo:hi]) == [pp.goodness for pp in pcr.pps], k
        want_sel = [id(pp) for score in sorted(pcr.pps_pruned, reverse=True) for pp in pcr.pps_pruned[score]]
        got_sel = [id(pcr.pps[int(i - lo)]) for i in res["sel"][k, :res["n_sel"][k]]]
        assert got_sel == want_sel, k
        total += len(want_repo)
        sel_total += len(want_sel)
    return total, sel_total


def test_array_pipeline_matches_object_pipeline_at_reference_defaults():
    from polyoligo_b200 import kasp_pipeline as kp
    from helpers import VIC, FAM
    n = 104
    data = kp.synth_config5(n, seed=2024)
    _kasp_defaults()
    try:
        res = kp.design_markers_arrays(data, n_primers=10, reporters=(VIC, FAM), chunk=40)
        pcrs = kasp.design_markers(_hrois(data, n), n_primers=10, reporters=(VIC, FAM))
        total, sel_total = _compare_array_and_object(data, n, res, pcrs)
    finally:
        p3b.resetP3Globals()
    assert total > 2000 and sel_total > 300


def test_marker_near_the_chromosome_start(ctx):
    # hroi.start <= PRIMER_MAX_SIZE: the window step `mmap[min(1, start-25):max(len, start+25)] = True`
    # gets a zero or NEGATIVE slice start, i.e. numpy slices from the END (reference _lib_kasp.py:681-684;
    # the reference's own edge-case input is test_data/markers_lim.txt): part of the homolog / mutation
    # masks then survives.  The array pipeline must make the object pipeline's decisions there too.
    from polyoligo_b200 import kasp_pipeline as kp
    from helpers import VIC, FAM
    n = 8
    data = kp.synth_config5(n, seed=99)
    shift = data["starts"] - np.array([3, 10, 20, 24, 25, 26, 27, 1], dtype=np.int64)
    data["mut_pos"] = data["mut_pos"] - np.repeat(shift, np.diff(data["mut_seg"]))
    data["starts"] = data["starts"] - shift          # regions start at 3, 10, 20, 24, 25, 26, 27, 1
    _kasp_defaults()
    try:
        res = kp.design_markers_arrays(data, n_primers=10, reporters=(VIC, FAM), chunk=8)
        pcrs = kasp.design_markers(_hrois(data, n), n_primers=10, reporters=(VIC, FAM))
        total, _ = _compare_array_and_object(data, n, res, pcrs)
    finally:
        p3b.resetP3Globals()
    assert total > 100


# ---- a4 on the device and the whole design behind one C call (po_kasp_masks_batch, po_kasp_design_batch) ----
def _bits_to_idx(words, n):
    bits = np.unpackbits(np.ascontiguousarray(words, dtype="<u4").view(np.uint8), bitorder="little")
    return [int(x) for x in np.flatnonzero(bits[:n])]


def test_mask_kernel_matches_reference_fixture(ctx):
    # tests/golden/masks.json is written by the reference's own include_mut_in_included_maps, the KASP
    # window step and get_sequence_excluded_regions (incl. regions that start 3 / 20 / 26 nt into a chromosome)
    import json
    from polyoligo_b200 import kasp_pipeline as kp, lib_utils
    with open(os.path.join(os.path.dirname(__file__), "golden", "masks.json")) as f:
        cases = json.load(f)["cases"]
    n, L = len(cases), cases[0]["n"]
    maps = {k: np.ones((n, L), dtype=bool) for k in ("mism", "partial_mism", "all")}
    for k, c in enumerate(cases):
        maps["mism"][k, c["mism"]] = False
        maps["partial_mism"][k, c["partial_mism"]] = False
    starts = np.array([c["start"] for c in cases], dtype=np.int64)
    seg = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(c["mutation_pos"]) for c in cases], out=seg[1:])
    pos = np.array([p for c in cases for p in c["mutation_pos"]], dtype=np.int64)
    names = ["mism_mut", "mism", "partial_mism_mut", "partial_mism", "all_mut", "all"]
    plain = ctx.kasp_masks(starts, L, kp.pack_maps(maps), pos, seg, kasp_window=False)
    kaspw = ctx.kasp_masks(starts, L, kp.pack_maps(maps), pos, seg, kasp_window=True, max_size=25)
    checked = 0
    for k, c in enumerate(cases):
        for t, name in enumerate(names):
            assert _bits_to_idx(plain[t, k], L) == c["excluded_idx_after_mut"][name], (k, name)
            idx = _bits_to_idx(kaspw[t, k], L)
            regions = [[int(a), int(b)] for a, b in lib_utils.list_2_ranges(idx)] if idx else []
            assert regions == c["kasp_excluded_regions"][name], (k, name)
            checked += len(idx)
    assert checked > 50
    three = ctx.kasp_masks(starts, L, kp.pack_maps(maps), kasp_window=False)      # no VCF hook: three types
    assert np.array_equal(three, plain[1::2])


def test_fused_design_call_matches_array_pipeline(ctx):
    from polyoligo_b200 import kasp_pipeline as kp, sharding
    from helpers import VIC, FAM
    n = 300
    data = kp.synth_config5(n, seed=4242)
    # a few regions at a chromosome start: part of the masks survives the window step there, so the later
    # search types ask new questions and the merge has to deduplicate
    near = np.array([3, 10, 20, 24, 25, 26], dtype=np.int64)
    shift = np.zeros(n, dtype=np.int64)
    shift[:len(near)] = data["starts"][:len(near)] - near
    data["mut_pos"] = data["mut_pos"] - np.repeat(shift, np.diff(data["mut_seg"]))
    data["starts"] = data["starts"] - shift
    _kasp_defaults()
    try:
        res = kp.design_markers_arrays(data, n_primers=10, reporters=(VIC, FAM), chunk=128)
        want, want_n = sharding.compact_selection(res, 10)
        got, got_n, merged = kp.design_markers_fused(data, n_primers=10, reporters=(VIC, FAM))
        # the same in blocks of 64 markers on three host threads / contexts / streams
        blk, blk_n, blk_merged = kp.design_markers_fused(data, n_primers=10, reporters=(VIC, FAM), workers=3, block=64)
        # and without the VCF hook (three search types)
        data3 = dict(data, mut_seg=None)
        res3 = kp.design_markers_arrays(data3, n_primers=10, reporters=(VIC, FAM), chunk=128)
        got3, got3_n, merged3 = kp.design_markers_fused(data3, n_primers=10, reporters=(VIC, FAM))
    finally:
        p3b.resetP3Globals()
    assert np.array_equal(got_n, want_n) and want_n.sum() > 1000
    assert np.array_equal(merged, np.diff(res["seg"]))
    assert got.tobytes() == want.tobytes()
    assert blk.tobytes() == want.tobytes() and np.array_equal(blk_n, want_n) and np.array_equal(blk_merged, merged)
    want3, want3_n = sharding.compact_selection(res3, 10)
    assert np.array_equal(got3_n, want3_n) and np.array_equal(merged3, np.diff(res3["seg"]))
    assert got3.tobytes() == want3.tobytes()
