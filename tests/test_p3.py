"""Primer picker (designPrimers replacement): GPU path against the
brute-force CPU oracle (oracle/p3_oracle.py).  libprimer3 itself is absent: parity with the
real library is unpinned, these tests pin the GPU path to the restated semantics."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import p3_oracle  # noqa: E402

import polyoligo_b200 as po  # noqa: E402
from polyoligo_b200 import _native, primer3_bindings as p3b  # noqa: E402

KASP_GLOBALS = {   # reference src/polyoligo/data/PRIMER3_KASP.yaml
    "PRIMER_OPT_SIZE": 20, "PRIMER_MIN_SIZE": 18, "PRIMER_MAX_SIZE": 25, "PRIMER_OPT_TM": 60.0,
    "PRIMER_MIN_TM": 55.0, "PRIMER_MAX_TM": 65.0, "PRIMER_MIN_GC": 20.0, "PRIMER_MAX_GC": 80.0,
    "PRIMER_MAX_POLY_X": 100, "PRIMER_SALT_MONOVALENT": 50.0, "PRIMER_DNA_CONC": 50.0,
    "PRIMER_MAX_NS_ACCEPTED": 0, "PRIMER_MAX_SELF_ANY": 12, "PRIMER_MAX_SELF_END": 8,
    "PRIMER_PAIR_MAX_DIFF_TM": 6.0, "PRIMER_PAIR_MAX_COMPL_ANY": 12, "PRIMER_PAIR_MAX_COMPL_END": 8,
    "PRIMER_PRODUCT_SIZE_RANGE": [[90, 150], [100, 200], [50, 250]], "PRIMER_FIRST_BASE_INDEX": 0,
    "PRIMER_PICK_INTERNAL_OLIGO": 0, "PRIMER_INTERNAL_MAX_SELF_END": 8, "PRIMER_INTERNAL_MAX_POLY_X": 100,
}


def oracle_settings(g, num_return):
    return dict(min_size=g["PRIMER_MIN_SIZE"], opt_size=g["PRIMER_OPT_SIZE"], max_size=g["PRIMER_MAX_SIZE"],
                num_return=num_return, min_tm=g["PRIMER_MIN_TM"], opt_tm=g["PRIMER_OPT_TM"], max_tm=g["PRIMER_MAX_TM"],
                min_gc=g["PRIMER_MIN_GC"], max_gc=g["PRIMER_MAX_GC"], max_poly_x=g["PRIMER_MAX_POLY_X"],
                pair_max_diff_tm=g["PRIMER_PAIR_MAX_DIFF_TM"],
                size_ranges=tuple(tuple(x) for x in g["PRIMER_PRODUCT_SIZE_RANGE"]))


def rand_template(rng, n, gc=0.45):
    p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    return "".join(np.array(list("ACGT"))[rng.choice(4, n, p=p)])


# ---- CPU: host logic ---------------------------------------------------------------------------
def test_settings_struct_matches_header():
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "polyoligo_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n    int32_t min_size, opt_size, max_size;"):hdr.index("} po_p3_settings;")]
    for name, _ in _native.P3Settings._fields_:
        assert name in body, name
    assert C.sizeof(_native.P3Settings) == 248


def test_globals_to_settings_and_job_masks():
    p3b.resetP3Globals()
    p3b.setP3Globals({**KASP_GLOBALS, "PRIMER_NUM_RETURN": 7})
    s = p3b.settings_from_globals()
    assert (s.min_size, s.opt_size, s.max_size, s.num_return) == (18, 20, 25, 7)
    assert (s.dv_conc, s.dntp_conc) == (1.5, 0.6)           # libprimer3 defaults the YAML does not override
    assert s.n_size_ranges == 3 and list(s.size_lo[:3]) == [90, 100, 50] and list(s.size_hi[:3]) == [150, 200, 250]
    with pytest.raises(ValueError):
        p3b.setP3Globals({"PRIMER_MAX_TEMPLATE_MISPRIMING": 12})
    t = "ACGT" * 30
    j = p3b._Job({"SEQUENCE_TEMPLATE": t, "SEQUENCE_EXCLUDED_REGION": [[5, 3], [100, 50]], "SEQUENCE_TARGET": [40, 2],
                  "SEQUENCE_PRIMER": t[10:30], "SEQUENCE_PRIMER_REVCOMP": p3b._revcomp(t[60:81])})
    assert j.mask[5:8].all() and j.mask[100:].all() and j.mask[40:42].all() and j.mask.sum() == 3 + 20 + 2
    assert j.left == (2, 20)            # first occurrence, like strstr
    assert j.right == (0 + 21 - 1, 21)
    assert p3b._Job({"SEQUENCE_TEMPLATE": t, "SEQUENCE_PRIMER": "GGGGGGGGGGGGGGGGGGGG"}).error
    p3b.resetP3Globals()


def test_oracle_design_properties(oracle):
    rng = np.random.default_rng(3)
    t = rand_template(rng, 160)
    s = oracle_settings(KASP_GLOBALS, 6)
    s["size_ranges"] = ((60, 90), (50, 140))
    res = p3_oracle.design(oracle, s, t, excluded=[(0, 5)], nthreads=4)
    assert 0 < len(res) <= 6
    keys = [(r["size_range"], r["pair_penalty"]) for r in res]
    assert keys == sorted(keys)
    seen = set()
    for r in res:
        l, rr = r["left"], r["right"]
        assert t[l["start"]:l["start"] + l["len"]] == l["seq"]
        assert p3b._revcomp(t[rr["start"] - rr["len"] + 1:rr["start"] + 1]) == rr["seq"]
        assert r["product_size"] == rr["start"] - l["start"] + 1
        lo, hi = s["size_ranges"][r["size_range"]]
        assert lo <= r["product_size"] <= hi and l["start"] >= 5
        assert abs(l["tm"] - rr["tm"]) <= 6.0 and r["pair_penalty"] == l["penalty"] + rr["penalty"]
        assert (l["start"], l["len"], rr["start"], rr["len"]) not in seen
        seen.add((l["start"], l["len"], rr["start"], rr["len"]))


# ---- GPU: parity with the oracle ----------------------------------------------------------------
def _gpu_settings(num_return):
    p3b.resetP3Globals()
    p3b.setP3Globals({**KASP_GLOBALS, "PRIMER_NUM_RETURN": num_return})
    return p3b.settings_from_globals()


def _same_oligo(rec, seq, ref):
    assert seq == ref["seq"]
    assert (int(rec["start"]), int(rec["len"])) == (ref["start"], ref["len"])
    for f in ("tm", "gc_percent", "penalty"):
        assert float(rec[f]) == ref[f], f


@pytest.mark.gpu
def test_candidate_lists_match_oracle(ctx):
    rng = np.random.default_rng(8)
    templates = [rand_template(rng, n) for n in (140, 90, 30, 17, 200)]
    templates[1] = templates[1][:40] + "N" + templates[1][41:]
    masks = [np.zeros(len(t), np.uint8) for t in templates]
    masks[0][50:70] = 1
    masks[4][:3] = 1
    s = _gpu_settings(5)
    list_off, recs = ctx.p3_candidates(s, templates, masks)
    seqs = _native.unpack(recs["seq"])
    so = oracle_settings(KASP_GLOBALS, 5)
    so = {**p3_oracle.DEFAULTS, **so}
    total = 0
    for k, (t, m) in enumerate(zip(templates, masks)):
        for strand, side in enumerate("LR"):
            ref = p3_oracle.candidates(so, t, list(m), side)
            lo, hi = list_off[2 * k + strand], list_off[2 * k + strand + 1]
            assert hi - lo == len(ref), (k, side)
            for i, r in enumerate(ref):
                _same_oligo(recs[lo + i], seqs[lo + i], r)
                assert recs[lo + i]["set"] == k
            total += len(ref)
    assert total > 200


def _check_pairs(rows, ref, what):
    assert len(rows) == len(ref), what
    for (l, lseq, r, rseq, pq, ca, ce, size, rng_), e in zip(rows, ref):
        _same_oligo(l, lseq, e["left"])
        _same_oligo(r, rseq, e["right"])
        assert float(pq) == e["pair_penalty"] and int(size) == e["product_size"] and int(rng_) == e["size_range"], what
        assert float(ca) == e["compl_any_th"] and float(ce) == e["compl_end_th"], what
        for rec, o in ((l, e["left"]), (r, e["right"])):
            for f in ("self_any_th", "self_end_th", "hairpin_th"):
                assert float(rec[f]) == o[f], (what, f)


@pytest.mark.gpu
def test_fixed_primer_design_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(12)
    s = _gpu_settings(12)
    so = oracle_settings(KASP_GLOBALS, 12)
    templates = [rand_template(rng, 320), rand_template(rng, 260, gc=0.5)]
    masks = [np.zeros(len(t), np.uint8) for t in templates]
    masks[0][200:215] = 1
    tasks, refs = [], []
    for k, t in enumerate(templates):
        full = {**p3_oracle.DEFAULTS, **so}
        lefts = [r for r in p3_oracle.candidates(full, t, list(masks[k]), "L") if 100 <= r["start"] <= 140][:3]
        rights = [r for r in p3_oracle.candidates(full, t, list(masks[k]), "R") if 150 <= r["start"] <= 190][:3]
        excluded = [(200, 15)] if k == 0 else []
        for r in lefts:
            tasks.append((k, 1, r["start"], r["len"]))
            refs.append(p3_oracle.design(oracle, so, t, excluded=excluded, left=r["seq"]))
        for r in rights:
            tasks.append((k, 2, r["start"], r["len"]))
            refs.append(p3_oracle.design(oracle, so, t, excluded=excluded, right=r["seq"]))
        # a given primer that violates the Tm limits yields nothing
        tasks.append((k, 1, 5, 18))
        refs.append(p3_oracle.design(oracle, so, t, excluded=excluded, left=t[5:23]))
    tasks = np.array(tasks, dtype=_native.P3_TASK_DTYPE)
    fixed, pairs, n_pairs = ctx.p3_design_fixed(s, templates, tasks, masks)
    assert sum(len(r) for r in refs) > 60
    for t in range(len(tasks)):
        n = int(n_pairs[t])
        others = pairs[t, :n]["other"]
        oseq = _native.unpack(others["seq"]) if n else []
        fseq = _native.unpack(fixed[t:t + 1]["seq"])[0]
        rows = []
        for i in range(n):
            p = pairs[t, i]
            if tasks[t]["side"] == 1:
                rows.append((fixed[t], fseq, others[i], oseq[i], p["pair_penalty"], p["compl_any_th"],
                             p["compl_end_th"], p["product_size"], p["size_range"]))
            else:
                rows.append((others[i], oseq[i], fixed[t], fseq, p["pair_penalty"], p["compl_any_th"],
                             p["compl_end_th"], p["product_size"], p["size_range"]))
        _check_pairs(rows, refs[t], "task %d" % t)


@pytest.mark.gpu
def test_design_primers_dict_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(5)
    t = rand_template(rng, 170)
    g = {**KASP_GLOBALS, "PRIMER_NUM_RETURN": 8, "PRIMER_PRODUCT_SIZE_RANGE": [[70, 100], [60, 150]]}
    p3b.resetP3Globals()
    p3b.setP3Globals(g)
    so = oracle_settings(g, 8)
    # both primers picked, with a target and an excluded region (the PCR / HRM / CAPS call shape)
    res = p3b.designPrimers({"SEQUENCE_TEMPLATE": t, "SEQUENCE_TARGET": [80, 4], "SEQUENCE_EXCLUDED_REGION": [[0, 8]]})
    ref = p3_oracle.design(oracle, so, t, excluded=[(0, 8)], target=(80, 4))
    assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref) > 0
    for i, e in enumerate(ref):
        assert res["PRIMER_LEFT_%d_SEQUENCE" % i] == e["left"]["seq"]
        assert res["PRIMER_RIGHT_%d_SEQUENCE" % i] == e["right"]["seq"]
        assert res["PRIMER_LEFT_%d" % i] == (e["left"]["start"], e["left"]["len"])
        assert res["PRIMER_RIGHT_%d" % i] == (e["right"]["start"], e["right"]["len"])
        assert res["PRIMER_PAIR_%d_PENALTY" % i] == e["pair_penalty"]
        assert res["PRIMER_PAIR_%d_COMPL_ANY_TH" % i] == e["compl_any_th"]
        assert res["PRIMER_PAIR_%d_COMPL_END_TH" % i] == e["compl_end_th"]
        assert res["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] == e["product_size"]
        assert res["PRIMER_LEFT_%d_TM" % i] == e["left"]["tm"] and res["PRIMER_RIGHT_%d_HAIRPIN_TH" % i] == e["right"]["hairpin_th"]
    # the KASP call shape through the same entry point
    left = ref[0]["left"]["seq"]
    res = p3b.designPrimers({"SEQUENCE_TEMPLATE": t, "SEQUENCE_PRIMER": left})
    ref2 = p3_oracle.design(oracle, so, t, left=left)
    assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref2) > 0
    assert [res["PRIMER_RIGHT_%d_SEQUENCE" % i] for i in range(len(ref2))] == [e["right"]["seq"] for e in ref2]
    assert p3b.designPrimers({"SEQUENCE_TEMPLATE": t, "SEQUENCE_PRIMER": "ACGTACGTACGTACGTACGTAA"}).get("PRIMER_ERROR")
    p3b.resetP3Globals()


def _rows_for_task(fixed, pairs, n_pairs, tasks, t):
    n = int(n_pairs[t])
    others = pairs[t, :n]["other"]
    oseq = _native.unpack(others["seq"]) if n else []
    fseq = _native.unpack(fixed[t:t + 1]["seq"])[0]
    rows = []
    for i in range(n):
        p = pairs[t, i]
        tail = (p["pair_penalty"], p["compl_any_th"], p["compl_end_th"], p["product_size"], p["size_range"])
        if tasks[t]["side"] == 1:
            rows.append((fixed[t], fseq, others[i], oseq[i]) + tail)
        else:
            rows.append((others[i], oseq[i], fixed[t], fseq) + tail)
    return rows


@pytest.mark.gpu
def test_fixed_primer_design_depth_100_and_ties(ctx, oracle):
    # the reference's default search depth (cli_kasp.py --depth 100): several proposal rounds per task
    rng = np.random.default_rng(41)
    s = _gpu_settings(100)
    so = oracle_settings(KASP_GLOBALS, 100)
    full = {**p3_oracle.DEFAULTS, **so}
    t1 = rand_template(rng, 500, gc=0.47)
    # a repeated block: candidates with identical sequence, Tm and penalty at two positions, so
    # pair penalties tie and the order falls to the position rules; a small depth cuts inside ties
    block = rand_template(rng, 45, gc=0.5)
    t2 = rand_template(rng, 130, gc=0.5) + block + rand_template(rng, 25) + block + rand_template(rng, 40)
    cases = []
    for k, (t, depth) in enumerate(((t1, 100), (t2, 100), (t2, 3), (t2, 7))):
        lefts = [r for r in p3_oracle.candidates(full, t, [0] * len(t), "L") if 40 <= r["start"] <= 110][:2]
        rights = [r for r in p3_oracle.candidates(full, t, [0] * len(t), "R") if len(t) - 60 <= r["start"]][:1]
        cases.append((t, depth, lefts, rights))
    n_checked = 0
    for t, depth, lefts, rights in cases:
        s.num_return = depth
        so["num_return"] = depth
        tasks = np.array([(0, 1, r["start"], r["len"]) for r in lefts] + [(0, 2, r["start"], r["len"]) for r in rights],
                         dtype=_native.P3_TASK_DTYPE)
        fixed, pairs, n_pairs = ctx.p3_design_fixed(s, [t], tasks)
        for i, r in enumerate(lefts + rights):
            kw = {"left": r["seq"]} if i < len(lefts) else {"right": r["seq"]}
            if t.upper().find(r["seq"] if i < len(lefts) else p3b._revcomp(r["seq"])) != \
                    (r["start"] if i < len(lefts) else r["start"] - r["len"] + 1):
                continue       # the given primer itself sits in the repeated block: the oracle finds the first copy
            ref = p3_oracle.design(oracle, so, t, **kw)
            _check_pairs(_rows_for_task(fixed, pairs, n_pairs, tasks, i), ref, "depth %d task %d" % (depth, i))
            n_checked += len(ref)
    assert n_checked > 150


@pytest.mark.gpu
def test_picker_edge_cases(ctx, oracle):
    rng = np.random.default_rng(2)
    s = _gpu_settings(4)
    # no templates, empty / too short / fully masked templates, template with N
    lo, recs = ctx.p3_candidates(s, [])
    assert list(lo) == [0] and len(recs) == 0
    t_ok = rand_template(rng, 120)
    templates = ["", "ACGTACGTAC", t_ok, t_ok, "N" * 80]
    masks = [np.zeros(len(t), np.uint8) for t in templates]
    masks[3][:] = 1
    lo, recs = ctx.p3_candidates(s, templates, masks)
    per_list = np.diff(lo)
    assert per_list[0] == per_list[1] == per_list[2] == per_list[3] == 0       # sets 0 and 1
    assert per_list[4] > 0 and per_list[5] > 0                                   # the plain template
    assert per_list[6] == per_list[7] == per_list[8] == per_list[9] == 0       # masked, all-N
    # tasks: primer outside the template, wrong length, on masked bases, no tasks at all
    tasks = np.array([(2, 1, 110, 20), (2, 1, 10, 5), (3, 1, 10, 20), (2, 2, 5, 20), (2, 3, 10, 20)],
                     dtype=_native.P3_TASK_DTYPE)
    fixed, pairs, n_pairs = ctx.p3_design_fixed(s, templates, tasks, masks)
    assert not (fixed["flags"] & _native.P3_OK).any() and not n_pairs.any()
    f0, p0, n0 = ctx.p3_design_fixed(s, templates, np.zeros(0, dtype=_native.P3_TASK_DTYPE), masks)
    assert len(f0) == 0 and p0.shape == (0, 4)
    with pytest.raises(RuntimeError):
        ctx.p3_design_fixed(s, templates, np.array([(9, 1, 0, 20)], dtype=_native.P3_TASK_DTYPE), masks)
    # a target between the given primer and the picked ones, and an included region (two-sided call)
    p3b.resetP3Globals()
    g = {**KASP_GLOBALS, "PRIMER_NUM_RETURN": 6, "PRIMER_PRODUCT_SIZE_RANGE": [[60, 170]]}
    p3b.setP3Globals(g)
    so = oracle_settings(g, 6)
    t = rand_template(rng, 200)
    full = {**p3_oracle.DEFAULTS, **so}
    for left in [r for r in p3_oracle.candidates(full, t, [0] * len(t), "L")
                 if r["start"] >= 30 and r["start"] + r["len"] <= 88]:
        ref = p3_oracle.design(oracle, so, t, target=(90, 10), left=left["seq"])
        if ref:       # skip given primers whose own self alignments exceed the limits
            break
    res = p3b.designPrimers({"SEQUENCE_TEMPLATE": t, "SEQUENCE_PRIMER": left["seq"], "SEQUENCE_TARGET": [90, 10]})
    assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref) > 0
    for i, e in enumerate(ref):
        assert res["PRIMER_RIGHT_%d_SEQUENCE" % i] == e["right"]["seq"]
        assert e["right"]["start"] - e["right"]["len"] + 1 >= 100
    res = p3b.designPrimers({"SEQUENCE_TEMPLATE": t, "SEQUENCE_INCLUDED_REGION": [20, 150]})
    ref = p3_oracle.design(oracle, so, t, excluded=[(0, 20), (170, 30)])
    assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref) > 0
    assert [res["PRIMER_LEFT_%d_SEQUENCE" % i] for i in range(len(ref))] == [e["left"]["seq"] for e in ref]
    assert [res["PRIMER_RIGHT_%d" % i] for i in range(len(ref))] == [(e["right"]["start"], e["right"]["len"]) for e in ref]
    p3b.resetP3Globals()


@pytest.mark.gpu
def test_two_sided_batch_matches_oracle(ctx, oracle):
    # several templates in one call (the CAPS / HRM call shape: both primers picked around a target),
    # including one whose first size range cannot fill the request and one with very few candidates
    rng = np.random.default_rng(23)
    g = {**KASP_GLOBALS, "PRIMER_NUM_RETURN": 9, "PRIMER_PRODUCT_SIZE_RANGE": [[60, 80], [50, 130]]}
    p3b.resetP3Globals()
    p3b.setP3Globals(g)
    so = oracle_settings(g, 9)
    cases = [(rand_template(rng, 190), (90, 5), [(0, 6)]), (rand_template(rng, 150, gc=0.5), (70, 2), []),
             (rand_template(rng, 110), (50, 4), [(0, 20), (90, 20)]), (rand_template(rng, 160, gc=0.3), None, [])]
    args = []
    for t, target, excl in cases:
        a = {"SEQUENCE_TEMPLATE": t}
        if target:
            a["SEQUENCE_TARGET"] = list(target)
        if excl:
            a["SEQUENCE_EXCLUDED_REGION"] = [list(x) for x in excl]
        args.append(a)
    got = p3b.design_primers_batch(args)
    n = 0
    for (t, target, excl), res in zip(cases, got):
        ref = p3_oracle.design(oracle, so, t, excluded=excl, target=target)
        assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref)
        for i, e in enumerate(ref):
            assert res["PRIMER_LEFT_%d" % i] == (e["left"]["start"], e["left"]["len"])
            assert res["PRIMER_RIGHT_%d" % i] == (e["right"]["start"], e["right"]["len"])
            assert res["PRIMER_PAIR_%d_PENALTY" % i] == e["pair_penalty"]
            assert res["PRIMER_PAIR_%d_COMPL_END_TH" % i] == e["compl_end_th"]
        n += len(ref)
    assert n >= 18
    p3b.resetP3Globals()


@pytest.mark.gpu
def test_long_template_lists_use_the_global_sort_path(ctx):
    # > 4096 candidates per list: the per-list bitonic sort leaves shared memory for global scratch
    rng = np.random.default_rng(77)
    t = rand_template(rng, 5200, gc=0.5)
    s = _gpu_settings(5)
    list_off, recs = ctx.p3_candidates(s, [t, rand_template(rng, 300)])
    assert list_off[1] - list_off[0] > 4096 and list_off[2] - list_off[1] > 4096
    so = {**p3_oracle.DEFAULTS, **oracle_settings(KASP_GLOBALS, 5)}
    seqs = _native.unpack(recs["seq"])
    for strand, side in enumerate("LR"):
        ref = p3_oracle.candidates(so, t, [0] * len(t), side)
        lo, hi = list_off[strand], list_off[strand + 1]
        assert hi - lo == len(ref)
        got = recs[lo:hi]
        assert [int(x) for x in got["start"]] == [r["start"] for r in ref]
        assert [int(x) for x in got["len"]] == [r["len"] for r in ref]
        assert [float(x) for x in got["penalty"]] == [r["penalty"] for r in ref]
        assert seqs[lo:hi] == [r["seq"] for r in ref]


@pytest.mark.gpu
def test_two_sided_per_left_quota_is_exact(ctx, oracle):
    # more pairs wanted than the per-left quota of the first pass (16): the quota bounds must either
    # prove the merge final or send the template to the unrestricted second pass
    rng = np.random.default_rng(101)
    g = {**KASP_GLOBALS, "PRIMER_NUM_RETURN": 40, "PRIMER_PRODUCT_SIZE_RANGE": [[60, 90], [50, 160]]}
    p3b.resetP3Globals()
    p3b.setP3Globals(g)
    so = oracle_settings(g, 40)
    cases = [(rand_template(rng, 300), None), (rand_template(rng, 260, gc=0.5), (125, 8)),
             (rand_template(rng, 140), None)]
    args = [{"SEQUENCE_TEMPLATE": t, **({"SEQUENCE_TARGET": list(tg)} if tg else {})} for t, tg in cases]
    got = p3b.design_primers_batch(args)
    n = 0
    for (t, tg), res in zip(cases, got):
        ref = p3_oracle.design(oracle, so, t, target=tg)
        assert res["PRIMER_PAIR_NUM_RETURNED"] == len(ref)
        for i, e in enumerate(ref):
            assert res["PRIMER_LEFT_%d" % i] == (e["left"]["start"], e["left"]["len"]), i
            assert res["PRIMER_RIGHT_%d" % i] == (e["right"]["start"], e["right"]["len"]), i
            assert res["PRIMER_PAIR_%d_PENALTY" % i] == e["pair_penalty"]
        n += len(ref)
    assert n >= 60
    p3b.resetP3Globals()
