#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ by running the REFERENCE's
own Python (/root/reference/src/polyoligo, imported here in the build
container) on seeded synthetic inputs.

The reference's thermodynamics come from primer3-py, which cannot be installed
offline, so ``primer3.thermoanalysis.ThermoAnalysis`` is backed by the CPU
oracle (tests/golden/ref_stubs.py).  Everything ABOVE that call -- candidate
enumeration, validity filter, heterodimer predicates, masks, scoring, classify
and prune -- is the reference's unmodified code, so these fixtures pin the
host-logic rows of SURVEY.md 8a (a2, a3, a4, a8, a9) to the reference itself.

Run (only where /root/reference exists):  python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.oracle import Oracle  # noqa: E402
import ref_stubs  # noqa: E402

ref_stubs.install(Oracle())
import polyoligo.lib_primer3 as lp  # noqa: E402
import polyoligo._lib_kasp as lk  # noqa: E402
import polyoligo.scoring as sc  # noqa: E402
import polyoligo.lib_utils as lu  # noqa: E402

KASP_GLOBALS = dict(PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0,
                    PRIMER_MIN_TM=55.0, PRIMER_MAX_TM=65.0)   # reference data/PRIMER3_KASP.yaml
VIC = "GAAGGTCGGAGTCAACGGATT"   # reference cli_kasp.py:65
FAM = "GAAGGTGACCAAGTTCATGCT"   # reference cli_kasp.py:72


def rand_seq(rng, n, gc=0.5):
    p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    return "".join(np.array(list("ACGT"))[rng.choice(4, n, p=p)])


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=0, sort_keys=True)
    print("wrote", name)


def make_hroi(seq, start, pos1, ref, alt):
    marker = types.SimpleNamespace(pos1=pos1, ref=ref, alt=alt)
    return types.SimpleNamespace(seq=seq, start=start, stop=start + len(seq) - 1, marker=marker,
                                 chrom="chrT", name="m")


def golden_marker_primers(rng):
    """a3: get_marker_primers (reference _lib_kasp.py:295-359), with and without the validity filter."""
    lp.set_tm_delta(5.0)
    lp.PRIMER3_GLOBALS.clear()
    lp.PRIMER3_GLOBALS.update(KASP_GLOBALS)
    cases = []
    specs = [("snp", 1, 1)] * 14 + [("del", 3, 1), ("del", 5, 2), ("ins", 1, 4), ("ins", 2, 6), ("mnp", 2, 2),
                                    ("mnp", 3, 3)]
    for kind, rl, al in specs:
        seq = rand_seq(rng, 500, gc=rng.choice([0.42, 0.5, 0.58]))
        start = int(rng.integers(1000, 100000))
        idx = 249
        ref = seq[idx:idx + rl]
        alt = ref
        while alt == ref or (al == rl == 1 and alt == ref):
            alt = rand_seq(rng, al)
        hroi = make_hroi(seq, start, start + idx, ref, alt)
        real_valid = lp.Primer.is_valid
        lp.Primer.is_valid = lambda self: True          # every enumerated candidate, in order
        allc = lk.get_marker_primers(hroi)
        lp.Primer.is_valid = real_valid
        kept = lk.get_marker_primers(hroi)
        cases.append({
            "seq": seq, "start": start, "stop": hroi.stop, "pos1": hroi.marker.pos1, "ref": ref, "alt": alt,
            "all": [{"dir": pp.dir, "ref": pp.primers[pp.dir].sequence, "alt": pp.primers["A"].sequence}
                    for pp in allc],
            "kept": [{"dir": pp.dir, "ref": pp.primers[pp.dir].sequence, "alt": pp.primers["A"].sequence,
                      "tm_ref": pp.primers[pp.dir].tm, "tm_alt": pp.primers["A"].tm} for pp in kept]})
    dump("kasp_marker_primers.json", {"globals": KASP_GLOBALS, "tm_delta": 5.0, "cases": cases})


def golden_marker_primers_edges(rng):
    """a3 again, away from the centre of the template: markers next to either template end (Python's
    negative-index slices come into play), longer indels, templates of other lengths."""
    lp.set_tm_delta(5.0)
    lp.PRIMER3_GLOBALS.clear()
    lp.PRIMER3_GLOBALS.update(KASP_GLOBALS)
    cases = []
    specs = []
    for idx in (0, 3, 17, 18, 24, 25, 26, 40):
        specs += [(500, idx, 1, 1), (500, 499 - idx, 1, 1)]
    specs += [(500, 20, 4, 1), (500, 22, 1, 5), (500, 478, 6, 2), (500, 470, 2, 9), (500, 249, 10, 1), (500, 249, 1, 12),
              (120, 60, 1, 1), (120, 30, 3, 1), (61, 30, 1, 1), (300, 290, 2, 2), (512, 256, 1, 1), (512, 500, 1, 3)]
    for n, idx, rl, al in specs:
        seq = rand_seq(rng, n, gc=rng.choice([0.42, 0.5, 0.58]))
        start = int(rng.integers(1000, 100000))
        rl = min(rl, n - idx)
        ref = seq[idx:idx + rl]
        alt = ref
        while alt == ref:
            alt = rand_seq(rng, al)
        hroi = make_hroi(seq, start, start + idx, ref, alt)
        real_valid = lp.Primer.is_valid
        lp.Primer.is_valid = lambda self: True
        try:
            allc = lk.get_marker_primers(hroi)
        finally:
            lp.Primer.is_valid = real_valid
        kept = lk.get_marker_primers(hroi)
        cases.append({
            "seq": seq, "start": start, "stop": hroi.stop, "pos1": hroi.marker.pos1, "ref": ref, "alt": alt,
            "all": [{"dir": pp.dir, "ref": pp.primers[pp.dir].sequence, "alt": pp.primers["A"].sequence}
                    for pp in allc],
            "kept": [{"dir": pp.dir, "ref": pp.primers[pp.dir].sequence, "alt": pp.primers["A"].sequence,
                      "tm_ref": pp.primers[pp.dir].tm, "tm_alt": pp.primers["A"].tm} for pp in kept]})
    dump("kasp_marker_primers_edges.json", {"globals": KASP_GLOBALS, "tm_delta": 5.0, "cases": cases})


def golden_is_valid(rng):
    """a1/a2: Primer.calc_thermals / is_valid (reference lib_primer3.py:56-97)."""
    lp.PRIMER3_GLOBALS.clear()
    lp.PRIMER3_GLOBALS.update(KASP_GLOBALS)
    rows = []
    for tm_delta in (5.0, 10.0):
        lp.set_tm_delta(tm_delta)
        for _ in range(150):
            s = rand_seq(rng, int(rng.integers(16, 28)), gc=rng.choice([0.3, 0.5, 0.7]))
            if rng.random() < 0.05:
                k = int(rng.integers(0, len(s)))
                s = s[:k] + "N" + s[k + 1:]
            p = lp.Primer(sequence=s)
            ok = p.is_valid()
            rows.append({"seq": s, "tm_delta": tm_delta, "valid": bool(ok), "tm": p.tm, "hairpin_tm": p.hairpin_tm,
                         "homodimer_tm": p.homodimer_tm, "gc": p.gc_content, "nN": p.nN, "length": p.length})
    lp.set_tm_delta(5.0)
    dump("is_valid.json", {"globals": KASP_GLOBALS, "rows": rows})


def golden_heterodimer(rng):
    """a8: PrimerPair.check_heterodimerization, PCR (lib_primer3.py:147-164) and KASP (_lib_kasp.py:89-117)."""
    lp.PRIMER3_GLOBALS.clear()
    lp.PRIMER3_GLOBALS.update(KASP_GLOBALS)
    pcr, kasp = [], []
    for tm_delta in (5.0, 2.0):
        lp.set_tm_delta(tm_delta)
        for k in range(60):
            f = rand_seq(rng, int(rng.integers(18, 26)))
            r = rand_seq(rng, int(rng.integers(18, 26)))
            if k % 7 == 0:   # force a strong dimer
                r = lp.Seq(f[-18:]).reverse_complement() + r[18:]
            pp = lp.PrimerPair()
            pp.tm_delta = tm_delta
            pp.primers["F"].sequence = f
            pp.primers["R"].sequence = str(r)
            for d in "FR":
                pp.primers[d].tm_delta = tm_delta
            pp.check_heterodimerization()
            pcr.append({"F": f, "R": str(r), "tm_delta": tm_delta, "dimer": bool(pp.dimer),
                        "tm_F": pp.primers["F"].tm, "tm_R": pp.primers["R"].tm})
        for k in range(60):
            direction = "F" if k % 2 == 0 else "R"
            core = rand_seq(rng, int(rng.integers(18, 26)))
            alt = core[:-1] + rng.choice([c for c in "ACGT" if c != core[-1]])
            com = rand_seq(rng, int(rng.integers(18, 26)))
            if k % 5 == 0:
                com = str(lp.Seq((VIC + core)[-20:]).reverse_complement()) + com[20:]
            pp = lk.PrimerPair()
            pp.tm_delta = tm_delta
            pp.dir = direction
            pp.primers[direction].sequence = core
            pp.primers["A"].sequence = alt
            pp.primers["R" if direction == "F" else "F"].sequence = com
            pp.add_reporter_dyes([VIC, FAM])
            pp.check_heterodimerization()
            kasp.append({"dir": direction, "F": pp.primers["F"].sequence, "R": pp.primers["R"].sequence,
                         "A": pp.primers["A"].sequence, "tm_delta": tm_delta, "reporters": [VIC, FAM],
                         "ref_dimer": bool(pp.ref_dimer), "alt_dimer": bool(pp.alt_dimer)})
    lp.set_tm_delta(5.0)
    dump("heterodimer_checks.json", {"pcr": pcr, "kasp": kasp})


def golden_scoring(rng):
    """a9: scoring.score_* (scoring.py:4-151), PCR.classify / prune (lib_primer3.py:425-463, _lib_kasp.py:278-291)."""
    rows = []
    pools = {"KASP": [], "PCR": [], "HRM": []}
    for k in range(240):
        assay = ["KASP", "PCR", "HRM", "CAPS"][k % 4]
        keys = ["F", "R", "A"] if assay == "KASP" else ["F", "R"]
        pp = types.SimpleNamespace()
        base = rng.uniform(55, 65)
        spread = rng.choice([0.5, 1.5, 3.0, 6.0])
        pp.primers = {d: types.SimpleNamespace(tm=float(base + rng.uniform(-spread, spread)),
                                               max_aaf=float(rng.choice([0.0, 0.0, 0.05, 0.099, 0.1, 0.3])))
                      for d in keys}
        pp.offtargets = ["x"] * int(rng.choice([0, 0, 1, 3]))
        pp.dimer = bool(rng.random() < 0.3)
        pp.ref_dimer = bool(rng.random() < 0.2)
        pp.alt_dimer = bool(rng.random() < 0.2)
        pp.max_indel_size = int(rng.choice([0, 0, 1, 49, 50, 120]))
        pp.hrm = {"delta": float(rng.choice([0.0, 0.3, 0.5, 0.7, 1.0, 1.8]))}
        goodness, qcode = sc.get_score_fnc(assay)(pp)
        row = {"assay": assay, "tms": [pp.primers[d].tm for d in keys],
               "max_aafs": [pp.primers[d].max_aaf for d in keys], "n_offtargets": len(pp.offtargets),
               "dimer": pp.dimer, "ref_dimer": pp.ref_dimer, "alt_dimer": pp.alt_dimer,
               "max_indel_size": pp.max_indel_size, "hrm_delta": pp.hrm["delta"],
               "goodness": int(goodness), "qcode": qcode}
        rows.append(row)
        if assay in pools:
            pp.goodness, pp.qcode, pp.uid = goodness, qcode, len(pools[assay])
            pp.score = lambda assay_name, pp=pp: None       # already scored
            pools[assay].append(pp)
    prune = {}
    for assay, pps in pools.items():
        for n in (3, 10, 1000):
            pcr = lk.PCR() if assay == "KASP" else lp.PCR()
            pcr.pps = pps
            pcr.classify(assay)
            if assay == "KASP":
                pcr.prune(n)
            else:
                pcr.prune(n, assay_name=assay)
            order = []
            for score in sorted(pcr.pps_pruned.keys(), reverse=True):   # report order (_lib_kasp.py:369)
                order += [pp.uid for pp in pcr.pps_pruned[score]]
            prune["%s/%d" % (assay, n)] = {
                "goodness": [int(pp.goodness) for pp in pps], "hrm_delta": [pp.hrm["delta"] for pp in pps],
                "order": order}
    dump("scoring.json", {"rows": rows, "prune": prune})


def golden_masks(rng):
    """a4: include_mut_in_included_maps, the KASP marker-window step, excluded regions,
    list_2_ranges, reduce_ivs (lib_primer3.py:564-600, _lib_kasp.py:681-684, lib_utils.py:144-261)."""
    cases = []
    for k in range(12):
        n = 500
        start = [int(rng.integers(1000, 10 ** 6)), 20, 3, 26][k % 4] if k >= 8 else int(rng.integers(1000, 10 ** 6))
        mism = rng.random(n) < 0.05
        partial = mism | (rng.random(n) < 0.15)
        allm = np.ones(n, dtype=bool)
        if k % 3 == 0:
            partial[:] = False
            partial[100:300] = True
        muts = [types.SimpleNamespace(pos=start + int(p)) for p in np.flatnonzero(rng.random(n) < 0.01)]
        hroi = types.SimpleNamespace(start=start, vcf_hook=True, mutations=muts,
                                     p3_sequence_included_maps={"all": allm.copy(), "partial_mism": partial.copy(),
                                                                "mism": mism.copy()})
        lp.include_mut_in_included_maps(hroi)
        before = {k2: [int(x) for x in np.flatnonzero(~v)] for k2, v in hroi.p3_sequence_included_maps.items()}
        pcr_regions = {}
        lp.get_sequence_excluded_regions(hroi)        # PCR / CAPS / HRM use the maps as they are
        for k2, v in hroi.p3_sequence_excluded_regions.items():
            pcr_regions[k2] = [[int(a), int(b)] for a, b in v]
        # KASP: marker window (reference _lib_kasp.py:681-684, restated literally incl. its quirk)
        for mmap in hroi.p3_sequence_included_maps.values():
            s = np.min([1, (hroi.start - 25)])
            e = np.max([len(mmap), (hroi.start + 25)])
            mmap[s:e] = True
        lp.get_sequence_excluded_regions(hroi)
        kasp_regions = {k2: [[int(a), int(b)] for a, b in v] for k2, v in hroi.p3_sequence_excluded_regions.items()}
        cases.append({"start": start, "n": n, "mism": [int(x) for x in np.flatnonzero(~mism)],
                      "partial_mism": [int(x) for x in np.flatnonzero(~partial)],
                      "mutation_pos": [int(m.pos) for m in muts], "excluded_idx_after_mut": before,
                      "pcr_excluded_regions": pcr_regions, "kasp_excluded_regions": kasp_regions})
    ivs = []
    for _ in range(10):
        m = int(rng.integers(150, 400))
        starts = np.sort(rng.choice(5000, m, replace=False))
        iv = [[int(s), int(rng.integers(1, 6))] for s in starts]
        red = lu.reduce_ivs([list(x) for x in iv], n=200)
        ivs.append({"in": iv, "out": [[int(a), int(b)] for a, b in red]})
    dump("masks.json", {"cases": cases, "reduce_ivs": ivs})


def golden_mutations(rng):
    """8f-3: PrimerPair.add_mutations (reference lib_primer3.py:166-210), PCR and KASP pairs."""
    import polyoligo.lib_vcf as lv
    cases = []
    for k in range(40):
        kasp = k % 2 == 1
        start = int(rng.integers(1000, 50000))
        muts = []
        for pos in np.sort(rng.choice(np.arange(start, start + 400), int(rng.integers(0, 12)), replace=False)):
            ref = rand_seq(rng, int(rng.choice([1, 1, 1, 2, 4])))
            alt = [rand_seq(rng, int(rng.choice([1, 1, 2, 60])))]
            muts.append(lv.Mutation("chrT", int(pos), ref, alt, aaf=float(rng.choice([0.0, 0.02, 0.099, 0.1, 0.4]))))
        pp = lk.PrimerPair() if kasp else lp.PrimerPair()
        fs = start + int(rng.integers(0, 100))
        fl, rl = int(rng.integers(18, 26)), int(rng.integers(18, 26))
        rs = fs + int(rng.integers(60, 250))
        pp.primers["F"].sequence = rand_seq(rng, fl)
        pp.primers["F"].start, pp.primers["F"].stop = fs, fs + fl - 1
        pp.primers["R"].sequence = rand_seq(rng, rl)
        pp.primers["R"].start, pp.primers["R"].stop = rs, rs + rl - 1
        if kasp:
            pp.dir = "F" if k % 4 == 1 else "R"
            pp.primers["A"].sequence = pp.primers[pp.dir].sequence[:-1] + "A"
            pp.primers["A"].start, pp.primers["A"].stop = pp.primers[pp.dir].start, pp.primers[pp.dir].stop
        pp.add_mutations(muts)
        keys = ["F", "R", "A"] if kasp else ["F", "R"]
        cases.append({"kasp": kasp, "mutations": [[m.pos, m.aaf, int(m.indel_size)] for m in muts],
                      "start": [int(pp.primers[d].start) for d in keys],
                      "stop": [int(pp.primers[d].stop) for d in keys],
                      "max_aaf": [float(pp.primers[d].max_aaf) for d in keys],
                      "max_indel": int(pp.max_indel_size)})
    dump("mutations.json", {"cases": cases})


def golden_offtargets(rng):
    """8f-4: PCR.list_offtargets (reference lib_primer3.py:392-419) on random BLAST-like hits."""
    lp.set_offtarget_size(40, 900)
    cases = []
    for _ in range(12):
        hits = {"F": {}, "R": {}}
        for d in "FR":
            # sorted: dump() writes dict keys sorted, and the iteration order decides the output order
            for chrom in sorted(rng.choice(["chr1", "chr2", "chr3", "scaffold_9"], size=int(rng.integers(1, 5)),
                                           replace=False)):
                n = int(rng.integers(0, 9))
                hits[d][str(chrom)] = [{"sstart": int(x)} for x in rng.integers(1000, 3000, n)]
        cases.append({"hits": hits, "offtargets": lp.PCR.list_offtargets(hits)})
    dump("offtargets.json", {"min_size": 40, "max_size": 900, "cases": cases})


def _p3_dict(pairs):
    """primer3-py style result dict (key order as primer3 emits it) from oracle picker pairs."""
    d = {"PRIMER_LEFT_EXPLAIN": "considered 0", "PRIMER_RIGHT_EXPLAIN": "considered 0",
         "PRIMER_PAIR_EXPLAIN": "considered 0", "PRIMER_LEFT_NUM_RETURNED": len(pairs),
         "PRIMER_RIGHT_NUM_RETURNED": len(pairs), "PRIMER_INTERNAL_NUM_RETURNED": 0,
         "PRIMER_PAIR_NUM_RETURNED": len(pairs)}
    for i, e in enumerate(pairs):
        d["PRIMER_PAIR_%d_PENALTY" % i] = e["pair_penalty"]
        for side, o in (("LEFT", e["left"]), ("RIGHT", e["right"])):
            d["PRIMER_%s_%d_PENALTY" % (side, i)] = o["penalty"]
        for side, o in (("LEFT", e["left"]), ("RIGHT", e["right"])):
            d["PRIMER_%s_%d_SEQUENCE" % (side, i)] = o["seq"]
        for side, o in (("LEFT", e["left"]), ("RIGHT", e["right"])):
            d["PRIMER_%s_%d" % (side, i)] = [o["start"], o["len"]]
        for f, k in (("TM", "tm"), ("GC_PERCENT", "gc_percent"), ("SELF_ANY_TH", "self_any_th"),
                     ("SELF_END_TH", "self_end_th"), ("HAIRPIN_TH", "hairpin_th")):
            for side, o in (("LEFT", e["left"]), ("RIGHT", e["right"])):
                d["PRIMER_%s_%d_%s" % (side, i, f)] = o[k]
        for side, o in (("LEFT", e["left"]), ("RIGHT", e["right"])):
            d["PRIMER_%s_%d_END_STABILITY" % (side, i)] = 3.21 + 0.01 * i
        d["PRIMER_PAIR_%d_COMPL_ANY_TH" % i] = e["compl_any_th"]
        d["PRIMER_PAIR_%d_COMPL_END_TH" % i] = e["compl_end_th"]
        d["PRIMER_PAIR_%d_PRODUCT_SIZE" % i] = e["product_size"]
    return d


def _pp_row(pp):
    row = {"dir": pp.dir, "penalty": pp.penalty, "product_size": pp.product_size,
           "compl_any_th": pp.compl_any_th, "compl_end_th": pp.compl_end_th}
    for d, p in pp.primers.items():
        row[d] = {"sequence": p.sequence, "start": p.start, "stop": p.stop, "id": p.id, "penalty": p.penalty,
                  "end_stability": p.end_stability, "hairpin_th": p.hairpin_th, "self_any_th": p.self_any_th,
                  "self_end_th": p.self_end_th}
    return row


def golden_design(rng):
    """a5 (result parsing), a6 and a7: the reference's get_primers / KASP design_primers + merge_primers
    (_lib_kasp.py:552-631) / PCR design_primers (_lib_pcr.py:167-219) run on CANNED designPrimers
    answers (made by the picker oracle; primer3 itself is not installable), so the selection logic
    around the picker is pinned to the reference's own code."""
    from oracle import p3_oracle
    import polyoligo._lib_pcr as lpcr
    oracle = Oracle()
    lp.set_tm_delta(5.0)
    lp.PRIMER3_GLOBALS.clear()
    lp.PRIMER3_GLOBALS.update(KASP_GLOBALS)
    so = dict(min_size=18, opt_size=20, max_size=25, num_return=6, min_tm=55.0, opt_tm=60.0, max_tm=65.0,
              min_gc=20.0, max_gc=80.0, max_poly_x=100, pair_max_diff_tm=6.0,
              size_ranges=((90, 150), (100, 200), (50, 250)),
              # same salt as ThermoAnalysis() so that many picked primers also pass Primer.is_valid
              # and the merge logic is exercised; what the picker answers is canned either way
              dv_conc=0.0, dntp_conc=0.0)
    canned = []

    def fake_design(seq_args):
        excluded = [tuple(x) for x in seq_args.get("SEQUENCE_EXCLUDED_REGION", [])]
        tgt = seq_args.get("SEQUENCE_TARGET")
        res = p3_oracle.design(oracle, so, seq_args["SEQUENCE_TEMPLATE"], excluded=excluded,
                               target=tuple(tgt) if tgt else None, left=seq_args.get("SEQUENCE_PRIMER"),
                               right=seq_args.get("SEQUENCE_PRIMER_REVCOMP"))
        d = _p3_dict(res)
        canned.append(list(d.items()))
        return d

    sys.modules["primer3"].bindings.designPrimers = fake_design
    kasp_cases = []
    for _ in range(2):
        seq = rand_seq(rng, 400, gc=0.48)
        start = int(rng.integers(1000, 100000))
        idx = 199
        ref = seq[idx]
        alt = rand_seq(rng, 1)
        while alt == ref:
            alt = rand_seq(rng, 1)
        hroi = make_hroi(seq, start, start + idx, ref, alt)
        mpps = lk.get_marker_primers(hroi)
        del canned[:]
        excluded = [[0, 12], [300, 25]]
        repo = lk.design_primers(pps_repo=[], mpps=mpps, roi=hroi, sequence_excluded_region=excluded, n_primers=10)
        first = [_pp_row(pp) for pp in repo]
        # second search pass into the same repository (as _lib_kasp.main does per search type)
        n1 = len(canned)
        repo = lk.design_primers(pps_repo=repo, mpps=mpps, roi=hroi, sequence_excluded_region=[[0, 5]],
                                 n_primers=14)
        second = [_pp_row(pp) for pp in repo]
        # the rest of _lib_kasp.main (:739-748) on that repository: dyes, heterodimers, mutations,
        # classify, prune, report
        import tempfile
        covered = sorted({int(p) for pp in repo for d in "FR" for p in range(pp.primers[d].start, pp.primers[d].stop + 1)})
        muts = []
        for k, pos in enumerate(rng.choice(covered, size=7, replace=False)):
            pos = int(pos)
            r0 = seq[pos - start]
            alts = [b for b in "ACGT" if b != r0]
            if k == 2:
                m = types.SimpleNamespace(pos=pos, ref=r0, alt=alts[:2], aaf=0.31, indel_size=0)
            elif k == 4:
                m = types.SimpleNamespace(pos=pos, ref=seq[pos - start:pos - start + 2], alt=[alts[0]], aaf=0.07,
                                          indel_size=1)
            else:
                m = types.SimpleNamespace(pos=pos, ref=r0, alt=[alts[k % 3]], aaf=float(rng.random() * 0.5),
                                          indel_size=0)
            muts.append(m)
        pcr = lk.PCR(snp_id="mk%d" % len(kasp_cases), chrom="chrT", pos=hroi.marker.pos1, ref=ref, alt=alt)
        pcr.pps = repo
        pcr.add_reporter_dyes([VIC, FAM])
        pcr.check_heterodimerization()
        pcr.add_mutations(muts)
        pcr.classify(assay_name="KASP")
        n_kept = pcr.prune(6)
        pcr.is_indel = None
        with tempfile.TemporaryDirectory() as tmp:
            lk.print_report(pcr, os.path.join(tmp, "rep"))
            lk.print_report_header(os.path.join(tmp, "hdr.txt"))
            report = {k: open(os.path.join(tmp, k)).read() for k in ("rep.txt", "rep.bed", "hdr.txt")}
        kasp_cases.append({"seq": seq, "start": start, "stop": hroi.stop, "pos1": hroi.marker.pos1, "ref": ref,
                           "alt": alt, "excluded": [excluded, [[0, 5]]], "n_primers": [10, 14],
                           "canned": [canned[:n1], canned[n1:]], "repo": [first, second],
                           "mutations": [{"pos": m.pos, "ref": m.ref, "alt": m.alt, "aaf": m.aaf,
                                          "indel_size": m.indel_size} for m in muts],
                           "n_kept": int(n_kept), "report": report})
    # PCR: both primers picked, diversity limit max_unique = 2
    pcr_cases = []
    for _ in range(2):
        seq = rand_seq(rng, 220, gc=0.5)
        roi = types.SimpleNamespace(seq=seq, chrom="chrP", left_roi=types.SimpleNamespace(start=5000))
        del canned[:]
        so["num_return"] = 30
        so["size_ranges"] = ((80, 160),)
        repo = lpcr.design_primers(pps_repo=[], roi=roi, sequence_target=[100, 6],
                                   sequence_excluded_region=[[0, 4]], n_primers=8, max_unique=2)
        pcr_cases.append({"seq": seq, "left_start": 5000, "target": [100, 6], "excluded": [[0, 4]], "n_primers": 8,
                          "max_unique": 2, "canned": canned[0], "repo": [_pp_row(pp) for pp in repo]})
    sys.modules["primer3"].bindings.designPrimers = lambda *a, **k: {}
    dump("design.json", {"globals": KASP_GLOBALS, "tm_delta": 5.0, "kasp": kasp_cases, "pcr": pcr_cases})


def golden_pcr_main(rng):
    """_lib_pcr.main (reference :222-356) and _lib_hrm.main (:190-341) from the point where the two
    homolog-mapped flank ROIs exist: include_mut_in_included_maps, Region.merge_with_primers
    (lib_markers.py:403-449), get_sequence_excluded_regions, the default product range (:323-324), one
    design_primers per search mask (picker answers CANNED from the picker oracle), heterodimers,
    add_mutations, (HRM: PCRproduct.get_hrm_temps,) classify, prune and the report writer -- all the
    reference's own code on a synthetic genome behind a stand-in BLAST hook (BLAST itself is external)."""
    import tempfile
    from oracle import p3_oracle
    import polyoligo._lib_pcr as lpcr
    import polyoligo._lib_hrm as lhrm
    import polyoligo.lib_markers as lm
    import polyoligo.lib_vcf as lv
    oracle = Oracle()
    pcr_globals = dict(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0,
                       PRIMER_MIN_TM=55.0, PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0)

    class FakeBlast:                      # lib_blast.BlastDB.fetch: {name: sequence} of 1-based closed ranges
        def __init__(self, genome):
            self.genome = genome

        def fetch(self, query):
            q = query[0]
            return {"%s:%d-%d" % (q["chr"], q["start"], q["stop"]): self.genome[q["start"] - 1:q["stop"]]}

    def hroi_at(blast, pos, outer, maps_p, muts, marker):
        lp_, rp_ = lu.get_n_padding(1, outer)
        h = lm.ROI(chrom="chrG", start=pos - lp_, stop=pos + rp_, blast_hook=blast, vcf_hook=True, name="t", marker=marker)
        n = h.n
        mism = rng.random(n) < maps_p[0]
        partial = mism | (rng.random(n) < maps_p[1])
        h.p3_sequence_included_maps = {"all": np.repeat(True, n), "partial_mism": partial, "mism": mism}
        h.mutations = [m for m in muts if h.start <= m.pos <= h.stop]
        return h

    cases = []
    specs = [("PCR", 300, 250, False), ("PCR", 90, 250, True), ("PCR", 1200, 250, False), ("HRM", 1, 1000, False),
             ("HRM", 1, 1000, True), ("CAPS", 1, 400, False), ("CAPS", 1, 400, True)]
    import polyoligo._lib_caps as lcaps
    wanted = ["EcoRI", "BamHI", "HindIII", "HpaII", "TaqI", "AluI", "AarI", "XbaI", "PstI", "KpnI", "SmaI", "NotI"]
    all_enzymes = lcaps.get_enzymes(included_enzymes=None)
    fragment_min_size = 100            # cli_caps.py --fragment_min_size (flanks and this kept small: the picker oracle is eager)
    for assay, tlen, outer, narrow in specs:
        genome = rand_seq(rng, 4000, gc=0.45)
        ts = int(rng.integers(1200, 1600))
        if assay == "CAPS":              # the REF allele completes the only EcoRI site of the neighbourhood
            genome = genome.replace("GAATTC", "GATATC")
            genome = genome[:ts - 3] + "GAATTC" + genome[ts + 3:]
        blast = FakeBlast(genome)
        marker = None
        if assay in ("HRM", "CAPS"):
            ref = genome[ts - 1]
            alt = rng.choice([c for c in "ACGT" if c != ref])
            marker = types.SimpleNamespace(name="mk", chrom="chrG", pos1=ts, ref=ref, alt=str(alt), n=1)
        muts = []
        for pos in np.sort(rng.choice(np.arange(ts - 600, ts + tlen + 600), 70, replace=False)):
            pos = int(pos)
            r0 = genome[pos - 1]
            kind = len(muts) % 9
            aaf = float(rng.choice([0.02, 0.099, 0.1, 0.3]))
            if kind == 4:                # deletion of 1-3 bases (VCF style: the anchor base stays)
                k = int(rng.integers(1, 4))
                muts.append(lv.Mutation("chrG", pos, genome[pos - 1:pos + k], [r0], aaf=aaf))
            elif kind == 7:              # insertion, one of them long enough for the 'I' quality flag
                k = 60 if len(muts) == 7 else int(rng.integers(1, 5))
                muts.append(lv.Mutation("chrG", pos, r0, [r0 + rand_seq(rng, k)], aaf=aaf))
            else:
                alts = [c for c in "ACGT" if c != r0]
                pick = [alts[int(rng.integers(0, 3))]] if kind != 2 else alts[:2]
                muts.append(lv.Mutation("chrG", pos, r0, pick, aaf=aaf))
        region = lm.Region(chrom="chrG", start=ts, stop=ts + tlen - 1, blast_hook=blast, vcf_hook=True, name="reg%d" % len(cases),
                           marker=marker)
        if assay in ("HRM", "CAPS"):     # _lib_hrm.main :231-252, _lib_caps.main :253-274: flanks at +- OUTER_LEN / 2
            lpos, rpos = int(region.start - outer / 2), int(region.start + outer / 2)
            flank = outer
        else:                            # _lib_pcr.main :257-278
            lpos, rpos, flank = region.start, region.stop, outer
        # narrow: a homolog exists, so "mism" / "partial_mism" keep only the few distinguishing bases
        # (lib_markers.py:236-243); otherwise no homolog was found and every map is all-True (:232-235)
        p = (1.0, 1.0) if not narrow else (0.10, 0.15)
        region.left_roi = hroi_at(blast, lpos, flank, p, muts, marker)
        region.right_roi = hroi_at(blast, rpos, flank, p, muts, marker)
        inputs = {side: {"start": h.start, "stop": h.stop,
                         "maps": {k: [int(x) for x in np.flatnonzero(~v)] for k, v in h.p3_sequence_included_maps.items()},
                         "mutations": [[m.pos, m.ref, m.alt, m.aaf, int(m.indel_size)] for m in h.mutations]}
                  for side, h in (("left", region.left_roi), ("right", region.right_roi))}
        lp.include_mut_in_included_maps(region.left_roi)
        lp.include_mut_in_included_maps(region.right_roi)
        if assay == "CAPS":              # _lib_caps.main :298-311: room for the fragments around the marker
            region = lm.Region(chrom=region.chrom, start=region.marker.pos1 - fragment_min_size,
                               stop=region.marker.pos1 + fragment_min_size, blast_hook=blast, vcf_hook=True,
                               name=region.name, marker=marker, left_roi=region.left_roi, right_roi=region.right_roi)
        target = {"start": region.start, "stop": region.stop}
        region = region.merge_with_primers()
        lp.get_sequence_excluded_regions(region)
        merged = {"start": region.start, "stop": region.stop, "n": region.n, "seq": region.seq,
                  "target": region.p3_sequence_target,
                  "excluded_idx": {k: [int(x) for x in np.flatnonzero(~v)] for k, v in region.p3_sequence_included_maps.items()},
                  "excluded_regions": {k: [[int(a), int(b)] for a, b in v] for k, v in region.p3_sequence_excluded_regions.items()}}
        # ---- the design loop, picker answers canned ----
        lp.PRIMER3_GLOBALS.clear()
        lp.PRIMER3_GLOBALS.update(pcr_globals)
        if assay == "HRM":
            lp.PRIMER3_GLOBALS["PRIMER_PRODUCT_SIZE_RANGE"] = [[40, 75], [40, 150]]
        if assay == "CAPS":              # _lib_caps.main :343 sets it unconditionally
            lp.PRIMER3_GLOBALS["PRIMER_PRODUCT_SIZE_RANGE"] = [[region.p3_sequence_target[0][1], region.n]]
        depth = 12
        lp.PRIMER3_GLOBALS["PRIMER_NUM_RETURN"] = depth
        lp.set_tm_delta(5.0)
        if "PRIMER_PRODUCT_SIZE_RANGE" not in lp.PRIMER3_GLOBALS:
            lp.PRIMER3_GLOBALS["PRIMER_PRODUCT_SIZE_RANGE"] = [[region.p3_sequence_target[0][1], region.n]]
        size_ranges = tuple(tuple(x) for x in lp.PRIMER3_GLOBALS["PRIMER_PRODUCT_SIZE_RANGE"])
        so = dict(min_size=18, opt_size=20, max_size=25, num_return=depth, min_tm=55.0, opt_tm=60.0, max_tm=65.0,
                  min_gc=20.0, max_gc=80.0, size_ranges=size_ranges,
                  # ThermoAnalysis()'s salt, so that picked primers also pass Primer.is_valid and the
                  # max_unique / search-order logic has something to select from (answers are canned)
                  dv_conc=0.0, dntp_conc=0.0)
        canned = []

        def fake_design(seq_args, so=so, canned=canned):
            excluded = [tuple(x) for x in seq_args.get("SEQUENCE_EXCLUDED_REGION", [])]
            tgt = seq_args.get("SEQUENCE_TARGET")
            res = p3_oracle.design(oracle, so, seq_args["SEQUENCE_TEMPLATE"], excluded=excluded,
                                   target=tuple(tgt[0]) if tgt else None)
            d = _p3_dict(res)
            canned.append({"excluded": [list(x) for x in excluded], "answer": list(d.items())})
            return d

        sys.modules["primer3"].bindings.designPrimers = fake_design
        if assay in ("HRM", "CAPS"):
            pcr = lhrm.PCR(snp_id=marker.name, chrom=marker.chrom, pos=marker.pos1, ref=marker.ref, alt=marker.alt)
        else:
            pcr = lp.PCR(chrom=region.chrom, name=region.name)
        for st in ["mism_mut", "mism", "partial_mism_mut", "partial_mism", "all_mut", "all"]:
            if len(pcr.pps) < depth:
                pcr.pps = lpcr.design_primers(pps_repo=pcr.pps, roi=region, sequence_target=region.p3_sequence_target,
                                              sequence_excluded_region=region.p3_sequence_excluded_regions[st],
                                              n_primers=depth, max_unique=2)
        pcr.check_heterodimerization()
        pcr.add_mutations(region.mutations)
        hrm_rows = None
        caps_rows = None
        enzymes = None
        repo_all = [_pp_row(pp) for pp in pcr.pps]
        if assay == "CAPS":              # _lib_caps.main :369-380
            enzymes = [dict(e) for e in all_enzymes if e["name"] in wanted]
            fixture_enzymes = [dict(e) for e in enzymes]
            pruned = []
            for pp in pcr.pps:
                prod = lm.PCRproduct(roi=region, pp=pp)
                prod.will_it_cut(enzymes=enzymes, min_fragment_size=fragment_min_size)
                if len(prod.enzymes) > 0:
                    pp.enzymes = prod.enzymes
                    pp.seq_x = prod.roi.seq_x
                    pruned.append(pp)
            pcr.pps = pruned
            caps_rows = [[e["name"] for e in pp.enzymes] for pp in pcr.pps]
            pcr.classify(assay_name="CAPS")
            n_kept = pcr.prune(5)
        elif assay == "HRM":
            for pp in pcr.pps:
                prod = lm.PCRproduct(roi=region, pp=pp)
                prod.get_hrm_temps()
                pp.hrm = prod.hrm
                pp.seq_x = prod.roi.seq_x
            hrm_rows = [dict(pp.hrm) for pp in pcr.pps]
            pcr.classify(assay_name="HRM")
            # PCR.prune orders a goodness class with np.argsort(-delta) "while keeping ties ordered"
            # (lib_primer3.py:444 and its comment); numpy's default sort is not stable (SIMD quicksort,
            # the tie order depends on the numpy build and the CPU), so the fixture is written with the
            # stable sort the comment asks for
            plain_argsort = np.argsort
            np.argsort = lambda a, *args, **kw: plain_argsort(a, *args, **{**kw, "kind": "stable"})
            try:
                n_kept = pcr.prune(5, assay_name="HRM")
            finally:
                np.argsort = plain_argsort
        else:
            pcr.classify(assay_name="PCR")
            n_kept = pcr.prune(5)
        repo = [_pp_row(pp) for pp in pcr.pps]
        with tempfile.TemporaryDirectory() as tmp:
            mod = {"HRM": lhrm, "CAPS": lcaps, "PCR": lpcr}[assay]
            mod.print_report(pcr, os.path.join(tmp, "rep"))
            mod.print_report_header(os.path.join(tmp, "hdr.txt"))
            report = {k: open(os.path.join(tmp, k)).read() for k in ("rep.txt", "rep.bed", "hdr.txt")}
        cases.append({"assay": assay, "genome": genome, "target": target, "outer": outer, "inputs": inputs,
                      "marker": None if marker is None else {"name": marker.name, "chrom": marker.chrom,
                                                              "pos1": marker.pos1, "ref": marker.ref, "alt": marker.alt},
                      "merged": merged, "size_ranges": [list(x) for x in size_ranges], "depth": depth, "canned": canned,
                      "repo": repo, "repo_all": repo_all, "hrm": hrm_rows, "caps": caps_rows,
                      "enzymes": fixture_enzymes if assay == "CAPS" else None,
                      "fragment_min_size": fragment_min_size if assay == "CAPS" else None, "goodness": [int(pp.goodness) for pp in pcr.pps],
                      "n_kept": int(n_kept), "report": report,
                      "mutations": [[m.pos, m.ref, m.alt, m.aaf, int(m.indel_size)] for m in region.mutations]})
    sys.modules["primer3"].bindings.designPrimers = lambda *a, **k: {}
    # the report of a region / marker without any pair (the writers' NA line)
    empty = {}
    with tempfile.TemporaryDirectory() as tmp:
        for assay, mod in (("PCR", lpcr), ("HRM", lhrm), ("CAPS", lcaps)):
            if assay == "PCR":
                pcr = lp.PCR(chrom="chrG", name="nothing")
            else:
                pcr = mod.PCR(snp_id="mk0", chrom="chrG", pos=1234, ref="A", alt="")
            mod.print_report(pcr, os.path.join(tmp, assay))
            empty[assay] = {k: open(os.path.join(tmp, assay + k)).read() for k in (".txt", ".bed")}
    dump("pcr_main.json", {"globals": pcr_globals, "tm_delta": 5.0, "cases": cases, "empty_reports": empty})


def golden_lib_markers(rng):
    """lib_markers.Marker (:297-389: '*' alleles, names, parse_indels, assert_marker), ROI sequences
    (:53-77: seq, seq_alt, seq_x) and ROI.upload_mutations (:279-294) on a synthetic genome."""
    import pandas as pd
    import polyoligo.lib_markers as lm

    class FakeBlast:
        def __init__(self, genome):
            self.genome = genome

        def fetch(self, query):
            q = query[0]
            return {"%s:%d-%d" % (q["chr"], q["start"], q["stop"]): self.genome[q["start"] - 1:q["stop"]]}

    genome = rand_seq(rng, 600, gc=0.5).lower()          # lower case on purpose: ROI upper-cases
    blast = FakeBlast(genome)
    cases = []
    specs = [("m_1", 100, None, "G"), (".", 200, "*", "ACG"), (None, 300, None, "*"), ("x", 301, "AC", "*"),
             ("snp_b", 50, None, "T"), ("mnp", 120, 3, "TT")]
    for name, pos, ref, alt in specs:
        if ref is None:
            ref = genome[pos].upper()
        elif isinstance(ref, int):
            ref = genome[pos:pos + ref].upper()
        elif ref == "AC":
            ref = genome[pos:pos + 2].upper()
        m = lm.Marker(name=name, chrom="chrM", pos=pos, ref_allele=ref, alt_allele=alt)
        before = dict(name=m.name, pos=m.pos, pos1=m.pos1, ref=m.ref, alt=m.alt, variant=m.variant, n=m.n)
        err = m.assert_marker(blast)
        m.parse_indels(blast)
        after = dict(ref=m.ref, alt=m.alt, variant=m.variant, n=m.n, is_indel=m.is_indel)
        rois = []
        for a, b in ((m.pos1 - 20, m.pos1 + 20), (m.pos1, m.pos1), (m.pos1 + 5, m.pos1 + 30)):
            roi = lm.ROI(chrom="chrM", start=a, stop=b, blast_hook=blast, marker=m)
            rois.append(dict(start=a, stop=b, name=roi.name, seq=roi.seq, seq_alt=roi.seq_alt, seq_x=roi.seq_x, n=roi.n))
        cases.append(dict(args=[name, pos, ref, alt], before=before, err=err, after=after, rois=rois))
    wrong = lm.Marker(name="bad", chrom="chrM", pos=10, ref_allele={"A": "C"}.get(genome[10].upper(), "A"), alt_allele="G")
    bad_err = wrong.assert_marker(blast)

    class FakeVcf:
        def fetch_genotypes(self, chrom, start, stop):
            pos = [p for p in (95, 101, 110, 140) if start <= p <= stop]
            G = pd.DataFrame({p: [0, 1, 2] for p in pos})
            return G, [types.SimpleNamespace(pos=p) for p in pos]

    ups = []
    for mpos in (100, 105):          # marker.pos1 = 101 is among the variants / 106 is not
        m = lm.Marker(name="u", chrom="chrM", pos=mpos, ref_allele=genome[mpos].upper(), alt_allele="T")
        roi = lm.ROI(chrom="chrM", start=90, stop=150, blast_hook=blast, vcf_hook=FakeVcf(), marker=m)
        roi.upload_mutations()
        ups.append(dict(marker_pos=mpos, mutation_pos=[x.pos for x in roi.mutations], columns=[int(c) for c in roi.G.columns]))
    dump("lib_markers.json", {"genome": genome, "cases": cases, "bad_args": ["bad", 10, wrong.ref, "G"], "bad_err": bad_err,
                              "upload": ups})


if __name__ == "__main__":
    if "--offtargets-only" in sys.argv:
        golden_offtargets(np.random.default_rng(20240611))
        sys.exit(0)
    if "--design-only" in sys.argv:
        golden_design(np.random.default_rng(20240610))
        sys.exit(0)
    if "--edges-only" in sys.argv:
        golden_marker_primers_edges(np.random.default_rng(20240614))
        sys.exit(0)
    if "--lib-markers-only" in sys.argv:
        golden_lib_markers(np.random.default_rng(20240613))
        sys.exit(0)
    if "--pcr-main-only" in sys.argv:
        golden_pcr_main(np.random.default_rng(20240612))
        sys.exit(0)
    rng = np.random.default_rng(20240608)
    golden_marker_primers(rng)
    golden_is_valid(rng)
    golden_heterodimer(rng)
    golden_scoring(rng)
    golden_masks(rng)
    golden_mutations(np.random.default_rng(20240609))
    golden_design(np.random.default_rng(20240610))
    golden_offtargets(np.random.default_rng(20240611))
    golden_pcr_main(np.random.default_rng(20240612))
    golden_lib_markers(np.random.default_rng(20240613))
    golden_marker_primers_edges(np.random.default_rng(20240614))
