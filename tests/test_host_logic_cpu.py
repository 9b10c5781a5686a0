"""CPU tests of the host logic against the golden fixtures generated from the
reference's own Python (tests/golden/make_golden.py): interval helpers, masks,
scoring, classify/prune order, shard partitioning and the multi-process gather."""
import json
import os
import types

import numpy as np
import pytest

from polyoligo_b200 import lib_primer3 as lp, lib_utils, scoring, sharding

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_reduce_ivs_matches_reference():
    for case in gold("masks.json")["reduce_ivs"]:
        assert lib_utils.reduce_ivs(case["in"], 200) == case["out"]
    assert lib_utils.reduce_ivs([], 200) == []
    assert lib_utils.reduce_ivs([[3, 2]], 200) == [[3, 2]]


def test_list_2_ranges():
    assert lib_utils.list_2_ranges([]) == []
    assert lib_utils.list_2_ranges([5]) == [[5, 1]]
    assert lib_utils.list_2_ranges([7, 1, 2, 3, 9, 8]) == [[1, 3], [7, 3]]


def test_masks_match_reference():
    lp.PRIMER3_GLOBALS.clear()
    lp.set_globals(PRIMER_MAX_SIZE=25)
    for case in gold("masks.json")["cases"]:
        n = case["n"]
        maps = {}
        for k in ("mism", "partial_mism"):
            m = np.ones(n, dtype=bool)
            m[case[k]] = False
            maps[k] = m
        maps["all"] = np.ones(n, dtype=bool)
        muts = [types.SimpleNamespace(pos=p) for p in case["mutation_pos"]]
        hroi = types.SimpleNamespace(start=case["start"], vcf_hook=True, mutations=muts,
                                     p3_sequence_included_maps=maps)
        lp.include_mut_in_included_maps(hroi)
        assert set(hroi.p3_sequence_included_maps) == set(case["excluded_idx_after_mut"])
        for k, idx in case["excluded_idx_after_mut"].items():
            assert list(np.flatnonzero(~hroi.p3_sequence_included_maps[k])) == idx
        lp.get_sequence_excluded_regions(hroi)      # PCR / CAPS / HRM: maps as they are
        assert hroi.p3_sequence_excluded_regions == case["pcr_excluded_regions"]
        lp.open_marker_window(hroi)                  # KASP quirk, restated literally
        lp.get_sequence_excluded_regions(hroi)
        assert hroi.p3_sequence_excluded_regions == case["kasp_excluded_regions"]


def fake_pp(row):
    keys = ["F", "R", "A"] if row["assay"] == "KASP" else ["F", "R"]
    pp = types.SimpleNamespace()
    pp.primers = {d: types.SimpleNamespace(tm=t, max_aaf=a) for d, t, a in zip(keys, row["tms"], row["max_aafs"])}
    pp.offtargets = ["x"] * row["n_offtargets"]
    pp.dimer, pp.ref_dimer, pp.alt_dimer = row["dimer"], row["ref_dimer"], row["alt_dimer"]
    pp.max_indel_size = row["max_indel_size"]
    pp.hrm = {"delta": row["hrm_delta"]}
    return pp


def test_scoring_matches_reference():
    rows = gold("scoring.json")["rows"]
    assert len(rows) == 240
    for row in rows:
        got = scoring.get_score_fnc(row["assay"])(fake_pp(row))
        assert (int(got[0]), got[1]) == (row["goodness"], row["qcode"]), row


def test_qcode_bits_roundtrip():
    from polyoligo_b200._native import Q_BITS
    for row in gold("scoring.json")["rows"]:
        bits = sum(Q_BITS[c] for c in row["qcode"])
        assert scoring.qcode_from_bits(bits) == row["qcode"]


def test_classify_prune_order_matches_reference():
    for key, case in gold("scoring.json")["prune"].items():
        assay, n = key.split("/")
        pcr = lp.PCR()
        for uid, (g, d) in enumerate(zip(case["goodness"], case["hrm_delta"])):
            pp = types.SimpleNamespace(goodness=g, hrm={"delta": d}, uid=uid)
            pp.score = lambda assay_name: None
            pcr.pps.append(pp)
        pcr.classify(assay)
        pcr.prune(int(n), assay_name=assay)
        order = [pp.uid for score in sorted(pcr.pps_pruned, reverse=True) for pp in pcr.pps_pruned[score]]
        if assay == "HRM":
            # The reference orders a goodness class with np.argsort(-delta), whose default
            # quicksort is not stable (its comment asks for "ties ordered"; the outcome for
            # equal deltas depends on the numpy build).  This package is stable; compare the
            # (goodness, delta) sequence, which is what the tie permutation cannot change.
            def keyseq(o):
                return [(case["goodness"][u], case["hrm_delta"][u]) for u in o]
            assert keyseq(order) == keyseq(case["order"]), key
            assert order == sorted(order, key=lambda u: (-case["goodness"][u], -case["hrm_delta"][u], u))
        else:
            assert order == case["order"], key


def test_partition_balances_and_covers():
    rng = np.random.default_rng(4)
    costs = rng.integers(1, 100, 1000)
    for shards in (1, 2, 3, 4, 8):
        b = sharding.partition(costs, shards)
        assert b[0][0] == 0 and b[-1][1] == 1000
        assert all(b[k][1] == b[k + 1][0] for k in range(shards - 1))
        loads = [costs[s:e].sum() for s, e in b]
        assert max(loads) <= costs.sum() / shards + costs.max()
    assert sharding.partition([], 4) == [(0, 0)] * 4
    assert sharding.partition([1, 1], 4)[-1][1] == 2


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = np.arange(101, dtype=np.int64)
    s, e = sharding.shard_bounds(len(units), rank, world, costs=units + 1.0)
    local = {"unit": units[s:e], "score": units[s:e] * 2.5}      # stand-in for a GPU scoring call
    merged = sharding.gather_to_rank0(local, dist)
    if rank == 0:
        q.put((merged["unit"].tolist(), merged["score"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    units, scores = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert units == list(range(101))
    assert scores == [u * 2.5 for u in range(101)]


def _set_kasp_design_globals():
    from polyoligo_b200 import lib_primer3 as lp, primer3_bindings as p3b
    p3b.resetP3Globals()
    lp.PRIMER3_GLOBALS.clear()
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0, PRIMER_MIN_TM=55.0,
                   PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0, PRIMER_MAX_POLY_X=100,
                   PRIMER_PAIR_MAX_DIFF_TM=6.0, PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200]],
                   PRIMER_NUM_RETURN=8)
    lp.set_tm_delta(5.0)


def _gloo_marker_worker(rank, world, port, q):
    # the real config-5 pipeline per rank; the CPU oracle context stands in for the GPU
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from oracle.cpu_pipeline import CpuContext
    from polyoligo_b200 import kasp_pipeline as kp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _set_kasp_design_globals()
    data = kp.synth_config5(5, seed=3, length=300)
    timing = {}
    merged = sharding.design_markers_sharded(data, rank, world, dist=dist, n_primers=3, ctx=CpuContext(nthreads=2),
                                             timing=timing)
    assert timing["design_s"] > 0
    if rank == 0:
        recs, n_sel = merged          # tensor gather (dist.gather of byte tensors), no pickling
        q.put((recs.tobytes(), n_sel.tolist()))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_marker_design_matches_single_process():
    import sys
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from oracle.cpu_pipeline import CpuContext
    from polyoligo_b200 import kasp_pipeline as kp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_marker_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _set_kasp_design_globals()
    data = kp.synth_config5(5, seed=3, length=300)
    one = kp.design_markers_arrays(data, n_primers=3, ctx=CpuContext(nthreads=2))
    assert len(one["f"]) > 5
    recs, n_sel = sharding.compact_selection(one, 3)
    assert recs.shape == (5, 3) and n_sel.sum() > 3
    assert got == (recs.tobytes(), n_sel.tolist())
    # the compact records are the selected pairs, in selection order
    m = int(np.argmax(n_sel > 0))
    first = one["sel"][m, 0]
    assert recs[m, 0]["f"].tobytes() == one["f"][first].tobytes() and recs[m, 0]["goodness"] == one["goodness"][first]
    # slicing keeps the mutation segments consistent
    part = sharding.slice_markers(data, 2, 5)
    assert len(part["alt"]) == 3 and part["mut_seg"][0] == 0 and part["mut_seg"][-1] == len(part["mut_pos"])


def test_ambiguous_sequences_and_mutation_labels():
    # string half of PrimerPair.add_mutations (reference lib_primer3.py:166-199) on hand-checked cases
    import types
    from polyoligo_b200 import lib_primer3 as lp, lib_utils
    assert lib_utils.seqs2ambiguous_dna(["ACGT", "ACGA", "TCGT"]) == "WCGW"
    assert lib_utils.seqs2ambiguous_dna(["A", "C", "G", "T"]) == "N"       # 'ACGT' is not a key of the inverted table
    assert lib_utils.reverse_complement("ACGTRYKM") == "KMRYACGT"
    pp = lp.PrimerPair()
    pp.dir = None
    pp.primers["F"] = lp.Primer(sequence="ACGTACGTAC", start=100, stop=109)
    pp.primers["R"] = lp.Primer(sequence="GGGTTTCCCA", start=200, stop=209)   # reverse strand: TGGGAAACCC on the template
    muts = [types.SimpleNamespace(pos=102, ref="G", alt=["A"], aaf=0.25, indel_size=0),
            types.SimpleNamespace(pos=200, ref="T", alt=["C", "G"], aaf=0.125, indel_size=3),
            types.SimpleNamespace(pos=150, ref="A", alt=["T"], aaf=0.4, indel_size=7)]
    pp.add_mutations(muts)
    assert pp.primers["F"].sequence_ambiguous == "ACRTACGTAC" and pp.primers["F"].mutations == ["G102A:0.25"]
    assert pp.primers["R"].sequence_ambiguous == "GGGTTTCCCV" and pp.primers["R"].mutations == ["T200C/G:0.12"]
    assert (pp.primers["F"].max_aaf, pp.primers["R"].max_aaf, pp.max_indel_size) == (0.25, 0.125, 7)


def test_list_offtargets_matches_reference_fixture():
    import json, os
    from polyoligo_b200 import lib_primer3 as lp
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "offtargets.json")))
    lp.set_offtarget_size(fx["min_size"], fx["max_size"])
    n = 0
    for case in fx["cases"]:
        assert lp.list_offtargets(case["hits"]) == case["offtargets"]
        n += len(case["offtargets"])
    assert n > 20
