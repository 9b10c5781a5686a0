"""SURVEY 8f "next" rows: HRM product Tm (Biopython Tm_NN) and mutation annotation."""
import json
import os
import types

import numpy as np
import pytest

from helpers import rand_seqs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_tm_nn_oracle_biopython_doc_anchor():
    # Biopython tutorial / MeltingTemp docstring: Tm_NN('CGTTCCAAAGATGTGGGCATGAGCTTAC') prints 60.32
    from oracle.polyoligo_oracle import tm_nn
    assert "%0.2f" % tm_nn("CGTTCCAAAGATGTGGGCATGAGCTTAC") == "60.32"
    # Tm_NN(seq, strict=False) as the reference calls it (lib_markers.py:586-587): _check() DROPS
    # ambiguous bases, it does not raise -- an N-padded product still gets its temperatures
    assert tm_nn("ACGTNACGTRY") == tm_nn("ACGTACGT")
    assert tm_nn("nnACGuACGT") == tm_nn("ACGTACGT")
    with pytest.raises(ValueError):
        tm_nn("NNNN")


def test_annotate_mutations_oracle_matches_reference_fixture():
    from oracle.polyoligo_oracle import annotate_mutations
    with open(os.path.join(GOLD, "mutations.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        win = list(zip(c["start"], c["stop"]))
        aaf, indel = annotate_mutations([tuple(m) for m in c["mutations"]], win, (c["start"][0], c["stop"][1]))
        assert aaf == c["max_aaf"] and indel == c["max_indel"]


@pytest.mark.gpu
def test_tm_nn_kernel_matches_oracle(ctx):
    from oracle.polyoligo_oracle import tm_nn
    rng = np.random.default_rng(12)
    seqs = rand_seqs(rng, 3000, 40, 150) + ["CGTTCCAAAGATGTGGGCATGAGCTTAC", "AT", "acgtacgtac", "A" * 500]
    got = ctx.tm_nn_batch(seqs)
    assert np.array_equal(got, np.array([tm_nn(s) for s in seqs]))
    assert "%0.2f" % got[3000] == "60.32"
    odd = ["ACGTNACGTRY", "NNACGTACGTNN", "ACGTIACGT", "IACGTACG", "nnacgu", "N" * 30 + seqs[0] + "N" * 30]
    assert np.array_equal(ctx.tm_nn_batch(odd), np.array([tm_nn(s) for s in odd]))
    assert np.isnan(ctx.tm_nn_batch(["NNNN", "ACGT"])[0])
    from polyoligo_b200 import lib_primer3 as lp
    out = lp.hrm_temps_batch(seqs[:5], [s[:20] + ("A" if s[20] != "A" else "C") + s[21:] for s in seqs[:5]])
    assert all(set(o) == {"ref", "alt", "delta"} and o["delta"] >= 0 for o in out)


@pytest.mark.gpu
def test_add_mutations_batch_matches_reference_fixture():
    from polyoligo_b200 import kasp, lib_primer3 as lp
    with open(os.path.join(GOLD, "mutations.json")) as f:
        cases = json.load(f)["cases"]
    for is_kasp in (False, True):
        sub = [c for c in cases if c["kasp"] == is_kasp]
        pcrs, muts = [], []
        for c in sub:
            pp = kasp.PrimerPair() if is_kasp else lp.PrimerPair()
            for d, s, e in zip(pp.primers, c["start"], c["stop"]):
                pp.primers[d].start, pp.primers[d].stop = s, e
            pcrs.append(lp.PCR(primer_pairs=[pp]))
            muts.append([types.SimpleNamespace(pos=m[0], aaf=m[1], indel_size=m[2]) for m in c["mutations"]])
        lp.add_mutations_batch(pcrs, muts)
        for c, pcr in zip(sub, pcrs):
            pp = pcr.pps[0]
            assert [pp.primers[d].max_aaf for d in pp.primers] == c["max_aaf"]
            assert pp.max_indel_size == c["max_indel"]
