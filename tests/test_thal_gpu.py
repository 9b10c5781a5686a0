"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): decisions bit-exact, dG within 1e-3 kcal/mol
(= 1 cal/mol in primer3-py units), Tm within 0.01 C.  The kernels are written
to reproduce the scalar algorithm's IEEE-double arithmetic operation for
operation, so these tests demand BIT-EXACT tm/dg/dh/ds, which implies the
tolerances above.
"""
import os

import numpy as np
import pytest

import polyoligo_b200 as po
from polyoligo_b200 import synth
from helpers import adversarial_pairs, config4_pairs, rand_seqs, revcomp, VIC, FAM

pytestmark = pytest.mark.gpu

DG_TOL_CAL = 1.0     # 1e-3 kcal/mol
TM_TOL_C = 0.01


def assert_parity(got, ref, what):
    sf = got["flags"] & po.ITEM_STRUCTURE_FOUND
    assert np.array_equal(sf, ref["structure_found"]), what + ": structure_found differs"
    assert np.max(np.abs(got["tm"] - ref["tm"]), initial=0) <= TM_TOL_C, what
    assert np.max(np.abs(got["dg"] - ref["dg"]), initial=0) <= DG_TOL_CAL, what
    for f in ("tm", "dg", "dh", "ds"):
        assert np.array_equal(got[f], ref[f]), "%s: %s not bit-exact" % (what, f)


def test_heterodimer_config4_shape(ctx, oracle):
    a, b = config4_pairs(20000)
    got = ctx.thal_batch(po.THAL_ANY, po.pack(a), po.pack(b))
    ref = oracle.thal_batch(a, b, 1, nthreads=8)
    assert_parity(got, ref, "heterodimer")
    found = ref["structure_found"] == 1
    assert np.array_equal(got["align_end_1"][found], ref["align_end_1"][found])
    assert np.array_equal(got["align_end_2"][found], ref["align_end_2"][found])


def test_homodimer_and_hairpin_config4_shape(ctx, oracle):
    a, _ = config4_pairs(20000, seed=11)
    pa = po.pack(a)
    assert_parity(ctx.thal_batch(po.THAL_ANY, pa), oracle.thal_batch(a, a, 1, nthreads=8), "homodimer")
    assert_parity(ctx.thal_batch(po.THAL_HAIRPIN, pa), oracle.thal_batch(a, None, 4, nthreads=8), "hairpin")


def test_adversarial_inputs(ctx, oracle):
    rng = np.random.default_rng(99)
    a, b = adversarial_pairs(rng)
    pa, pb = po.pack(a), po.pack(b)
    assert_parity(ctx.thal_batch(po.THAL_ANY, pa, pb), oracle.thal_batch(a, b, 1, nthreads=8), "adversarial het")
    assert_parity(ctx.thal_batch(po.THAL_ANY, pa), oracle.thal_batch(a, a, 1, nthreads=8), "adversarial homodimer")
    assert_parity(ctx.thal_batch(po.THAL_HAIRPIN, pa), oracle.thal_batch(a, None, 4, nthreads=8), "adversarial hairpin")
    assert_parity(ctx.thal_batch(po.THAL_HAIRPIN, pb), oracle.thal_batch(b, None, 4, nthreads=8), "adversarial hairpin b")


def test_end_anchored_types(ctx, oracle):
    # thal END1 / END2 (3'-anchored): what primer3's self_end_th / compl_end_th are built from
    a, b = config4_pairs(6000, seed=21)
    rng = np.random.default_rng(17)
    x, y = adversarial_pairs(rng)
    a, b = a + x, b + y
    pa, pb = po.pack(a), po.pack(b)
    for t in (po.THAL_END1, po.THAL_END2):
        got = ctx.thal_batch(t, pa, pb)
        ref = oracle.thal_batch(a, b, t, nthreads=8)
        assert_parity(got, ref, "end type %d" % t)
        found = ref["structure_found"] == 1
        assert np.array_equal(got["align_end_1"][found], ref["align_end_1"][found])
        assert np.array_equal(got["align_end_2"][found], ref["align_end_2"][found])
    assert_parity(ctx.thal_batch(po.THAL_END1, pa), oracle.thal_batch(a, a, 2, nthreads=8), "self end")


def test_kasp_tailed_shape(ctx, oracle):
    # reporter + allele-specific primer against the common primer (reference _lib_kasp.py:95-105)
    rng = np.random.default_rng(5)
    core = rand_seqs(rng, 4000, 18, 26)
    com = rand_seqs(rng, 4000, 18, 25)
    tailed = [(VIC if k % 2 else FAM) + s for k, s in enumerate(core)]
    got = ctx.thal_batch(po.THAL_ANY, po.pack(tailed), po.pack(com))
    assert_parity(got, oracle.thal_batch(tailed, com, 1, nthreads=8), "tailed F/R")
    got = ctx.thal_batch(po.THAL_ANY, po.pack(com), po.pack(tailed))   # dir == "R": (F, A) order
    assert_parity(got, oracle.thal_batch(com, tailed, 1, nthreads=8), "tailed R order")


def test_tm_and_gc(ctx, oracle):
    rng = np.random.default_rng(1)
    seqs = rand_seqs(rng, 5000, 2, 60) + ["GTAAAACGACGGCCAGT", "ACGT", "GAATTC", "AT", "GC" * 30]
    tm, gc = ctx.tm_batch(po.pack(seqs), with_gc=True)
    assert np.array_equal(tm, oracle.tm_batch(seqs))
    assert tm[5000] == 49.16808228911765   # primer3-py README anchor
    expect_gc = np.array([100.0 * (s.count("G") + s.count("C")) / len(s) for s in seqs])
    assert np.array_equal(gc, expect_gc)


def _anchor_ctx(ctx):
    d = os.environ.get("POLYOLIGO_P3_CONFIG")
    return (po.Context(0, d), True) if d else (ctx, False)


def test_readme_hairpin_anchor_structure(ctx):
    c, own = _anchor_ctx(ctx)
    out = c.thal_batch(po.THAL_HAIRPIN, po.pack(["CCCCCATCCGATCAGGGGG"]))
    if own:
        c.close()
    assert out["flags"][0] & po.ITEM_STRUCTURE_FOUND
    assert out["dh"][0] == -36300.0


@pytest.mark.xfail(condition=not os.environ.get("POLYOLIGO_P3_CONFIG"), strict=True,
                   reason="thal PARITY UNPINNED: the embedded literature tables miss upstream's ds by 10.02 e.u. "
                          "(scripts/localise_anchor.py); a plain assertion with POLYOLIGO_P3_CONFIG=<primer3_config>")
def test_readme_hairpin_anchor_upstream_values(ctx):
    # primer3-py README: calcHairpin('CCCCCATCCGATCAGGGGG') -> tm=34.15, dg=337.09, dh=-36300.00, ds=-118.13
    c, own = _anchor_ctx(ctx)
    out = c.thal_batch(po.THAL_HAIRPIN, po.pack(["CCCCCATCCGATCAGGGGG"]))
    if own:
        c.close()
    assert round(float(out["tm"][0]), 2) == 34.15
    assert round(float(out["dg"][0]), 2) == 337.09
    assert round(float(out["ds"][0]), 2) == -118.13


def test_skipped_items_and_empty_batch(ctx):
    seqs = ["ACGTNACGTACGTACGTAC", "", "A" * 61, "ACGTACGTACGTACGTACGT", "ACG-ACGT"]
    out = ctx.thal_batch(po.THAL_HAIRPIN, po.pack(seqs))
    assert list((out["flags"] & po.ITEM_SKIPPED) != 0) == [True, True, True, False, True]
    assert np.all(out["tm"][[0, 1, 2, 4]] == 0.0)
    out = ctx.thal_batch(po.THAL_ANY, po.pack(seqs), po.pack(seqs[::-1]))
    assert np.all((out["flags"] & po.ITEM_SKIPPED) != 0)
    assert len(ctx.thal_batch(po.THAL_ANY, po.pack([]), po.pack([]))) == 0
    assert len(ctx.tm_batch(po.pack([]))) == 0


def test_conditions_follow_set_conditions(oracle):
    c = po.Context(0)
    try:
        c.set_conditions(mv=75.0, dv=1.5, dntp=0.6, dna_conc=200.0, temp_c=25.0, max_loop=12)
        a, b = config4_pairs(3000, seed=3)
        kw = dict(mv=75.0, dv=1.5, dntp=0.6, dna_conc=200.0, temp_k=25.0 + 273.15, max_loop=12)
        assert_parity(c.thal_batch(po.THAL_ANY, po.pack(a), po.pack(b)),
                      oracle.thal_batch(a, b, 1, nthreads=8, **kw), "het, custom conditions")
        assert_parity(c.thal_batch(po.THAL_HAIRPIN, po.pack(a)),
                      oracle.thal_batch(a, None, 4, nthreads=8, **kw), "hairpin, custom conditions")
        assert np.array_equal(c.tm_batch(po.pack(a)),
                              oracle.tm_batch(a, mv=75.0, dv=1.5, dntp=0.6, dna_conc=200.0))
    finally:
        c.close()


def test_synthetic_tetraloop_table_path(tmp_path, oracle):
    # the default set ships no tetraloop list; exercise the lookup with a synthetic one
    import os
    import shutil
    from oracle.oracle import Oracle
    src = os.path.join(os.path.dirname(po.__file__), "params")
    dst = str(tmp_path / "params")
    shutil.copytree(src, dst)
    loops = ["CGAAAG", "GGAAAC", "AGCAAT", "TTTTTA", "CTTCGG"]
    with open(os.path.join(dst, "tetraloop.dh"), "w") as f:
        f.write("".join("%s\t%d\n" % (k, -1500 - 100 * i) for i, k in enumerate(loops)))
    with open(os.path.join(dst, "tetraloop.ds"), "w") as f:
        f.write("".join("%s\t%.2f\n" % (k, -0.65 * (i + 1)) for i, k in enumerate(loops)))
    rng = np.random.default_rng(8)
    stems = rand_seqs(rng, 600, 4, 10)
    seqs = [s + loops[k % 5][1:5] + revcomp(s) for k, s in enumerate(stems)]
    seqs = [s for s in seqs if len(s) <= 60]
    c = po.Context(0, param_dir=dst)
    try:
        o2 = Oracle(dst)
        got = c.thal_batch(po.THAL_HAIRPIN, po.pack(seqs))
        assert_parity(got, o2.thal_batch(seqs, None, 4, nthreads=8), "hairpin with tetraloop table")
        base = oracle.thal_batch(seqs, None, 4, nthreads=8)
        assert np.any(base["dh"] != got["dh"])   # the bonus table really took part
    finally:
        c.close()


def test_score_pairs_and_props(ctx, oracle):
    a, b = synth.config4(6000, seed=77)
    out = ctx.score_pairs(a, b)
    sa, sb = synth.to_strings(a), synth.to_strings(b)
    assert_parity(out["het"], oracle.thal_batch(sa, sb, 1, nthreads=8), "score_pairs het")
    assert_parity(out["hairpin_a"], oracle.thal_batch(sa, None, 4, nthreads=8), "score_pairs hairpin a")
    assert_parity(out["hairpin_b"], oracle.thal_batch(sb, None, 4, nthreads=8), "score_pairs hairpin b")
    assert np.array_equal(out["tm_a"], oracle.tm_batch(sa)) and np.array_equal(out["tm_b"], oracle.tm_batch(sb))
    # Primer.calc_thermals / is_valid (reference lib_primer3.py:56-97) with the KASP YAML thresholds
    v = po.ValidParams(18, 25, 20.0, 80.0, 55.0, 65.0, 5.0)
    props = ctx.oligo_props_batch(a, v)
    tm = oracle.tm_batch(sa)
    hp = oracle.thal_batch(sa, None, 4, nthreads=8)["tm"]
    hd = oracle.thal_batch(sa, sa, 1, nthreads=8)["tm"]
    lens = np.array([len(s) for s in sa])
    gc = np.array([100.0 * (s.count("G") + s.count("C")) / len(s) for s in sa])
    ok = (lens >= 18) & (lens <= 25) & (gc >= 20.0) & (gc <= 80.0) & (tm >= 55.0) & (tm <= 65.0) \
        & ~(hp >= tm - 5.0) & ~(hd >= tm - 5.0)
    assert np.array_equal(props["tm"], tm) and np.array_equal(props["hairpin_tm"], hp)
    assert np.array_equal(props["homodimer_tm"], hd) and np.array_equal(props["gc_content"], gc)
    assert np.array_equal((props["flags"] & po.PRIMER_VALID) != 0, ok)
    assert 0 < ok.sum() < len(ok)


def test_full_size_properties(ctx, oracle):
    """BASELINE config 4 at a large size (2M pairs here; bench.py runs the full 10^7):
    size-independent properties instead of a full oracle pass."""
    n = 2_000_000
    a, b = synth.config4(n, seed=synth.CONFIG4_SEED)
    het = ctx.thal_batch(po.THAL_ANY, a, b)
    # determinism: the dynamic work queue must not influence results
    again = ctx.thal_batch(po.THAL_ANY, a, b)
    assert het.tobytes() == again.tobytes()
    # thermodynamic identities on every item
    found = (het["flags"] & po.ITEM_STRUCTURE_FOUND) != 0
    assert np.all(het["tm"][~found] == 0.0)
    assert np.allclose(het["dg"][found], het["dh"][found] - 310.15 * het["ds"][found], rtol=0, atol=1e-6)
    assert np.all(np.isfinite(het["dh"][found])) and np.all(np.isfinite(het["tm"][found]))
    assert np.all((het["flags"] & po.ITEM_SKIPPED) == 0)
    # homodimer(a) == heterodimer(a, a)
    k = 200_000
    assert ctx.thal_batch(po.THAL_ANY, a[:k]).tobytes() == ctx.thal_batch(po.THAL_ANY, a[:k], a[:k]).tobytes()
    # batch composition must not matter: a strided subset recomputed alone is identical
    idx = np.arange(0, n, 97)
    sub = ctx.thal_batch(po.THAL_ANY, a[idx], b[idx])
    assert sub.tobytes() == het[idx].tobytes()
    # and a seeded sample agrees with the oracle
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(n, 5000, replace=False))
    ref = oracle.thal_batch(synth.to_strings(a[pick]), synth.to_strings(b[pick]), 1, nthreads=8)
    assert_parity(het[pick], ref, "full-size sample")


def test_relaxation_counts_match_oracle(ctx, oracle):
    a, b = config4_pairs(3000, seed=21)
    ref = oracle.thal_batch(a, b, 1, nthreads=8)
    assert ctx.count_relaxations(po.THAL_ANY, po.pack(a), po.pack(b)) == int(ref["relaxations"].sum())
    refh = oracle.thal_batch(a, None, 4, nthreads=8)
    assert ctx.count_relaxations(po.THAL_HAIRPIN, po.pack(a)) == int(refh["relaxations"].sum())
