"""Parity gate against vectors from the REAL primer3 (written by scripts/repin.py).

The vectors do not exist until someone runs ``python scripts/repin.py`` where ``import primer3``
works (or with POLYOLIGO_P3_CONFIG=<primer3_config dir>): until then these tests SKIP and thal /
the picker stay parity-unpinned.  Tolerances are north_star's: Tm 0.01 C,
dG 1e-3 kcal/mol (1 cal/mol), structure_found and the selected primers identical."""
import json
import os

import numpy as np
import pytest

import polyoligo_b200 as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.environ.get("POLYOLIGO_GOLDEN_P3", os.path.join(ROOT, "tests", "golden", "primer3"))
PARAMS = os.environ.get("POLYOLIGO_P3_CONFIG")
KINDS = ("any", "end1", "end2", "hairpin", "tailed")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "source.json")),
                                reason="no primer3 golden vectors (run scripts/repin.py): PARITY UNPINNED")


def _load(kind):
    z = np.load(os.path.join(GOLDEN, "thal_%s.npz" % kind))
    a = [str(x) for x in z["a"]]
    b = [str(x) for x in z["b"]] if len(z["b"]) else None
    return a, b, z["res"]


def _assert_close(got, want, what):
    assert np.array_equal(got[:, 4], want[:, 4]), what + ": structure_found differs"
    assert np.max(np.abs(got[:, 0] - want[:, 0])) <= 0.01, what + ": Tm differs by more than 0.01 C"
    assert np.max(np.abs(got[:, 1] - want[:, 1])) <= 1.0, what + ": dG differs by more than 1 cal/mol"


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_primer3(kind):
    from oracle.oracle import Oracle
    orc = Oracle(PARAMS)
    a, b, want = _load(kind)
    t = {"any": 1, "tailed": 1, "end1": 2, "end2": 3, "hairpin": 4}[kind]
    r = orc.thal_batch(a, None if kind == "hairpin" else b, t, nthreads=8)
    got = np.stack([r["tm"], r["dg"], r["dh"], r["ds"], r["structure_found"].astype(np.float64)], axis=1)
    _assert_close(got, want, "oracle thal " + kind)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_matches_primer3(kind):
    ctx = po.Context(0, PARAMS)
    try:
        a, b, want = _load(kind)
        t = {"any": po.THAL_ANY, "tailed": po.THAL_ANY, "end1": po.THAL_END1, "end2": po.THAL_END2,
             "hairpin": po.THAL_HAIRPIN}[kind]
        r = ctx.thal_batch(t, po.pack(a), po.pack(b) if b is not None else None)
        got = np.stack([r["tm"], r["dg"], r["dh"], r["ds"], (r["flags"] & 1).astype(np.float64)], axis=1)
        _assert_close(got, want, "CUDA thal " + kind)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_picker_matches_primer3():
    from polyoligo_b200 import primer3_bindings as p3b
    with open(os.path.join(GOLDEN, "design_kasp.json")) as f:
        d = json.load(f)
    p3b.resetP3Globals()
    p3b.setP3Globals(d["globals"])
    try:
        res = p3b.design_primers_batch([{"SEQUENCE_TEMPLATE": c["template"], "SEQUENCE_PRIMER": c["fixed_left"],
                                         "SEQUENCE_EXCLUDED_REGION": c["excluded"]} for c in d["cases"]])
    finally:
        p3b.resetP3Globals()
    for c, r in zip(d["cases"], res):
        n = int(r.get("PRIMER_PAIR_NUM_RETURNED", 0))
        assert [r["PRIMER_RIGHT_%d_SEQUENCE" % i] for i in range(n)] == [p["right"] for p in c["pairs"]]
