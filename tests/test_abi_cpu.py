"""CPU tests of the drop-in boundary: the shared library loads without a GPU and
exports every symbol include/polyoligo_b200.h declares; host-only entry points
(packing) behave; GPU entry points fail loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

import polyoligo_b200 as po
from polyoligo_b200 import _native, synth
from helpers import rand_seqs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "polyoligo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(po_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    assert set(names) == set(_native.EXPORTS), "python binding table and header disagree"


def test_struct_layouts_match_header():
    assert po.OLIGO_DTYPE.itemsize == 16
    assert po.THAL_OUT_DTYPE.itemsize == 40
    assert po.PROPS_DTYPE.itemsize == 40
    assert ctypes.sizeof(po.ValidParams) == 48


def test_pack_roundtrip_and_flags():
    rng = np.random.default_rng(2)
    seqs = rand_seqs(rng, 2000, 0, 60)
    rec = po.pack(seqs)
    assert po.unpack(rec) == seqs
    meta = rec["w"][:, 3] >> 24
    assert np.array_equal(meta & 0x3F, [len(s) for s in seqs])
    assert np.all((meta & 0xC0) == 0)
    rec = po.pack(["acgtn", "ACGTRYKM", "AC-GT", "A" * 61, "acgt"])
    meta = rec["w"][:, 3] >> 24
    assert list(meta & 0x80) == [0x80, 0x80, 0, 0, 0]
    assert list(meta & 0x40) == [0, 0, 0x40, 0x40, 0]
    assert po.unpack(rec[4:]) == ["ACGT"]          # case-insensitive


def test_synth_pack_equals_c_pack():
    a, b = synth.config4(5000)
    assert np.array_equal(po.pack(synth.to_strings(a))["w"], a["w"])
    lens = (a["w"][:, 3] >> 24) & 0x3F
    assert lens.min() >= 18 and lens.max() <= 30
    a2, _ = synth.config4(5000)
    assert a.tobytes() == a2.tobytes()             # seeded, reproducible


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="po_init failed"):
        po.Context(0)


def test_bad_param_dir_reports_error(tmp_path):
    h = ctypes.c_void_p()
    lib = _native.load()
    rc = lib.po_init(0, os.fsencode(str(tmp_path)), ctypes.byref(h))
    assert rc == 3      # PO_E_PARAMS, reported before any CUDA call
    assert b"cannot open" in lib.po_last_error(None)
