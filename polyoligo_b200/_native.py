"""ctypes binding of libpolyoligo_b200.so (C ABI in include/polyoligo_b200.h).

There is no CPU fallback: if the CUDA library is missing or no context can be
created the calls raise.  Packing (``pack`` / ``unpack``) is host-only and
works without a GPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POLYOLIGO_B200_LIB", os.path.join(_HERE, "libpolyoligo_b200.so"))

THAL_ANY, THAL_END1, THAL_END2, THAL_HAIRPIN = 1, 2, 3, 4
ITEM_STRUCTURE_FOUND, ITEM_SKIPPED, ITEM_SYMMETRIC, PRIMER_VALID = 0x1, 0x2, 0x4, 0x100
MAX_OLIGO_LEN = 60

OLIGO_DTYPE = np.dtype([("w", "<u4", (4,))])
THAL_OUT_DTYPE = np.dtype([("tm", "<f8"), ("dg", "<f8"), ("dh", "<f8"), ("ds", "<f8"),
                           ("flags", "<i4"), ("align_end_1", "<i2"), ("align_end_2", "<i2")])
PROPS_DTYPE = np.dtype([("tm", "<f8"), ("hairpin_tm", "<f8"), ("homodimer_tm", "<f8"),
                        ("gc_content", "<f8"), ("length", "<i4"), ("flags", "<i4")])
KASP_MARKER_DTYPE = np.dtype([("tmpl", "<u4", (32,)), ("nmask", "<u4", (16,)), ("alt", "<u4", (2,)),
                              ("alt_nmask", "<u4"), ("tmpl_len", "<i4"), ("marker_pos", "<i4"),
                              ("marker_pos_rc", "<i4"), ("ref_len", "<i4"), ("alt_len", "<i4"),
                              ("reserved", "<i4", (8,))])
assert OLIGO_DTYPE.itemsize == 16 and THAL_OUT_DTYPE.itemsize == 40 and PROPS_DTYPE.itemsize == 40
assert KASP_MARKER_DTYPE.itemsize == 256

ASSAY_PCR, ASSAY_KASP, ASSAY_HRM = 0, 1, 2
Q_BITS = {"t": 0x001, "O": 0x002, "d": 0x004, "m": 0x008, "M": 0x010, "i": 0x020, "I": 0x040, "h": 0x080,
          "H": 0x100}
Q_REF_DIMER, Q_ALT_DIMER = 0x1000, 0x2000


class RankIn(C.Structure):
    _fields_ = [("assay", C.c_int32), ("top_n", C.c_int32), ("n_pairs", C.c_int64), ("n_markers", C.c_int64),
                ("f", C.c_void_p), ("r", C.c_void_p), ("a", C.c_void_p), ("dir", C.c_void_p), ("seg", C.c_void_p),
                ("n_offtargets", C.c_void_p), ("max_aaf", C.c_void_p), ("max_indel", C.c_void_p),
                ("hrm_delta", C.c_void_p), ("reporters", C.c_uint32 * 8), ("tm_delta", C.c_double)]


class RankOut(C.Structure):
    _fields_ = [("goodness", C.c_void_p), ("qbits", C.c_void_p), ("tm", C.c_void_p), ("het_tm", C.c_void_p),
                ("sel", C.c_void_p), ("n_sel", C.c_void_p)]


class ValidParams(C.Structure):
    _fields_ = [("min_size", C.c_int32), ("max_size", C.c_int32),
                ("min_gc", C.c_double), ("max_gc", C.c_double),
                ("min_tm", C.c_double), ("max_tm", C.c_double),
                ("tm_delta", C.c_double)]


class P3Settings(C.Structure):
    """po_p3_settings; defaults are libprimer3's (version-2 defaults, thermodynamic alignment)."""
    _fields_ = [("min_size", C.c_int32), ("opt_size", C.c_int32), ("max_size", C.c_int32), ("num_return", C.c_int32),
                ("min_tm", C.c_double), ("opt_tm", C.c_double), ("max_tm", C.c_double),
                ("min_gc", C.c_double), ("max_gc", C.c_double),
                ("max_self_any_th", C.c_double), ("max_self_end_th", C.c_double), ("max_hairpin_th", C.c_double),
                ("pair_max_compl_any_th", C.c_double), ("pair_max_compl_end_th", C.c_double),
                ("pair_max_diff_tm", C.c_double),
                ("wt_tm_gt", C.c_double), ("wt_tm_lt", C.c_double), ("wt_size_lt", C.c_double),
                ("wt_size_gt", C.c_double),
                ("mv_conc", C.c_double), ("dv_conc", C.c_double), ("dntp_conc", C.c_double), ("dna_conc", C.c_double),
                ("max_poly_x", C.c_int32), ("gc_clamp", C.c_int32), ("max_end_gc", C.c_int32),
                ("n_size_ranges", C.c_int32), ("size_lo", C.c_int32 * 8), ("size_hi", C.c_int32 * 8)]


class KaspDesignParams(C.Structure):
    """po_kasp_design_params"""
    _fields_ = [("min_size", C.c_int32), ("max_size", C.c_int32), ("top_n", C.c_int32), ("with_mutations", C.c_int32),
                ("tm_delta", C.c_double), ("valid", ValidParams), ("reporters", C.c_uint32 * 8)]


# po_kasp_pair: one selected primer triple of a marker (also sharding.SEL_DTYPE)
KASP_PAIR_DTYPE = np.dtype([("f", "<u4", (4,)), ("r", "<u4", (4,)), ("a", "<u4", (4,)), ("goodness", "<i4"),
                            ("qbits", "<i4"), ("dir", "u1"), ("pad", "u1", (7,))])
assert KASP_PAIR_DTYPE.itemsize == 64

P3_OLIGO_DTYPE = np.dtype([("seq", OLIGO_DTYPE), ("tm", "<f8"), ("gc_percent", "<f8"), ("penalty", "<f8"),
                           ("self_any_th", "<f8"), ("self_end_th", "<f8"), ("hairpin_th", "<f8"),
                           ("start", "<i4"), ("len", "<i4"), ("set", "<i4"), ("flags", "<i4")])
P3_TASK_DTYPE = np.dtype([("set", "<i4"), ("side", "<i4"), ("start", "<i4"), ("len", "<i4")])
P3_PAIR_DTYPE = np.dtype([("other", P3_OLIGO_DTYPE), ("pair_penalty", "<f8"), ("compl_any_th", "<f8"),
                          ("compl_end_th", "<f8"), ("product_size", "<i4"), ("size_range", "<i4")])
assert P3_OLIGO_DTYPE.itemsize == 80 and P3_PAIR_DTYPE.itemsize == 112 and P3_TASK_DTYPE.itemsize == 16
P3_OK, P3_SELF_DONE, P3_SELF_FAILED, P3_PRIMER_VALID = 1, 2, 4, 16


# every symbol include/polyoligo_b200.h declares
EXPORTS = {
    "po_init": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "po_destroy": (None, [C.c_void_p]),
    "po_last_error": (C.c_char_p, [C.c_void_p]),
    "po_set_conditions": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.c_int, C.c_int]),
    "po_pack": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "po_unpack": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "po_tm_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "po_thal_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "po_oligo_props_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "po_tm_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_thal_batch_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]),
    "po_score_pairs_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_score_pairs_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_kasp_enumerate_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "po_rank_pairs_batch": (C.c_int, [C.c_void_p, C.POINTER(RankIn), C.POINTER(RankOut)]),
    "po_tm_nn_batch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "po_annotate_mutations_batch": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p,
                                                                               C.c_int, C.c_void_p, C.c_void_p]),
    "po_p3_candidates_batch": (C.c_int, [C.c_void_p, C.POINTER(P3Settings), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]),
    "po_p3_design_fixed_batch": (C.c_int, [C.c_void_p, C.POINTER(P3Settings), C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(ValidParams),
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_kasp_design_batch": (C.c_int, [C.c_void_p, C.POINTER(P3Settings), C.POINTER(KaspDesignParams), C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_kasp_masks_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "po_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "po_launch_count": (C.c_int64, [C.c_void_p]),
    "po_thal_item_count": (C.c_int64, [C.c_void_p]),
    "po_thal_count_relaxations": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.POINTER(C.c_int64)]),
    "po_measure_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Loads the shared library (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "polyoligo_b200: %s is missing -- run `python -m polyoligo_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _seq_blob(seqs):
    if isinstance(seqs, tuple):
        blob, off = seqs
        return np.ascontiguousarray(blob, dtype=np.uint8), np.ascontiguousarray(off, dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        np.cumsum([len(s) for s in seqs], out=off[1:])
    blob = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8)
    return blob, off


def pack(seqs):
    """list[str] (or (uint8 blob, int64 offsets)) -> structured array of 16-byte oligo records."""
    blob, off = _seq_blob(seqs)
    n = len(off) - 1
    out = np.zeros(n, dtype=OLIGO_DTYPE)
    buf = blob.tobytes() if blob.size else b""
    rc = load().po_pack(buf, off.ctypes.data, n, out.ctypes.data)
    if rc:
        raise RuntimeError("po_pack failed (%d)" % rc)
    return out


def unpack(oligos):
    oligos = np.ascontiguousarray(oligos, dtype=OLIGO_DTYPE)
    n = len(oligos)
    buf = np.zeros(61 * n, dtype=np.uint8)
    rc = load().po_unpack(oligos.ctypes.data, n, buf.ctypes.data)
    if rc:
        raise RuntimeError("po_unpack failed (%d)" % rc)
    raw = buf.reshape(n, 61)
    return [bytes(r[:int(np.argmax(r == 0))]).decode("ascii") for r in raw]


class ParityWarning(UserWarning):
    """thal runs on the embedded literature tables, not on primer3's own primer3_config/ files."""


_warned = False


def _warn_embedded_tables():
    """Once per process: thal results on the embedded tables are UNPINNED against primer3 (DESIGN.md
    section 2; the one published anchor differs by 10 e.u. of dS).  Pass primer3's primer3_config/
    directory as ``param_dir`` or set POLYOLIGO_P3_CONFIG to reproduce an installed primer3."""
    global _warned
    if not _warned:
        _warned = True
        import warnings
        warnings.warn("polyoligo_b200: thal uses the embedded literature parameter tables; hairpin / dimer Tm are "
                      "not pinned against primer3 -- set POLYOLIGO_P3_CONFIG=<primer3_config dir> (or param_dir) "
                      "to load upstream's tables", ParityWarning, stacklevel=3)


class Context:
    """One CUDA context of the scoring library on one GPU."""

    def __init__(self, device=0, param_dir=None):
        lib = load()
        h = C.c_void_p()
        if param_dir is None:
            param_dir = os.environ.get("POLYOLIGO_P3_CONFIG") or None    # upstream's primer3_config/ directory
        if param_dir is None:
            _warn_embedded_tables()
        d = None if param_dir is None else os.fsencode(param_dir)
        rc = lib.po_init(int(device), d, C.byref(h))
        if rc:
            raise RuntimeError("po_init failed (%d): %s" % (rc, lib.po_last_error(None).decode()))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            load().po_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, load().po_last_error(self._h).decode()))

    def set_conditions(self, mv=50.0, dv=0.0, dntp=0.8, dna_conc=50.0, temp_c=37.0, max_loop=30,
                       max_nn_length=60):
        self._check(load().po_set_conditions(self._h, mv, dv, dntp, dna_conc, temp_c, max_loop, max_nn_length),
                    "po_set_conditions")

    # ---- host-buffer calls ---------------------------------------------------
    def tm_batch(self, oligos, with_gc=False):
        oligos = np.ascontiguousarray(oligos, dtype=OLIGO_DTYPE)
        n = len(oligos)
        tm = np.empty(n, dtype=np.float64)
        gc = np.empty(n, dtype=np.float64) if with_gc else None
        self._check(load().po_tm_batch(self._h, oligos.ctypes.data, n, tm.ctypes.data,
                                       gc.ctypes.data if with_gc else None), "po_tm_batch")
        return (tm, gc) if with_gc else tm

    def thal_batch(self, type_, a, b=None, out=None):
        a = np.ascontiguousarray(a, dtype=OLIGO_DTYPE)
        n = len(a)
        if b is not None:
            b = np.ascontiguousarray(b, dtype=OLIGO_DTYPE)
            if len(b) != n:
                raise ValueError("a and b differ in length")
        if out is None:
            out = np.empty(n, dtype=THAL_OUT_DTYPE)
        self._check(load().po_thal_batch(self._h, type_, a.ctypes.data, b.ctypes.data if b is not None else None,
                                         n, out.ctypes.data), "po_thal_batch")
        return out

    def oligo_props_batch(self, oligos, valid=None):
        oligos = np.ascontiguousarray(oligos, dtype=OLIGO_DTYPE)
        n = len(oligos)
        out = np.empty(n, dtype=PROPS_DTYPE)
        self._check(load().po_oligo_props_batch(self._h, oligos.ctypes.data, n,
                                                C.byref(valid) if valid is not None else None,
                                                out.ctypes.data), "po_oligo_props_batch")
        return out

    def score_pairs(self, a, b, out=None):
        """Heterodimer(a,b) + hairpin(a) + hairpin(b) + Tm(a) + Tm(b) for host arrays.

        ``out`` may carry preallocated (e.g. pinned) arrays under the keys
        het / hairpin_a / hairpin_b / tm_a / tm_b."""
        a = np.ascontiguousarray(a, dtype=OLIGO_DTYPE)
        b = np.ascontiguousarray(b, dtype=OLIGO_DTYPE)
        n = len(a)
        if len(b) != n:
            raise ValueError("a and b differ in length")
        if out is None:
            out = {"het": np.empty(n, dtype=THAL_OUT_DTYPE), "hairpin_a": np.empty(n, dtype=THAL_OUT_DTYPE),
                   "hairpin_b": np.empty(n, dtype=THAL_OUT_DTYPE), "tm_a": np.empty(n, dtype=np.float64),
                   "tm_b": np.empty(n, dtype=np.float64)}
        self._check(load().po_score_pairs_batch(self._h, a.ctypes.data, b.ctypes.data, n,
                                                out["het"].ctypes.data, out["hairpin_a"].ctypes.data,
                                                out["hairpin_b"].ctypes.data, out["tm_a"].ctypes.data,
                                                out["tm_b"].ctypes.data), "po_score_pairs_batch")
        return out

    def score_pairs_dev(self, d_a, d_b, n, d_het, d_hp_a, d_hp_b, d_tm_a, d_tm_b, stream=None):
        self._check(load().po_score_pairs_batch_dev(self._h, d_a, d_b, n, d_het, d_hp_a, d_hp_b, d_tm_a, d_tm_b,
                                                    stream), "po_score_pairs_batch_dev")

    def count_relaxations(self, type_, a, b=None):
        a = np.ascontiguousarray(a, dtype=OLIGO_DTYPE)
        if b is not None:
            b = np.ascontiguousarray(b, dtype=OLIGO_DTYPE)
        v = C.c_int64()
        self._check(load().po_thal_count_relaxations(self._h, type_, a.ctypes.data,
                                                     b.ctypes.data if b is not None else None, len(a),
                                                     C.byref(v)), "po_thal_count_relaxations")
        return v.value

    # ---- primer picking (po_p3_*) -------------------------------------------------
    @staticmethod
    def _p3_inputs(templates, masks):
        blob, off = _seq_blob(templates)
        if masks is None:
            m = None
        elif isinstance(masks, np.ndarray) and masks.dtype == np.uint8 and masks.size == blob.size:
            m = np.ascontiguousarray(masks).reshape(-1)     # already concatenated
        else:
            m = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8) for x in masks])
                                     if len(masks) else np.zeros(0, np.uint8))
            if m.size != blob.size:
                raise ValueError("masks and templates differ in length")
        return np.ascontiguousarray(blob), off, m

    def p3_candidates(self, settings, templates, masks=None):
        """Sorted candidate lists of every template: (list_off[2*n+1], records)."""
        blob, off, m = self._p3_inputs(templates, masks)
        n = len(off) - 1
        list_off = np.zeros(2 * n + 1, dtype=np.int64)
        args = (self._h, C.byref(settings), blob.ctypes.data if blob.size else None,
                m.ctypes.data if m is not None and m.size else None, off.ctypes.data, n, list_off.ctypes.data)
        self._check(load().po_p3_candidates_batch(*args, None, 0), "po_p3_candidates_batch")
        out = np.zeros(int(list_off[-1]), dtype=P3_OLIGO_DTYPE)
        if out.size:
            self._check(load().po_p3_candidates_batch(*args, out.ctypes.data, out.size), "po_p3_candidates_batch")
        return list_off, out

    def p3_design_fixed(self, settings, templates, tasks, masks=None, targets=None, valid=None):
        """designPrimers with one primer given: (fixed records, pairs[n_tasks, num_return], n_pairs).
        ``valid`` (ValidParams) also evaluates Primer.is_valid of the picked primers (P3_PRIMER_VALID)."""
        blob, off, m = self._p3_inputs(templates, masks)
        n = len(off) - 1
        tasks = np.ascontiguousarray(tasks, dtype=P3_TASK_DTYPE)
        nt = len(tasks)
        tg = None if targets is None else np.ascontiguousarray(targets, dtype=np.int32).reshape(n, 2)
        fixed = np.zeros(nt, dtype=P3_OLIGO_DTYPE)
        pairs = np.empty((nt, settings.num_return), dtype=P3_PAIR_DTYPE)   # rows beyond n_pairs stay unset
        n_pairs = np.zeros(nt, dtype=np.int32)
        self._check(load().po_p3_design_fixed_batch(
            self._h, C.byref(settings), blob.ctypes.data if blob.size else None,
            m.ctypes.data if m is not None and m.size else None, off.ctypes.data,
            tg.ctypes.data if tg is not None else None, n, tasks.ctypes.data, nt,
            C.byref(valid) if valid is not None else None, fixed.ctypes.data,
            pairs.ctypes.data, n_pairs.ctypes.data), "po_p3_design_fixed_batch")
        return fixed, pairs, n_pairs

    # ---- device-pointer calls (torch tensors / raw pointers) --------------------
    def thal_batch_dev(self, type_, d_a, d_b, n, d_out, stream=None):
        self._check(load().po_thal_batch_dev(self._h, type_, d_a, d_b, n, d_out, stream), "po_thal_batch_dev")

    def tm_batch_dev(self, d_oligos, n, d_tm, d_gc=None, stream=None):
        self._check(load().po_tm_batch_dev(self._h, d_oligos, n, d_tm, d_gc, stream), "po_tm_batch_dev")

    def sync(self, stream=None):
        self._check(load().po_sync(self._h, stream), "po_sync")

    def launch_count(self):
        return load().po_launch_count(self._h)

    def thal_item_count(self):
        return load().po_thal_item_count(self._h)

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(load().po_measure_fp64_peak(self._h, C.byref(v)), "po_measure_fp64_peak")
        return v.value


    # ---- KASP enumeration and pair ranking ---------------------------------------
    def kasp_enumerate(self, markers, min_size, max_size):
        """markers: KASP_MARKER_DTYPE array -> (ref, alt) oligo arrays of shape (n, slots)."""
        markers = np.ascontiguousarray(markers, dtype=KASP_MARKER_DTYPE)
        n = len(markers)
        slots = 2 * (max_size - min_size + 1)
        ref = np.zeros((n, slots), dtype=OLIGO_DTYPE)
        alt = np.zeros((n, slots), dtype=OLIGO_DTYPE)
        self._check(load().po_kasp_enumerate_batch(self._h, markers.ctypes.data, n, min_size, max_size,
                                                   ref.ctypes.data, alt.ctypes.data), "po_kasp_enumerate_batch")
        return ref, alt

    def rank_pairs(self, assay, f, r, seg, n_offtargets, max_aaf, max_indel, tm_delta, top_n, a=None, direction=None,
                   hrm_delta=None, reporters=None):
        """Heterodimer predicate + goodness + per-marker top-n for a batch of markers.

        f, r (, a): OLIGO_DTYPE arrays, one entry per candidate pair; seg: n_markers+1 offsets."""
        f = np.ascontiguousarray(f, dtype=OLIGO_DTYPE)
        r = np.ascontiguousarray(r, dtype=OLIGO_DTYPE)
        n = len(f)
        seg = np.ascontiguousarray(seg, dtype=np.int64)
        nm = len(seg) - 1
        keep = [f, r, seg]
        arg = RankIn()
        arg.assay, arg.top_n, arg.n_pairs, arg.n_markers = assay, top_n, n, nm
        arg.f, arg.r, arg.seg = f.ctypes.data, r.ctypes.data, seg.ctypes.data

        def put(name, arr, dtype, shape=None):
            if arr is None:
                return None
            arr = np.ascontiguousarray(arr, dtype=dtype)
            if shape is not None:
                arr = arr.reshape(shape)
            keep.append(arr)
            setattr(arg, name, arr.ctypes.data)
            return arr
        put("a", a, OLIGO_DTYPE)
        put("dir", direction, np.uint8)
        put("n_offtargets", n_offtargets, np.int32)
        put("max_aaf", max_aaf, np.float64, (n, 3))
        put("max_indel", max_indel, np.int32)
        put("hrm_delta", hrm_delta, np.float64)
        if reporters is not None:
            rep = pack(list(reporters))
            for k in range(8):
                arg.reporters[k] = int(rep["w"].reshape(-1)[k])
        arg.tm_delta = float(tm_delta)
        res = {"goodness": np.zeros(n, dtype=np.int32), "qbits": np.zeros(n, dtype=np.int32),
               "tm": np.zeros((n, 3)), "het_tm": np.zeros((n, 2)),
               "sel": np.full((nm, top_n), -1, dtype=np.int64), "n_sel": np.zeros(nm, dtype=np.int32)}
        out = RankOut(*[res[k].ctypes.data for k in ("goodness", "qbits", "tm", "het_tm", "sel", "n_sel")])
        self._check(load().po_rank_pairs_batch(self._h, C.byref(arg), C.byref(out)), "po_rank_pairs_batch")
        return res

    def kasp_design(self, settings, params, markers, starts, maps_bits, mut_pos=None, mut_aaf=None, mut_indel=None,
                    mut_seg=None):
        """po_kasp_design_batch: (records[n, top_n] KASP_PAIR_DTYPE, n_sel[n], n_merged[n])."""
        markers = np.ascontiguousarray(markers, dtype=KASP_MARKER_DTYPE)
        n = len(markers)
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        maps_bits = np.ascontiguousarray(maps_bits, dtype=np.uint32).reshape(n, 3, 16)
        out = np.zeros((n, params.top_n), dtype=KASP_PAIR_DTYPE)
        n_out = np.zeros(n, dtype=np.int32)
        n_merged = np.zeros(n, dtype=np.int32)
        keep = []

        def ptr(a, dtype):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data
        self._check(load().po_kasp_design_batch(
            self._h, C.byref(settings), C.byref(params), markers.ctypes.data, starts.ctypes.data, n,
            maps_bits.ctypes.data, ptr(mut_pos, np.int64), ptr(mut_aaf, np.float64), ptr(mut_indel, np.int32),
            ptr(mut_seg, np.int64), out.ctypes.data, n_out.ctypes.data, n_merged.ctypes.data), "po_kasp_design_batch")
        return out, n_out, n_merged

    def kasp_masks(self, starts, tmpl_len, maps_bits, mut_pos=None, mut_seg=None, kasp_window=True, max_size=25):
        """po_kasp_masks_batch: excluded-base bit maps excl[n_types, n, 16] of every search type."""
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        n = len(starts)
        maps_bits = np.ascontiguousarray(maps_bits, dtype=np.uint32).reshape(n, 3, 16)
        with_mut = mut_seg is not None
        excl = np.zeros((6 if with_mut else 3, n, 16), dtype=np.uint32)
        mp = np.ascontiguousarray(mut_pos, dtype=np.int64) if with_mut else None
        ms = np.ascontiguousarray(mut_seg, dtype=np.int64) if with_mut else None
        self._check(load().po_kasp_masks_batch(self._h, starts.ctypes.data, n, int(tmpl_len), maps_bits.ctypes.data,
                                               mp.ctypes.data if with_mut else None, ms.ctypes.data if with_mut else None,
                                               1 if with_mut else 0, 1 if kasp_window else 0, int(max_size),
                                               excl.ctypes.data), "po_kasp_masks_batch")
        return excl

    # ---- "next" rows: HRM product Tm, mutation annotation ---------------------------
    def tm_nn_batch(self, seqs):
        """Biopython Tm_NN (defaults) of each sequence; NaN where Biopython would raise."""
        blob, off = _seq_blob(seqs)
        n = len(off) - 1
        out = np.empty(n, dtype=np.float64)
        self._check(load().po_tm_nn_batch(self._h, blob.tobytes() if blob.size else b"", off.ctypes.data, n,
                                          out.ctypes.data), "po_tm_nn_batch")
        return out

    def annotate_mutations(self, mut_pos, mut_aaf, mut_indel, mut_seg, pair_seg, primer_start, primer_stop,
                           n_primers):
        mut_pos = np.ascontiguousarray(mut_pos, dtype=np.int64)
        mut_aaf = np.ascontiguousarray(mut_aaf, dtype=np.float64)
        mut_indel = np.ascontiguousarray(mut_indel, dtype=np.int32)
        mut_seg = np.ascontiguousarray(mut_seg, dtype=np.int64)
        pair_seg = np.ascontiguousarray(pair_seg, dtype=np.int64)
        n_pairs = int(pair_seg[-1]) if len(pair_seg) else 0
        ps = np.ascontiguousarray(primer_start, dtype=np.int64).reshape(n_pairs, 3)
        pe = np.ascontiguousarray(primer_stop, dtype=np.int64).reshape(n_pairs, 3)
        aaf = np.zeros((n_pairs, 3))
        indel = np.zeros(n_pairs, dtype=np.int32)
        self._check(load().po_annotate_mutations_batch(
            self._h, mut_pos.ctypes.data, mut_aaf.ctypes.data, mut_indel.ctypes.data, mut_seg.ctypes.data,
            pair_seg.ctypes.data, len(mut_seg) - 1, ps.ctypes.data, pe.ctypes.data, n_primers, aaf.ctypes.data,
            indel.ctypes.data), "po_annotate_mutations_batch")
        return aaf, indel
