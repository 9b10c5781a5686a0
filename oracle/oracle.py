"""ctypes front end of the CPU ORACLE (test infrastructure, NOT product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
reference arm may import this module.  See ``oracle/po_oracle.h`` for scope,
provenance and parity status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpo_oracle.so")


class Args(C.Structure):
    _fields_ = [("mv", C.c_double), ("dv", C.c_double), ("dntp", C.c_double),
                ("dna_conc", C.c_double), ("temp_k", C.c_double),
                ("max_loop", C.c_int32), ("type", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("tm", C.c_double), ("dg", C.c_double), ("dh", C.c_double), ("ds", C.c_double),
                ("structure_found", C.c_int32), ("align_end_1", C.c_int32),
                ("align_end_2", C.c_int32), ("status", C.c_int32), ("relaxations", C.c_int64)]


RESULT_DTYPE = np.dtype([("tm", "f8"), ("dg", "f8"), ("dh", "f8"), ("ds", "f8"),
                         ("structure_found", "i4"), ("align_end_1", "i4"),
                         ("align_end_2", "i4"), ("status", "i4"), ("relaxations", "i8")])

ANY, END1, END2, HAIRPIN = 1, 2, 3, 4


def build(force=False):
    if force or not os.path.exists(_LIB) or \
            os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "po_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.po_oracle_params_load.restype = C.c_void_p
        L.po_oracle_params_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.po_oracle_params_free.argtypes = [C.c_void_p]
        L.po_oracle_params_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.po_oracle_default_args.argtypes = [C.POINTER(Args), C.c_int]
        L.po_oracle_seqtm.restype = C.c_double
        L.po_oracle_seqtm.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.po_oracle_thal.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(Args), C.POINTER(Result)]
        L.po_oracle_tm_batch.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int]
        L.po_oracle_thal_batch.argtypes = [C.c_void_p, C.POINTER(Args), C.c_char_p, C.c_void_p,
                                           C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.po_oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _blob(seqs):
    """list[str] or (uint8 blob, int64 offsets) -> (bytes, offsets array)."""
    if isinstance(seqs, tuple):
        blob, off = seqs
        return np.ascontiguousarray(blob, dtype=np.uint8).tobytes(), np.ascontiguousarray(off, dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return "".join(seqs).encode("ascii"), off


class Oracle:
    """CPU oracle with one parameter set (``param_dir=None`` = embedded defaults)."""

    def __init__(self, param_dir=None):
        err = C.create_string_buffer(512)
        if param_dir is None:      # the same default as the product's Context: upstream's tables when named
            param_dir = os.environ.get("POLYOLIGO_P3_CONFIG") or None
        d = None if param_dir is None else os.fsencode(param_dir)
        self._p = lib().po_oracle_params_load(d, err, 512)
        if not self._p:
            raise RuntimeError("oracle parameter load failed: " + err.value.decode())

    def __del__(self):
        if getattr(self, "_p", None):
            lib().po_oracle_params_free(self._p)
            self._p = None

    @staticmethod
    def args(type_=ANY, **kw):
        a = Args()
        lib().po_oracle_default_args(C.byref(a), type_)
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    def table(self, idx):
        n = {0: 625, 1: 625, 2: 625, 3: 625, 4: 125, 5: 125, 6: 30, 7: 30, 8: 30}[idx]
        dh, ds = np.empty(n), np.empty(n)
        lib().po_oracle_params_get(self._p, idx, dh.ctypes.data, ds.ctypes.data)
        return dh, ds

    # -- scalar calls, named after the primer3-py calls the reference makes ----
    @staticmethod
    def calcTm(seq, mv=50.0, dv=0.0, dntp=0.8, dna_conc=50.0, max_nn_length=60):
        return lib().po_oracle_seqtm(seq.encode("ascii"), dna_conc, mv, dv, dntp, max_nn_length)

    def thal(self, s1, s2, type_=ANY, **kw):
        a = self.args(type_, **kw)
        r = Result()
        lib().po_oracle_thal(self._p, s1.encode("ascii"), s2.encode("ascii"), C.byref(a), C.byref(r))
        return r

    def calcHairpin(self, seq, **kw):
        return self.thal(seq, seq, HAIRPIN, **kw)

    def calcHomodimer(self, seq, **kw):
        return self.thal(seq, seq, ANY, **kw)

    def calcHeterodimer(self, s1, s2, **kw):
        return self.thal(s1, s2, ANY, **kw)

    # -- batch calls -----------------------------------------------------------
    @staticmethod
    def tm_batch(seqs, nthreads=1, mv=50.0, dv=0.0, dntp=0.8, dna_conc=50.0, max_nn_length=60):
        blob, off = _blob(seqs)
        n = len(off) - 1
        out = np.empty(n, dtype=np.float64)
        lib().po_oracle_tm_batch(blob, off.ctypes.data, n, dna_conc, mv, dv, dntp, max_nn_length,
                                 out.ctypes.data, nthreads)
        return out

    def thal_batch(self, seqs1, seqs2=None, type_=ANY, nthreads=1, **kw):
        a = self.args(type_, **kw)
        b1, o1 = _blob(seqs1)
        n = len(o1) - 1
        out = np.zeros(n, dtype=RESULT_DTYPE)
        if seqs2 is None:
            lib().po_oracle_thal_batch(self._p, C.byref(a), b1, o1.ctypes.data, None, None, n,
                                       out.ctypes.data, nthreads)
        else:
            b2, o2 = _blob(seqs2)
            assert len(o2) - 1 == n
            lib().po_oracle_thal_batch(self._p, C.byref(a), b1, o1.ctypes.data, b2, o2.ctypes.data, n,
                                       out.ctypes.data, nthreads)
        return out


def max_threads():
    return lib().po_oracle_max_threads()
