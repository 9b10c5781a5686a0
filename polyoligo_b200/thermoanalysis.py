"""Drop-in for ``primer3.thermoanalysis.ThermoAnalysis`` as polyoligo uses it.

The reference constructs ``ThermoAnalysis()`` with defaults and calls
``calcTm`` / ``calcHairpin`` / ``calcHomodimer`` / ``calcHeterodimer``
(reference src/polyoligo/lib_primer3.py:58-61,157-158; _lib_kasp.py:99-105;
_lib_crispr.py:41-42), reading only ``.tm`` of the results.  This module offers
the same names backed by the CUDA library, plus ``*_batch`` variants that take
lists: the scalar calls are batches of one and exist for drop-in use, the
batched calls are what the pipelines in this package use.

There is no CPU fallback: without the CUDA library or a GPU the first call raises.
"""
import os
import threading


from . import _native

_lock = threading.Lock()
_contexts = {}


DEFAULT_CONDITIONS = (50.0, 0.0, 0.8, 50.0, 37.0, 30, 60)   # primer3-py 0.6.0 ThermoAnalysis() defaults


def get_context(device=None, conditions=None):
    """Process-wide scoring context, created lazily (after any fork: the reference
    forks one worker per marker, reference src/polyoligo/cli_kasp.py:354).

    One context per (process, device, conditions): the ``conditions=None`` context keeps the
    ``ThermoAnalysis()`` defaults for ever -- it is the one ``lib_primer3`` / ``kasp`` use, like the
    reference, which builds a fresh default ``ThermoAnalysis()`` per call -- and a non-default
    ``ThermoAnalysis(mv_conc=...)`` gets its own context instead of mutating the shared one."""
    if device is None:
        device = int(os.environ.get("POLYOLIGO_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if conditions is not None:
        conditions = tuple(conditions)
        if conditions == DEFAULT_CONDITIONS:
            conditions = None
    key = (os.getpid(), device, conditions)
    with _lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = _native.Context(device)
            if conditions is not None:
                ctx.set_conditions(*conditions)
            _contexts[key] = ctx
        return ctx


class ThermoResult:
    """Mirror of primer3-py's ThermoResult: tm (deg C), dg and dh (cal/mol), ds (cal/K/mol)."""
    __slots__ = ("structure_found", "tm", "dg", "dh", "ds", "align_end_1", "align_end_2")

    def __init__(self, rec):
        self.structure_found = bool(rec["flags"] & _native.ITEM_STRUCTURE_FOUND)
        self.tm = float(rec["tm"])
        self.dg = float(rec["dg"])
        self.dh = float(rec["dh"])
        self.ds = float(rec["ds"])
        self.align_end_1 = int(rec["align_end_1"])
        self.align_end_2 = int(rec["align_end_2"])

    def __repr__(self):
        return "ThermoResult(structure_found={}, tm={:0.2f}, dg={:0.2f}, dh={:0.2f}, ds={:0.2f})".format(
            self.structure_found, self.tm, self.dg, self.dh, self.ds)


class ThermoAnalysis:
    """Same constructor keywords and defaults as primer3-py 0.6.0's ThermoAnalysis."""

    def __init__(self, mv_conc=50.0, dv_conc=0.0, dntp_conc=0.8, dna_conc=50.0, temp_c=37.0, max_loop=30,
                 max_nn_length=60, tm_method="santalucia", salt_correction_method="santalucia", device=None):
        if tm_method not in ("santalucia", 1) or salt_correction_method not in ("santalucia", 1):
            raise ValueError("only the santalucia Tm method and salt correction are implemented "
                             "(the ones polyoligo uses)")
        self.mv_conc, self.dv_conc, self.dntp_conc, self.dna_conc = mv_conc, dv_conc, dntp_conc, dna_conc
        self.temp_c, self.max_loop, self.max_nn_length = temp_c, max_loop, max_nn_length
        self._device = device

    def _ctx(self):
        return get_context(self._device, (float(self.mv_conc), float(self.dv_conc), float(self.dntp_conc),
                                          float(self.dna_conc), float(self.temp_c), int(self.max_loop),
                                          int(self.max_nn_length)))

    @staticmethod
    def _check(seqs):
        for s in seqs:
            if len(s) > _native.MAX_OLIGO_LEN:
                # primer3-py raises for over-long thal inputs as well
                raise RuntimeError("sequence longer than %d nt for thermodynamic alignment" % _native.MAX_OLIGO_LEN)

    # ---- batched -------------------------------------------------------------
    def calcTm_batch(self, seqs):
        return self._ctx().tm_batch(_native.pack(seqs))

    def calcHairpin_batch(self, seqs):
        self._check(seqs)
        return self._ctx().thal_batch(_native.THAL_HAIRPIN, _native.pack(seqs))

    def calcHomodimer_batch(self, seqs):
        self._check(seqs)
        return self._ctx().thal_batch(_native.THAL_ANY, _native.pack(seqs))

    def calcHeterodimer_batch(self, seqs1, seqs2):
        self._check(seqs1)
        self._check(seqs2)
        return self._ctx().thal_batch(_native.THAL_ANY, _native.pack(seqs1), _native.pack(seqs2))

    # ---- scalar, as the reference calls them ------------------------------------
    def calcTm(self, seq):
        return float(self.calcTm_batch([seq])[0])

    def calcHairpin(self, seq):
        return ThermoResult(self.calcHairpin_batch([seq])[0])

    def calcHomodimer(self, seq):
        return ThermoResult(self.calcHomodimer_batch([seq])[0])

    def calcHeterodimer(self, seq1, seq2):
        return ThermoResult(self.calcHeterodimer_batch([seq1], [seq2])[0])


# module-level helpers in the style of primer3.bindings
def calcTm(seq, **kw):
    return ThermoAnalysis(**kw).calcTm(seq)


def calcHairpin(seq, **kw):
    return ThermoAnalysis(**kw).calcHairpin(seq)


def calcHomodimer(seq, **kw):
    return ThermoAnalysis(**kw).calcHomodimer(seq)


def calcHeterodimer(seq1, seq2, **kw):
    return ThermoAnalysis(**kw).calcHeterodimer(seq1, seq2)


def gc_content(seq):
    """Bio.SeqUtils.GC as used at reference src/polyoligo/lib_primer3.py:63."""
    n = sum(seq.count(x) for x in "GCgcSs")
    return n * 100.0 / len(seq) if seq else 0.0


__all__ = ["ThermoAnalysis", "ThermoResult", "calcTm", "calcHairpin", "calcHomodimer", "calcHeterodimer",
           "get_context", "gc_content"]
