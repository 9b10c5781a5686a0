"""Host-side mirror of the hot-path parts of reference src/polyoligo/_lib_kasp.py:
allele-specific marker-primer enumeration (get_marker_primers, :295-359) and
the reporter-tailed heterodimer check (:75-117), batched over markers.

``get_marker_primers_batch`` enumerates every marker's 2 x (MAX-MIN+1)
candidates with the CUDA enumeration kernel, scores all of them with one
oligo-properties pass and keeps, per marker and in the reference's order, the
slots whose REF and ALT primers are both valid.
"""
import numpy as np

from . import _native, lib_primer3
from .thermoanalysis import get_context

_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    _CODE[_c] = _v


def _pack_bits(seq, n_words):
    """ASCII -> (2-bit words, N mask words) as uint32 arrays."""
    raw = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    codes = _CODE[raw]
    isn = codes == 255
    codes = np.where(isn, 0, codes).astype(np.uint64)
    idx = np.arange(len(raw))
    words = np.zeros(n_words, dtype=np.uint64)
    np.add.at(words, idx >> 4, codes << (2 * (idx & 15)).astype(np.uint64))
    nmask = np.zeros((n_words + 1) // 2, dtype=np.uint64)
    np.add.at(nmask, idx >> 5, isn.astype(np.uint64) << (idx & 31).astype(np.uint64))
    return words.astype(np.uint32), nmask.astype(np.uint32)


def marker_record(hroi):
    """One po_kasp_marker from a homolog-mapped ROI (attributes seq, start, stop,
    marker.pos1 / ref / alt as in reference lib_markers.ROI / Marker)."""
    rec = np.zeros((), dtype=_native.KASP_MARKER_DTYPE)
    seq = hroi.seq
    if len(seq) > 512 or len(hroi.marker.alt) > 32 or len(hroi.marker.ref) > 32:
        raise ValueError("template longer than 512 nt or allele longer than 32 nt")
    rec["tmpl"], rec["nmask"] = _pack_bits(seq, 32)
    alt_w, alt_n = _pack_bits(hroi.marker.alt, 2)
    rec["alt"], rec["alt_nmask"] = alt_w, alt_n[0]
    rec["tmpl_len"] = len(seq)
    rec["marker_pos"] = hroi.marker.pos1 - hroi.start           # _lib_kasp.py:297
    rec["marker_pos_rc"] = hroi.stop - hroi.marker.pos1 + 1     # _lib_kasp.py:298
    rec["ref_len"] = len(hroi.marker.ref)
    rec["alt_len"] = len(hroi.marker.alt)
    return rec


class PrimerPair(lib_primer3.PrimerPair):
    def __init__(self):
        super().__init__()
        self.snp_id = None
        self.primers = {"F": lib_primer3.Primer(), "R": lib_primer3.Primer(), "A": lib_primer3.Primer()}
        self.ref_dimer = None
        self.alt_dimer = None

    def add_reporter_dyes(self, reporters):
        """reference _lib_kasp.py:75-86"""
        marker, common = (("F", "R") if self.dir == "F" else ("R", "F"))
        self.primers[marker].reporter = reporters[0]
        self.primers["A"].reporter = reporters[1]
        self.primers[common].reporter = ""

    def check_heterodimerization(self):
        check_heterodimerization_batch([self])


def check_heterodimerization_batch(pps, device=None):
    """KASP PrimerPair.check_heterodimerization (reference _lib_kasp.py:89-117) for many
    pairs: reporter + primer sequences; tm_delta enters both sides of the comparison."""
    if not pps:
        return
    lib_primer3.calc_thermals_batch([p for pp in pps for p in pp.primers.values()], device)
    tailed = {d: [pp.primers[d].reporter + pp.primers[d].sequence for pp in pps] for d in "FRA"}
    first = [tailed["A"][k] if pp.dir == "F" else tailed["F"][k] for k, pp in enumerate(pps)]
    second = [tailed["R"][k] if pp.dir == "F" else tailed["A"][k] for k, pp in enumerate(pps)]
    ctx = get_context(device)
    ref = ctx.thal_batch(_native.THAL_ANY, _native.pack(tailed["F"]), _native.pack(tailed["R"]))
    alt = ctx.thal_batch(_native.THAL_ANY, _native.pack(first), _native.pack(second))
    for pp, r, a in zip(pps, ref, alt):
        tm_ref = float(r["tm"]) - pp.tm_delta
        tm_alt = float(a["tm"]) - pp.tm_delta
        lim_f, lim_r = pp.primers["F"].tm - pp.tm_delta, pp.primers["R"].tm - pp.tm_delta
        pp.ref_dimer = not ((tm_ref <= lim_f) and (tm_ref <= lim_r))
        pp.alt_dimer = not ((tm_alt <= lim_f) and (tm_alt <= lim_r))


class PCR(lib_primer3.PCR):
    def check_heterodimerization(self):
        check_heterodimerization_batch(self.pps)

    def add_reporter_dyes(self, reporters):
        for pp in self.pps:
            pp.add_reporter_dyes(reporters)

    def prune(self, n):
        return super().prune(n, assay_name="KASP")


def get_marker_primers_batch(hrois, device=None, return_all=False):
    """get_marker_primers (reference _lib_kasp.py:295-359) for many markers.

    Returns one list of PrimerPair per marker: the slots whose REF and ALT primers
    both pass Primer.is_valid, in the reference's (padding ascending, F then R) order.
    ``return_all`` also returns every enumerated (dir, ref, alt) triple per marker."""
    if not hrois:
        return ([], []) if return_all else []
    g = lib_primer3.PRIMER3_GLOBALS
    lo, hi = int(g["PRIMER_MIN_SIZE"]), int(g["PRIMER_MAX_SIZE"])
    ctx = get_context(device)
    markers = np.array([marker_record(h) for h in hrois], dtype=_native.KASP_MARKER_DTYPE)
    ref, alt = ctx.kasp_enumerate(markers, lo, hi)
    slots = ref.shape[1]
    v = lib_primer3.valid_params()
    pr = ctx.oligo_props_batch(ref.reshape(-1), v).reshape(ref.shape)
    pa = ctx.oligo_props_batch(alt.reshape(-1), v).reshape(alt.shape)
    keep = ((pr["flags"] & _native.PRIMER_VALID) != 0) & ((pa["flags"] & _native.PRIMER_VALID) != 0)
    ref_seq = np.array(_native.unpack(ref.reshape(-1)), dtype=object).reshape(ref.shape)
    alt_seq = np.array(_native.unpack(alt.reshape(-1)), dtype=object).reshape(alt.shape)
    out, everything = [], []
    for m, hroi in enumerate(hrois):
        pps = []
        for s in range(slots):
            if not keep[m, s]:
                continue
            direction = "F" if s % 2 == 0 else "R"
            pp = PrimerPair()
            pp.chrom, pp.snp_id, pp.dir = getattr(hroi, "chrom", None), getattr(hroi, "name", None), direction
            for key, seq, rec in ((direction, ref_seq[m, s], pr[m, s]), ("A", alt_seq[m, s], pa[m, s])):
                p = lib_primer3.Primer(sequence=seq)
                p.tm, p.hairpin_tm, p.homodimer_tm = float(rec["tm"]), float(rec["hairpin_tm"]), float(rec["homodimer_tm"])
                p.gc_content, p.length, p.nN = float(rec["gc_content"]), int(rec["length"]), 0
                pp.primers[key] = p
            pps.append(pp)
        out.append(pps)
        if return_all:
            everything.append([("F" if s % 2 == 0 else "R", ref_seq[m, s], alt_seq[m, s]) for s in range(slots)])
    return (out, everything) if return_all else out


def get_marker_primers(hroi, device=None):
    return get_marker_primers_batch([hroi], device)[0]
