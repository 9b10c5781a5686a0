"""Host-side mirror of the hot-path parts of reference src/polyoligo/lib_primer3.py.

Same names, attributes and decisions as the reference's ``Primer`` /
``PrimerPair`` / ``PCR`` (lib_primer3.py:24-97, 111-214, 217-463) and its mask
helpers (:564-600), but every thermodynamic quantity comes from BATCHED calls
into the CUDA library: ``calc_thermals_batch`` evaluates any number of primers
in one go, and the per-object methods are thin wrappers over batches of one so
existing callers keep working.  ``get_primers`` (:466-529) goes through
``primer3_bindings`` (the GPU primer picker).  BLAST off-target search
(check_offtargeting) stays the reference's external glue and is not here.
"""
from copy import deepcopy

import numpy as np

from . import _native, lib_utils, primer3_bindings, scoring
from .thermoanalysis import get_context, gc_content

PRIMER3_GLOBALS = {}     # the PRIMER_* thresholds of the assay YAML (reference data/PRIMER3_*.yaml)
TM_DELTA = 5             # reference lib_primer3.py:21
MIN_OFFTARGET_SIZE = 0   # reference lib_primer3.py:17-18 (set_offtarget_size)
MAX_OFFTARGET_SIZE = 1000


def set_globals(**kwargs):
    """reference lib_primer3.py:539-544: remembers the tags and hands them to the primer picker."""
    PRIMER3_GLOBALS.update(kwargs)
    primer3_bindings.setP3Globals(PRIMER3_GLOBALS)


def set_offtarget_size(min_size, max_size):
    global MIN_OFFTARGET_SIZE, MAX_OFFTARGET_SIZE
    MIN_OFFTARGET_SIZE, MAX_OFFTARGET_SIZE = min_size, max_size


def list_offtargets(pp_hits):
    """Off-target PCR products of one primer pair from its BLAST hits (reference
    lib_primer3.py:392-419; BLAST itself stays external): pp_hits[d][chrom] is a list of hits
    with 'sstart'.  Per chromosome the distinct forward starts i and reverse starts j are sorted;
    for each i the stops j > i are walked in order and reported while MIN < j - i < MAX -- the
    walk STOPS at the first j > i outside that range, also when it is still too close (the
    reference's `else: break`), which is kept."""
    out = []
    for chrom in pp_hits["F"]:
        if chrom not in pp_hits["R"]:
            continue
        starts = np.unique([h["sstart"] for h in pp_hits["F"][chrom]])
        stops = np.unique([h["sstart"] for h in pp_hits["R"][chrom]])
        for i in starts:
            for j in stops[np.searchsorted(stops, i, side="right"):]:
                delta = j - i
                if MIN_OFFTARGET_SIZE < delta < MAX_OFFTARGET_SIZE:
                    out.append("{}:{}-{}".format(chrom, i, j))
                else:
                    break
    return out


def set_tm_delta(v):
    global TM_DELTA
    TM_DELTA = v


def valid_params(tm_delta=None):
    g = PRIMER3_GLOBALS
    return _native.ValidParams(int(g["PRIMER_MIN_SIZE"]), int(g["PRIMER_MAX_SIZE"]), float(g["PRIMER_MIN_GC"]),
                               float(g["PRIMER_MAX_GC"]), float(g["PRIMER_MIN_TM"]), float(g["PRIMER_MAX_TM"]),
                               float(TM_DELTA if tm_delta is None else tm_delta))


class Primer:
    def __init__(self, **kwargs):
        self.id = None
        self.start = None
        self.stop = None
        self.length = None
        self.direction = None
        self.hairpin_tm = None
        self.homodimer_tm = None
        self.gc_content = None
        self.has_thermals = False       # never set True, as in the reference (lib_primer3.py:36)
        self.tm_delta = TM_DELTA
        self.tm = None
        self.mutations = []
        self.max_aaf = 0
        self.nN = 0
        self.sequence = None
        self.sequence_ambiguous = None
        self.reporter = ""
        # attributes shared with the primer picker (reference lib_primer3.py:45-51)
        self.end_stability = None
        self.hairpin_th = None
        self.penalty = None
        self.self_any_th = None
        self.self_end_th = None
        for k, v in kwargs.items():
            if k in self.__dict__:
                setattr(self, k, deepcopy(v))

    def calc_thermals(self):
        calc_thermals_batch([self])

    def is_valid(self):
        return bool(is_valid_batch([self])[0])

    def add_primer3_attributes(self, p3):
        """reference lib_primer3.py:99-108"""
        for k in ("id", "start", "stop", "sequence", "end_stability", "hairpin_th", "penalty", "self_any_th",
                  "self_end_th"):
            setattr(self, k, getattr(p3, k))


def calc_thermals_batch(primers, device=None):
    """Primer.calc_thermals (reference lib_primer3.py:56-64) for many primers at once.

    Primers containing N (or longer than 60 nt) are not evaluated on the device;
    the reference rejects them regardless of their thermodynamics (:72-74), so
    their tm / hairpin_tm / homodimer_tm are reported as 0.0."""
    if not primers:
        return np.zeros(0, dtype=_native.PROPS_DTYPE)
    seqs = [p.sequence for p in primers]
    props = get_context(device).oligo_props_batch(_native.pack(seqs))
    skipped = (props["flags"] & _native.ITEM_SKIPPED) != 0
    for p, s, rec, skip in zip(primers, seqs, props, skipped):
        p.length = len(s)
        p.gc_content = gc_content(s)
        p.nN = s.count("N")
        p.tm = 0.0 if skip else float(rec["tm"])
        p.hairpin_tm = 0.0 if skip else float(rec["hairpin_tm"])
        p.homodimer_tm = 0.0 if skip else float(rec["homodimer_tm"])
    return props


def is_valid_batch(primers, device=None):
    """Primer.is_valid (reference lib_primer3.py:66-97) for many primers: bool array."""
    calc_thermals_batch(primers, device)
    g = PRIMER3_GLOBALS
    ok = np.ones(len(primers), dtype=bool)
    for k, p in enumerate(primers):
        if p.nN > 0:
            ok[k] = False
        if p.length < g["PRIMER_MIN_SIZE"] or p.length > g["PRIMER_MAX_SIZE"]:
            ok[k] = False
        if p.gc_content < g["PRIMER_MIN_GC"] or p.gc_content > g["PRIMER_MAX_GC"]:
            ok[k] = False
        if p.tm < g["PRIMER_MIN_TM"] or p.tm > g["PRIMER_MAX_TM"]:
            ok[k] = False
        if p.hairpin_tm >= (p.tm - p.tm_delta):
            ok[k] = False
        if p.homodimer_tm >= (p.tm - p.tm_delta):
            ok[k] = False
    return ok


class PrimerPair:
    def __init__(self):
        self.tm_delta = TM_DELTA
        self.id = None
        self.primers = {"F": Primer(), "R": Primer()}
        self.offtargets = []
        self.max_indel_size = 0
        self.chrom = None
        self.dimer = None
        self.goodness = None
        self.qcode = None
        self.hrm = None
        self.compl_any_th = None
        self.compl_end_th = None
        self.product_size = None
        self.penalty = None
        self.dir = None

    def add_primer3_attributes(self, pp3):
        """reference lib_primer3.py:139-145"""
        for d in ("F", "R"):
            self.primers[d].add_primer3_attributes(pp3.primers[d])
        self.compl_any_th = pp3.compl_any_th
        self.compl_end_th = pp3.compl_end_th
        self.product_size = pp3.product_size
        self.penalty = pp3.penalty

    def check_heterodimerization(self):
        check_heterodimerization_batch([self])

    def add_mutations(self, mutations):
        """reference lib_primer3.py:166-210, for one pair on the host (many pairs: add_mutations_batch
        for the numbers, add_mutation_strings for the text)."""
        add_mutation_strings(self, mutations)
        for d, p in self.primers.items():
            for m in mutations:
                if p.start <= m.pos <= p.stop:
                    p.max_aaf = max(p.max_aaf, m.aaf)
        for m in mutations:
            if self.primers["F"].start <= m.pos <= self.primers["R"].stop:
                self.max_indel_size = max(self.max_indel_size, m.indel_size)

    def score(self, assay_name):
        self.goodness, self.qcode = scoring.get_score_fnc(assay_name)(self)


def check_heterodimerization_batch(pps, device=None):
    """PrimerPair.check_heterodimerization (reference lib_primer3.py:147-164) for many pairs."""
    if not pps:
        return
    calc_thermals_batch([p for pp in pps for p in pp.primers.values()], device)
    f = _native.pack([pp.primers["F"].sequence for pp in pps])
    r = _native.pack([pp.primers["R"].sequence for pp in pps])
    het = get_context(device).thal_batch(_native.THAL_ANY, f, r)
    for pp, rec in zip(pps, het):
        tm_ref = float(rec["tm"])
        pp.dimer = not ((tm_ref <= (pp.primers["F"].tm - pp.tm_delta)) and
                        (tm_ref <= (pp.primers["R"].tm - pp.tm_delta)))


def _rank_inputs(pps, keys):
    n = len(pps)
    aaf = np.zeros((n, 3))
    for k, pp in enumerate(pps):
        for c, d in enumerate(keys):
            aaf[k, c] = pp.primers[d].max_aaf
    return (np.array([len(pp.offtargets) for pp in pps], dtype=np.int32), aaf,
            np.array([pp.max_indel_size for pp in pps], dtype=np.int32))


class PCR:
    def __init__(self, chrom=None, primer_pairs=None, name=None, snp_id=None, pos=None, ref=None, alt=None):
        self.chrom, self.name, self.snp_id, self.pos, self.ref, self.alt = chrom, name, snp_id, pos, ref, alt
        self.ref_n = len(ref) if ref is not None else 0       # reference lib_primer3.py:229-239
        self.alt_n = len(alt) if alt is not None else 0
        if self.ref == "":
            self.ref = "*"
        if self.alt == "":
            self.alt = "*"
        self.pps = primer_pairs if primer_pairs is not None else []
        self.pps_classified = {}
        self.pps_pruned = {}

    def check_heterodimerization(self):
        check_heterodimerization_batch(self.pps)

    def classify(self, assay_name):
        for pp in self.pps:
            pp.score(assay_name)
            self.pps_classified.setdefault(pp.goodness, []).append(pp)

    def prune(self, n, assay_name="generic"):
        """Best n pairs: descending goodness, insertion order inside a class (HRM:
        larger delta first).  reference lib_primer3.py:433-463."""
        nprim = 0
        for score in sorted(self.pps_classified, reverse=True):
            group = self.pps_classified[score]
            if assay_name == "HRM":
                order = np.argsort(-np.array([pp.hrm["delta"] for pp in group]), kind="stable")
                group = [group[i] for i in order]
            self.pps_pruned[score] = group[:max(0, n - nprim)]
            nprim += len(self.pps_pruned[score])
        return nprim

    def rank(self, n, assay_name, device=None):
        """check_heterodimerization + classify + prune in one device pass
        (po_rank_pairs_batch); fills the same attributes the three calls would."""
        rank_batch([self], n, assay_name, device)
        return sum(len(v) for v in self.pps_pruned.values())


def rank_batch(pcrs, n, assay_name, device=None, reporters=None):
    """Heterodimer check, goodness and top-n for MANY markers/regions in one call."""
    kasp = assay_name == "KASP"
    assay = {"KASP": _native.ASSAY_KASP, "HRM": _native.ASSAY_HRM}.get(assay_name, _native.ASSAY_PCR)
    keys = ["F", "R", "A"] if kasp else ["F", "R"]
    pps = [pp for pcr in pcrs for pp in pcr.pps]
    seg = np.zeros(len(pcrs) + 1, dtype=np.int64)
    np.cumsum([len(pcr.pps) for pcr in pcrs], out=seg[1:])
    if not pps:
        return
    n_off, aaf, indel = _rank_inputs(pps, keys)
    kw = {}
    if len({float(pp.tm_delta) for pp in pps}) != 1:
        raise ValueError("rank_batch: tm_delta differs between the pairs of one batch")
    if kasp:
        kw["a"] = _native.pack([pp.primers["A"].sequence for pp in pps])
        kw["direction"] = np.array([0 if pp.dir == "F" else 1 for pp in pps], dtype=np.uint8)
        # The tails are the ones add_reporter_dyes put on the primers (reference _lib_kasp.py:75-86,
        # 95-97: reporter + sequence): marker primer -> reporters[0], ALT -> reporters[1], common -> "".
        on_primers = {(pp.primers[pp.dir].reporter, pp.primers["A"].reporter) for pp in pps}
        if len(on_primers) != 1:
            raise ValueError("rank_batch: the pairs of one batch carry different reporter dyes")
        own = on_primers.pop()
        if reporters is None:
            reporters = own if any(own) else None
        elif any(own) and tuple(reporters) != own:
            raise ValueError("rank_batch: reporters argument %r conflicts with the primers' reporter dyes %r"
                             % (tuple(reporters), own))
        kw["reporters"] = reporters
    if assay_name == "HRM":
        kw["hrm_delta"] = np.array([pp.hrm["delta"] for pp in pps])
    res = get_context(device).rank_pairs(assay, _native.pack([pp.primers["F"].sequence for pp in pps]),
                                         _native.pack([pp.primers["R"].sequence for pp in pps]), seg, n_off, aaf,
                                         indel, pps[0].tm_delta, n, **kw)
    for k, pp in enumerate(pps):
        for c, d in enumerate(keys):
            pp.primers[d].tm = float(res["tm"][k, c])
        bits = int(res["qbits"][k])
        if kasp:
            pp.ref_dimer = bool(bits & _native.Q_REF_DIMER)
            pp.alt_dimer = bool(bits & _native.Q_ALT_DIMER)
        else:
            pp.dimer = bool(bits & _native.Q_BITS["d"])
        pp.goodness = int(res["goodness"][k])
        pp.qcode = scoring.qcode_from_bits(bits)
    for m, pcr in enumerate(pcrs):
        pcr.pps_classified, pcr.pps_pruned = {}, {}
        for pp in pcr.pps:
            pcr.pps_classified.setdefault(pp.goodness, []).append(pp)
        for idx in res["sel"][m, :res["n_sel"][m]]:
            pp = pps[int(idx)]
            pcr.pps_pruned.setdefault(pp.goodness, []).append(pp)


# ---- masks ---------------------------------------------------
def get_exclusion_zone(v, hard_exclude=None):
    """Positions a primer may not cover -> [[start, length], ...] (reference lib_primer3.py:564-573)."""
    idx = np.flatnonzero(~np.asarray(v, dtype=bool))
    if hard_exclude is not None and len(hard_exclude):
        idx = np.union1d(idx, np.asarray(hard_exclude, dtype=np.int64))
    return lib_utils.list_2_ranges(idx)


def include_mut_in_included_maps(hroi):
    """Adds a '<name>_mut' twin of every inclusion map with the VCF mutation positions
    cleared (reference lib_primer3.py:576-593)."""
    maps = {}
    for k, mmap in hroi.p3_sequence_included_maps.items():
        maps[k] = np.array(mmap, dtype=bool, copy=True)
        if hroi.vcf_hook:
            twin = np.array(mmap, dtype=bool, copy=True)
            pos = np.array([m.pos - hroi.start for m in hroi.mutations], dtype=np.int64)
            twin[pos] = False
            maps[k + "_mut"] = twin
    hroi.p3_sequence_included_maps = maps


def get_sequence_excluded_regions(hroi):
    hroi.p3_sequence_excluded_regions = {k: get_exclusion_zone(m) for k, m in hroi.p3_sequence_included_maps.items()}


def open_marker_window(hroi):
    """The KASP step 'make sure the marker region of the primer is made available'
    (reference _lib_kasp.py:681-684), restated literally: ``start`` is computed
    from the GENOME coordinate hroi.start, so for any region further than
    PRIMER_MAX_SIZE from the chromosome start the slice is [1:] and every map
    is opened except index 0 (not 'fixed')."""
    size = PRIMER3_GLOBALS["PRIMER_MAX_SIZE"]
    for mmap in hroi.p3_sequence_included_maps.values():
        start = min(1, hroi.start - size)
        stop = max(len(mmap), hroi.start + size)
        mmap[start:stop] = True


# ---- "next" rows: mutation annotation and HRM product Tm ----------------
def add_mutations_batch(pcrs, mutations_per_pcr, device=None):
    """Numeric part of PrimerPair.add_mutations (reference lib_primer3.py:166-210) for many
    markers: fills primers[d].max_aaf and pp.max_indel_size.  ``mutations_per_pcr[m]`` is the
    mutation list (objects with pos / aaf / indel_size) of pcrs[m].  The mutation text and the
    IUPAC-ambiguous primer strings stay with the reference's report writer."""
    pps = [pp for pcr in pcrs for pp in pcr.pps]
    if not pps:
        return
    keys = list(pps[0].primers.keys())
    order = [k for k in ("F", "R", "A") if k in keys]
    pair_seg = np.zeros(len(pcrs) + 1, dtype=np.int64)
    np.cumsum([len(pcr.pps) for pcr in pcrs], out=pair_seg[1:])
    mut_seg = np.zeros(len(pcrs) + 1, dtype=np.int64)
    np.cumsum([len(m) for m in mutations_per_pcr], out=mut_seg[1:])
    muts = [m for ml in mutations_per_pcr for m in ml]
    start = np.zeros((len(pps), 3), dtype=np.int64)
    stop = np.full((len(pps), 3), -1, dtype=np.int64)
    for k, pp in enumerate(pps):
        for c, d in enumerate(order):
            start[k, c], stop[k, c] = pp.primers[d].start, pp.primers[d].stop
    aaf, indel = get_context(device).annotate_mutations(
        [m.pos for m in muts], [m.aaf for m in muts], [m.indel_size for m in muts], mut_seg, pair_seg, start, stop,
        len(order))
    for k, pp in enumerate(pps):
        for c, d in enumerate(order):
            pp.primers[d].max_aaf = max(pp.primers[d].max_aaf, float(aaf[k, c]))
        pp.max_indel_size = max(pp.max_indel_size, int(indel[k]))


def add_mutation_strings(pp, mutations):
    """String half of PrimerPair.add_mutations (reference lib_primer3.py:166-199): per primer the
    mutation labels 'REFposALT1/ALT2:aaf' and the IUPAC-ambiguous sequence.  R primers (and the ALT
    primer of an R-direction KASP pair) are handled on the forward strand and flipped back."""
    for d, p in pp.primers.items():
        flip = d == "R" or (d == "A" and pp.dir == "R")
        fwd = lib_utils.reverse_complement(p.sequence) if flip else p.sequence
        variants = [fwd]
        for m in mutations:
            if p.start <= m.pos <= p.stop:
                base = list(fwd)
                ppos = m.pos - p.start
                for malt in m.alt:
                    base[ppos:(ppos + len(m.ref))] = list(malt)     # edits accumulate, as in the reference
                    if len(base) == len(fwd):
                        variants.append("".join(base))
                p.mutations.append("{}{:d}{}:{}".format(m.ref, m.pos, "/".join(m.alt), lib_utils.round_tidy(m.aaf, 2)))
        amb = lib_utils.seqs2ambiguous_dna(variants)
        p.sequence_ambiguous = lib_utils.reverse_complement(amb) if flip else amb


def hrm_temps_batch(ref_products, alt_products, device=None):
    """PCRproduct.get_hrm_temps (reference lib_markers.py:518-529) for many products:
    list of {"ref", "alt", "delta"} with the reference's round_tidy digits."""
    n = len(ref_products)
    tm = get_context(device).tm_nn_batch(list(ref_products) + list(alt_products))
    out = []
    for r, a in zip(tm[:n], tm[n:]):
        if np.isnan(r) or np.isnan(a):
            raise IndexError("string index out of range")   # Biopython: no A/C/G/T/I base left after _check()
        out.append({"ref": lib_utils.round_tidy(r, 4), "alt": lib_utils.round_tidy(a, 4),
                    "delta": lib_utils.round_tidy(np.abs(r - a), 2)})
    return out


def parse_p3_result(p3_primers, target_start=1):
    """The dict -> [PrimerPair] part of get_primers (reference lib_primer3.py:481-529): pairs in
    primer-pair-id order, coordinates shifted by target_start, every PRIMER_{LEFT,RIGHT,PAIR}_i_*
    field except *_TM stored under its lower-cased name."""
    pps = {}
    for k, v in p3_primers.items():
        if k in ("PRIMER_WARNING", "PRIMER_ERROR"):
            return []          # the whole design run is void
        fields = k.split("_")
        try:
            pid = int(fields[2])
        except (ValueError, IndexError):
            continue
        name = "_".join(fields[3:]).lower()
        if name == "tm":
            continue
        if pid not in pps:
            pps[pid] = PrimerPair()
            for d in ("F", "R"):
                pps[pid].primers[d].id = pid
        pp = pps[pid]
        if fields[1] == "LEFT":
            if name:
                pp.primers["F"].__dict__[name] = v
            else:
                pp.primers["F"].start = target_start + v[0]
                pp.primers["F"].stop = pp.primers["F"].start + (v[1] - 1)
        elif fields[1] == "RIGHT":
            if name:
                pp.primers["R"].__dict__[name] = v
            else:
                pp.primers["R"].stop = target_start + v[0]
                pp.primers["R"].start = pp.primers["R"].stop - (v[1] - 1)
        elif fields[1] == "PAIR":
            pp.__dict__[name] = v
    return [pps[i] for i in sorted(pps)]


def get_primers_batch(seq_args_list, target_starts, device=None):
    """get_primers (reference lib_primer3.py:466-529) for many designPrimers calls at once."""
    prepared = []
    for args in seq_args_list:
        args = dict(args)
        for k, v in PRIMER3_GLOBALS.items():
            if k.startswith("SEQUENCE") and k not in args:
                args[k] = v
        for k in ("SEQUENCE_EXCLUDED_REGION", "SEQUENCE_INCLUDED_REGION"):
            if k in args:
                args[k] = lib_utils.reduce_ivs(args[k], n=200)   # primer3 refuses more than 200 intervals
        prepared.append(args)
    results = primer3_bindings.design_primers_batch(prepared, device=device)
    return [parse_p3_result(r, ts) for r, ts in zip(results, target_starts)]


def get_primers(primer3_seq_args, target_start=1):
    return get_primers_batch([primer3_seq_args], [target_start])[0]
