"""Host-side mirror of the hot-path parts of reference src/polyoligo/_lib_kasp.py:
allele-specific marker-primer enumeration (get_marker_primers, :295-359) and
the reporter-tailed heterodimer check (:75-117), batched over markers.

``get_marker_primers_batch`` enumerates every marker's 2 x (MAX-MIN+1)
candidates with the CUDA enumeration kernel, scores all of them with one
oligo-properties pass and keeps, per marker and in the reference's order, the
slots whose REF and ALT primers are both valid.
"""
from copy import deepcopy

import numpy as np

from . import _native, lib_primer3, lib_utils
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
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.is_indel = None

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


def merge_primers(mpp, p3_primers):
    """reference _lib_kasp.py:552-566: one PrimerPair per primer3 pair, the marker primer pair
    completed with the picked common primer; the ALT primer takes the marker primer's coordinates."""
    pps = []
    for pp3 in p3_primers:
        pp = deepcopy(mpp)
        pp.add_primer3_attributes(pp3)
        pp.primers["A"].start = pp.primers[pp.dir].start
        pp.primers["A"].stop = pp.primers[pp.dir].stop
        pp.primers["A"].id = pp.primers[pp.dir].id
        pps.append(pp)
    return pps


def design_primers_multi(jobs, n_primers=10, device=None):
    """design_primers (reference _lib_kasp.py:568-631) for many markers: ``jobs`` is a list of
    dicts(pps_repo, mpps, roi, sequence_target=None, sequence_excluded_region=None).  Every
    designPrimers call of every marker goes to the GPU in one batch, the validity of every
    picked common primer is evaluated in one batch, and the round-robin merge then makes the
    reference's decisions in the reference's order."""
    seq_args, starts, owner = [], [], []
    for j, job in enumerate(jobs):
        roi = job["roi"]
        base = {"SEQUENCE_TEMPLATE": roi.seq}
        if job.get("sequence_target") is not None:
            base["SEQUENCE_TARGET"] = job["sequence_target"]
        if job.get("sequence_excluded_region") is not None:
            base["SEQUENCE_EXCLUDED_REGION"] = job["sequence_excluded_region"]
        for mpp in job["mpps"]:
            args = dict(base)
            key = "SEQUENCE_PRIMER" if mpp.dir == "F" else "SEQUENCE_PRIMER_REVCOMP"
            args[key] = mpp.primers[mpp.dir].sequence
            seq_args.append(args)
            starts.append(roi.start)
            owner.append(j)
    p3_lists = lib_primer3.get_primers_batch(seq_args, starts, device) if seq_args else []
    repos = [{} for _ in jobs]
    k = 0
    for j, job in enumerate(jobs):
        for pid, mpp in enumerate(job["mpps"]):
            repos[j][pid] = merge_primers(mpp, p3_lists[k])
            k += 1
    # Primer.is_valid of every picked common primer (:621-625), once per distinct sequence
    commons = [pp.primers["R" if pp.dir == "F" else "F"] for repo in repos for lst in repo.values() for pp in lst]
    by_seq = {}
    for p in commons:
        by_seq.setdefault(p.sequence, lib_primer3.Primer(sequence=p.sequence))
    uniq = list(by_seq.values())
    ok = dict(zip(by_seq.keys(), lib_primer3.is_valid_batch(uniq, device))) if uniq else {}
    for p in commons:
        src = by_seq[p.sequence]
        for a in ("length", "gc_content", "nN", "tm", "hairpin_tm", "homodimer_tm"):
            setattr(p, a, getattr(src, a))
    out = []
    for job, p3_repo in zip(jobs, repos):
        pps_repo = job["pps_repo"]
        flag_continue = True
        while flag_continue:
            for mpp_id in [m for m in p3_repo if len(p3_repo[m]) == 0]:
                del p3_repo[mpp_id]
            if len(p3_repo) == 0:
                flag_continue = False
            for mpp_id, p3_pps in p3_repo.items():
                pp = p3_pps.pop(0)
                is_new = not any(all(rpp.primers[d].sequence == pp.primers[d].sequence for d in "FRA")
                                 for rpp in pps_repo)
                if is_new:
                    common = pp.primers["R" if pp.dir == "F" else "F"]
                    if ok[common.sequence]:
                        pps_repo.append(pp)
                    if len(pps_repo) == n_primers:
                        flag_continue = False
                        break
        out.append(pps_repo)
    return out


def design_primers(pps_repo, mpps, roi, sequence_target=None, sequence_excluded_region=None, n_primers=10,
                   device=None):
    return design_primers_multi([dict(pps_repo=pps_repo, mpps=mpps, roi=roi, sequence_target=sequence_target,
                                      sequence_excluded_region=sequence_excluded_region)], n_primers, device)[0]


SEARCH_TYPES_VCF = ["mism_mut", "mism", "partial_mism_mut", "partial_mism", "all_mut", "all"]
SEARCH_TYPES = ["mism", "partial_mism", "all"]


def design_markers(hrois, n_primers=10, reporters=None, device=None):
    """The design part of _lib_kasp.main (reference _lib_kasp.py:671-743) for many homolog-mapped
    ROIs at once, with the BLAST off-target search left out (no off-targets).  Each hroi carries
    seq / start / stop / marker / chrom, p3_sequence_included_maps, mutations and vcf_hook like the
    reference's lib_markers.ROI after map_homologs + upload_mutations.  Returns one PCR per marker,
    classified and pruned."""
    depth = lib_primer3.PRIMER3_GLOBALS["PRIMER_NUM_RETURN"]
    for hroi in hrois:
        lib_primer3.include_mut_in_included_maps(hroi)
        lib_primer3.open_marker_window(hroi)
        lib_primer3.get_sequence_excluded_regions(hroi)
    mpps_all = get_marker_primers_batch(hrois, device)
    pcrs = [PCR(snp_id=h.marker.name if hasattr(h.marker, "name") else None, chrom=getattr(h, "chrom", None),
                pos=h.marker.pos1, ref=h.marker.ref, alt=h.marker.alt) for h in hrois]
    n_types = max(len(h.p3_sequence_included_maps) for h in hrois) if hrois else 0
    for step in range(n_types):
        jobs, who = [], []
        for m, (hroi, pcr) in enumerate(zip(hrois, pcrs)):
            types_ = SEARCH_TYPES_VCF if "mism_mut" in hroi.p3_sequence_included_maps else SEARCH_TYPES
            if step < len(types_) and len(pcr.pps) < depth:
                jobs.append(dict(pps_repo=pcr.pps, mpps=mpps_all[m], roi=hroi,
                                 sequence_excluded_region=hroi.p3_sequence_excluded_regions[types_[step]]))
                who.append(m)
        if jobs:
            for m, repo in zip(who, design_primers_multi(jobs, n_primers=depth, device=device)):
                pcrs[m].pps = repo
    for pcr in pcrs:
        pcr.add_reporter_dyes(reporters or ("", ""))
    lib_primer3.add_mutations_batch(pcrs, [list(h.mutations) for h in hrois], device)
    for pcr, hroi in zip(pcrs, hrois):
        if hroi.mutations and hasattr(hroi.mutations[0], "alt"):
            for pp in pcr.pps:
                lib_primer3.add_mutation_strings(pp, hroi.mutations)
    lib_primer3.rank_batch(pcrs, n_primers, "KASP", device, reporters=reporters)
    return pcrs


# ---- report writer (reference _lib_kasp.py:14-45, 360-530) --------------------------------------
DELIMITER = "\t"
HEADER = ["marker", "chr", "pos", "ref", "alt", "start", "end", "direction", "type", "assay_id", "seq_5_3",
          "seq_5_3_ambiguous", "primer_id", "goodness", "qcode", "length", "prod_size", "tm", "gc_content",
          "will_dimerize", "n_offtargets", "max_aaf", "indels", "offtarget_products", "mutations"]


def print_report_header(fp):
    with open(fp, "w") as f:
        f.write("{}\n".format(DELIMITER.join(HEADER)))


def _allele_case(seq, n):
    """lower case except the last n bases (the allele at the 3' end)"""
    return seq.lower() if n == 0 else seq[:-n].lower() + seq[-n:]


def print_report(pcr, fp):
    """Per-marker TSV (<fp>.txt) and BED (<fp>.bed) of the pruned pairs, byte-identical to
    reference _lib_kasp.py:367-530: REF / ALT / COM rows per assay, primer ids numbered by first
    appearance of the (case-marked) sequence, round_tidy'd Tm / GC / aaf."""
    order = ["REF", "ALT", "COM"]
    seq_ids = {pt: [] for pt in order}
    ppid = 0
    with open(fp + ".txt", "w") as f, open(fp + ".bed", "w") as f_bed:
        for score in sorted(set(pcr.pps_pruned.keys()), reverse=True):
            for pp in pcr.pps_pruned[score]:
                pp.id = ppid
                dyes = {d: pp.primers[d].reporter for d in pp.primers}
                marker_d, common_d = ("F", "R") if pp.dir == "F" else ("R", "F")
                seqs = {"REF": _allele_case(pp.primers[marker_d].sequence, pcr.ref_n),
                        "COM": pp.primers[common_d].sequence.lower(),
                        "ALT": _allele_case(pp.primers["A"].sequence, pcr.alt_n)}
                seqs_amb = {"REF": _allele_case(pp.primers[marker_d].sequence_ambiguous, pcr.ref_n),
                            "COM": pp.primers[common_d].sequence_ambiguous.lower(),
                            "ALT": _allele_case(pp.primers["A"].sequence_ambiguous, pcr.alt_n)}
                ids = {}
                for pt in order:
                    if seqs[pt] not in seq_ids[pt]:
                        seq_ids[pt].append(seqs[pt])
                    ids[pt] = pt + "_" + str(seq_ids[pt].index(seqs[pt]))
                offtargets = ",".join(pp.offtargets) or "NA"
                for ptype, d in zip(order, [marker_d, "A", common_d]):
                    mutations = ",".join(pp.primers[d].mutations) or "NA"
                    will_dimerize, direction = (pp.alt_dimer, pp.dir) if d == "A" else (pp.ref_dimer, d)
                    if pp.qcode == "":
                        pp.qcode = "."
                    ref, alt = pcr.ref, pcr.alt
                    if pcr.is_indel == "ins":
                        ref = "*"
                    elif pcr.is_indel == "del":
                        alt = "*"
                    p = pp.primers[d]
                    fields = [pcr.snp_id, pcr.chrom, int(pcr.pos), ref, alt, int(p.start), int(p.stop), direction,
                              ptype, pp.id, dyes[d] + seqs[ptype], dyes[d] + seqs_amb[ptype],
                              pcr.snp_id + "_" + ids[ptype], pp.goodness, pp.qcode, p.length, pp.product_size,
                              lib_utils.round_tidy(p.tm, 3), lib_utils.round_tidy(p.gc_content, 3), will_dimerize,
                              len(pp.offtargets), lib_utils.round_tidy(p.max_aaf, 2), pp.max_indel_size,
                              offtargets, mutations]
                    f.write("{}\n".format(DELIMITER.join(str(x) for x in fields)))
                    bed = [pcr.chrom, int(p.start), int(p.stop),
                           "{}_{}-{}".format(pcr.snp_id, ids[ptype], pp.goodness), "0",
                           "+" if direction == "F" else "-"]
                    f_bed.write("{}\n".format("\t".join(str(x) for x in bed)))
                f.write("\n")
                ppid += 1
        if ppid == 0:
            fields = [str(x) for x in (pcr.snp_id, pcr.chrom, int(pcr.pos), pcr.ref, pcr.alt)]
            f.write(DELIMITER.join(fields + (len(HEADER) - 5) * ["NA"]) + "\n")
            f.write("\n")
        f.write("\n")

