"""Host-side mirror of reference src/polyoligo/_lib_pcr.py: the primer-pair selection
``design_primers`` (:167-219; also what the CAPS and HRM assays call), the design part of ``main``
(:222-356) for MANY regions at once (``design_regions``; the CAPS and HRM variants of
``_lib_caps.main`` :214-389 / ``_lib_hrm.main`` :190-341 are the same flow with one extra step and
live in ``caps.py`` / ``hrm.py``) and the report writer (:62-164).

Both primers are picked by the GPU primer picker (one batch per search mask over every region that
still needs pairs), validity, heterodimers, mutation annotation, product melting temperatures,
scoring and top-n run as one device batch each.  BLAST / MUSCLE homolog mapping and BLAST
off-target search stay external: the flank ROIs arrive with their inclusion maps, and
``offtarget_hook(pcrs)`` is called where the reference calls ``check_offtargeting``."""
import numpy as np

from . import lib_markers, lib_primer3, lib_utils

DELIMITER = "\t"
HEADER = ["target", "chr", "start", "end", "direction", "assay_id", "seq_5_3", "seq_5_3_ambiguous", "primer_id",
          "goodness", "qcode", "length", "prod_size", "tm", "gc_content", "n_offtargets", "max_aaf", "indels",
          "offtarget_products", "mutations"]            # reference _lib_pcr.py:17-38
SEARCH_TYPES = ["mism_mut", "mism", "partial_mism_mut", "partial_mism", "all_mut", "all"]   # :316-319


def _select(pps_repo, p3_repo, valid, roi, n_primers, max_unique):
    """The selection loop of design_primers (reference :182-219) given the picker's pairs and the
    validity of their primers."""
    p_counts = {}
    if not np.isinf(max_unique):
        for pp in pps_repo:
            for d in ("F", "R"):
                seq = pp.primers[d].sequence
                p_counts[seq] = p_counts.get(seq, 0) + 1
    for pp in p3_repo:
        if len(pps_repo) == n_primers:
            break
        pp.chrom = roi.chrom
        is_pp_valid = True
        for d in pp.primers:
            if not valid[id(pp.primers[d])]:
                is_pp_valid = False
            if not np.isinf(max_unique):
                seq = pp.primers[d].sequence
                if seq in p_counts:
                    if p_counts[seq] >= max_unique:
                        is_pp_valid = False
                    else:
                        p_counts[seq] += 1     # counted even if the pair is rejected (reference quirk)
                else:
                    p_counts[seq] = 1
        if is_pp_valid:
            pps_repo.append(pp)
    return pps_repo


def design_primers_multi(jobs, n_primers=10, max_unique=np.inf, device=None):
    """design_primers for many regions in one picker batch and one validity batch.
    jobs: dicts(pps_repo, roi, sequence_target, sequence_excluded_region); returns the repositories."""
    seq_args_list, starts = [], []
    for job in jobs:
        seq_args = {"SEQUENCE_TEMPLATE": job["roi"].seq}
        if job.get("sequence_target") is not None:
            seq_args["SEQUENCE_TARGET"] = job["sequence_target"]
        if job.get("sequence_excluded_region") is not None:
            seq_args["SEQUENCE_EXCLUDED_REGION"] = job["sequence_excluded_region"]
        seq_args_list.append(seq_args)
        starts.append(job["roi"].left_roi.start)
    try:
        p3_repos = lib_primer3.get_primers_batch(seq_args_list, starts, device)
    except Exception:          # the reference swallows designPrimers failures per call (:177-180)
        p3_repos = []
        for seq_args, start in zip(seq_args_list, starts):
            try:
                p3_repos.append(lib_primer3.get_primers_batch([seq_args], [start], device)[0])
            except Exception:
                p3_repos.append([])
    # validity of every returned primer in one batch (the reference asks per pair, :206-208)
    primers = [pp.primers[d] for repo in p3_repos for pp in repo for d in pp.primers]
    ok = lib_primer3.is_valid_batch(primers, device) if primers else []
    valid = {id(p): bool(v) for p, v in zip(primers, ok)}
    return [_select(job["pps_repo"], repo, valid, job["roi"], n_primers, max_unique)
            for job, repo in zip(jobs, p3_repos)]


def design_primers(pps_repo, roi, sequence_target=None, sequence_excluded_region=None, n_primers=10,
                   max_unique=np.inf, device=None):
    return design_primers_multi([dict(pps_repo=pps_repo, roi=roi, sequence_target=sequence_target,
                                      sequence_excluded_region=sequence_excluded_region)],
                                n_primers, max_unique, device)[0]


def prepare_region(region, fragment_min_size=None):
    """reference _lib_pcr.main :285-296 (_lib_caps.main :292-317, _lib_hrm.main :262-272): the
    '_mut' twins of the flank maps, the merged template and its excluded intervals per search mask.
    ``region.left_roi`` / ``right_roi`` carry p3_sequence_included_maps and mutations."""
    for flank in (region.left_roi, region.right_roi):
        if flank.mutations is None:
            flank.upload_mutations()
        lib_primer3.include_mut_in_included_maps(flank)
    if fragment_min_size is not None:     # CAPS: room for both fragments around the marker
        region = lib_markers.Region(chrom=region.chrom, start=region.marker.pos1 - fragment_min_size,
                                    stop=region.marker.pos1 + fragment_min_size, blast_hook=region.blast_hook,
                                    malign_hook=region.malign_hook, vcf_hook=region.vcf_hook, name=region.name,
                                    marker=region.marker, do_print_alt_subjects=region.do_print_alt_subjects,
                                    left_roi=region.left_roi, right_roi=region.right_roi)
    merged = region.merge_with_primers()
    lib_primer3.get_sequence_excluded_regions(merged)
    return merged


def search_masks(pcrs, merged, product_ranges=None, device=None):
    """The search-mask loop of main (:316-338) for many regions: per mask ONE picker batch over the
    regions whose repository is still short of PRIMER_NUM_RETURN (regions that need their own
    PRIMER_PRODUCT_SIZE_RANGE are grouped by it).  Returns the number of new pairs per region and mask."""
    depth = lib_primer3.PRIMER3_GLOBALS["PRIMER_NUM_RETURN"]
    new_pairs = [dict() for _ in pcrs]
    kept_range = lib_primer3.PRIMER3_GLOBALS.get("PRIMER_PRODUCT_SIZE_RANGE")
    for search_type in SEARCH_TYPES:
        todo = [k for k, (pcr, reg) in enumerate(zip(pcrs, merged))
                if search_type in reg.p3_sequence_excluded_regions and len(pcr.pps) < depth]
        groups = {}
        for k in todo:
            key = None if product_ranges is None else tuple(tuple(x) for x in product_ranges[k])
            groups.setdefault(key, []).append(k)
        for key, ks in groups.items():
            if key is not None:
                lib_primer3.set_globals(PRIMER_PRODUCT_SIZE_RANGE=[list(x) for x in key])
            before = [len(pcrs[k].pps) for k in ks]
            jobs = [dict(pps_repo=pcrs[k].pps, roi=merged[k], sequence_target=merged[k].p3_sequence_target,
                         sequence_excluded_region=merged[k].p3_sequence_excluded_regions[search_type]) for k in ks]
            for k, repo, b in zip(ks, design_primers_multi(jobs, depth, 2, device), before):
                pcrs[k].pps = repo
                new_pairs[k][search_type] = len(repo) - b
    if product_ranges is not None and kept_range is not None:
        lib_primer3.set_globals(PRIMER_PRODUCT_SIZE_RANGE=kept_range)
    return new_pairs


def finish(pcrs, merged, n_primers, assay_name, device=None):
    """Mutation annotation, heterodimers, goodness and top-n of every region in one device pass each
    (reference main :347-350)."""
    lib_primer3.add_mutations_batch(pcrs, [reg.mutations or [] for reg in merged], device)
    for pcr, reg in zip(pcrs, merged):
        for pp in pcr.pps:
            lib_primer3.add_mutation_strings(pp, reg.mutations or [])
    lib_primer3.rank_batch(pcrs, n_primers, assay_name, device)
    for pcr in pcrs:
        if not pcr.pps:
            pcr.pps_classified, pcr.pps_pruned = {}, {}
    return [sum(len(v) for v in pcr.pps_pruned.values()) for pcr in pcrs]


def design_regions(regions, n_primers=10, offtarget_hook=None, per_region_range=False, device=None):
    """_lib_pcr.main from the homolog-mapped flanks to the ranked pairs, for many regions.
    PRIMER_PRODUCT_SIZE_RANGE: when the globals do not hold one, the reference sets it from the FIRST
    region a process sees ([[target length, template length]], :323-324) and keeps it for the regions
    that follow; that is mirrored unless per_region_range=True gives every region its own."""
    merged = [prepare_region(r) for r in regions]
    pcrs = [lib_primer3.PCR(chrom=reg.chrom, name=reg.name) for reg in merged]
    ranges = None
    if "PRIMER_PRODUCT_SIZE_RANGE" not in lib_primer3.PRIMER3_GLOBALS and merged:
        if per_region_range:
            ranges = [[[reg.p3_sequence_target[0][1], reg.n]] for reg in merged]
        else:
            lib_primer3.set_globals(PRIMER_PRODUCT_SIZE_RANGE=[[merged[0].p3_sequence_target[0][1], merged[0].n]])
    search_masks(pcrs, merged, ranges, device)
    if offtarget_hook is not None:
        offtarget_hook(pcrs)
    finish(pcrs, merged, n_primers, "PCR", device)
    return pcrs, merged


def print_report_header(fp, header=HEADER):
    with open(fp, "w") as f:
        f.write("{}\n".format(DELIMITER.join(header)))


def write_pair_report(pcr, fp, n_header, lead, bed_name, mid=None, tail=None, per_pair=None):
    """The F / R report of the PCR, HRM and CAPS assays (reference _lib_pcr.py:70-164,
    _lib_hrm.py:63-165, _lib_caps.py:83-195): ``lead`` are the region / marker columns, ``mid(pp, x)``
    the assay columns between prod_size and tm, ``tail(pp)`` the trailing ones; ``per_pair(pp)`` yields
    the items a pair is written once for (CAPS: its enzymes).  Primer ids number the distinct sequences
    per direction by first appearance; assay ids count written blocks."""
    seq_ids = {"F": [], "R": []}
    ppid = 0
    with open(fp + ".txt", "w") as f, open(fp + ".bed", "w") as f_bed:
        for score in sorted(set(pcr.pps_pruned.keys()), reverse=True):
            for pp in pcr.pps_pruned[score]:
                pp.id = ppid
                ids = {}
                for d in ("F", "R"):
                    seq = pp.primers[d].sequence
                    if seq not in seq_ids[d]:
                        seq_ids[d].append(seq)
                    ids[d] = d + "_" + str(seq_ids[d].index(seq))
                offtargets = ",".join(pp.offtargets) or "NA"
                for item in (per_pair(pp) if per_pair else [None]):
                    for d in ("F", "R"):
                        p = pp.primers[d]
                        if pp.qcode == "":
                            pp.qcode = "."
                        fields = list(lead) + (mid(pp, item)[0] if mid else [])
                        fields += [int(p.start), int(p.stop), d, pp.id, p.sequence, p.sequence_ambiguous, ids[d],
                                   pp.goodness, pp.qcode, p.length, pp.product_size]
                        fields += mid(pp, item)[1] if mid else []
                        fields += [lib_utils.round_tidy(p.tm, 3), lib_utils.round_tidy(p.gc_content, 3),
                                   len(pp.offtargets), lib_utils.round_tidy(p.max_aaf, 2), pp.max_indel_size,
                                   offtargets, ",".join(p.mutations) or "NA"]
                        fields += tail(pp) if tail else []
                        f.write("{}\n".format(DELIMITER.join(str(x) for x in fields)))
                        bed = [pcr.chrom, int(p.start), int(p.stop), "{}_{}-{}".format(bed_name, ids[d], pp.goodness),
                               "0", "+" if d == "F" else "-"]
                        f_bed.write("{}\n".format("\t".join(str(x) for x in bed)))
                    f.write("\n")
                    ppid += 1
        if ppid == 0:
            f.write(DELIMITER.join([str(x) for x in lead] + (n_header - len(lead)) * ["NA"]) + "\n")
            f.write("\n")
        f.write("\n")


def print_report(pcr, fp):
    write_pair_report(pcr, fp, len(HEADER), [pcr.name, pcr.chrom], pcr.name)
