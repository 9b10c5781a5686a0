"""Host-side mirror of reference src/polyoligo/_lib_caps.py: the CAPS design flow from the
homolog-mapped flank ROIs to the ranked pairs (``main`` :214-389) for many markers at once, the
restriction-enzyme filter and the report writer (:76-195).  Same flow as ``pcr.design_regions`` with
the template widened to marker -/+ fragment_min_size, the product range always [[target, template]]
(:343) and only the pairs kept whose product exactly one allele lets an enzyme cut once
(``PCRproduct.will_it_cut``, host regular expressions)."""
import json
import os

from . import lib_markers, lib_primer3, pcr as _pcr

OUTER_LEN = 1000
ENZYME_FILENAME = os.path.join(os.path.dirname(__file__), "data", "type2.json")
DELIMITER = "\t"
HEADER = ["marker", "chr", "pos", "ref", "alt", "enzymes", "enzyme_suppliers", "start", "end", "direction", "assay_id",
          "seq_5_3", "seq_5_3_ambiguous", "primer_id", "goodness", "qcode", "length", "prod_size", "allele_cut",
          "fragment_1", "fragment_2", "tm", "gc_content", "n_offtargets", "max_aaf", "indels", "offtarget_products",
          "mutations", "PCR_product"]


class PCR(lib_primer3.PCR):
    pass


def get_enzymes(included_enzymes=None, filename=None):
    """reference _lib_caps.py:55-70.  The enzyme table (REBASE type II: name, supplier, regex F / R)
    is third-party data the reference ships as data/type2.json; it is read from ``filename`` (same JSON
    format), by default from this package's data/ directory if the file has been placed there."""
    filename = filename or ENZYME_FILENAME
    if not os.path.exists(filename):
        raise FileNotFoundError("restriction enzyme table %s not found (reference src/polyoligo/data/type2.json)"
                                % filename)
    with open(filename) as f:
        enzymes = json.load(f)
    if included_enzymes is not None:
        enzymes = [e for e in enzymes if e["name"] in included_enzymes]
    return enzymes


def design_regions(regions, enzymes, fragment_min_size, n_primers=10, offtarget_hook=None, device=None):
    merged = [_pcr.prepare_region(r, fragment_min_size) for r in regions]
    pcrs = [PCR(snp_id=reg.marker.name, chrom=reg.marker.chrom, pos=reg.marker.pos1, ref=reg.marker.ref,
                alt=reg.marker.alt) for reg in merged]
    ranges = [[[reg.p3_sequence_target[0][1], reg.n]] for reg in merged]
    _pcr.search_masks(pcrs, merged, ranges, device)
    if merged:      # the reference leaves the last region's range in the globals (:343)
        lib_primer3.set_globals(PRIMER_PRODUCT_SIZE_RANGE=ranges[-1])
    if offtarget_hook is not None:
        offtarget_hook(pcrs)
    for pcr, reg in zip(pcrs, merged):
        kept = []
        # one enzyme list per region, as every _lib_caps.main call loads its own (:251); WITHIN a region
        # the accepted enzyme dicts are shared between the pairs (see PCRproduct.will_it_cut)
        region_enzymes = [dict(e) for e in enzymes]
        for pp in pcr.pps:
            product = lib_markers.PCRproduct(roi=reg, pp=pp)
            product.will_it_cut(enzymes=region_enzymes, min_fragment_size=fragment_min_size)
            if product.enzymes:
                pp.enzymes = product.enzymes
                pp.seq_x = product.roi.seq_x
                kept.append(pp)
        pcr.pps = kept
    _pcr.finish(pcrs, merged, n_primers, "CAPS", device)
    return pcrs, merged


def print_report_header(fp):
    _pcr.print_report_header(fp, HEADER)


def print_report(pcr, fp):
    lead = [pcr.snp_id, pcr.chrom, int(pcr.pos), pcr.ref, pcr.alt]
    _pcr.write_pair_report(pcr, fp, len(HEADER), lead, pcr.snp_id,
                           mid=lambda pp, e: ([e["name"], e["supplier"]], [e["allele"], e["fragl"], e["fragr"]]),
                           tail=lambda pp: [pp.seq_x], per_pair=lambda pp: pp.enzymes)
