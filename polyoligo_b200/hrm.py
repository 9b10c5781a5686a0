"""Host-side mirror of reference src/polyoligo/_lib_hrm.py: the HRM design flow from the
homolog-mapped flank ROIs to the ranked pairs (``main`` :190-341) for many markers at once, and the
report writer (:56-165).  Same flow as ``pcr.design_regions`` plus the product melting temperatures
of both alleles (``PCRproduct.get_hrm_temps``, one ``po_tm_nn_batch`` call for every product of the
batch) and the HRM score / delta-ordered prune."""
from . import lib_markers, lib_primer3, pcr as _pcr

OUTER_LEN = 1000         # reference _lib_hrm.py:14: the flanks sit at marker -/+ OUTER_LEN / 2
DELIMITER = "\t"
HEADER = ["marker", "chr", "pos", "ref", "alt", "delta_Tm", "ref_Tm", "alt_Tm", "start", "end", "direction",
          "assay_id", "seq_5_3", "seq_5_3_ambiguous", "primer_id", "goodness", "qcode", "length", "prod_size", "tm",
          "gc_content", "n_offtargets", "max_aaf", "indels", "offtarget_products", "mutations", "PCR_product"]


class PCR(lib_primer3.PCR):
    pass


def marker_pcr(region):
    m = region.marker
    return PCR(snp_id=m.name, chrom=m.chrom, pos=m.pos1, ref=m.ref, alt=m.alt)


def design_regions(regions, n_primers=10, offtarget_hook=None, device=None):
    merged = [_pcr.prepare_region(r) for r in regions]
    pcrs = [marker_pcr(reg) for reg in merged]
    _pcr.search_masks(pcrs, merged, None, device)
    if offtarget_hook is not None:
        offtarget_hook(pcrs)
    products = [lib_markers.PCRproduct(roi=reg, pp=pp) for pcr, reg in zip(pcrs, merged) for pp in pcr.pps]
    temps = lib_primer3.hrm_temps_batch([p.roi.seq for p in products], [p.roi.seq_alt for p in products],
                                        device) if products else []
    for prod, hrm in zip(products, temps):
        prod.hrm = prod.pp.hrm = hrm
        prod.pp.seq_x = prod.roi.seq_x
    _pcr.finish(pcrs, merged, n_primers, "HRM", device)
    return pcrs, merged


def print_report_header(fp):
    _pcr.print_report_header(fp, HEADER)


def print_report(pcr, fp):
    lead = [pcr.snp_id, pcr.chrom, int(pcr.pos), pcr.ref, pcr.alt]
    _pcr.write_pair_report(pcr, fp, len(HEADER), lead, pcr.snp_id,
                           mid=lambda pp, _: ([pp.hrm["delta"], pp.hrm["ref"], pp.hrm["alt"]], []),
                           tail=lambda pp: [pp.seq_x])
