"""Host-side mirror of the primer-pair selection of reference src/polyoligo/_lib_pcr.py:167-219
(design_primers; also what the CAPS and HRM assays call through their own wrappers): both
primers are picked by the primer picker, then pairs are kept in penalty order while both
primers pass Primer.is_valid and no primer sequence is used more than ``max_unique`` times."""
import numpy as np

from . import lib_primer3


def design_primers(pps_repo, roi, sequence_target=None, sequence_excluded_region=None, n_primers=10,
                   max_unique=np.inf, device=None):
    seq_args = {"SEQUENCE_TEMPLATE": roi.seq}
    if sequence_target is not None:
        seq_args["SEQUENCE_TARGET"] = sequence_target
    if sequence_excluded_region is not None:
        seq_args["SEQUENCE_EXCLUDED_REGION"] = sequence_excluded_region
    try:
        p3_repo = lib_primer3.get_primers(seq_args, target_start=roi.left_roi.start)
    except Exception:          # the reference swallows designPrimers failures here (:177-180)
        p3_repo = []
    # validity of every returned primer in one batch (the reference asks per pair, :206-208)
    primers = [pp.primers[d] for pp in p3_repo for d in pp.primers]
    ok = lib_primer3.is_valid_batch(primers, device) if primers else []
    valid = {id(p): bool(v) for p, v in zip(primers, ok)}
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
