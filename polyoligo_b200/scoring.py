"""Heuristic goodness score of a primer pair: same names and results as
reference src/polyoligo/scoring.py, plus a decoder for the bit form the CUDA
ranking kernel returns (po_rank_pairs_batch)."""
import numpy as np

from ._native import Q_BITS

_QCODE_ORDER = "hHtOdmMiI"      # the order in which scoring.py appends letters


def qcode_from_bits(bits):
    return "".join(c for c in _QCODE_ORDER if bits & Q_BITS[c])


def _common(pp, dimer, w_off, w_aaf):
    score, q = 0, ""
    tms = np.array([p.tm for p in pp.primers.values()])
    if np.sum(np.abs(tms - np.mean(tms))) <= 5:
        score += 1
    else:
        q += "t"
    if len(pp.offtargets) == 0:
        score += w_off
    else:
        q += "O"
    if not dimer:
        score += 1
    else:
        q += "d"
    aaf = np.array([p.max_aaf for p in pp.primers.values()])
    if np.all(aaf < 0.1):
        score += w_aaf
        if np.all(aaf == 0):
            score += 1
        else:
            q += "m"
    else:
        q += "M"
    if pp.max_indel_size < 50:
        score += 1
        if pp.max_indel_size == 0:
            score += 1
        else:
            q += "i"
    else:
        q += "I"
    return score, q


def score_base(pp):
    return _common(pp, pp.dimer, 3, 2)


def score_kasp(pp):
    return _common(pp, pp.ref_dimer or pp.alt_dimer, 3, 2)


def score_hrm(pp):
    score, q = 0, ""
    if pp.hrm["delta"] >= 0.5:
        score += 2
        if pp.hrm["delta"] >= 1:
            score += 1
        else:
            q += "h"
    else:
        q += "H"
    s2, q2 = _common(pp, pp.dimer, 1, 1)
    return score + s2, q + q2


def get_score_fnc(assay_name):
    return {"KASP": score_kasp, "PCR": score_base, "CAPS": score_base, "HRM": score_hrm}.get(assay_name)
