"""CPU ORACLE of the primer picker (test infrastructure, NOT product code).

Brute-force restatement of what libprimer3's choose_primers does for the tags the
reference sets (reference src/polyoligo/lib_primer3.py:466-529,539-544 and
src/polyoligo/data/PRIMER3_*.yaml), i.e. the semantics DESIGN.md section 7 lists:
every candidate oligo is enumerated, every alignment of every cheap-valid pair is
evaluated eagerly with the C oracle's thal, and the answer is read off a full
sort.  libprimer3 itself (primer3-py 0.6.0) is absent from this machine, so this
restatement is written from the published algorithm: PARITY UNPINNED.

Only tests/ may import this module.
"""
import math

from .oracle import Oracle, ANY, END1, END2, HAIRPIN

_RC = str.maketrans("ACGT", "TGCA")

DEFAULTS = dict(
    min_size=18, opt_size=20, max_size=27, num_return=5, min_tm=57.0, opt_tm=60.0, max_tm=63.0,
    min_gc=20.0, max_gc=80.0, max_self_any_th=47.0, max_self_end_th=47.0, max_hairpin_th=47.0,
    pair_max_compl_any_th=47.0, pair_max_compl_end_th=47.0, pair_max_diff_tm=100.0,
    wt_tm_gt=1.0, wt_tm_lt=1.0, wt_size_lt=1.0, wt_size_gt=1.0,
    mv_conc=50.0, dv_conc=1.5, dntp_conc=0.6, dna_conc=50.0, max_poly_x=5, gc_clamp=0, max_end_gc=5,
    size_ranges=((100, 300),))


def penalty(s, tm, ln):
    """libprimer3 p_obj_fn with the default weights (Tm and size terms only)."""
    q = 0.0
    if s["wt_tm_gt"] and tm > s["opt_tm"]:
        q += s["wt_tm_gt"] * (tm - s["opt_tm"])
    if s["wt_tm_lt"] and tm < s["opt_tm"]:
        q += s["wt_tm_lt"] * (s["opt_tm"] - tm)
    if s["wt_size_lt"] and ln < s["opt_size"]:
        q += s["wt_size_lt"] * (s["opt_size"] - ln)
    if s["wt_size_gt"] and ln > s["opt_size"]:
        q += s["wt_size_gt"] * (ln - s["opt_size"])
    return q


def _sequence_rules(s, seq):
    if s["gc_clamp"] > 0 and any(c not in "GC" for c in seq[-s["gc_clamp"]:]):
        return False
    if sum(c in "GC" for c in seq[-5:]) > s["max_end_gc"]:
        return False
    run, best = 1, 1
    for a, b in zip(seq, seq[1:]):
        run = run + 1 if a == b else 1
        best = max(best, run)
    return best <= s["max_poly_x"]


def cheap_record(s, template, mask, side, start, ln):
    """Position / composition / Tm constraints of one oligo; None when rejected.
    side 'L': bases [start, start+ln); side 'R': bases [start-ln+1, start], reverse-complemented."""
    lo, hi = (start, start + ln) if side == "L" else (start - ln + 1, start + 1)
    if lo < 0 or hi > len(template):
        return None
    window = template[lo:hi].upper()
    if any(c not in "ACGT" for c in window) or any(mask[lo:hi]):
        return None
    seq = window if side == "L" else window[::-1].translate(_RC)
    gc = 100.0 * sum(c in "GC" for c in seq) / ln
    if gc < s["min_gc"] or gc > s["max_gc"]:
        return None
    tm = Oracle.calcTm(seq, mv=s["mv_conc"], dv=s["dv_conc"], dntp=s["dntp_conc"], dna_conc=s["dna_conc"],
                       max_nn_length=36)
    if tm < s["min_tm"] or tm > s["max_tm"]:
        return None
    if not _sequence_rules(s, seq):
        return None
    return dict(seq=seq, start=start, len=ln, tm=tm, gc_percent=gc, penalty=penalty(s, tm, ln))


def candidates(s, template, mask, side):
    """All candidates of one strand, in libprimer3's list order (penalty, start descending, length)."""
    out = []
    for start in range(len(template)):
        for ln in range(s["min_size"], s["max_size"] + 1):
            r = cheap_record(s, template, mask, side, start, ln)
            if r:
                out.append(r)
    out.sort(key=lambda r: (r["penalty"], -r["start"], r["len"]))
    return out


def _thal_kw(s):
    return dict(mv=s["mv_conc"], dv=s["dv_conc"], dntp=s["dntp_conc"], dna_conc=s["dna_conc"])


def add_self(oracle, s, recs, nthreads=8):
    seqs = [r["seq"] for r in recs]
    if not seqs:
        return
    kw = _thal_kw(s)
    a = oracle.thal_batch(seqs, seqs, ANY, nthreads=nthreads, **kw)["tm"]
    e = oracle.thal_batch(seqs, seqs, END1, nthreads=nthreads, **kw)["tm"]
    h = oracle.thal_batch(seqs, None, HAIRPIN, nthreads=nthreads, **kw)["tm"]
    for r, x, y, z in zip(recs, a, e, h):
        r["self_any_th"], r["self_end_th"], r["hairpin_th"] = float(x), float(y), float(z)
        r["self_ok"] = (x <= s["max_self_any_th"] and y <= s["max_self_end_th"] and z <= s["max_hairpin_th"])


def design(oracle, settings, template, excluded=(), target=None, left=None, right=None, nthreads=8):
    """Returns the chosen pairs as dicts(left, right, pair_penalty, compl_any_th, compl_end_th,
    product_size, size_range).  left / right: given primer sequences (SEQUENCE_PRIMER /
    SEQUENCE_PRIMER_REVCOMP)."""
    s = {**DEFAULTS, **settings}
    mask = [0] * len(template)
    for a, l in excluded:
        for k in range(max(a, 0), min(a + l, len(template))):
            mask[k] = 1
    if target:
        for k in range(max(target[0], 0), min(target[0] + target[1], len(template))):
            mask[k] = 1
    up = template.upper()
    if left is not None:
        k = up.find(left.upper())
        r = cheap_record(s, template, mask, "L", k, len(left)) if k >= 0 else None
        L = [r] if r else []
    else:
        L = candidates(s, template, mask, "L")
    if right is not None:
        k = up.find(right.upper()[::-1].translate(_RC))
        r = cheap_record(s, template, mask, "R", k + len(right) - 1, len(right)) if k >= 0 else None
        R = [r] if r else []
    else:
        R = candidates(s, template, mask, "R")
    add_self(oracle, s, L, nthreads)
    add_self(oracle, s, R, nthreads)
    chosen, taken = [], set()
    for interval, (lo, hi) in enumerate(s["size_ranges"]):
        want = s["num_return"] - len(chosen)
        if want <= 0:
            break
        cand = []
        for li, l in enumerate(L):
            if not l["self_ok"]:
                continue
            for ri, r in enumerate(R):
                if not r["self_ok"] or (li, ri) in taken:
                    continue
                product = r["start"] - l["start"] + 1
                if product < lo or product > hi:
                    continue
                # sizes an earlier range covered were searched there already
                if any(a <= product <= b for a, b in s["size_ranges"][:interval]):
                    continue
                if math.fabs(l["tm"] - r["tm"]) > s["pair_max_diff_tm"]:
                    continue
                if target and target[1] > 0:
                    if not (l["start"] + l["len"] - 1 < target[0] and r["start"] - r["len"] + 1 >= target[0] + target[1]):
                        continue
                cand.append((li, ri, product))
        if not cand:
            continue
        a = [L[li]["seq"] for li, _, _ in cand]
        b = [R[ri]["seq"] for _, ri, _ in cand]
        kw = _thal_kw(s)
        any_ = oracle.thal_batch(a, b, ANY, nthreads=nthreads, **kw)["tm"]
        e1 = oracle.thal_batch(a, b, END1, nthreads=nthreads, **kw)["tm"]
        e2 = oracle.thal_batch(a, b, END2, nthreads=nthreads, **kw)["tm"]
        rows = []
        for (li, ri, product), x, y, z in zip(cand, any_, e1, e2):
            end = max(float(y), float(z))
            if x > s["pair_max_compl_any_th"] or end > s["pair_max_compl_end_th"]:
                continue
            l, r = L[li], R[ri]
            pq = l["penalty"] + r["penalty"]
            cm = (r["self_end_th"] + l["self_end_th"] + end) * 1.1 + r["self_any_th"] + l["self_any_th"] + float(x)
            rows.append(((pq, cm, -l["start"], r["start"], l["len"], r["len"]), li, ri, product, float(x), end))
        rows.sort(key=lambda t: t[0])
        for key, li, ri, product, x, end in rows[:want]:
            taken.add((li, ri))
            chosen.append(dict(left=L[li], right=R[ri], pair_penalty=key[0], compl_any_th=x, compl_end_th=end,
                               product_size=product, size_range=interval))
    return chosen
