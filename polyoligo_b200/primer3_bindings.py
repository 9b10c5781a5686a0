"""Drop-in for the parts of ``primer3.bindings`` the reference uses.

The reference reaches libprimer3's primer picker through
``primer3.bindings.designPrimers(seq_args)`` and ``primer3.setP3Globals(dict)``
(reference src/polyoligo/lib_primer3.py:479,544), with these sequence tags:
``SEQUENCE_TEMPLATE``, ``SEQUENCE_TARGET``, ``SEQUENCE_EXCLUDED_REGION`` (PCR /
CAPS / HRM, _lib_pcr.py:167-180) plus ``SEQUENCE_PRIMER`` or
``SEQUENCE_PRIMER_REVCOMP`` (KASP: the allele-specific primer is given and the
common primer is picked, _lib_kasp.py:567-590).  The global tags it sets are the
ones in src/polyoligo/data/PRIMER3_*.yaml plus ``PRIMER_NUM_RETURN``.

libprimer3 is third-party code that is not in the reference tree; what is
implemented here is the restatement described in DESIGN.md section 7 (parity
with the real library UNPINNED).  Candidate enumeration, sorting, the
fixed-primer search and every thermodynamic alignment run on the GPU; the
two-sided search (both primers picked) turns the best left candidates of every
template into fixed-primer tasks of one device call and merges their answers on
the host under an exactness bound (_search_two_sided_batch).

``designPrimers`` returns a primer3-py style dict (``PRIMER_LEFT_0_SEQUENCE``,
``PRIMER_PAIR_0_PENALTY``, ...), so ``lib_primer3.get_primers`` parses it
unchanged.  ``design_primers_batch`` takes many ``seq_args`` at once.
"""
import numpy as np

from . import _native
from .thermoanalysis import get_context

# libprimer3 defaults ("version 2" global settings) of the tags that are honoured
DEFAULT_GLOBALS = {
    "PRIMER_MIN_SIZE": 18, "PRIMER_OPT_SIZE": 20, "PRIMER_MAX_SIZE": 27,
    "PRIMER_MIN_TM": 57.0, "PRIMER_OPT_TM": 60.0, "PRIMER_MAX_TM": 63.0,
    "PRIMER_MIN_GC": 20.0, "PRIMER_MAX_GC": 80.0,
    "PRIMER_MAX_POLY_X": 5, "PRIMER_GC_CLAMP": 0, "PRIMER_MAX_END_GC": 5, "PRIMER_MAX_NS_ACCEPTED": 0,
    "PRIMER_SALT_MONOVALENT": 50.0, "PRIMER_SALT_DIVALENT": 1.5, "PRIMER_DNTP_CONC": 0.6, "PRIMER_DNA_CONC": 50.0,
    "PRIMER_MAX_SELF_ANY_TH": 47.0, "PRIMER_MAX_SELF_END_TH": 47.0, "PRIMER_MAX_HAIRPIN_TH": 47.0,
    "PRIMER_PAIR_MAX_COMPL_ANY_TH": 47.0, "PRIMER_PAIR_MAX_COMPL_END_TH": 47.0,
    "PRIMER_PAIR_MAX_DIFF_TM": 100.0,
    "PRIMER_WT_TM_GT": 1.0, "PRIMER_WT_TM_LT": 1.0, "PRIMER_WT_SIZE_LT": 1.0, "PRIMER_WT_SIZE_GT": 1.0,
    "PRIMER_PRODUCT_SIZE_RANGE": [[100, 300]],
    "PRIMER_NUM_RETURN": 5, "PRIMER_FIRST_BASE_INDEX": 0,
    "PRIMER_THERMODYNAMIC_OLIGO_ALIGNMENT": 1, "PRIMER_TM_FORMULA": 1, "PRIMER_SALT_CORRECTIONS": 1,
}
# tags that are accepted and have no effect on this path (the old dpal limits are unused when
# PRIMER_THERMODYNAMIC_OLIGO_ALIGNMENT = 1; internal oligos are never picked)
IGNORED_GLOBALS = {"PRIMER_PICK_INTERNAL_OLIGO", "PRIMER_INTERNAL_MAX_SELF_END", "PRIMER_INTERNAL_MAX_POLY_X",
                   "PRIMER_MAX_SELF_ANY", "PRIMER_MAX_SELF_END", "PRIMER_PAIR_MAX_COMPL_ANY",
                   "PRIMER_PAIR_MAX_COMPL_END"}

_globals = dict(DEFAULT_GLOBALS)
_p3_contexts = {}


def setP3Globals(global_args):
    """primer3.setP3Globals: updates the process-wide settings (values accumulate, as in primer3-py)."""
    for k, v in global_args.items():
        if k.startswith("SEQUENCE"):
            continue
        if k not in DEFAULT_GLOBALS and k not in IGNORED_GLOBALS:
            raise ValueError("primer3 tag %s is not supported by polyoligo_b200.primer3_bindings" % k)
        _globals[k] = v


def resetP3Globals():
    _globals.clear()
    _globals.update(DEFAULT_GLOBALS)


def settings_from_globals(g=None):
    g = _globals if g is None else {**DEFAULT_GLOBALS, **g}
    for tag, want in (("PRIMER_THERMODYNAMIC_OLIGO_ALIGNMENT", 1), ("PRIMER_TM_FORMULA", 1),
                      ("PRIMER_SALT_CORRECTIONS", 1), ("PRIMER_MAX_NS_ACCEPTED", 0), ("PRIMER_FIRST_BASE_INDEX", 0)):
        if int(g[tag]) != want:
            raise ValueError("%s = %r is not supported (only %d)" % (tag, g[tag], want))
    s = _native.P3Settings()
    s.min_size, s.opt_size, s.max_size = int(g["PRIMER_MIN_SIZE"]), int(g["PRIMER_OPT_SIZE"]), int(g["PRIMER_MAX_SIZE"])
    s.num_return = int(g["PRIMER_NUM_RETURN"])
    s.min_tm, s.opt_tm, s.max_tm = float(g["PRIMER_MIN_TM"]), float(g["PRIMER_OPT_TM"]), float(g["PRIMER_MAX_TM"])
    s.min_gc, s.max_gc = float(g["PRIMER_MIN_GC"]), float(g["PRIMER_MAX_GC"])
    s.max_self_any_th = float(g["PRIMER_MAX_SELF_ANY_TH"])
    s.max_self_end_th = float(g["PRIMER_MAX_SELF_END_TH"])
    s.max_hairpin_th = float(g["PRIMER_MAX_HAIRPIN_TH"])
    s.pair_max_compl_any_th = float(g["PRIMER_PAIR_MAX_COMPL_ANY_TH"])
    s.pair_max_compl_end_th = float(g["PRIMER_PAIR_MAX_COMPL_END_TH"])
    s.pair_max_diff_tm = float(g["PRIMER_PAIR_MAX_DIFF_TM"])
    s.wt_tm_gt, s.wt_tm_lt = float(g["PRIMER_WT_TM_GT"]), float(g["PRIMER_WT_TM_LT"])
    s.wt_size_lt, s.wt_size_gt = float(g["PRIMER_WT_SIZE_LT"]), float(g["PRIMER_WT_SIZE_GT"])
    s.mv_conc, s.dv_conc = float(g["PRIMER_SALT_MONOVALENT"]), float(g["PRIMER_SALT_DIVALENT"])
    s.dntp_conc, s.dna_conc = float(g["PRIMER_DNTP_CONC"]), float(g["PRIMER_DNA_CONC"])
    s.max_poly_x, s.gc_clamp, s.max_end_gc = int(g["PRIMER_MAX_POLY_X"]), int(g["PRIMER_GC_CLAMP"]), int(g["PRIMER_MAX_END_GC"])
    ranges = g["PRIMER_PRODUCT_SIZE_RANGE"]
    if ranges and not isinstance(ranges[0], (list, tuple)):
        ranges = [ranges]
    if len(ranges) > 8:
        raise ValueError("at most 8 PRIMER_PRODUCT_SIZE_RANGE entries are supported")
    s.n_size_ranges = len(ranges)
    for k, (lo, hi) in enumerate(ranges):
        s.size_lo[k], s.size_hi[k] = int(lo), int(hi)
    return s


_RC = str.maketrans("ACGTacgt", "TGCAtgca")


def _revcomp(s):
    return s[::-1].translate(_RC)


def _intervals(v):
    if v is None:
        return []
    v = list(v)
    if v and not isinstance(v[0], (list, tuple, np.ndarray)):
        v = [v]
    return [(int(a), int(b)) for a, b in v]


class _Job:
    """One designPrimers call, normalised."""

    def __init__(self, seq_args):
        self.template = seq_args["SEQUENCE_TEMPLATE"]
        n = len(self.template)
        mask = np.zeros(n, dtype=np.uint8)
        for a, l in _intervals(seq_args.get("SEQUENCE_EXCLUDED_REGION")):
            mask[max(a, 0):max(a + l, 0)] = 1
        inc = _intervals(seq_args.get("SEQUENCE_INCLUDED_REGION"))
        if inc:
            a, l = inc[0]
            mask[:max(a, 0)] = 1
            mask[max(a + l, 0):] = 1
        tg = _intervals(seq_args.get("SEQUENCE_TARGET"))
        if len(tg) > 1:
            raise ValueError("more than one SEQUENCE_TARGET is not supported")
        self.target = tg[0] if tg else (0, 0)
        if tg:
            mask[max(tg[0][0], 0):max(tg[0][0] + tg[0][1], 0)] = 1   # a primer may not overlap the target
        self.mask = mask
        self.error = None
        self.left = self.right = None     # (start, len) of a given primer, primer3 convention
        up = self.template.upper()
        lp, rp = seq_args.get("SEQUENCE_PRIMER"), seq_args.get("SEQUENCE_PRIMER_REVCOMP")
        if lp:
            k = up.find(lp.upper())
            if k < 0:
                self.error = "Specified left primer not in sequence"
            self.left = (k, len(lp))
        if rp:
            k = up.find(_revcomp(rp.upper()))
            if k < 0:
                self.error = "Specified right primer not in sequence"
            self.right = (k + len(rp) - 1, len(rp))


def _p3_context(settings, device=None):
    """Scoring context holding libprimer3's thal conditions (they differ from ThermoAnalysis()'s)."""
    base = get_context(device)
    key = (id(base), settings.mv_conc, settings.dv_conc, settings.dntp_conc, settings.dna_conc)
    ctx = _p3_contexts.get(key)
    if ctx is None:
        ctx = _native.Context(base.device)
        ctx.set_conditions(settings.mv_conc, settings.dv_conc, settings.dntp_conc, settings.dna_conc, 37.0, 30, 36)
        _p3_contexts[key] = ctx
    return ctx


# SantaLucia 1998 nearest-neighbour -dG37 in cal/mol (oligotm.c oligodg(), tm_method santalucia)
_NN_DG = {"AA": 1000, "AC": 1440, "AG": 1280, "AT": 880, "CA": 1450, "CC": 1840, "CG": 2170, "CT": 1280,
          "GA": 1300, "GC": 2240, "GG": 1840, "GT": 1440, "TA": 580, "TC": 1300, "TG": 1450, "TT": 1000}


def end_stability(seq):
    """libprimer3 end_oligodg(seq, 5): duplex-disruption dG (kcal/mol) of the last five 3' bases.
    primer3's manual pins the scale: GCGCG = 6.86, TATAT = 0.86."""
    s = seq[-5:].upper()
    if len(s) < 2:
        return 0.0
    dg = -1960                                   # initiation
    dg += sum(-50 for c in (s[0], s[-1]) if c in "AT")   # terminal AT penalty
    dg += sum(_NN_DG[s[k:k + 2]] for k in range(len(s) - 1))
    if len(s) % 2 == 0 and s == _revcomp(s):
        dg += -430                               # symmetry correction
    return dg / 1000.0


def _oligo_fields(side, i, rec, seq):
    p = "PRIMER_%s_%d" % (side, i)
    return {p: (int(rec["start"]), int(rec["len"])), p + "_SEQUENCE": seq, p + "_PENALTY": float(rec["penalty"]),
            p + "_TM": float(rec["tm"]), p + "_GC_PERCENT": float(rec["gc_percent"]),
            p + "_END_STABILITY": end_stability(seq),
            p + "_SELF_ANY_TH": float(rec["self_any_th"]), p + "_SELF_END_TH": float(rec["self_end_th"]),
            p + "_HAIRPIN_TH": float(rec["hairpin_th"])}


def _result_dict(pairs):
    """pairs: list of (left record, right record, pair penalty, compl_any, compl_end, product size)."""
    out = {"PRIMER_LEFT_NUM_RETURNED": len(pairs), "PRIMER_RIGHT_NUM_RETURNED": len(pairs),
           "PRIMER_INTERNAL_NUM_RETURNED": 0, "PRIMER_PAIR_NUM_RETURNED": len(pairs)}
    if not pairs:
        return out
    seqs = _native.unpack(np.array([p[0]["seq"] for p in pairs] + [p[1]["seq"] for p in pairs],
                                   dtype=_native.OLIGO_DTYPE))
    n = len(pairs)
    for i, (l, r, pq, ca, ce, size) in enumerate(pairs):
        p = "PRIMER_PAIR_%d" % i
        out[p + "_PENALTY"] = float(pq)
        out.update(_oligo_fields("LEFT", i, l, seqs[i]))
        out.update(_oligo_fields("RIGHT", i, r, seqs[n + i]))
        out[p + "_COMPL_ANY_TH"] = float(ca)
        out[p + "_COMPL_END_TH"] = float(ce)
        out[p + "_PRODUCT_SIZE"] = int(size)
    return out


# ---- two-sided search (both primers picked) ------------------------------------------------
def _size_ok(s, interval, product):
    ok = (product >= s.size_lo[interval]) & (product <= s.size_hi[interval])
    for e in range(interval):
        ok &= ~((product >= s.size_lo[e]) & (product <= s.size_hi[e]))
    return ok


def _ensure_self(ctx, s, recs, idx):
    """Lazily evaluates the self alignments of recs[idx] (in place)."""
    idx = np.unique(idx[np.isnan(recs["self_any_th"][idx])])
    if idx.size == 0:
        return
    seq = np.ascontiguousarray(recs["seq"][idx])
    recs["self_any_th"][idx] = ctx.thal_batch(_native.THAL_ANY, seq)["tm"]
    recs["self_end_th"][idx] = ctx.thal_batch(_native.THAL_END1, seq)["tm"]
    recs["hairpin_th"][idx] = ctx.thal_batch(_native.THAL_HAIRPIN, seq)["tm"]


def _self_ok(s, recs, idx):
    return ((recs["self_any_th"][idx] <= s.max_self_any_th) & (recs["self_end_th"][idx] <= s.max_self_end_th) &
            (recs["hairpin_th"][idx] <= s.max_hairpin_th))


def _search_two_sided(ctx, s, L, R, target):
    """Best num_return (left, right) pairs in libprimer3's order.  L / R are the sorted candidate
    records.  The sorted lists are explored corner by corner: every pair outside the top-M x top-M
    corner has a penalty of at least min(L[M] + R[0], L[0] + R[M]), so pairs below that bound
    are final."""
    accepted = []
    if len(L) == 0 or len(R) == 0:
        return accepted
    tar_s, tar_l = target
    for interval in range(s.n_size_ranges):
        want = s.num_return - len(accepted)
        if want <= 0:
            break
        found = []
        M = 64
        done_pairs = {}
        while True:
            ml, mr = min(M, len(L)), min(M, len(R))
            bound = np.inf
            if ml < len(L):
                bound = min(bound, L["penalty"][ml] + R["penalty"][0])
            if mr < len(R):
                bound = min(bound, L["penalty"][0] + R["penalty"][mr])
            li, ri = np.meshgrid(np.arange(ml), np.arange(mr), indexing="ij")
            li, ri = li.ravel(), ri.ravel()
            product = R["start"][ri] - L["start"][li] + 1
            ok = _size_ok(s, interval, product)
            ok &= ~(np.abs(L["tm"][li] - R["tm"][ri]) > s.pair_max_diff_tm)
            if tar_l > 0:
                ok &= (L["start"][li] + L["len"][li] - 1 < tar_s) & (R["start"][ri] - R["len"][ri] + 1 >= tar_s + tar_l)
            li, ri, product = li[ok], ri[ok], product[ok]
            pq = L["penalty"][li] + R["penalty"][ri]
            order = np.lexsort((ri, li, pq))
            li, ri, product, pq = li[order], ri[order], product[order], pq[order]
            n_inside = int(np.count_nonzero(pq < bound))   # sorted: the safe pairs are a prefix
            found = []
            pos, complete = 0, False
            while pos < n_inside and not complete:
                end = min(pos + 256, n_inside)
                cl, cr = li[pos:end], ri[pos:end]
                _ensure_self(ctx, s, L, cl)
                _ensure_self(ctx, s, R, cr)
                todo = [k for k in range(len(cl)) if (int(cl[k]), int(cr[k])) not in done_pairs]
                if todo:
                    a = np.ascontiguousarray(L["seq"][cl[todo]])
                    b = np.ascontiguousarray(R["seq"][cr[todo]])
                    any_ = ctx.thal_batch(_native.THAL_ANY, a, b)["tm"]
                    end_ = np.maximum(ctx.thal_batch(_native.THAL_END1, a, b)["tm"],
                                      ctx.thal_batch(_native.THAL_END2, a, b)["tm"])
                    for k, x, y in zip(todo, any_, end_):
                        done_pairs[(int(cl[k]), int(cr[k]))] = (float(x), float(y))
                sok_l, sok_r = _self_ok(s, L, cl), _self_ok(s, R, cr)
                for k in range(len(cl)):
                    # enough pairs, and the next one is strictly worse than the last accepted: final
                    if len(found) >= want and pq[pos + k] > found[-1][2]:
                        complete = True
                        break
                    ca, ce = done_pairs[(int(cl[k]), int(cr[k]))]
                    if sok_l[k] and sok_r[k] and ca <= s.pair_max_compl_any_th and ce <= s.pair_max_compl_end_th:
                        found.append((int(cl[k]), int(cr[k]), float(pq[pos + k]), ca, ce, int(product[pos + k])))
                pos = end
            if complete or (ml == len(L) and mr == len(R)):
                break
            M *= 2

        def key(f):
            l, r = L[f[0]], R[f[1]]
            cm = (r["self_end_th"] + l["self_end_th"] + f[4]) * 1.1 + r["self_any_th"] + l["self_any_th"] + f[3]
            return (f[2], cm, -int(l["start"]), int(r["start"]), int(l["len"]), int(r["len"]))
        found.sort(key=key)
        accepted += [(L[f[0]].copy(), R[f[1]].copy(), f[2], f[3], f[4], f[5]) for f in found[:want]]
    return accepted


def _search_two_sided_batch(ctx, s, templates, masks, targets, lefts, right_min_penalty, per_left=None):
    """Both primers picked, for many templates at once, on top of the device's fixed-primer search:
    every one of the best M left candidates of a template becomes a 'left primer given' task (one
    device call for all templates), and the per-left answers are merged.  The merge is final when
    the wanted number of pairs lies strictly below every bound on what has not been seen:
      * left candidates not yet covered: pair penalty >= L[M].penalty + R[0].penalty (else M grows);
      * with ``per_left`` < num_return each task returns only its best ``per_left`` pairs; a task
        that filled its quota may hold more pairs, all at or above the penalty of its last one.
    A template that cannot be settled under the per-left quota is returned as None (the caller runs
    it again without the quota).  Size ranges are honoured in order: a later range only fills what
    the earlier ones left open."""
    n = len(templates)
    want_total = s.num_return
    quota = want_total if per_left is None else min(per_left, want_total)
    dev = _native.P3Settings.from_buffer_copy(s)
    dev.num_return = quota
    calls = []                        # (fixed, pairs) of every device call: rows are gathered from here at the end
    acc = [[] for _ in range(n)]      # per template: (call index, row slice) into that call's key arrays
    keys = []                         # per call: dict of flat per-row key arrays (task index K, rank R, sort keys)
    full = [[] for _ in range(n)]     # per template: (size range, penalty) of the last pair of every full task
    covered = [0] * n
    result = [None] * n
    given_up = [False] * n
    M = 64
    while any(result[t] is None and not given_up[t] for t in range(n)):
        tasks, owner = [], []
        for t in range(n):
            if result[t] is not None or given_up[t]:
                continue
            hi = min(M, len(lefts[t]))
            for li in range(covered[t], hi):
                tasks.append((t, 1, int(lefts[t]["start"][li]), int(lefts[t]["len"][li])))
                owner.append(t)
            covered[t] = hi
        if tasks:
            tasks = np.array(tasks, dtype=_native.P3_TASK_DTYPE)
            owner = np.array(owner)
            fixed, pairs, n_pairs = ctx.p3_design_fixed(dev, templates, tasks, masks, targets)
            K, R = np.nonzero(np.arange(quota)[None, :] < n_pairs[:, None])
            oth = pairs["other"]
            # only the scalar sort keys are gathered per row; whole records only for the chosen rows
            kd = {"K": K, "R": R, "range": pairs["size_range"][K, R], "pq": pairs["pair_penalty"][K, R],
                  "cm": (oth["self_end_th"][K, R] + fixed["self_end_th"][K] + pairs["compl_end_th"][K, R]) * 1.1 +
                        oth["self_any_th"][K, R] + fixed["self_any_th"][K] + pairs["compl_any_th"][K, R],
                  "nls": -fixed["start"][K].astype(np.int64), "rs": oth["start"][K, R],
                  "ll": fixed["len"][K], "rl": oth["len"][K, R]}
            calls.append((fixed, pairs))
            keys.append(kd)
            cuts = np.searchsorted(owner[K], np.arange(n + 1))      # owner[K] is non-decreasing
            for t in range(n):
                if cuts[t + 1] > cuts[t]:
                    acc[t].append((len(calls) - 1, slice(cuts[t], cuts[t + 1])))
            if quota < want_total:
                for k in np.flatnonzero(n_pairs == quota):
                    last = pairs[k, quota - 1]
                    full[owner[k]].append((int(last["size_range"]), float(last["pair_penalty"])))
        for t in range(n):
            if result[t] is not None or given_up[t]:
                continue
            all_left = covered[t] >= len(lefts[t])
            bound = np.inf if all_left else float(lefts[t]["penalty"][covered[t]]) + right_min_penalty[t]
            col = {f: (np.concatenate([keys[c][f][sl] for c, sl in acc[t]]) if acc[t] else np.zeros(0))
                   for f in ("range", "pq", "cm", "nls", "rs", "ll", "rl")}
            src = np.concatenate([np.full(sl.stop - sl.start, c) for c, sl in acc[t]]) if acc[t] else np.zeros(0, int)
            row = np.concatenate([np.arange(sl.start, sl.stop) for c, sl in acc[t]]) if acc[t] else np.zeros(0, int)
            # pair order inside a size range: penalty, compl_measure, left start descending, right start
            # ascending, left length, right length (np.lexsort: the last key is the primary one)
            order = np.lexsort((col["rl"], col["ll"], col["rs"], col["nls"], col["cm"], col["pq"], col["range"]))
            rng_sorted, pq_sorted = col["range"][order], col["pq"][order]
            chosen, complete = [], True
            for interval in range(s.n_size_ranges):
                want = want_total - len(chosen)
                if want <= 0:
                    break
                idx = np.flatnonzero(rng_sorted == interval)
                # tasks that filled their quota in an EARLIER range never showed this range at all
                blocked = any(k < interval for k, _ in full[t])
                quota_bound = min([q for k, q in full[t] if k == interval], default=np.inf)
                if all_left and not blocked and quota_bound == np.inf:
                    chosen += list(idx[:want])                        # everything there is
                elif not blocked and len(idx) >= want and pq_sorted[idx[want - 1]] < min(bound, quota_bound):
                    chosen += list(idx[:want])
                else:
                    complete = False
                    if blocked or (len(idx) >= want and pq_sorted[idx[want - 1]] >= quota_bound) or all_left:
                        given_up[t] = True            # more left candidates cannot settle it: the quota is in the way
                    break
            if complete:
                rows_ = []
                for i in chosen[:want_total]:
                    c, rr = int(src[order[i]]), int(row[order[i]])
                    fx, pr = calls[c]
                    k, r = int(keys[c]["K"][rr]), int(keys[c]["R"][rr])
                    p = pr[k, r]
                    rows_.append((fx[k].copy(), p["other"].copy(), p["pair_penalty"], p["compl_any_th"],
                                  p["compl_end_th"], p["product_size"]))
                result[t] = rows_
        M *= 4
    return result


# ---- public entry points ------------------------------------------------------------------
def design_primers_batch(seq_args_list, global_args=None, device=None):
    """designPrimers for many seq_args at once; returns one primer3-style dict per call."""
    s = settings_from_globals(global_args)
    ctx = _p3_context(s, device)
    jobs = [_Job(a) for a in seq_args_list]
    out = [None] * len(jobs)
    fixed_ix = [k for k, j in enumerate(jobs) if not j.error and ((j.left is None) != (j.right is None))]
    other_ix = [k for k, j in enumerate(jobs) if not j.error and k not in set(fixed_ix)]
    for k, j in enumerate(jobs):
        if j.error:
            out[k] = {"PRIMER_ERROR": j.error}
    if fixed_ix:
        # templates shared by several calls (KASP: every marker primer of a marker) are uploaded once
        sets, set_of = {}, []
        for k in fixed_ix:
            j = jobs[k]
            key = (j.template, j.mask.tobytes(), j.target)
            set_of.append(sets.setdefault(key, len(sets)))
        keys = list(sets.keys())
        templates = [key[0] for key in keys]
        masks = [np.frombuffer(key[1], dtype=np.uint8) for key in keys]
        targets = np.array([key[2] for key in keys], dtype=np.int32)
        tasks = np.zeros(len(fixed_ix), dtype=_native.P3_TASK_DTYPE)
        for t, k in enumerate(fixed_ix):
            j = jobs[k]
            side, (start, ln) = (1, j.left) if j.left is not None else (2, j.right)
            tasks[t] = (set_of[t], side, start, ln)
        fixed, pairs, n_pairs = ctx.p3_design_fixed(s, templates, tasks, masks, targets)
        for t, k in enumerate(fixed_ix):
            rows = []
            for p in pairs[t, :n_pairs[t]]:
                l, r = (fixed[t], p["other"]) if tasks[t]["side"] == 1 else (p["other"], fixed[t])
                rows.append((l, r, p["pair_penalty"], p["compl_any_th"], p["compl_end_th"], p["product_size"]))
            out[k] = _result_dict(rows)
    if other_ix:
        templates = [jobs[k].template for k in other_ix]
        masks = [jobs[k].mask for k in other_ix]
        list_off, recs = ctx.p3_candidates(s, templates, masks)
        free, lefts, rmin = [], [], []
        for t, k in enumerate(other_ix):
            j = jobs[k]
            L = recs[list_off[2 * t]:list_off[2 * t + 1]]
            R = recs[list_off[2 * t + 1]:list_off[2 * t + 2]]
            if j.left is not None:     # both primers given: libprimer3 evaluates that single pair
                L = L[(L["start"] == j.left[0]) & (L["len"] == j.left[1])].copy()
                R = R[(R["start"] == j.right[0]) & (R["len"] == j.right[1])].copy()
                out[k] = _result_dict(_search_two_sided(ctx, s, L, R, j.target))
            elif len(L) == 0 or len(R) == 0:
                out[k] = _result_dict([])
            else:
                # left candidates that no right candidate can reach with an allowed product size (or
                # that do not lie before the target) can never be in a pair: drop them up front
                rs = np.sort(R["start"].astype(np.int64))
                ls = L["start"].astype(np.int64)
                lo_ = min(s.size_lo[:s.n_size_ranges]) if s.n_size_ranges else 0
                hi_ = max(s.size_hi[:s.n_size_ranges]) if s.n_size_ranges else -1
                reach = np.searchsorted(rs, ls + hi_ - 1, side="right") > np.searchsorted(rs, ls + lo_ - 1, side="left")
                if j.target[1] > 0:
                    reach &= (ls + L["len"] - 1 < j.target[0]) & (ls + hi_ - 1 >= j.target[0] + j.target[1])
                L = L[reach]
                if len(L) == 0:
                    out[k] = _result_dict([])
                    continue
                free.append(t)
                lefts.append(L)
                rmin.append(float(R["penalty"][0]))
        if free:
            targets = np.array([jobs[other_ix[t]].target for t in free], dtype=np.int32)
            sub_t, sub_m = [templates[t] for t in free], [masks[t] for t in free]
            # first with a small per-left quota (the best pairs of a template rarely share a left primer
            # more than a few times), then without it for the templates the quota could not settle
            res = _search_two_sided_batch(ctx, s, sub_t, sub_m, targets, lefts, rmin, per_left=16)
            redo = [i for i, r_ in enumerate(res) if r_ is None]
            if redo:
                again = _search_two_sided_batch(ctx, s, [sub_t[i] for i in redo], [sub_m[i] for i in redo],
                                                targets[redo], [lefts[i] for i in redo], [rmin[i] for i in redo])
                for i, r_ in zip(redo, again):
                    res[i] = r_
            for t, rows_ in zip(free, res):
                out[other_ix[t]] = _result_dict(rows_)
    return out


def designPrimers(seq_args, global_args=None, misprime_lib=None, mishyb_lib=None, debug=False):
    """primer3.bindings.designPrimers(seq_args, global_args=None, ...)."""
    if misprime_lib or mishyb_lib:
        raise ValueError("mispriming libraries are not supported")
    if global_args:
        setP3Globals(global_args)
    return design_primers_batch([seq_args])[0]
