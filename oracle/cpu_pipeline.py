"""CPU ORACLE of the config-5 array pipeline (test infrastructure and bench.py's CPU baseline; NOT
product code, nothing under polyoligo_b200/ imports it).

``CpuContext`` offers the calls ``polyoligo_b200.kasp_pipeline`` makes on a scoring context --
kasp_enumerate, oligo_props_batch, p3_design_fixed, annotate_mutations, rank_pairs -- each
restated for the host: the thermodynamics come from the C oracle (po_oracle.c, all host cores),
everything else is plain Python / numpy written from the reference's files:

* kasp_enumerate     reference src/polyoligo/_lib_kasp.py:295-337 (get_marker_primers slicing)
* oligo_props_batch  reference src/polyoligo/lib_primer3.py:56-97 (calc_thermals + is_valid)
* p3_design_fixed    the picker semantics of DESIGN.md section 7, with libprimer3's LAZY
                     evaluation order (alignments only for oligos that reach a candidate pair);
                     oracle/p3_oracle.py is the eager brute-force statement of the same rules
* rank_pairs         reference _lib_kasp.py:75-117 (reporter-tailed heterodimers), scoring.py:105-151
                     (score_kasp), lib_primer3.py:425-431 + _lib_kasp.py:278-291 (classify, prune)

Records travel in the product's array formats (polyoligo_b200._native dtypes, pack / unpack).
"""
import os

import numpy as np

from polyoligo_b200 import _native

from . import p3_oracle
from .oracle import Oracle, ANY, END1, END2, HAIRPIN
from .polyoligo_oracle import annotate_mutations as _annotate_one

_RC = str.maketrans("ACGT", "TGCA")
_LET = "ACGT"


def _revcomp(s):
    return s[::-1].translate(_RC)


def _words_to_str(words, n):
    return "".join(_LET[(int(words[p >> 4]) >> (2 * (p & 15))) & 3] for p in range(n))


class CpuContext:
    def __init__(self, nthreads=None):
        self.orc = Oracle()
        self.nthreads = nthreads or os.cpu_count() or 1
        self.thal_items = 0
        self.device = -1

    # ---- instrumentation the bench reads -------------------------------------------------------
    def thal_item_count(self):
        return self.thal_items

    def launch_count(self):
        return 0

    def _thal(self, a, b, type_, **kw):
        self.thal_items += len(a)
        return self.orc.thal_batch(a, b, type_, nthreads=self.nthreads, **kw)["tm"]

    # ---- a3 --------------------------------------------------------------------------------------
    def kasp_enumerate(self, markers, min_size, max_size):
        n = len(markers)
        slots = 2 * (max_size - min_size + 1)
        ref, alt = [], []
        for mk in markers:
            seq = _words_to_str(mk["tmpl"], int(mk["tmpl_len"]))
            seq_rc = _revcomp(seq)
            ref_len, alt_len = int(mk["ref_len"]), int(mk["alt_len"])
            alt_a = _words_to_str(mk["alt"], alt_len)
            d = ref_len - alt_len
            for padding in range(min_size, max_size + 1):
                x2 = int(mk["marker_pos"]) + ref_len
                x1 = x2 - padding - abs(d)
                y2 = int(mk["marker_pos_rc"])
                y1 = y2 - padding - abs(d) - ref_len
                for direction in "FR":
                    cseq = seq[x1:x2] if direction == "F" else seq_rc[y1:y2]
                    calt = alt_a if direction == "F" else _revcomp(alt_a)
                    pos_a = len(cseq) - ref_len
                    if pos_a < 0 or not cseq:
                        ref.append("")
                        alt.append("")
                        continue
                    palt = cseq[:pos_a] + calt
                    ref.append(cseq[d:] if d > 0 else cseq)
                    alt.append(palt if d > 0 else palt[-d:] if d < 0 else palt)
        return (_native.pack(ref).reshape(n, slots), _native.pack(alt).reshape(n, slots))

    # ---- a1 / a2 -----------------------------------------------------------------------------------
    def oligo_props_batch(self, oligos, valid=None):
        oligos = np.ascontiguousarray(oligos, dtype=_native.OLIGO_DTYPE)
        seqs = _native.unpack(oligos)
        out = np.zeros(len(seqs), dtype=_native.PROPS_DTYPE)
        ok = [k for k, s in enumerate(seqs) if s]
        good = [seqs[k] for k in ok]
        if good:
            out["tm"][ok] = self.orc.tm_batch(good, nthreads=self.nthreads)
            out["hairpin_tm"][ok] = self._thal(good, None, HAIRPIN)
            out["homodimer_tm"][ok] = self._thal(good, good, ANY)
        for k, s in enumerate(seqs):
            out["length"][k] = len(s)
            out["gc_content"][k] = 100.0 * (s.count("G") + s.count("C")) / len(s) if s else 0.0
            if not s:
                out["flags"][k] = _native.ITEM_SKIPPED
                continue
            if valid is None:
                continue
            r = out[k]
            good_ = (valid.min_size <= len(s) <= valid.max_size and valid.min_gc <= r["gc_content"] <= valid.max_gc
                     and valid.min_tm <= r["tm"] <= valid.max_tm and r["hairpin_tm"] < r["tm"] - valid.tm_delta
                     and r["homodimer_tm"] < r["tm"] - valid.tm_delta)
            if good_:
                out["flags"][k] = _native.PRIMER_VALID
        return out

    # ---- a5 (+ the validity of the picked primers, a6) -----------------------------------------------
    def p3_design_fixed(self, settings, templates, tasks, masks=None, targets=None, valid=None):
        s = {k: getattr(settings, k) for k, _ in settings._fields_ if k not in ("size_lo", "size_hi", "n_size_ranges")}
        s["size_ranges"] = tuple((settings.size_lo[k], settings.size_hi[k]) for k in range(settings.n_size_ranges))
        s = {**p3_oracle.DEFAULTS, **s}
        kw = dict(mv=s["mv_conc"], dv=s["dv_conc"], dntp=s["dntp_conc"], dna_conc=s["dna_conc"])
        blob, off = templates
        blob = np.asarray(blob, dtype=np.uint8)
        m = None if masks is None else np.asarray(masks, dtype=np.uint8).reshape(-1)
        tasks = np.asarray(tasks, dtype=_native.P3_TASK_DTYPE)
        nt, depth = len(tasks), settings.num_return
        pairs = np.zeros((nt, depth), dtype=_native.P3_PAIR_DTYPE)
        n_pairs = np.zeros(nt, dtype=np.int32)
        fixed = np.zeros(nt, dtype=_native.P3_OLIGO_DTYPE)
        lists = {}      # (set, side) -> sorted candidate dicts, shared by the tasks of a template
        for t, tk in enumerate(tasks):
            k = int(tk["set"])
            tmpl = blob[off[k]:off[k + 1]].tobytes().decode("ascii")
            mask = [0] * len(tmpl) if m is None else list(m[off[k]:off[k + 1]])
            side = "L" if tk["side"] == 1 else "R"
            # designPrimers locates the given primer by string search: first occurrence
            # (reference lib_primer3.py:479 passes the sequence only)
            st, ln = int(tk["start"]), int(tk["len"])
            lo = st if side == "L" else st - ln + 1
            if 0 <= lo and lo + ln <= len(tmpl):
                first = tmpl.upper().find(tmpl[lo:lo + ln].upper())
                st = first if side == "L" else first + ln - 1
            f = p3_oracle.cheap_record(s, tmpl, mask, side, st, ln)
            if f is None:
                continue
            p3_oracle.add_self(self.orc, s, [f], self.nthreads)
            self.thal_items += 3
            if not f["self_ok"]:
                continue
            oside = "R" if side == "L" else "L"
            key = (k, oside)
            if key not in lists:
                lists[key] = p3_oracle.candidates(s, tmpl, mask, oside)
            cand = lists[key]
            found = []
            for interval, (lo, hi) in enumerate(s["size_ranges"]):
                if len(found) >= depth:
                    break
                rows, pos = [], 0
                while pos < len(cand):
                    # the next cheap-valid candidates in list order, a small batch at a time
                    batch = []
                    while pos < len(cand) and len(batch) < 16:
                        c = cand[pos]
                        pos += 1
                        product = c["start"] - f["start"] + 1 if side == "L" else f["start"] - c["start"] + 1
                        if product < lo or product > hi:
                            continue
                        if any(a <= product <= b for a, b in s["size_ranges"][:interval]):
                            continue
                        if abs(f["tm"] - c["tm"]) > s["pair_max_diff_tm"] or c.get("self_ok") is False:
                            continue
                        batch.append((c, product))
                    if not batch:
                        continue
                    need_self = [c for c, _ in batch if "self_ok" not in c]
                    if need_self:
                        p3_oracle.add_self(self.orc, s, need_self, self.nthreads)
                        self.thal_items += 3 * len(need_self)
                    live = [(c, p) for c, p in batch if c["self_ok"]]
                    if live:
                        a = [f["seq"] if side == "L" else c["seq"] for c, _ in live]
                        b = [c["seq"] if side == "L" else f["seq"] for c, _ in live]
                        any_ = self._thal(a, b, ANY, **kw)
                        end_ = np.maximum(self._thal(a, b, END1, **kw), self._thal(a, b, END2, **kw))
                        for (c, product), x, y in zip(live, any_, end_):
                            if x > s["pair_max_compl_any_th"] or y > s["pair_max_compl_end_th"]:
                                continue
                            l, r = (f, c) if side == "L" else (c, f)
                            pq = l["penalty"] + r["penalty"]
                            cm = (r["self_end_th"] + l["self_end_th"] + float(y)) * 1.1 + r["self_any_th"] + \
                                l["self_any_th"] + float(x)
                            rows.append(((pq, cm, -l["start"], r["start"], l["len"], r["len"]), c, product,
                                         float(x), float(y)))
                    want = depth - len(found)
                    # enough pairs, and everything still ahead is strictly worse than the last kept one
                    if len(rows) >= want:
                        rows.sort(key=lambda z: z[0])
                        if pos >= len(cand):
                            break
                        nxt = cand[pos]          # the list is sorted: its penalty bounds everything ahead
                        ahead = (f["penalty"] + nxt["penalty"]) if side == "L" else (nxt["penalty"] + f["penalty"])
                        if ahead > rows[want - 1][0][0]:
                            break
                rows.sort(key=lambda z: z[0])
                found += [(interval,) + r for r in rows[:depth - len(found)]]
            rec = self._record(f, k, _native.P3_OK | _native.P3_SELF_DONE)
            fixed[t] = rec
            for i, (interval, key_, c, product, x, y) in enumerate(found):
                pairs[t, i]["other"] = self._record(c, k, _native.P3_OK | _native.P3_SELF_DONE)
                pairs[t, i]["pair_penalty"], pairs[t, i]["compl_any_th"], pairs[t, i]["compl_end_th"] = key_[0], x, y
                pairs[t, i]["product_size"], pairs[t, i]["size_range"] = product, interval
            n_pairs[t] = len(found)
        if valid is not None and n_pairs.sum():
            t_idx, r_idx = np.nonzero(np.arange(depth)[None, :] < n_pairs[:, None])
            oth = pairs["other"][t_idx, r_idx]
            seqs = _native.unpack(np.ascontiguousarray(oth["seq"]))
            uniq = sorted(set(seqs))
            props = self.oligo_props_batch(_native.pack(uniq), valid)
            okmap = {q: bool(p["flags"] & _native.PRIMER_VALID) for q, p in zip(uniq, props)}
            flags = pairs["other"]["flags"]
            for ti, ri, q in zip(t_idx, r_idx, seqs):
                if okmap[q]:
                    flags[ti, ri] |= _native.P3_PRIMER_VALID
        return fixed, pairs, n_pairs

    @staticmethod
    def _record(c, set_id, flags):
        rec = np.zeros((), dtype=_native.P3_OLIGO_DTYPE)
        rec["seq"] = _native.pack([c["seq"]])[0]
        for f in ("tm", "gc_percent", "penalty", "self_any_th", "self_end_th", "hairpin_th"):
            rec[f] = c[f]
        rec["start"], rec["len"], rec["set"], rec["flags"] = c["start"], c["len"], set_id, flags
        return rec

    # ---- 8f-3 ----------------------------------------------------------------------------------------
    def annotate_mutations(self, mut_pos, mut_aaf, mut_indel, mut_seg, pair_seg, primer_start, primer_stop,
                           n_primers):
        n_pairs = int(pair_seg[-1])
        ps = np.asarray(primer_start).reshape(n_pairs, 3)
        pe = np.asarray(primer_stop).reshape(n_pairs, 3)
        aaf = np.zeros((n_pairs, 3))
        indel = np.zeros(n_pairs, dtype=np.int32)
        for m in range(len(mut_seg) - 1):
            muts = [(int(mut_pos[i]), float(mut_aaf[i]), int(mut_indel[i])) for i in range(mut_seg[m], mut_seg[m + 1])]
            for k in range(pair_seg[m], pair_seg[m + 1]):
                windows = [(int(ps[k, c]), int(pe[k, c])) for c in range(n_primers)]
                a, ind = _annotate_one(muts, windows, (int(ps[k, 0]), int(pe[k, 1])))
                aaf[k, :n_primers] = a
                indel[k] = ind
        return aaf, indel

    # ---- a8 / a9 (KASP) --------------------------------------------------------------------------------
    def rank_pairs(self, assay, f, r, seg, n_offtargets, max_aaf, max_indel, tm_delta, top_n, a=None, direction=None,
                   hrm_delta=None, reporters=None):
        assert assay == _native.ASSAY_KASP, "the CPU pipeline oracle covers the KASP assay"
        fs, rs, as_ = _native.unpack(f), _native.unpack(r), _native.unpack(a)
        n = len(fs)
        tm = np.stack([self.orc.tm_batch(x, nthreads=self.nthreads) if n else np.zeros(0) for x in (fs, rs, as_)], axis=1)
        dye1, dye2 = reporters
        ft, rt, at, first, second = [], [], [], [], []
        for k in range(n):
            marker_is_f = direction[k] == 0
            tf = (dye1 if marker_is_f else "") + fs[k]       # add_reporter_dyes: marker primer dye1, ALT dye2
            tr = ("" if marker_is_f else dye1) + rs[k]
            ta = dye2 + as_[k]
            ft.append(tf)
            rt.append(tr)
            first.append(ta if marker_is_f else tf)          # alt pair: (A, R) for F markers, (F, A) for R markers
            second.append(tr if marker_is_f else ta)
        het_ref = self._thal(ft, rt, ANY) if n else np.zeros(0)
        het_alt = self._thal(first, second, ANY) if n else np.zeros(0)
        goodness = np.zeros(n, dtype=np.int32)
        qbits = np.zeros(n, dtype=np.int32)
        aaf = np.asarray(max_aaf).reshape(n, 3)
        for k in range(n):
            t_ref, t_alt = het_ref[k] - tm_delta, het_alt[k] - tm_delta
            lim_f, lim_r = tm[k, 0] - tm_delta, tm[k, 1] - tm_delta
            ref_dimer = not (t_ref <= lim_f and t_ref <= lim_r)
            alt_dimer = not (t_alt <= lim_f and t_alt <= lim_r)
            score, bits = 0, 0
            tms = np.array([tm[k, 0], tm[k, 1], tm[k, 2]])
            if np.sum(np.abs(tms - np.mean(tms))) <= 5:
                score += 1
            else:
                bits |= _native.Q_BITS["t"]
            if n_offtargets[k] == 0:
                score += 3
            else:
                bits |= _native.Q_BITS["O"]
            if not ref_dimer and not alt_dimer:
                score += 1
            else:
                bits |= _native.Q_BITS["d"]
            if np.all(aaf[k] < 0.1):
                score += 2
                if np.all(aaf[k] == 0):
                    score += 1
                else:
                    bits |= _native.Q_BITS["m"]
            else:
                bits |= _native.Q_BITS["M"]
            if max_indel[k] < 50:
                score += 1
                if max_indel[k] == 0:
                    score += 1
                else:
                    bits |= _native.Q_BITS["i"]
            else:
                bits |= _native.Q_BITS["I"]
            if ref_dimer:
                bits |= _native.Q_REF_DIMER
            if alt_dimer:
                bits |= _native.Q_ALT_DIMER
            goodness[k], qbits[k] = score, bits
        nm = len(seg) - 1
        sel = np.full((nm, top_n), -1, dtype=np.int64)
        n_sel = np.zeros(nm, dtype=np.int32)
        for m in range(nm):
            idx = np.arange(seg[m], seg[m + 1])
            order = idx[np.argsort(-goodness[idx], kind="stable")][:top_n]    # classes descending, insertion order inside
            sel[m, :len(order)] = order
            n_sel[m] = len(order)
        return {"goodness": goodness, "qbits": qbits, "tm": tm, "het_tm": np.stack([het_ref, het_alt], axis=1),
                "sel": sel, "n_sel": n_sel}
