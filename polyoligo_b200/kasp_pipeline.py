"""Array-level KASP design pipeline: the design part of reference
src/polyoligo/_lib_kasp.py:671-743 (main) for 10^4..10^5 markers per call, BASELINE config 5.

``kasp.design_markers`` makes the same decisions with the reference's objects (Primer /
PrimerPair / PCR) and is what the fixtures pin; this module is the throughput path: every
stage works on flat numpy arrays and batched C-ABI calls, nothing is per-marker Python.

    a3  marker-primer enumeration          po_kasp_enumerate_batch
    a1/a2 validity of REF / ALT primers      po_oligo_props_batch
    a5  common primer for every valid marker primer, per search mask
                                             po_p3_design_fixed_batch
    a6  validity of the picked common primers + round-robin merge
                                             po_oligo_props_batch + a sort
    a8/a9 reporter-tailed heterodimers, mutation annotation, goodness, top-n
                                             po_annotate_mutations_batch, po_rank_pairs_batch

Scope notes: templates are plain ACGT of one common length (<= 512), alleles are SNPs or
indels <= 32 nt, the given marker primer is located like the reference's designPrimers does it
(first occurrence of its sequence in the template; po_p3_design_fixed_batch relocates the task),
and BLAST off-targets are fixed to none.
"""
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _native, lib_primer3, primer3_bindings
from .thermoanalysis import get_context

_worker_contexts = {}
_worker_lock = threading.Lock()


def _worker_context(device, slot):
    """One scoring context per (process, device, worker slot): chunks are processed by a few host
    threads so that one chunk's numpy glue overlaps another chunk's kernels (the C calls release
    the GIL)."""
    if slot == 0:
        return get_context(device)
    base = get_context(device)
    key = (os.getpid(), base.device, slot)
    with _worker_lock:
        ctx = _worker_contexts.get(key)
        if ctx is None:
            ctx = _native.Context(base.device)
            _worker_contexts[key] = ctx
        return ctx


CONFIG5_SEED = 20240608
_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 255, dtype=np.uint8)
_CODE[_LETTERS] = np.arange(4, dtype=np.uint8)


def synth_config5(n_markers, seed=CONFIG5_SEED, length=500, gc=0.42):
    """BASELINE config 5 marker set: 500-nt templates (GC 0.42), SNP at index 249,
    homolog maps mism ~ Bernoulli(0.05), partial_mism a superset at 0.15, all = 1, mutations
    Bernoulli(0.01) per position with aaf ~ U(0, 0.5), indel size 0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    codes = rng.choice(4, size=(n_markers, length), p=p).astype(np.uint8)
    idx = (length - 1) // 2
    shift = rng.integers(1, 4, n_markers).astype(np.uint8)
    alt = (codes[:, idx] + shift) % 4                         # uniform over the three other bases
    mism = rng.random((n_markers, length)) < 0.05
    partial = mism | (rng.random((n_markers, length)) < 0.10 / 0.95)
    maps = {"mism": mism, "partial_mism": partial, "all": np.ones((n_markers, length), dtype=bool)}
    is_mut = rng.random((n_markers, length)) < 0.01
    is_mut[:, idx] = False
    m_idx, pos = np.nonzero(is_mut)
    starts = 100000 + 1000 * np.arange(n_markers, dtype=np.int64)
    mut_seg = np.zeros(n_markers + 1, dtype=np.int64)
    np.cumsum(np.bincount(m_idx, minlength=n_markers), out=mut_seg[1:])
    return dict(templates=_LETTERS[codes], starts=starts, marker_idx=np.full(n_markers, idx, dtype=np.int64),
                ref_len=np.ones(n_markers, dtype=np.int32), alt=[chr(_LETTERS[a]) for a in alt], maps=maps,
                mut_pos=starts[m_idx] + pos, mut_aaf=rng.random(len(pos)) * 0.5,
                mut_indel=np.zeros(len(pos), dtype=np.int32), mut_seg=mut_seg)


def _pack_rows(ascii_rows, n_words):
    """(n, L) ASCII -> (n, n_words) uint32 2-bit words (L <= 16 * n_words, ACGT only)."""
    n, width = ascii_rows.shape
    codes = _CODE[ascii_rows]
    if (codes > 3).any():
        raise ValueError("templates must be plain ACGT")
    full = np.zeros((n, 16 * n_words), dtype=np.uint32)
    full[:, :width] = codes
    shifts = (2 * (np.arange(16 * n_words) % 16)).astype(np.uint32)
    return np.bitwise_or.reduce((full << shifts[None, :]).reshape(n, n_words, 16), axis=2).astype(np.uint32)


def marker_records(templates, marker_idx, ref_len, alts):
    """Vectorised kasp.marker_record: po_kasp_marker records from (n, L) ASCII templates."""
    n, length = templates.shape
    if length > 512:
        raise ValueError("templates longer than 512 nt")
    rec = np.zeros(n, dtype=_native.KASP_MARKER_DTYPE)
    rec["tmpl"] = _pack_rows(templates, 32)
    alt_len = np.array([len(a) for a in alts], dtype=np.int32)
    if alt_len.max(initial=0) > 32 or int(np.max(ref_len, initial=0)) > 32:
        raise ValueError("allele longer than 32 nt")
    alt_rows = np.full((n, 32), ord("A"), dtype=np.uint8)
    for k, a in enumerate(alts):
        alt_rows[k, :len(a)] = np.frombuffer(a.encode("ascii"), dtype=np.uint8)
    rec["alt"] = _pack_rows(alt_rows, 2)
    rec["tmpl_len"] = length
    rec["marker_pos"] = marker_idx                       # pos1 - hroi.start  (_lib_kasp.py:297)
    rec["marker_pos_rc"] = length - marker_idx           # hroi.stop - pos1 + 1 (_lib_kasp.py:298)
    rec["ref_len"] = ref_len
    rec["alt_len"] = alt_len
    return rec


def _slot_geometry(marker_idx, ref_len, alt_len, length, lo, hi):
    """Template position (primer3 convention) and length of the REF marker primer of every
    enumeration slot: slot 2*k is the F primer of padding lo+k, slot 2*k+1 the R primer
    (reference _lib_kasp.py:309-337)."""
    pad = np.repeat(np.arange(lo, hi + 1), 2)[None, :]
    is_r = (np.arange(2 * (hi - lo + 1)) % 2 == 1)[None, :]
    d = (ref_len - alt_len)[:, None]
    trim = np.maximum(d, 0)                              # REF primer loses its first d bases
    mp, rl = marker_idx[:, None], ref_len[:, None]
    x2 = mp + rl
    x1 = x2 - pad - np.abs(d) + trim
    y2 = (length - marker_idx)[:, None]
    y1 = y2 - pad - np.abs(d) - rl + trim
    start = np.where(is_r, length - 1 - y1, x1)
    ln = np.where(is_r, y2 - y1, x2 - x1)
    return start.astype(np.int32), ln.astype(np.int32)


def pack_maps(maps):
    """The three inclusion maps (n, L) bool -> (n, 3, 16) uint32 bit maps (mism, partial_mism, all): bit p % 32 of
    word p / 32 set iff a primer may cover base p."""
    out = []
    for name in ("mism", "partial_mism", "all"):
        m = np.asarray(maps[name], dtype=bool)
        n, length = m.shape
        full = np.zeros((n, 512), dtype=np.uint8)
        full[:, :length] = m
        out.append(np.packbits(full, axis=1, bitorder="little").view("<u4").reshape(n, 16))
    return np.ascontiguousarray(np.stack(out, axis=1))


def design_markers_fused(data, n_primers=10, reporters=None, device=None, ctx=None, workers=2, block=4096):
    """The same design as design_markers_arrays through ONE C call per block of markers
    (po_kasp_design_batch): masks, marker-primer tasks, the round-robin merge, coordinates and the
    ranking all run on the device; the host packs the inputs and receives the selected triples.
    Blocks go to ``workers`` host threads, each with its own context and stream, so that the tail of
    one block's picker rounds (few unfinished tasks, host-driven) overlaps another block's kernels.
    Returns ``(records, n_sel, n_merged)``: records[n, n_primers] of _native.KASP_PAIR_DTYPE in report
    order (sharding.compact_selection layout)."""
    g = lib_primer3.PRIMER3_GLOBALS
    settings = primer3_bindings.settings_from_globals()
    templates = np.ascontiguousarray(data["templates"], dtype=np.uint8)
    n = len(templates)
    params = _native.KaspDesignParams()
    params.min_size, params.max_size = int(g["PRIMER_MIN_SIZE"]), int(g["PRIMER_MAX_SIZE"])
    params.top_n = int(n_primers)
    params.with_mutations = 1 if data.get("mut_seg") is not None else 0
    params.tm_delta = float(lib_primer3.TM_DELTA)
    params.valid = lib_primer3.valid_params()
    rep = _native.pack(list(reporters or ("", "")))
    for k in range(8):
        params.reporters[k] = int(rep["w"].reshape(-1)[k])
    marker_idx = np.asarray(data["marker_idx"], dtype=np.int64)
    ref_len = np.asarray(data["ref_len"], dtype=np.int64)
    starts = np.ascontiguousarray(data["starts"], dtype=np.int64)
    mut = None
    if params.with_mutations:
        mut = (np.ascontiguousarray(data["mut_pos"], dtype=np.int64), np.ascontiguousarray(data["mut_aaf"], dtype=np.float64),
               np.ascontiguousarray(data["mut_indel"], dtype=np.int32), np.ascontiguousarray(data["mut_seg"], dtype=np.int64))

    def call(c, lo, hi):
        # the block's inputs are packed here, inside the worker: one block's packing overlaps another's kernels
        rec = marker_records(templates[lo:hi], marker_idx[lo:hi], ref_len[lo:hi], data["alt"][lo:hi])
        maps_bits = pack_maps({k: m[lo:hi] for k, m in data["maps"].items()})
        kw = {}
        if mut is not None:
            a0, a1 = int(mut[3][lo]), int(mut[3][hi])
            kw = dict(mut_pos=mut[0][a0:a1], mut_aaf=mut[1][a0:a1], mut_indel=mut[2][a0:a1], mut_seg=mut[3][lo:hi + 1] - a0)
        return c.kasp_design(settings, params, rec, starts[lo:hi], maps_bits, **kw)

    bounds = [(b0, min(n, b0 + block)) for b0 in range(0, n, block)]
    workers = 1 if ctx is not None else max(1, min(workers, len(bounds)))
    if workers == 1:
        return call(ctx if ctx is not None else get_context(device), 0, n)
    slots = list(range(workers))
    slot_lock = threading.Lock()

    def run(b):
        with slot_lock:
            slot = slots.pop()
        try:
            return call(_worker_context(device, slot), b[0], b[1])
        finally:
            with slot_lock:
                slots.append(slot)

    with ThreadPoolExecutor(max_workers=workers) as pool:
        parts = list(pool.map(run, bounds))
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))


def design_markers_arrays(data, n_primers=10, reporters=None, device=None, chunk=2048, stats=None, workers=2,
                          ctx=None):
    """Runs the pipeline on a marker set laid out like synth_config5()'s result.  Thresholds come
    from lib_primer3.PRIMER3_GLOBALS / TM_DELTA and primer3_bindings' globals (set_globals).
    ``ctx`` replaces the scoring context (tests and the bench's CPU baseline inject the CPU
    oracle's context there; the product never does).
    Returns per-marker arrays: ``sel`` (n, n_primers) indices into the merged pair arrays
    ``f`` / ``r`` / ``a`` (oligo records), ``dir``, ``goodness``, ``qbits``, ``seg``."""
    g = lib_primer3.PRIMER3_GLOBALS
    lo, hi = int(g["PRIMER_MIN_SIZE"]), int(g["PRIMER_MAX_SIZE"])
    settings = primer3_bindings.settings_from_globals()
    depth = settings.num_return
    vparams = lib_primer3.valid_params()
    templates = np.ascontiguousarray(data["templates"], dtype=np.uint8)
    n, length = templates.shape
    if 2 * (hi - lo + 1) > 32:
        raise ValueError("PRIMER_MAX_SIZE - PRIMER_MIN_SIZE + 1 > 16: the merge keys hold 32 marker-primer slots")
    out = {k: [] for k in ("f", "r", "a", "dir", "goodness", "qbits", "sel", "n_sel", "count")}
    types_ = ["mism_mut", "mism", "partial_mism_mut", "partial_mism", "all_mut", "all"] \
        if data.get("mut_seg") is not None else ["mism", "partial_mism", "all"]
    bounds = [(c0, min(n, c0 + chunk)) for c0 in range(0, n, chunk)]
    workers = 1 if ctx is not None else max(1, min(workers, len(bounds)))
    slots = list(range(workers))
    slot_lock = threading.Lock()

    def run(b):
        with slot_lock:
            slot = slots.pop()
        try:
            st = {} if stats is not None else None
            res = _design_chunk(ctx if ctx is not None else _worker_context(device, slot), data, b[0], b[1],
                                templates[b[0]:b[1]], length, lo, hi,
                                settings, depth, vparams, types_, n_primers, reporters, st)
            return res, st
        finally:
            with slot_lock:
                slots.append(slot)

    if workers == 1:
        results = [run(b) for b in bounds]
    else:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            results = list(pool.map(run, bounds))
    for res, st in results:
        for k in out:
            out[k].append(res[k])
        if stats is not None:
            for k, v in st.items():
                stats[k] = stats.get(k, 0) + v
    seg = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.concatenate(out["count"]), out=seg[1:])
    off = 0
    sel = []
    for cnt, s_ in zip(out["count"], out["sel"]):          # chunk-local pair indices -> global
        sel.append(np.where(s_ >= 0, s_ + off, -1))
        off += int(cnt.sum())
    return dict(f=np.concatenate(out["f"]), r=np.concatenate(out["r"]), a=np.concatenate(out["a"]),
                dir=np.concatenate(out["dir"]), goodness=np.concatenate(out["goodness"]),
                qbits=np.concatenate(out["qbits"]), sel=np.concatenate(sel), n_sel=np.concatenate(out["n_sel"]),
                seg=seg)


def _design_chunk(ctx, data, c0, c1, templates, length, lo, hi, settings, depth, vparams, types_, n_primers,
                  reporters, stats):
    m = c1 - c0
    slots = 2 * (hi - lo + 1)
    marker_idx = np.asarray(data["marker_idx"][c0:c1], dtype=np.int64)
    ref_len = np.asarray(data["ref_len"][c0:c1], dtype=np.int64)
    alts = data["alt"][c0:c1]
    alt_len = np.array([len(a) for a in alts], dtype=np.int64)
    starts = np.asarray(data["starts"][c0:c1], dtype=np.int64)
    # ---- a3 + a1/a2 ----------------------------------------------------------------------------
    ref, alt = ctx.kasp_enumerate(marker_records(templates, marker_idx, ref_len, alts), lo, hi)
    pr = ctx.oligo_props_batch(ref.reshape(-1), vparams).reshape(ref.shape)
    pa = ctx.oligo_props_batch(alt.reshape(-1), vparams).reshape(alt.shape)
    keep = ((pr["flags"] & _native.PRIMER_VALID) != 0) & ((pa["flags"] & _native.PRIMER_VALID) != 0)
    p_start, p_len = _slot_geometry(marker_idx, ref_len, alt_len, length, lo, hi)
    mk, sl = np.nonzero(keep)                             # marker-major, slot order = mpp order
    # ---- search masks (a4): inclusion maps, mutation twins, the marker-window quirk --------------
    size = int(lib_primer3.PRIMER3_GLOBALS["PRIMER_MAX_SIZE"])
    # _lib_kasp.py:681-684, literally: mmap[min(1, start - size):max(len, start + size)] = True.  `start` is a
    # GENOME coordinate: the slice is [1:] for every region further than `size` from the chromosome start, [0:]
    # at start == size, and for a NEGATIVE first index numpy counts from the end ([len + s:], clamped at 0).
    s_idx = np.minimum(1, starts - size)
    win_lo = np.where(s_idx >= 0, s_idx, np.maximum(length + s_idx, 0))
    in_window = np.arange(length)[None, :] >= win_lo[:, None]
    masks = {}
    mut_rows = None
    if data.get("mut_seg") is not None:
        ms = np.asarray(data["mut_seg"], dtype=np.int64)
        a0, a1 = ms[c0], ms[c1]
        mut_marker = np.repeat(np.arange(m), np.diff(ms[c0:c1 + 1]))
        mut_rows = (mut_marker, np.asarray(data["mut_pos"][a0:a1], dtype=np.int64) - starts[mut_marker])
    for name in types_:
        base = name[:-4] if name.endswith("_mut") else name
        mp = np.array(data["maps"][base][c0:c1], dtype=bool, copy=True)
        if name.endswith("_mut"):
            mp[mut_rows] = False
        mp |= in_window
        masks[name] = (~mp).astype(np.uint8)
    blob = templates.reshape(-1)
    off = np.arange(m + 1, dtype=np.int64) * length
    # ---- a5 + a6, one pass per search type --------------------------------------------------------
    count = np.zeros(m, dtype=np.int64)
    acc = {k: [] for k in ("marker", "slot", "other")}
    seen_keys = np.zeros(0, dtype=np.int64)
    done_masks = []
    # A primer that can pair with a marker primer lies inside the largest product that starts / ends
    # at that marker primer; mask differences outside that window cannot change the picker's answer.
    maxprod = int(max(settings.size_hi[:settings.n_size_ranges]))
    slot_is_r = (np.arange(slots) % 2 == 1)[None, :]
    big = 1 << 30
    lo_pos = np.where(keep & slot_is_r, p_start - maxprod + 1, big).min(axis=1)      # left candidates of R tasks
    hi_pos = np.where(keep & ~slot_is_r, p_start + maxprod - 1, -big).max(axis=1)    # right candidates of F tasks
    # the marker primers themselves must stay clear of excluded bases as well
    own_lo = np.where(keep, np.where(slot_is_r, p_start - p_len + 1, p_start), big).min(axis=1)
    own_hi = np.where(keep, np.where(slot_is_r, p_start, p_start + p_len - 1), -big).max(axis=1)
    pos = np.arange(length)[None, :]
    window = (pos >= np.minimum(lo_pos, own_lo)[:, None]) & (pos <= np.maximum(hi_pos, own_hi)[:, None])
    for step, name in enumerate(types_):
        mask = masks[name]
        active = count < depth
        for prev in done_masks:                           # same mask, same answer: nothing new to merge
            active &= ((mask != prev) & window).any(axis=1)
        done_masks.append(mask)
        tsel = active[mk]
        if not tsel.any():
            continue
        tm_, ts_ = mk[tsel], sl[tsel]
        tasks = np.zeros(len(tm_), dtype=_native.P3_TASK_DTYPE)
        tasks["set"] = tm_
        tasks["side"] = 1 + (ts_ % 2)
        tasks["start"] = p_start[tm_, ts_]
        tasks["len"] = p_len[tm_, ts_]
        # picker + Primer.is_valid of every picked common primer, in one device call
        _, pairs, n_pairs = ctx.p3_design_fixed(settings, (blob, off), tasks, mask, valid=vparams)
        if stats is not None:
            stats["p3_tasks"] = stats.get("p3_tasks", 0) + len(tasks)
            stats["p3_pairs"] = stats.get("p3_pairs", 0) + int(n_pairs.sum())
        oth = pairs["other"]
        cand2d = (np.arange(depth)[None, :] < n_pairs[:, None]) & ((oth["flags"] & _native.P3_PRIMER_VALID) != 0)
        t_idx, r_idx = np.nonzero(cand2d)
        if len(t_idx) == 0:
            continue
        e_marker, e_slot = tm_[t_idx], ts_[t_idx]
        if len(seen_keys) or step + 1 < len(types_):
            # same marker primer slot and same common primer: already in the repository?
            pkey = (e_marker * 32 + e_slot) * 1024 * 64 + oth["start"][t_idx, r_idx].astype(np.int64) * 64 + \
                oth["len"][t_idx, r_idx]
            if len(seen_keys):
                fresh = ~np.isin(pkey, seen_keys)
                t_idx, r_idx, e_marker, e_slot, pkey = (x[fresh] for x in (t_idx, r_idx, e_marker, e_slot, pkey))
        else:
            pkey = None
        # round-robin pop order: rank r first, then marker-primer order; cut at `depth` per marker
        order = np.argsort((e_marker * depth + r_idx) * 32 + e_slot, kind="stable")
        cm = e_marker[order]
        rank = np.arange(len(order)) - np.searchsorted(cm, cm, side="left")
        ci = order[rank < (depth - count[cm])]
        # (pairs the reference pops but rejects are never added later either: identical inputs
        # give identical answers in every later pass)
        acc["marker"].append(e_marker[ci])
        acc["slot"].append(e_slot[ci])
        acc["other"].append(oth[t_idx[ci], r_idx[ci]])
        if pkey is not None:
            seen_keys = np.concatenate([seen_keys, pkey[ci]])
        count += np.bincount(e_marker[ci], minlength=m)
    if acc["marker"]:
        e_marker = np.concatenate(acc["marker"])
        e_slot = np.concatenate(acc["slot"])
        other = np.concatenate(acc["other"])
        order = np.argsort(e_marker, kind="stable")       # pass order, then merge order, inside a marker
        e_marker, e_slot, other = e_marker[order], e_slot[order], other[order]
    else:
        e_marker = e_slot = np.zeros(0, dtype=np.int64)
        other = np.zeros(0, dtype=_native.P3_OLIGO_DTYPE)
    npair = len(e_marker)
    seg = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(e_marker, minlength=m), out=seg[1:])
    is_r = (e_slot % 2 == 1)
    mark = ref[e_marker, e_slot]
    f = np.where(is_r, other["seq"]["w"].T, mark["w"].T).T
    r = np.where(is_r, mark["w"].T, other["seq"]["w"].T).T
    fo = np.zeros(npair, dtype=_native.OLIGO_DTYPE)
    ro = np.zeros(npair, dtype=_native.OLIGO_DTYPE)
    fo["w"], ro["w"] = f, r
    ao = np.ascontiguousarray(alt[e_marker, e_slot])
    res = dict(f=fo, r=ro, a=ao, dir=is_r.astype(np.uint8), count=np.diff(seg))
    if npair == 0:
        res.update(goodness=np.zeros(0, np.int32), qbits=np.zeros(0, np.int32),
                   sel=np.full((m, n_primers), -1, dtype=np.int64), n_sel=np.zeros(m, dtype=np.int32))
        return res
    # ---- genome coordinates of F / R / A (get_primers :505-516, merge_primers :559-562) ------------
    g0 = starts[e_marker]
    ms_, ml_ = p_start[e_marker, e_slot].astype(np.int64), p_len[e_marker, e_slot].astype(np.int64)
    os_, ol_ = other["start"].astype(np.int64), other["len"].astype(np.int64)
    left_s = np.where(is_r, os_, ms_)
    left_l = np.where(is_r, ol_, ml_)
    right_s = np.where(is_r, ms_, os_)
    right_l = np.where(is_r, ml_, ol_)
    f_start, f_stop = g0 + left_s, g0 + left_s + left_l - 1
    r_stop = g0 + right_s
    r_start = r_stop - (right_l - 1)
    a_start, a_stop = np.where(is_r, r_start, f_start), np.where(is_r, r_stop, f_stop)
    pstart = np.stack([f_start, r_start, a_start], axis=1)
    pstop = np.stack([f_stop, r_stop, a_stop], axis=1)
    if data.get("mut_seg") is not None:
        ms = np.asarray(data["mut_seg"], dtype=np.int64)
        a0, a1 = ms[c0], ms[c1]
        aaf, indel = ctx.annotate_mutations(data["mut_pos"][a0:a1], data["mut_aaf"][a0:a1],
                                            data["mut_indel"][a0:a1], ms[c0:c1 + 1] - a0, seg, pstart, pstop, 3)
    else:
        aaf, indel = np.zeros((npair, 3)), np.zeros(npair, dtype=np.int32)
    rk = ctx.rank_pairs(_native.ASSAY_KASP, fo, ro, seg, np.zeros(npair, dtype=np.int32), aaf, indel,
                        lib_primer3.TM_DELTA, n_primers, a=ao, direction=res["dir"], reporters=reporters or ("", ""))
    res.update(goodness=rk["goodness"], qbits=rk["qbits"], sel=rk["sel"], n_sel=rk["n_sel"])
    return res
