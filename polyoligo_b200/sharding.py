"""Multi-GPU sharding of the scoring path.

Every oligo, pair and marker is independent (the reference runs one process per
marker and exchanges nothing but result files, reference src/polyoligo/
cli_kasp.py:354, _lib_kasp.py:533-549), so the units are split into contiguous,
cost-balanced ranges, one per GPU; each rank scores its range with its own
context and the per-unit results are gathered on the host in input order.
There is no data-path collective and no NCCL traffic.

Two drivers:
  * ``score_pairs_multi_gpu``  one process, one host thread + context per GPU
  * ``shard_bounds`` / ``gather_to_rank0``  one process per GPU under torchrun
    (torch.distributed is used for the host-side gather only; gloo or nccl)
"""
import threading

import numpy as np


def partition(costs, n_shards):
    """Contiguous ranges [(start, stop), ...] whose summed costs are as equal as the
    order allows: range k ends where the cost prefix first reaches (k+1)/n of the total."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    prefix = np.concatenate([[0.0], np.cumsum(costs)])
    total = prefix[-1]
    cuts = [0]
    for k in range(1, n_shards):
        cuts.append(int(np.searchsorted(prefix, total * k / n_shards, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(n_shards)]


def shard_bounds(n, rank, world, costs=None):
    """The contiguous unit range of ``rank`` among ``world`` ranks."""
    if costs is None:
        costs = np.ones(n)
    return partition(costs, world)[rank]


def pair_costs(a, b):
    """Cost model of one heterodimer evaluation: ~ (len1 * len2)^2 DP relaxations."""
    la = ((a["w"][:, 3] >> 24) & 0x3F).astype(np.float64)
    lb = ((b["w"][:, 3] >> 24) & 0x3F).astype(np.float64)
    return (la * lb) ** 2 + 1.0


def score_pairs_multi_gpu(a, b, devices):
    """Config-4 style scoring of (a[k], b[k]) pairs over several GPUs of one box from ONE
    process: one thread and one context per device, results concatenated in input order."""
    from . import _native
    bounds = partition(pair_costs(a, b), len(devices))
    parts = [None] * len(devices)
    errors = []

    def work(k, dev):
        try:
            ctx = _native.Context(dev)
            s, e = bounds[k]
            parts[k] = ctx.score_pairs(a[s:e], b[s:e])
            ctx.close()
        except Exception as exc:  # noqa: BLE001 - re-raised below
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k, d)) for k, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return {key: np.concatenate([p[key] for p in parts]) for key in parts[0]}


def gather_to_rank0(local, dist):
    """Host-side gather of per-rank result arrays (dict of numpy arrays) in rank order.
    ``dist`` is torch.distributed with an initialised process group (gloo or nccl)."""
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(local, bucket, dst=0)
    if dist.get_rank() != 0:
        return None
    return {key: np.concatenate([b[key] for b in bucket]) for key in bucket[0]}


# ---- marker sets (BASELINE config 5) -----------------------------------------------------------
def slice_markers(data, start, stop):
    """The markers [start, stop) of a kasp_pipeline marker set (same dict layout)."""
    out = {}
    for k in ("templates", "starts", "marker_idx", "ref_len"):
        out[k] = data[k][start:stop]
    out["alt"] = list(data["alt"][start:stop])
    out["maps"] = {name: m[start:stop] for name, m in data["maps"].items()}
    if data.get("mut_seg") is not None:
        seg = np.asarray(data["mut_seg"], dtype=np.int64)
        a0, a1 = int(seg[start]), int(seg[stop])
        out["mut_seg"] = seg[start:stop + 1] - a0
        for k in ("mut_pos", "mut_aaf", "mut_indel"):
            out[k] = data[k][a0:a1]
    else:
        out["mut_seg"] = None
    return out


def merge_marker_results(parts):
    """Concatenates per-shard results of kasp_pipeline.design_markers_arrays in shard order:
    pair arrays are appended, per-marker segment offsets and selected-pair indices are shifted."""
    out = {k: np.concatenate([p[k] for p in parts]) for k in ("f", "r", "a", "dir", "goodness", "qbits", "n_sel")}
    seg, sel, off = [np.zeros(1, dtype=np.int64)], [], 0
    for p in parts:
        seg.append(p["seg"][1:] + off)
        sel.append(np.where(p["sel"] >= 0, p["sel"] + off, -1))
        off += int(p["seg"][-1])
    out["seg"] = np.concatenate(seg)
    out["sel"] = np.concatenate(sel)
    return out


# per-marker result record of a design: the selected (top-n) primer pairs, 64 bytes each
SEL_DTYPE = np.dtype([("f", "<u4", (4,)), ("r", "<u4", (4,)), ("a", "<u4", (4,)), ("goodness", "<i4"),
                      ("qbits", "<i4"), ("dir", "u1"), ("pad", "u1", (7,))])
assert SEL_DTYPE.itemsize == 64


def compact_selection(res, n_primers):
    """The per-marker results worth sending home: ``(records[n_markers, n_primers], n_sel)``, the
    selected pairs of every marker in rank order (rows beyond n_sel are zero)."""
    sel = np.asarray(res["sel"])
    n = sel.shape[0]
    out = np.zeros((n, n_primers), dtype=SEL_DTYPE)
    ok = sel >= 0
    idx = sel[ok]
    out["f"][ok] = res["f"]["w"][idx]
    out["r"][ok] = res["r"]["w"][idx]
    out["a"][ok] = res["a"]["w"][idx]
    out["goodness"][ok] = res["goodness"][idx]
    out["qbits"][ok] = res["qbits"][idx]
    out["dir"][ok] = res["dir"][idx]
    return out, np.asarray(res["n_sel"], dtype=np.int32)


_gather_buffers = {}


def _staging(key, shape, pinned, device=None):
    """Staging tensors of the gather, allocated once per process and shape (pinned host memory and
    device buffers are expensive to allocate and would otherwise dominate a sub-second shard)."""
    import torch
    t = _gather_buffers.get(key + (shape,))
    if t is None:
        if device is not None:
            t = torch.empty(shape, dtype=torch.uint8, device=device)
        else:
            t = torch.empty(shape, dtype=torch.uint8)
            if pinned:
                t = t.pin_memory()
        _gather_buffers[key + (shape,)] = t
    return t


def gather_selection(records, n_sel, bounds, dist, device=None):
    """Gather of the per-marker results of all ranks on rank 0, in marker order, as TENSORS (no
    pickling): every rank contributes one byte tensor padded to the largest shard; with the nccl
    backend the bytes travel device to device (``device``: this rank's cuda device) through
    staging buffers that are kept between calls, with gloo they stay on the host.  ``bounds``:
    [(start, stop), ...] of every rank's shard.
    Returns (records[n_total, n_primers], n_sel[n_total]) on rank 0, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    n_primers = records.shape[1]
    rows = max(e - s for s, e in bounds)
    row_bytes = n_primers * SEL_DTYPE.itemsize + 4
    on_gpu = dist.get_backend() == "nccl"
    m = len(n_sel)
    host_in = _staging(("in",), (rows, row_bytes), on_gpu)
    buf = host_in.numpy()
    buf[:m, :row_bytes - 4] = records.view(np.uint8).reshape(m, -1)
    buf[:m, row_bytes - 4:] = n_sel.astype("<i4").view(np.uint8).reshape(m, 4)
    buf[m:] = 0
    if on_gpu:
        dev = torch.device("cuda", device if device is not None else torch.cuda.current_device())
        t = _staging(("dev_in",), (rows, row_bytes), False, dev)
        t.copy_(host_in, non_blocking=True)
        out = _staging(("dev_out",), (world, rows, row_bytes), False, dev) if rank == 0 else None
    else:
        t = host_in
        out = _staging(("out",), (world, rows, row_bytes), False) if rank == 0 else None
    dist.gather(t, list(out.unbind(0)) if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    if on_gpu:
        host_out = _staging(("host_out",), (world, rows, row_bytes), True)
        host_out.copy_(out)
        parts = host_out.numpy()
    else:
        parts = out.numpy()
    n_total = bounds[-1][1] - bounds[0][0]
    recs = np.empty((n_total, n_primers), dtype=SEL_DTYPE)
    cnts = np.empty(n_total, dtype=np.int32)
    for (s, e), p in zip(bounds, parts):
        k = e - s
        recs[s - bounds[0][0]:e - bounds[0][0]].view(np.uint8).reshape(k, -1)[:] = p[:k, :row_bytes - 4]
        cnts[s - bounds[0][0]:e - bounds[0][0]].view(np.uint8).reshape(k, 4)[:] = p[:k, row_bytes - 4:]
    return recs, cnts


def design_markers_sharded(data, rank, world, dist=None, n_primers=10, device=None, timing=None, fused=None, **kw):
    """One rank's share of a marker set under torchrun -- the north_star multi-GPU path: a host-side
    partition of the marker list (contiguous, equally sized ranges: every marker of the synthetic set
    costs about the same) plus a gather of the per-marker results on rank 0.  No data-path collective.
    Returns ``(records, n_sel)`` of ALL markers on rank 0 (``compact_selection`` layout), the local
    share without ``dist``, None on the other ranks.  ``timing`` (dict) receives the shard's own
    design time in seconds.  ``fused`` (default: on, unless a ``ctx`` is injected): one
    po_kasp_design_batch call per shard instead of the array pipeline's call sequence."""
    import time
    from . import kasp_pipeline
    n = len(data["starts"])
    bounds = partition(np.ones(n), world)
    lo, hi = bounds[rank]
    if fused is None:
        fused = kw.get("ctx") is None      # the single-call device path, unless a stand-in context is injected
    t0 = time.perf_counter()
    if fused:
        recs, cnt, _ = kasp_pipeline.design_markers_fused(slice_markers(data, lo, hi), n_primers=n_primers,
                                                          reporters=kw.get("reporters"), device=device)
    else:
        local = kasp_pipeline.design_markers_arrays(slice_markers(data, lo, hi), n_primers=n_primers, device=device,
                                                    **kw)
        recs, cnt = compact_selection(local, n_primers)
    if timing is not None:
        timing["design_s"] = time.perf_counter() - t0
    if dist is None:
        return recs, cnt
    return gather_selection(recs, cnt, bounds, dist, device)
