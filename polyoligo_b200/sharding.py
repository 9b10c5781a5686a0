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
