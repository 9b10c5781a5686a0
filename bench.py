#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 thermodynamic scoring path.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on
that fits one GPU): a synthetic batch of 10^7 random 18-30 nt oligo pairs; one
"step" scores the whole batch: calcHeterodimer(A,B) + calcHairpin(A) +
calcHairpin(B) (3 thal evaluations per pair, the metric's unit) + calcTm(A) +
calcTm(B).  Under torchrun every rank scores its own batch of the same shape
(independent shards, no data-path collective: weak scaling).

Secondary (BASELINE.json configs[4]): the KASP design of 10^5 synthetic markers,
partitioned over the ranks (strong scaling) and gathered on rank 0 as tensors --
north_star's multi-GPU path -- reported under ``config5_kasp_design``.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line on rank 0 (contract in the task description):
  value        whole-job thal evaluations/s with inputs resident in HBM
  e2e          same metric through the public API with pinned HOST buffers
  roofline     dominant kernel (heterodimer DP) against the measured FP64 peak
  cpu_baseline the CPU oracle, one process per host core, on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "oligo_pair_thal_evaluations_per_sec"
UNIT = "thal evaluations/s"
EVALS_PER_PAIR = 3
FLOP_PER_RELAXATION = 24  # SURVEY.md 8d convention
DYES = ("GAAGGTCGGAGTCAACGGATT", "GAAGGTGACCAAGTTCATGCT")  # reference cli_kasp.py:65,72


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU per step")
    p.add_argument("--cpu-pairs", type=int, default=1_000_000,
                   help="config-4 prefix timed on the CPU (BASELINE.md section 3: 10^6, best of 3)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-markers", type=int, default=1024, help="config-5 prefix timed on the CPU")
    p.add_argument("--markers", type=int, default=100_000,
                   help="TOTAL markers of the config-5 (KASP design) measurement, partitioned over the ranks; 0 skips it")
    p.add_argument("--marker-passes", type=int, default=3)
    p.add_argument("--secondary-pairs", type=int, default=2_000_000,
                   help="pairs of the KASP-tailed and homodimer side measurements; 0 skips them")
    p.add_argument("--cpu-leg", default=None, choices=["config4", "config5"], help=argparse.SUPPRESS)
    p.add_argument("--cpu-passes", type=int, default=3, help=argparse.SUPPRESS)
    return p.parse_args()


def workload_config(args, world):
    return {"workload": "config4: synthetic random 18-30 nt oligo pairs, heterodimer + 2x hairpin + 2x Tm per pair",
            "pairs_per_gpu": args.pairs, "global_pairs": args.pairs * world, "evals_per_pair": EVALS_PER_PAIR,
            "conditions": "mv 50 mM, dv 0, dNTP 0.8 mM, DNA 50 nM, 37 C, max_loop 30 (ThermoAnalysis defaults)",
            "parallelism": "independent shards, %d rank(s), no collective" % world,
            "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % (args.pairs * (32 + 136) / 1e6)}


def set_kasp_globals():
    """reference src/polyoligo/data/PRIMER3_KASP.yaml + cli_kasp.py defaults (--depth 100, -n 10, dyes)"""
    from polyoligo_b200 import lib_primer3 as lp, primer3_bindings as p3b
    p3b.resetP3Globals()
    lp.PRIMER3_GLOBALS.clear()
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0,
                   PRIMER_MIN_TM=55.0, PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0,
                   PRIMER_MAX_POLY_X=100, PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0,
                   PRIMER_MAX_NS_ACCEPTED=0, PRIMER_PAIR_MAX_DIFF_TM=6.0,
                   PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]], PRIMER_NUM_RETURN=100)
    lp.set_tm_delta(5.0)


# ----------------------------------------------------------------------------
# CPU side (oracle): one worker PROCESS per host core over a static partition, the way the
# reference fans out (mp.Pool, reference src/polyoligo/cli_kasp.py:354); a pass takes as long as its
# slowest worker.  Runs in its own interpreter (--cpu-leg), which never touches CUDA.
# ----------------------------------------------------------------------------
def _cpu4_worker(job):
    a_str, b_str = job
    from oracle.oracle import Oracle, ANY, HAIRPIN
    orc = Oracle()
    t0 = time.perf_counter()
    orc.thal_batch(a_str, b_str, ANY, nthreads=1)
    orc.thal_batch(a_str, None, HAIRPIN, nthreads=1)
    orc.thal_batch(b_str, None, HAIRPIN, nthreads=1)
    orc.tm_batch(a_str, nthreads=1)
    orc.tm_batch(b_str, nthreads=1)
    return time.perf_counter() - t0


def _cpu5_worker(job):
    seed_n, lo, hi = job
    from oracle.cpu_pipeline import CpuContext
    from polyoligo_b200 import kasp_pipeline as kp, sharding
    set_kasp_globals()
    data = sharding.slice_markers(kp.synth_config5(seed_n, seed=kp.CONFIG5_SEED), lo, hi)
    cctx = CpuContext(nthreads=1)
    t0 = time.perf_counter()
    res = kp.design_markers_arrays(data, n_primers=10, ctx=cctx, reporters=DYES)
    return time.perf_counter() - t0, int(cctx.thal_item_count()), int(len(res["f"]))


def cpu_leg(args):
    """Body of ``bench.py --cpu-leg ...``: prints one JSON object."""
    import multiprocessing as mp
    from oracle.oracle import build
    from polyoligo_b200 import sharding, synth
    build()
    procs = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(procs)
    if args.cpu_leg == "config4":
        a, b = synth.config4(args.cpu_pairs)
        a_str, b_str = synth.to_strings(a), synth.to_strings(b)
        bounds = sharding.partition(sharding.pair_costs(a, b), procs)
        jobs = [(a_str[s:e], b_str[s:e]) for s, e in bounds]
        pool.map(_cpu4_worker, [(j[0][:50], j[1][:50]) for j in jobs], chunksize=1)  # warm-up: load the library
        passes = [max(pool.map(_cpu4_worker, jobs, chunksize=1)) for _ in range(args.cpu_passes)]
        out = {"pass_seconds": passes, "pairs": args.cpu_pairs, "procs": procs}
    else:
        n = args.cpu_markers
        bounds = sharding.partition(np.ones(n), procs)
        res = pool.map(_cpu5_worker, [(n, s, e) for s, e in bounds if e > s], chunksize=1)
        out = {"pass_seconds": [max(r[0] for r in res)], "markers": n, "procs": procs,
               "thal_evaluations": sum(r[1] for r in res), "pairs_ranked": sum(r[2] for r in res)}
    pool.close()
    print(json.dumps(out), flush=True)


def run_cpu_leg(which, args, passes=3, pairs=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--cpu-leg", which, "--cpu-passes", str(passes),
           "--cpu-pairs", str(pairs if pairs is not None else args.cpu_pairs), "--cpu-markers", str(args.cpu_markers)]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    if out.returncode != 0:
        raise RuntimeError("cpu leg failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_reference(args, rank, world):
    """The reference's CPU path: libprimer3 cannot be installed offline, so this is the oracle port
    of it, one process per host core, on a bounded prefix of the workload per step."""
    if rank != 0:
        return
    # a budget of 3 x 10^6 pairs over the K steps keeps the whole arm within a few minutes
    per_step = max(50_000, min(args.cpu_pairs, 3_000_000 // max(args.steps, 1)))
    r = run_cpu_leg("config4", args, passes=args.steps, pairs=per_step)
    dt = sum(r["pass_seconds"])
    value = EVALS_PER_PAIR * per_step * args.steps / dt
    sample = ("first %d pairs of the config-4 generator per step, %d worker processes (one per host core), static "
              "cost-balanced partition, wall clock of the slowest worker" % (per_step, r["procs"]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["procs"], "kind": "port", "sample": sample,
                             "best_step_value": EVALS_PER_PAIR * per_step / min(r["pass_seconds"])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def current_profile():
    """profiles/current.json: instruction and DRAM figures of the thal kernels from the ncu capture of
    THIS source (stamped with the hash of csrc/ + the public header by scripts/make_current_profile.py).
    Returns None when the sources changed since: stale figures are never reported."""
    from polyoligo_b200 import build as po_build
    try:
        with open(os.path.join(ROOT, "profiles", "current.json")) as f:
            prof = json.load(f)
    except Exception:
        return None
    return prof if prof.get("source_hash") == po_build.source_hash() else None


def run_config5(args, rank, world, local_rank, barrier, dist, reduce_max):
    """BASELINE config 5 as north_star words it: ``--markers`` synthetic KASP markers in total,
    partitioned over the ranks by sharding.partition (strong scaling), every shard designed through
    the public array API with HOST buffers (enumeration, validity, common-primer picking, merge,
    reporter-tailed heterodimers, goodness, top 10), the per-marker results gathered on rank 0 as
    tensors; the gather is inside the timed region.  Wall clock per pass, max over ranks."""
    from polyoligo_b200 import kasp_pipeline as kp, sharding
    set_kasp_globals()
    workers = 2      # design_markers_fused's default: blocks of 4096 markers on two host threads / contexts / streams
    kctxs = [kp._worker_context(local_rank, slot) for slot in range(workers)]
    kp.design_markers_fused(kp.synth_config5(min(4096, max(256, args.markers // world)), seed=1), n_primers=10,
                            reporters=DYES, device=local_rank)  # warm-up (allocations)
    data = kp.synth_config5(args.markers, seed=kp.CONFIG5_SEED)      # the same set on every rank; each takes its range
    bounds = sharding.partition(np.ones(args.markers), world)
    pass_ms, shard_s = [], []
    items = launches = 0
    merged = None
    for _ in range(max(1, args.marker_passes)):
        timing = {}
        items0 = sum(c.thal_item_count() for c in kctxs)
        launches0 = sum(c.launch_count() for c in kctxs)
        barrier()
        t0 = time.perf_counter()
        merged = sharding.design_markers_sharded(data, rank, world, dist=dist, n_primers=10, device=local_rank,
                                                 timing=timing, reporters=DYES)
        barrier()
        pass_ms.append(reduce_max((time.perf_counter() - t0) * 1e3))
        shard_s.append(timing["design_s"])
        items = sum(c.thal_item_count() for c in kctxs) - items0
        launches = sum(c.launch_count() for c in kctxs) - launches0
    if dist is not None:
        import torch
        t = torch.tensor([float(np.median(shard_s)), float(items), float(launches)], dtype=torch.float64, device="cuda")
        allv = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per_rank = [[float(x) for x in v.cpu()] for v in allv]
    else:
        per_rank = [[float(np.median(shard_s)), float(items), float(launches)]]
    if rank != 0:
        return None
    recs, n_sel = merged
    best = min(pass_ms)
    shard = [p[0] for p in per_rank]
    return {"workload": "synthetic KASP design, %d markers in total x 500 nt, SNP at 249, depth 100, top 10, no "
                        "off-targets (BASELINE config 5 shape), contiguous equal shards over %d rank(s)" %
                        (args.markers, world),
            "scaling": "strong",
            "timing": "host wall clock of one pass: sharding.design_markers_sharded = input packing + one "
                      "po_kasp_design_batch call per block of 4096 markers, blocks on two host threads / contexts / "
                      "streams (host buffers in, top-10 triples out) + gather of the per-marker results on rank 0 "
                      "(dist.gather of byte tensors); max over ranks; best of %d passes" % len(pass_ms),
            "markers_per_s": args.markers / (best * 1e-3), "ms": best, "pass_ms": pass_ms,
            "thal_evaluations_per_s": sum(p[1] for p in per_rank) / (best * 1e-3),
            "thal_evaluations": int(sum(p[1] for p in per_rank)), "gpu_launches": int(sum(p[2] for p in per_rank)),
            "shard_design_s": {"max": max(shard), "mean": float(np.mean(shard))},
            "gather_bytes_per_rank": int(max(e - s for s, e in bounds) * (10 * 64 + 4)),
            "host_threads_per_rank": workers, "markers_with_pairs": int((n_sel > 0).sum()),
            "pairs_selected": int(n_sel.sum())}


def run_b200(args, rank, world, local_rank):
    import torch
    import polyoligo_b200 as po
    from polyoligo_b200 import synth

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = po.Context(local_rank)   # raises if the CUDA library is missing: no CPU fallback
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    n = args.pairs

    # ---- workload: pinned host records and device-resident copies ------------
    a_np, b_np = synth.config4(n, seed=synth.CONFIG4_SEED + rank)
    h_a = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
    h_a.numpy()[:] = a_np.view(np.uint8)
    h_b.numpy()[:] = b_np.view(np.uint8)
    d_a = h_a.cuda(non_blocking=True)
    d_b = h_b.cuda(non_blocking=True)
    d_het = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    d_hpa = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    d_hpb = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    d_tma = torch.empty(n, dtype=torch.float64, device="cuda")
    d_tmb = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    # a dedicated non-default stream: its handle is passed to the C ABI, and the CUDA
    # events below are recorded on the same stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def step_dev(events=None):
        """One step on resident inputs; optional per-kernel CUDA events."""
        def mark(k):
            if events is not None:
                events[k].record(stream)
        mark(0)
        ctx.thal_batch_dev(po.THAL_ANY, d_a.data_ptr(), d_b.data_ptr(), n, d_het.data_ptr(), sp)
        mark(1)
        ctx.thal_batch_dev(po.THAL_HAIRPIN, d_a.data_ptr(), None, n, d_hpa.data_ptr(), sp)
        ctx.thal_batch_dev(po.THAL_HAIRPIN, d_b.data_ptr(), None, n, d_hpb.data_ptr(), sp)
        mark(2)
        ctx.tm_batch_dev(d_a.data_ptr(), n, d_tma.data_ptr(), None, sp)
        ctx.tm_batch_dev(d_b.data_ptr(), n, d_tmb.data_ptr(), None, sp)
        mark(3)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if dist is None:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- untimed: algorithmic work of one step and the compute peak ------------
    relax_het = ctx.count_relaxations(po.THAL_ANY, a_np, b_np)
    relax_hp = ctx.count_relaxations(po.THAL_HAIRPIN, a_np) + ctx.count_relaxations(po.THAL_HAIRPIN, b_np)
    fp64_peak = ctx.measure_fp64_peak()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()

    # ---- timed region 1: device-resident -----------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for k in range(args.steps):
        step_dev(ev[k])
    t_end.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    dev_ms = t_begin.elapsed_time(t_end)
    het_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    hp_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    tm_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))

    # ---- timed region 2: end to end through the public host-buffer API -------------
    h_out = {"het": torch.empty(n * 40, dtype=torch.uint8, pin_memory=True),
             "hairpin_a": torch.empty(n * 40, dtype=torch.uint8, pin_memory=True),
             "hairpin_b": torch.empty(n * 40, dtype=torch.uint8, pin_memory=True),
             "tm_a": torch.empty(n, dtype=torch.float64, pin_memory=True),
             "tm_b": torch.empty(n, dtype=torch.float64, pin_memory=True)}
    out_np = {k: (v.numpy().view(po.THAL_OUT_DTYPE) if v.dtype == torch.uint8 else v.numpy()) for k, v in h_out.items()}
    a_pin = h_a.numpy().view(po.OLIGO_DTYPE)
    b_pin = h_b.numpy().view(po.OLIGO_DTYPE)
    ctx.score_pairs(a_pin, b_pin, out=out_np)  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        ctx.score_pairs(a_pin, b_pin, out=out_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    # results of the two paths must agree (same kernels): cheap integrity check
    het_dev = d_het.cpu().numpy().view(po.THAL_OUT_DTYPE)
    assert het_dev.tobytes() == out_np["het"].tobytes(), "device-resident and host-buffer paths disagree"

    # ---- side shapes: reporter-tailed KASP heterodimers and homodimers, device-resident ----
    side = {}
    if args.secondary_pairs > 0:
        m = min(args.secondary_pairs, n)
        rng = np.random.Generator(np.random.PCG64(20240609 + rank))
        core, _ = synth.random_oligos(rng, min(m, 200000), 18, 25)
        com, _ = synth.random_oligos(rng, m, 18, 25)
        core_s = synth.to_strings(core)   # tail + core built on the host for a 200k block, tiled on the device
        tailed = po.pack([DYES[k % 2] + s for k, s in enumerate(core_s)])
        reps = (m + len(tailed) - 1) // len(tailed)
        d_t = torch.from_numpy(np.tile(tailed.view(np.uint8).reshape(-1), reps)[:m * 16].copy()).cuda()
        d_c = torch.from_numpy(com.view(np.uint8).reshape(-1).copy()).cuda()
        d_o = torch.empty(m * 40, dtype=torch.uint8, device="cuda")
        for name, pa, pb in (("kasp_tailed_heterodimers_per_s", d_t, d_c), ("homodimers_per_s", d_a, None)):
            for _ in range(2):
                ctx.thal_batch_dev(po.THAL_ANY, pa.data_ptr(), pb.data_ptr() if pb is not None else None, m,
                                   d_o.data_ptr(), sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            ctx.thal_batch_dev(po.THAL_ANY, pa.data_ptr(), pb.data_ptr() if pb is not None else None, m,
                               d_o.data_ptr(), sp)
            e1.record(stream)
            barrier()
            side[name] = m * world / (reduce_max(e0.elapsed_time(e1)) * 1e-3)
        side["side_shapes"] = ("%d items per GPU each: 21-nt VIC/FAM reporter + 18-25 nt against 18-25 nt (reference "
                               "_lib_kasp.py:95-105); calcHomodimer of the config-4 A oligos" % m)
        del d_t, d_c, d_o

    # ---- secondary: BASELINE config 5, KASP design of synthetic markers (markers/s) ------------
    c5 = None
    if args.markers > 0:
        c5 = run_config5(args, rank, world, local_rank, barrier, dist, reduce_max)

    # ---- reduce over ranks: max time ------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, het_ms, hp_ms, tm_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, het_ms, hp_ms, tm_ms = [float(x) for x in times.cpu()]

    if rank == 0:
        evals = EVALS_PER_PAIR * n * world * args.steps
        value = evals / (dev_ms * 1e-3)
        e2e_value = evals / (e2e_ms * 1e-3)
        achieved = relax_het * FLOP_PER_RELAXATION / (het_ms * 1e-3) / 1e12
        achieved_hp = relax_hp * FLOP_PER_RELAXATION / (hp_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = n * (32 + 40)  # two records in, one result out per heterodimer
        prof = current_profile()
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issue_peak = 4 * sm_count * sm_hz / 1e9

        def issue(inst_per_item, items, ms):
            return {"warp_instructions_per_item": inst_per_item,
                    "achieved_ginst_s": inst_per_item * items / (ms * 1e-3) / 1e9, "peak_ginst_s": issue_peak,
                    "frac": inst_per_item * items / (ms * 1e-3) / 1e9 / issue_peak,
                    "source": "instruction count from profiles/current.json (ncu capture of this source, %s), time "
                              "from this run's CUDA events" % prof.get("capture", "?")}
        roofline = {"bound": "fp64", "kernel": "k_thal<256, dimer, first tier> (heterodimer DP)",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                    "peak_source": "measured here: register-resident DFMA microbenchmark (po_measure_fp64_peak)",
                    "relaxations_per_launch": relax_het, "flop_per_relaxation": FLOP_PER_RELAXATION,
                    "kernel_ms": het_ms, "share_of_step": het_ms / (dev_ms / args.steps),
                    "traffic": prof["dimer"]["dram_bytes_per_pair"] * n if prof else None,
                    "traffic_note": "DRAM bytes per launch = ncu dram__bytes_read+write per pair of the captured launch "
                                    "x this launch's pairs (profiles/current.json; null when the sources changed since "
                                    "the capture); algorithmic bytes are 72 B/pair (32 in + 40 out)",
                    "hbm": {"achieved_gbs": hbm_bytes / (het_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "frac": hbm_bytes / (het_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650"},
                    "hairpin": {"kernel": "k_thal<128, hairpin, first tier> fill + finish, two oligos per pair",
                                "achieved": achieved_hp, "peak": fp64_peak, "unit": "TFLOP/s",
                                "frac": achieved_hp / fp64_peak, "relaxations_per_step": relax_hp, "kernel_ms": hp_ms,
                                "share_of_step": hp_ms / (dev_ms / args.steps),
                                # the cell values cross HBM between the fill and the finish kernel
                                "traffic": prof["hairpin"]["dram_bytes_per_oligo"] * 2 * n if prof else None},
                    "note": "compute-bound max-plus style DP on CUDA cores; not an HBM or tensor-core kernel (SURVEY 8d); "
                            "the kernels are bound by instruction issue (issue_slots), the FP64 share of the "
                            "instructions is small by construction"}
        if prof:
            roofline["issue_slots"] = issue(prof["dimer"]["warp_instructions_per_pair"], n, het_ms)
            roofline["hairpin"]["issue_slots"] = issue(prof["hairpin"]["warp_instructions_per_oligo"], 2 * n, hp_ms)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 136 * n,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "kernel_ms": {"heterodimer": het_ms, "hairpin_x2": hp_ms, "tm_x2": tm_ms},
                "heterodimers_per_s": n * world / (het_ms * 1e-3), "hairpins_per_s": 2 * n * world / (hp_ms * 1e-3)}
        line.update(side)
        if c5 is not None:
            line["config5_kasp_design"] = c5
        if world == 1 and not args.no_cpu_baseline:
            from oracle.oracle import Oracle, build
            build()
            threads = os.cpu_count() or 1
            # cross-check a prefix of the GPU batch against the oracle
            orc = Oracle()
            ref = orc.thal_batch(synth.to_strings(a_np[:20000]), synth.to_strings(b_np[:20000]), 1, nthreads=threads)
            got = out_np["het"][:20000]
            parity = bool(np.array_equal(got["tm"], ref["tm"]) and np.array_equal(got["dg"], ref["dg"]))
            r4 = run_cpu_leg("config4", args, passes=3)
            best = min(r4["pass_seconds"])
            line["cpu_baseline"] = {"value": EVALS_PER_PAIR * r4["pairs"] / best, "unit": UNIT, "cores": r4["procs"],
                                    "kind": "port",
                                    "sample": "first %d pairs of the config-4 generator, one worker process per host "
                                              "core over a static cost-balanced partition, wall clock of the slowest "
                                              "worker, best of 3 passes (BASELINE.md section 3)" % r4["pairs"],
                                    "pass_seconds": r4["pass_seconds"], "gpu_matches_oracle_on_sample": parity}
            if c5 is not None and args.cpu_markers > 0:
                # the same pipeline with the CPU oracle's context, one worker process per core over markers
                # (the reference's fan-out, cli_kasp.py:354)
                r5 = run_cpu_leg("config5", args)
                c5["cpu_baseline"] = {"value": r5["markers"] / r5["pass_seconds"][0], "unit": "markers/s",
                                      "cores": r5["procs"], "kind": "port",
                                      "thal_evaluations": r5["thal_evaluations"], "pairs_ranked": r5["pairs_ranked"],
                                      "sample": "first %d markers of the same generator, one worker process per host "
                                                "core (oracle/cpu_pipeline.py, single-threaded), wall clock of the "
                                                "slowest worker" % r5["markers"]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.cpu_leg:
        cpu_leg(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
